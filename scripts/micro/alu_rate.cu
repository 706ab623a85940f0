// Micro-benchmark: per-SM issue rate of the instructions the GELU epilogue is made of (MUFU.EX2 in f32 / f16 / f16x2 /
// bf16x2, HFMA2, FFMA, the f16 <-> f32 / bf16 conversions), 8 or 16 resident warps, independent chains.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 alu_rate.cu -o alu_rate
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x) {
  uint32_t y;
  if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 1) asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 2) {
    unsigned short h = static_cast<unsigned short>(x), o;
    asm volatile("ex2.approx.f16 %0, %1;" : "=h"(o) : "h"(h));
    y = o;
  } else if (OP == 3) asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 4) asm volatile("fma.rn.f16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  else if (OP == 5) asm volatile("fma.rn.f32 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  else if (OP == 6) asm volatile("tanh.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 7) asm volatile("tanh.approx.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 8) asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=r"(y) : "r"(x));
  else if (OP == 9) asm volatile("fma.rn.bf16x2 %0, %1, %1, %1;" : "=r"(y) : "r"(x));
  else if (OP == 10) {  // f16x2 -> two f32 -> bf16x2 (what the epilogue does to store half results as bf16)
    float lo, hi;
    asm volatile("{.reg .b16 a, b; mov.b32 {a, b}, %2; cvt.f32.f16 %0, a; cvt.f32.f16 %1, b;}" : "=f"(lo), "=f"(hi) : "r"(x));
    asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(hi), "f"(lo));
  } else if (OP == 11) {  // two f32 -> f16x2
    asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(y) : "f"(__uint_as_float(x)), "f"(__uint_as_float(x ^ 1u)));
  }
  return y;
}

template <int OP>
__global__ void __launch_bounds__(512, 1) k(int iters, long long* out, uint32_t seed) {
  uint32_t r[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) r[j] = seed + threadIdx.x * 8 + j;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = op<OP>(r[j]);
  }
  const long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) acc ^= r[j];
  if (acc == 0x12345u) out[1] = acc;
  if (threadIdx.x == 0) out[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int threads) {
  long long* d;
  cudaMalloc(&d, 16);
  const int iters = 2000;
  k<OP><<<1, threads>>>(iters, d, 12345u);
  k<OP><<<1, threads>>>(iters, d, 12345u);
  cudaDeviceSynchronize();
  long long h = 0;
  cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
  const double warp_instr = static_cast<double>(iters) * 8 * (threads / 32);
  printf("%-28s %3d threads: %7.2f clk per warp-instruction per SM  (%5.1f lanes/clk/SM)\n", name, threads,
         h / warp_instr, 32.0 * warp_instr / h);
  cudaFree(d);
}

int main() {
  for (int threads : {256, 512}) {
    run<0>("ex2.approx.ftz.f32", threads);
    run<1>("ex2.approx.f16x2", threads);
    run<2>("ex2.approx.f16", threads);
    run<3>("ex2.approx.ftz.bf16x2", threads);
    run<6>("tanh.approx.f16x2", threads);
    run<7>("tanh.approx.f32", threads);
    run<8>("rcp.approx.ftz.f32", threads);
    run<4>("fma.rn.f16x2", threads);
    run<9>("fma.rn.bf16x2", threads);
    run<5>("fma.rn.f32", threads);
    run<10>("f16x2 -> 2 f32 -> bf16x2", threads);
    run<11>("2 f32 -> f16x2", threads);
  }
  return 0;
}

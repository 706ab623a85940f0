// Micro-benchmark: tcgen05.mma issue/execute rate for small-N tiles, dependent vs independent accumulators.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I <csrc> mma_rate.cu -o mma_rate
#include <cstdio>
#include <cuda_runtime.h>
#include "ptx.cuh"
using namespace vitatk;

// converged-warp issue: every lane runs the loop, one elected lane issues (no divergent region around the MMA)
__device__ __forceinline__ void umma_elect(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t"
      ".reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// mode: bit3 = converged-warp issue; bit0 = alternate between 2 accumulators, bit1 = A from TMEM (TS), bit2 = 4 accumulators round-robin
__global__ void __launch_bounds__(128, 1) k(int N, int count, int mode, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_mbar_init(); }
  if (warp == 0) { ptx::tmem_alloc(&slot, 512); ptx::tmem_relinquish(); }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (warp == 2 && (mode & 8) && !(mode & 16)) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t a = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint64_t b = ptx::make_smem_desc_sw128(ptx::smem_u32(smem) + 16384);
    const int nacc = (mode & 1) ? 2 : 1;
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < count; ++i) umma_elect(tmem + (i % nacc) * 64, a + 2 * (i & 3), b + 2 * (i & 3), idesc, 1u);
      long long t1 = clock64();
      if (lane == 0) ptx::umma_commit(&bar);
      __syncwarp();
      ptx::mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      if (lane == 0) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  // mode 16: TWO warps issue concurrently (converged-elect) into different accumulators, one barrier each: does the issue
  // rate double, i.e. is the ~50 clk per tcgen05.mma a per-warp cost or a limit of the tensor-pipe front end?
  if ((warp == 2 || warp == 3) && (mode & 16)) {
    __shared__ uint64_t bar2[2];
    if (lane == 0) { ptx::mbar_init(&bar2[warp - 2], 1); ptx::fence_mbar_init(); }
    __syncwarp();
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t a = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint64_t b = ptx::make_smem_desc_sw128(ptx::smem_u32(smem) + 16384);
    const uint32_t acc = tmem + (warp - 2) * 256;
    for (int rep = 0; rep < 3; ++rep) {
      asm volatile("bar.sync 1, 64;" ::: "memory");
      long long t0 = clock64();
      for (int i = 0; i < count; ++i) umma_elect(acc + (i & 1) * 64, a + 2 * (i & 3), b + 2 * (i & 3), idesc, 1u);
      long long t1 = clock64();
      if (lane == 0) ptx::umma_commit(&bar2[warp - 2]);
      __syncwarp();
      ptx::mbar_wait(&bar2[warp - 2], rep & 1);
      long long t2 = clock64();
      if (lane == 0 && warp == 2) { out[rep * 2] = t1 - t0; out[rep * 2 + 1] = t2 - t0; }
    }
  }
  if (warp == 1 && lane == 0 && !(mode & 8)) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, N);
    const uint64_t a = ptx::make_smem_desc_sw128(ptx::smem_u32(smem));
    const uint64_t b = ptx::make_smem_desc_sw128(ptx::smem_u32(smem) + 16384);
    const int nacc = (mode & 4) ? 4 : ((mode & 1) ? 2 : 1);
    for (int rep = 0; rep < 3; ++rep) {
      long long t0 = clock64();
      for (int i = 0; i < count; ++i) {
        const uint32_t d = tmem + (i % nacc) * 64;
        if (mode & 2) ptx::umma_bf16_ts(d, tmem + 448, b + 2 * (i & 3), idesc, 1u);
        else ptx::umma_bf16(d, a + 2 * (i & 3), b + 2 * (i & 3), idesc, 1u);
      }
      long long t1 = clock64();
      ptx::umma_commit(&bar);
      ptx::mbar_wait(&bar, rep & 1);
      long long t2 = clock64();
      out[rep * 2] = t1 - t0;
      out[rep * 2 + 1] = t2 - t0;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 0) { ptx::tc_fence_after(); ptx::tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
  const int Ns[] = {16, 64, 128, 256};
  for (int mode = 0; mode < 25; ++mode) {
    if (mode == 5 || mode == 7 || (mode >= 10 && mode != 24)) continue;
    for (int N : Ns) {
      if ((mode & 5) && N > 64 && mode < 8) continue;
      if (mode == 24 && N > 64) continue;
      const int count = 64;
      k<<<1, 128, 100 * 1024>>>(N, count, mode, d);
      long long h[6];
      cudaError_t e = cudaMemcpy(h, d, 48, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      printf("mode %d (%s%s) N=%3d: issue %6.1f clk/MMA, issue+complete %6.1f clk/MMA\n", mode,
             (mode & 16) ? "SS-elect x2 warps" : (mode & 8) ? "SS-elect" : (mode & 2) ? "TS" : "SS", (mode & 4) ? ",4acc" : ((mode & 1) ? ",2acc" : ",1acc"), N, h[4] / (double)count,
             h[5] / (double)count);
    }
  }
  return 0;
}

// HBM-bound kernels of the attack step: LayerNorm forward/backward (warp-shuffle reductions, 16-byte
// vector accesses), classifier head + cross-entropy + its backward, and the fused PGD/FGSM pixel kernels
// (sign step + L-inf projection + [0,1] clamp + ImageNet renormalisation + im2col, one pass).
//
// Reference call sites replaced: torch LayerNorm (HF modeling_vit.py:325-326,333,340,455), classifier
// (HF:613,642), F.cross_entropy (whitebox_attacks.py:29), normalisation (whitebox_attacks.py:26),
// the FGSM tail (whitebox_attacks.py:32-38) and the torchattacks PGD loop body.
#include <cuda_fp16.h>

#include "vitatk_internal.h"
#include "mma_sync.cuh"

namespace vitatk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void unpack8(const uint4& q, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// The two residual streams (forward h, backward dh) are stored as IEEE fp16 instead of bf16: same bytes,
// 8x smaller rounding step on the only tensors whose rounding accumulates over all 24 residual adds.  f16 != 0 selects it.
__device__ __forceinline__ void unpack8(const uint4& q, float* f, int f16) {
  if (f16) {
    const __half2* p = reinterpret_cast<const __half2*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 t = __half22float2(p[i]);
      f[2 * i] = t.x;
      f[2 * i + 1] = t.y;
    }
  } else {
    unpack8(q, f);
  }
}
__device__ __forceinline__ uint4 pack8(const float* f, int f16);
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

__device__ __forceinline__ uint4 pack8(const float* f, int f16) {
  if (!f16) return pack8(f);
  uint4 q;
  __half2* p = reinterpret_cast<__half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2half2_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm: one warp per row, CH = cols / 256 sixteen-byte chunks per lane
// ------------------------------------------------------------------------------------------------
template <int CH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y,
                                                     float2* __restrict__ stats, int rows, float eps, int x_f16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int COLS = CH * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * COLS);
  float v[CH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    unpack8(__ldg(xr + lane + 32 * i), v[i], x_f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  }
  const float mean = warp_sum(s) * (1.f / COLS);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / COLS) + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * COLS);
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = (lane + 32 * i) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[i][j] - mean) * rstd * gg[j] + bb[j];
    yr[lane + 32 * i] = pack8(o);
  }
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

template <int CH>
__global__ void __launch_bounds__(256) ln_stats_kernel(const bf16* __restrict__ x, float2* __restrict__ stats, int rows,
                                                       float eps, int x_f16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int COLS = CH * 256;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * COLS);
  float v[CH][8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    unpack8(__ldg(xr + lane + 32 * i), v[i], x_f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[i][j];
  }
  const float mean = warp_sum(s) * (1.f / COLS);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float d = v[i][j] - mean;
      q += d * d;
    }
  const float rstd = rsqrtf(warp_sum(q) * (1.f / COLS) + eps);
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

template <int CH>
__global__ void __launch_bounds__(256) ln_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                     const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                     const bf16* __restrict__ dres, bf16* __restrict__ dx, int rows,
                                                     int streaming, int x_f16, int g_f16) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  constexpr int COLS = CH * 256;
  // the three inputs are read exactly once: with `streaming` they are loaded evict-first (ld.global.cs) so that they do
  // not push the freshly written output -- which the next two kernels read -- out of L2
  auto ld = [&](const uint4* p) { return streaming ? __ldcs(p) : __ldg(p); };
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * COLS);
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * COLS);
  const float2 st = stats[row];
  float xh[CH][8], gd[CH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    const int c = (lane + 32 * i) * 8;
    float xv[8], dv[8];
    unpack8(ld(xr + lane + 32 * i), xv, x_f16);
    unpack8(ld(dyr + lane + 32 * i), dv);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[i][j] = (xv[j] - st.x) * st.y;
      gd[i][j] = dv[j] * gg[j];
      s1 += gd[i][j];
      s2 += gd[i][j] * xh[i][j];
    }
  }
  const float m1 = warp_sum(s1) * (1.f / COLS);
  const float m2 = warp_sum(s2) * (1.f / COLS);
  uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * COLS);
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * COLS) : nullptr;
#pragma unroll
  for (int i = 0; i < CH; ++i) {
    float o[8];
    float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (rr) unpack8(ld(rr + lane + 32 * i), r, g_f16);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = r[j] + st.y * (gd[i][j] - m1 - xh[i][j] * m2);
    dxr[lane + 32 * i] = pack8(o, g_f16);
  }
}

// LayerNorm backward that also produces the NEXT LoRA site's down-projection of its output:  T = dx * LB^T, LB [16 KS, COLS]
// (the adapter's B^T rows as packed for the skinny GEMM it replaces: bt_proj after LN2-backward, the previous layer's
// bt_fc2 after LN1-backward).  The kernel holds every dx row it writes, so the 77 MB re-read and the launch of the skinny
// GEMM go away.  The eight rows of a CTA are parked in shared memory (16-bit, as stored); warp w then owns columns
// [96 w, 96 w + 96) of the reduction and runs it as m16n8k16 MMAs (rows 8..15 of the A operand are zero) with the
// permuted-k trick of train.cu -- lane t owns 8 consecutive columns of each 32-column block: one 16-byte shared load for
// the A fragment pair, one 16-byte global load (L1-resident, the same 24 KB for every CTA) per 8 adapter rows; the eight
// [8, 16 KS] partials are summed through shared memory in a fixed order.
template <int KS>
__global__ void __launch_bounds__(256) ln_bwd_bt_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                        const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                        const bf16* __restrict__ dres, bf16* __restrict__ dx, int rows,
                                                        int streaming, int x_f16, int g_f16, const bf16* __restrict__ LB,
                                                        bf16* __restrict__ T, int ldt) {
  constexpr int CH = 3, COLS = 768, XLD = COLS + 32;  // 1600-byte rows: two consecutive rows cover all 32 banks
  constexpr int NT = 2 * KS;                          // 8-column tiles of T
  __shared__ __align__(16) uint16_t Xs[8 * XLD];
  __shared__ __align__(16) float Ps[8][8][8 * NT];    // [warp][row][column]
  pdl_wait();
  pdl_launch_dependents();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * 8, row = row0 + warp;
  auto ld = [&](const uint4* p) { return streaming ? __ldcs(p) : __ldg(p); };
  if (row < rows) {
    const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * COLS);
    const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * COLS);
    const float2 st = stats[row];
    float xh[CH][8], gd[CH][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      const int c = (lane + 32 * i) * 8;
      float xv[8], dv[8];
      unpack8(ld(xr + lane + 32 * i), xv, x_f16);
      unpack8(ld(dyr + lane + 32 * i), dv);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[i][j] = (xv[j] - st.x) * st.y;
        gd[i][j] = dv[j] * gg[j];
        s1 += gd[i][j];
        s2 += gd[i][j] * xh[i][j];
      }
    }
    const float m1 = warp_sum(s1) * (1.f / COLS);
    const float m2 = warp_sum(s2) * (1.f / COLS);
    uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * COLS);
    const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * COLS) : nullptr;
#pragma unroll
    for (int i = 0; i < CH; ++i) {
      float o[8];
      float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (rr) unpack8(ld(rr + lane + 32 * i), r, g_f16);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = r[j] + st.y * (gd[i][j] - m1 - xh[i][j] * m2);
      const uint4 pk = pack8(o, g_f16);
      dxr[lane + 32 * i] = pk;
      *reinterpret_cast<uint4*>(Xs + warp * XLD + (lane + 32 * i) * 8) = pk;
    }
  } else {
#pragma unroll
    for (int i = 0; i < CH; ++i) *reinterpret_cast<uint4*>(Xs + warp * XLD + (lane + 32 * i) * 8) = make_uint4(0u, 0u, 0u, 0u);
  }
  __syncthreads();
  {
    const int g = lane >> 2, t = lane & 3;
    float acc[NT][4];
#pragma unroll
    for (int n = 0; n < NT; ++n)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;
#pragma unroll
    for (int kb = 0; kb < 3; ++kb) {
      const int k0 = warp * 96 + kb * 32 + 8 * t;
      const uint4 a = *reinterpret_cast<const uint4*>(Xs + g * XLD + k0);
      uint4 b[NT];
#pragma unroll
      for (int n = 0; n < NT; ++n) b[n] = __ldg(reinterpret_cast<const uint4*>(LB + static_cast<size_t>(n * 8 + g) * COLS + k0));
      const uint32_t a_lo[4] = {a.x, 0u, a.y, 0u}, a_hi[4] = {a.z, 0u, a.w, 0u};  // rows 8..15 of the tile do not exist
#pragma unroll
      for (int n = 0; n < NT; ++n) {
        if (g_f16) {
          mma16816_f16(acc[n], a_lo, b[n].x, b[n].y);
          mma16816_f16(acc[n], a_hi, b[n].z, b[n].w);
        } else {
          mma16816(acc[n], a_lo, b[n].x, b[n].y);
          mma16816(acc[n], a_hi, b[n].z, b[n].w);
        }
      }
    }
#pragma unroll
    for (int n = 0; n < NT; ++n) *reinterpret_cast<float2*>(&Ps[warp][g][n * 8 + 2 * t]) = make_float2(acc[n][0], acc[n][1]);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 8 * 8 * NT; i += 256) {
    const int r = i / (8 * NT), c = i % (8 * NT);
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) s += Ps[w][r][c];
    if (row0 + r < rows) T[static_cast<size_t>(row0 + r) * ldt + c] = __float2bfloat16(s);
  }
}

int layernorm_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float2* stats, int rows, int cols,
                  float eps, cudaStream_t stream, int x_f16) {
  const int grid = (rows + 7) / 8;
  switch (cols) {
    case 768: VITATK_CUDA_OK(launch_pdl(ln_fwd_kernel<3>, dim3(grid), dim3(256), 0, stream, 1, x, gamma, beta, y, stats, rows, eps, x_f16)); break;
    case 1024: VITATK_CUDA_OK(launch_pdl(ln_fwd_kernel<4>, dim3(grid), dim3(256), 0, stream, 1, x, gamma, beta, y, stats, rows, eps, x_f16)); break;
    case 512: VITATK_CUDA_OK(launch_pdl(ln_fwd_kernel<2>, dim3(grid), dim3(256), 0, stream, 1, x, gamma, beta, y, stats, rows, eps, x_f16)); break;
    case 256: VITATK_CUDA_OK(launch_pdl(ln_fwd_kernel<1>, dim3(grid), dim3(256), 0, stream, 1, x, gamma, beta, y, stats, rows, eps, x_f16)); break;
    default: set_error("layernorm_fwd: cols=%d unsupported", cols); return 1;
  }
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_stats(const bf16* x, float2* stats, int rows, int cols, float eps, cudaStream_t stream, int x_f16) {
  const int grid = (rows + 7) / 8;
  switch (cols) {
    case 768: VITATK_CUDA_OK(launch_pdl(ln_stats_kernel<3>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, eps, x_f16)); break;
    case 1024: VITATK_CUDA_OK(launch_pdl(ln_stats_kernel<4>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, eps, x_f16)); break;
    case 512: VITATK_CUDA_OK(launch_pdl(ln_stats_kernel<2>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, eps, x_f16)); break;
    case 256: VITATK_CUDA_OK(launch_pdl(ln_stats_kernel<1>, dim3(grid), dim3(256), 0, stream, 1, x, stats, rows, eps, x_f16)); break;
    default: set_error("layernorm_stats: cols=%d unsupported", cols); return 1;
  }
  return 0;
}

int layernorm_bwd(const bf16* dy, const bf16* x, const float2* stats, const float* gamma, const bf16* dres,
                  bf16* dx_out, int rows, int cols, cudaStream_t stream, int x_f16, int g_f16) {
  const int grid = (rows + 7) / 8;
  static int streaming = -1;
  if (streaming < 0) {
    const char* e = getenv("VITATK_LN_STREAM");
    streaming = (e && e[0] == '0') ? 0 : 1;
  }
  switch (cols) {
    case 768: VITATK_CUDA_OK(launch_pdl(ln_bwd_kernel<3>, dim3(grid), dim3(256), 0, stream, 1, dy, x, stats, gamma, dres, dx_out, rows, streaming, x_f16, g_f16)); break;
    case 1024: VITATK_CUDA_OK(launch_pdl(ln_bwd_kernel<4>, dim3(grid), dim3(256), 0, stream, 1, dy, x, stats, gamma, dres, dx_out, rows, streaming, x_f16, g_f16)); break;
    case 512: VITATK_CUDA_OK(launch_pdl(ln_bwd_kernel<2>, dim3(grid), dim3(256), 0, stream, 1, dy, x, stats, gamma, dres, dx_out, rows, streaming, x_f16, g_f16)); break;
    case 256: VITATK_CUDA_OK(launch_pdl(ln_bwd_kernel<1>, dim3(grid), dim3(256), 0, stream, 1, dy, x, stats, gamma, dres, dx_out, rows, streaming, x_f16, g_f16)); break;
    default: set_error("layernorm_bwd: cols=%d unsupported", cols); return 1;
  }
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int layernorm_bwd_bt(const bf16* dy, const bf16* x, const float2* stats, const float* gamma, const bf16* dres, bf16* dx_out,
                     int rows, int cols, cudaStream_t stream, int x_f16, int g_f16, const bf16* LB, int ksteps, bf16* T, int ldt) {
  if (cols != 768 || ksteps < 1 || ksteps > 4 || LB == nullptr || T == nullptr || (reinterpret_cast<uintptr_t>(LB) & 15) != 0) {
    set_error("layernorm_bwd_bt: cols=%d ksteps=%d unsupported", cols, ksteps);
    return 1;
  }
  const int grid = (rows + 7) / 8;
  static int streaming = -1;
  if (streaming < 0) {
    const char* e = getenv("VITATK_LN_STREAM");
    streaming = (e && e[0] == '0') ? 0 : 1;
  }
#define LNBT(KS) \
  VITATK_CUDA_OK(launch_pdl(ln_bwd_bt_kernel<KS>, dim3(grid), dim3(256), 0, stream, 1, dy, x, stats, gamma, dres, dx_out, rows, streaming, x_f16, g_f16, LB, T, ldt))
  switch (ksteps) {
    case 1: LNBT(1); break;
    case 2: LNBT(2); break;
    case 3: LNBT(3); break;
    default: LNBT(4); break;
  }
#undef LNBT
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// classifier head: final LN on the CLS row + Linear + softmax-CE (+ backward to the hidden state)
// one CTA (256 threads) per image
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
  for (int i = 0; i < static_cast<int>(blockDim.x >> 5); ++i) t += red[i];
  return t;
}

__global__ void __launch_bounds__(256) head_kernel(const bf16* __restrict__ h, const float* __restrict__ gamma,
                                                   const float* __restrict__ beta, const float* __restrict__ Wc,
                                                   const float* __restrict__ bc, const int64_t* __restrict__ labels,
                                                   float* __restrict__ logits, float* __restrict__ loss,
                                                   bf16* __restrict__ dh, int tokens, int dim, int classes, float eps,
                                                   float grad_scale, const float* __restrict__ dlogits, int h_f16,
                                                   int dh_f16, float* __restrict__ y_out, float* __restrict__ dlogits_out) {
  pdl_wait();
  pdl_launch_dependents();
  extern __shared__ float hs[];
  float* xn = hs;              // [dim] normalised (x-mean)*rstd
  float* yv = xn + dim;        // [dim] LN output
  float* dyv = yv + dim;       // [dim] grad wrt LN output
  float* lg = dyv + dim;       // [classes]
  __shared__ float red[8];
  const int b = blockIdx.x, tid = threadIdx.x;
  const bf16* row = h + static_cast<size_t>(b) * tokens * dim;
  float s = 0.f;
  for (int k = tid; k < dim; k += blockDim.x) {
    const float v = h_f16 ? __half2float(reinterpret_cast<const __half*>(row)[k]) : __bfloat162float(row[k]);
    xn[k] = v;
    s += v;
  }
  const float mean = block_sum(s, red) / dim;
  float q = 0.f;
  for (int k = tid; k < dim; k += blockDim.x) {
    const float d = xn[k] - mean;
    q += d * d;
  }
  const float rstd = rsqrtf(block_sum(q, red) / dim + eps);
  for (int k = tid; k < dim; k += blockDim.x) {
    const float xh = (xn[k] - mean) * rstd;
    xn[k] = xh;
    yv[k] = xh * gamma[k] + beta[k];
  }
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  for (int c = warp; c < classes; c += nw) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += yv[k] * __ldg(Wc + static_cast<size_t>(c) * dim + k);
    acc = warp_sum(acc);
    if (lane == 0) lg[c] = acc + bc[c];
  }
  __syncthreads();
  // softmax-CE (every thread redundantly: classes is tiny)
  float mx = -INFINITY;
  for (int c = 0; c < classes; ++c) mx = fmaxf(mx, lg[c]);
  float se = 0.f;
  for (int c = 0; c < classes; ++c) se += expf(lg[c] - mx);
  const float lse = mx + logf(se);
  // label out of [0, classes): F.cross_entropy raises (whitebox_attacks.py:29).  Here the image gets a NaN loss and no
  // one-hot term; nothing is read out of bounds.  (Host labels are range-checked before the launch.)
  const long long yl = labels ? static_cast<long long>(labels[b]) : -1;
  const bool bad_label = labels && (yl < 0 || yl >= classes);
  const int y = (labels && !bad_label) ? static_cast<int>(yl) : -1;
  if (tid < classes) logits[static_cast<size_t>(b) * classes + tid] = lg[tid];
  for (int c = tid + blockDim.x; c < classes; c += blockDim.x) logits[static_cast<size_t>(b) * classes + c] = lg[c];
  if (tid == 0 && loss && y >= 0) loss[b] = lse - lg[y];
  if (tid == 0 && loss && bad_label) loss[b] = __int_as_float(0x7fc00000);
  if (y_out != nullptr)  // LoRA training: the classifier's input (final-LayerNorm output of the CLS row)
    for (int k = tid; k < dim; k += blockDim.x) y_out[static_cast<size_t>(b) * dim + k] = yv[k];
  if (dh == nullptr) return;
  // dlogits = (softmax - onehot) * grad_scale, or the caller's cotangent (vector-Jacobian product); dy = Wc^T dlogits
  __syncthreads();
  for (int c = tid; c < classes; c += blockDim.x)
    lg[c] = dlogits ? dlogits[static_cast<size_t>(b) * classes + c] : expf(lg[c] - lse) - (c == y ? 1.f : 0.f);
  __syncthreads();
  if (dlogits_out != nullptr)  // LoRA training: the classifier's output cotangent (unscaled)
    for (int c = tid; c < classes; c += blockDim.x) dlogits_out[static_cast<size_t>(b) * classes + c] = lg[c];
  for (int k = tid; k < dim; k += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < classes; ++c) acc += lg[c] * __ldg(Wc + static_cast<size_t>(c) * dim + k);
    dyv[k] = acc * grad_scale;
  }
  __syncthreads();
  float s1 = 0.f, s2 = 0.f;
  for (int k = tid; k < dim; k += blockDim.x) {
    const float gd = dyv[k] * gamma[k];
    s1 += gd;
    s2 += gd * xn[k];
  }
  const float m1 = block_sum(s1, red) / dim;
  const float m2 = block_sum(s2, red) / dim;
  bf16* drow = dh + static_cast<size_t>(b) * tokens * dim;
  for (int k = tid; k < dim; k += blockDim.x) {
    const float o = rstd * (dyv[k] * gamma[k] - m1 - xn[k] * m2);
    if (dh_f16) reinterpret_cast<__half*>(drow)[k] = __float2half_rn(o);
    else drow[k] = __float2bfloat16(o);
  }
  // every other token of this image receives zero gradient from the head
  uint4* z = reinterpret_cast<uint4*>(drow + dim);
  const int nz = (tokens - 1) * dim / 8;
  for (int i = tid; i < nz; i += blockDim.x) z[i] = make_uint4(0, 0, 0, 0);
}

int head_fwd_bwd(const bf16* h, const float* gamma, const float* beta, const float* Wc, const float* bc,
                 const int64_t* labels, float* logits, float* loss, bf16* dh, int batch, int tokens, int dim,
                 int classes, float eps, float grad_scale, cudaStream_t stream, const float* dlogits, int h_f16, int dh_f16,
                 float* y_out, float* dlogits_out) {
  if (dim % 8 != 0 || classes > 4096) {
    set_error("head_fwd_bwd: dim=%d classes=%d unsupported", dim, classes);
    return 1;
  }
  const size_t smem = (3 * dim + classes) * sizeof(float);
  VITATK_CUDA_OK(launch_pdl(head_kernel, dim3(batch), dim3(256), smem, stream, 1, h, gamma, beta, Wc, bc, labels, logits,
                            loss, dh, tokens, dim, classes, eps, grad_scale, dlogits, h_f16, dh_f16, y_out, dlogits_out));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// pixel kernels. Image NCHW fp32 [B,3,224,224]; im2col matrix bf16 [B*197, 768]:
//   row = b*197 + 1 + (y/16)*14 + x/16 (row b*197 is the CLS slot, all zero), col = c*256 + (y%16)*16 + x%16
// (the flattening of the HF patch conv weight [768,3,16,16], HF modeling_vit.py:151,166).
// One thread = 8 consecutive pixels of one image row = one 16-byte bf16 chunk of the im2col matrix.
// ------------------------------------------------------------------------------------------------
static constexpr int IMG = 224, PATCH = 16, GRID_P = 14, TOK = 197, PDIM = 768;

__device__ __forceinline__ size_t cols_offset(int b, int c, int y, int x8) {
  const int x = x8 * 8;
  const int rowi = b * TOK + 1 + (y >> 4) * GRID_P + (x >> 4);
  return static_cast<size_t>(rowi) * PDIM + c * 256 + (y & 15) * 16 + (x & 15);
}
__device__ __forceinline__ void decode_idx(long long idx, int& b, int& c, int& y, int& x8) {
  x8 = static_cast<int>(idx % 28);
  idx /= 28;
  y = static_cast<int>(idx % IMG);
  idx /= IMG;
  c = static_cast<int>(idx % 3);
  b = static_cast<int>(idx / 3);
}
__device__ __forceinline__ float u01_hash(uint64_t seed, uint64_t ctr) {
  // splitmix64 finaliser on (seed, counter): stateless, independent of launch geometry / GPU count
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}


// Strict L-inf ball: the reference arithmetic adv = clamp(x0 + delta, 0, 1) can leave fl32(adv - x0) a few
// ulps of eps outside [-eps, eps] (x0 + delta is rounded at the ulp of adv, ~6e-8, eps's ulp is ~2e-9).
// BASELINE's gate is ||adv - x0||_inf <= eps exactly, so move adv one ulp towards x0 when that happens.
__device__ __forceinline__ float enforce_ball(float adv, float x0, float eps) {
  const float d = adv - x0;
  if (d > eps) adv = __uint_as_float(__float_as_uint(adv) - 1u);        // adv > x0 >= 0: one ulp down
  else if (d < -eps) adv = __uint_as_float(__float_as_uint(adv) + 1u);  // 0 <= adv < x0: one ulp up
  return adv;
}

__global__ void __launch_bounds__(256) pgd_init_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                                       float* __restrict__ adv, bf16* __restrict__ cols, int batch,
                                                       PixelNorm nrm, float eps, int use_rng, uint64_t seed,
                                                       uint64_t image_index0) {
  const long long total = static_cast<long long>(batch) * 3 * IMG * 28;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int b, c, y, x8;
  decode_idx(idx, b, c, y, x8);
  const size_t p = (static_cast<size_t>(b * 3 + c) * IMG + y) * IMG + x8 * 8;
  const float4 a0 = __ldg(reinterpret_cast<const float4*>(x0 + p));
  const float4 a1 = __ldg(reinterpret_cast<const float4*>(x0 + p + 4));
  float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  if (noise != nullptr) {
    const float4 n0 = __ldg(reinterpret_cast<const float4*>(noise + p));
    const float4 n1 = __ldg(reinterpret_cast<const float4*>(noise + p + 4));
    const float nn[8] = {n0.x, n0.y, n0.z, n0.w, n1.x, n1.y, n1.z, n1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = enforce_ball(fminf(fmaxf(v[j] + nn[j], 0.f), 1.f), v[j], eps);
  } else if (use_rng) {
    const uint64_t e0 = (image_index0 + b) * (3ull * IMG * IMG) + (static_cast<uint64_t>(c) * IMG + y) * IMG + x8 * 8;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float n = (2.f * u01_hash(seed, e0 + j) - 1.f) * eps;
      v[j] = enforce_ball(fminf(fmaxf(v[j] + n, 0.f), 1.f), v[j], eps);
    }
  }
  *reinterpret_cast<float4*>(adv + p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(adv + p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = (v[j] - nrm.mean[c]) * nrm.inv_std[c];
  *reinterpret_cast<uint4*>(cols + cols_offset(b, c, y, x8)) = pack8(o);
  // CLS slot of the im2col matrix stays zero: thread (c=0,y=0) of each 8-pixel column group clears a share
  if (c == 0 && y < 4) {
    // 4 rows x 28 threads = 112 threads >= 96 chunks of the 768-wide CLS row
    const int chunk = y * 28 + x8;
    if (chunk < PDIM / 8)
      *reinterpret_cast<uint4*>(cols + static_cast<size_t>(b) * TOK * PDIM + chunk * 8) = make_uint4(0, 0, 0, 0);
  }
}

__global__ void __launch_bounds__(256) pgd_update_kernel(const bf16* __restrict__ dcols, const float* __restrict__ x0,
                                                         float* __restrict__ adv, bf16* __restrict__ cols, int batch,
                                                         PixelNorm nrm, float eps, float alpha) {
  pdl_wait();
  pdl_launch_dependents();
  const long long total = static_cast<long long>(batch) * 3 * IMG * 28;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int b, c, y, x8;
  decode_idx(idx, b, c, y, x8);
  const size_t p = (static_cast<size_t>(b * 3 + c) * IMG + y) * IMG + x8 * 8;
  const size_t co = cols_offset(b, c, y, x8);
  float g[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(dcols + co)), g);
  const float4 o0 = __ldg(reinterpret_cast<const float4*>(x0 + p));
  const float4 o1 = __ldg(reinterpret_cast<const float4*>(x0 + p + 4));
  const float4 a0 = *reinterpret_cast<const float4*>(adv + p);
  const float4 a1 = *reinterpret_cast<const float4*>(adv + p + 4);
  const float xo[8] = {o0.x, o0.y, o0.z, o0.w, o1.x, o1.y, o1.z, o1.w};
  float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    // torchattacks order: adv + alpha*sign(g); delta = clamp(adv - x, -eps, eps); adv = clamp(x + delta, 0, 1)
    const float sg = (g[j] > 0.f) ? 1.f : ((g[j] < 0.f) ? -1.f : 0.f);
    const float stepped = v[j] + alpha * sg;
    const float delta = fminf(fmaxf(stepped - xo[j], -eps), eps);
    v[j] = enforce_ball(fminf(fmaxf(xo[j] + delta, 0.f), 1.f), xo[j], eps);
    o[j] = (v[j] - nrm.mean[c]) * nrm.inv_std[c];
  }
  *reinterpret_cast<float4*>(adv + p) = make_float4(v[0], v[1], v[2], v[3]);
  *reinterpret_cast<float4*>(adv + p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  *reinterpret_cast<uint4*>(cols + co) = pack8(o);
}

__global__ void __launch_bounds__(256) grad_to_image_kernel(const bf16* __restrict__ dcols, float* __restrict__ grad,
                                                            int batch, PixelNorm nrm, float scale) {
  const long long total = static_cast<long long>(batch) * 3 * IMG * 28;
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  int b, c, y, x8;
  decode_idx(idx, b, c, y, x8);
  const size_t p = (static_cast<size_t>(b * 3 + c) * IMG + y) * IMG + x8 * 8;
  float g[8];
  unpack8(__ldg(reinterpret_cast<const uint4*>(dcols + cols_offset(b, c, y, x8))), g);
  const float k = nrm.inv_std[c] * scale;
  *reinterpret_cast<float4*>(grad + p) = make_float4(g[0] * k, g[1] * k, g[2] * k, g[3] * k);
  *reinterpret_cast<float4*>(grad + p + 4) = make_float4(g[4] * k, g[5] * k, g[6] * k, g[7] * k);
}

__global__ void count_correct_kernel(const float* __restrict__ logits, const int64_t* __restrict__ labels, int batch,
                                     int classes, unsigned long long* counts) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int ok = 0;
  if (b < batch) {
    const float* l = logits + static_cast<size_t>(b) * classes;
    int best = 0;
    float bv = l[0];
    for (int c = 1; c < classes; ++c)
      if (l[c] > bv) {  // first maximum wins, like torch.argmax
        bv = l[c];
        best = c;
      }
    ok = (best == static_cast<int>(labels[b])) ? 1 : 0;
  }
  const unsigned m = __ballot_sync(0xffffffffu, ok);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(counts, static_cast<unsigned long long>(__popc(m)));
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(counts + 1, static_cast<unsigned long long>(batch));
}

static inline int pixel_grid(int batch) {
  const long long total = static_cast<long long>(batch) * 3 * IMG * 28;
  return static_cast<int>((total + 255) / 256);
}

int pgd_init(const float* x0, const float* noise, float* adv, bf16* cols, int batch, PixelNorm nrm, float eps,
             int use_rng, uint64_t seed, uint64_t image_index0, cudaStream_t stream) {
  pgd_init_kernel<<<pixel_grid(batch), 256, 0, stream>>>(x0, noise, adv, cols, batch, nrm, eps, use_rng, seed,
                                                         image_index0);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}
int pgd_update(const bf16* dcols, const float* x0, float* adv, bf16* cols, int batch, PixelNorm nrm, float eps,
               float alpha, cudaStream_t stream) {
  VITATK_CUDA_OK(launch_pdl(pgd_update_kernel, dim3(pixel_grid(batch)), dim3(256), 0, stream, 1, dcols, x0, adv, cols,
                            batch, nrm, eps, alpha));
  return 0;
}
// Utils.py:106-113 save_images (clamp to [0,1], *255, truncate to uint8) followed by what re-loading the PNG with ToTensor
// gives back (/255): the pixel values the reference's evaluation scripts actually see (train_loras.py:56-76 reads the
// saved files).  HBM-bound: 4 B read + 4 B write (+ 1 B for the optional uint8 HWC copy) per pixel-channel.
__global__ void __launch_bounds__(256) png_roundtrip_kernel(const float4* __restrict__ in, float4* __restrict__ out,
                                                            size_t n4) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const float4 v = __ldg(in + i);
    float4 o;
    o.x = __fdiv_rn(static_cast<float>(__float2uint_rz(__fmul_rn(fminf(fmaxf(v.x, 0.f), 1.f), 255.f))), 255.f);
    o.y = __fdiv_rn(static_cast<float>(__float2uint_rz(__fmul_rn(fminf(fmaxf(v.y, 0.f), 1.f), 255.f))), 255.f);
    o.z = __fdiv_rn(static_cast<float>(__float2uint_rz(__fmul_rn(fminf(fmaxf(v.z, 0.f), 1.f), 255.f))), 255.f);
    o.w = __fdiv_rn(static_cast<float>(__float2uint_rz(__fmul_rn(fminf(fmaxf(v.w, 0.f), 1.f), 255.f))), 255.f);
    out[i] = o;
  }
}

// uint8 HWC image the reference hands to PIL (Utils.py:111-113): [B,224,224,3] from NCHW fp32
__global__ void __launch_bounds__(256) to_uint8_hwc_kernel(const float* __restrict__ in, uint8_t* __restrict__ out,
                                                           int batch) {
  const size_t total = static_cast<size_t>(batch) * 224 * 224;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const size_t b = i / (224 * 224), px = i % (224 * 224);
    const float* src = in + b * 3 * 224 * 224 + px;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float v = fminf(fmaxf(__ldg(src + c * 224 * 224), 0.f), 1.f);
      out[i * 3 + c] = static_cast<uint8_t>(__float2uint_rz(__fmul_rn(v, 255.f)));
    }
  }
}

int png_roundtrip(const float* images, float* out, uint8_t* u8_hwc, int batch, cudaStream_t stream) {
  const size_t n = static_cast<size_t>(batch) * 3 * 224 * 224;
  if (out) {
    png_roundtrip_kernel<<<pixel_grid(batch), 256, 0, stream>>>(reinterpret_cast<const float4*>(images),
                                                                reinterpret_cast<float4*>(out), n / 4);
    VITATK_CUDA_OK(cudaGetLastError());
  }
  if (u8_hwc) {
    to_uint8_hwc_kernel<<<pixel_grid(batch), 256, 0, stream>>>(images, u8_hwc, batch);
    VITATK_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

// One-off (finalize): copy of a LoRA up-projection lb [rows, 64] whose columns col.. carry the consumer GEMM's per-column
// constants as bf16 hi/lo splits, so the tensor core adds them (rank-1 updates against matching columns of T):
//   c1 != null: [c1_hi, c1_lo, c1_hi, c2_hi, c2_lo, c2_hi]   (folded LayerNorm; T holds [-mh, -mh, -ml, sh, sh, sl])
//   c1 == null: [c2_hi, c2_lo]                               (plain bias; T holds [1, 1])
__global__ void lora_const_columns_kernel(const bf16* __restrict__ lb, const float* __restrict__ c1,
                                          const float* __restrict__ c2, bf16* __restrict__ out, int rows, int col) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= rows) return;
  for (int j = 0; j < 64; ++j) out[static_cast<size_t>(n) * 64 + j] = lb[static_cast<size_t>(n) * 64 + j];
  bf16* o = out + static_cast<size_t>(n) * 64 + col;
  const float b = c2[n];
  const bf16 bh = __float2bfloat16(b), bl = __float2bfloat16(b - __bfloat162float(bh));
  if (c1 != nullptr) {
    const float a = c1[n];
    const bf16 ah = __float2bfloat16(a), al = __float2bfloat16(a - __bfloat162float(ah));
    o[0] = ah; o[1] = al; o[2] = ah; o[3] = bh; o[4] = bl; o[5] = bh;
  } else {
    o[0] = bh; o[1] = bl;
  }
}

int lora_const_columns(const bf16* lb, const float* c1, const float* c2, bf16* out, int rows, int col,
                       cudaStream_t stream) {
  lora_const_columns_kernel<<<(rows + 127) / 128, 128, 0, stream>>>(lb, c1, c2, out, rows, col);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int grad_to_image(const bf16* dcols, float* grad, int batch, PixelNorm nrm, float scale, cudaStream_t stream) {
  grad_to_image_kernel<<<pixel_grid(batch), 256, 0, stream>>>(dcols, grad, batch, nrm, scale);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}
int count_correct(const float* logits, const int64_t* labels, int batch, int classes, long long* counts,
                  cudaStream_t stream) {
  count_correct_kernel<<<(batch + 127) / 128, 128, 0, stream>>>(logits, labels, batch, classes,
                                                                reinterpret_cast<unsigned long long*>(counts));
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

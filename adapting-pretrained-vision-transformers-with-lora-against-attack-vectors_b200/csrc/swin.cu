// Swin Transformer (shifted-window attention) on the attack engine (BASELINE configs[2]: PGD-20 on LoRA
// Swin-B).  The reference only names the model family (README.md:53); the arithmetic restated here is HF
// transformers' SwinForImageClassification (modeling_swin.py: window_partition :141-150, patch merging :326-349,
// window attention with relative-position bias and the -100 region mask :410-459,:556-582, cyclic shift :615-636,
// pooled head :870-915), forward and input-gradient backward.
//
// Every Linear (qkv, proj, fc1, fc2, patch embed, patch-merging reduction) runs on the same tcgen05 GEMM as the ViT path
// (bias / residual / GELU+GELU' / LoRA k-block epilogues; stage widths 128 / 256 / 512 / 1024).  New here:
//   win_attn_fwd / win_attn_bwd   7x7-window attention, head dim 32: one four-warp CTA per (window, head), the five small
//                                 products on warp-level bf16 MMAs; the cyclic shift, the window partition and their
//                                 inverses are index arithmetic on the token-major q|k|v rows (nothing is rolled or
//                                 re-laid out in memory), the relative-position bias comes from a [heads, 49, 49] table,
//                                 the shifted-window mask from each token's region id.  The backward recomputes the
//                                 49x49 probabilities (they are never stored).
//   merge_gather / merge_scatter  2x2 patch merging as a row permutation ([B,H,W,C] -> [B,H/2,W/2,4C]) and its inverse
//   ln_any_fwd / ln_any_bwd       LayerNorm for any width (128 ... 2048 here), one warp per row
//   swin_head                     final LayerNorm on all 49 tokens + mean pool + classifier + CE (+ backward)
//   swin pixel kernels            PGD init / update / gradient materialisation for the 4x4 patch-embedding layout
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <vector>

#include "../../include/vitatk.h"
#include "vitatk_internal.h"
#include "mma_sync.cuh"

namespace vitatk {
namespace {

constexpr int WIN = 7, WT = 49, HD = 32, IMG = 224;

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
__device__ __forceinline__ uint4 pack8(const float* f) {
  uint4 q;
  __nv_bfloat162* p = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) p[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  return q;
}

// ------------------------------------------------------------------------------------------------
// LayerNorm for any width (cols % 8 == 0): one warp per row, two passes over the row (the second one hits L1 / L2)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_any_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, bf16* __restrict__ y,
                                                         float2* __restrict__ stats, int rows, int cols, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * cols);
  const int nch = cols >> 3;
  float s = 0.f;
  for (int c = lane; c < nch; c += 32) {
    float v[8];
    unpack8(__ldg(xr + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[j];
  }
  const float mean = warp_sum(s) / cols;
  float q = 0.f;
  for (int c = lane; c < nch; c += 32) {
    float v[8];
    unpack8(__ldg(xr + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) q = fmaf(v[j] - mean, v[j] - mean, q);
  }
  const float rstd = rsqrtf(warp_sum(q) / cols + eps);
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * cols);
  for (int c = lane; c < nch; c += 32) {
    float v[8], o[8];
    unpack8(__ldg(xr + c), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = (v[j] - mean) * rstd * __ldg(gamma + c * 8 + j) + __ldg(beta + c * 8 + j);
    yr[c] = pack8(o);
  }
  if (lane == 0) stats[row] = make_float2(mean, rstd);
}

// dx = dres + LN_backward(dy)   (dres may be null)
__global__ void __launch_bounds__(256) ln_any_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                         const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                         const bf16* __restrict__ dres, bf16* __restrict__ dx, int rows, int cols) {
  pdl_wait();
  pdl_launch_dependents();
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(row) * cols);
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + static_cast<size_t>(row) * cols);
  const int nch = cols >> 3;
  const float2 st = stats[row];
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < nch; c += 32) {
    float xv[8], dv[8];
    unpack8(__ldg(xr + c), xv);
    unpack8(__ldg(dyr + c), dv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gd = dv[j] * __ldg(gamma + c * 8 + j);
      s1 += gd;
      s2 = fmaf(gd, (xv[j] - st.x) * st.y, s2);
    }
  }
  const float m1 = warp_sum(s1) / cols, m2 = warp_sum(s2) / cols;
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + static_cast<size_t>(row) * cols) : nullptr;
  uint4* dxr = reinterpret_cast<uint4*>(dx + static_cast<size_t>(row) * cols);
  for (int c = lane; c < nch; c += 32) {
    float xv[8], dv[8], r[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
    unpack8(__ldg(xr + c), xv);
    unpack8(__ldg(dyr + c), dv);
    if (rr) unpack8(__ldg(rr + c), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float gd = dv[j] * __ldg(gamma + c * 8 + j);
      o[j] = r[j] + st.y * (gd - m1 - (xv[j] - st.x) * st.y * m2);
    }
    dxr[c] = pack8(o);
  }
}

// The widths the Swin path uses (128 ... 2048 = 8 * LPR * CPL columns): LPR lanes per row, CPL 16-byte chunks per lane, the
// row lives in registers -- one global read, one write (the kernels are pure HBM traffic: 4 B per element forward, 8 B
// backward with the residual gradient).
template <int LPR>
__device__ __forceinline__ float row_sum(float v) {
#pragma unroll
  for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <int LPR, int CPL>
__global__ void __launch_bounds__(256) ln_reg_fwd_kernel(const bf16* __restrict__ x, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, bf16* __restrict__ y,
                                                         float2* __restrict__ stats, int rows, float eps) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int COLS = 8 * LPR * CPL, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  const uint4* xr = reinterpret_cast<const uint4*>(x + static_cast<size_t>(ok ? row : 0) * COLS);
  float v[CPL][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    unpack8(__ldg(xr + sub + k * LPR), v[k]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += v[k][j];
  }
  const float mean = row_sum<LPR>(s) * (1.f / COLS);
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k)
#pragma unroll
    for (int j = 0; j < 8; ++j) q = fmaf(v[k][j] - mean, v[k][j] - mean, q);
  const float rstd = rsqrtf(row_sum<LPR>(q) * (1.f / COLS) + eps);
  if (!ok) return;
  uint4* yr = reinterpret_cast<uint4*>(y + static_cast<size_t>(row) * COLS);
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = sub + k * LPR;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c), b1 = __ldg(reinterpret_cast<const float4*>(beta) + 2 * c + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w}, bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = fmaf((v[k][j] - mean) * rstd, gg[j], bb[j]);
    yr[c] = pack8(o);
  }
  if (sub == 0) stats[row] = make_float2(mean, rstd);
}
template <int LPR, int CPL>
__global__ void __launch_bounds__(256) ln_reg_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                         const float2* __restrict__ stats, const float* __restrict__ gamma,
                                                         const bf16* __restrict__ dres, bf16* __restrict__ dx, int rows) {
  pdl_wait();
  pdl_launch_dependents();
  constexpr int COLS = 8 * LPR * CPL, RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, sub = lane % LPR;
  const int row = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW + lane / LPR;
  const bool ok = row < rows;
  const size_t off = static_cast<size_t>(ok ? row : 0) * COLS;
  const uint4* xr = reinterpret_cast<const uint4*>(x + off);
  const uint4* dyr = reinterpret_cast<const uint4*>(dy + off);
  const float2 st = stats[ok ? row : 0];
  float xh[CPL][8], gd[CPL][8];  // (x - mean) * rstd and gamma * dy
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = sub + k * LPR;
    unpack8(__ldg(xr + c), xh[k]);
    unpack8(__ldg(dyr + c), gd[k]);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c), g1 = __ldg(reinterpret_cast<const float4*>(gamma) + 2 * c + 1);
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      xh[k][j] = (xh[k][j] - st.x) * st.y;
      gd[k][j] *= gg[j];
      s1 += gd[k][j];
      s2 = fmaf(gd[k][j], xh[k][j], s2);
    }
  }
  const float m1 = row_sum<LPR>(s1) * (1.f / COLS), m2 = row_sum<LPR>(s2) * (1.f / COLS);
  if (!ok) return;
  const uint4* rr = dres ? reinterpret_cast<const uint4*>(dres + off) : nullptr;
  uint4* dxr = reinterpret_cast<uint4*>(dx + off);
#pragma unroll
  for (int k = 0; k < CPL; ++k) {
    const int c = sub + k * LPR;
    float r[8] = {0, 0, 0, 0, 0, 0, 0, 0}, o[8];
    if (rr) unpack8(__ldg(rr + c), r);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = r[j] + st.y * (gd[k][j] - m1 - xh[k][j] * m2);
    dxr[c] = pack8(o);
  }
}

int ln_any_fwd(const bf16* x, const float* g, const float* b, bf16* y, float2* st, int rows, int cols, float eps, cudaStream_t s) {
#define LN_FWD(LPR, CPL)                                                                                                    \
  VITATK_CUDA_OK(launch_pdl(ln_reg_fwd_kernel<LPR, CPL>, dim3((rows + 8 * (32 / LPR) - 1) / (8 * (32 / LPR))), dim3(256), 0, s, 1, x, g, b, y, st, rows, eps))
  switch (((reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(b)) & 15) ? 0 : cols) {
    case 128: LN_FWD(16, 1); break;
    case 256: LN_FWD(32, 1); break;
    case 512: LN_FWD(32, 2); break;
    case 1024: LN_FWD(32, 4); break;
    case 2048: LN_FWD(32, 8); break;
    default: VITATK_CUDA_OK(launch_pdl(ln_any_fwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, 1, x, g, b, y, st, rows, cols, eps));
  }
#undef LN_FWD
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}
int ln_any_bwd(const bf16* dy, const bf16* x, const float2* st, const float* g, const bf16* dres, bf16* dx, int rows, int cols,
               cudaStream_t s) {
#define LN_BWD(LPR, CPL)                                                                                                    \
  VITATK_CUDA_OK(launch_pdl(ln_reg_bwd_kernel<LPR, CPL>, dim3((rows + 8 * (32 / LPR) - 1) / (8 * (32 / LPR))), dim3(256), 0, s, 1, dy, x, st, g, dres, dx, rows))
  switch ((reinterpret_cast<uintptr_t>(g) & 15) ? 0 : cols) {
    case 128: LN_BWD(16, 1); break;
    case 256: LN_BWD(32, 1); break;
    case 512: LN_BWD(32, 2); break;
    case 1024: LN_BWD(32, 4); break;
    case 2048: LN_BWD(32, 8); break;
    default: VITATK_CUDA_OK(launch_pdl(ln_any_bwd_kernel, dim3((rows + 7) / 8), dim3(256), 0, s, 1, dy, x, st, g, dres, dx, rows, cols));
  }
#undef LN_BWD
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// 2x2 patch merging (modeling_swin.py:326-345): out[b, y2, x2, k*C + c] = in[b, 2*y2 + (k & 1), 2*x2 + (k >> 1), c]
// (k = 0..3 in HF's concat order: (0,0), (1,0), (0,1), (1,1)).  One thread per 16-byte chunk.  scatter = the inverse.
// ------------------------------------------------------------------------------------------------
__global__ void merge_permute_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int batch, int R, int C, int scatter) {
  pdl_wait();
  pdl_launch_dependents();
  const int chunks = C >> 3, R2 = R >> 1;
  const long long total = static_cast<long long>(batch) * R2 * R2 * 4 * chunks;
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % chunks);
  long long t = i / chunks;
  const int k = static_cast<int>(t % 4);
  t /= 4;
  const int x2 = static_cast<int>(t % R2);
  t /= R2;
  const int y2 = static_cast<int>(t % R2), b = static_cast<int>(t / R2);
  const size_t fine = ((static_cast<size_t>(b) * R + 2 * y2 + (k & 1)) * R + 2 * x2 + (k >> 1)) * C + ch * 8;
  const size_t coarse = ((static_cast<size_t>(b) * R2 + y2) * R2 + x2) * (4 * C) + k * C + ch * 8;
  if (scatter) *reinterpret_cast<uint4*>(out + fine) = *reinterpret_cast<const uint4*>(in + coarse);
  else *reinterpret_cast<uint4*>(out + coarse) = *reinterpret_cast<const uint4*>(in + fine);
}
int merge_permute(const bf16* in, bf16* out, int batch, int R, int C, int scatter, cudaStream_t s) {
  const long long total = static_cast<long long>(batch) * (R / 2) * (R / 2) * 4 * (C / 8);
  VITATK_CUDA_OK(launch_pdl(merge_permute_kernel, dim3(static_cast<unsigned>((total + 255) / 256)), dim3(256), 0, s, 1, in, out, batch, R, C, scatter));
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Window attention.  Token rows are (b, y, x) row-major over an R x R grid; q|k|v packed [M, 3C], head h = columns
// h*32 .. h*32+31 of each third.  Window (wy, wx) holds the tokens whose SHIFTED coordinates y' = (y - shift) mod R fall
// in [7 wy, 7 wy + 7): torch.roll(-shift) + window_partition without moving data.
//
// A (window, head) unit is 49 x 49 x 32: 12 KB of HBM traffic forward, 24 KB backward, against 0.3 / 0.8 MFLOP -- the
// kernels are HBM-bound byte work once the arithmetic is off the FP32 pipe.  One CTA of four warps per unit; the 49
// tokens are padded to 64 rows in shared memory; each warp owns 16 query rows (16 key rows in the second backward
// phase) and runs the five small products on warp-level m16n8k16 bf16 MMAs (fp32 accumulate) straight from ldmatrix
// fragments -- a 64 x 64 x 32 product is far below one tcgen05 tile, and the TMEM / mbarrier hand-offs of the UMMA path
// cost more than the whole unit.  The probabilities never leave registers in the forward (the S accumulator fragments
// are re-packed as the A operand of P*V); the backward recomputes them.
// ------------------------------------------------------------------------------------------------
struct WinGeom {
  int R, nW, heads, C, shift;
};
__device__ __forceinline__ int win_token_row(const WinGeom& g, int b, int win, int t) {
  const int wpr = g.R / WIN;
  const int wy = win / wpr, wx = win % wpr, iy = t / WIN, ix = t % WIN;
  int y = wy * WIN + iy + g.shift, x = wx * WIN + ix + g.shift;
  if (y >= g.R) y -= g.R;
  if (x >= g.R) x -= g.R;
  return (b * g.R + y) * g.R + x;
}
constexpr int WPAD = 64;  // tokens of a window, padded to four 16-row MMA tiles
constexpr int QS = 40;    // row stride (elements) of the [64][32] operand tiles: 80 B, ldmatrix conflict-free
constexpr int PS = 72;    // row stride of the [64][64] P / dS tiles: 144 B
constexpr int WA_THREADS = 128;


// Loading a unit: thread tid owns the 16-byte chunk (tid & 3) of token rows t0 = tid >> 2 (always a real token) and
// t1 = t0 + 32 (a real token when < 49, else a zero row of the padding).  All global loads of a thread are issued before
// the first shared-memory store.
struct WinRows {
  size_t r0, r1;  // global token rows
  bool has1;
};
__device__ __forceinline__ WinRows win_rows(const WinGeom& g, int b, int win, int tid, int* trow) {
  const int t0 = tid >> 2, t1 = t0 + 32;
  WinRows w;
  w.has1 = t1 < WT;
  const int a0 = win_token_row(g, b, win, t0), a1 = w.has1 ? win_token_row(g, b, win, t1) : 0;
  if ((tid & 3) == 0) {
    trow[t0] = a0;
    trow[t1] = a1;
  }
  w.r0 = static_cast<size_t>(a0);
  w.r1 = static_cast<size_t>(a1);
  return w;
}
template <int N>
__device__ __forceinline__ void win_load_tiles(bf16* const (&dst)[N], const bf16* const (&src)[N], const size_t (&ld)[N], const WinRows& w,
                                               int tid) {
  const int d8 = (tid & 3) * 8, t0 = tid >> 2;
  uint4 v0[N], v1[N];
#pragma unroll
  for (int i = 0; i < N; ++i) {
    v0[i] = __ldg(reinterpret_cast<const uint4*>(src[i] + w.r0 * ld[i] + d8));
    v1[i] = w.has1 ? __ldg(reinterpret_cast<const uint4*>(src[i] + w.r1 * ld[i] + d8)) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int i = 0; i < N; ++i) {
    *reinterpret_cast<uint4*>(dst[i] + t0 * QS + d8) = v0[i];
    *reinterpret_cast<uint4*>(dst[i] + (t0 + 32) * QS + d8) = v1[i];
  }
}
// 16 rows [m0, m0 + 16) of a staged [64][QS] tile -> global rows (row stride ld), 16-byte stores
__device__ __forceinline__ void win_store_rows(bf16* dst, size_t ld, const bf16* tile, const int* trow, int m0, int lane) {
#pragma unroll
  for (int c = lane; c < 64; c += 32) {
    const int t = m0 + (c >> 2), d8 = (c & 3) * 8;
    if (t < WT) *reinterpret_cast<uint4*>(dst + static_cast<size_t>(trow[t]) * ld + d8) = *reinterpret_cast<const uint4*>(tile + t * QS + d8);
  }
}
// accumulator fragments of a [16][32] tile (four n-tiles) -> rows m0 + g, m0 + g + 8 of a staged tile, scaled
__device__ __forceinline__ void win_stage_acc(bf16* tile, const float (&o)[4][4], int m0, int lane, float s_lo, float s_hi) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int nt = 0; nt < 4; ++nt) {
    *reinterpret_cast<uint32_t*>(tile + (m0 + g) * QS + nt * 8 + 2 * t) = pack_bf16x2(o[nt][0] * s_lo, o[nt][1] * s_lo);
    *reinterpret_cast<uint32_t*>(tile + (m0 + g + 8) * QS + nt * 8 + 2 * t) = pack_bf16x2(o[nt][2] * s_hi, o[nt][3] * s_hi);
  }
}
// Relative-position bias + shifted-window mask in FRAGMENT order: for every (variant, head) a [warp 4][n-tile 7][lane 32]
// array of float4 = the four accumulator elements (rows g, g + 8; key columns 2t, 2t + 1) of that lane, so the softmax
// prologue is seven coalesced 16-byte loads and 28 FMAs with no index arithmetic, compares or branches.  Key columns
// >= 49 hold -inf, padded query rows 0.  variant: bit 1 = window in the last window row, bit 0 = in the last window column
// of a shifted block (modeling_swin.py:556-575: only those windows straddle the cyclic seam and mask by region).
constexpr int BIAS_FRAG_F4 = 4 * 7 * 32;  // float4 per (variant, head)
__global__ void relbias_frag_kernel(const float* __restrict__ bias, float4* __restrict__ out, int heads, int shift, int variants) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= variants * heads * BIAS_FRAG_F4) return;
  const int lane = idx & 31, nt = (idx >> 5) % 7, warp = (idx / 224) & 3, h = (idx / BIAS_FRAG_F4) % heads, v = idx / (BIAS_FRAG_F4 * heads);
  const int g = lane >> 2, t = lane & 3;
  float val[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const int i = warp * 16 + g + (e >> 1) * 8, j = nt * 8 + 2 * t + (e & 1);
    float x = 0.f;
    if (j >= WT) x = -INFINITY;
    else if (i < WT) {
      x = bias[(static_cast<size_t>(h) * WT + i) * WT + j];
      const int ryi = (v & 2) ? (i / WIN < WIN - shift ? 1 : 2) : 0, ryj = (v & 2) ? (j / WIN < WIN - shift ? 1 : 2) : 0;
      const int rxi = (v & 1) ? (i % WIN < WIN - shift ? 1 : 2) : 0, rxj = (v & 1) ? (j % WIN < WIN - shift ? 1 : 2) : 0;
      if (ryi != ryj || rxi != rxj) x -= 100.f;
    }
    val[e] = x;
  }
  out[idx] = make_float4(val[0], val[1], val[2], val[3]);
}
int relbias_frag(const float* bias, float* out, int heads, int shift, cudaStream_t s) {
  const int variants = shift ? 4 : 1, total = variants * heads * BIAS_FRAG_F4;
  relbias_frag_kernel<<<(total + 255) / 256, 256, 0, s>>>(bias, reinterpret_cast<float4*>(out), heads, shift, variants);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}
size_t relbias_frag_bytes(int heads, int shift) { return static_cast<size_t>(shift ? 4 : 1) * heads * BIAS_FRAG_F4 * sizeof(float4); }
__device__ __forceinline__ const float4* win_bias_tab(const float* tab, const WinGeom& g, int win, int h, int warp) {
  const int wpr = g.R / WIN;
  const int v = g.shift ? ((win / wpr == wpr - 1 ? 2 : 0) | (win % wpr == wpr - 1 ? 1 : 0)) : 0;
  return reinterpret_cast<const float4*>(tab) + (static_cast<size_t>(v) * g.heads + h) * BIAS_FRAG_F4 + warp * (7 * 32);
}

// S = scale * Q K^T + (bias + region mask) for the warp's 16 query rows, then the row softmax statistics.
// On return p[nt][e] = exp(s - rowmax) (0 for key columns >= 49), inv_* = 1 / rowsum (0 for query rows >= 49).
__device__ __forceinline__ void win_scores(float (&p)[8][4], float& inv_lo, float& inv_hi, const bf16* Qs, const bf16* Ks,
                                           const float4* __restrict__ tab, float scale, int m0, int lane) {
  const int g = lane >> 2;
  float4 bz[7];
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) bz[nt] = __ldg(tab + nt * 32 + lane);
#pragma unroll
  for (int nt = 0; nt < 8; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) p[nt][e] = 0.f;
  uint32_t aq[2][4];
#pragma unroll
  for (int ks = 0; ks < 2; ++ks) ldsm_x4(aq[ks], smem_u32(Qs + (m0 + frag_r_lo(lane)) * QS + ks * 16 + frag_c_hi(lane)));
#pragma unroll
  for (int np = 0; np < 4; ++np) {
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      uint32_t bk[4];
      ldsm_x4(bk, smem_u32(Ks + (np * 16 + frag_r_hi(lane)) * QS + ks * 16 + frag_c_lo(lane)));
      mma16816(p[2 * np], aq[ks], bk[0], bk[1]);
      if (np < 3) mma16816(p[2 * np + 1], aq[ks], bk[2], bk[3]);
    }
  }
  float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) {
    p[nt][0] = fmaf(p[nt][0], scale, bz[nt].x);
    p[nt][1] = fmaf(p[nt][1], scale, bz[nt].y);
    p[nt][2] = fmaf(p[nt][2], scale, bz[nt].z);
    p[nt][3] = fmaf(p[nt][3], scale, bz[nt].w);
    mx_lo = fmaxf(mx_lo, fmaxf(p[nt][0], p[nt][1]));
    mx_hi = fmaxf(mx_hi, fmaxf(p[nt][2], p[nt][3]));
  }
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
  mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
  mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
  // exp(s - mx) = 2^(s * log2e - mx * log2e): one FFMA + one MUFU.EX2 per element
  constexpr float LOG2E = 1.4426950408889634f;
  const float nm_lo = -mx_lo * LOG2E, nm_hi = -mx_hi * LOG2E;
  float sum_lo = 0.f, sum_hi = 0.f;
#pragma unroll
  for (int nt = 0; nt < 7; ++nt) {
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      p[nt][e] = ex2_approx(fmaf(p[nt][e], LOG2E, nm_lo));
      p[nt][2 + e] = ex2_approx(fmaf(p[nt][2 + e], LOG2E, nm_hi));
      sum_lo += p[nt][e];
      sum_hi += p[nt][2 + e];
    }
  }
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 1);
  sum_lo += __shfl_xor_sync(0xffffffffu, sum_lo, 2);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 1);
  sum_hi += __shfl_xor_sync(0xffffffffu, sum_hi, 2);
  inv_lo = m0 + g < WT ? 1.f / sum_lo : 0.f;
  inv_hi = m0 + g + 8 < WT ? 1.f / sum_hi : 0.f;
}
// acc[16][32] += X[16][64] * Y[64][32]: X from accumulator-layout registers (eight n-tiles = four k-steps), Y stored [k][n]
__device__ __forceinline__ void win_mma_regA(float (&acc)[4][4], const float (&x)[8][4], const bf16* Ys, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    a[0] = pack_bf16x2(x[2 * ks][0], x[2 * ks][1]);
    a[1] = pack_bf16x2(x[2 * ks][2], x[2 * ks][3]);
    a[2] = pack_bf16x2(x[2 * ks + 1][0], x[2 * ks + 1][1]);
    a[3] = pack_bf16x2(x[2 * ks + 1][2], x[2 * ks + 1][3]);
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(b, smem_u32(Ys + (ks * 16 + frag_r_lo(lane)) * QS + dp * 16 + frag_c_hi(lane)));
      mma16816(acc[2 * dp], a, b[0], b[1]);
      mma16816(acc[2 * dp + 1], a, b[2], b[3]);
    }
  }
}
// acc[16][32] += X^T[16][64] * Y[64][32] for output rows [m0, m0 + 16): X stored [k = 64][m = 64] (stride PS), Y stored [k][n]
__device__ __forceinline__ void win_mma_transA(float (&acc)[4][4], const bf16* Xs, const bf16* Ys, int m0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
    uint32_t a[4];
    ldsm_x4_t(a, smem_u32(Xs + (ks * 16 + frag_r_hi(lane)) * PS + m0 + frag_c_lo(lane)));
#pragma unroll
    for (int dp = 0; dp < 2; ++dp) {
      uint32_t b[4];
      ldsm_x4_t(b, smem_u32(Ys + (ks * 16 + frag_r_lo(lane)) * QS + dp * 16 + frag_c_hi(lane)));
      mma16816(acc[2 * dp], a, b[0], b[1]);
      mma16816(acc[2 * dp + 1], a, b[2], b[3]);
    }
  }
}

__global__ void __launch_bounds__(WA_THREADS) win_attn_fwd_kernel(const bf16* __restrict__ qkv, const float* __restrict__ bias,
                                                                  bf16* __restrict__ out, WinGeom g, float scale) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) bf16 Qs[WPAD * QS], Ks[WPAD * QS], Vs[WPAD * QS];
  __shared__ int trow[WPAD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x;
  const int h = item % g.heads, win = (item / g.heads) % g.nW, b = item / (g.heads * g.nW);
  const size_t ld = 3 * static_cast<size_t>(g.C);
  {
    const WinRows wr = win_rows(g, b, win, tid, trow);
    const bf16* base = qkv + h * HD;
    bf16* const dst[3] = {Qs, Ks, Vs};
    const bf16* const srcs[3] = {base, base + g.C, base + 2 * g.C};
    const size_t lds[3] = {ld, ld, ld};
    win_load_tiles<3>(dst, srcs, lds, wr, tid);
  }
  __syncthreads();
  const int m0 = warp * 16;
  float p[8][4], inv_lo, inv_hi;
  win_scores(p, inv_lo, inv_hi, Qs, Ks, win_bias_tab(bias, g, win, h, warp), scale, m0, lane);
  float o[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[nt][e] = 0.f;
  win_mma_regA(o, p, Vs, lane);
  // the warp's own 16 Q rows are dead (only this warp read them): stage O there, then 16-byte stores
  __syncwarp();
  win_stage_acc(Qs, o, m0, lane, inv_lo, inv_hi);
  __syncwarp();
  win_store_rows(out + h * HD, static_cast<size_t>(g.C), Qs, trow, m0, lane);
}

// Backward: recompute P, dP = dO V^T, D = rowsum(P o dP), dS = scale * P o (dP - D); dQ = dS K (phase 1, warp = 16 query
// rows); dK = dS^T Q, dV = P^T dO (phase 2, warp = 16 key rows, P and dS through shared memory).
__global__ void __launch_bounds__(WA_THREADS) win_attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout,
                                                                  const float* __restrict__ bias, bf16* __restrict__ dqkv,
                                                                  WinGeom g, float scale) {
  pdl_wait();
  pdl_launch_dependents();
  __shared__ __align__(16) bf16 Qs[WPAD * QS], Ks[WPAD * QS], Vs[WPAD * QS], Gs[WPAD * QS];
  __shared__ __align__(16) bf16 Ps[WPAD * PS], Ds[WPAD * PS];
  __shared__ int trow[WPAD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int item = blockIdx.x;
  const int h = item % g.heads, win = (item / g.heads) % g.nW, b = item / (g.heads * g.nW);
  const size_t ld = 3 * static_cast<size_t>(g.C);
  {
    const WinRows wr = win_rows(g, b, win, tid, trow);
    const bf16* base = qkv + h * HD;
    bf16* const dst[4] = {Qs, Ks, Vs, Gs};
    const bf16* const srcs[4] = {base, base + g.C, base + 2 * g.C, dout + h * HD};
    const size_t lds[4] = {ld, ld, ld, static_cast<size_t>(g.C)};
    win_load_tiles<4>(dst, srcs, lds, wr, tid);
  }
  __syncthreads();
  const int m0 = warp * 16, gq = lane >> 2, tq = lane & 3;
  float dq[4][4];
  {
    float p[8][4], inv_lo, inv_hi;
    win_scores(p, inv_lo, inv_hi, Qs, Ks, win_bias_tab(bias, g, win, h, warp), scale, m0, lane);
    // dP = dO V^T (V stored [key][dim] = the "col" B operand, like K in the scores)
    float dp[8][4];
#pragma unroll
    for (int nt = 0; nt < 8; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dp[nt][e] = 0.f;
    uint32_t ag[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) ldsm_x4(ag[ks], smem_u32(Gs + (m0 + frag_r_lo(lane)) * QS + ks * 16 + frag_c_hi(lane)));
#pragma unroll
    for (int np = 0; np < 4; ++np) {
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        uint32_t bv[4];
        ldsm_x4(bv, smem_u32(Vs + (np * 16 + frag_r_hi(lane)) * QS + ks * 16 + frag_c_lo(lane)));
        mma16816(dp[2 * np], ag[ks], bv[0], bv[1]);
        if (np < 3) mma16816(dp[2 * np + 1], ag[ks], bv[2], bv[3]);
      }
    }
    float d_lo = 0.f, d_hi = 0.f;
#pragma unroll
    for (int nt = 0; nt < 7; ++nt) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        p[nt][e] *= inv_lo;
        p[nt][2 + e] *= inv_hi;
        d_lo = fmaf(p[nt][e], dp[nt][e], d_lo);
        d_hi = fmaf(p[nt][2 + e], dp[nt][2 + e], d_hi);
      }
    }
    d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 1);
    d_lo += __shfl_xor_sync(0xffffffffu, d_lo, 2);
    d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 1);
    d_hi += __shfl_xor_sync(0xffffffffu, d_hi, 2);
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      // P (bf16) for dV, then dS in place of dP
      *reinterpret_cast<uint32_t*>(Ps + (m0 + gq) * PS + nt * 8 + 2 * tq) = pack_bf16x2(p[nt][0], p[nt][1]);
      *reinterpret_cast<uint32_t*>(Ps + (m0 + gq + 8) * PS + nt * 8 + 2 * tq) = pack_bf16x2(p[nt][2], p[nt][3]);
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        dp[nt][e] = p[nt][e] * (dp[nt][e] - d_lo) * scale;
        dp[nt][2 + e] = p[nt][2 + e] * (dp[nt][2 + e] - d_hi) * scale;
      }
      *reinterpret_cast<uint32_t*>(Ds + (m0 + gq) * PS + nt * 8 + 2 * tq) = pack_bf16x2(dp[nt][0], dp[nt][1]);
      *reinterpret_cast<uint32_t*>(Ds + (m0 + gq + 8) * PS + nt * 8 + 2 * tq) = pack_bf16x2(dp[nt][2], dp[nt][3]);
    }
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) dq[nt][e] = 0.f;
    win_mma_regA(dq, dp, Ks, lane);
  }
  __syncthreads();  // P, dS complete; K, V no longer read
  float dk[4][4], dv[4][4];
#pragma unroll
  for (int nt = 0; nt < 4; ++nt)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[nt][e] = dv[nt][e] = 0.f;
  win_mma_transA(dv, Ps, Gs, m0, lane);
  win_mma_transA(dk, Ds, Qs, m0, lane);
  win_stage_acc(Ks, dk, m0, lane, 1.f, 1.f);
  win_stage_acc(Vs, dv, m0, lane, 1.f, 1.f);
  __syncthreads();  // every warp is done with Q (and dO)
  win_stage_acc(Qs, dq, m0, lane, 1.f, 1.f);
  __syncwarp();
  bf16* dbase = dqkv + h * HD;
  win_store_rows(dbase, ld, Qs, trow, m0, lane);
  win_store_rows(dbase + g.C, ld, Ks, trow, m0, lane);
  win_store_rows(dbase + 2 * g.C, ld, Vs, trow, m0, lane);
}

int win_attn_fwd(const bf16* qkv, const float* bias_frag, bf16* out, int batch, int R, int C, int heads, int shift, cudaStream_t s) {
  WinGeom g = {R, (R / WIN) * (R / WIN), heads, C, shift};
  const int items = batch * g.nW * heads;
  VITATK_CUDA_OK(launch_pdl(win_attn_fwd_kernel, dim3(items), dim3(WA_THREADS), 0, s, 1, qkv, bias_frag, out, g, 1.0f / sqrtf(static_cast<float>(HD))));
  return 0;
}
int win_attn_bwd(const bf16* qkv, const bf16* dout, const float* bias_frag, bf16* dqkv, int batch, int R, int C, int heads, int shift,
                 cudaStream_t s) {
  WinGeom g = {R, (R / WIN) * (R / WIN), heads, C, shift};
  const int items = batch * g.nW * heads;
  VITATK_CUDA_OK(launch_pdl(win_attn_bwd_kernel, dim3(items), dim3(WA_THREADS), 0, s, 1, qkv, dout, bias_frag, dqkv, g, 1.0f / sqrtf(static_cast<float>(HD))));
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Head (modeling_swin.py:886-895, 1075-1080): LayerNorm on each of the T tokens of the last stage, mean over tokens,
// classifier, CE; backward to the last stage's hidden state.  One CTA per image.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) swin_head_kernel(const bf16* __restrict__ h, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ Wc,
                                                        const float* __restrict__ bc, const int64_t* __restrict__ labels,
                                                        float* __restrict__ logits, float* __restrict__ loss, bf16* __restrict__ dh,
                                                        int tokens, int dim, int classes, float eps, float grad_scale) {
  extern __shared__ float hs[];
  float* pooled = hs;             // [dim]
  float* dpool = pooled + dim;    // [dim]
  float* lg = dpool + dim;        // [classes]
  float* tst = lg + classes;      // [tokens][2] mean, rstd
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, nw = blockDim.x >> 5;
  const bf16* base = h + static_cast<size_t>(b) * tokens * dim;
  for (int k = tid; k < dim; k += blockDim.x) pooled[k] = 0.f;
  __syncthreads();
  // per-token statistics (warp per token)
  for (int t = warp; t < tokens; t += nw) {
    float s = 0.f;
    for (int k = lane; k < dim; k += 32) s += __bfloat162float(base[static_cast<size_t>(t) * dim + k]);
    const float mean = warp_sum(s) / dim;
    float q = 0.f;
    for (int k = lane; k < dim; k += 32) {
      const float d = __bfloat162float(base[static_cast<size_t>(t) * dim + k]) - mean;
      q = fmaf(d, d, q);
    }
    const float rstd = rsqrtf(warp_sum(q) / dim + eps);
    if (lane == 0) {
      tst[2 * t] = mean;
      tst[2 * t + 1] = rstd;
    }
  }
  __syncthreads();
  for (int k = tid; k < dim; k += blockDim.x) {
    float acc = 0.f;
    for (int t = 0; t < tokens; ++t)
      acc += (__bfloat162float(base[static_cast<size_t>(t) * dim + k]) - tst[2 * t]) * tst[2 * t + 1];
    pooled[k] = acc / tokens * gamma[k] + beta[k];
  }
  __syncthreads();
  for (int c = warp; c < classes; c += nw) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc = fmaf(pooled[k], __ldg(Wc + static_cast<size_t>(c) * dim + k), acc);
    acc = warp_sum(acc);
    if (lane == 0) lg[c] = acc + bc[c];
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int c = 0; c < classes; ++c) mx = fmaxf(mx, lg[c]);
  float se = 0.f;
  for (int c = 0; c < classes; ++c) se += expf(lg[c] - mx);
  const float lse = mx + logf(se);
  const long long yl = labels ? static_cast<long long>(labels[b]) : -1;
  const int y = (labels && yl >= 0 && yl < classes) ? static_cast<int>(yl) : -1;
  for (int c = tid; c < classes; c += blockDim.x) logits[static_cast<size_t>(b) * classes + c] = lg[c];
  if (tid == 0 && loss) loss[b] = y >= 0 ? lse - lg[y] : __int_as_float(0x7fc00000);
  if (dh == nullptr) return;
  __syncthreads();
  for (int c = tid; c < classes; c += blockDim.x) lg[c] = expf(lg[c] - lse) - (c == y ? 1.f : 0.f);
  __syncthreads();
  // d pooled-LN-output[k] (same for every token, / tokens), times gamma
  for (int k = tid; k < dim; k += blockDim.x) {
    float acc = 0.f;
    for (int c = 0; c < classes; ++c) acc = fmaf(lg[c], __ldg(Wc + static_cast<size_t>(c) * dim + k), acc);
    dpool[k] = acc * grad_scale / tokens * gamma[k];
  }
  __syncthreads();
  // LayerNorm backward per token (warp per token): dx = rstd * (g - mean(g) - xhat * mean(g * xhat)), g = dpool
  for (int t = warp; t < tokens; t += nw) {
    const float mean = tst[2 * t], rstd = tst[2 * t + 1];
    float s1 = 0.f, s2 = 0.f;
    for (int k = lane; k < dim; k += 32) {
      const float xh = (__bfloat162float(base[static_cast<size_t>(t) * dim + k]) - mean) * rstd;
      s1 += dpool[k];
      s2 = fmaf(dpool[k], xh, s2);
    }
    const float m1 = warp_sum(s1) / dim, m2 = warp_sum(s2) / dim;
    bf16* drow = dh + (static_cast<size_t>(b) * tokens + t) * dim;
    for (int k = lane; k < dim; k += 32) {
      const float xh = (__bfloat162float(base[static_cast<size_t>(t) * dim + k]) - mean) * rstd;
      drow[k] = __float2bfloat16(rstd * (dpool[k] - m1 - xh * m2));
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Pixel kernels for the 4x4 patch embedding: im2col row = b * 3136 + (y / 4) * 56 + x / 4, column = c * 16 + (y % 4) * 4 + x % 4
// (flattening of the conv weight [128, 3, 4, 4]); the matrix is 64 columns wide (48 used, the rest stay zero).
// ------------------------------------------------------------------------------------------------
constexpr int SW_P = 4, SW_G = IMG / SW_P, SW_K = 64;
__device__ __forceinline__ size_t swin_cols_off(int b, int c, int y, int x) {
  return (static_cast<size_t>(b) * SW_G * SW_G + (y >> 2) * SW_G + (x >> 2)) * SW_K + c * 16 + (y & 3) * 4 + (x & 3);
}
__device__ __forceinline__ float u01_hash(uint64_t seed, uint64_t ctr) {
  uint64_t z = seed + 0x9E3779B97F4A7C15ull * (ctr + 1);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return static_cast<float>(z >> 40) * (1.0f / 16777216.0f);
}
__device__ __forceinline__ float enforce_ball(float adv, float x0, float eps) {
  const float d = adv - x0;
  if (d > eps) adv = __uint_as_float(__float_as_uint(adv) - 1u);
  else if (d < -eps) adv = __uint_as_float(__float_as_uint(adv) + 1u);
  return adv;
}
// mode 0: init (adv = clamp(x0 + noise | rng, 0, 1)); 1: PGD update from dcols; 2: materialise the gradient image.
// One thread = 4 consecutive pixels (one patch row).
__global__ void __launch_bounds__(256) swin_pixel_kernel(int mode, const float* __restrict__ x0, const float* __restrict__ noise,
                                                         float* __restrict__ adv, bf16* __restrict__ cols,
                                                         const bf16* __restrict__ dcols, float* __restrict__ grad, int batch,
                                                         PixelNorm nrm, float eps, float alpha, int use_rng, uint64_t seed,
                                                         uint64_t index0, float gscale) {
  const long long total = static_cast<long long>(batch) * 3 * IMG * (IMG / 4);
  const long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int x4 = static_cast<int>(idx % (IMG / 4));
  long long t = idx / (IMG / 4);
  const int y = static_cast<int>(t % IMG);
  t /= IMG;
  const int c = static_cast<int>(t % 3), b = static_cast<int>(t / 3);
  const size_t p = (static_cast<size_t>(b * 3 + c) * IMG + y) * IMG + x4 * 4;
  const size_t co = swin_cols_off(b, c, y, x4 * 4);
  if (mode == 2) {
    const uint2 q = *reinterpret_cast<const uint2*>(dcols + co);
    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    const float2 a = __bfloat1622float2(q2[0]), bb = __bfloat1622float2(q2[1]);
    const float k = nrm.inv_std[c] * gscale;
    *reinterpret_cast<float4*>(grad + p) = make_float4(a.x * k, a.y * k, bb.x * k, bb.y * k);
    return;
  }
  const float4 o4 = __ldg(reinterpret_cast<const float4*>(x0 + p));
  const float xo[4] = {o4.x, o4.y, o4.z, o4.w};
  float v[4];
  if (mode == 0) {
    for (int j = 0; j < 4; ++j) v[j] = xo[j];
    if (noise != nullptr) {
      const float4 n4 = __ldg(reinterpret_cast<const float4*>(noise + p));
      const float nn[4] = {n4.x, n4.y, n4.z, n4.w};
      for (int j = 0; j < 4; ++j) v[j] = enforce_ball(fminf(fmaxf(xo[j] + nn[j], 0.f), 1.f), xo[j], eps);
    } else if (use_rng) {
      const uint64_t e0 = (index0 + b) * (3ull * IMG * IMG) + (static_cast<uint64_t>(c) * IMG + y) * IMG + x4 * 4;
      for (int j = 0; j < 4; ++j) {
        const float n = (2.f * u01_hash(seed, e0 + j) - 1.f) * eps;
        v[j] = enforce_ball(fminf(fmaxf(xo[j] + n, 0.f), 1.f), xo[j], eps);
      }
    }
  } else {
    const float4 a4 = *reinterpret_cast<const float4*>(adv + p);
    const float av[4] = {a4.x, a4.y, a4.z, a4.w};
    const uint2 q = *reinterpret_cast<const uint2*>(dcols + co);
    const __nv_bfloat162* q2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    const float2 g0 = __bfloat1622float2(q2[0]), g1 = __bfloat1622float2(q2[1]);
    const float gg[4] = {g0.x, g0.y, g1.x, g1.y};
    for (int j = 0; j < 4; ++j) {
      const float sg = (gg[j] > 0.f) ? 1.f : ((gg[j] < 0.f) ? -1.f : 0.f);
      const float stepped = av[j] + alpha * sg;
      const float delta = fminf(fmaxf(stepped - xo[j], -eps), eps);
      v[j] = enforce_ball(fminf(fmaxf(xo[j] + delta, 0.f), 1.f), xo[j], eps);
    }
  }
  *reinterpret_cast<float4*>(adv + p) = make_float4(v[0], v[1], v[2], v[3]);
  uint2 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
  o2[0] = __floats2bfloat162_rn((v[0] - nrm.mean[c]) * nrm.inv_std[c], (v[1] - nrm.mean[c]) * nrm.inv_std[c]);
  o2[1] = __floats2bfloat162_rn((v[2] - nrm.mean[c]) * nrm.inv_std[c], (v[3] - nrm.mean[c]) * nrm.inv_std[c]);
  *reinterpret_cast<uint2*>(cols + co) = o;
}

struct BlockW {
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  const bf16 *qkv_w = nullptr, *qkv_wt = nullptr, *proj_w = nullptr, *proj_wt = nullptr, *fc1_w = nullptr, *fc1_wt = nullptr,
             *fc2_w = nullptr, *fc2_wt = nullptr;
  const float *qkv_b = nullptr, *proj_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr, *relbias = nullptr;
  float* bias_frag = nullptr;  // engine-owned fragment-order copy of relbias (+ region masks), relbias_frag()
  struct Lora {
    int rank = 0;
    const bf16 *la_fwd = nullptr, *lb_fwd = nullptr, *lb_bwd = nullptr, *la_bwd = nullptr;
  } lora[4];
  // saved activations
  bf16 *h_in = nullptr, *qkv = nullptr, *ao = nullptr, *h_mid = nullptr, *u = nullptr;
  float2 *st1 = nullptr, *st2 = nullptr;
};
struct BlockPlans {
  GemmPlan t_qkv, qkv, t_proj, proj, t_fc1, fc1, t_fc2, fc2;
  GemmPlan bt_fc2, bfc2, bt_fc1, bfc1, bt_proj, bproj, bt_qkv, bqkv;
};
struct StageW {
  const float *mln_g = nullptr, *mln_b = nullptr;
  const bf16 *m_w = nullptr, *m_wt = nullptr;
  float2* stm = nullptr;
  GemmPlan red, bred;
};
struct SwinPlans {
  GemmPlan patch, bpatch;
  std::vector<std::vector<BlockPlans>> blocks;
  std::vector<GemmPlan> red, bred;
};

}  // namespace
}  // namespace vitatk

using namespace vitatk;

struct vitatk_swin {
  vitatk_swin_config cfg;
  int num_sms = 148;
  bool finalized = false;
  const bf16 *patch_w = nullptr, *patch_wt = nullptr;
  const float *patch_b = nullptr, *eln_g = nullptr, *eln_b = nullptr, *fln_g = nullptr, *fln_b = nullptr, *head_w = nullptr,
              *head_b = nullptr;
  std::vector<std::vector<BlockW>> blk;  // [stage][block]
  StageW stg[3];
  char* ws = nullptr;
  long long ws_bytes = 0;
  char* bias_tab = nullptr;  // fragment-order relative-position bias tables of every block
  // scratch sized for the largest stage
  bf16 *cols = nullptr, *e0 = nullptr, *xn = nullptr, *g = nullptr, *T = nullptr, *dA = nullptr, *dB = nullptr, *dxn = nullptr,
       *dao = nullptr, *du = nullptr, *dqkv = nullptr, *xm = nullptr;
  bf16* sout[4] = {nullptr, nullptr, nullptr, nullptr};  // output of each stage (input of its patch merging / of the head)
  float2* st_e = nullptr;
  float *logits = nullptr, *loss = nullptr, *scratch_img = nullptr;
  std::map<int, SwinPlans*> plans;
  long long launches = 0;
  PixelNorm nrm;
};

namespace vitatk {
namespace {

constexpr int LORA_PAD = 64;
int ksteps_of(int r) { return (r + 15) / 16; }

int swin_dims(const vitatk_swin* e, int s, int* C, int* R) {
  *C = e->cfg.embed_dim << s;
  *R = (e->cfg.image_size / e->cfg.patch_size) >> s;
  return 0;
}
// cyclic shift of block bi of stage s (modeling_swin.py:592-600: odd blocks, unless the stage is a single window)
int swin_shift(const vitatk_swin* e, int s, int bi) {
  int C, R;
  swin_dims(e, s, &C, &R);
  return (bi % 2 == 1 && R > e->cfg.window) ? e->cfg.window / 2 : 0;
}

int build_swin_plans(vitatk_swin* e, int batch, SwinPlans** out) {
  auto it = e->plans.find(batch);
  if (it != e->plans.end()) {
    *out = it->second;
    return 0;
  }
  std::unique_ptr<SwinPlans> owner(new SwinPlans());
  SwinPlans* ps = owner.get();
  const vitatk_swin_config& c = e->cfg;
  GemmEpilogue plain = {};
  plain.mode = EPI_PLAIN;
  const int R0 = c.image_size / c.patch_size, C0 = c.embed_dim;
  const int M0 = batch * R0 * R0;
  {
    GemmEpilogue ep = plain;
    ep.bias = e->patch_b;
    if (gemm_plan_init(&ps->patch, M0, C0, SW_K, e->cols, SW_K, e->patch_w, SW_K, e->e0, C0, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0,
                       0, ep))
      return 1;
    if (gemm_plan_init(&ps->bpatch, M0, SW_K, C0, e->dxn, C0, e->patch_wt, C0, e->dqkv, SW_K, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0,
                       0, plain))
      return 1;
  }
  ps->blocks.resize(4);
  ps->red.resize(3);
  ps->bred.resize(3);
  for (int s = 0; s < 4; ++s) {
    int C, R;
    swin_dims(e, s, &C, &R);
    const int M = batch * R * R, F = 4 * C;
    ps->blocks[s].resize(c.depths[s]);
    for (int bi = 0; bi < c.depths[s]; ++bi) {
      BlockW& w = e->blk[s][bi];
      BlockPlans& p = ps->blocks[s][bi];
      const BlockW::Lora& sq = w.lora[VITATK_SITE_QKV];
      const BlockW::Lora& sp = w.lora[VITATK_SITE_PROJ];
      const BlockW::Lora& s1 = w.lora[VITATK_SITE_FC1];
      const BlockW::Lora& s2 = w.lora[VITATK_SITE_FC2];
      bf16* h_out = (bi + 1 < c.depths[s]) ? e->blk[s][bi + 1].h_in : e->sout[s];  // last block of a stage -> stage output
      auto skinny = [&](GemmPlan* pl, int K, const bf16* A, const bf16* B) {
        return gemm_plan_init(pl, M, LORA_PAD, K, A, K, B, K, e->T, LORA_PAD, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0, 0, plain);
      };
      auto lora_args = [&](const BlockW::Lora& l, const bf16* LB, int* nkb, int* ks) {
        *nkb = l.rank > 0 ? 1 : 0;
        *ks = ksteps_of(l.rank);
        return l.rank > 0 ? LB : nullptr;
      };
      int nkb, ks;
      // ---- forward ----
      if (sq.rank > 0 && skinny(&p.t_qkv, C, e->xn, sq.la_fwd)) return 1;
      {
        GemmEpilogue ep = plain;
        ep.bias = w.qkv_b;
        const bf16* LB = lora_args(sq, sq.lb_fwd, &nkb, &ks);
        if (gemm_plan_init(&p.qkv, M, 3 * C, C, e->xn, C, w.qkv_w, C, w.qkv, 3 * C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks,
                           0, ep))
          return 1;
      }
      if (sp.rank > 0 && skinny(&p.t_proj, C, w.ao, sp.la_fwd)) return 1;
      {
        GemmEpilogue ep = {EPI_RESIDUAL, w.proj_b, w.h_in, C, nullptr, 0};
        const bf16* LB = lora_args(sp, sp.lb_fwd, &nkb, &ks);
        if (gemm_plan_init(&p.proj, M, C, C, w.ao, C, w.proj_w, C, w.h_mid, C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0, ep))
          return 1;
      }
      if (s1.rank > 0 && skinny(&p.t_fc1, C, e->xn, s1.la_fwd)) return 1;
      {
        GemmEpilogue ep = plain;
        ep.mode = EPI_GELU_DUAL;
        ep.bias = w.fc1_b;
        const bf16* LB = lora_args(s1, s1.lb_fwd, &nkb, &ks);
        if (gemm_plan_init(&p.fc1, M, F, C, e->xn, C, w.fc1_w, C, e->g, F, w.u, F, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0, ep)) return 1;
      }
      if (s2.rank > 0 && skinny(&p.t_fc2, F, e->g, s2.la_fwd)) return 1;
      {
        GemmEpilogue ep = {EPI_RESIDUAL, w.fc2_b, w.h_mid, C, nullptr, 0};
        const bf16* LB = lora_args(s2, s2.lb_fwd, &nkb, &ks);
        if (gemm_plan_init(&p.fc2, M, C, F, e->g, F, w.fc2_w, F, h_out, C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0, ep))
          return 1;
      }
      // ---- backward (gradient wrt the block output arrives in dA; dB holds the mid-block gradient) ----
      if (s2.rank > 0 && skinny(&p.bt_fc2, C, e->dA, s2.lb_bwd)) return 1;
      {
        GemmEpilogue ep = {EPI_MUL, nullptr, w.u, F, nullptr, 0};
        const bf16* LB = lora_args(s2, s2.la_bwd, &nkb, &ks);
        if (gemm_plan_init(&p.bfc2, M, F, C, e->dA, C, w.fc2_wt, C, e->du, F, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0, ep))
          return 1;
      }
      if (s1.rank > 0 && skinny(&p.bt_fc1, F, e->du, s1.lb_bwd)) return 1;
      {
        const bf16* LB = lora_args(s1, s1.la_bwd, &nkb, &ks);
        if (gemm_plan_init(&p.bfc1, M, C, F, e->du, F, w.fc1_wt, F, e->dxn, C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0,
                           plain))
          return 1;
      }
      if (sp.rank > 0 && skinny(&p.bt_proj, C, e->dB, sp.lb_bwd)) return 1;
      {
        const bf16* LB = lora_args(sp, sp.la_bwd, &nkb, &ks);
        if (gemm_plan_init(&p.bproj, M, C, C, e->dB, C, w.proj_wt, C, e->dao, C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb, ks, 0,
                           plain))
          return 1;
      }
      if (sq.rank > 0 && skinny(&p.bt_qkv, 3 * C, e->dqkv, sq.lb_bwd)) return 1;
      {
        const bf16* LB = lora_args(sq, sq.la_bwd, &nkb, &ks);
        if (gemm_plan_init(&p.bqkv, M, C, 3 * C, e->dqkv, 3 * C, w.qkv_wt, 3 * C, e->dxn, C, nullptr, 0, e->T, LORA_PAD, LB, LORA_PAD, nkb,
                           ks, 0, plain))
          return 1;
      }
    }
    if (s < 3) {  // patch merging: [M/4, 4C] -> LN -> [M/4, 2C]
      const int M4 = M / 4;
      if (gemm_plan_init(&ps->red[s], M4, 2 * C, 4 * C, e->xn, 4 * C, e->stg[s].m_w, 4 * C, e->blk[s + 1][0].h_in, 2 * C, nullptr, 0,
                         nullptr, 0, nullptr, 0, 0, 0, 0, plain))
        return 1;
      if (gemm_plan_init(&ps->bred[s], M4, 4 * C, 2 * C, e->dA, 2 * C, e->stg[s].m_wt, 2 * C, e->dxn, 4 * C, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, plain))
        return 1;
    }
  }
  e->plans[batch] = owner.release();
  *out = ps;
  return 0;
}

#define SRUN(expr)         \
  do {                     \
    if (expr) return 1;    \
    ++e->launches;         \
  } while (0)
#define SGEMM(plan) SRUN(gemm_launch((plan), s, e->num_sms))

// forward from e->cols (normalised im2col of the input) to the last stage's output in e->hfin
int swin_forward(vitatk_swin* e, SwinPlans* ps, int batch, cudaStream_t s) {
  PdlScope pdl(true);  // short kernels: overlap each launch ramp with the previous kernel (vitatk_internal.h)
  const vitatk_swin_config& c = e->cfg;
  const int R0 = c.image_size / c.patch_size, C0 = c.embed_dim;
  SGEMM(&ps->patch);
  SRUN(ln_any_fwd(e->e0, e->eln_g, e->eln_b, e->blk[0][0].h_in, e->st_e, batch * R0 * R0, C0, c.ln_eps, s));
  for (int st = 0; st < 4; ++st) {
    int C, R;
    swin_dims(e, st, &C, &R);
    const int M = batch * R * R, heads = c.heads[st];
    for (int bi = 0; bi < c.depths[st]; ++bi) {
      BlockW& w = e->blk[st][bi];
      BlockPlans& p = ps->blocks[st][bi];
      const int shift = swin_shift(e, st, bi);
      SRUN(ln_any_fwd(w.h_in, w.ln1_g, w.ln1_b, e->xn, w.st1, M, C, c.ln_eps, s));
      if (w.lora[VITATK_SITE_QKV].rank > 0) SGEMM(&p.t_qkv);
      SGEMM(&p.qkv);
      SRUN(win_attn_fwd(w.qkv, w.bias_frag, w.ao, batch, R, C, heads, shift, s));
      if (w.lora[VITATK_SITE_PROJ].rank > 0) SGEMM(&p.t_proj);
      SGEMM(&p.proj);
      SRUN(ln_any_fwd(w.h_mid, w.ln2_g, w.ln2_b, e->xn, w.st2, M, C, c.ln_eps, s));
      if (w.lora[VITATK_SITE_FC1].rank > 0) SGEMM(&p.t_fc1);
      SGEMM(&p.fc1);
      if (w.lora[VITATK_SITE_FC2].rank > 0) SGEMM(&p.t_fc2);
      SGEMM(&p.fc2);
    }
    if (st < 3) {
      SRUN(merge_permute(e->sout[st], e->xm, batch, R, C, 0, s));
      SRUN(ln_any_fwd(e->xm, e->stg[st].mln_g, e->stg[st].mln_b, e->xn, e->stg[st].stm, M / 4, 4 * C, c.ln_eps, s));
      SGEMM(&ps->red[st]);
    }
  }
  return 0;
}

// backward from dA = d loss / d (last stage output) down to e->dqkv = d loss / d cols
int swin_backward(vitatk_swin* e, SwinPlans* ps, int batch, cudaStream_t s) {
  PdlScope pdl(true);  // short kernels: overlap each launch ramp with the previous kernel (vitatk_internal.h)
  const vitatk_swin_config& c = e->cfg;
  const int R0 = c.image_size / c.patch_size, C0 = c.embed_dim;
  for (int st = 3; st >= 0; --st) {
    int C, R;
    swin_dims(e, st, &C, &R);
    const int M = batch * R * R, heads = c.heads[st];
    if (st < 3) {
      // dA holds the gradient wrt the NEXT stage's input [M/4, 2C]: back through reduction, LayerNorm and the 2x2 gather
      SGEMM(&ps->bred[st]);                                                          // dxn [M/4, 4C]
      SRUN(merge_permute(e->sout[st], e->xm, batch, R, C, 0, s));                    // the LayerNorm's input, re-gathered
      SRUN(ln_any_bwd(e->dxn, e->xm, e->stg[st].stm, e->stg[st].mln_g, nullptr, e->du, M / 4, 4 * C, s));
      SRUN(merge_permute(e->du, e->dA, batch, R, C, 1, s));                          // dA [M, C]
    }
    for (int bi = c.depths[st] - 1; bi >= 0; --bi) {
      BlockW& w = e->blk[st][bi];
      BlockPlans& p = ps->blocks[st][bi];
      const int shift = swin_shift(e, st, bi);
      if (w.lora[VITATK_SITE_FC2].rank > 0) SGEMM(&p.bt_fc2);
      SGEMM(&p.bfc2);
      if (w.lora[VITATK_SITE_FC1].rank > 0) SGEMM(&p.bt_fc1);
      SGEMM(&p.bfc1);
      SRUN(ln_any_bwd(e->dxn, w.h_mid, w.st2, w.ln2_g, e->dA, e->dB, M, C, s));
      if (w.lora[VITATK_SITE_PROJ].rank > 0) SGEMM(&p.bt_proj);
      SGEMM(&p.bproj);
      SRUN(win_attn_bwd(w.qkv, e->dao, w.bias_frag, e->dqkv, batch, R, C, heads, shift, s));
      if (w.lora[VITATK_SITE_QKV].rank > 0) SGEMM(&p.bt_qkv);
      SGEMM(&p.bqkv);
      SRUN(ln_any_bwd(e->dxn, w.h_in, w.st1, w.ln1_g, e->dB, e->dA, M, C, s));
    }
  }
  SRUN(ln_any_bwd(e->dA, e->e0, e->st_e, e->eln_g, nullptr, e->dxn, batch * R0 * R0, C0, s));
  SGEMM(&ps->bpatch);  // e->dqkv <- d loss / d cols  [M0, 64]
  return 0;
}

int swin_pixels(vitatk_swin* e, int mode, const float* x0, const float* noise, float* adv, float* grad, int batch, float eps,
                float alpha, int use_rng, uint64_t seed, uint64_t index0, float gscale, cudaStream_t s) {
  const long long total = static_cast<long long>(batch) * 3 * IMG * (IMG / 4);
  swin_pixel_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(mode, x0, noise, adv, e->cols, e->dqkv, grad, batch,
                                                                               e->nrm, eps, alpha, use_rng, seed, index0, gscale);
  VITATK_CUDA_OK(cudaGetLastError());
  ++e->launches;
  return 0;
}

int swin_head(vitatk_swin* e, const int64_t* labels, float* logits, float* loss, bf16* dh, int batch, float gscale, cudaStream_t s) {
  const vitatk_swin_config& c = e->cfg;
  const int C = c.embed_dim << 3, R = (c.image_size / c.patch_size) >> 3;
  const int tokens = R * R;
  const size_t smem = (2 * C + c.num_classes + 2 * tokens) * sizeof(float);
  swin_head_kernel<<<batch, 256, smem, s>>>(e->sout[3], e->fln_g, e->fln_b, e->head_w, e->head_b, labels, logits, loss, dh, tokens, C,
                                            c.num_classes, c.ln_eps, gscale);
  VITATK_CUDA_OK(cudaGetLastError());
  ++e->launches;
  return 0;
}

int swin_check(vitatk_swin* e, int batch) {
  if (!e || !e->finalized) {
    set_error("swin engine not finalized");
    return 1;
  }
  if (batch < 1 || batch > e->cfg.max_batch) {
    set_error("batch %d outside [1, max_batch=%d]", batch, e->cfg.max_batch);
    return 1;
  }
  return 0;
}

}  // namespace
}  // namespace vitatk

extern "C" {

int vitatk_swin_create(const vitatk_swin_config* cfg, vitatk_swin** out) {
  if (!cfg || !out) {
    set_error("vitatk_swin_create: null argument");
    return 1;
  }
  bool ok = cfg->image_size == 224 && cfg->patch_size == 4 && cfg->window == 7 && cfg->embed_dim % 64 == 0 && cfg->num_classes >= 1 &&
            cfg->max_batch >= 1;
  for (int s = 0; s < 4 && ok; ++s) ok = cfg->depths[s] >= 1 && cfg->heads[s] * HD == (cfg->embed_dim << s);
  if (!ok) {
    set_error("vitatk_swin_create: unsupported geometry (image 224, patch 4, window 7, head dim 32, embed_dim %% 64 == 0)");
    return 1;
  }
  int dev = 0;
  VITATK_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VITATK_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("vitatk requires an sm_100a (B200) device; found sm_%d%d. There is no fallback path.", prop.major, prop.minor);
    return 1;
  }
  vitatk_swin* e = new vitatk_swin();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  e->blk.resize(4);
  for (int s = 0; s < 4; ++s) e->blk[s].resize(cfg->depths[s]);
  for (int i = 0; i < 3; ++i) {
    e->nrm.mean[i] = cfg->mean[i];
    e->nrm.inv_std[i] = 1.0f / cfg->std[i];
  }
  *out = e;
  return 0;
}

int vitatk_swin_destroy(vitatk_swin* e) {
  if (!e) return 0;
  for (auto& kv : e->plans) delete kv.second;
  if (e->ws) cudaFree(e->ws);
  if (e->bias_tab) cudaFree(e->bias_tab);
  delete e;
  return 0;
}

int vitatk_swin_set_tensor(vitatk_swin* e, int id, int stage, int block, const void* p, long long nbytes) {
  if (!e || !p) {
    set_error("vitatk_swin_set_tensor: null argument");
    return 1;
  }
  const vitatk_swin_config& c = e->cfg;
  const float* pf = static_cast<const float*>(p);
  const bf16* pb = static_cast<const bf16*>(p);
  long long want = -1;
  const long long C0 = c.embed_dim, CL = C0 << 3;
  if (id < 16) {
    switch (id) {
      case VITATK_SWIN_PATCH_W: e->patch_w = pb; want = C0 * SW_K * 2; break;
      case VITATK_SWIN_PATCH_WT: e->patch_wt = pb; want = C0 * SW_K * 2; break;
      case VITATK_SWIN_PATCH_B: e->patch_b = pf; want = C0 * 4; break;
      case VITATK_SWIN_EMB_LN_G: e->eln_g = pf; want = C0 * 4; break;
      case VITATK_SWIN_EMB_LN_B: e->eln_b = pf; want = C0 * 4; break;
      case VITATK_SWIN_FINAL_LN_G: e->fln_g = pf; want = CL * 4; break;
      case VITATK_SWIN_FINAL_LN_B: e->fln_b = pf; want = CL * 4; break;
      case VITATK_SWIN_HEAD_W: e->head_w = pf; want = c.num_classes * CL * 4; break;
      case VITATK_SWIN_HEAD_B: e->head_b = pf; want = c.num_classes * 4; break;
      default: break;
    }
  } else if (id >= 64) {
    if (stage < 0 || stage > 2) {
      set_error("vitatk_swin_set_tensor: merge stage %d out of range", stage);
      return 1;
    }
    const long long C = C0 << stage;
    StageW& w = e->stg[stage];
    switch (id) {
      case VITATK_SWIN_MERGE_LN_G: w.mln_g = pf; want = 4 * C * 4; break;
      case VITATK_SWIN_MERGE_LN_B: w.mln_b = pf; want = 4 * C * 4; break;
      case VITATK_SWIN_MERGE_W: w.m_w = pb; want = 2 * C * 4 * C * 2; break;
      case VITATK_SWIN_MERGE_WT: w.m_wt = pb; want = 2 * C * 4 * C * 2; break;
      default: break;
    }
  } else {
    if (stage < 0 || stage > 3 || block < 0 || block >= c.depths[stage]) {
      set_error("vitatk_swin_set_tensor: stage %d block %d out of range", stage, block);
      return 1;
    }
    const long long C = C0 << stage, F = 4 * C;
    BlockW& w = e->blk[stage][block];
    switch (id) {
      case VITATK_SWIN_LN1_G: w.ln1_g = pf; want = C * 4; break;
      case VITATK_SWIN_LN1_B: w.ln1_b = pf; want = C * 4; break;
      case VITATK_SWIN_QKV_W: w.qkv_w = pb; want = 3 * C * C * 2; break;
      case VITATK_SWIN_QKV_WT: w.qkv_wt = pb; want = 3 * C * C * 2; break;
      case VITATK_SWIN_QKV_B: w.qkv_b = pf; want = 3 * C * 4; break;
      case VITATK_SWIN_RELBIAS:
        w.relbias = pf;
        want = static_cast<long long>(c.heads[stage]) * WT * WT * 4;
        if (e->finalized && nbytes == want) {  // re-derive the engine-owned fragment-order copy
          if (relbias_frag(pf, w.bias_frag, c.heads[stage], swin_shift(e, stage, block), nullptr)) return 1;
          VITATK_CUDA_OK(cudaStreamSynchronize(nullptr));
        }
        break;
      case VITATK_SWIN_PROJ_W: w.proj_w = pb; want = C * C * 2; break;
      case VITATK_SWIN_PROJ_WT: w.proj_wt = pb; want = C * C * 2; break;
      case VITATK_SWIN_PROJ_B: w.proj_b = pf; want = C * 4; break;
      case VITATK_SWIN_LN2_G: w.ln2_g = pf; want = C * 4; break;
      case VITATK_SWIN_LN2_B: w.ln2_b = pf; want = C * 4; break;
      case VITATK_SWIN_FC1_W: w.fc1_w = pb; want = F * C * 2; break;
      case VITATK_SWIN_FC1_WT: w.fc1_wt = pb; want = F * C * 2; break;
      case VITATK_SWIN_FC1_B: w.fc1_b = pf; want = F * 4; break;
      case VITATK_SWIN_FC2_W: w.fc2_w = pb; want = F * C * 2; break;
      case VITATK_SWIN_FC2_WT: w.fc2_wt = pb; want = F * C * 2; break;
      case VITATK_SWIN_FC2_B: w.fc2_b = pf; want = C * 4; break;
      default: break;
    }
  }
  if (want < 0) {
    set_error("vitatk_swin_set_tensor: unknown tensor id %d", id);
    return 1;
  }
  if (nbytes != want) {
    set_error("vitatk_swin_set_tensor: id %d stage %d block %d expects %lld bytes, got %lld", id, stage, block, want, nbytes);
    return 1;
  }
  for (auto& kv : e->plans) delete kv.second;
  e->plans.clear();
  return 0;
}

int vitatk_swin_set_lora(vitatk_swin* e, int stage, int block, int site, int rank, const void* la_fwd, const void* lb_fwd,
                         const void* lb_bwd, const void* la_bwd) {
  if (!e || stage < 0 || stage > 3 || block < 0 || block >= e->cfg.depths[stage] || site < 0 || site > 3 || rank < 0 ||
      rank > LORA_PAD || (rank > 0 && (!la_fwd || !lb_fwd || !lb_bwd || !la_bwd))) {
    set_error("vitatk_swin_set_lora: bad arguments");
    return 1;
  }
  BlockW::Lora& l = e->blk[stage][block].lora[site];
  l.rank = rank;
  l.la_fwd = static_cast<const bf16*>(la_fwd);
  l.lb_fwd = static_cast<const bf16*>(lb_fwd);
  l.lb_bwd = static_cast<const bf16*>(lb_bwd);
  l.la_bwd = static_cast<const bf16*>(la_bwd);
  for (auto& kv : e->plans) delete kv.second;
  e->plans.clear();
  return 0;
}

int vitatk_swin_finalize(vitatk_swin* e) {
  if (!e) {
    set_error("vitatk_swin_finalize: null engine");
    return 1;
  }
  if (e->finalized) return 0;
  const vitatk_swin_config& c = e->cfg;
  bool ok = e->patch_w && e->patch_wt && e->patch_b && e->eln_g && e->eln_b && e->fln_g && e->fln_b && e->head_w && e->head_b;
  for (int s = 0; s < 4 && ok; ++s) {
    for (const BlockW& w : e->blk[s])
      ok = ok && w.ln1_g && w.ln1_b && w.ln2_g && w.ln2_b && w.qkv_w && w.qkv_wt && w.qkv_b && w.relbias && w.proj_w && w.proj_wt &&
           w.proj_b && w.fc1_w && w.fc1_wt && w.fc1_b && w.fc2_w && w.fc2_wt && w.fc2_b;
    if (s < 3) ok = ok && e->stg[s].mln_g && e->stg[s].mln_b && e->stg[s].m_w && e->stg[s].m_wt;
  }
  if (!ok) {
    set_error("vitatk_swin_finalize: a tensor is missing");
    return 1;
  }
  auto al = [](long long b) { return (b + 1023) / 1024 * 1024; };
  const long long B = c.max_batch;
  const int R0 = c.image_size / c.patch_size;
  const long long M0 = B * R0 * R0, C0 = c.embed_dim;
  // per-stage token counts and widths; M * C is the same (M0 * C0 / 2^s) ... sizes below use stage 0 as the maximum
  long long total = 0;
  for (int s = 0; s < 4; ++s) {
    const long long M = M0 >> (2 * s), C = C0 << s;
    total += c.depths[s] * (al(M * C * 2) * 3 + al(M * 3 * C * 2) + al(M * 4 * C * 2) + 2 * al(M * 8));
    if (s < 3) total += al((M / 4) * 8);
  }
  const long long szC = al(M0 * C0 * 2), sz3 = al(M0 * 3 * C0 * 2), sz4 = al(M0 * 4 * C0 * 2);
  const long long sz_img = al(B * 3 * IMG * IMG * 4);
  total += al(M0 * SW_K * 2) + szC /*e0*/ + sz4 /*xn (also [M/4, 4C] rows)*/ + sz4 /*g*/ + al(M0 * LORA_PAD * 2) /*T*/ +
           3 * szC /*dA dB dao*/ + 2 * szC /*stage outputs: M C halves per stage*/ + sz4 /*dxn (also [M/4, 4C])*/ + sz4 /*du*/ + sz3 /*dqkv*/ + sz4 /*xm*/ + al(M0 * 8) /*st_e*/ +
           al(B * c.num_classes * 4) + al(B * 4) + sz_img;
  VITATK_CUDA_OK(cudaMalloc(&e->ws, total));
  VITATK_CUDA_OK(cudaMemset(e->ws, 0, total));
  e->ws_bytes = total;
  char* p = e->ws;
  auto take = [&](long long b) {
    char* r = p;
    p += al(b);
    return r;
  };
  for (int s = 0; s < 4; ++s) {
    const long long M = M0 >> (2 * s), C = C0 << s;
    for (BlockW& w : e->blk[s]) {
      w.h_in = reinterpret_cast<bf16*>(take(M * C * 2));
      w.ao = reinterpret_cast<bf16*>(take(M * C * 2));
      w.h_mid = reinterpret_cast<bf16*>(take(M * C * 2));
      w.qkv = reinterpret_cast<bf16*>(take(M * 3 * C * 2));
      w.u = reinterpret_cast<bf16*>(take(M * 4 * C * 2));
      w.st1 = reinterpret_cast<float2*>(take(M * 8));
      w.st2 = reinterpret_cast<float2*>(take(M * 8));
    }
    if (s < 3) e->stg[s].stm = reinterpret_cast<float2*>(take((M / 4) * 8));
  }
  e->cols = reinterpret_cast<bf16*>(take(M0 * SW_K * 2));
  e->e0 = reinterpret_cast<bf16*>(take(M0 * C0 * 2));
  e->xn = reinterpret_cast<bf16*>(take(M0 * 4 * C0 * 2));
  e->g = reinterpret_cast<bf16*>(take(M0 * 4 * C0 * 2));
  e->T = reinterpret_cast<bf16*>(take(M0 * LORA_PAD * 2));
  e->dA = reinterpret_cast<bf16*>(take(M0 * C0 * 2));
  e->dB = reinterpret_cast<bf16*>(take(M0 * C0 * 2));
  e->dao = reinterpret_cast<bf16*>(take(M0 * C0 * 2));
  for (int s = 0; s < 4; ++s) e->sout[s] = reinterpret_cast<bf16*>(take((M0 * C0 * 2) >> s));
  e->dxn = reinterpret_cast<bf16*>(take(M0 * 4 * C0 * 2));
  e->du = reinterpret_cast<bf16*>(take(M0 * 4 * C0 * 2));
  e->dqkv = reinterpret_cast<bf16*>(take(M0 * 3 * C0 * 2));
  e->xm = reinterpret_cast<bf16*>(take(M0 * 4 * C0 * 2));
  e->st_e = reinterpret_cast<float2*>(take(M0 * 8));
  e->logits = reinterpret_cast<float*>(take(B * c.num_classes * 4));
  e->loss = reinterpret_cast<float*>(take(B * 4));
  e->scratch_img = reinterpret_cast<float*>(take(B * 3 * IMG * IMG * 4));
  // relative-position bias + region masks in MMA fragment order (relbias_frag_kernel), one table per block
  size_t tab_bytes = 0;
  for (int s = 0; s < 4; ++s)
    for (int bi = 0; bi < c.depths[s]; ++bi) tab_bytes += relbias_frag_bytes(c.heads[s], swin_shift(e, s, bi));
  VITATK_CUDA_OK(cudaMalloc(&e->bias_tab, tab_bytes));
  char* tp = e->bias_tab;
  for (int s = 0; s < 4; ++s) {
    for (int bi = 0; bi < c.depths[s]; ++bi) {
      BlockW& w = e->blk[s][bi];
      w.bias_frag = reinterpret_cast<float*>(tp);
      tp += relbias_frag_bytes(c.heads[s], swin_shift(e, s, bi));
      if (relbias_frag(w.relbias, w.bias_frag, c.heads[s], swin_shift(e, s, bi), nullptr)) return 1;
    }
  }
  VITATK_CUDA_OK(cudaStreamSynchronize(nullptr));
  e->ws_bytes += static_cast<long long>(tab_bytes);
  e->finalized = true;
  return 0;
}

long long vitatk_swin_workspace_bytes(const vitatk_swin* e) { return e ? e->ws_bytes : 0; }
long long vitatk_swin_launch_count(const vitatk_swin* e) { return e ? e->launches : 0; }

int vitatk_swin_set_normalization(vitatk_swin* e, const float* mean3, const float* std3) {
  if (!e || !mean3 || !std3) {
    set_error("vitatk_swin_set_normalization: null argument");
    return 1;
  }
  for (int i = 0; i < 3; ++i) {
    if (!(std3[i] > 0.f)) {
      set_error("vitatk_swin_set_normalization: std[%d] must be > 0", i);
      return 1;
    }
    e->nrm.mean[i] = mean3[i];
    e->nrm.inv_std[i] = 1.0f / std3[i];
  }
  return 0;
}

int vitatk_swin_forward(vitatk_swin* e, const float* images, int batch, float* logits_out, void* stream) {
  if (swin_check(e, batch)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SwinPlans* ps = nullptr;
  if (build_swin_plans(e, batch, &ps)) return 1;
  if (swin_pixels(e, 0, images, nullptr, e->scratch_img, nullptr, batch, 0.f, 0.f, 0, 0, 0, 0.f, s)) return 1;
  if (swin_forward(e, ps, batch, s)) return 1;
  return swin_head(e, nullptr, logits_out, nullptr, nullptr, batch, 0.f, s);
}

int vitatk_swin_input_grad(vitatk_swin* e, const float* images, const int64_t* labels, int batch, float* grad_out, float* logits_out,
                           float* loss_out, void* stream) {
  if (swin_check(e, batch)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SwinPlans* ps = nullptr;
  if (build_swin_plans(e, batch, &ps)) return 1;
  if (swin_pixels(e, 0, images, nullptr, e->scratch_img, nullptr, batch, 0.f, 0.f, 0, 0, 0, 0.f, s)) return 1;
  if (swin_forward(e, ps, batch, s)) return 1;
  if (swin_head(e, labels, logits_out ? logits_out : e->logits, loss_out ? loss_out : e->loss, e->dA, batch, 1.0f, s)) return 1;
  if (swin_backward(e, ps, batch, s)) return 1;
  return swin_pixels(e, 2, images, nullptr, nullptr, grad_out, batch, 0.f, 0.f, 0, 0, 0, 1.0f / batch, s);
}

int vitatk_swin_attack(vitatk_swin* e, const float* images, const int64_t* labels, int batch, float eps, float alpha, int steps,
                       int start, const float* noise, uint64_t seed, uint64_t image_index0, float* adv, void* stream) {
  if (swin_check(e, batch)) return 1;
  if (steps < 1 || !images || !labels || !adv || images == adv || (start == VITATK_START_NOISE && !noise)) {
    set_error("vitatk_swin_attack: bad arguments");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SwinPlans* ps = nullptr;
  if (build_swin_plans(e, batch, &ps)) return 1;
  if (swin_pixels(e, 0, images, start == VITATK_START_NOISE ? noise : nullptr, adv, nullptr, batch, eps, 0.f,
                  start == VITATK_START_RNG ? 1 : 0, seed, image_index0, 0.f, s))
    return 1;
  for (int it = 0; it < steps; ++it) {
    if (swin_forward(e, ps, batch, s)) return 1;
    if (swin_head(e, labels, e->logits, e->loss, e->dA, batch, 1.0f, s)) return 1;
    if (swin_backward(e, ps, batch, s)) return 1;
    if (swin_pixels(e, 1, images, nullptr, adv, nullptr, batch, eps, alpha, 0, 0, 0, 0.f, s)) return 1;
  }
  return 0;
}

int vitatk_swin_count_correct(vitatk_swin* e, const float* images, const int64_t* labels, int batch, long long* counts, void* stream) {
  if (swin_check(e, batch)) return 1;
  if (!images || !labels || !counts) {
    set_error("vitatk_swin_count_correct: null argument");
    return 1;
  }
  if (vitatk_swin_forward(e, images, batch, e->logits, stream)) return 1;
  ++e->launches;
  return count_correct(e->logits, labels, batch, e->cfg.num_classes, counts, static_cast<cudaStream_t>(stream));
}

long long vitatk_k_win_bias_table(const float* bias, float* table, int heads, int shift, void* stream) {
  if (heads <= 0 || shift < 0 || shift >= WIN) {
    set_error("vitatk_k_win_bias_table: bad arguments");
    return -1;
  }
  if (bias && table && relbias_frag(bias, table, heads, shift, static_cast<cudaStream_t>(stream))) return -1;
  return static_cast<long long>(relbias_frag_bytes(heads, shift));
}
// bias_is_table = 0: bias is the [heads, 49, 49] table (a temporary fragment-order copy is made per call); 1: bias is the
// fragment-order table vitatk_k_win_bias_table wrote for the same (heads, shift)
int vitatk_k_win_attn_fwd(const void* qkv, const float* bias, int bias_is_table, void* out, int batch, int R, int C, int heads, int shift,
                          void* stream) {
  if (!qkv || !bias || !out || batch <= 0 || R % WIN || C != heads * HD || shift < 0 || shift >= WIN) {
    set_error("vitatk_k_win_attn_fwd: bad arguments");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bias_is_table) return win_attn_fwd(static_cast<const bf16*>(qkv), bias, static_cast<bf16*>(out), batch, R, C, heads, shift, s);
  float* tab = nullptr;
  VITATK_CUDA_OK(cudaMallocAsync(&tab, relbias_frag_bytes(heads, shift), s));
  int rc = relbias_frag(bias, tab, heads, shift, s);
  if (!rc) rc = win_attn_fwd(static_cast<const bf16*>(qkv), tab, static_cast<bf16*>(out), batch, R, C, heads, shift, s);
  cudaFreeAsync(tab, s);
  return rc;
}
int vitatk_k_win_attn_bwd(const void* qkv, const void* dout, const float* bias, int bias_is_table, void* dqkv, int batch, int R, int C,
                          int heads, int shift, void* stream) {
  if (!qkv || !dout || !bias || !dqkv || batch <= 0 || R % WIN || C != heads * HD || shift < 0 || shift >= WIN) {
    set_error("vitatk_k_win_attn_bwd: bad arguments");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (bias_is_table)
    return win_attn_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout), bias, static_cast<bf16*>(dqkv), batch, R, C, heads, shift, s);
  float* tab = nullptr;
  VITATK_CUDA_OK(cudaMallocAsync(&tab, relbias_frag_bytes(heads, shift), s));
  int rc = relbias_frag(bias, tab, heads, shift, s);
  if (!rc) rc = win_attn_bwd(static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout), tab, static_cast<bf16*>(dqkv), batch, R, C, heads, shift, s);
  cudaFreeAsync(tab, s);
  return rc;
}

}  // extern "C"

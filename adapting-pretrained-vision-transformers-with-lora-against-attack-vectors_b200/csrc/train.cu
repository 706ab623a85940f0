// LoRA fine-tuning step on the engine's activations: what train_loras.py:295-324 does per batch --
//   peft_model.train(); logits = model(x); loss = CE(logits, y); loss.backward(); Adam.step()   (lr 1e-4, train_loras.py:284)
// with peft's LoRA layer in train mode (train_loras.py:79-95):  y = W x + b + s * B (A dropout_p(x)),  p = 0.1,
// trainable = every lora_A / lora_B plus a full copy of the classifier (task_type SEQ_CLS).
//
// The frozen-weight forward / input-gradient path is the engine's tcgen05 GEMM / attention / LayerNorm kernels; this file
// adds what only training needs.  All of it is HBM-bound streaming work with tiny reductions (rank r <= 64), so the
// kernels are plain coalesced CUDA-core kernels sized to the 148 SMs -- not GEMMs reshaped for the tensor cores:
//   lora_down_kernel   T[m, c0 + j] = sum_k drop(x[m, k]) * A[j, k]                  (dropout mask regenerated, never stored)
//   lora_dx_kernel     dX[m, k] (+)= drop'(m, k) * sum_j BT[m, c0 + j] * s A[j, k]   (LoRA share of the input gradient; the
//                      mask applies to this share only, so it cannot ride in the frozen GEMM's accumulator), optional * mul
//   wgrad_kernel       P[chunk][n][j] = sum_{m in chunk} drop?(X[m, n]) * S[m, c0 + j]   (dB = dY^T T, dA = BT^T drop(x))
//   wgrad_reduce_kernel fixed-order sum over chunks (deterministic, no atomics), scale, optional transpose -> fp32 grads
//   head_wgrad_kernel  classifier dW, db from the head kernel's saved LN output and dlogits
//   adam_kernel        torch.optim.Adam semantics (bias correction, eps outside the sqrt) over the flat parameter buffer
//   lora_repack_kernel fp32 masters -> the bf16 operand layouts of vitatk_set_lora (s*B, B^T, s*A^T, A)
// Dropout masks are counter-based: keep(seed, m * K + k) from a 32-bit integer hash, identical in oracle/train_oracle.py.
#include <cuda_fp16.h>
#include <stdint.h>

#include "vitatk_internal.h"

namespace vitatk {

__host__ __device__ __forceinline__ uint32_t hash32(uint32_t h) {  // "lowbias32" integer finaliser
  h ^= h >> 16;
  h *= 0x7feb352du;
  h ^= h >> 15;
  h *= 0x846ca68bu;
  h ^= h >> 16;
  return h;
}
// keep-probability 1 - p: element idx of the adapter keyed by `seed` survives iff hash >= p * 2^32
__device__ __forceinline__ bool drop_keep(uint32_t seed, uint32_t idx, uint32_t thresh) { return hash32(idx ^ seed) >= thresh; }

static uint32_t drop_threshold(float p) {
  if (!(p > 0.f)) return 0u;
  double t = static_cast<double>(p) * 4294967296.0;
  return t >= 4294967295.0 ? 0xffffffffu : static_cast<uint32_t>(t);
}

// ------------------------------------------------------------------------------------------------
// T[m, c0 + j] = sum_k drop(x[m, k]) * A[j, k], j < r.  One warp per row; lane l owns columns {8 (l + 32 i)} (16-byte loads);
// A (fp32 master, [r, K]) is read through L1 (it is tiny and shared by every row).  R = padded rank handled per pass.
// ------------------------------------------------------------------------------------------------
template <int R>
__global__ void __launch_bounds__(256) lora_down_kernel(const bf16* __restrict__ x, int ldx, int K, const float* __restrict__ A,
                                                        int r, bf16* __restrict__ T, int ldt, int c0, int rows,
                                                        uint32_t seed, uint32_t thresh, float inv_keep, long long row0) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  const bf16* xr = x + static_cast<size_t>(row) * ldx;
  const uint32_t base = static_cast<uint32_t>((row0 + row) * static_cast<long long>(K));
  for (int k0 = lane * 8; k0 < K; k0 += 256) {
    const uint4 q = *reinterpret_cast<const uint4*>(xr + k0);
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
    float xv[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(p2[i]);
      xv[2 * i] = f.x;
      xv[2 * i + 1] = f.y;
    }
    if (thresh != 0u) {
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = drop_keep(seed, base + k0 + i, thresh) ? xv[i] * inv_keep : 0.f;
    }
#pragma unroll
    for (int j = 0; j < R; ++j) {
      if (j < r) {
        const float4 a0 = __ldg(reinterpret_cast<const float4*>(A + static_cast<size_t>(j) * K + k0));
        const float4 a1 = __ldg(reinterpret_cast<const float4*>(A + static_cast<size_t>(j) * K + k0 + 4));
        acc[j] = fmaf(xv[0], a0.x, acc[j]);
        acc[j] = fmaf(xv[1], a0.y, acc[j]);
        acc[j] = fmaf(xv[2], a0.z, acc[j]);
        acc[j] = fmaf(xv[3], a0.w, acc[j]);
        acc[j] = fmaf(xv[4], a1.x, acc[j]);
        acc[j] = fmaf(xv[5], a1.y, acc[j]);
        acc[j] = fmaf(xv[6], a1.z, acc[j]);
        acc[j] = fmaf(xv[7], a1.w, acc[j]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < R; ++j) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], o);
  }
  if (lane == 0) {
    bf16* tr = T + static_cast<size_t>(row) * ldt + c0;
#pragma unroll
    for (int j = 0; j < R; ++j)
      if (j < r) tr[j] = __float2bfloat16(acc[j]);
  }
}

int lora_down(const bf16* x, int ldx, int K, const float* A, int r, bf16* T, int ldt, int c0, int rows, uint32_t seed,
              float p, long long row0, cudaStream_t stream) {
  if (K % 8 != 0 || r < 1 || r > 64) {
    set_error("lora_down: unsupported K=%d r=%d", K, r);
    return 1;
  }
  const uint32_t th = drop_threshold(p);
  const float inv = th ? 1.0f / (1.0f - p) : 1.0f;
  const dim3 grid((rows + 7) / 8), block(256);
  // ranks above 16 run as several 16-wide passes over the row (x stays L1/L2-hot)
  for (int j0 = 0; j0 < r; j0 += 16) {
    const int rr = r - j0 < 16 ? r - j0 : 16;
    if (rr <= 8)
      lora_down_kernel<8><<<grid, block, 0, stream>>>(x, ldx, K, A + static_cast<size_t>(j0) * K, rr, T, ldt, c0 + j0, rows, seed,
                                                      th, inv, row0);
    else
      lora_down_kernel<16><<<grid, block, 0, stream>>>(x, ldx, K, A + static_cast<size_t>(j0) * K, rr, T, ldt, c0 + j0, rows,
                                                       seed, th, inv, row0);
  }
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// dX[m, k] = (accumulate ? dX[m, k] : 0) + sum_a drop'_a(m, k) * sum_j BT[m, c0_a + j] * (s_a A_a[j, k]);  then * mul[m, k].
// Up to three adapters share the input (q, k, v read the same LayerNorm output, each with its own mask).  One thread per
// 8 consecutive columns of one row; the row's BT values are fetched once per thread (<= 3 * 16 floats).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) lora_dx_kernel(bf16* __restrict__ dX, int ldx, int K, const bf16* __restrict__ BT, int ldt,
                                                      LoraDxArgs args, const bf16* __restrict__ mul, int ldm, int rows,
                                                      int accumulate, uint32_t thresh, float inv_keep, long long row0) {
  const int chunks = K >> 3;
  const long long gid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (gid >= static_cast<long long>(rows) * chunks) return;
  const int row = static_cast<int>(gid / chunks), k0 = static_cast<int>(gid % chunks) * 8;
  float out[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) out[i] = 0.f;
  const uint32_t base = static_cast<uint32_t>((row0 + row) * static_cast<long long>(K)) + k0;
  for (int a = 0; a < args.n; ++a) {
    const LoraDxAdapter ad = args.ad[a];
    float part[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) part[i] = 0.f;
    const bf16* bt = BT + static_cast<size_t>(row) * ldt + ad.c0;
    for (int j = 0; j < ad.r; ++j) {
      const float b = __bfloat162float(bt[j]) * ad.scale;
      const float4 a0 = __ldg(reinterpret_cast<const float4*>(ad.A + static_cast<size_t>(j) * K + k0));
      const float4 a1 = __ldg(reinterpret_cast<const float4*>(ad.A + static_cast<size_t>(j) * K + k0 + 4));
      part[0] = fmaf(b, a0.x, part[0]);
      part[1] = fmaf(b, a0.y, part[1]);
      part[2] = fmaf(b, a0.z, part[2]);
      part[3] = fmaf(b, a0.w, part[3]);
      part[4] = fmaf(b, a1.x, part[4]);
      part[5] = fmaf(b, a1.y, part[5]);
      part[6] = fmaf(b, a1.z, part[6]);
      part[7] = fmaf(b, a1.w, part[7]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (thresh == 0u) out[i] += part[i];
      else if (drop_keep(ad.seed, base + i, thresh)) out[i] = fmaf(part[i], inv_keep, out[i]);
    }
  }
  uint4* dst = reinterpret_cast<uint4*>(dX + static_cast<size_t>(row) * ldx + k0);
  if (accumulate) {
    const uint4 q = *dst;
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(p2[i]);
      out[2 * i] += f.x;
      out[2 * i + 1] += f.y;
    }
  }
  if (mul != nullptr) {
    const uint4 q = *reinterpret_cast<const uint4*>(mul + static_cast<size_t>(row) * ldm + k0);
    const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float2 f = __bfloat1622float2(p2[i]);
      out[2 * i] *= f.x;
      out[2 * i + 1] *= f.y;
    }
  }
  uint4 o;
  __nv_bfloat162* o2 = reinterpret_cast<__nv_bfloat162*>(&o);
#pragma unroll
  for (int i = 0; i < 4; ++i) o2[i] = __floats2bfloat162_rn(out[2 * i], out[2 * i + 1]);
  *dst = o;
}

int lora_dx(bf16* dX, int ldx, int K, const bf16* BT, int ldt, const LoraDxArgs& args, const bf16* mul, int ldm, int rows,
            int accumulate, float p, long long row0, cudaStream_t stream) {
  if (K % 8 != 0 || args.n < 1 || args.n > 3) {
    set_error("lora_dx: unsupported K=%d adapters=%d", K, args.n);
    return 1;
  }
  const uint32_t th = drop_threshold(p);
  const float inv = th ? 1.0f / (1.0f - p) : 1.0f;
  const long long total = static_cast<long long>(rows) * (K / 8);
  lora_dx_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(dX, ldx, K, BT, ldt, args, mul, ldm, rows,
                                                                                 accumulate, th, inv, row0);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Weight gradients of a rank-r adapter:  G[n, j] = sum_m drop?(X[m, n]) * S[m, c0 + j]   (reduction over the M token rows).
//   dB[out, r] = s * dY^T T      X = dY [M, out],  S = T = drop(x) A^T
//   dA[r, in]  = s * BT^T drop(x)  X = x  [M, in],   S = BT = dY B          (transposed on the way out)
// Split over row chunks: CTA (slab, chunk) = 128 columns x WG_ROWS rows; thread = one column with r accumulators, the
// chunk's S rows staged in shared memory (broadcast reads), X read as coalesced 256-byte row segments.  Partials go to
// P[chunk][n][16]; wgrad_reduce_kernel sums the chunks in a fixed order (deterministic) and applies the scale.
// ------------------------------------------------------------------------------------------------
static constexpr int WG_ROWS = 256;
static constexpr int WG_COLS = 128;

template <int R>
__global__ void __launch_bounds__(WG_COLS) wgrad_kernel(const bf16* __restrict__ X, int ldx, int N, const bf16* __restrict__ S,
                                                        int lds, int c0, int r, int rows, float* __restrict__ P,
                                                        uint32_t seed, uint32_t thresh, float inv_keep, long long row0,
                                                        int mask_ld, int x_f16) {
  __shared__ float s_sh[WG_ROWS][R];
  const int n = blockIdx.x * WG_COLS + threadIdx.x;
  const int m_begin = blockIdx.y * WG_ROWS;
  const int m_end = min(rows, m_begin + WG_ROWS);
  for (int i = threadIdx.x; i < WG_ROWS * R; i += WG_COLS) {
    const int mm = i / R, j = i % R;
    const int m = m_begin + mm;
    s_sh[mm][j] = (m < m_end && j < r) ? __bfloat162float(S[static_cast<size_t>(m) * lds + c0 + j]) : 0.f;
  }
  __syncthreads();
  float acc[R];
#pragma unroll
  for (int j = 0; j < R; ++j) acc[j] = 0.f;
  if (n < N) {
    const bf16* xp = X + static_cast<size_t>(m_begin) * ldx + n;
    for (int m = m_begin; m < m_end; ++m, xp += ldx) {
      float xv = x_f16 ? __half2float(*reinterpret_cast<const __half*>(xp)) : __bfloat162float(*xp);
      if (thresh != 0u) {
        const uint32_t idx = static_cast<uint32_t>((row0 + m) * static_cast<long long>(mask_ld)) + n;
        xv = drop_keep(seed, idx, thresh) ? xv * inv_keep : 0.f;
      }
      const float* sr = s_sh[m - m_begin];
#pragma unroll
      for (int j = 0; j < R; ++j) acc[j] = fmaf(xv, sr[j], acc[j]);
    }
    float* pp = P + (static_cast<size_t>(blockIdx.y) * N + n) * 16;
#pragma unroll
    for (int j = 0; j < R; ++j) pp[j] = acc[j];
  }
}

// G[n * r + j] (or G[j * N + n] when transpose) = scale * sum_chunks P[chunk][n][j]
__global__ void wgrad_reduce_kernel(const float* __restrict__ P, int chunks, int N, int rr, int r_total, int j0, float scale,
                                    int transpose, float* __restrict__ G) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * rr) return;
  const int n = i / rr, j = i % rr;
  float acc = 0.f;
  for (int c = 0; c < chunks; ++c) acc += P[(static_cast<size_t>(c) * N + n) * 16 + j];
  G[transpose ? static_cast<size_t>(j0 + j) * N + n : static_cast<size_t>(n) * r_total + j0 + j] = acc * scale;
}

int wgrad(const bf16* X, int ldx, int N, const bf16* S, int lds, int c0, int r, int rows, float* partial, float scale,
          int transpose, float* G, uint32_t seed, float p, long long row0, int mask_ld, int x_f16, cudaStream_t stream) {
  if (r < 1 || r > 64) {
    set_error("wgrad: unsupported rank %d", r);
    return 1;
  }
  const uint32_t th = drop_threshold(p);
  const float inv = th ? 1.0f / (1.0f - p) : 1.0f;
  const int chunks = (rows + WG_ROWS - 1) / WG_ROWS;
  const dim3 grid((N + WG_COLS - 1) / WG_COLS, chunks);
  for (int j0 = 0; j0 < r; j0 += 16) {  // ranks above 16: 16 columns of S per pass
    const int rr = r - j0 < 16 ? r - j0 : 16;
    if (rr <= 8)
      wgrad_kernel<8><<<grid, WG_COLS, 0, stream>>>(X, ldx, N, S, lds, c0 + j0, rr, rows, partial, seed, th, inv, row0, mask_ld, x_f16);
    else
      wgrad_kernel<16><<<grid, WG_COLS, 0, stream>>>(X, ldx, N, S, lds, c0 + j0, rr, rows, partial, seed, th, inv, row0, mask_ld, x_f16);
    // the reduce of pass j0 fills rank columns [j0, j0 + rr) of G ([N, r] row-major, or [r, N] when transposed)
    wgrad_reduce_kernel<<<(N * rr + 255) / 256, 256, 0, stream>>>(partial, chunks, N, rr, r, j0, scale, transpose, G);
  }
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// Classifier (peft modules_to_save copy, train_loras.py:84): dW[c, k] = scale * sum_b dlogits[b, c] * y[b, k], db likewise.
// y = final-LayerNorm output of the CLS row and dlogits = softmax - onehot are saved by the head kernel.
// ------------------------------------------------------------------------------------------------
__global__ void head_wgrad_kernel(const float* __restrict__ y, const float* __restrict__ dlogits, int batch, int dim, int classes,
                                  float scale, float* __restrict__ dW, float* __restrict__ db) {
  const int c = blockIdx.x;
  for (int k = threadIdx.x; k < dim; k += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < batch; ++b) acc = fmaf(dlogits[static_cast<size_t>(b) * classes + c], y[static_cast<size_t>(b) * dim + k], acc);
    dW[static_cast<size_t>(c) * dim + k] = acc * scale;
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int b = 0; b < batch; ++b) acc += dlogits[static_cast<size_t>(b) * classes + c];
    db[c] = acc * scale;
  }
}

int head_wgrad(const float* y, const float* dlogits, int batch, int dim, int classes, float scale, float* dW, float* db,
               cudaStream_t stream) {
  head_wgrad_kernel<<<classes, 256, 0, stream>>>(y, dlogits, batch, dim, classes, scale, dW, db);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// torch.optim.Adam (train_loras.py:284; betas 0.9 / 0.999, eps 1e-8, no weight decay, no amsgrad):
//   m = b1 m + (1 - b1) g ; v = b2 v + (1 - b2) g^2 ; p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps)
// ------------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  p[i] -= (lr / bc1) * (mi / denom);
}

int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int step,
              cudaStream_t stream) {
  const float bc1 = 1.f - powf(b1, static_cast<float>(step));
  const float bc2 = sqrtf(1.f - powf(b2, static_cast<float>(step)));
  adam_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, stream>>>(p, g, m, v, n, lr, b1, b2, eps, bc1, bc2);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------------
// fp32 masters -> bf16 operands of one adapter inside its site's packed buffers (layouts of vitatk_set_lora):
//   la_fwd[(row0 + j) * in + k]            = A[j, k]                 (only used by the eval / attack path)
//   lb_fwd[(out0 + n) * 64 + col0 + j]     = s * B[n, j]
//   lb_bwd[(row0 + j) * ld_lbb + out0 + n] = B[n, j]
//   la_bwd[k * ld_lab + row0 + j]          = s * A[j, k]
// ------------------------------------------------------------------------------------------------
__global__ void lora_repack_kernel(const float* __restrict__ A, const float* __restrict__ B, int r, int in, int out, float s,
                                   bf16* __restrict__ la_fwd, bf16* __restrict__ lb_fwd, bf16* __restrict__ lb_bwd,
                                   bf16* __restrict__ la_bwd, int row0, int col0, int out0, int ld_lbb, int ld_lab,
                                   const float* __restrict__ gamma, int fmt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < r * in) {
    const int j = i / in, k = i % in;
    const float a = A[i];
    const float af = gamma ? a * gamma[k] : a;
    if (fmt & 1) reinterpret_cast<__half*>(la_fwd)[static_cast<size_t>(row0 + j) * in + k] = __float2half_rn(af);
    else la_fwd[static_cast<size_t>(row0 + j) * in + k] = __float2bfloat16(af);
    la_bwd[static_cast<size_t>(k) * ld_lab + row0 + j] = __float2bfloat16(s * a);
  }
  if (i < out * r) {
    const int n = i / r, j = i % r;
    const float b = B[i];
    lb_fwd[static_cast<size_t>(out0 + n) * 64 + col0 + j] = __float2bfloat16(s * b);
    if (fmt & 2) reinterpret_cast<__half*>(lb_bwd)[static_cast<size_t>(row0 + j) * ld_lbb + out0 + n] = __float2half_rn(b);
    else lb_bwd[static_cast<size_t>(row0 + j) * ld_lbb + out0 + n] = __float2bfloat16(b);
  }
}

int lora_repack(const float* A, const float* B, int r, int in, int out, float s, bf16* la_fwd, bf16* lb_fwd, bf16* lb_bwd,
                bf16* la_bwd, int row0, int col0, int out0, int ld_lbb, int ld_lab, const float* gamma, int fmt,
                cudaStream_t stream) {
  const int n = r * (in > out ? in : out);
  lora_repack_kernel<<<(n + 255) / 256, 256, 0, stream>>>(A, B, r, in, out, s, la_fwd, lb_fwd, lb_bwd, la_bwd, row0, col0, out0,
                                                          ld_lbb, ld_lab, gamma, fmt);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

uint32_t train_mask_seed(uint64_t seed, uint64_t step, int layer, int adapter) {
  // splitmix64 of (seed, step, layer, adapter) -> 32 bits; restated in oracle/train_oracle.py
  uint64_t z = seed + 0x9e3779b97f4a7c15ull * (step * 1024ull + static_cast<uint64_t>(layer) * 8ull + static_cast<uint64_t>(adapter) + 1ull);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  z = z ^ (z >> 31);
  return static_cast<uint32_t>(z);
}

}  // namespace vitatk

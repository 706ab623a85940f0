// LoRA fine-tuning step on the engine's activations: what train_loras.py:295-324 does per batch --
//   peft_model.train(); logits = model(x); loss = CE(logits, y); loss.backward(); Adam.step()   (lr 1e-4, train_loras.py:284)
// with peft's LoRA layer in train mode (train_loras.py:79-95):  y = W x + b + s * B (A dropout_p(x)),  p = 0.1,
// trainable = every lora_A / lora_B plus a full copy of the classifier (task_type SEQ_CLS).
//
// The frozen-weight forward / input-gradient path is the engine's tcgen05 GEMM / attention / LayerNorm kernels; this file
// adds what only training needs.  All of it is HBM-bound streaming work with tiny reductions (rank r <= 64): coalesced
// streaming kernels sized to the 148 SMs, not GEMM launches.  The rank-r products inside them run as warp-level
// m16n8k16 MMAs (*_mma_kernel below, mma_sync.cuh) -- the first, CUDA-core versions were FMA-issue bound at 8-14x their
// HBM time -- and the CUDA-core kernels remain as the fallback for shapes the MMA versions do not take:
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
#include "mma_sync.cuh"

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

// Tensor-core version (K % 32 == 0): one warp per 16 rows, the whole reduction in mma.sync m16n8k16 with the row in the A
// operand and up to 16 adapter rows in the B operand.  The MMA's k index is a free permutation as long as both operands
// agree, so lane t of a quad owns the 8 CONSECUTIVE columns k0 + 8t .. 8t + 7 of each 32-column block (one 16-byte load
// per row for x, two for the fp32 adapter row) instead of the strided (2t, 2t + 1, 2t + 8, 2t + 9) of the canonical
// fragment.  Dropped elements are zeroed in the packed registers; 1 / (1 - p) is applied to the 16 accumulators.
template <bool DROP>
__global__ void __launch_bounds__(128) lora_down_mma_kernel(const bf16* __restrict__ x, int ldx, int K, const float* __restrict__ A,
                                                            int r, bf16* __restrict__ T, int ldt, int c0, int rows, uint32_t seed,
                                                            uint32_t thresh, float inv_keep, long long row0) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (blockIdx.x * 4 + warp) * 16;
  if (m0 >= rows) return;
  const int r_lo = m0 + g, r_hi = r_lo + 8;
  const bool ok_lo = r_lo < rows, ok_hi = r_hi < rows;
  const bf16* x_lo = x + static_cast<size_t>(ok_lo ? r_lo : m0) * ldx + 8 * t;
  const bf16* x_hi = x + static_cast<size_t>(ok_hi ? r_hi : m0) * ldx + 8 * t;
  const bool a0_ok = g < r, a1_ok = g + 8 < r;
  const float* a0p = A + static_cast<size_t>(a0_ok ? g : 0) * K + 8 * t;
  const float* a1p = A + static_cast<size_t>(a1_ok ? g + 8 : 0) * K + 8 * t;
  const uint32_t base_lo = static_cast<uint32_t>((row0 + r_lo) * static_cast<long long>(K)) + 8 * t;
  const uint32_t base_hi = static_cast<uint32_t>((row0 + r_hi) * static_cast<long long>(K)) + 8 * t;
  float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
  for (int k0 = 0; k0 < K; k0 += 32) {
    uint4 q_lo = __ldg(reinterpret_cast<const uint4*>(x_lo + k0));
    uint4 q_hi = __ldg(reinterpret_cast<const uint4*>(x_hi + k0));
    const float4 f00 = __ldg(reinterpret_cast<const float4*>(a0p + k0)), f01 = __ldg(reinterpret_cast<const float4*>(a0p + k0 + 4));
    const float4 f10 = __ldg(reinterpret_cast<const float4*>(a1p + k0)), f11 = __ldg(reinterpret_cast<const float4*>(a1p + k0 + 4));
    uint32_t xl[4] = {q_lo.x, q_lo.y, q_lo.z, q_lo.w}, xh[4] = {q_hi.x, q_hi.y, q_hi.z, q_hi.w};
    if (DROP) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint32_t kl = (drop_keep(seed, base_lo + k0 + 2 * i, thresh) ? 0x0000ffffu : 0u) |
                            (drop_keep(seed, base_lo + k0 + 2 * i + 1, thresh) ? 0xffff0000u : 0u);
        const uint32_t kh = (drop_keep(seed, base_hi + k0 + 2 * i, thresh) ? 0x0000ffffu : 0u) |
                            (drop_keep(seed, base_hi + k0 + 2 * i + 1, thresh) ? 0xffff0000u : 0u);
        xl[i] &= kl;
        xh[i] &= kh;
      }
    }
    uint32_t b0[4], b1[4];
    b0[0] = a0_ok ? pack_bf16x2(f00.x, f00.y) : 0u;
    b0[1] = a0_ok ? pack_bf16x2(f00.z, f00.w) : 0u;
    b0[2] = a0_ok ? pack_bf16x2(f01.x, f01.y) : 0u;
    b0[3] = a0_ok ? pack_bf16x2(f01.z, f01.w) : 0u;
    b1[0] = a1_ok ? pack_bf16x2(f10.x, f10.y) : 0u;
    b1[1] = a1_ok ? pack_bf16x2(f10.z, f10.w) : 0u;
    b1[2] = a1_ok ? pack_bf16x2(f11.x, f11.y) : 0u;
    b1[3] = a1_ok ? pack_bf16x2(f11.z, f11.w) : 0u;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      const uint32_t af[4] = {xl[2 * s], xh[2 * s], xl[2 * s + 1], xh[2 * s + 1]};
      mma16816(acc0, af, b0[2 * s], b0[2 * s + 1]);
      mma16816(acc1, af, b1[2 * s], b1[2 * s + 1]);
    }
  }
  const float sc = DROP ? inv_keep : 1.0f;
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int j0 = 2 * t + e, j1 = j0 + 8;
    if (ok_lo) {
      bf16* tr = T + static_cast<size_t>(r_lo) * ldt + c0;
      if (j0 < r) tr[j0] = __float2bfloat16(acc0[e] * sc);
      if (j1 < r) tr[j1] = __float2bfloat16(acc1[e] * sc);
    }
    if (ok_hi) {
      bf16* tr = T + static_cast<size_t>(r_hi) * ldt + c0;
      if (j0 < r) tr[j0] = __float2bfloat16(acc0[2 + e] * sc);
      if (j1 < r) tr[j1] = __float2bfloat16(acc1[2 + e] * sc);
    }
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
  const bool mma_ok = K % 32 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(A) & 15) == 0;
  const dim3 grid((rows + 7) / 8), block(256);
  // ranks above 16 run as several 16-wide passes over the row (x stays L1/L2-hot)
  for (int j0 = 0; j0 < r; j0 += 16) {
    const int rr = r - j0 < 16 ? r - j0 : 16;
    const float* Aj = A + static_cast<size_t>(j0) * K;
    if (mma_ok) {
      const dim3 g2((rows + 63) / 64);
      if (th) lora_down_mma_kernel<true><<<g2, 128, 0, stream>>>(x, ldx, K, Aj, rr, T, ldt, c0 + j0, rows, seed, th, inv, row0);
      else lora_down_mma_kernel<false><<<g2, 128, 0, stream>>>(x, ldx, K, Aj, rr, T, ldt, c0 + j0, rows, seed, th, inv, row0);
    } else if (rr <= 8) {
      lora_down_kernel<8><<<grid, block, 0, stream>>>(x, ldx, K, Aj, rr, T, ldt, c0 + j0, rows, seed, th, inv, row0);
    } else {
      lora_down_kernel<16><<<grid, block, 0, stream>>>(x, ldx, K, Aj, rr, T, ldt, c0 + j0, rows, seed, th, inv, row0);
    }
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

// Tensor-core version (K % 32 == 0, ranks <= 16 * KS): one warp per 16 rows x a slice of the columns.  BT's fragments
// (16 x r per adapter) are loaded once; per 32-column block each adapter costs 4 * KS MMAs.  The MMA's n index is a free
// permutation too: output slot (tile i, column 2t + e) stands for column 8t + 2i + e of the block, so every lane ends up
// with 8 CONSECUTIVE columns of its two rows -- one 16-byte read-modify-write per row.
template <int KS, bool DROP>
__global__ void __launch_bounds__(128) lora_dx_mma_kernel(bf16* __restrict__ dX, int ldx, int K, const bf16* __restrict__ BT, int ldt,
                                                          LoraDxArgs args, const bf16* __restrict__ mul, int ldm, int rows,
                                                          int accumulate, uint32_t thresh, float inv_keep, long long row0, int kslice) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int m0 = (blockIdx.x * 4 + warp) * 16;
  if (m0 >= rows) return;
  const int r_lo = m0 + g, r_hi = r_lo + 8;
  const bool ok_lo = r_lo < rows, ok_hi = r_hi < rows;
  const int kb_begin = blockIdx.y * kslice, kb_end = min(K, kb_begin + kslice);
  auto bt2 = [&](int row, bool ok, int c, int rem) -> uint32_t {  // BT[row][c], BT[row][c + 1], zero beyond the rank
    if (!ok || rem <= 0) return 0u;
    const uint32_t v = *reinterpret_cast<const uint32_t*>(BT + static_cast<size_t>(row) * ldt + c);
    return rem >= 2 ? v : (v & 0x0000ffffu);
  };
  uint32_t af[3][KS][4];
#pragma unroll
  for (int a = 0; a < 3; ++a) {
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int j = ks * 16 + 2 * t;
      const bool on = a < args.n;
      const int c0 = on ? args.ad[a].c0 : 0, r = on ? args.ad[a].r : 0;
      af[a][ks][0] = bt2(r_lo, ok_lo, c0 + j, r - j);
      af[a][ks][1] = bt2(r_hi, ok_hi, c0 + j, r - j);
      af[a][ks][2] = bt2(r_lo, ok_lo, c0 + j + 8, r - j - 8);
      af[a][ks][3] = bt2(r_hi, ok_hi, c0 + j + 8, r - j - 8);
    }
  }
  const int bcol = 8 * (g >> 1);  // B fragment: this lane supplies column 8 (g >> 1) + 2 i + (g & 1) of tile i
  const bool odd = g & 1;
  for (int kb = kb_begin; kb < kb_end; kb += 32) {
    float o_lo[8], o_hi[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) o_lo[c] = o_hi[c] = 0.f;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      if (a < args.n) {
        const LoraDxAdapter ad = args.ad[a];
        float part[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int e = 0; e < 4; ++e) part[i][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KS; ++ks) {
          if (ks * 16 < ad.r) {
            float v[4][4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
              const int j = ks * 16 + 2 * t + (jj & 1) + (jj >> 1) * 8;
              float4 u0 = make_float4(0.f, 0.f, 0.f, 0.f), u1 = u0;
              if (j < ad.r) {
                const float* ap = ad.A + static_cast<size_t>(j) * K + kb + bcol;
                u0 = __ldg(reinterpret_cast<const float4*>(ap));
                u1 = __ldg(reinterpret_cast<const float4*>(ap + 4));
              }
              v[jj][0] = odd ? u0.y : u0.x;
              v[jj][1] = odd ? u0.w : u0.z;
              v[jj][2] = odd ? u1.y : u1.x;
              v[jj][3] = odd ? u1.w : u1.z;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) mma16816(part[i], af[a][ks], pack_bf16x2(v[0][i], v[1][i]), pack_bf16x2(v[2][i], v[3][i]));
          }
        }
        const float coef = ad.scale * (DROP ? inv_keep : 1.0f);
        const uint32_t base_lo = static_cast<uint32_t>((row0 + r_lo) * static_cast<long long>(K)) + kb + 8 * t;
        const uint32_t base_hi = static_cast<uint32_t>((row0 + r_hi) * static_cast<long long>(K)) + kb + 8 * t;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const int c = 2 * i + e;
            const bool k_lo = !DROP || drop_keep(ad.seed, base_lo + c, thresh), k_hi = !DROP || drop_keep(ad.seed, base_hi + c, thresh);
            if (k_lo) o_lo[c] = fmaf(part[i][e], coef, o_lo[c]);
            if (k_hi) o_hi[c] = fmaf(part[i][2 + e], coef, o_hi[c]);
          }
        }
      }
    }
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int row = hb ? r_hi : r_lo;
      if (hb ? ok_hi : ok_lo) {
        float* o = hb ? o_hi : o_lo;
        uint4* dst = reinterpret_cast<uint4*>(dX + static_cast<size_t>(row) * ldx + kb + 8 * t);
        if (accumulate) {
          const uint4 q = *dst;
          const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(p2[i]);
            o[2 * i] += f.x;
            o[2 * i + 1] += f.y;
          }
        }
        if (mul != nullptr) {
          const uint4 q = __ldg(reinterpret_cast<const uint4*>(mul + static_cast<size_t>(row) * ldm + kb + 8 * t));
          const __nv_bfloat162* p2 = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(p2[i]);
            o[2 * i] *= f.x;
            o[2 * i + 1] *= f.y;
          }
        }
        uint4 w;
        w.x = pack_bf16x2(o[0], o[1]);
        w.y = pack_bf16x2(o[2], o[3]);
        w.z = pack_bf16x2(o[4], o[5]);
        w.w = pack_bf16x2(o[6], o[7]);
        *dst = w;
      }
    }
  }
}

int lora_dx(bf16* dX, int ldx, int K, const bf16* BT, int ldt, const LoraDxArgs& args, const bf16* mul, int ldm, int rows,
            int accumulate, float p, long long row0, cudaStream_t stream) {
  if (K % 8 != 0 || args.n < 1 || args.n > 3) {
    set_error("lora_dx: unsupported K=%d adapters=%d", K, args.n);
    return 1;
  }
  const uint32_t th = drop_threshold(p);
  const float inv = th ? 1.0f / (1.0f - p) : 1.0f;
  int rmax = 0;
  bool aligned = K % 32 == 0 && ldx % 8 == 0 && ldt % 2 == 0 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(BT) & 3) == 0 && (mul == nullptr || (ldm % 8 == 0 && (reinterpret_cast<uintptr_t>(mul) & 15) == 0));
  for (int a = 0; a < args.n; ++a) {
    rmax = args.ad[a].r > rmax ? args.ad[a].r : rmax;
    aligned = aligned && args.ad[a].c0 % 2 == 0 && (reinterpret_cast<uintptr_t>(args.ad[a].A) & 15) == 0;
  }
  if (aligned && rmax <= 32) {
    int kslice = K;  // split the columns until the grid has a few CTAs per SM
    const int row_ctas = (rows + 63) / 64;
    while (kslice % 64 == 0 && row_ctas * (K / kslice) < 592) kslice /= 2;
    const dim3 grid(row_ctas, (K + kslice - 1) / kslice);
#define DX_LAUNCH(KS, DROP) \
  lora_dx_mma_kernel<KS, DROP><<<grid, 128, 0, stream>>>(dX, ldx, K, BT, ldt, args, mul, ldm, rows, accumulate, th, inv, row0, kslice)
    if (rmax <= 16) {
      if (th) DX_LAUNCH(1, true);
      else DX_LAUNCH(1, false);
    } else {
      if (th) DX_LAUNCH(2, true);
      else DX_LAUNCH(2, false);
    }
#undef DX_LAUNCH
  } else {
    const long long total = static_cast<long long>(rows) * (K / 8);
    lora_dx_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(dX, ldx, K, BT, ldt, args, mul, ldm, rows,
                                                                                   accumulate, th, inv, row0);
  }
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

// Tensor-core version: the same (128 columns x WG_ROWS rows) CTA and partial layout, four warps of 32 columns each.  Per
// 64-row stage the (masked) X rows and the S rows go to shared memory once (16-byte loads), and the reduction over the
// rows runs as m16n8k16 MMAs with BOTH operands read through ldmatrix.trans (the reduction index m is the row index of
// both tiles): C[n, j] += X^T[n, m] S[m, j].  F16: X is an fp16 stream (S is converted to fp16 while staging).
constexpr int WGX_LD = WG_COLS + 8;  // 272-byte rows: ldmatrix conflict-free
constexpr int WGS_LD = 24;           // 48-byte rows
template <bool F16>
__global__ void __launch_bounds__(128) wgrad_mma_kernel(const bf16* __restrict__ X, int ldx, int N, const bf16* __restrict__ S, int lds,
                                                        int c0, int r, int rows, float* __restrict__ P, uint32_t seed,
                                                        uint32_t thresh, float inv_keep, long long row0, int mask_ld) {
  __shared__ __align__(16) uint16_t Xs[64 * WGX_LD];
  __shared__ __align__(16) uint16_t Ss[64 * WGS_LD];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_blk = blockIdx.x * WG_COLS, m_begin = blockIdx.y * WG_ROWS;
  float acc[2][2][4];
#pragma unroll
  for (int a = 0; a < 2; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;
  const int xc = (tid & 15) * 8, xr0 = tid >> 4;   // X staging: 16 chunks per row, rows xr0 + 8 i
  const int sr = tid >> 1, sc8 = (tid & 1) * 8;    // S staging: row sr, columns sc8 .. sc8 + 7
  for (int ms = 0; ms < WG_ROWS; ms += 64) {
    __syncthreads();
    uint4 xv[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = m_begin + ms + xr0 + 8 * i;
      xv[i] = (m < rows && n_blk + xc < N) ? __ldg(reinterpret_cast<const uint4*>(X + static_cast<size_t>(m) * ldx + n_blk + xc))
                                           : make_uint4(0u, 0u, 0u, 0u);
    }
    {
      const int m = m_begin + ms + sr;
      uint32_t w[4] = {0u, 0u, 0u, 0u};
      if (m < rows) {
        const bf16* sp = S + static_cast<size_t>(m) * lds + c0 + sc8;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int j = sc8 + 2 * i;
          const float v0 = j < r ? __bfloat162float(sp[2 * i]) : 0.f, v1 = j + 1 < r ? __bfloat162float(sp[2 * i + 1]) : 0.f;
          if (F16) {
            const __half2 h = __floats2half2_rn(v0, v1);
            w[i] = *reinterpret_cast<const uint32_t*>(&h);
          } else {
            w[i] = pack_bf16x2(v0, v1);
          }
        }
      }
      *reinterpret_cast<uint4*>(Ss + sr * WGS_LD + sc8) = make_uint4(w[0], w[1], w[2], w[3]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t w[4] = {xv[i].x, xv[i].y, xv[i].z, xv[i].w};
      if (thresh != 0u) {
        const int m = m_begin + ms + xr0 + 8 * i;
        const uint32_t idx = static_cast<uint32_t>((row0 + m) * static_cast<long long>(mask_ld)) + n_blk + xc;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          w[q] &= (drop_keep(seed, idx + 2 * q, thresh) ? 0x0000ffffu : 0u) | (drop_keep(seed, idx + 2 * q + 1, thresh) ? 0xffff0000u : 0u);
      }
      *reinterpret_cast<uint4*>(Xs + (xr0 + 8 * i) * WGX_LD + xc) = make_uint4(w[0], w[1], w[2], w[3]);
    }
    __syncthreads();
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {
      uint32_t bq[4];
      ldsm_x4_t(bq, smem_u32(Ss + (ks * 16 + frag_r_lo(lane)) * WGS_LD + frag_c_hi(lane)));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        uint32_t aq[4];
        ldsm_x4_t(aq, smem_u32(Xs + (ks * 16 + frag_r_hi(lane)) * WGX_LD + warp * 32 + mt * 16 + frag_c_lo(lane)));
        if (F16) {
          mma16816_f16(acc[mt][0], aq, bq[0], bq[1]);
          mma16816_f16(acc[mt][1], aq, bq[2], bq[3]);
        } else {
          mma16816(acc[mt][0], aq, bq[0], bq[1]);
          mma16816(acc[mt][1], aq, bq[2], bq[3]);
        }
      }
    }
  }
  const float sc = thresh != 0u ? inv_keep : 1.0f;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int hb = 0; hb < 2; ++hb) {
      const int n = n_blk + warp * 32 + mt * 16 + g + hb * 8;
      if (n < N) {
        float* pp = P + (static_cast<size_t>(blockIdx.y) * N + n) * 16;
#pragma unroll
        for (int jt = 0; jt < 2; ++jt)
          *reinterpret_cast<float2*>(pp + jt * 8 + 2 * t) = make_float2(acc[mt][jt][2 * hb] * sc, acc[mt][jt][2 * hb + 1] * sc);
      }
    }
  }
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
  const bool mma_ok = N % 8 == 0 && ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(X) & 15) == 0 && (reinterpret_cast<uintptr_t>(partial) & 7) == 0;
  for (int j0 = 0; j0 < r; j0 += 16) {  // ranks above 16: 16 columns of S per pass
    const int rr = r - j0 < 16 ? r - j0 : 16;
    if (mma_ok && x_f16)
      wgrad_mma_kernel<true><<<grid, 128, 0, stream>>>(X, ldx, N, S, lds, c0 + j0, rr, rows, partial, seed, th, inv, row0, mask_ld);
    else if (mma_ok)
      wgrad_mma_kernel<false><<<grid, 128, 0, stream>>>(X, ldx, N, S, lds, c0 + j0, rr, rows, partial, seed, th, inv, row0, mask_ld);
    else if (rr <= 8)
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

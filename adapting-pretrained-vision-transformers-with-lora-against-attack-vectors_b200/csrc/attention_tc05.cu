// tcgen05 attention forward for ViT (T <= 208 tokens, head dim 64): persistent CTAs (one per SM) looping over
// (image, head) pairs, both 128-row query tiles of a head in flight at once (one softmax warp group each).
//
//   S = Q K^T        tcgen05.mma SS, M=128 (two query tiles), N=208 keys, K=64      -> TMEM cols [0,208) fp32
//   P = softmax(S)   4 warps, one query row per thread: tcgen05.ld -> exp2 -> bf16 -> tcgen05.st
//                    P overwrites S in place (two bf16 per 32-bit column, cols [0,104))
//   O = P V          tcgen05.mma TS: A = P from TMEM, B = V from smem (MN-major, 128B swizzle), N=64, K=208
//                    -> TMEM cols [128,192) fp32
// Q/K/V tiles arrive by 3-D TMA from the packed [B, T, 3*D] QKV GEMM output (rows >= T are zero-filled by
// the tensor map, so no masking of the operands is needed, only of the score columns).
// Replaces HF eager/sdpa attention forward (HF modeling_vit.py:185-193,228-249).
#include "ptx.cuh"
#include "vitatk_internal.h"

namespace vitatk {

static constexpr int A_HD = 64;
static constexpr int A_TPAD = 208;
static constexpr int A_THREADS = 320;      // 2 x 4 softmax warps (one query tile each) + TMA warp + MMA warp
static constexpr int A_TMA_WARP = 8, A_MMA_WARP = 9;
static constexpr int A_TMEM_COLS = 512;    // tile t: S at [256t, 256t+208), P in place, O at [256t+128, 256t+192)
static constexpr int A_TILE_COLS = 256;
static constexpr int A_O_COL = 128;
static constexpr int Q_BYTES = 2 * 128 * 128;
static constexpr int KV_BYTES = A_TPAD * 128;
static constexpr int A_STAGE_BYTES = Q_BYTES + 2 * KV_BYTES;
static constexpr int A_SMEM = 1024 + 2 * A_STAGE_BYTES + 256;

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Persistent: CTA c handles (image, head) pairs c, c + gridDim.x, ...; the next head's Q/K/V tiles are
// prefetched by TMA into the other smem stage while the current head is in the tensor cores / softmax warps.
__global__ void __launch_bounds__(A_THREADS, 1)
attn_fwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     bf16* __restrict__ out, float* __restrict__ lse2, int tokens, int heads, int num_items, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * A_STAGE_BYTES);
  uint64_t* load_full = bars;        // [2] TMA bytes landed                      (TMA -> MMA)
  uint64_t* load_empty = bars + 2;   // [2] all MMAs of the head retired           (MMA -> TMA)
  uint64_t* s_full = bars + 4;       // [2 tiles] S complete                       (MMA -> softmax)
  uint64_t* p_full = bars + 6;       // [2 tiles] P written to TMEM, 4 warp arrivals (softmax -> MMA)
  uint64_t* o_full = bars + 8;       // [2 tiles] O complete                       (MMA -> softmax)
  uint64_t* tmem_free = bars + 10;   // [2 tiles] O read out, 4 warp arrivals      (softmax -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * A_HD;
  const int ntiles = tokens > 128 ? 2 : 1;

  if (warp == A_MMA_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&load_full[i], 1);
      ptx::mbar_init(&load_empty[i], 1);
      ptx::mbar_init(&s_full[i], 1);
      ptx::mbar_init(&p_full[i], 4);
      ptx::mbar_init(&o_full[i], 1);
      ptx::mbar_init(&tmem_free[i], 4);
    }
    ptx::fence_mbar_init();
  }
  if (warp == A_TMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmQ);
      ptx::prefetch_tmap(&tmKV);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, A_TMEM_COLS);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == A_TMA_WARP) {
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const int b = item / heads, h = item % heads;
        ptx::mbar_wait(&load_empty[st], ((n >> 1) & 1) ^ 1);
        uint8_t* Qs = smem + st * A_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&load_full[st], A_STAGE_BYTES);
        ptx::tma_load_3d(Qs, &tmQ, &load_full[st], h * A_HD, 0, b);
        ptx::tma_load_3d(Qs + 128 * 128, &tmQ, &load_full[st], h * A_HD, 128, b);
        ptx::tma_load_3d(Qs + Q_BYTES, &tmKV, &load_full[st], D + h * A_HD, 0, b);
        ptx::tma_load_3d(Qs + Q_BYTES + KV_BYTES, &tmKV, &load_full[st], 2 * D + h * A_HD, 0, b);
      }
    }
  } else if (warp == A_MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, A_TPAD);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const uint32_t par = n & 1;
        const uint32_t qs = ptx::smem_u32(smem + st * A_STAGE_BYTES);
        const uint64_t kdesc = ptx::make_smem_desc_sw128(qs + Q_BYTES);
        const uint64_t vdesc = ptx::make_smem_desc_mn_sw128(qs + Q_BYTES + KV_BYTES, 1024);
        ptx::mbar_wait(&load_full[st], (n >> 1) & 1);
        for (int t = 0; t < ntiles; ++t) {
          ptx::mbar_wait(&tmem_free[t], par ^ 1);  // previous head's O of this tile has been read out
          ptx::tc_fence_after();
          const uint64_t qdesc = ptx::make_smem_desc_sw128(qs + t * 128 * 128);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            ptx::umma_bf16(tmem + t * A_TILE_COLS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
          ptx::umma_commit(&s_full[t]);
        }
        for (int t = 0; t < ntiles; ++t) {
          ptx::mbar_wait(&p_full[t], par);
          ptx::tc_fence_after();
          const uint32_t tt = tmem + t * A_TILE_COLS;
#pragma unroll 1
          for (int ks = 0; ks < A_TPAD / 16; ++ks)  // 16 keys per step: 8 TMEM columns of P, 16 rows (2 KB) of V
            ptx::umma_bf16_ts(tt + A_O_COL, tt + ks * 8, vdesc + ks * (2048 >> 4), idesc_pv, ks > 0 ? 1u : 0u);
          ptx::umma_commit(&o_full[t]);
        }
        ptx::umma_commit(&load_empty[st]);  // every MMA reading this stage has retired
      }
    }
  } else {
    const int t = warp >> 2;        // query tile of this warp group
    const int w4 = warp & 3;        // TMEM lane quarter
    if (t < ntiles) {
      const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16) + t * A_TILE_COLS;
      const int i = t * 128 + w4 * 32 + lane;                  // query row
      const bool warp_active = t * 128 + w4 * 32 < tokens;     // warp-uniform
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const uint32_t par = n & 1;
        const int b = item / heads, h = item % heads;
        ptx::mbar_wait(&s_full[t], par);
        ptx::tc_fence_after();
        float inv_sum = 0.f, lse = 0.f;
        if (warp_active) {
          float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, r);
            ptx::tmem_ld_wait();
            const int limit = tokens - c * 32;
            if (limit >= 32) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                mx0 = fmaxf(mx0, __uint_as_float(r[j]));
                mx1 = fmaxf(mx1, __uint_as_float(r[j + 1]));
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < limit) mx0 = fmaxf(mx0, __uint_as_float(r[j]));
            }
          }
          const float m2 = fmaxf(mx0, mx1) * sl2;
          float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            uint32_t r[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, r);
            ptx::tmem_ld_wait();
            const int limit = tokens - c * 32;
            uint32_t pk[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float p0 = ex2_approx(fmaf(__uint_as_float(r[2 * j]), sl2, -m2));
              float p1 = ex2_approx(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -m2));
              if (limit < 32) {
                if (2 * j >= limit) p0 = 0.f;
                if (2 * j + 1 >= limit) p1 = 0.f;
              }
              sum0 += p0;
              sum1 += p1;
              __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1);
              pk[j] = *reinterpret_cast<uint32_t*>(&v);
            }
            ptx::tmem_st_32x32b_x16(lane_addr + c * 16, pk);
          }
          ptx::tmem_st_wait();
          const float sum = sum0 + sum1;
          inv_sum = 1.f / sum;
          lse = m2 + log2f(sum);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&p_full[t]);
        ptx::mbar_wait(&o_full[t], par);
        ptx::tc_fence_after();
        if (warp_active) {
          uint32_t o0[32], o1[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + A_O_COL, o0);
          ptx::tmem_ld_32x32b_x32(lane_addr + A_O_COL + 32, o1);
          ptx::tmem_ld_wait();
          if (i < tokens) {
            uint4* dst = reinterpret_cast<uint4*>(out + (static_cast<size_t>(b) * tokens + i) * D + h * A_HD);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(o0[8 * j + 2 * k]) * inv_sum,
                                                         __uint_as_float(o0[8 * j + 2 * k + 1]) * inv_sum);
                w[k] = *reinterpret_cast<uint32_t*>(&v);
              }
              dst[j] = make_uint4(w[0], w[1], w[2], w[3]);
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint32_t w[4];
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(o1[8 * j + 2 * k]) * inv_sum,
                                                         __uint_as_float(o1[8 * j + 2 * k + 1]) * inv_sum);
                w[k] = *reinterpret_cast<uint32_t*>(&v);
              }
              dst[4 + j] = make_uint4(w[0], w[1], w[2], w[3]);
            }
            if (lse2) lse2[static_cast<size_t>(item) * A_TPAD + i] = lse;
          }
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_free[t]);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == A_TMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, A_TMEM_COLS);
  }
}

int attention_fwd_plan_init(AttnFwdPlan* p, const bf16* qkv, bf16* out, float* lse2, int batch, int tokens, int heads) {
  if (tokens < 1 || tokens > A_TPAD) {
    set_error("attention_fwd_tc05: tokens=%d unsupported (max %d)", tokens, A_TPAD);
    return 1;
  }
  p->batch = batch;
  p->tokens = tokens;
  p->heads = heads;
  p->qkv = qkv;
  p->out = out;
  p->lse2 = lse2;
  const uint64_t ld = 3ull * heads * A_HD;
  if (make_tmap_3d(&p->tmQ, qkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, A_HD, 128)) return 1;
  if (make_tmap_3d(&p->tmKV, qkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, A_HD, A_TPAD)) return 1;
  return 0;
}

int attention_fwd_tc05(const AttnFwdPlan* p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
    attr = true;
  }
  const float sl2 = 1.4426950408889634f / sqrtf(static_cast<float>(A_HD));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = p->batch * p->heads;
  attn_fwd_tc05_kernel<<<items < sms ? items : sms, A_THREADS, A_SMEM, stream>>>(p->tmQ, p->tmKV, p->out, p->lse2,
                                                                                p->tokens, p->heads, items, sl2);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}


// =================================================================================================
// Backward (input gradients), two persistent tcgen05 kernels that never transpose through shared memory:
//   attn_bwd_dq_kernel  work item = (image, head, 128-query tile):  S = Q K^T, dP = dO V^T  ->
//                       dS = P o (dP - delta) / sqrt(d)  (bf16, written back to TMEM)  ->  dQ = dS K   (TS MMA)
//                       also produces delta_i = sum_d dO[i,d] O[i,d] for the second kernel
//   attn_bwd_dkv_kernel work item = (image, head, 128-key tile):    S^T = K Q^T, dP^T = V dO^T  ->
//                       P^T and dS^T (bf16 in TMEM)  ->  dV = P^T dO,  dK = dS^T Q             (TS MMAs)
// The softmax statistics come from the forward (lse2, log2 domain).  Eight element-wise warps split every tile
// by columns (no row reductions are needed in the backward), the next item's operands are prefetched by TMA.
// Zero-filled operand rows (>= T) make every padded row/column contribute exactly zero, so no masks are needed.
// Replaces the autograd backward of HF attention (HF modeling_vit.py:185-193) w.r.t. q, k, v.
// =================================================================================================
static constexpr int B_THREADS = 320;
static constexpr int B_TMA_WARP = 8, B_MMA_WARP = 9;
static constexpr int B_A_BYTES = 128 * 128;                       // one 128-row operand tile
static constexpr int B_STAGE_BYTES = 2 * B_A_BYTES + 2 * KV_BYTES;  // A0, A1, B0[208], B1[208]
static constexpr int B_SMEM = 1024 + 2 * B_STAGE_BYTES + 256;

struct BwdBars {
  uint64_t *load_full, *load_empty, *mm1_full, *ew_full, *mm2_full, *tmem_free;
  uint32_t* tmem_slot;
};

__device__ __forceinline__ BwdBars bwd_setup(uint8_t* smem, int warp, int lane, int free_count, const CUtensorMap* m0,
                                             const CUtensorMap* m1, const CUtensorMap* m2, const CUtensorMap* m3) {
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * B_STAGE_BYTES);
  BwdBars b;
  b.load_full = bars;       // [2]
  b.load_empty = bars + 2;  // [2]
  b.mm1_full = bars + 4;    // score-like accumulators complete (MMA -> element-wise warps)
  b.ew_full = bars + 5;     // bf16 operands written back to TMEM, 8 warp arrivals (element-wise -> MMA)
  b.mm2_full = bars + 6;    // output accumulators complete (MMA -> epilogue)
  b.tmem_free = bars + 7;   // outputs read out (epilogue -> MMA)
  b.tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  if (warp == B_MMA_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&b.load_full[i], 1);
      ptx::mbar_init(&b.load_empty[i], 1);
    }
    ptx::mbar_init(b.mm1_full, 1);
    ptx::mbar_init(b.ew_full, 8);
    ptx::mbar_init(b.mm2_full, 1);
    ptx::mbar_init(b.tmem_free, free_count);
    ptx::fence_mbar_init();
  }
  if (warp == B_TMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(m0);
      ptx::prefetch_tmap(m1);
      ptx::prefetch_tmap(m2);
      ptx::prefetch_tmap(m3);
    }
    __syncwarp();
    ptx::tmem_alloc(b.tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  return b;
}

__device__ __forceinline__ void store_row64(bf16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32]) {
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(a[8 * j + 2 * k]), __uint_as_float(a[8 * j + 2 * k + 1]));
      w[k] = *reinterpret_cast<uint32_t*>(&v);
    }
    d[j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      __nv_bfloat162 v = __floats2bfloat162_rn(__uint_as_float(b[8 * j + 2 * k]), __uint_as_float(b[8 * j + 2 * k + 1]));
      w[k] = *reinterpret_cast<uint32_t*>(&v);
    }
    d[4 + j] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// TMEM map (dQ kernel): S [0,208)  dP [256,464)  dS(bf16): keys 0..127 at [0,64), keys 128..207 at [128,176)
//                       dQ accumulator [256,320)
__global__ void __launch_bounds__(B_THREADS, 1)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ128, const __grid_constant__ CUtensorMap tmDO128,
                   const __grid_constant__ CUtensorMap tmKV208, const bf16* __restrict__ dout,
                   const bf16* __restrict__ o, const float* __restrict__ lse2, float* __restrict__ delta,
                   bf16* __restrict__ dqkv, int tokens, int heads, int ntiles, int num_items, float sl2, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * A_HD;
  BwdBars bar = bwd_setup(smem, warp, lane, 4, &tmQ128, &tmDO128, &tmKV208, &tmKV208);
  const uint32_t tmem = *bar.tmem_slot;

  if (warp == B_TMA_WARP) {
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1, hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
        ptx::mbar_wait(&bar.load_empty[st], ((n >> 1) & 1) ^ 1);
        uint8_t* s0 = smem + st * B_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&bar.load_full[st], B_STAGE_BYTES);
        ptx::tma_load_3d(s0, &tmQ128, &bar.load_full[st], h * A_HD, t * 128, b);                     // Q tile
        ptx::tma_load_3d(s0 + B_A_BYTES, &tmDO128, &bar.load_full[st], h * A_HD, t * 128, b);        // dO tile
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES, &tmKV208, &bar.load_full[st], D + h * A_HD, 0, b);      // K
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES + KV_BYTES, &tmKV208, &bar.load_full[st], 2 * D + h * A_HD, 0, b);  // V
      }
    }
  } else if (warp == B_MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, A_TPAD);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const uint32_t par = n & 1;
        const uint32_t s0 = ptx::smem_u32(smem + st * B_STAGE_BYTES);
        const uint64_t qdesc = ptx::make_smem_desc_sw128(s0);
        const uint64_t dodesc = ptx::make_smem_desc_sw128(s0 + B_A_BYTES);
        const uint64_t kdesc = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES);
        const uint64_t vdesc = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES + KV_BYTES);
        const uint64_t kdesc_mn = ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES, 1024);
        ptx::mbar_wait(&bar.load_full[st], (n >> 1) & 1);
        ptx::mbar_wait(bar.tmem_free, par ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, dodesc + 2 * k, vdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(bar.mm1_full);
        ptx::mbar_wait(bar.ew_full, par);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < A_TPAD / 16; ++ks) {
          const uint32_t a = ks < 8 ? tmem + ks * 8 : tmem + 128 + (ks - 8) * 8;
          ptx::umma_bf16_ts(tmem + 256, a, kdesc_mn + ks * (2048 >> 4), idesc_o, ks > 0 ? 1u : 0u);
        }
        ptx::umma_commit(bar.mm2_full);
        ptx::umma_commit(&bar.load_empty[st]);
      }
    }
  } else {
    const int w4 = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    int n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      const int hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
      const int i = t * 128 + w4 * 32 + lane;
      const bool valid = i < tokens;
      float l = 0.f, dlt = 0.f;
      if (valid) {  // row statistics; overlaps the tensor-core phase
        l = __ldg(lse2 + static_cast<size_t>(hd) * A_TPAD + i);
        const size_t off = (static_cast<size_t>(b) * tokens + i) * D + h * A_HD;
        const uint4* po = reinterpret_cast<const uint4*>(o + off);
        const uint4* pd = reinterpret_cast<const uint4*>(dout + off);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 a = __ldg(po + j), c = __ldg(pd + j);
          const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
          const __nv_bfloat162* pc = reinterpret_cast<const __nv_bfloat162*>(&c);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float2 x = __bfloat1622float2(pa[k]), y = __bfloat1622float2(pc[k]);
            dlt = fmaf(x.x, y.x, dlt);
            dlt = fmaf(x.y, y.y, dlt);
          }
        }
        if (half == 0) delta[static_cast<size_t>(hd) * A_TPAD + i] = dlt;
      }
      ptx::mbar_wait(bar.mm1_full, par);
      ptx::tc_fence_after();
      const int c0 = half == 0 ? 0 : 4, c1 = half == 0 ? 4 : 7;
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        uint32_t s[32], dp[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, s);
        ptx::tmem_ld_32x32b_x32(lane_addr + 256 + c * 32, dp);
        ptx::tmem_ld_wait();
        uint32_t pk[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[2 * j]), sl2, -l));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[2 * j + 1]), sl2, -l));
          const float d0 = p0 * (__uint_as_float(dp[2 * j]) - dlt) * scale;
          const float d1 = p1 * (__uint_as_float(dp[2 * j + 1]) - dlt) * scale;
          __nv_bfloat162 v = __floats2bfloat162_rn(d0, d1);
          pk[j] = *reinterpret_cast<uint32_t*>(&v);
        }
        ptx::tmem_st_32x32b_x16(lane_addr + (half == 0 ? c * 16 : 128 + (c - 4) * 16), pk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar.ew_full);
      if (half == 0) {
        ptx::mbar_wait(bar.mm2_full, par);
        ptx::tc_fence_after();
        uint32_t a0[32], a1[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + 256, a0);
        ptx::tmem_ld_32x32b_x32(lane_addr + 256 + 32, a1);
        ptx::tmem_ld_wait();
        if (valid) store_row64(dqkv + (static_cast<size_t>(b) * tokens + i) * 3 * D + h * A_HD, a0, a1);
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(bar.tmem_free);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == B_TMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// TMEM map (dK/dV kernel): S^T [0,208)  dP^T [256,464)
//   P^T(bf16):  queries 0..127 at [0,64),    queries 128..207 at [128,176)      dV accumulator [64,128)
//   dS^T(bf16): queries 0..127 at [256,320), queries 128..207 at [384,432)      dK accumulator [320,384)
__global__ void __launch_bounds__(B_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ208,
                    const __grid_constant__ CUtensorMap tmDO208, const float* __restrict__ lse2,
                    const float* __restrict__ delta, bf16* __restrict__ dqkv, int tokens, int heads, int ntiles,
                    int num_items, float sl2, float scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * A_HD;
  BwdBars bar = bwd_setup(smem, warp, lane, 8, &tmKV128, &tmQ208, &tmDO208, &tmDO208);
  const uint32_t tmem = *bar.tmem_slot;

  if (warp == B_TMA_WARP) {
    if (lane == 0) {
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1, hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
        ptx::mbar_wait(&bar.load_empty[st], ((n >> 1) & 1) ^ 1);
        uint8_t* s0 = smem + st * B_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx(&bar.load_full[st], B_STAGE_BYTES);
        ptx::tma_load_3d(s0, &tmKV128, &bar.load_full[st], D + h * A_HD, t * 128, b);                  // K tile
        ptx::tma_load_3d(s0 + B_A_BYTES, &tmKV128, &bar.load_full[st], 2 * D + h * A_HD, t * 128, b);  // V tile
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES, &tmQ208, &bar.load_full[st], h * A_HD, 0, b);             // Q
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES + KV_BYTES, &tmDO208, &bar.load_full[st], h * A_HD, 0, b);  // dO
      }
    }
  } else if (warp == B_MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, A_TPAD);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const uint32_t par = n & 1;
        const uint32_t s0 = ptx::smem_u32(smem + st * B_STAGE_BYTES);
        const uint64_t kdesc = ptx::make_smem_desc_sw128(s0);
        const uint64_t vdesc = ptx::make_smem_desc_sw128(s0 + B_A_BYTES);
        const uint64_t qdesc = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES);
        const uint64_t dodesc = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES + KV_BYTES);
        const uint64_t qdesc_mn = ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES, 1024);
        const uint64_t dodesc_mn = ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES + KV_BYTES, 1024);
        ptx::mbar_wait(&bar.load_full[st], (n >> 1) & 1);
        ptx::mbar_wait(bar.tmem_free, par ^ 1);
        ptx::tc_fence_after();
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, kdesc + 2 * k, qdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem + 256, vdesc + 2 * k, dodesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(bar.mm1_full);
        ptx::mbar_wait(bar.ew_full, par);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < A_TPAD / 16; ++ks) {
          const uint32_t off = ks < 8 ? ks * 8 : 128 + (ks - 8) * 8;
          ptx::umma_bf16_ts(tmem + 64, tmem + off, dodesc_mn + ks * (2048 >> 4), idesc_o, ks > 0 ? 1u : 0u);       // dV
          ptx::umma_bf16_ts(tmem + 320, tmem + 256 + off, qdesc_mn + ks * (2048 >> 4), idesc_o, ks > 0 ? 1u : 0u);  // dK
        }
        ptx::umma_commit(bar.mm2_full);
        ptx::umma_commit(&bar.load_empty[st]);
      }
    }
  } else {
    const int w4 = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    int n = 0;
    for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
      const uint32_t par = n & 1;
      const int hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
      const int jrow = t * 128 + w4 * 32 + lane;  // key row
      const float* lrow = lse2 + static_cast<size_t>(hd) * A_TPAD;
      const float* drow = delta + static_cast<size_t>(hd) * A_TPAD;
      ptx::mbar_wait(bar.mm1_full, par);
      ptx::tc_fence_after();
      const int c0 = half == 0 ? 0 : 4, c1 = half == 0 ? 4 : 7;
#pragma unroll 1
      for (int c = c0; c < c1; ++c) {
        uint32_t s[32], dp[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, s);
        ptx::tmem_ld_32x32b_x32(lane_addr + 256 + c * 32, dp);
        uint32_t pk[16], dk[16];
        const bool in_range = c * 32 + 32 <= A_TPAD;  // chunk 6 covers queries 192..223: 208.. do not exist
        ptx::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 l4 = make_float4(0.f, 0.f, 0.f, 0.f), d4 = l4;
          if (in_range || j < 4) {
            l4 = __ldg(reinterpret_cast<const float4*>(lrow + c * 32) + j);
            d4 = __ldg(reinterpret_cast<const float4*>(drow + c * 32) + j);
          }
          const float p0 = ex2_approx(fmaf(__uint_as_float(s[4 * j]), sl2, -l4.x));
          const float p1 = ex2_approx(fmaf(__uint_as_float(s[4 * j + 1]), sl2, -l4.y));
          const float p2 = ex2_approx(fmaf(__uint_as_float(s[4 * j + 2]), sl2, -l4.z));
          const float p3 = ex2_approx(fmaf(__uint_as_float(s[4 * j + 3]), sl2, -l4.w));
          __nv_bfloat162 v0 = __floats2bfloat162_rn(p0, p1), v1 = __floats2bfloat162_rn(p2, p3);
          pk[2 * j] = *reinterpret_cast<uint32_t*>(&v0);
          pk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&v1);
          __nv_bfloat162 w0 = __floats2bfloat162_rn(p0 * (__uint_as_float(dp[4 * j]) - d4.x) * scale,
                                                    p1 * (__uint_as_float(dp[4 * j + 1]) - d4.y) * scale);
          __nv_bfloat162 w1 = __floats2bfloat162_rn(p2 * (__uint_as_float(dp[4 * j + 2]) - d4.z) * scale,
                                                    p3 * (__uint_as_float(dp[4 * j + 3]) - d4.w) * scale);
          dk[2 * j] = *reinterpret_cast<uint32_t*>(&w0);
          dk[2 * j + 1] = *reinterpret_cast<uint32_t*>(&w1);
        }
        const uint32_t off = half == 0 ? c * 16 : 128 + (c - 4) * 16;
        ptx::tmem_st_32x32b_x16(lane_addr + off, pk);
        ptx::tmem_st_32x32b_x16(lane_addr + 256 + off, dk);
      }
      ptx::tmem_st_wait();
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar.ew_full);
      ptx::mbar_wait(bar.mm2_full, par);
      ptx::tc_fence_after();
      {
        // half 0 stores dV (accumulator [64,128)), half 1 stores dK ([320,384))
        const uint32_t col = half == 0 ? 64 : 320;
        uint32_t a0[32], a1[32];
        ptx::tmem_ld_32x32b_x32(lane_addr + col, a0);
        ptx::tmem_ld_32x32b_x32(lane_addr + col + 32, a1);
        ptx::tmem_ld_wait();
        if (jrow < tokens)
          store_row64(dqkv + (static_cast<size_t>(b) * tokens + jrow) * 3 * D + (half == 0 ? 2 * D : D) + h * A_HD, a0, a1);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar.tmem_free);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == B_TMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

int attention_bwd_plan_init(AttnBwdPlan* p, const bf16* qkv, const bf16* dout, const bf16* o, const float* lse2,
                            float* delta, bf16* dqkv, int batch, int tokens, int heads) {
  if (tokens < 1 || tokens > A_TPAD) {
    set_error("attention_bwd_tc05: tokens=%d unsupported (max %d)", tokens, A_TPAD);
    return 1;
  }
  p->batch = batch;
  p->tokens = tokens;
  p->heads = heads;
  p->dout = dout;
  p->o = o;
  p->lse2 = lse2;
  p->delta = delta;
  p->dqkv = dqkv;
  const uint64_t ld = 3ull * heads * A_HD, ldo = 1ull * heads * A_HD;
  if (make_tmap_3d(&p->tmQKV128, qkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, A_HD, 128)) return 1;
  if (make_tmap_3d(&p->tmQKV208, qkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, A_HD, A_TPAD)) return 1;
  if (make_tmap_3d(&p->tmDO128, dout, ldo, tokens, batch, ldo * 2, ldo * 2 * tokens, A_HD, 128)) return 1;
  if (make_tmap_3d(&p->tmDO208, dout, ldo, tokens, batch, ldo * 2, ldo * 2 * tokens, A_HD, A_TPAD)) return 1;
  return 0;
}

int attention_bwd_tc05(const AttnBwdPlan* p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dq_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_dkv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    attr = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(A_HD));
  const float sl2 = scale * 1.4426950408889634f;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ntiles = p->tokens > 128 ? 2 : 1;
  const int items = p->batch * p->heads * ntiles;
  const int grid = items < sms ? items : sms;
  attn_bwd_dq_kernel<<<grid, B_THREADS, B_SMEM, stream>>>(p->tmQKV128, p->tmDO128, p->tmQKV208, p->dout, p->o, p->lse2,
                                                         p->delta, p->dqkv, p->tokens, p->heads, ntiles, items, sl2, scale);
  VITATK_CUDA_OK(cudaGetLastError());
  attn_bwd_dkv_kernel<<<grid, B_THREADS, B_SMEM, stream>>>(p->tmQKV128, p->tmQKV208, p->tmDO208, p->lse2, p->delta,
                                                          p->dqkv, p->tokens, p->heads, ntiles, items, sl2, scale);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

// tcgen05 attention forward for ViT (T <= 208 tokens, head dim 64): one CTA per (image, head), two CTAs per SM.
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
static constexpr int A_THREADS = 160;      // 4 softmax warps + 1 TMA/MMA warp
static constexpr int A_TMEM_COLS = 256;
static constexpr int A_O_COL = 128;
static constexpr int Q_BYTES = 2 * 128 * 128;
static constexpr int KV_BYTES = A_TPAD * 128;
static constexpr int A_SMEM = 1024 + Q_BYTES + 2 * KV_BYTES + 128;

__global__ void __launch_bounds__(A_THREADS, 2)
attn_fwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     bf16* __restrict__ out, float* __restrict__ lse2, int tokens, int heads, float sl2) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* Qs = smem;                    // [2][128 rows][128 B]
  uint8_t* Ks = Qs + Q_BYTES;            // [208][128 B]
  uint8_t* Vs = Ks + KV_BYTES;           // [208][128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(Vs + KV_BYTES);
  uint64_t* bar_load = bars;      // TMA bytes landed
  uint64_t* bar_s = bars + 1;     // S tile complete (MMA -> softmax)
  uint64_t* bar_p = bars + 2;     // P written to TMEM (softmax -> MMA), 4 warp arrivals
  uint64_t* bar_o = bars + 3;     // O tile complete (MMA -> softmax)
  uint64_t* bar_done = bars + 4;  // O read out, TMEM reusable (softmax -> MMA), 4 warp arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 5);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * A_HD;

  if (warp == 4) {
    if (lane == 0) {
      ptx::mbar_init(bar_load, 1);
      ptx::mbar_init(bar_s, 1);
      ptx::mbar_init(bar_p, 4);
      ptx::mbar_init(bar_o, 1);
      ptx::mbar_init(bar_done, 4);
      ptx::fence_mbar_init();
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

  if (warp == 4) {
    if (lane == 0) {
      ptx::mbar_arrive_expect_tx(bar_load, Q_BYTES + 2 * KV_BYTES);
      ptx::tma_load_3d(Qs, &tmQ, bar_load, h * A_HD, 0, b);
      ptx::tma_load_3d(Qs + 128 * 128, &tmQ, bar_load, h * A_HD, 128, b);
      ptx::tma_load_3d(Ks, &tmKV, bar_load, D + h * A_HD, 0, b);
      ptx::tma_load_3d(Vs, &tmKV, bar_load, 2 * D + h * A_HD, 0, b);
      ptx::mbar_wait(bar_load, 0);
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, A_TPAD);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      const uint64_t kdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(Ks));
      const uint64_t vdesc = ptx::make_smem_desc_mn_sw128(ptx::smem_u32(Vs), 1024);
      for (int t = 0; t < 2; ++t) {
        if (t * 128 >= tokens) break;
        if (t == 1) {  // tile 0's O has been read out of TMEM
          ptx::mbar_wait(bar_done, 0);
          ptx::tc_fence_after();
        }
        const uint64_t qdesc = ptx::make_smem_desc_sw128(ptx::smem_u32(Qs + t * 128 * 128));
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(tmem, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit(bar_s);
        ptx::mbar_wait(bar_p, t);
        ptx::tc_fence_after();
#pragma unroll 1
        for (int ks = 0; ks < A_TPAD / 16; ++ks)  // 16 keys per step: 8 TMEM columns of P, 16 rows (2 KB) of V
          ptx::umma_bf16_ts(tmem + A_O_COL, tmem + ks * 8, vdesc + ks * (2048 >> 4), idesc_pv, ks > 0 ? 1u : 0u);
        ptx::umma_commit(bar_o);
      }
    }
  } else {
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(warp * 32) << 16);
    for (int t = 0; t < 2; ++t) {
      if (t * 128 >= tokens) break;
      const int i = t * 128 + warp * 32 + lane;           // query row
      const bool warp_active = t * 128 + warp * 32 < tokens;  // warp-uniform
      ptx::mbar_wait(bar_s, t);
      ptx::tc_fence_after();
      float inv_sum = 0.f, lse = 0.f;
      if (warp_active) {
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 7; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, r);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < tokens) mx = fmaxf(mx, __uint_as_float(r[j]));
        }
        const float m2 = mx * sl2;
        float sum = 0.f;
#pragma unroll 1
        for (int c = 0; c < 7; ++c) {
          uint32_t r[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + c * 32, r);
          ptx::tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const int col = c * 32 + 2 * j;
            const float p0 = col < tokens ? exp2f(fmaf(__uint_as_float(r[2 * j]), sl2, -m2)) : 0.f;
            const float p1 = col + 1 < tokens ? exp2f(fmaf(__uint_as_float(r[2 * j + 1]), sl2, -m2)) : 0.f;
            sum += p0 + p1;
            __nv_bfloat162 v = __floats2bfloat162_rn(p0, p1);
            pk[j] = *reinterpret_cast<uint32_t*>(&v);
          }
          ptx::tmem_st_32x32b_x16(lane_addr + c * 16, pk);
        }
        ptx::tmem_st_wait();
        inv_sum = 1.f / sum;
        lse = m2 + log2f(sum);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_p);
      ptx::mbar_wait(bar_o, t);
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
          if (lse2) lse2[static_cast<size_t>(blockIdx.x) * A_TPAD + i] = lse;
        }
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(bar_done);
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 4) {
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
  attn_fwd_tc05_kernel<<<p->batch * p->heads, A_THREADS, A_SMEM, stream>>>(p->tmQ, p->tmKV, p->out, p->lse2, p->tokens,
                                                                          p->heads, sl2);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

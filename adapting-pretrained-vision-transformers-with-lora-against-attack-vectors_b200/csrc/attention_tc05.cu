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
#include <stdlib.h>

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
static constexpr int A_OUT_OFF = 2 * A_STAGE_BYTES;          // 8 warps x 4 KB output staging
static constexpr int A_SMEM = 1024 + 2 * A_STAGE_BYTES + 8 * 4096 + 256;

// optional in-kernel timeline of the forward kernel (timing experiments only, VITATK_ATTN_DBG & 32): CTA 0 records
// (event, unit, clock) triples into a device buffer set with attention_fwd_set_trace()
__device__ long long* g_fwd_trace = nullptr;
__device__ __forceinline__ void ftrace(int slot0, int& idx, int ev, int u) {
  if (g_fwd_trace != nullptr && blockIdx.x == 0 && idx < 680) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    g_fwd_trace[(slot0 + idx) * 2] = (static_cast<long long>(ev) << 32) | static_cast<unsigned>(u);
    g_fwd_trace[(slot0 + idx) * 2 + 1] = t;
    ++idx;
  }
}
#define FTR(slot0, ev, u) do { if (trace_on) ftrace(slot0, tr_idx, ev, u); } while (0)

__device__ __forceinline__ float max3(float a, float b, float c) {
  return fmaxf(a, fmaxf(b, c));  // (the 3-input max.f32 / FMNMX3 form measured slower on B200)
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}


// One warp: 32 rows x 64 bf16 columns (fp32 accumulators a|b, optional scale) -> 128B-swizzled smem tile ->
// 3-D TMA store into a [B, T, ld] tensor at (col, row0, b).  Rows >= T are clipped by the tensor map.
__device__ __forceinline__ void stage_store_64(uint8_t* stage, const uint32_t (&a)[32], const uint32_t (&b)[32],
                                               float mul, const CUtensorMap* tm, int col, int row0, int img, int lane) {
  if (lane == 0) ptx::tma_store_wait_read<0>();  // previous store from this buffer has been read out
  __syncwarp();
  const uint32_t row_base = ptx::smem_u32(stage) + lane * 128;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int e = (8 * j + 2 * k) & 31;
      const float lo = __uint_as_float(j < 4 ? a[e] : b[e]) * mul;
      const float hi = __uint_as_float(j < 4 ? a[e + 1] : b[e + 1]) * mul;
      __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
      w[k] = *reinterpret_cast<uint32_t*>(&v);
    }
    const uint32_t addr = row_base + ((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3])
                 : "memory");
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_3d(tm, stage, col, row0, img);
    ptx::tma_store_commit();
  }
}

// Persistent: CTA c handles (image, head) pairs c, c + gridDim.x, ...; the next head's Q/K/V tiles are
// prefetched by TMA into the other smem stage while the current head is in the tensor cores / softmax warps.
__global__ void __launch_bounds__(A_THREADS, 1)
attn_fwd_tc05_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmKV,
                     const __grid_constant__ CUtensorMap tmO, float* __restrict__ lse2, int tokens, int heads,
                     int num_items, float sl2, int trace_on) {
  int tr_idx = 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_OUT_OFF + 8 * 4096);
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
  (void)D;

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
  pdl_wait();
  pdl_launch_dependents();

  if (warp == A_TMA_WARP) {
    {
      const uint32_t leader = ptx::elect_leader();
      int n = 0;
      for (int item = blockIdx.x; item < num_items; item += gridDim.x, ++n) {
        const int st = n & 1;
        const int b = item / heads, h = item % heads;
        ptx::mbar_wait(&load_empty[st], ((n >> 1) & 1) ^ 1);
        uint8_t* Qs = smem + st * A_STAGE_BYTES;
        ptx::mbar_arrive_expect_tx_p(leader, &load_full[st], A_STAGE_BYTES);
        ptx::tma_load_3d_p(leader, Qs, &tmQ, &load_full[st], h * A_HD, 0, b);
        ptx::tma_load_3d_p(leader, Qs + 128 * 128, &tmQ, &load_full[st], h * A_HD, 128, b);
        ptx::tma_load_3d_p(leader, Qs + Q_BYTES, &tmKV, &load_full[st], D + h * A_HD, 0, b);
        ptx::tma_load_3d_p(leader, Qs + Q_BYTES + KV_BYTES, &tmKV, &load_full[st], 2 * D + h * A_HD, 0, b);
      }
    }
  } else if (warp == A_MMA_WARP) {
    {
      const uint32_t leader = ptx::elect_leader();
      constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, A_TPAD);
      constexpr uint32_t idesc_pv = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      // Work units are (head, query tile) pairs; S of unit u is issued BEFORE P*V of unit u - 1, so the two TMEM slots
      // run half a period out of phase: while one tile's softmax owns the MUFU pipe, the other tile's S / P*V MMAs and
      // read-out proceed (the previous in-order S,S,PV,PV schedule kept both tiles in lockstep: ~8.5K clk per head).
      const int my_items =
          (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
      const int units = my_items * ntiles;
      auto issue_pv = [&](int v) {
        const int n = v / ntiles, t = v - n * ntiles, st = n & 1;
        const uint32_t qs = ptx::smem_u32(smem + st * A_STAGE_BYTES);
        const uint64_t vdesc = ptx::make_smem_desc_mn_sw128(qs + Q_BYTES + KV_BYTES, 1024);
        ptx::mbar_wait(&p_full[t], n & 1);
        ptx::tc_fence_after();
        FTR(0, 4, v);
        const uint32_t tt = tmem + t * A_TILE_COLS;
#pragma unroll
        for (int ks = 0; ks < A_TPAD / 16; ++ks)  // 16 keys per step: 8 TMEM columns of P, 16 rows (2 KB) of V
          ptx::umma_bf16_ts_p(leader, tt + A_O_COL, tt + ks * 8, vdesc + ks * (2048 >> 4), idesc_pv, ks > 0 ? 1u : 0u);
        ptx::umma_commit_p(leader, &o_full[t]);
        FTR(0, 5, v);
        if (t == ntiles - 1) ptx::umma_commit_p(leader, &load_empty[st]);  // every MMA reading this stage has retired
      };
      for (int u = 0; u < units; ++u) {
        const int n = u / ntiles, t = u - n * ntiles, st = n & 1;
        const uint32_t qs = ptx::smem_u32(smem + st * A_STAGE_BYTES);
        const uint64_t kdesc = ptx::make_smem_desc_sw128(qs + Q_BYTES);
        FTR(0, 1, u);
        if (t == 0) ptx::mbar_wait(&load_full[st], (n >> 1) & 1);
        ptx::mbar_wait(&tmem_free[t], (n & 1) ^ 1);  // previous head's O of this tile has been read out
        ptx::tc_fence_after();
        FTR(0, 2, u);
        const uint64_t qdesc = ptx::make_smem_desc_sw128(qs + t * 128 * 128);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          ptx::umma_bf16_p(leader, tmem + t * A_TILE_COLS, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k > 0 ? 1u : 0u);
        ptx::umma_commit_p(leader, &s_full[t]);
        FTR(0, 3, u);
        if (u > 0) issue_pv(u - 1);
      }
      if (units > 0) issue_pv(units - 1);
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
        const int tslot = (w4 == 0 && lane == 0) ? 680 * (1 + t) : -1;
        ptx::mbar_wait(&s_full[t], par);
        ptx::tc_fence_after();
        if (tslot >= 0) FTR(tslot, 11, 2 * n + t);
        float inv_sum = 0.f, lse = 0.f;
        if (warp_active) {
          // pass 1 (row max): the 208 score columns are read in two batches (4 + 3 chunks of 32) so only two TMEM
          // round trips are exposed instead of seven
          float mx0 = -INFINITY, mx1 = -INFINITY;
          {
            uint32_t r0[32], r1[32], r2[32], r3[32];
            ptx::tmem_ld_32x32b_x32(lane_addr, r0);
            ptx::tmem_ld_32x32b_x32(lane_addr + 32, r1);
            ptx::tmem_ld_32x32b_x32(lane_addr + 64, r2);
            ptx::tmem_ld_32x32b_x32(lane_addr + 96, r3);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              mx0 = max3(mx0, __uint_as_float(r0[j]), __uint_as_float(r1[j]));
              mx1 = max3(mx1, __uint_as_float(r0[j + 1]), __uint_as_float(r1[j + 1]));
            }
            if (tokens >= 128) {
#pragma unroll
              for (int j = 0; j < 32; j += 2) {
                mx0 = max3(mx0, __uint_as_float(r2[j]), __uint_as_float(r3[j]));
                mx1 = max3(mx1, __uint_as_float(r2[j + 1]), __uint_as_float(r3[j + 1]));
              }
            } else {
              // fewer than 128 keys: redo the first four chunks with per-column limits
              mx0 = mx1 = -INFINITY;
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                if (j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r0[j]));
                if (32 + j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r1[j]));
                if (64 + j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r2[j]));
                if (96 + j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r3[j]));
              }
            }
          }
          if (tokens > 128) {
            uint32_t r0[32], r1[32], r2[32];
            ptx::tmem_ld_32x32b_x32(lane_addr + 128, r0);
            ptx::tmem_ld_32x32b_x32(lane_addr + 160, r1);
            ptx::tmem_ld_32x32b_x32(lane_addr + 192, r2);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (128 + j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r0[j]));
              if (160 + j < tokens) mx1 = fmaxf(mx1, __uint_as_float(r1[j]));
              if (192 + j < tokens) mx0 = fmaxf(mx0, __uint_as_float(r2[j]));
            }
          }
          const float m2 = fmaxf(mx0, mx1) * sl2;
          if (tslot >= 0) FTR(tslot, 12, 2 * n + t);
          // pass 2 (exp, row sum, bf16 P written over S): chunk c + 1 is in flight while chunk c is processed
          float sum0 = 0.f, sum1 = 0.f;
          uint32_t ra[32], rb[32];
          ptx::tmem_ld_32x32b_x32(lane_addr, ra);
#pragma unroll
          for (int c = 0; c < 7; ++c) {
            uint32_t(&r)[32] = (c & 1) ? rb : ra;
            uint32_t(&rn)[32] = (c & 1) ? ra : rb;
            ptx::tmem_ld_wait();
            if (c + 1 < 7) ptx::tmem_ld_32x32b_x32(lane_addr + (c + 1) * 32, rn);
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
        if (tslot >= 0) FTR(tslot, 13, 2 * n + t);
        ptx::mbar_wait(&o_full[t], par);
        ptx::tc_fence_after();
        if (tslot >= 0) FTR(tslot, 14, 2 * n + t);
        if (warp_active) {
          uint32_t o0[32], o1[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + A_O_COL, o0);
          ptx::tmem_ld_32x32b_x32(lane_addr + A_O_COL + 32, o1);
          ptx::tmem_ld_wait();
          stage_store_64(smem + A_OUT_OFF + warp * 4096, o0, o1, inv_sum, &tmO, h * A_HD, t * 128 + w4 * 32, b, lane);
          if (i < tokens) {
            if (lse2) lse2[static_cast<size_t>(item) * A_TPAD + i] = lse;
          }
        }
        // (releasing the slot before the store was measured SLOWER, 143 vs 102 us: it lets the two tiles' MUFU-bound
        // softmax phases drift into lockstep again)
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&tmem_free[t]);
        if (tslot >= 0) FTR(tslot, 15, 2 * n + t);
      }
      if (lane == 0) ptx::tma_store_wait_all<0>();
      __syncwarp();
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
  const uint64_t ldo = 1ull * heads * A_HD;
  if (make_tmap_3d(&p->tmO, out, ldo, tokens, batch, ldo * 2, ldo * 2 * tokens, A_HD, 32)) return 1;
  return 0;
}

int attention_fwd_set_trace(long long* dev_buf) {
  VITATK_CUDA_OK(cudaMemcpyToSymbol(g_fwd_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
}

int attention_fwd_tc05(const AttnFwdPlan* p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
    attr = true;
  }
  const float sl2 = 1.4426950408889634f / sqrtf(static_cast<float>(A_HD));
  static int trace_on = -1;
  if (trace_on < 0) {
    const char* e = getenv("VITATK_ATTN_DBG");
    trace_on = (e && (atoi(e) & 32)) ? 1 : 0;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = p->batch * p->heads;
  VITATK_CUDA_OK(launch_pdl(attn_fwd_tc05_kernel, dim3(items < sms ? items : sms), dim3(A_THREADS), A_SMEM, stream, 1,
                            p->tmQ, p->tmKV, p->tmO, p->lse2, p->tokens, p->heads, items, sl2, trace_on));
  return 0;
}


// =================================================================================================
// Backward (input gradients), two persistent tcgen05 kernels that never transpose through shared memory:
//   attn_bwd_kernel<0>  work item = (image, head, 128-query tile):  S = Q K^T, dP = dO V^T  ->
//                       dS = P o (dP - delta) / sqrt(d)  (bf16, written back to TMEM)  ->  dQ = dS K   (TS MMA)
//   attn_bwd_kernel<1>  work item = (image, head, 128-key tile):    S^T = K Q^T, dP^T = V dO^T  ->
//                       P^T and dS^T (bf16 in TMEM)  ->  dV = P^T dO,  dK = dS^T Q             (TS MMAs)
//   attn_delta_kernel   delta_i = sum_d dO[i,d] O[i,d] per (image, head, query), shared by both
// The softmax statistics come from the forward (lse2, log2 domain).  Eight element-wise warps split every tile
// by columns (no row reductions are needed in the backward), the next item's operands are prefetched by TMA.
// Zero-filled operand rows (>= T) make every padded row/column contribute exactly zero, so no masks are needed.
// Replaces the autograd backward of HF attention (HF modeling_vit.py:185-193) w.r.t. q, k, v.
// =================================================================================================
static constexpr int B_THREADS = 320;
static constexpr int B_TMA_WARP = 8, B_MMA_WARP = 9;
static constexpr int B_A_BYTES = 128 * 128;                               // one 128-row operand tile
static constexpr int B_STATS_OFF = 2 * B_A_BYTES + 2 * KV_BYTES;          // lse2 row, then delta row (1 KB apart)
static constexpr int B_STATS_BYTES = A_TPAD * 4;                          // 832
static constexpr int B_STAGE_BYTES = B_STATS_OFF + 2048;
static constexpr int B_TX_BYTES = 2 * B_A_BYTES + 2 * KV_BYTES + 2 * B_STATS_BYTES;
static constexpr int B_OUT_OFF = 2 * B_STAGE_BYTES;          // 8 warps x 4 KB output staging
static constexpr int B_SMEM = 1024 + 2 * B_STAGE_BYTES + 8 * 4096 + 256;
// TMEM: two ping-pong buffers of 192 columns, one per 64-column sub-block of the score matrix:
//   [0,64) X = S (or S^T) sub-block, [64,128) Y = dP (or dP^T), [128,160) bf16 operand 1, [160,192) bf16 operand 2
// and two 64-column output accumulators that live for the whole item.
static constexpr int B_BUFW = 192;
static constexpr int B_ACC0 = 384, B_ACC1 = 448;
static constexpr int B_NSB = 4;  // sub-blocks of the 208 columns: 64, 64, 64, 16

__device__ __forceinline__ void store_row32(bf16* dst, const uint32_t (&a)[32]) {
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
}

__device__ __forceinline__ void store_row64(bf16* dst, const uint32_t (&a)[32], const uint32_t (&b)[32]) {
  store_row32(dst, a);
  store_row32(dst + 32, b);
}

// MODE 0 (dQ):    rows = queries.  A0 = Q tile, A1 = dO tile, B0 = K, B1 = V.
//                 X = S = Q K^T, Y = dP = dO V^T, dS = P o (dP - delta)/sqrt(d) -> TMEM, ACC0 += dS K.
// MODE 1 (dK,dV): rows = keys.     A0 = K tile, A1 = V tile, B0 = Q, B1 = dO.
//                 X = S^T = K Q^T, Y = dP^T = V dO^T, P^T and dS^T -> TMEM, ACC0 (dV) += P^T dO, ACC1 (dK) += dS^T Q.
// The 208 score columns are processed as four sub-blocks through two ping-pong TMEM buffers, so the tensor
// cores compute sub-block g+1 / g+2 (also of the NEXT work item) while the element-wise warps are on g.
template <int MODE>
__global__ void __launch_bounds__(B_THREADS, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap mA0, const __grid_constant__ CUtensorMap mA1,
                const __grid_constant__ CUtensorMap mB0, const __grid_constant__ CUtensorMap mB1,
                const __grid_constant__ CUtensorMap mOut, int cA0, int cA1, int cB0, int cB1,
                const float* __restrict__ lse2, const float* __restrict__ delta, bf16* __restrict__ dqkv, int tokens, int heads, int ntiles, int num_items, float sl2, float scale,
                int dbg) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OUT_OFF + 8 * 4096);
  uint64_t* load_full = bars;       // [2] TMA -> MMA / element-wise
  uint64_t* load_empty = bars + 2;  // [2] MMA -> TMA
  uint64_t* xy_full = bars + 4;     // [2] X,Y of a sub-block complete        (MMA -> element-wise)
  uint64_t* a_full = bars + 6;      // [2] bf16 operands written, 8 arrivals   (element-wise -> MMA)
  uint64_t* acc_full = bars + 8;    // output accumulators complete           (MMA -> epilogue)
  uint64_t* acc_free = bars + 9;    // accumulators read out, 8 arrivals       (epilogue -> MMA)
  uint64_t* xy_used = bars + 10;    // [2] X,Y of a sub-block are in registers, 8 arrivals (element-wise -> MMA):
                                    //     lets the tensor cores refill the buffer while the warps still compute
  uint64_t* ts_done = bars + 12;    // [2] the TS MMAs that read a buffer's bf16 operands retired (MMA -> element-wise)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * A_HD;

  if (warp == B_MMA_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&load_full[i], 1);
      ptx::mbar_init(&load_empty[i], 1);
      ptx::mbar_init(&xy_full[i], 1);
      ptx::mbar_init(&a_full[i], 8);
      ptx::mbar_init(&xy_used[i], 8);
      ptx::mbar_init(&ts_done[i], 1);
    }
    ptx::mbar_init(acc_full, 1);
    ptx::mbar_init(acc_free, 8);
    ptx::fence_mbar_init();
  }
  if (warp == B_TMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(&mA0);
      ptx::prefetch_tmap(&mA1);
      ptx::prefetch_tmap(&mB0);
      ptx::prefetch_tmap(&mB1);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int my_items = (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == B_TMA_WARP) {
    if (lane == 0) {
      for (int n = 0; n < my_items; ++n) {
        const int item = blockIdx.x + n * gridDim.x;
        const int st = n & 1, hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
        ptx::mbar_wait(&load_empty[st], ((n >> 1) & 1) ^ 1);
        uint8_t* s0 = smem + st * B_STAGE_BYTES;
        if ((dbg & 1) && n >= 2) {  // timing experiment: no loads after the first two items
          ptx::mbar_arrive(&load_full[st]);
          continue;
        }
        ptx::mbar_arrive_expect_tx(&load_full[st], B_TX_BYTES);
        ptx::tma_load_3d(s0, &mA0, &load_full[st], cA0 + h * A_HD, t * 128, b);
        ptx::tma_load_3d(s0 + B_A_BYTES, &mA1, &load_full[st], cA1 + h * A_HD, t * 128, b);
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES, &mB0, &load_full[st], cB0 + h * A_HD, 0, b);
        ptx::tma_load_3d(s0 + 2 * B_A_BYTES + KV_BYTES, &mB1, &load_full[st], cB1 + h * A_HD, 0, b);
        ptx::bulk_load_1d(s0 + B_STATS_OFF, lse2 + static_cast<size_t>(hd) * A_TPAD, B_STATS_BYTES, &load_full[st]);
        ptx::bulk_load_1d(s0 + B_STATS_OFF + 1024, delta + static_cast<size_t>(hd) * A_TPAD, B_STATS_BYTES,
                          &load_full[st]);
      }
    }
  } else if (warp == B_MMA_WARP) {
    if (lane == 0) {
      constexpr uint32_t idesc_64 = ptx::make_idesc_bf16(128, 64);
      constexpr uint32_t idesc_16 = ptx::make_idesc_bf16(128, 16);
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, A_HD) | ptx::IDESC_B_MN_MAJOR;
      const int total_g = my_items * B_NSB;
      auto issue_xy = [&](int g) {
        const int n = g >> 2, sb = g & 3, st = n & 1;
        if (sb == 0) ptx::mbar_wait(&load_full[st], (n >> 1) & 1);
        const uint32_t s0 = ptx::smem_u32(smem + st * B_STAGE_BYTES);
        const uint64_t a0 = ptx::make_smem_desc_sw128(s0);
        const uint64_t a1 = ptx::make_smem_desc_sw128(s0 + B_A_BYTES);
        const uint64_t b0 = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES + sb * 64 * 128);
        const uint64_t b1 = ptx::make_smem_desc_sw128(s0 + 2 * B_A_BYTES + KV_BYTES + sb * 64 * 128);
        const uint32_t buf = tmem + (g & 1) * B_BUFW;
        const uint32_t idesc = sb < 3 ? idesc_64 : idesc_16;
        if (!(dbg & 16)) {
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(buf, a0 + 2 * k, b0 + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) ptx::umma_bf16(buf + 64, a1 + 2 * k, b1 + 2 * k, idesc, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&xy_full[g & 1]);
      };
      if (total_g > 0) issue_xy(0);
      if (total_g > 1) issue_xy(1);
      for (int g = 0; g < total_g; ++g) {
        const int n = g >> 2, sb = g & 3, st = n & 1;
        // refill this buffer's X/Y as soon as the element-wise warps hold sub-block g in registers
        ptx::mbar_wait(&xy_used[g & 1], (g >> 1) & 1);
        ptx::tc_fence_after();
        if (g + 2 < total_g) issue_xy(g + 2);
        ptx::mbar_wait(&a_full[g & 1], (g >> 1) & 1);
        if (sb == 0) ptx::mbar_wait(acc_free, (n & 1) ^ 1);  // previous item's outputs have been read out
        ptx::tc_fence_after();
        const uint32_t s0 = ptx::smem_u32(smem + st * B_STAGE_BYTES);
        const uint32_t buf = tmem + (g & 1) * B_BUFW;
        const int ksteps = sb < 3 ? 4 : 1;
        for (int ks = 0; ks < ((dbg & 16) ? 0 : ksteps); ++ks) {
          const uint32_t acc = (sb > 0 || ks > 0) ? 1u : 0u;
          const uint32_t rows = (sb * 64 + ks * 16) * 128;  // 16 reduction rows of the MN-major B operand
          if (MODE == 0) {
            ptx::umma_bf16_ts(tmem + B_ACC0, buf + 128 + ks * 8,
                              ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES + rows, 1024), idesc_o, acc);
          } else {
            ptx::umma_bf16_ts(tmem + B_ACC0, buf + 128 + ks * 8,
                              ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES + KV_BYTES + rows, 1024), idesc_o, acc);
            ptx::umma_bf16_ts(tmem + B_ACC1, buf + 160 + ks * 8,
                              ptx::make_smem_desc_mn_sw128(s0 + 2 * B_A_BYTES + rows, 1024), idesc_o, acc);
          }
        }
        ptx::umma_commit(&ts_done[g & 1]);
        if (sb == 3) {
          ptx::umma_commit(acc_full);
          ptx::umma_commit(&load_empty[st]);
        }
      }
    }
  } else {
    const int w4 = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    const int r = w4 * 32 + lane;
    const float nscale = -scale;
    for (int n = 0; n < my_items; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int st = n & 1, hd = item / ntiles, t = item % ntiles, b = hd / heads, h = hd % heads;
      const int row = t * 128 + r;  // query (MODE 0) or key (MODE 1) index of this thread
      (void)b;
      const float* sl = reinterpret_cast<const float*>(smem + st * B_STAGE_BYTES + B_STATS_OFF);
      const float* sd = sl + 256;
      ptx::mbar_wait(&load_full[st], (n >> 1) & 1);
      float l_row = 0.f, nds_row = 0.f;
      if (MODE == 0 && row < A_TPAD) {
        l_row = sl[row];
        nds_row = sd[row] * nscale;
      }
#pragma unroll 1
      for (int sb = 0; sb < B_NSB; ++sb) {
        const int g = n * B_NSB + sb;
        ptx::mbar_wait(&xy_full[g & 1], (g >> 1) & 1);
        ptx::tc_fence_after();
        const uint32_t buf = lane_addr + (g & 1) * B_BUFW;
        if (sb < 3 || half == 0) {
          const int coff = sb < 3 ? 32 * half : 0;
          uint32_t x[32], y[32];
          if (!(dbg & 8)) {
          ptx::tmem_ld_32x32b_x32(buf + coff, x);
          ptx::tmem_ld_32x32b_x32(buf + 64 + coff, y);
          ptx::tmem_ld_wait();
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) x[j] = y[j] = j + lane;
          }
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&xy_used[g & 1]);
          uint32_t o1[16], o2[16];
          const float* cl = sl + sb * 64 + coff;  // column statistics (MODE 1)
          const float* cd = sd + sb * 64 + coff;
          if (dbg & 2) {  // timing experiment: no element-wise math
#pragma unroll
            for (int j = 0; j < 16; ++j) o1[j] = o2[j] = x[j] ^ y[j];
          } else
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float4 l4, d4;
            if (MODE == 0) {
              l4 = make_float4(l_row, l_row, l_row, l_row);
              d4 = make_float4(nds_row, nds_row, nds_row, nds_row);
            } else {
              l4 = *reinterpret_cast<const float4*>(cl + 4 * j);
              d4 = *reinterpret_cast<const float4*>(cd + 4 * j);
              d4.x *= nscale; d4.y *= nscale; d4.z *= nscale; d4.w *= nscale;
            }
            const float p0 = ex2_approx(fmaf(__uint_as_float(x[4 * j]), sl2, -l4.x));
            const float p1 = ex2_approx(fmaf(__uint_as_float(x[4 * j + 1]), sl2, -l4.y));
            const float p2 = ex2_approx(fmaf(__uint_as_float(x[4 * j + 2]), sl2, -l4.z));
            const float p3 = ex2_approx(fmaf(__uint_as_float(x[4 * j + 3]), sl2, -l4.w));
            // dS = P * (dP - delta) * scale
            const float e0 = p0 * fmaf(__uint_as_float(y[4 * j]), scale, d4.x);
            const float e1 = p1 * fmaf(__uint_as_float(y[4 * j + 1]), scale, d4.y);
            const float e2 = p2 * fmaf(__uint_as_float(y[4 * j + 2]), scale, d4.z);
            const float e3 = p3 * fmaf(__uint_as_float(y[4 * j + 3]), scale, d4.w);
            __nv_bfloat162 v0 = __floats2bfloat162_rn(e0, e1), v1 = __floats2bfloat162_rn(e2, e3);
            if (MODE == 0) {
              o1[2 * j] = *reinterpret_cast<uint32_t*>(&v0);
              o1[2 * j + 1] = *reinterpret_cast<uint32_t*>(&v1);
            } else {
              __nv_bfloat162 w0 = __floats2bfloat162_rn(p0, p1), w1 = __floats2bfloat162_rn(p2, p3);
              o1[2 * j] = *reinterpret_cast<uint32_t*>(&w0);
              o1[2 * j + 1] = *reinterpret_cast<uint32_t*>(&w1);
              o2[2 * j] = *reinterpret_cast<uint32_t*>(&v0);
              o2[2 * j + 1] = *reinterpret_cast<uint32_t*>(&v1);
            }
          }
          ptx::mbar_wait(&ts_done[g & 1], ((g >> 1) & 1) ^ 1);  // TS MMAs of sub-block g-2 no longer read this buffer
          ptx::tc_fence_after();
          if (!(dbg & 8)) {
          ptx::tmem_st_32x32b_x16(buf + 128 + (sb < 3 ? 16 * half : 0), o1);
          if (MODE == 1) ptx::tmem_st_32x32b_x16(buf + 160 + (sb < 3 ? 16 * half : 0), o2);
          ptx::tmem_st_wait();
          } else if (o1[3] == 0x12345 && o2[5] == 0x777) {
            dqkv[0] = __float2bfloat16(1.f);
          }
        } else {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&xy_used[g & 1]);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(&a_full[g & 1]);
      }
      // ---- epilogue: accumulators -> bf16 -> dqkv ----
      ptx::mbar_wait(acc_full, n & 1);
      ptx::tc_fence_after();
      if (MODE == 0) {
        if (half == 0) {  // dQ: 64 columns of ACC0 -> q slot
          uint32_t a0[32], a1[32];
          ptx::tmem_ld_32x32b_x32(lane_addr + B_ACC0, a0);
          ptx::tmem_ld_32x32b_x32(lane_addr + B_ACC0 + 32, a1);
          ptx::tmem_ld_wait();
          if (!(dbg & 4))
            stage_store_64(smem + B_OUT_OFF + warp * 4096, a0, a1, 1.0f, &mOut, h * A_HD, t * 128 + w4 * 32, b, lane);
        }
      } else {  // half 0: dV (ACC0) -> v slot, half 1: dK (ACC1) -> k slot
        uint32_t a0[32], a1[32];
        const uint32_t col = half == 0 ? B_ACC0 : B_ACC1;
        ptx::tmem_ld_32x32b_x32(lane_addr + col, a0);
        ptx::tmem_ld_32x32b_x32(lane_addr + col + 32, a1);
        ptx::tmem_ld_wait();
        if (!(dbg & 4))
          stage_store_64(smem + B_OUT_OFF + warp * 4096, a0, a1, 1.0f, &mOut, (half == 0 ? 2 * D : D) + h * A_HD,
                         t * 128 + w4 * 32, b, lane);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(acc_free);
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == B_TMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

// delta[b,h,i] = sum_d dO[i, h*64+d] * O[i, h*64+d]   (one warp per token row, 8 lanes per head)
__global__ void __launch_bounds__(256) attn_delta_kernel_2k(const bf16* __restrict__ dout, const bf16* __restrict__ o,
                                                         float* __restrict__ delta, int rows, int tokens, int heads) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int D = heads * A_HD;
  const int b = row / tokens, i = row % tokens;
  const uint4* pa = reinterpret_cast<const uint4*>(dout + static_cast<size_t>(row) * D);
  const uint4* pb = reinterpret_cast<const uint4*>(o + static_cast<size_t>(row) * D);
  for (int c = lane; c < D / 8; c += 32) {  // chunk c holds 8 elements of head c / 8
    const uint4 x = __ldg(pa + c), y = __ldg(pb + c);
    const __nv_bfloat162* px = reinterpret_cast<const __nv_bfloat162*>(&x);
    const __nv_bfloat162* py = reinterpret_cast<const __nv_bfloat162*>(&y);
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float2 u = __bfloat1622float2(px[k]), v = __bfloat1622float2(py[k]);
      acc = fmaf(u.x, v.x, acc);
      acc = fmaf(u.y, v.y, acc);
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    if ((lane & 7) == 0) delta[(static_cast<size_t>(b) * heads + c / 8) * A_TPAD + i] = acc;
  }
}

int attention_delta(const AttnBwdPlan* p, cudaStream_t stream) {
  const int rows = p->batch * p->tokens;
  attn_delta_kernel_2k<<<(rows + 7) / 8, 256, 0, stream>>>(p->dout, p->o, p->delta, rows, p->tokens, p->heads);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
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
  if (make_tmap_3d(&p->tmDqkv, dqkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, A_HD, 32)) return 1;
  if (make_tmap_3d(&p->tmDqkv32, dqkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, 32, 32, 64)) return 1;
  return 0;
}

int attention_bwd_tc05(const AttnBwdPlan* p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
    attr = true;
  }
  if (p->heads * A_HD % 256 != 0) {
    set_error("attention_bwd_tc05: heads*64 must be a multiple of 256");
    return 1;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(A_HD));
  const float sl2 = scale * 1.4426950408889634f;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int D = p->heads * A_HD;
  const int ntiles = p->tokens > 128 ? 2 : 1;
  const int items = p->batch * p->heads * ntiles;
  const int grid = items < sms ? items : sms;
  const int rows = p->batch * p->tokens;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("VITATK_ATTN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  attn_delta_kernel_2k<<<(rows + 7) / 8, 256, 0, stream>>>(p->dout, p->o, p->delta, rows, p->tokens, p->heads);
  VITATK_CUDA_OK(cudaGetLastError());
  // dQ: A0 = Q tile, A1 = dO tile, B0 = K, B1 = V
  attn_bwd_kernel<0><<<grid, B_THREADS, B_SMEM, stream>>>(p->tmQKV128, p->tmDO128, p->tmQKV208, p->tmQKV208, p->tmDqkv, 0, 0, D,
                                                         2 * D, p->lse2, p->delta, p->dqkv, p->tokens, p->heads, ntiles,
                                                         items, sl2, scale, dbg);
  VITATK_CUDA_OK(cudaGetLastError());
  // dK,dV: A0 = K tile, A1 = V tile, B0 = Q, B1 = dO
  attn_bwd_kernel<1><<<grid, B_THREADS, B_SMEM, stream>>>(p->tmQKV128, p->tmQKV128, p->tmQKV208, p->tmDO208, p->tmDqkv, D, 2 * D, 0,
                                                         0, p->lse2, p->delta, p->dqkv, p->tokens, p->heads, ntiles,
                                                         items, sl2, scale, dbg);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

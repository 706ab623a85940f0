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
                     int num_items, float sl2, int trace_arg) {
#ifdef VITATK_DBG_KERNELS
  const int trace_on = trace_arg;
#else
  constexpr int trace_on = 0;  // the product build carries no timing-experiment branches
  (void)trace_arg;
#endif
  int tr_idx = 0;
  (void)tr_idx;
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
#ifdef VITATK_DBG_KERNELS
  VITATK_CUDA_OK(cudaMemcpyToSymbol(g_fwd_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
#else
  (void)dev_buf;
  set_error("attention_fwd_set_trace: build with -DVITATK_DBG_KERNELS (VITATK_DBG_BUILD=1) for the in-kernel timeline");
  return 1;
#endif
}

int attention_fwd_tc05(const AttnFwdPlan* p, cudaStream_t stream) {
  static PerDeviceOnce once;
  if (once.need())
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_fwd_tc05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
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
// Backward-side helpers shared with attention_bwd_fused.cu: the delta kernel (only used when the proj-backward GEMM
// epilogue does not produce delta) and the tensor maps of the fused single-pass backward.
// =================================================================================================
// delta[b,h,i] = sum_d dO[i, h*64+d] * O[i, h*64+d]   (one warp per token row, 8 lanes per head)
__global__ void __launch_bounds__(256) attn_delta_kernel(const bf16* __restrict__ dout, const bf16* __restrict__ o,
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
  attn_delta_kernel<<<(rows + 7) / 8, 256, 0, stream>>>(p->dout, p->o, p->delta, rows, p->tokens, p->heads);
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
  if (make_tmap_3d(&p->tmDO208, dout, ldo, tokens, batch, ldo * 2, ldo * 2 * tokens, A_HD, A_TPAD)) return 1;
  if (make_tmap_3d(&p->tmDqkv32, dqkv, ld, tokens, batch, ld * 2, ld * 2 * tokens, 32, 32, 64)) return 1;
  return 0;
}

}  // namespace vitatk

// Single-pass tcgen05 attention backward (input gradients dQ, dK, dV) for ViT heads (T <= 208 tokens, d = 64).
//
// One persistent CTA per SM loops over (image, head) pairs.  TMEM lanes are KEYS; the 208 query columns are walked in
// sub-blocks of <= 64, twice (once per 128-key tile):
//
//   X = K_t Q_sb^T,  Y = V_t dO_sb^T                 SS MMAs  -> TMEM ping-pong buffer (fp32, 64 + 64 columns)
//   P^T = exp2(X*c - lse),  dS^T = P^T o (Y - delta)  8 element-wise warps, computed ONCE per (key, query) pair;
//                                                    bf16 results overwrite X / Y in place (TMEM) and dS^T is also
//                                                    written to shared memory in the MN-major 128B-swizzle layout
//   dV_t += P^T dO_sb,  dK_t += dS^T Q_sb            TS MMAs (A operand straight from TMEM)
//   dQ_qt += dS K_t                                   SS MMA, A = dS^T tile read transposed (MN-major descriptor)
//
// so the softmax probabilities and dS are evaluated once (the two-kernel version in attention_tc05.cu evaluates them
// twice and issues 7 GEMMs instead of 5), nothing is transposed through registers, and dQ needs no atomics because one
// CTA owns the whole head.  1/sqrt(d) is applied when the dQ / dK accumulators are read out.  Operand rows >= T are
// zero-filled by the TMA tensor maps, which makes every padded row / column contribute exactly zero: no masks.
// Replaces the autograd backward of HF attention (HF modeling_vit.py:185-193) w.r.t. q, k, v.
#include <stdlib.h>

#include "ptx.cuh"
#include "vitatk_internal.h"

namespace vitatk {
namespace {

constexpr int HD = 64;
constexpr int TPAD = 208;
constexpr int F_THREADS = 320;  // 8 element-wise / epilogue warps + TMA warp + MMA warp
constexpr int F_TMA_WARP = 8, F_MMA_WARP = 9;
constexpr int KV_TILE = 128 * 128;   // one 128-row K or V tile (128 B per row)
constexpr int QD_BYTES = TPAD * 128;  // all query rows of Q or dO
constexpr int STAT_BYTES = TPAD * 4;
// shared-memory map (offsets from the 1024-aligned base)
constexpr int OFF_K = 0;                            // [2 key tiles]
constexpr int OFF_V = 2 * KV_TILE;                  // [2 key tiles]
constexpr int OFF_Q = 4 * KV_TILE;                  // [2 stages]
constexpr int OFF_DO = OFF_Q + 2 * QD_BYTES;        // [2 stages]
constexpr int OFF_DS = OFF_DO + 2 * QD_BYTES;       // dS^T: 2 blocks of [128 keys x 64 queries] bf16
constexpr int OFF_STATS = OFF_DS + 2 * KV_TILE;     // [2 stages][lse2 1 KB | delta 1 KB]
constexpr int OFF_STAGE = OFF_STATS + 4096;         // 8 warps x 2 KB output staging
constexpr int OFF_BARS = OFF_STAGE + 8 * 2048;
constexpr int F_SMEM = 1024 + OFF_BARS + 256;
// TMEM columns
constexpr int TM_BUF = 128;  // buffer b at 128*b: X at +0, Y at +64 (bf16 P^T / dS^T overwrite them in place)
constexpr int TM_DV = 256, TM_DK = 320, TM_DQ = 384;  // dQ tile qt at TM_DQ + 64*qt

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// one warp: 32 rows x 32 fp32 accumulator columns -> bf16 -> 64B-swizzled smem tile -> 3-D TMA store
__device__ __forceinline__ void stage_store_32(uint8_t* stage, const uint32_t (&a)[32], float mul, const CUtensorMap* tm,
                                               int col, int row0, int img, int lane) {
  if (lane == 0) ptx::tma_store_wait_read<0>();
  __syncwarp();
  const uint32_t row_base = ptx::smem_u32(stage) + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
      w[k] = pack2(__uint_as_float(a[8 * j + 2 * k]) * mul, __uint_as_float(a[8 * j + 2 * k + 1]) * mul);
    const uint32_t addr = row_base + ((j ^ sw) << 4);
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

__global__ void __launch_bounds__(F_THREADS, 1)
attn_bwd_fused_kernel(const __grid_constant__ CUtensorMap tmKV128, const __grid_constant__ CUtensorMap tmQ208,
                      const __grid_constant__ CUtensorMap tmDO208, const __grid_constant__ CUtensorMap tmOut,
                      const float* __restrict__ lse2, const float* __restrict__ delta, int tokens, int heads,
                      int num_items, float sl2, float scale, int dbg) {
  // dbg (VITATK_ATTN_DBG, timing experiments only): 1 no accumulator read-out, 2 no element-wise math,
  // 4 no dS^T smem tile / dQ MMAs, 8 no X/Y/TS MMAs
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* qdo_full = bars;        // [2] Q, dO, stats of a head landed                  (TMA -> MMA, element-wise)
  uint64_t* qdo_empty = bars + 2;   // [2] every MMA of the head retired                   (MMA -> TMA)
  uint64_t* kv_full = bars + 4;     // [2 key tiles] K_t, V_t landed                       (TMA -> MMA)
  uint64_t* kv_empty = bars + 6;    // [2 key tiles] every MMA reading K_t / V_t retired   (MMA -> TMA)
  uint64_t* xy_full = bars + 8;     // [2 buffers] X, Y of a step complete                 (MMA -> element-wise)
  uint64_t* a_full = bars + 10;     // [2 buffers] bf16 operands written, 8 arrivals       (element-wise -> MMA)
  uint64_t* kvacc_full = bars + 12; // dV_t, dK_t accumulators complete                    (MMA -> epilogue)
  uint64_t* dq_full = bars + 13;    // dQ accumulators complete                            (MMA -> epilogue)
  uint64_t* ds_free = bars + 14;    // the dQ MMAs that read the dS^T smem tile retired    (MMA -> element-wise)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int ncols = (tokens + 15) & ~15;             // query columns actually processed
  const int nsb = (ncols + 63) >> 6;                 // sub-blocks of <= 64 query columns
  const int nkt = tokens > 128 ? 2 : 1;              // 128-key tiles
  const int nsteps = nkt * nsb;

  if (warp == F_MMA_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&qdo_full[i], 1);
      ptx::mbar_init(&qdo_empty[i], 1);
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&xy_full[i], 1);
      ptx::mbar_init(&a_full[i], 8);
    }
    ptx::mbar_init(kvacc_full, 1);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(ds_free, 1);
    ptx::fence_mbar_init();
  }
  if (warp == F_TMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmKV128);
      ptx::prefetch_tmap(&tmQ208);
      ptx::prefetch_tmap(&tmDO208);
      ptx::prefetch_tmap(&tmOut);
    }
    __syncwarp();
    ptx::tmem_alloc(tmem_slot, 512);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const int my_items =
      (num_items - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);

  if (warp == F_TMA_WARP) {
    // ======================================= TMA producer =======================================
    if (lane == 0) {
      for (int n = 0; n < my_items; ++n) {
        const int item = blockIdx.x + n * gridDim.x;
        const int b = item / heads, h = item % heads, st = n & 1;
        ptx::mbar_wait(&qdo_empty[st], ((n >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx(&qdo_full[st], 2 * QD_BYTES + 2 * STAT_BYTES);
        ptx::tma_load_3d(smem + OFF_Q + st * QD_BYTES, &tmQ208, &qdo_full[st], h * HD, 0, b);
        ptx::tma_load_3d(smem + OFF_DO + st * QD_BYTES, &tmDO208, &qdo_full[st], h * HD, 0, b);
        ptx::bulk_load_1d(smem + OFF_STATS + st * 2048, lse2 + static_cast<size_t>(item) * TPAD, STAT_BYTES,
                          &qdo_full[st]);
        ptx::bulk_load_1d(smem + OFF_STATS + st * 2048 + 1024, delta + static_cast<size_t>(item) * TPAD, STAT_BYTES,
                          &qdo_full[st]);
        for (int kt = 0; kt < nkt; ++kt) {
          ptx::mbar_wait(&kv_empty[kt], (n & 1) ^ 1);
          ptx::mbar_arrive_expect_tx(&kv_full[kt], 2 * KV_TILE);
          ptx::tma_load_3d(smem + OFF_K + kt * KV_TILE, &tmKV128, &kv_full[kt], D + h * HD, 128 * kt, b);
          ptx::tma_load_3d(smem + OFF_V + kt * KV_TILE, &tmKV128, &kv_full[kt], 2 * D + h * HD, 128 * kt, b);
        }
      }
    }
  } else if (warp == F_MMA_WARP) {
    // ======================================= MMA issuer =======================================
    if (lane == 0) {
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HD) | ptx::IDESC_B_MN_MAJOR;
      constexpr uint32_t idesc_q = ptx::make_idesc_bf16(128, HD) | ptx::IDESC_A_MN_MAJOR | ptx::IDESC_B_MN_MAJOR;
      const uint32_t sbase = ptx::smem_u32(smem);
      const int total = my_items * nsteps;
      auto issue_xy = [&](int G) {
        const int n = G / nsteps, r = G % nsteps, kt = r / nsb, sb = r % nsb, st = n & 1;
        if (r == 0) ptx::mbar_wait(&qdo_full[st], (n >> 1) & 1);
        if (sb == 0) ptx::mbar_wait(&kv_full[kt], n & 1);
        ptx::tc_fence_after();
        const int w = min(64, ncols - 64 * sb);
        const uint32_t idesc = ptx::make_idesc_bf16(128, static_cast<uint32_t>(w));
        const uint64_t a0 = ptx::make_smem_desc_sw128(sbase + OFF_K + kt * KV_TILE);
        const uint64_t a1 = ptx::make_smem_desc_sw128(sbase + OFF_V + kt * KV_TILE);
        const uint64_t b0 = ptx::make_smem_desc_sw128(sbase + OFF_Q + st * QD_BYTES + sb * 64 * 128);
        const uint64_t b1 = ptx::make_smem_desc_sw128(sbase + OFF_DO + st * QD_BYTES + sb * 64 * 128);
        const uint32_t buf = tmem + (G & 1) * TM_BUF;
        if (!(dbg & 8)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(buf, a0 + 2 * k, b0 + 2 * k, idesc, k > 0 ? 1u : 0u);
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16(buf + 64, a1 + 2 * k, b1 + 2 * k, idesc, k > 0 ? 1u : 0u);
        }
        ptx::umma_commit(&xy_full[G & 1]);
      };
      if (total > 0) issue_xy(0);
      if (total > 1) issue_xy(1);
      for (int G = 0; G < total; ++G) {
        const int n = G / nsteps, r = G % nsteps, kt = r / nsb, sb = r % nsb, st = n & 1;
        ptx::mbar_wait(&a_full[G & 1], (G >> 1) & 1);
        ptx::tc_fence_after();
        const int w = min(64, ncols - 64 * sb);
        const uint32_t buf = tmem + (G & 1) * TM_BUF;
        const uint32_t q_s = sbase + OFF_Q + st * QD_BYTES, do_s = sbase + OFF_DO + st * QD_BYTES;
        for (int ks = 0; ks < ((dbg & 8) ? 0 : (w >> 4)); ++ks) {
          const uint32_t acc = (sb > 0 || ks > 0) ? 1u : 0u;
          const uint32_t rows = (sb * 64 + ks * 16) * 128;  // 16 query rows = one k-step of the MN-major B operand
          const uint32_t acol = 32 * (ks >> 1) + 8 * (ks & 1);
          ptx::umma_bf16_ts(tmem + TM_DV, buf + acol, ptx::make_smem_desc_mn_sw128(do_s + rows, 1024), idesc_o, acc);
          ptx::umma_bf16_ts(tmem + TM_DK, buf + 64 + acol, ptx::make_smem_desc_mn_sw128(q_s + rows, 1024), idesc_o,
                            acc);
        }
        const bool last_sb = sb == nsb - 1;
        if ((sb & 1) || last_sb) {
          // dQ_qt += dS[queries of this pair of sub-blocks, keys of tile kt] * K_kt
          const int qt = sb >> 1;
          const int kvalid = min(128, tokens - 128 * kt);
          const int ks5 = (kvalid + 15) >> 4;
          for (int ks = 0; ks < ((dbg & 4) ? 0 : ks5); ++ks)
            ptx::umma_bf16(tmem + TM_DQ + 64 * qt, ptx::make_smem_desc_mn_sw128(sbase + OFF_DS + ks * 2048, KV_TILE),
                           ptx::make_smem_desc_mn_sw128(sbase + OFF_K + kt * KV_TILE + ks * 2048, 1024), idesc_q,
                           (kt > 0 || ks > 0) ? 1u : 0u);
          ptx::umma_commit(ds_free);
        }
        if (last_sb) {
          ptx::umma_commit(kvacc_full);
          ptx::umma_commit(&kv_empty[kt]);
          if (kt == nkt - 1) {
            ptx::umma_commit(&qdo_empty[st]);
            ptx::umma_commit(dq_full);
          }
        }
        if (G + 2 < total) issue_xy(G + 2);
      }
    }
  } else {
    // ======================================= element-wise + epilogue warps =======================================
    const int w4 = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    const int row = w4 * 32 + lane;  // key row inside the 128-key tile
    uint8_t* my_stage = smem + OFF_STAGE + warp * 2048;
    const uint32_t ds_row = ptx::smem_u32(smem + OFF_DS) + row * 128;
    int G = 0, pairs = 0, tiles = 0;
    // deferred read-out of finished accumulators (hidden behind the next step's math)
    bool pend_kv = false, pend_dq = false;
    int pkv_kt = 0, pkv_b = 0, pkv_h = 0, pkv_par = 0, pdq_b = 0, pdq_h = 0, pdq_par = 0;

    auto flush_kv = [&]() {
      ptx::mbar_wait(kvacc_full, pkv_par);
      ptx::tc_fence_after();
      if (128 * pkv_kt + 32 * w4 < tokens && !(dbg & 1)) {
        uint32_t a[32];
        const int col = (half == 0 ? 2 * D : D) + pkv_h * HD;  // half 0: dV -> v slot, half 1: dK -> k slot
        const float mul = half == 0 ? 1.0f : scale;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          ptx::tmem_ld_32x32b_x32(lane_addr + (half == 0 ? TM_DV : TM_DK) + 32 * c, a);
          ptx::tmem_ld_wait();
          stage_store_32(my_stage, a, mul, &tmOut, col + 32 * c, 128 * pkv_kt + 32 * w4, pkv_b, lane);
        }
      }
      pend_kv = false;
    };
    auto flush_dq = [&]() {
      ptx::mbar_wait(dq_full, pdq_par);
      ptx::tc_fence_after();
      if (128 * half + 32 * w4 < tokens && !(dbg & 1)) {  // half selects the 128-query tile
        uint32_t a[32];
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          ptx::tmem_ld_32x32b_x32(lane_addr + TM_DQ + 64 * half + 32 * c, a);
          ptx::tmem_ld_wait();
          stage_store_32(my_stage, a, scale, &tmOut, pdq_h * HD + 32 * c, 128 * half + 32 * w4, pdq_b, lane);
        }
      }
      pend_dq = false;
    };

    for (int n = 0; n < my_items; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / heads, h = item % heads, st = n & 1;
      const float* sl = reinterpret_cast<const float*>(smem + OFF_STATS + st * 2048);
      const float* sd = sl + 256;
      ptx::mbar_wait(&qdo_full[st], (n >> 1) & 1);  // statistics are visible
      for (int kt = 0; kt < nkt; ++kt) {
        const int kvalid = min(128, tokens - 128 * kt);
        const bool warp_active = 32 * w4 < ((kvalid + 15) & ~15);  // rows the dQ MMA will read
#pragma unroll 1
        for (int sb = 0; sb < nsb; ++sb, ++G) {
          const int w = min(64, ncols - 64 * sb);
          const int my_w = max(0, min(32, w - 32 * half));  // 0, 16 or 32 query columns for this warp
          const uint32_t buf = lane_addr + (G & 1) * TM_BUF;
          ptx::mbar_wait(&xy_full[G & 1], (G >> 1) & 1);
          ptx::tc_fence_after();
          if (warp_active && my_w > 0) {
            uint32_t x[32], y[32];
            ptx::tmem_ld_32x32b_x32(buf + 32 * half, x);
            ptx::tmem_ld_32x32b_x32(buf + 64 + 32 * half, y);
            ptx::tmem_ld_wait();
            uint32_t o1[16], o2[16];
            const float4* cl = reinterpret_cast<const float4*>(sl + 64 * sb + 32 * half);
            const float4* cd = reinterpret_cast<const float4*>(sd + 64 * sb + 32 * half);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (dbg & 2) {
                o1[2 * j] = x[4 * j] ^ y[4 * j + 1];
                o1[2 * j + 1] = x[4 * j + 2] ^ y[4 * j + 3];
                o2[2 * j] = x[4 * j + 1] ^ y[4 * j];
                o2[2 * j + 1] = x[4 * j + 3] ^ y[4 * j + 2];
              } else if (j < 4 || my_w == 32) {
                const float4 l4 = cl[j], d4 = cd[j];
                const float p0 = ex2f(fmaf(__uint_as_float(x[4 * j]), sl2, -l4.x));
                const float p1 = ex2f(fmaf(__uint_as_float(x[4 * j + 1]), sl2, -l4.y));
                const float p2 = ex2f(fmaf(__uint_as_float(x[4 * j + 2]), sl2, -l4.z));
                const float p3 = ex2f(fmaf(__uint_as_float(x[4 * j + 3]), sl2, -l4.w));
                // unscaled dS = P o (dP - delta); 1/sqrt(d) is applied when dQ / dK are read out
                const float e0 = p0 * (__uint_as_float(y[4 * j]) - d4.x);
                const float e1 = p1 * (__uint_as_float(y[4 * j + 1]) - d4.y);
                const float e2 = p2 * (__uint_as_float(y[4 * j + 2]) - d4.z);
                const float e3 = p3 * (__uint_as_float(y[4 * j + 3]) - d4.w);
                o1[2 * j] = pack2(p0, p1);
                o1[2 * j + 1] = pack2(p2, p3);
                o2[2 * j] = pack2(e0, e1);
                o2[2 * j + 1] = pack2(e2, e3);
              } else {
                o1[2 * j] = o1[2 * j + 1] = o2[2 * j] = o2[2 * j + 1] = 0u;
              }
            }
            // bf16 operands of the TS MMAs overwrite this warp's own X / Y columns
            if (my_w == 32) {
              ptx::tmem_st_32x32b_x16(buf + 32 * half, o1);
              ptx::tmem_st_32x32b_x16(buf + 64 + 32 * half, o2);
            } else {
              ptx::tmem_st_32x32b_x8(buf + 32 * half, o1);
              ptx::tmem_st_32x32b_x8(buf + 64 + 32 * half, o2);
            }
            // dS^T tile for the dQ MMA: row = key (128 B = 64 queries), 16-byte chunks XOR-swizzled by row & 7
            if (!(sb & 1) && pairs > 0) ptx::mbar_wait(ds_free, (pairs - 1) & 1);
            const uint32_t blk = ds_row + (sb & 1) * KV_TILE;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if ((i < 2 || my_w == 32) && !(dbg & 4)) {
                const uint32_t addr = blk + (((4 * half + i) ^ (row & 7)) << 4);
                asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o2[4 * i]), "r"(o2[4 * i + 1]),
                             "r"(o2[4 * i + 2]), "r"(o2[4 * i + 3])
                             : "memory");
              }
            }
            ptx::tmem_st_wait();
            ptx::fence_proxy_async_smem();
          }
          if (pend_kv) flush_kv();
          if (pend_dq) flush_dq();
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&a_full[G & 1]);
          if ((sb & 1) || sb == nsb - 1) ++pairs;
        }
        pend_kv = true;
        pkv_kt = kt;
        pkv_b = b;
        pkv_h = h;
        pkv_par = tiles & 1;
        ++tiles;
      }
      pend_dq = true;
      pdq_b = b;
      pdq_h = h;
      pdq_par = n & 1;
    }
    if (pend_kv) flush_kv();
    if (pend_dq) flush_dq();
    if (lane == 0) ptx::tma_store_wait_all<0>();
    __syncwarp();
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == F_TMA_WARP) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem, 512);
  }
}

}  // namespace

int attention_delta(const AttnBwdPlan* p, cudaStream_t stream);  // attention_tc05.cu

int attention_bwd_fused(const AttnBwdPlan* p, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
    attr = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  const float sl2 = scale * 1.4426950408889634f;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int items = p->batch * p->heads;
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("VITATK_ATTN_DBG");
    dbg = e ? atoi(e) : 0;
  }
  if (!(dbg & 16) && attention_delta(p, stream)) return 1;
  attn_bwd_fused_kernel<<<items < sms ? items : sms, F_THREADS, F_SMEM, stream>>>(
      p->tmQKV128, p->tmQKV208, p->tmDO208, p->tmDqkv32, p->lse2, p->delta, p->tokens, p->heads, items, sl2, scale,
      dbg);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

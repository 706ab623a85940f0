// Single-pass tcgen05 attention backward (input gradients dQ, dK, dV) for ViT heads (T <= 208 tokens, d = 64).
//
// One persistent CTA per SM loops over (image, head) pairs.  TMEM lanes are KEYS; the 208 query columns are walked in
// sub-blocks of <= 64, twice (once per 128-key tile):
//
//   X = K_t Q_sb^T,  Y = V_t dO_sb^T                 SS MMAs  -> TMEM ping-pong buffer (fp32, 64 + 64 columns)
//   P^T = exp2(X*c - lse),  dS^T = P^T o (Y - delta)  8 element-wise warps, computed ONCE per (key, query) pair;
//                                                    bf16 P^T overwrites X in place (TMEM); dS^T goes to shared
//                                                    memory as a [keys x queries] 128B-swizzle tile
//   dV_t += P^T dO_sb                                 TS MMA (A operand straight from TMEM)
//   dK_t += dS^T Q_sb                                 SS MMA, A = the dS^T smem tile read as a K-major operand
//   dQ_qt += dS K_t                                   SS MMA, A = the same tile read transposed (MN-major descriptor)
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
constexpr int F_THREADS = 448;  // 8 element-wise warps + TMA warp + MMA warp + 4 accumulator read-out warps
constexpr int F_TMA_WARP = 8, F_MMA_WARP = 9, F_EPI_WARP0 = 10;
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
constexpr int OFF_STAGE = OFF_STATS + 4096;         // 4 read-out warps x 2 x 2 KB output staging
constexpr int OFF_BARS = OFF_STAGE + 8 * 2048;
constexpr int F_SMEM = 1024 + OFF_BARS + 256;
// TMEM columns
constexpr int TM_BUF = 128;  // buffer b at 128*b: X at +0, Y at +64 (bf16 P^T overwrites X in place)
constexpr int TM_DV = 256, TM_DK = 320, TM_DQ = 384;  // dQ tile qt at TM_DQ + 64*qt

// optional in-kernel timeline (timing experiments only): CTA 0 records (event, step, clock) triples
__device__ long long* g_trace = nullptr;
__device__ __forceinline__ void trace(int slot0, int& idx, int ev, int G) {
  if (g_trace != nullptr && blockIdx.x == 0 && idx < 1024) {
    long long t;
    asm volatile("mov.u64 %0, %%clock64;" : "=l"(t));
    g_trace[(slot0 + idx) * 2] = (static_cast<long long>(ev) << 32) | static_cast<unsigned>(G);
    g_trace[(slot0 + idx) * 2 + 1] = t;
    ++idx;
  }
}
#define TR(slot0, ev, G) do { if (dbg & 32) trace(slot0, tr_idx, ev, G); } while (0)

__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {  // explicit shared-space load (a generic LD is slower)
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}

// 32 fp32 accumulator columns of one lane -> 16 packed bf16 pairs
__device__ __forceinline__ void pack_chunk(const uint32_t (&a)[32], float mul, uint32_t (&p)[16]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) p[j] = pack2(__uint_as_float(a[2 * j]) * mul, __uint_as_float(a[2 * j + 1]) * mul);
}
// TMEM -> registers -> packed bf16 for one 32-column accumulator chunk (one chunk at a time keeps the read-out
// warps inside their 128-register budget)
__device__ __forceinline__ void read_pack(uint32_t taddr, float mul, uint32_t (&p)[16]) {
  uint32_t a[32];
  ptx::tmem_ld_32x32b_x32(taddr, a);
  ptx::tmem_ld_wait();
  pack_chunk(a, mul, p);
}
// one warp: 32 rows x 32 bf16 columns from registers -> 64B-swizzled smem tile -> 3-D TMA store
__device__ __forceinline__ void stage_store_32(uint8_t* stage, const uint32_t (&p)[16], const CUtensorMap* tm, int col,
                                               int row0, int img, int lane) {
  if (lane == 0) ptx::tma_store_wait_read<1>();  // the buffer used two stores ago has been read out
  __syncwarp();
  const uint32_t row_base = ptx::smem_u32(stage) + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t addr = row_base + ((j ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(p[4 * j]), "r"(p[4 * j + 1]),
                 "r"(p[4 * j + 2]), "r"(p[4 * j + 3])
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
                      int num_items, float sl2, float scale, int dbg_arg) {
#ifdef VITATK_DBG_KERNELS
  const int dbg = dbg_arg;
#else
  constexpr int dbg = 0;  // the product build carries no timing-experiment branches
  (void)dbg_arg;
#endif
  // dbg (VITATK_ATTN_DBG with -DVITATK_DBG_KERNELS, timing experiments only): 1 no accumulator read-out, 2 no element-wise math,
  // 4 no dS^T smem tile / dK / dQ MMAs, 8 no X/Y/dV MMAs, 16 no delta, 32 timeline, 64 no statistics loads.
  // Measured (B = 256): 276 us; without the statistics loads 270; without ANY element-wise math 225 -- the kernel is
  // bound by the MMA-issue / mbarrier hand-off chain of its 64-query steps, not by the element-wise work.
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BARS);
  uint64_t* qdo_full = bars;         // [2] Q, dO, stats of a head landed                  (TMA -> MMA, element-wise)
  uint64_t* qdo_empty = bars + 2;    // [2] every MMA of the head retired                   (MMA -> TMA)
  uint64_t* kv_full = bars + 4;      // [2 key tiles] K_t, V_t landed                       (TMA -> MMA)
  uint64_t* kv_empty = bars + 6;     // [2 key tiles] every MMA reading K_t / V_t retired   (MMA -> TMA)
  uint64_t* xy_full = bars + 8;      // [2 buffers] X, Y of a step complete                 (MMA -> element-wise)
  uint64_t* a_full = bars + 10;      // [2 buffers] bf16 P^T written to TMEM, 8 arrivals    (element-wise -> MMA)
  uint64_t* ds_full = bars + 12;     // [2 buffers] dS^T block written to smem, 8 arrivals  (element-wise -> MMA)
  uint64_t* ds_free = bars + 14;     // the MMAs that read the dS^T smem tile retired       (MMA -> element-wise)
  uint64_t* kvacc_full = bars + 15;  // dV_t, dK_t accumulators complete                    (MMA -> read-out)
  uint64_t* acc_free = bars + 16;    // dV_t, dK_t are in registers, 4 arrivals             (read-out -> MMA)
  uint64_t* dq_full = bars + 17;     // dQ accumulators complete                            (MMA -> read-out)
  uint64_t* dq_free = bars + 18;     // dQ accumulators are in registers, 4 arrivals        (read-out -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int D = heads * HD;
  const int ncols = (tokens + 15) & ~15;  // query columns actually processed
  const int nsb = (ncols + 63) >> 6;      // sub-blocks of <= 64 query columns
  const int nkt = tokens > 128 ? 2 : 1;   // 128-key tiles
  const int nsteps = nkt * nsb;

  if (warp == F_MMA_WARP && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      ptx::mbar_init(&qdo_full[i], 1);
      ptx::mbar_init(&qdo_empty[i], 1);
      ptx::mbar_init(&kv_full[i], 1);
      ptx::mbar_init(&kv_empty[i], 1);
      ptx::mbar_init(&xy_full[i], 1);
      ptx::mbar_init(&a_full[i], 8);
      ptx::mbar_init(&ds_full[i], 8);
    }
    ptx::mbar_init(ds_free, 1);
    ptx::mbar_init(kvacc_full, 1);
    ptx::mbar_init(acc_free, 4);
    ptx::mbar_init(dq_full, 1);
    ptx::mbar_init(dq_free, 4);
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
  pdl_wait();
  pdl_launch_dependents();

  if (warp == F_TMA_WARP) {
    // ======================================= TMA producer (converged warp) =======================================
    {
      const uint32_t leader = ptx::elect_leader();
      for (int n = 0; n < my_items; ++n) {
        const int item = blockIdx.x + n * gridDim.x;
        const int b = item / heads, h = item % heads, st = n & 1;
        ptx::mbar_wait(&qdo_empty[st], ((n >> 1) & 1) ^ 1);
        ptx::mbar_arrive_expect_tx_p(leader, &qdo_full[st], 2 * QD_BYTES + 2 * STAT_BYTES);
        ptx::tma_load_3d_p(leader, smem + OFF_Q + st * QD_BYTES, &tmQ208, &qdo_full[st], h * HD, 0, b);
        ptx::tma_load_3d_p(leader, smem + OFF_DO + st * QD_BYTES, &tmDO208, &qdo_full[st], h * HD, 0, b);
        ptx::bulk_load_1d_p(leader, smem + OFF_STATS + st * 2048, lse2 + static_cast<size_t>(item) * TPAD, STAT_BYTES,
                          &qdo_full[st]);
        ptx::bulk_load_1d_p(leader, smem + OFF_STATS + st * 2048 + 1024, delta + static_cast<size_t>(item) * TPAD, STAT_BYTES,
                          &qdo_full[st]);
        for (int kt = 0; kt < nkt; ++kt) {
          ptx::mbar_wait(&kv_empty[kt], (n & 1) ^ 1);
          ptx::mbar_arrive_expect_tx_p(leader, &kv_full[kt], 2 * KV_TILE);
          ptx::tma_load_3d_p(leader, smem + OFF_K + kt * KV_TILE, &tmKV128, &kv_full[kt], D + h * HD, 128 * kt, b);
          ptx::tma_load_3d_p(leader, smem + OFF_V + kt * KV_TILE, &tmKV128, &kv_full[kt], 2 * D + h * HD, 128 * kt, b);
        }
      }
    }
  } else if (warp == F_MMA_WARP) {
    // ======================================= MMA issuer (converged warp) =======================================
    // The single issuing lane is the scarce resource of this kernel (~20 MMAs per step), so descriptors are kept as
    // precomputed low words (+ a constant high word) and the step counters advance without integer divisions.
    {
      const uint32_t leader = ptx::elect_leader();
      constexpr uint32_t idesc_o = ptx::make_idesc_bf16(128, HD) | ptx::IDESC_B_MN_MAJOR;
      constexpr uint32_t idesc_q = ptx::make_idesc_bf16(128, HD) | ptx::IDESC_A_MN_MAJOR | ptx::IDESC_B_MN_MAJOR;
      constexpr uint32_t DESC_HI = (1024u >> 4) | (1u << 14) | (2u << 29);  // SBO 1024 B, version 1, SWIZZLE_128B
      auto mk = [](uint32_t lo) -> uint64_t { return (static_cast<uint64_t>(DESC_HI) << 32) | lo; };
      auto lo_k = [](uint32_t addr) -> uint32_t { return (addr >> 4) | (1u << 16); };               // K-major
      auto lo_mn = [](uint32_t addr, uint32_t lbo) -> uint32_t { return (addr >> 4) | ((lbo >> 4) << 16); };  // MN-major
      const uint32_t sbase = ptx::smem_u32(smem);
      // two-entry tables as selects (a runtime-indexed local array would live in local memory)
      const uint32_t k_lo0 = lo_k(sbase + OFF_K), v_lo0 = lo_k(sbase + OFF_V), kmn_lo0 = lo_mn(sbase + OFF_K, 1024);
      const uint32_t q_lo0 = lo_k(sbase + OFF_Q), do_lo0 = lo_k(sbase + OFF_DO);
      const uint32_t qmn_lo0 = lo_mn(sbase + OFF_Q, 1024), domn_lo0 = lo_mn(sbase + OFF_DO, 1024);
      const uint32_t dsk_lo0 = lo_k(sbase + OFF_DS), dsmn_lo = lo_mn(sbase + OFF_DS, KV_TILE);
      constexpr uint32_t KV_STEP = KV_TILE >> 4, QD_STEP = QD_BYTES >> 4;  // second tile / stage in descriptor units
      const int total = my_items * nsteps;
      struct Step { int n, kt, sb; };
      auto advance = [&](Step& s) {
        if (++s.sb == nsb) {
          s.sb = 0;
          if (++s.kt == nkt) {
            s.kt = 0;
            ++s.n;
          }
        }
      };
      auto issue_xy = [&](const Step& s, int G) {
        const int st = s.n & 1;
        if (s.kt == 0 && s.sb == 0) ptx::mbar_wait(&qdo_full[st], (s.n >> 1) & 1);
        if (s.sb == 0) ptx::mbar_wait(&kv_full[s.kt], s.n & 1);
        ptx::tc_fence_after();
        const int w = min(64, ncols - 64 * s.sb);
        const uint32_t idesc = ptx::make_idesc_bf16(128, static_cast<uint32_t>(w));
        const uint32_t a0 = k_lo0 + s.kt * KV_STEP, a1 = v_lo0 + s.kt * KV_STEP;
        const uint32_t b0 = q_lo0 + st * QD_STEP + s.sb * 512, b1 = do_lo0 + st * QD_STEP + s.sb * 512;  // 64 rows * 128 B >> 4
        const uint32_t buf = tmem + (G & 1) * TM_BUF;
        if (!(dbg & 8)) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {  // X and Y are independent accumulation chains: interleave them
            ptx::umma_bf16_p(leader, buf, mk(a0 + 2 * k), mk(b0 + 2 * k), idesc, k > 0 ? 1u : 0u);
            ptx::umma_bf16_p(leader, buf + 64, mk(a1 + 2 * k), mk(b1 + 2 * k), idesc, k > 0 ? 1u : 0u);
          }
        }
        ptx::umma_commit_p(leader, &xy_full[G & 1]);
      };
      Step cur = {0, 0, 0}, ahead = {0, 0, 0};
      if (total > 0) {
        issue_xy(ahead, 0);
        advance(ahead);
      }
      if (total > 1) {
        issue_xy(ahead, 1);
        advance(ahead);
      }
      int tile_idx = 0;
      int tr_idx = 0;
      for (int G = 0; G < total; ++G, advance(cur)) {
        const int n = cur.n, kt = cur.kt, sb = cur.sb, st = n & 1;
        const int w = min(64, ncols - 64 * sb);
        const int nks = w >> 4;
        TR(0, 1, G);
        const uint32_t buf = tmem + (G & 1) * TM_BUF;
        const bool last_sb = sb == nsb - 1;
        const uint32_t brow = (sb * 64) * 8;  // first query row of the sub-block in an MN-major B operand, (row * 128 B) >> 4
        // ---- P^T is in TMEM: dV_t += P^T dO_sb, then the X/Y buffer can be refilled for step G + 2 ----
        ptx::mbar_wait(&a_full[G & 1], (G >> 1) & 1);
        TR(0, 2, G);
        if (sb == 0 && tile_idx > 0) ptx::mbar_wait(acc_free, (tile_idx - 1) & 1);  // previous dV/dK read out
        ptx::tc_fence_after();
        TR(0, 3, G);
        if (!(dbg & 8)) {
          const uint32_t b = domn_lo0 + st * QD_STEP + brow;
          if (nks == 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16_ts_p(leader, tmem + TM_DV, buf + 32 * (ks >> 1) + 8 * (ks & 1), mk(b + ks * 128), idesc_o,
                                  (sb > 0 || ks > 0) ? 1u : 0u);
          } else {
            for (int ks = 0; ks < nks; ++ks)
              ptx::umma_bf16_ts_p(leader, tmem + TM_DV, buf + 32 * (ks >> 1) + 8 * (ks & 1), mk(b + ks * 128), idesc_o,
                                  (sb > 0 || ks > 0) ? 1u : 0u);
          }
        }
        TR(0, 4, G);
        if (G + 2 < total) {
          issue_xy(ahead, G + 2);
          advance(ahead);
        }
        TR(0, 5, G);
        // ---- dS^T block is in smem: dK_t += dS^T Q_sb, and after a pair of sub-blocks dQ_qt += dS K_t ----
        ptx::mbar_wait(&ds_full[G & 1], (G >> 1) & 1);
        ptx::tc_fence_after();
        TR(0, 6, G);
        if (!(dbg & 4)) {
          const uint32_t a = dsk_lo0 + (sb & 1) * KV_STEP, b = qmn_lo0 + st * QD_STEP + brow;
          if (nks == 4) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks)
              ptx::umma_bf16_p(leader, tmem + TM_DK, mk(a + 2 * ks), mk(b + ks * 128), idesc_o,
                               (sb > 0 || ks > 0) ? 1u : 0u);
          } else {
            for (int ks = 0; ks < nks; ++ks)
              ptx::umma_bf16_p(leader, tmem + TM_DK, mk(a + 2 * ks), mk(b + ks * 128), idesc_o,
                               (sb > 0 || ks > 0) ? 1u : 0u);
          }
        }
        if ((sb & 1) || last_sb) {
          const int qt = sb >> 1;
          const int kvalid = min(128, tokens - 128 * kt);
          const int ks5 = (kvalid + 15) >> 4;
          if (kt == 0 && qt == 0 && n > 0) {  // previous head's dQ has been read out
            ptx::mbar_wait(dq_free, (n - 1) & 1);
            ptx::tc_fence_after();
          }
          if (!(dbg & 4)) {
            const uint32_t d = tmem + TM_DQ + 64 * qt, b = kmn_lo0 + kt * KV_STEP;
            if (ks5 == 8) {
#pragma unroll
              for (int ks = 0; ks < 8; ++ks)
                ptx::umma_bf16_p(leader, d, mk(dsmn_lo + ks * 128), mk(b + ks * 128), idesc_q, (kt > 0 || ks > 0) ? 1u : 0u);
            } else {
              for (int ks = 0; ks < ks5; ++ks)
                ptx::umma_bf16_p(leader, d, mk(dsmn_lo + ks * 128), mk(b + ks * 128), idesc_q, (kt > 0 || ks > 0) ? 1u : 0u);
            }
          }
          ptx::umma_commit_p(leader, ds_free);
        }
        TR(0, 7, G);
        if (last_sb) {
          ptx::umma_commit_p(leader, kvacc_full);
          ptx::umma_commit_p(leader, &kv_empty[kt]);
          ++tile_idx;
          if (kt == nkt - 1) {
            ptx::umma_commit_p(leader, &qdo_empty[st]);
            ptx::umma_commit_p(leader, dq_full);
          }
        }
      }
    }
  } else if (warp >= F_EPI_WARP0) {
    // ======================================= accumulator read-out warps =======================================
    const int w4 = warp & 3;  // TMEM lane quarter this warp may access
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    uint8_t* my_stage = smem + OFF_STAGE + (warp - F_EPI_WARP0) * 4096;
    uint32_t nstore = 0;
    int tile_idx = 0;
    for (int n = 0; n < my_items; ++n) {
      const int item = blockIdx.x + n * gridDim.x;
      const int b = item / heads, h = item % heads;
      for (int kt = 0; kt < nkt; ++kt, ++tile_idx) {
        ptx::mbar_wait(kvacc_full, tile_idx & 1);
        ptx::tc_fence_after();
        const bool rows_ok = 128 * kt + 32 * w4 < tokens && !(dbg & 1);
        uint32_t pv0[16], pv1[16], pk0[16], pk1[16];
        if (rows_ok) {
          read_pack(lane_addr + TM_DV, 1.0f, pv0);
          read_pack(lane_addr + TM_DV + 32, 1.0f, pv1);
          read_pack(lane_addr + TM_DK, scale, pk0);
          read_pack(lane_addr + TM_DK + 32, scale, pk1);
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive(acc_free);
        if (rows_ok) {
          const int row0 = 128 * kt + 32 * w4;
          stage_store_32(my_stage + (nstore++ & 1) * 2048, pv0, &tmOut, 2 * D + h * HD, row0, b, lane);
          stage_store_32(my_stage + (nstore++ & 1) * 2048, pv1, &tmOut, 2 * D + h * HD + 32, row0, b, lane);
          stage_store_32(my_stage + (nstore++ & 1) * 2048, pk0, &tmOut, D + h * HD, row0, b, lane);
          stage_store_32(my_stage + (nstore++ & 1) * 2048, pk1, &tmOut, D + h * HD + 32, row0, b, lane);
        }
      }
      ptx::mbar_wait(dq_full, n & 1);
      ptx::tc_fence_after();
      const bool q0_ok = 32 * w4 < tokens && !(dbg & 1), q1_ok = 128 + 32 * w4 < tokens && !(dbg & 1);
      uint32_t pa0[16], pa1[16], pc0[16], pc1[16];
      if (q0_ok) {
        read_pack(lane_addr + TM_DQ, scale, pa0);
        read_pack(lane_addr + TM_DQ + 32, scale, pa1);
      }
      if (q1_ok) {
        read_pack(lane_addr + TM_DQ + 64, scale, pc0);
        read_pack(lane_addr + TM_DQ + 96, scale, pc1);
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(dq_free);
      if (q0_ok) {
        stage_store_32(my_stage + (nstore++ & 1) * 2048, pa0, &tmOut, h * HD, 32 * w4, b, lane);
        stage_store_32(my_stage + (nstore++ & 1) * 2048, pa1, &tmOut, h * HD + 32, 32 * w4, b, lane);
      }
      if (q1_ok) {
        stage_store_32(my_stage + (nstore++ & 1) * 2048, pc0, &tmOut, h * HD, 128 + 32 * w4, b, lane);
        stage_store_32(my_stage + (nstore++ & 1) * 2048, pc1, &tmOut, h * HD + 32, 128 + 32 * w4, b, lane);
      }
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
    __syncwarp();
  } else {
    // ======================================= element-wise warps =======================================
    // Software-pipelined: the X / Y rows of step G + 1 are requested from TMEM right after step G's P^T has been
    // handed to the MMA warp, so their latency hides behind the dS^T smem write of step G.
    const int w4 = warp & 3, half = warp >> 2;
    const uint32_t lane_addr = tmem + (static_cast<uint32_t>(w4 * 32) << 16);
    const int row = w4 * 32 + lane;  // key row inside the 128-key tile
    const uint32_t ds_row = ptx::smem_u32(smem + OFF_DS) + row * 128;
    const int total = my_items * nsteps;
    int pairs = 0;
    int tr_idx = (warp == 0 && lane == 0) ? 0 : 1 << 20;
    auto my_width = [&](int sb) { return max(0, min(32, min(64, ncols - 64 * sb) - 32 * half)); };
    auto tile_active = [&](int kt) { return 32 * w4 < ((min(128, tokens - 128 * kt) + 15) & ~15); };  // rows the dQ MMA reads
    uint32_t x[32], y[32];
    int n = 0, kt = 0, sb = 0;
    bool work = false;
    if (total > 0) {
      ptx::mbar_wait(&qdo_full[0], 0);  // statistics of the first head are visible
      ptx::mbar_wait(&xy_full[0], 0);
      ptx::tc_fence_after();
      work = tile_active(0) && my_width(0) > 0;
      if (work) {
        ptx::tmem_ld_32x32b_x32(lane_addr + 32 * half, x);
        ptx::tmem_ld_32x32b_x32(lane_addr + 64 + 32 * half, y);
      }
    }
#pragma unroll 1
    for (int G = 0; G < total; ++G) {
      const int st = n & 1;
      const uint32_t sl = ptx::smem_u32(smem + OFF_STATS + st * 2048) + (64 * sb + 32 * half) * 4;  // lse2 of my columns
      const uint32_t sd = sl + 1024;                                                                // delta
      const int my_w = my_width(sb);
      const uint32_t buf = lane_addr + (G & 1) * TM_BUF;
      uint32_t o2[16];
      TR(1024, 12, G);
      if (work) {
        ptx::tmem_ld_wait();
        TR(1024, 13, G);
        uint32_t o1[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (dbg & 2) {
            o1[2 * j] = x[4 * j] ^ y[4 * j + 1];
            o1[2 * j + 1] = x[4 * j + 2] ^ y[4 * j + 3];
            o2[2 * j] = x[4 * j + 1] ^ y[4 * j];
            o2[2 * j + 1] = x[4 * j + 3] ^ y[4 * j + 2];
          } else if (j < 4 || my_w == 32) {
            float4 l4, d4;
            if (dbg & 64) {  // timing experiment: no statistics loads (results are garbage)
              l4 = make_float4(1.f, 2.f, 3.f, 4.f);
              d4 = make_float4(0.5f, 0.25f, 0.125f, 0.0625f);
            } else {
              l4 = lds128(sl + 16 * j);
              d4 = lds128(sd + 16 * j);
            }
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
        TR(1024, 14, G);
        // bf16 P^T (A operand of the dV MMA) overwrites this warp's own X columns
        if (my_w == 32) ptx::tmem_st_32x32b_x16(buf + 32 * half, o1);
        else ptx::tmem_st_32x32b_x8(buf + 32 * half, o1);
        ptx::tmem_st_wait();
      }
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&a_full[G & 1]);
      TR(1024, 15, G);
      // ---- request the next step's X / Y rows ----
      const bool pair_done = (sb & 1) || sb == nsb - 1;
      const int cur_sb = sb;
      const bool cur_work = work;
      if (++sb == nsb) {
        sb = 0;
        if (++kt == nkt) {
          kt = 0;
          ++n;
        }
      }
      if (G + 1 < total) {
        if (kt == 0 && sb == 0) ptx::mbar_wait(&qdo_full[n & 1], (n >> 1) & 1);  // next head's statistics
        ptx::mbar_wait(&xy_full[(G + 1) & 1], ((G + 1) >> 1) & 1);
        ptx::tc_fence_after();
        work = tile_active(kt) && my_width(sb) > 0;
        if (work) {
          const uint32_t nbuf = lane_addr + ((G + 1) & 1) * TM_BUF;
          ptx::tmem_ld_32x32b_x32(nbuf + 32 * half, x);
          ptx::tmem_ld_32x32b_x32(nbuf + 64 + 32 * half, y);
        }
      }
      // ---- dS^T block of step G: row = key (128 B = 64 queries), 16-byte chunks XOR-swizzled by row & 7.  It is the
      // K-major A operand of the dK MMA and, read transposed, the MN-major A operand of the dQ MMA. ----
      if (cur_work) {
        if (!(cur_sb & 1) && pairs > 0) ptx::mbar_wait(ds_free, (pairs - 1) & 1);
        TR(1024, 16, G);
        const uint32_t blk = ds_row + (cur_sb & 1) * KV_TILE;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((i < 2 || my_w == 32) && !(dbg & 4)) {
            const uint32_t addr = blk + (((4 * half + i) ^ (row & 7)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(o2[4 * i]), "r"(o2[4 * i + 1]),
                         "r"(o2[4 * i + 2]), "r"(o2[4 * i + 3])
                         : "memory");
          }
        }
        ptx::fence_proxy_async_smem();
      }
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ds_full[G & 1]);
      TR(1024, 17, G);
      if (pair_done) ++pairs;
    }
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

// timing experiments: device buffer of 2 * 2048 * 2 int64 receiving CTA 0's event timeline (with VITATK_ATTN_DBG & 32)
int attention_bwd_set_trace(long long* dev_buf) {
#ifdef VITATK_DBG_KERNELS
  VITATK_CUDA_OK(cudaMemcpyToSymbol(g_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
#else
  (void)dev_buf;
  set_error("attention_bwd_set_trace: build with -DVITATK_DBG_KERNELS (VITATK_DBG_BUILD=1) for the in-kernel timeline");
  return 1;
#endif
}

int attention_bwd_fused(const AttnBwdPlan* p, cudaStream_t stream, bool compute_delta) {
  static PerDeviceOnce once;
  if (once.need())
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, F_SMEM));
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
  if (compute_delta && !(dbg & 16) && attention_delta(p, stream)) return 1;
  VITATK_CUDA_OK(launch_pdl(attn_bwd_fused_kernel, dim3(items < sms ? items : sms), dim3(F_THREADS), F_SMEM, stream, 1,
                            p->tmQKV128, p->tmQKV208, p->tmDO208, p->tmDqkv32, p->lse2, p->delta, p->tokens, p->heads,
                            items, sl2, scale, dbg));
  return 0;
}

}  // namespace vitatk

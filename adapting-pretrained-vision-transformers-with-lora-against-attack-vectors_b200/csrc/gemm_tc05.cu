// Persistent, warp-specialised tcgen05 GEMM for sm_100a with the LoRA product and the
// bias / residual / GELU / GELU' / row-table epilogues fused into the accumulator tile.
//
//   out[M,N] = epi( A[M,K] * B[N,K]^T + T[M,64*nkb] * LB[N,64*nkb]^T )
//
// Replaces (reference call sites): every nn.Linear of HF ViTLayer (HF modeling_vit.py:216-218,262,
// 290-298,305-311) plus the peft LoRA branch configured at train_loras.py:79-95, forward and
// input-gradient backward (dX = dY*W + s*(dY*B)*A).
//
// Structure (one CTA per SM, 320 threads):
//   warps 0-7 : epilogue   TMEM -> registers (tcgen05.ld) -> fused math -> swizzled smem -> TMA store
//               (two warps per TMEM lane quarter, each taking one column half of the tile; residual /
//               multiplier rows are software-pipelined one 32-column chunk ahead in registers)
//   warp  8   : TMA producer (cp.async.bulk.tensor, 128B swizzle, mbarrier complete_tx); owns TMEM alloc
//   warp  9   : MMA issuer   (one thread, tcgen05.mma cta_group::1 kind::f16, M=128 N=BN K=16)
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM accumulator double buffer full/empty
// (MMA <-> epilogue), static persistent tile scheduler (tile = blockIdx.x + i*gridDim.x).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_fp16.h>
#include <type_traits>

#include "ptx.cuh"
#include "vitatk_internal.h"

namespace vitatk {

static constexpr int BM = 128;
static constexpr int BK = 64;  // 64 bf16 = 128 B = one swizzle row
static constexpr int NUM_EPI_WARPS = 8;  // single-CTA kernel and pair kernel with EW == 1; the pair kernel with EW == 2 has 16
// warp roles: epilogue warps 0..NEPI-1, then the TMA producer warp, then the MMA issuer warp
__host__ __device__ constexpr int gemm_epi_warps(bool two, int ew) { return (two && ew == 2) ? 16 : NUM_EPI_WARPS; }
// (+ in the pair kernel, a "signal" warp that publishes finished T-tiles with a gpu-scope release, so that no epilogue,
// producer or MMA thread ever blocks on that fence)
__host__ __device__ constexpr int gemm_threads(bool two, int ew) { return (gemm_epi_warps(two, ew) + (two ? 3 : 2)) * 32; }
static constexpr int CHUNK = 32;                  // epilogue column granule = one tcgen05.ld.32x32b.x32
static constexpr int STAGE_OUT_BYTES = 32 * 64;   // one warp's 32-row x 32-col bf16 store tile (64 B swizzle)
static constexpr int SLAB_BYTES = 128 * 128;      // pair kernel: 128-row x 64-col bf16 slab (128 B swizzle)

// TWO = CTA-pair variant (cta_group::2): the pair computes a 256 x BN tile, each CTA stages its 128 rows of A and
// HALF of the B tile per k-block (32 KB instead of 48 KB at BN = 256), which is what relieves the L2 -> SM operand path.
template <int BN, bool TWO = false>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_ROWS = TWO ? BN / 2 : BN;
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = TWO ? 5 : ((BN == 256) ? 4 : (BN == 192 ? 4 : (BN == 128 ? 6 : 8)));
  static constexpr int TMEM_COLS = (2 * BN <= 32) ? 32 : (2 * BN <= 64 ? 64 : (2 * BN <= 128 ? 128 : (2 * BN <= 256 ? 256 : 512)));
  // epilogue staging: single kernel = 2 x (32 rows x 32 cols) per warp; pair kernel = 2 column groups x 2 slabs of
  // (128 rows x 64 cols, 128B swizzle) shared by the four warps of a group
  static constexpr int OUT_BYTES = TWO ? 2 * 2 * SLAB_BYTES : NUM_EPI_WARPS * 2 * STAGE_OUT_BYTES;
  static constexpr int BAR_BYTES = 256;
  static constexpr int XCH_BYTES = TWO ? 0 : 1024;  // single-CTA kernel: 128 x (mean, M2) hand-over between the two statistics groups
  static constexpr int SMEM_BYTES = 1024 /*align slack*/ + STAGES * STAGE_BYTES + OUT_BYTES + BAR_BYTES + XCH_BYTES;
  static_assert(STAGES % 2 == 0 || TWO, "the statistics groups of the single-CTA kernel own stages by parity");
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of shared memory per CTA");
};

struct GemmKernelArgs {
  int M, N, K;
  int lora_nkb, lora_ksteps, lora_group_cols;
  GemmEpilogue epi;
  int reverse_m;  // 1: walk the M-blocks from the last to the first (the input was just written in ascending order by the
                  // previous kernel, so its tail is still in L2)
  // T-tiles (pair kernel; see GemmTT in vitatk_internal.h).  tt_n = 0: off.
  int tt_n;
  bf16* tt_out;
  int tt_ld;
  const float* tt_bias;
  unsigned int* tt_flags;
  int a_f16, out_f16, res_f16;  // 16-bit formats of the A and B operands / output / EPI_RESIDUAL input (0 bf16, 1 fp16)
  int dbg_rt;    // run-time experiment switches that exist in the product build (VITATK_GEMM_RT): 1 = L2 prefetch of T-tiles
  int gelu_f32;  // GELU in the pair epilogue: 1 = fp32 Abramowitz-Stegun (VITATK_GELU=f32), otherwise the fp32 2^P fit
  int dbg;  // timing experiments (DBG instantiation only, VITATK_GEMM_DBG): 1 no aux loads, 2 no stores,
            // 4 no TMA loads after the first ring fill, 8 no MMA issue, 16 epilogue skips TMEM reads and math,
            // 32 no staging writes, 64 no bias, 128 no group barriers, 256 no GELU math, 512 epilogue timeline,
            // 1024 no L1 prefetch of the epilogue constants, 2048 per-row scale only (no per-column constants)
};

// Exact-erf GELU and its derivative from ONE exponential and ONE reciprocal (Abramowitz-Stegun 7.1.26,
// |erf error| < 1.5e-7, far below the bf16 output resolution):
//   x = |u|/sqrt2, t = 1/(1+p x), e = exp(-x^2) = exp(-u^2/2), erfc(x) = poly(t) e
//   Phi(u) = u<0 ? erfc/2 : 1 - erfc/2 ;  gelu = u Phi ;  gelu' = Phi + u e / sqrt(2 pi)
__device__ __forceinline__ void gelu_and_grad(float u, float& g, float& d) {
  const float ax = fabsf(u) * 0.70710678118654752f;
  float t, e;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.0f)));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-1.4426950408889634f * ax * ax));
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float half_erfc = 0.5f * p * t * e;
  const float phi = u < 0.f ? half_erfc : 1.0f - half_erfc;
  g = u * phi;
  d = fmaf(u * e, 0.3989422804014327f, phi);
}
// fp32 variant of the 2^P fit below: 12 FMA/ALU-pipe instructions + 2 MUFU.EX2 per element and NO f16 conversions.
// On B200 HFMA2 issues at half the FFMA rate (scripts/micro/alu_rate.cu: 64 vs 125 lanes/clk/SM) and ex2.f16x2 is two
// MUFU ops, so packed half buys no throughput; in fp32 the same degree-4 fit is also more accurate
// (|gelu error| <= 9.2e-5, |gelu' error| <= 7.2e-4 over [-8, 8]).
__device__ __forceinline__ void gelu_and_grad_p(float u, float& g, float& d) {
  const float ax = fminf(fabsf(u), 4.25f);
  float p = fmaf(0.00238781f, ax, -0.03508832f);
  p = fmaf(p, ax, -0.48539043f);
  p = fmaf(p, ax, -1.13599843f);
  p = fmaf(p, ax, -1.00205332f);
  float e, pb;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(p));                                    // Phi(-|u|) < 0.5
  const float hm = 0.5f - e;
  const float phi = 0.5f + __uint_as_float(__float_as_uint(hm) | (__float_as_uint(u) & 0x80000000u));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(pb) : "f"(fmaf(u * u, -0.72134752f, -1.32574806f)));  // pdf(u)
  g = u * phi;
  d = fmaf(u, pb, phi);
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi);
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ uint32_t pack_f16x2(float lo, float hi) {
  __half2 v = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void unpack_f16x8(const uint4& q, float* f) {
  const __half2* p = reinterpret_cast<const __half2*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __half22float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
__device__ __forceinline__ void unpack_bf16x8(const uint4& q, float* f) {
  const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&q);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float2 t = __bfloat1622float2(p[i]);
    f[2 * i] = t.x;
    f[2 * i + 1] = t.y;
  }
}
// one warp: 32 rows x 32 bf16 columns from registers -> 64B-swizzled smem tile -> TMA store
__device__ __forceinline__ void stage_and_store(uint8_t* stage, const uint32_t (&src)[16], const CUtensorMap* tm, int c0,
                                                int c1, int lane) {
  if (lane == 0) ptx::tma_store_wait_read<1>();  // the buffer used two stores ago has been read out
  __syncwarp();
  const uint32_t row_base = ptx::smem_u32(stage) + lane * 64;
  const uint32_t sw = (lane >> 1) & 3;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t addr = row_base + ((j ^ sw) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(src[4 * j]), "r"(src[4 * j + 1]),
                 "r"(src[4 * j + 2]), "r"(src[4 * j + 3])
                 : "memory");
  }
  ptx::fence_proxy_async_smem();
  __syncwarp();
  if (lane == 0) {
    ptx::tma_store_2d(tm, stage, c0, c1);
    ptx::tma_store_commit();
  }
}

// in-kernel timeline (DBG instantiation, dbg & 512): CTA 0, lane 0 of epilogue warps 0 and 5 write clock64() per event
__device__ long long* g_gemm_trace = nullptr;

template <int BN, bool DBG, bool TWO, int EW = 1>
__global__ void __launch_bounds__(gemm_threads(TWO, EW), 1)
gemm_tc05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmLA, const __grid_constant__ CUtensorMap tmLB,
                 const __grid_constant__ CUtensorMap tmOut, const __grid_constant__ CUtensorMap tmOut2,
                 const __grid_constant__ CUtensorMap tmAux, const __grid_constant__ CUtensorMap tmTB,
                 const GemmKernelArgs args) {
  using Cfg = GemmCfg<BN, TWO>;
  // CTA pair: rank 0 is the leader (issues the M = 256 MMAs); unit = CTA (single) or cluster (pair)
  const uint32_t rank = TWO ? ptx::cluster_ctarank() : 0u;
  const int unit = TWO ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int num_units = TWO ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  constexpr int TILE_M = TWO ? 2 * BM : BM;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NEPI = gemm_epi_warps(TWO, EW);
  constexpr int TMA_WARP = NEPI;
  constexpr int MMA_WARP = NEPI + 1;
  extern __shared__ uint8_t smem_raw[];
  // 1024 B alignment: required by the 128B swizzle pattern shared by TMA and the UMMA descriptors
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_stage = smem;
  uint8_t* smem_out = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_out + Cfg::OUT_BYTES);
  uint8_t* smem_xch = smem_out + Cfg::OUT_BYTES + Cfg::BAR_BYTES;
  (void)smem_xch;
  uint64_t* full_bar = bars;                   // [STAGES]
  uint64_t* empty_bar = bars + STAGES;         // [STAGES]
  uint64_t* tmem_full = bars + 2 * STAGES;     // [2]
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;  // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  uint64_t* aux_bar = bars + 2 * STAGES + 5;   // [2 groups][2 slabs] residual / multiplier slab landed (pair kernel)
  uint64_t* tt_done = bars + 2 * STAGES + 9;   // the T-tile's rows of this CTA are in global memory (pair kernel)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int tiles_m = (args.M + TILE_M - 1) / TILE_M;
  const int tiles_n = args.N / BN;
  const int num_tiles = tiles_m * tiles_n;
  const int main_kb = args.K / BK;
  const int num_kb = main_kb + args.lora_nkb;
  auto mblock = [&](int tile_) { return args.reverse_m ? tiles_m - 1 - tile_ / tiles_n : tile_ / tiles_n; };
  // T-tiles (pair kernel, args.tt_n > 0): every M-block gets one extra work item that computes T = A * TB^T for the block
  // (same A rows, the adapter's down-projection as B operand, full K).  Schedule: the T-tile of block m runs ONE WAVE of
  // output tiles ahead of the block's own output tiles -- units 0 .. tt_pro-1 start with the T-tiles of the first wave's
  // blocks, and the unit that owns output tile (m, m % tiles_n) first computes the T-tile of block m + tt_shift -- so its
  // rows are published long before a LoRA k-block needs them, while the A block is still in L2 when the output tiles
  // stream it again.  Every role derives the same item sequence from (unit, tile) alone.
  const bool tt_on = TWO && args.tt_n > 0;
  const int tt_shift = (num_units + tiles_n - 1) / tiles_n;        // M-blocks per wave of output tiles
  const int tt_pro = tt_on ? min(tt_shift, tiles_m) : 0;
  auto ttile_of = [&](int tile_) -> int {  // linear block index whose T-tile precedes output tile `tile_`, or -1
    if (!tt_on) return -1;
    const int ml = tile_ / tiles_n;
    if (tile_ - ml * tiles_n != ml % tiles_n) return -1;
    const int mt = ml + tt_shift;
    return mt < tiles_m ? mt : -1;
  };
  auto block_row0 = [&](int mlin) {  // first row of this CTA's 128 rows of (linear) block mlin
    return (args.reverse_m ? tiles_m - 1 - mlin : mlin) * TILE_M + static_cast<int>(rank) * BM;
  };
  uint32_t* tt_ready = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 10);  // output tiles of this unit whose T rows are acquired

  if (warp == MMA_WARP && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      ptx::mbar_init(&full_bar[s], 1);  // pair: the leader's producer expects the bytes of BOTH CTAs' loads
      ptx::mbar_init(&empty_bar[s], (!TWO && args.epi.stats_out != nullptr) ? 5 : 1);  // + the four row-statistics warps
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&tmem_full[b], 1);
      ptx::mbar_init(&tmem_empty[b], TWO ? 2 * NEPI : NEPI);  // pair: both CTAs' epilogue warps
    }
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&aux_bar[i], 1);
    ptx::mbar_init(tt_done, NEPI / 2);  // one arrival per warp of column group 0
    *tt_ready = 0u;
    ptx::fence_mbar_init();
  }
  if (warp == TMA_WARP) {
    if (lane == 0) {
      ptx::prefetch_tmap(&tmA);
      ptx::prefetch_tmap(&tmB);
      ptx::prefetch_tmap(&tmOut);
      if (TWO && (args.epi.mode == EPI_RESIDUAL || args.epi.mode == EPI_MUL || args.epi.mode == EPI_ROWDOT))
        ptx::prefetch_tmap(&tmAux);
      if (args.lora_nkb > 0) {
        ptx::prefetch_tmap(&tmLA);
        ptx::prefetch_tmap(&tmLB);
      }
      if (tt_on) ptx::prefetch_tmap(&tmTB);
    }
    if constexpr (TWO) {
      ptx::tmem_alloc_2cta(tmem_slot, Cfg::TMEM_COLS);
      ptx::tmem_relinquish_2cta();
    } else {
      ptx::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      ptx::tmem_relinquish();
    }
  }
  ptx::tc_fence_before();
  if constexpr (TWO) ptx::cluster_sync();  // barrier inits of both CTAs are visible before any remote arrive / TMA signal
  else __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();               // everything above overlapped the previous kernel; global memory is touched only below
  pdl_launch_dependents();

  if (warp == TMA_WARP) {
    // ================================= TMA producer (converged warp, elected lane issues) =================================
    const uint32_t leader = ptx::elect_leader();
    uint32_t cnt = 0;
    uint32_t n_main = 0;  // output tiles this unit has started
    auto load_ttile = [&](int mlin) {  // A rows of block mlin x the adapter's down-projection (tt_n / 2 rows of TB per CTA)
      const uint32_t tb_bytes = static_cast<uint32_t>(args.tt_n) * 64u;  // (tt_n / 2 rows) x 128 B
      const uint32_t fb0 = ptx::mapa_shared(ptx::smem_u32(&full_bar[0]), 0);
      const int tm0 = block_row0(mlin);
      for (int kb = 0; kb < main_kb; ++kb, ++cnt) {
        const int s = cnt % STAGES;
        ptx::mbar_wait(&empty_bar[s], ((cnt / STAGES) & 1) ^ 1);
        uint8_t* sa = smem_stage + s * Cfg::STAGE_BYTES;
        const uint32_t fb = fb0 + s * 8;
        if (rank == 0) ptx::mbar_arrive_expect_tx_p(leader, &full_bar[s], 2 * (Cfg::A_BYTES + tb_bytes));
        ptx::tma_load_2d_2cta_p(leader, sa, &tmA, fb, kb * BK, tm0);
        ptx::tma_load_2d_2cta_p(leader, sa + Cfg::A_BYTES, &tmTB, fb, kb * BK, static_cast<int>(rank) * (args.tt_n >> 1));
      }
    };
    if constexpr (TWO) {
      if (unit < tt_pro) load_ttile(unit);
    }
    for (int tile = unit; tile < num_tiles; tile += num_units, ++n_main) {
      const int m0 = mblock(tile) * TILE_M + static_cast<int>(rank) * BM;              // this CTA's 128 rows of A
      const int n0 = (tile % tiles_n) * BN;
      const int nb0 = n0 + static_cast<int>(rank) * Cfg::B_ROWS;                     // pair: this CTA's half of B
      const int tcol0 = args.lora_group_cols > 0 ? (n0 / args.lora_group_cols) * BK : 0;
      int pf_row0 = -1;  // rows of the T-tile that follows this output tile in the unit's sequence
      if constexpr (TWO) {
        const int mt = ttile_of(tile);
        if (mt >= 0) load_ttile(mt);
        // Experiment (VITATK_GEMM_RT=1, off by default): pull the A rows of the unit's NEXT T-tile into L2 while this
        // output tile runs.  Measured on B200 it makes T-tiles SLOWER (fc2: 70 instead of 50 us per launch): a T-tile is
        // bound by the ring depth x the loaded L2 -> SM latency (~680 clk per k-block whatever its size), not by HBM, and
        // the prefetch traffic only delays the output tile it rides on.
        if (tt_on && tile + num_units < num_tiles && (args.dbg_rt & 1)) {
          const int nmt = ttile_of(tile + num_units);
          if (nmt >= 0) pf_row0 = block_row0(nmt);
        }
      }
      for (int kb = 0; kb < num_kb; ++kb, ++cnt) {
        const int s = cnt % STAGES;
        if (TWO && pf_row0 >= 0 && kb < main_kb) ptx::tma_prefetch_2d_p(leader, &tmA, kb * BK, pf_row0);
        if (TWO && tt_on && kb == main_kb) {
          // The LoRA k-block reads T rows produced by a T-tile (usually of another unit).  The publisher warp has
          // acquired the block's flag ahead of time and counts the output tiles that may proceed; all that is left here
          // is a shared-memory wait and ordering this thread's TMA (async proxy) read behind the acquired writes.
          uint32_t spins = 0;
          while (ptx::ld_acquire_cta_shared(tt_ready) <= n_main) {
            if (++spins > (1u << 30)) __trap();  // a T-tile that never arrives must not hang the GPU
          }
          ptx::fence_proxy_async_global();
        }
        const uint32_t ph = (cnt / STAGES) & 1;
        ptx::mbar_wait(&empty_bar[s], ph ^ 1);
        uint8_t* sa = smem_stage + s * Cfg::STAGE_BYTES;
        uint8_t* sb = sa + Cfg::A_BYTES;
        if constexpr (TWO) {
          const uint32_t fb = ptx::mapa_shared(ptx::smem_u32(&full_bar[s]), 0);  // the leader CTA's full barrier
          if (DBG && (args.dbg & 4) && cnt >= STAGES) {
            if (rank == 0) ptx::mbar_arrive_p(leader, &full_bar[s]);
            continue;
          }
          // Only the leader arrives (no remote round trip per k-block).  The peer's complete_tx may land first: the
          // tx-count is signed and the phase cannot complete before the leader's pending arrival.
          if (rank == 0) ptx::mbar_arrive_expect_tx_p(leader, &full_bar[s], 2 * Cfg::STAGE_BYTES);
          if (kb < main_kb) {
            ptx::tma_load_2d_2cta_p(leader, sa, &tmA, fb, kb * BK, m0);
            ptx::tma_load_2d_2cta_p(leader, sb, &tmB, fb, kb * BK, nb0);
          } else {
            const int j = kb - main_kb;
            ptx::tma_load_2d_2cta_p(leader, sa, &tmLA, fb, tcol0 + j * BK, m0);
            ptx::tma_load_2d_2cta_p(leader, sb, &tmLB, fb, j * BK, nb0);
          }
        } else {
          if (DBG && (args.dbg & 4) && cnt >= STAGES) {
            ptx::mbar_arrive_p(leader, &full_bar[s]);
            continue;
          }
          ptx::mbar_arrive_expect_tx_p(leader, &full_bar[s], Cfg::STAGE_BYTES);
          if (kb < main_kb) {
            ptx::tma_load_2d_p(leader, sa, &tmA, &full_bar[s], kb * BK, m0);
            ptx::tma_load_2d_p(leader, sb, &tmB, &full_bar[s], kb * BK, n0);
          } else {
            const int j = kb - main_kb;
            ptx::tma_load_2d_p(leader, sa, &tmLA, &full_bar[s], tcol0 + j * BK, m0);
            ptx::tma_load_2d_p(leader, sb, &tmLB, &full_bar[s], j * BK, n0);
          }
        }
      }
    }
  } else if (warp == MMA_WARP) {
    // ================================= MMA issuer (converged warp, elected lane issues) =================================
    const uint32_t leader = ptx::elect_leader();
    constexpr uint32_t idesc = ptx::make_idesc_bf16(TILE_M, BN);          // A = T (bf16): the LoRA k-blocks
    // a_f16: A AND B of the main k-blocks are IEEE fp16 (a_format / b_format fields: 0 = F16, 1 = BF16).  kind::f16 wants
    // both operands in the same format -- a mixed fp16 x bf16 descriptor is an illegal instruction on sm_100a (measured)
    const uint32_t idesc_a = args.a_f16 ? (idesc & ~((7u << 7) | (7u << 10))) : idesc;
    uint32_t cnt = 0;
    uint32_t it = 0;
    auto mma_ttile = [&]() {  // T-tile: M = 256, N = tt_n, K = the GEMM's K, into the first columns of accumulator it & 1
      const uint32_t tbuf = it & 1;
      ptx::mbar_wait(&tmem_empty[tbuf], ((it >> 1) & 1) ^ 1);
      ptx::tc_fence_after();
      const uint32_t idesc_t =
          ptx::make_idesc_bf16(TILE_M, static_cast<uint32_t>(args.tt_n)) & (args.a_f16 ? ~((7u << 7) | (7u << 10)) : ~0u);
      for (int kb = 0; kb < main_kb; ++kb, ++cnt) {
        const int s = cnt % STAGES;
        ptx::mbar_wait(&full_bar[s], (cnt / STAGES) & 1);
        ptx::tc_fence_after();
        const uint32_t sa = ptx::smem_u32(smem_stage + s * Cfg::STAGE_BYTES);
        const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
        const uint64_t bdesc = ptx::make_smem_desc_sw128(sa + Cfg::A_BYTES);
#pragma unroll
        for (int k = 0; k < BK / 16; ++k)
          ptx::umma_bf16_2cta_p(leader, tmem_base + tbuf * BN, adesc + 2 * k, bdesc + 2 * k, idesc_t,
                                (kb > 0 || k > 0) ? 1u : 0u);
        ptx::umma_commit_2cta_mc_p(leader, &empty_bar[s], 3);
      }
      ptx::umma_commit_2cta_mc_p(leader, &tmem_full[tbuf], 3);
      ++it;
    };
    if (!TWO || rank == 0) {
      if constexpr (TWO) {
        if (unit < tt_pro) mma_ttile();
      }
      for (int tile = unit; tile < num_tiles; tile += num_units, ++it) {
        if constexpr (TWO) {
          if (ttile_of(tile) >= 0) mma_ttile();
        }
        const uint32_t buf = it & 1;
        const uint32_t use = it >> 1;
        // the epilogue warps (of both CTAs of a pair) have drained this accumulator
        ptx::mbar_wait(&tmem_empty[buf], (use & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tmem_acc = tmem_base + buf * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++cnt) {
          const int s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1;
          ptx::mbar_wait(&full_bar[s], ph);
          ptx::tc_fence_after();
          const uint32_t sa = ptx::smem_u32(smem_stage + s * Cfg::STAGE_BYTES);
          const uint32_t sb = sa + Cfg::A_BYTES;
          const uint64_t adesc = ptx::make_smem_desc_sw128(sa);
          const uint64_t bdesc = ptx::make_smem_desc_sw128(sb);
          const int ksteps = (kb < main_kb) ? BK / 16 : args.lora_ksteps;
          if (!(DBG && (args.dbg & 8))) {
            if (kb < main_kb) {
#pragma unroll
              for (int k = 0; k < BK / 16; ++k) {  // +16 bf16 = 32 B along K inside the 128 B swizzle row: +2 in (addr>>4)
                if constexpr (TWO)
                  ptx::umma_bf16_2cta_p(leader, tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc_a, (kb > 0 || k > 0) ? 1u : 0u);
                else
                  ptx::umma_bf16_p(leader, tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc_a, (kb > 0 || k > 0) ? 1u : 0u);
              }
            } else {
              for (int k = 0; k < ksteps; ++k) {
                if constexpr (TWO)
                  ptx::umma_bf16_2cta_p(leader, tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
                else
                  ptx::umma_bf16_p(leader, tmem_acc, adesc + 2 * k, bdesc + 2 * k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
              }
            }
          }
          // smem slot reusable (in both CTAs of a pair) once these MMAs retire
          if constexpr (TWO) ptx::umma_commit_2cta_mc_p(leader, &empty_bar[s], 3);
          else ptx::umma_commit_p(leader, &empty_bar[s]);
        }
        // accumulator complete
        if constexpr (TWO) ptx::umma_commit_2cta_mc_p(leader, &tmem_full[buf], 3);
        else ptx::umma_commit_p(leader, &tmem_full[buf]);
      }
    }
  } else if (TWO && warp == MMA_WARP + 1) {
    // ================================= T-tile publisher (pair kernel) =================================
    // Waits until column group 0 has stored this CTA's rows of a T-tile, then publishes them with a gpu-scope release
    // on the M-block's flag.  A dedicated warp, so that the fence (which waits for the stores to be visible device-wide)
    // never stalls the epilogue, the producer or the MMA issuer.
    // It also acquires, ahead of the producer, the flags of the blocks this unit's output tiles belong to (tt_ready counts
    // the tiles that may load their LoRA k-block) and keeps the flag array self-resetting: the last of a block's
    // 2 * tiles_n consumers (tiles_n output tiles x 2 CTAs) zeroes the block's pair of counters for the next launch.
    if (tt_on) {
      int pt = unit;                             // tile cursor of the publish side
      int pblk = unit < tt_pro ? unit : -1;      // block of the next T-tile this CTA publishes
      auto seek = [&]() {
        while (pblk < 0 && pt < num_tiles) {
          pblk = ttile_of(pt);
          pt += num_units;
        }
      };
      seek();
      uint32_t npub = 0, nacq = 0;
      int at = unit;                             // next output tile whose block flag is to be acquired
      unsigned int pend_old = 0;                 // (lane 0) result of the previous consumer count, examined one round later
      int pend_blk = -1;
      auto settle = [&]() {
        if (lane == 0 && pend_blk >= 0 && pend_old == 2u * tiles_n - 1u) {
          args.tt_flags[tiles_m + pend_blk] = 0u;
          args.tt_flags[pend_blk] = 0u;
        }
      };
      uint32_t idle = 0;
      while (pblk >= 0 || at < num_tiles) {
        if (++idle > (1u << 25)) __trap();  // (every poll is a device-memory round trip: tens of seconds without progress)
        if (pblk >= 0 && ptx::mbar_try_wait(tt_done, npub & 1)) {
          idle = 0;
          ++npub;
          __threadfence();
          if (lane == 0) ptx::red_release_gpu_add(args.tt_flags + pblk, 1u);
          __syncwarp();
          pblk = -1;
          seek();
        }
        if (at < num_tiles) {
          const int ml = at / tiles_n;
          if (ptx::ld_acquire_gpu(args.tt_flags + ml) >= 2u) {  // both CTAs of the producing pair have published
            idle = 0;
            settle();
            if (lane == 0) {
              pend_old = atomicAdd(args.tt_flags + tiles_m + ml, 1u);
              pend_blk = ml;
            }
            ++nacq;
            __syncwarp();
            if (lane == 0) ptx::st_release_cta_shared(tt_ready, nacq);
            at += num_units;
          }
        }
      }
      settle();
    }
  } else if constexpr (TWO) {
    // ================================= slab epilogue (pair kernel), warps 0..NEPI-1 =================================
    // Column group g owns columns [128 g, 128 g + 128) of the tile as two 64-column slabs; its warps fill one 128-row x
    // 64-column slab in 128B-swizzled smem and ONE thread issues ONE TMA store per slab (4 per tile instead of 32 small
    // ones, full 128-byte lines).  Residual / multiplier slabs are TMA-loaded into the same buffer one slab ahead and
    // combined in place, so no strided per-lane global loads remain.
    // EW = warps per TMEM lane quarter inside a group: 1 -> four warps per group, every thread owns a full 64-column
    // slab row; 2 -> eight warps per group, warp (q, hh) owns columns [32 hh, 32 hh + 32) of the slab row (half the
    // registers per thread and four instead of two warps per scheduler to hide the epilogue's dependent-issue latency).
    constexpr int NC = 64 / EW;            // slab columns per thread
    constexpr int NP = NC / 2;             // packed bf16x2 registers per thread per slab
    constexpr int NQ = NC / 8;             // 16-byte chunks per thread per slab row
    const int q = warp & 3;
    const int hh = (EW == 2) ? ((warp >> 2) & 1) : 0;
    const int g = (EW == 2) ? (warp >> 3) : (warp >> 2);
    const int trow = q * 32 + lane;  // row inside this CTA's 128-row block
    const bool issuer = (q == 0) && (hh == 0) && (lane == 0);
    const uint32_t gbuf = ptx::smem_u32(smem_out) + g * 2 * SLAB_BYTES;
    uint64_t* aux_full = aux_bar + g * 2;
    const uint32_t bar_id = 1 + g;
    const uint32_t swz = trow & 7;
    const GemmEpilogue epi = args.epi;
    const bool has_aux = (epi.mode == EPI_RESIDUAL || epi.mode == EPI_MUL || epi.mode == EPI_ROWDOT);
    const bool no_store = DBG && (args.dbg & 2);
    const uint32_t tmem_empty_remote = ptx::mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0);  // the leader's barrier
    uint32_t it = 0;
    uint32_t c = 0;  // slabs this group has produced; slab c uses buffer c & 1
    // timeline slot: [tile iteration < 24][slab][event < 12][warp 0 | warp 5]
    auto GTR = [&](int sl_, int ev) {
      if (DBG && (args.dbg & 512) && g_gemm_trace != nullptr && blockIdx.x == 0 && lane == 0 && (warp == 0 || warp == 5) &&
          it < 24)
        g_gemm_trace[((it * 2 + sl_) * 12 + ev) * 2 + (warp == 0 ? 0 : 1)] = clock64();
    };
    auto group_sync = [&]() {
      if (DBG && (args.dbg & 128)) return;  // timing experiment: no group barriers (results are garbage)
      asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "n"(128 * EW) : "memory");
    };
    auto issue_aux = [&](uint32_t cc, int tile_, int sl_) {  // issuer thread only
      const int am0 = mblock(tile_) * TILE_M + static_cast<int>(rank) * BM;
      const int acol = (tile_ % tiles_n) * BN + g * 128 + sl_ * 64;
      const uint32_t b = cc & 1;
      ptx::mbar_arrive_expect_tx(&aux_full[b], SLAB_BYTES);
      ptx::tma_load_2d(smem_out + g * 2 * SLAB_BYTES + b * SLAB_BYTES, &tmAux, &aux_full[b], acol, am0);
    };
    // write this thread's NC packed columns of its row into slab buffer b
    auto write_row = [&](uint32_t b, const uint32_t (&pk)[NP]) {
      if (DBG && (args.dbg & 32)) {  // timing experiment: no staging writes (keep the values alive)
        if (pk[3] == 0x12345678u && pk[NP - 1] == 0x9abcdef0u) tmem_slot[1] = pk[5];
        return;
      }
      const uint32_t rowaddr = gbuf + b * SLAB_BYTES + trow * 128;
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const uint32_t addr = rowaddr + (((hh * NQ + j) ^ swz) << 4);
        asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pk[4 * j]), "r"(pk[4 * j + 1]),
                     "r"(pk[4 * j + 2]), "r"(pk[4 * j + 3])
                     : "memory");
      }
    };
    // hand a filled slab to the TMA store engine (all warps of the group call; one thread issues)
    auto store_slab = [&](uint32_t b, const CUtensorMap* tm, int col, int row0) {
      ptx::fence_proxy_async_smem();
      group_sync();
      if (issuer && !no_store) {
        ptx::tma_store_2d(tm, smem_out + g * 2 * SLAB_BYTES + b * SLAB_BYTES, col, row0);
        ptx::tma_store_commit();
      }
    };
    auto epi_ttile = [&](int mlin) {
      {
        const int m0 = block_row0(mlin);
        // ---- T-tile epilogue: columns [0, tt_n) of this accumulator are T = A * TB^T for this CTA's 128 rows.  Column
        // group 0 converts them (+ the optional per-column bias) and writes its rows straight to global memory (16 KB per
        // CTA: no staging buffer, no TMA store to wait for); group 1 only returns the accumulator. ----
        const uint32_t tbuf = it & 1;
        ptx::mbar_wait(&tmem_full[tbuf], (it >> 1) & 1);
        ptx::tc_fence_after();
        // eight columns at a time (tcgen05.ld.x8 -> bias -> bf16 -> one 16-byte store): the T-tile is ~1 % of the work
        // and must not add register pressure to the main epilogue below
        const int my_cols = (g == 0) ? max(0, min(NC, args.tt_n - hh * NC)) : 0;
        const bool row_in = m0 + trow < args.M;
        const uint32_t ta = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + tbuf * BN + hh * NC;
        bf16* trow_ptr = args.tt_out + static_cast<size_t>(m0 + trow) * args.tt_ld + hh * NC;
#pragma unroll 1
        for (int j = 0; j < my_cols; j += 8) {
          uint32_t r8[8];
          ptx::tmem_ld_32x32b_x8(ta + j, r8);
          ptx::tmem_ld_wait();
          float f[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) f[k] = __uint_as_float(r8[k]);
          if (args.tt_bias != nullptr) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(args.tt_bias + hh * NC + j));
            const float4 b1 = __ldg(reinterpret_cast<const float4*>(args.tt_bias + hh * NC + j + 4));
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
          }
          if (row_in)
            ptx::st_global_v4(trow_ptr + j, pack_bf16x2(f[0], f[1]), pack_bf16x2(f[2], f[3]), pack_bf16x2(f[4], f[5]),
                              pack_bf16x2(f[6], f[7]));
        }
        ptx::tc_fence_before();
        __syncwarp();
        if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_remote + tbuf * 8);
        if (g == 0) {
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(tt_done);
        }
        ++it;
      }
    };
    if (has_aux && issuer && unit < num_tiles) issue_aux(0, unit, 0);
    if (unit < tt_pro) epi_ttile(unit);
    for (int tile = unit; tile < num_tiles; tile += num_units, ++it) {
      const int m0 = mblock(tile) * TILE_M + static_cast<int>(rank) * BM;
      const int n0 = (tile % tiles_n) * BN;
      {
        const int mt = ttile_of(tile);
        if (mt >= 0) epi_ttile(mt);
      }
      const uint32_t buf = it & 1;
      const uint32_t use = it >> 1;
      const int row = m0 + trow;
      const bool row_ok = row < args.M;
      // folded LayerNorm: this row's (mean, rstd), fetched before the accumulator is awaited
      const float2 st = (epi.row_stats != nullptr && row_ok) ? __ldg(epi.row_stats + row) : make_float2(0.f, 1.f);
      // the group's 128 columns of c1 / bias (4 lines each) are pulled into L1 before the accumulator is awaited: under
      // the main loop's shared-memory traffic an L1 miss in the middle of the epilogue costs thousands of clocks
      if (!(DBG && (args.dbg & 1024)) && q == 0 && hh == 0 && lane < 8) {
        const float* base = (lane < 4) ? epi.c1 : epi.bias;
        if (base != nullptr) ptx::prefetch_l1(base + n0 + g * 128 + (lane & 3) * 32);
      }
      GTR(0, 0);
      ptx::mbar_wait(&tmem_full[buf], use & 1);
      ptx::tc_fence_after();
      GTR(0, 1);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN + g * 128 + hh * NC;
      uint32_t rn[EW == 2 ? 32 : 1];  // EW == 2: slab 1's accumulator columns, fetched together with slab 0's
      (void)rn;
#pragma unroll(EW == 2 ? 2 : 1)
      for (int sl = 0; sl < 2; ++sl) {
        const int scol = n0 + g * 128 + sl * 64;  // first column of the slab (TMA coordinates)
        const int ncol = scol + hh * NC;          // first column this thread owns
        float v[NC];
        bool drained = false;  // every tcgen05.ld of this accumulator has completed
        if (DBG && (args.dbg & 16)) {
#pragma unroll
          for (int j = 0; j < NC; ++j) v[j] = 0.f;
          drained = (sl == 1);
        } else if constexpr (EW == 1) {
          uint32_t r0[32], r1[32];
          ptx::tmem_ld_32x32b_x32(taddr + sl * 64, r0);
          ptx::tmem_ld_32x32b_x32(taddr + sl * 64 + 32, r1);
          ptx::tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            v[j] = __uint_as_float(r0[j]);
            v[NC / 2 + j] = __uint_as_float(r1[j]);
          }
          drained = (sl == 1);
        } else {
          // slab 1's columns are requested while slab 0 is being staged / stored (prefetch_slab1 below), so the second
          // half of the tile starts without a TMEM round trip
          if (sl == 0) {
            uint32_t r0[32];
            ptx::tmem_ld_32x32b_x32(taddr, r0);
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r0[j]);
          } else {
            ptx::tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rn[j]);
            drained = true;
          }
        }
        // EW == 2: issue slab 1's TMEM load once slab 0's values are packed (v is dead, registers are free again)
        auto prefetch_slab1 = [&]() {
          if constexpr (EW == 2) {
            if (sl == 0 && !(DBG && (args.dbg & 16))) ptx::tmem_ld_32x32b_x32(taddr + 64, rn);
          }
        };
        GTR(sl, 2);
        if (drained) {  // hand the accumulator back to the MMA warp
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(tmem_empty_remote + buf * 8);
        }
        if (DBG && (args.dbg & 2048)) {  // timing experiment: per-row scale only, no per-column constants
#pragma unroll
          for (int j = 0; j < NC; ++j) v[j] *= st.y;
        } else if (epi.row_stats != nullptr && epi.c1 == nullptr) {
          // folded LayerNorm with the constants already in the accumulator (tensor-core rank-1 updates): rstd only
#pragma unroll
          for (int j = 0; j < NC; ++j) v[j] *= st.y;
        } else if (epi.row_stats != nullptr) {  // LayerNorm folded into this GEMM: acc <- rstd * (acc - mean * c1[n])
          const float nm = -st.x * st.y;
          const float4* cp = reinterpret_cast<const float4*>(epi.c1 + ncol);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            const float4 cc = __ldg(cp + j);
            v[4 * j] = fmaf(v[4 * j], st.y, nm * cc.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], st.y, nm * cc.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], st.y, nm * cc.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], st.y, nm * cc.w);
          }
        }
        if (epi.bias != nullptr && !(DBG && (args.dbg & (64 | 2048)))) {
          const float4* bp = reinterpret_cast<const float4*>(epi.bias + ncol);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            const float4 b = __ldg(bp + j);
            v[4 * j] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
        if (epi.mode == EPI_ROWTABLE && row_ok) {
          const float4* tp =
              reinterpret_cast<const float4*>(epi.table + static_cast<size_t>(row % epi.table_rows) * args.N + ncol);
#pragma unroll
          for (int j = 0; j < NC / 4; ++j) {
            const float4 b = __ldg(tp + j);
            v[4 * j] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
        uint32_t pk[NP];
        GTR(sl, 3);
        if (epi.mode == EPI_GELU_DUAL) {
          uint32_t pk2[NP];
          if (DBG && (args.dbg & 256)) {  // timing experiment: no GELU math
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
              pk2[j] = pk[j] ^ 0x00010001u;
            }
          } else if (args.gelu_f32 == 1) {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              float g0, d0, g1, d1;
              gelu_and_grad(v[2 * j], g0, d0);
              gelu_and_grad(v[2 * j + 1], g1, d1);
              pk[j] = pack_bf16x2(g0, g1);
              pk2[j] = pack_bf16x2(d0, d1);
            }
          } else {
#pragma unroll
            for (int j = 0; j < NP; ++j) {
              float g0, d0, g1, d1;
              gelu_and_grad_p(v[2 * j], g0, d0);
              gelu_and_grad_p(v[2 * j + 1], g1, d1);
              pk[j] = pack_bf16x2(g0, g1);
              pk2[j] = pack_bf16x2(d0, d1);
            }
          }
          prefetch_slab1();
          GTR(sl, 4);
          // two slabs per column slab: gelu(u) -> out (buffer 0), gelu'(u) -> out2 (buffer 1); both buffers were last
          // used one column slab ago, so one wait + two group barriers cover both stores
          if (issuer) ptx::tma_store_wait_read<0>();
          GTR(sl, 5);
          group_sync();
          GTR(sl, 6);
          write_row(0, pk);
          write_row(1, pk2);
          GTR(sl, 7);
          ptx::fence_proxy_async_smem();
          GTR(sl, 8);
          group_sync();
          GTR(sl, 9);
          if (issuer && !no_store) {
            ptx::tma_store_2d(&tmOut, smem_out + g * 2 * SLAB_BYTES, scol, m0);
            ptx::tma_store_2d(&tmOut2, smem_out + g * 2 * SLAB_BYTES + SLAB_BYTES, scol, m0);
            ptx::tma_store_commit();
          }
          GTR(sl, 10);
          c += 2;
        } else if (has_aux) {
          const uint32_t b = c & 1;
          ptx::mbar_wait(&aux_full[b], (c >> 1) & 1);  // residual / multiplier slab landed (=> the buffer was free)
          const uint32_t rowaddr = gbuf + b * SLAB_BYTES + trow * 128;
          float dot = 0.f;
#pragma unroll
          for (int j = 0; j < NQ; ++j) {
            uint4 a;
            const uint32_t addr = rowaddr + (((hh * NQ + j) ^ swz) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
            float f[8];
            if (args.res_f16 && epi.mode == EPI_RESIDUAL) unpack_f16x8(a, f);
            else unpack_bf16x8(a, f);
            if (epi.mode == EPI_ROWDOT) {
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                pk[4 * j + k] = pack_bf16x2(v[8 * j + 2 * k], v[8 * j + 2 * k + 1]);
                // the consumer reads the bf16-rounded values, so the dot product uses them too
                const float lo = __uint_as_float(pk[4 * j + k] << 16), hi = __uint_as_float(pk[4 * j + k] & 0xffff0000u);
                dot = fmaf(lo, f[2 * k], dot);
                dot = fmaf(hi, f[2 * k + 1], dot);
              }
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                if (epi.mode == EPI_RESIDUAL) v[8 * j + k] += f[k];
                else v[8 * j + k] *= f[k];
              }
            }
          }
          if (epi.mode == EPI_ROWDOT) {
            // one slab row = one head's 64 columns: only the EW == 1 layout holds them in one thread (the host never
            // selects EW == 2 for this mode)
            if (EW == 1 && row_ok)
              epi.rowdot[(static_cast<size_t>(row / epi.rowdot_rows) * (args.N >> 6) + (scol >> 6)) * epi.rowdot_pad +
                         row % epi.rowdot_rows] = dot;
          } else if (args.out_f16) {
#pragma unroll
            for (int j = 0; j < NP; ++j) pk[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < NP; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          }
          prefetch_slab1();
          write_row(b, pk);  // in place: each thread only ever touches its own part of its own row of the slab
          store_slab(b, &tmOut, scol, m0);
          ++c;
          if (issuer) {
            // the other buffer's last store has been read out -> fetch the next slab's residual / multiplier into it
            ptx::tma_store_wait_read<1>();
            if (sl == 0) issue_aux(c, tile, 1);
            else if (tile + num_units < num_tiles) issue_aux(c, tile + num_units, 0);
          }
        } else {
          if (args.out_f16) {
#pragma unroll
            for (int j = 0; j < NP; ++j) pk[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < NP; ++j) pk[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
          }
          prefetch_slab1();
          GTR(sl, 4);
          // One group barrier per slab: buffer c & 1 was last read by the store of slab c - 2, and the issuer only passes
          // the barrier of slab c - 1 (below) once that store has been read out, so nobody needs to wait here.
          const uint32_t b = c & 1;
          write_row(b, pk);
          GTR(sl, 7);
          ptx::fence_proxy_async_smem();
          if (issuer) ptx::tma_store_wait_read<0>();  // store c - 1: the buffer slab c + 1 will overwrite
          group_sync();
          if (issuer && !no_store) {
            ptx::tma_store_2d(&tmOut, smem_out + g * 2 * SLAB_BYTES + b * SLAB_BYTES, scol, m0);
            ptx::tma_store_commit();
          }
          GTR(sl, 10);
          ++c;
        }
      }
    }
    if (issuer) ptx::tma_store_wait_all<0>();
    __syncwarp();
  } else {
    // ================================= chunk epilogue (single-CTA kernel), warps 0..7 =================================
    // warp w owns TMEM lanes 32*(w%4).. (hardware rule) and column half w/4 of every tile.
    const int q = warp & 3;
    const int hsel = warp >> 2;
    constexpr int CPW = BN / (2 * CHUNK);  // 32-column chunks per warp per tile
    uint8_t* my_out = smem_out + warp * 2 * STAGE_OUT_BYTES;
    const GemmEpilogue epi = args.epi;
    const bool has_aux = (epi.mode == EPI_RESIDUAL || epi.mode == EPI_MUL) && !(DBG && (args.dbg & 1));
    uint32_t it = 0;
    uint32_t store_idx = 0;
    uint4 aux_next[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) aux_next[i] = make_uint4(0, 0, 0, 0);
    // prefetch the aux (residual / multiplier) rows of the first chunk before the accumulator is ready
    const uint32_t tmem_empty_remote = TWO ? ptx::mapa_shared(ptx::smem_u32(&tmem_empty[0]), 0) : 0u;  // leader's barrier
    if (has_aux && unit < num_tiles) {
      const int m0 = mblock(unit) * TILE_M + static_cast<int>(rank) * BM, n0 = (unit % tiles_n) * BN;
      const int row = m0 + q * 32 + lane;
      if (row < args.M) {
        const uint4* rp = reinterpret_cast<const uint4*>(epi.res + static_cast<size_t>(row) * epi.ld_res + n0 +
                                                         hsel * CPW * CHUNK);
#pragma unroll
        for (int i = 0; i < 4; ++i) aux_next[i] = __ldg(rp + i);
      }
    }
    uint32_t scnt = 0;  // k-block counter of the row-statistics warps (same sequence as the producer's)
    if (!TWO && epi.stats_out != nullptr && hsel == 0) asm volatile("bar.arrive 2, 256;" ::: "memory");
    for (int tile = unit; tile < num_tiles; tile += num_units, ++it) {
      const int m0 = mblock(tile) * TILE_M + static_cast<int>(rank) * BM;
      const int n0 = (tile % tiles_n) * BN;
      const uint32_t buf = it & 1;
      const uint32_t use = it >> 1;
      float2 row_mr = make_float2(0.f, 1.f);
      if (!TWO && epi.stats_out != nullptr) {
        // LayerNorm statistics of this tile's 128 A rows, read from the operand stages as they land (thread = row).
        // All eight epilogue warps take part: warp (q, hsel) owns the k-blocks that land in stages of parity hsel (STAGES
        // is even, so a stage always belongs to the same group and its empty barrier counts the MMA + that group's four
        // warps).  Each group accumulates shifted sums (d = x - its first element: well conditioned when |mean| >> std);
        // group 1 hands its (mean, M2) to group 0 through shared memory and the two are merged (Chan et al.).
        const int srow = q * 32 + lane;
        float shift = 0.f, sd = 0.f, sdd = 0.f;
        int nproc = 0;
        for (int kb = 0; kb < num_kb; ++kb, ++scnt) {
          const int s = scnt % STAGES;
          if ((s & 1) != hsel) continue;
          ptx::mbar_wait(&full_bar[s], (scnt / STAGES) & 1);
          if (kb < main_kb) {
            const uint32_t rowaddr = ptx::smem_u32(smem_stage + s * Cfg::STAGE_BYTES) + srow * 128;
#pragma unroll
            for (int c = 0; c < 8; ++c) {
              uint4 a;
              const uint32_t addr = rowaddr + ((c ^ (srow & 7)) << 4);
              asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "r"(addr));
              float f[8];
              if (args.a_f16) unpack_f16x8(a, f);
              else unpack_bf16x8(a, f);
              if (nproc == 0 && c == 0) shift = f[0];
#pragma unroll
              for (int k = 0; k < 8; ++k) {
                const float d = f[k] - shift;
                sd += d;
                sdd = fmaf(d, d, sdd);
              }
            }
            ++nproc;
          }
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive(&empty_bar[s]);
        }
        const float n_g = static_cast<float>(nproc * BK);
        const float mean_g = nproc ? shift + sd / n_g : 0.f;
        const float m2_g = nproc ? fmaxf(sdd - sd * sd / n_g, 0.f) : 0.f;
        float2* xch = reinterpret_cast<float2*>(smem_xch);
        if (hsel == 1) {
          asm volatile("bar.sync 2, 256;" ::: "memory");   // group 0 has read the previous tile's hand-over
          xch[srow] = make_float2(mean_g, m2_g);
          asm volatile("bar.sync 1, 256;" ::: "memory");
        } else {
          asm volatile("bar.sync 1, 256;" ::: "memory");
          const float2 o = xch[srow];
          asm volatile("bar.arrive 2, 256;" ::: "memory");
          const float n_o = static_cast<float>(args.K) - n_g, n = static_cast<float>(args.K);
          const float delta = o.x - mean_g;
          const float mean = mean_g + delta * (n_o / n);
          const float var = fmaxf((m2_g + o.y + delta * delta * (n_g * n_o / n)) / n, 0.f);
          const float2 mr = make_float2(mean, rsqrtf(var + epi.stats_eps));
          if (m0 + srow < args.M) epi.stats_out[m0 + srow] = mr;
          row_mr = mr;
        }
      }
      ptx::mbar_wait(&tmem_full[buf], use & 1);
      ptx::tc_fence_after();
      const int row = m0 + q * 32 + lane;
      const bool row_ok = row < args.M;
      const uint32_t taddr_row = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + buf * BN;

#pragma unroll 1
      for (int i = 0; i < ((DBG && (args.dbg & 16)) ? 0 : CPW); ++i) {
        // T's constant columns live in the first chunk of every 64-column group: with them enabled the chunks are
        // interleaved so that those (even) chunks all belong to the statistics warps 0..3, which hold (mean, rstd) of
        // their rows in registers -- no exchange with warps 4..7 needed
        const int c = (epi.stat_col > 0) ? (2 * i + hsel) : (hsel * CPW + i);
        const int ncol = n0 + c * CHUNK;
        uint32_t r[32];
        ptx::tmem_ld_32x32b_x32(taddr_row + c * CHUNK, r);
        uint4 aux[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) aux[j] = aux_next[j];
        if (has_aux) {
          // software pipeline: fetch the next chunk's aux rows (possibly of the next tile) while this one computes
          int nrow = row, ncol_next = ncol + CHUNK;
          bool ok = row_ok;
          if (i + 1 == CPW) {
            const int nt = tile + num_units;
            ok = nt < num_tiles;
            nrow = mblock(nt) * TILE_M + static_cast<int>(rank) * BM + q * 32 + lane;
            ncol_next = (nt % tiles_n) * BN + hsel * CPW * CHUNK;
            ok = ok && nrow < args.M;
          }
          if (ok) {
            const uint4* rp = reinterpret_cast<const uint4*>(epi.res + static_cast<size_t>(nrow) * epi.ld_res + ncol_next);
#pragma unroll
            for (int j = 0; j < 4; ++j) aux_next[j] = __ldg(rp + j);
          }
        }
        ptx::tmem_ld_wait();
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (epi.row_stats != nullptr) {  // LayerNorm folded into this GEMM
          const float2 st = row_ok ? __ldg(epi.row_stats + row) : make_float2(0.f, 1.f);
          const float nm = -st.x * st.y;
          const float4* cp = reinterpret_cast<const float4*>(epi.c1 + ncol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 c = __ldg(cp + j);
            v[4 * j] = fmaf(v[4 * j], st.y, nm * c.x);
            v[4 * j + 1] = fmaf(v[4 * j + 1], st.y, nm * c.y);
            v[4 * j + 2] = fmaf(v[4 * j + 2], st.y, nm * c.z);
            v[4 * j + 3] = fmaf(v[4 * j + 3], st.y, nm * c.w);
          }
        }
        if (epi.bias != nullptr) {
          const float4* bp = reinterpret_cast<const float4*>(epi.bias + ncol);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 b = __ldg(bp + j);
            v[4 * j] += b.x;
            v[4 * j + 1] += b.y;
            v[4 * j + 2] += b.z;
            v[4 * j + 3] += b.w;
          }
        }
        if (has_aux) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float f[8];
            if (args.res_f16 && epi.mode == EPI_RESIDUAL) unpack_f16x8(aux[j], f);
            else unpack_bf16x8(aux[j], f);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              if (epi.mode == EPI_RESIDUAL) v[8 * j + k] += f[k];
              else v[8 * j + k] *= f[k];
            }
          }
        } else if (epi.mode == EPI_ROWTABLE) {
          if (row_ok) {
            const float4* tp = reinterpret_cast<const float4*>(
                epi.table + static_cast<size_t>(row % epi.table_rows) * args.N + ncol);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = __ldg(tp + j);
              v[4 * j] += b.x;
              v[4 * j + 1] += b.y;
              v[4 * j + 2] += b.z;
              v[4 * j + 3] += b.w;
            }
          }
        }
        if (!TWO && epi.stats_out != nullptr && epi.stat_col > 0 && ((c & 1) == 0)) {
          // first chunk of a 64-column group: columns stat_col .. stat_col + 5 become the per-row factors of the
          // consumer's tensor-core constants (hi / lo bf16 splits keep ~16 mantissa bits)
          const float sg = 1.0f / row_mr.y;
          const float mh = __bfloat162float(__float2bfloat16(row_mr.x)), ml = row_mr.x - mh;
          const float sh = __bfloat162float(__float2bfloat16(sg)), sl = sg - sh;
          auto fill = [&](auto off) {  // compile-time offset: six register moves
            constexpr int o = decltype(off)::value;
            v[o] = -mh; v[o + 1] = -mh; v[o + 2] = -ml; v[o + 3] = sh; v[o + 4] = sh; v[o + 5] = sl;
          };
          switch (epi.stat_col) {  // ranks are multiples of 8 in practice
            case 8: fill(std::integral_constant<int, 8>{}); break;
            case 16: fill(std::integral_constant<int, 16>{}); break;
            case 24: fill(std::integral_constant<int, 24>{}); break;
            default:
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int k = j - epi.stat_col;
                if (static_cast<unsigned>(k) < 6u) v[j] = (k < 2) ? -mh : ((k == 2) ? -ml : ((k < 5) ? sh : sl));
              }
          }
        }
        uint32_t packed[16];
        if (epi.mode == EPI_GELU_DUAL) {
          uint32_t packed2[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            float g0, d0, g1, d1;
            gelu_and_grad(v[2 * j], g0, d0);
            gelu_and_grad(v[2 * j + 1], g1, d1);
            packed[j] = pack_bf16x2(g0, g1);
            packed2[j] = pack_bf16x2(d0, d1);
          }
          if (!(DBG && (args.dbg & 2)))
            stage_and_store(my_out + (store_idx & 1) * STAGE_OUT_BYTES, packed2, &tmOut2, ncol, m0 + q * 32, lane);
          ++store_idx;
        } else if (args.out_f16) {
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[j] = pack_f16x2(v[2 * j], v[2 * j + 1]);
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
        }
        if (!(DBG && (args.dbg & 2)))
          stage_and_store(my_out + (store_idx & 1) * STAGE_OUT_BYTES, packed, &tmOut, ncol, m0 + q * 32, lane);
        else if (packed[3] == 0x12345678u && packed[7] == 0x9abcdef0u)
          tmem_slot[1] = packed[5];  // keep the math alive when stores are disabled
        ++store_idx;
      }
      // all tcgen05.ld of this accumulator have completed (wait::ld above) -> hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (TWO) ptx::mbar_arrive_cluster(tmem_empty_remote + buf * 8);
        else ptx::mbar_arrive(&tmem_empty[buf]);
      }
    }
    if (lane == 0) ptx::tma_store_wait_all<0>();
    __syncwarp();
  }

  ptx::tc_fence_before();
  if constexpr (TWO) ptx::cluster_sync();  // the peer's smem / TMEM stay valid until the leader's last MMA retired
  else __syncthreads();
  if (warp == TMA_WARP) {
    ptx::tc_fence_after();
    if constexpr (TWO) ptx::tmem_dealloc_2cta(tmem_base, Cfg::TMEM_COLS);
    else ptx::tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || p == nullptr) {
    set_error("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: %s", cudaGetErrorString(e));
    return nullptr;
  }
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// 2D bf16 row-major tensor [rows, cols] with leading dimension ld (elements); box = [box_cols, box_rows]
static int make_tmap_2d(CUtensorMap* tm, const void* base, uint64_t rows, uint64_t cols, uint64_t ld,
                        uint32_t box_cols, uint32_t box_rows, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || ((ld * 2) & 15)) {
    set_error("tensor map: base %p / ld %llu not 16-byte aligned", base, (unsigned long long)ld);
    return 1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) rows=%llu cols=%llu ld=%llu box=%ux%u", (int)r,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_cols, box_rows);
    return 1;
  }
  return 0;
}

// 3D bf16 tensor [d2][d1][d0] (d0 contiguous) with byte strides s1 (between d1 rows) and s2 (between d2 slabs);
// box = [b0, b1, 1], 128B (default) or 64B swizzle, out-of-bounds elements read as zero / are not written.
int make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, int swizzle_bytes) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return 1;
  if ((reinterpret_cast<uintptr_t>(base) & 15) || (s1_bytes & 15) || (s2_bytes & 15)) {
    set_error("tensor map 3d: base/strides not 16-byte aligned");
    return 1;
  }
  cuuint64_t gdim[3] = {d0, d1, d2};
  cuuint64_t gstride[2] = {s1_bytes, s2_bytes};
  cuuint32_t box[3] = {b0, b1, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled(3d) failed (%d) dims=%llu,%llu,%llu box=%u,%u", (int)r, (unsigned long long)d0,
              (unsigned long long)d1, (unsigned long long)d2, b0, b1);
    return 1;
  }
  return 0;
}

static bool gemm_two_cta_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VITATK_GEMM_2CTA");  // "0" selects the single-CTA kernel for every shape
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

int gemm_plan_init(GemmPlan* p, int M, int N, int K, const bf16* A, int lda, const bf16* B, int ldb, bf16* out,
                   int ldo, bf16* out2, int ldo2, const bf16* T, int ldt, const bf16* LB, int ldlb, int lora_nkb,
                   int lora_ksteps, int lora_group_cols, GemmEpilogue epi, const GemmTT* tt) {
  if (K % BK != 0 || N % 64 != 0 || M <= 0) {
    set_error("gemm_plan_init: unsupported shape M=%d N=%d K=%d (K%%64, N%%64 must be 0)", M, N, K);
    return 1;
  }
  p->M = M;
  p->N = N;
  p->K = K;
  p->BN = (N % 256 == 0) ? 256 : (N % 192 == 0 ? 192 : (N % 128 == 0 ? 128 : 64));
  p->lora_nkb = lora_nkb;
  p->lora_ksteps = lora_ksteps;
  p->lora_group_cols = lora_group_cols;
  p->epi = epi;
  p->reverse_m = 0;
  p->a_f16 = p->out_f16 = p->res_f16 = 0;
  p->tt = GemmTT{};
  if (tt != nullptr && tt->n > 0) p->tt = *tt;
  if (lora_group_cols > 0 && lora_group_cols % p->BN != 0) {
    set_error("gemm_plan_init: lora_group_cols %d not a multiple of BN %d", lora_group_cols, p->BN);
    return 1;
  }
  if (make_tmap_2d(&p->tmA, A, M, K, lda, BK, BM)) return 1;
  p->two_cta = (p->BN == 256 && gemm_two_cta_enabled()) ? 1 : 0;
  const int b_rows = p->two_cta ? p->BN / 2 : p->BN;
  if (make_tmap_2d(&p->tmB, B, N, K, ldb, BK, b_rows)) return 1;
  // store boxes: pair kernel = 64-col x 128-row slabs (128B swizzle); single kernel = 32 x 32 chunks (64B swizzle)
  const uint32_t obc = p->two_cta ? 64 : CHUNK, obr = p->two_cta ? 128 : 32;
  const CUtensorMapSwizzle osw = p->two_cta ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  if (make_tmap_2d(&p->tmOut, out, M, N, ldo, obc, obr, osw)) return 1;
  if (out2) {
    if (make_tmap_2d(&p->tmOut2, out2, M, N, ldo2, obc, obr, osw)) return 1;
  } else {
    p->tmOut2 = p->tmOut;
  }
  if (epi.stats_out && (p->two_cta || N != p->BN || lora_nkb != 0)) {
    set_error("gemm_plan_init: row statistics need the single-CTA kernel with one N-tile and no LoRA k-blocks");
    return 1;
  }
  if (epi.stat_col != 0 && (!epi.stats_out || epi.stat_col < 1 || epi.stat_col + 6 > 32 || epi.mode != EPI_PLAIN)) {
    set_error("gemm_plan_init: stat_col %d needs stats_out, EPI_PLAIN and columns stat_col..stat_col+5 inside the first "
              "32 of a group", epi.stat_col);
    return 1;
  }
  if (epi.row_stats && !epi.c1 && (!p->two_cta || lora_nkb == 0)) {
    set_error("gemm_plan_init: the scale-only LayerNorm fold needs the pair kernel and a LoRA k-block carrying the constants");
    return 1;
  }
  if (epi.mode == EPI_ROWDOT && (!p->two_cta || !epi.rowdot || epi.rowdot_rows <= 0)) {
    set_error("gemm_plan_init: EPI_ROWDOT needs the pair kernel (N %% 256 == 0, VITATK_GEMM_2CTA != 0) and a side buffer");
    return 1;
  }
  if (p->two_cta && (epi.mode == EPI_RESIDUAL || epi.mode == EPI_MUL || epi.mode == EPI_ROWDOT)) {
    if (make_tmap_2d(&p->tmAux, epi.res, M, N, epi.ld_res, 64, 128)) return 1;
  } else {
    p->tmAux = p->tmOut;
  }
  p->tmTB = p->tmB;
  if (p->tt.n > 0) {
    const GemmTT& t = p->tt;
    if (!p->two_cta || lora_nkb != 1 || lora_group_cols != 0 || (t.n != 32 && t.n != 64) || lora_ksteps * 16 > t.n ||
        t.out != T || t.ld_out != ldt || t.tb == nullptr || t.flags == nullptr) {
      set_error("gemm_plan_init: T-tiles need the pair kernel, one LoRA k-block reading the T they produce, n in {32, 64} "
                ">= 16 * lora_ksteps, and a flag array");
      return 1;
    }
    if (make_tmap_2d(&p->tmTB, t.tb, 64, K, t.ld_tb, BK, t.n / 2)) return 1;
  }
  if (lora_nkb > 0) {
    const int tcols = lora_group_cols > 0 ? (N / lora_group_cols) * 64 : lora_nkb * 64;
    if (make_tmap_2d(&p->tmLA, T, M, tcols, ldt, BK, BM)) return 1;
    if (make_tmap_2d(&p->tmLB, LB, N, lora_nkb * 64, ldlb, BK, b_rows)) return 1;
  } else {
    p->tmLA = p->tmA;
    p->tmLB = p->tmB;
  }
  return 0;
}

static int gemm_gelu_f32() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VITATK_GELU");
    v = (e && strcmp(e, "f32") == 0) ? 1 : 2;  // default: fp32 2^P fit
  }
  return v;
}

static bool gemm_epi16_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VITATK_GEMM_EPI16");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}

static int gemm_epi16_max_k() {  // VITATK_GEMM_EPI16_MAXK: largest K that still gets the 16-warp epilogue (default 1024)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VITATK_GEMM_EPI16_MAXK");
    v = e ? atoi(e) : 1024;
  }
  return v;
}

static int gemm_dbg_flags() {
  static int dbg = -1;
  if (dbg < 0) {
    const char* e = getenv("VITATK_GEMM_DBG");
    dbg = e ? atoi(e) : 0;
  }
  return dbg;
}

template <int BN, bool TWO, int EW = 1>
static int launch_bn(const GemmPlan* p, cudaStream_t stream, int num_sms) {
  using Cfg = GemmCfg<BN, TWO>;
  static PerDeviceOnce once;
  if (once.need()) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(gemm_tc05_kernel<BN, false, TWO, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
#ifdef VITATK_DBG_KERNELS
    VITATK_CUDA_OK(cudaFuncSetAttribute(gemm_tc05_kernel<BN, true, TWO, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
#endif
  }
  const int tile_m = TWO ? 2 * BM : BM;
  const int tiles = ((p->M + tile_m - 1) / tile_m) * (p->N / BN);
  int max_units = TWO ? num_sms / 2 : num_sms;
  if constexpr (TWO) {
    // Units wait on one another (T-tiles), so every cluster of the grid must be resident at once: ask the runtime how
    // many 2-CTA clusters of this kernel the device can hold (a GPC with an odd number of SMs leaves one unpaired).
    static int max_clusters[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    int& mc = max_clusters[dev & 63];
    if (mc == 0) {
      cudaLaunchConfig_t qc = {};
      qc.gridDim = dim3(2 * (num_sms / 2), 1, 1);
      qc.blockDim = dim3(gemm_threads(TWO, EW), 1, 1);
      qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
      cudaLaunchAttribute qa[1];
      qa[0].id = cudaLaunchAttributeClusterDimension;
      qa[0].val.clusterDim.x = 2;
      qa[0].val.clusterDim.y = 1;
      qa[0].val.clusterDim.z = 1;
      qc.attrs = qa;
      qc.numAttrs = 1;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, gemm_tc05_kernel<BN, false, TWO, EW>, &qc) != cudaSuccess || n < 1) {
        cudaGetLastError();
        n = num_sms / 2;
      }
      mc = n;
    }
    if (mc < max_units) max_units = mc;
  }
  const int units = tiles < max_units ? tiles : max_units;
  GemmKernelArgs a;
  a.M = p->M;
  a.N = p->N;
  a.K = p->K;
  a.lora_nkb = p->lora_nkb;
  a.lora_ksteps = p->lora_ksteps;
  a.lora_group_cols = p->lora_group_cols;
  a.epi = p->epi;
  a.dbg = gemm_dbg_flags();
  a.gelu_f32 = gemm_gelu_f32();
  a.reverse_m = p->reverse_m;
  a.tt_n = TWO ? p->tt.n : 0;
  a.tt_out = p->tt.out;
  a.tt_ld = p->tt.ld_out;
  a.tt_bias = p->tt.bias;
  a.tt_flags = p->tt.flags;
  a.a_f16 = p->a_f16;
  a.out_f16 = p->out_f16;
  a.res_f16 = p->res_f16;
  {
    static int rt = -1;
    if (rt < 0) {
      const char* e = getenv("VITATK_GEMM_RT");
      rt = e ? atoi(e) : 0;
    }
    a.dbg_rt = rt;
  }
  const dim3 grid(TWO ? 2 * units : units, 1, 1), block(gemm_threads(TWO, EW), 1, 1);
#ifdef VITATK_DBG_KERNELS  // timing-experiment instantiations (VITATK_GEMM_DBG switches) are not part of the product build
  if (a.dbg) {
    VITATK_CUDA_OK(launch_pdl(gemm_tc05_kernel<BN, true, TWO, EW>, grid, block, Cfg::SMEM_BYTES, stream, TWO ? 2 : 1, p->tmA,
                              p->tmB, p->tmLA, p->tmLB, p->tmOut, p->tmOut2, p->tmAux, p->tmTB, a));
    return 0;
  }
#endif
  VITATK_CUDA_OK(launch_pdl(gemm_tc05_kernel<BN, false, TWO, EW>, grid, block, Cfg::SMEM_BYTES, stream, TWO ? 2 : 1, p->tmA,
                              p->tmB, p->tmLA, p->tmLB, p->tmOut, p->tmOut2, p->tmAux, p->tmTB, a));
  return 0;
}

int gemm_set_trace(long long* dev_buf) {
#ifdef VITATK_DBG_KERNELS
  VITATK_CUDA_OK(cudaMemcpyToSymbol(g_gemm_trace, &dev_buf, sizeof(dev_buf)));
  return 0;
#else
  (void)dev_buf;
  set_error("gemm_set_trace: build with -DVITATK_DBG_KERNELS (VITATK_DBG_BUILD=1) for the in-kernel timeline");
  return 1;
#endif
}

int gemm_launch(const GemmPlan* p, cudaStream_t stream, int num_sms) {
  switch (p->BN) {
    case 256:
      if (!p->two_cta) return launch_bn<256, false>(p, stream, num_sms);
      // 16 epilogue warps (two per TMEM lane quarter and column group) for the short-K GEMMs, whose epilogue is as long
      // as their main loop (measured: fc1 -7 %, qkv -4 %, proj -4 %, bfc2 -3 %); the K >= 2304 GEMMs are main-loop
      // bound and lose ~3 % to the extra warps, and ROWDOT needs a whole slab row in one thread
      if (gemm_epi16_enabled() && p->epi.mode != EPI_ROWDOT && p->K <= gemm_epi16_max_k())
        return launch_bn<256, true, 2>(p, stream, num_sms);
      return launch_bn<256, true, 1>(p, stream, num_sms);
    case 192: return launch_bn<192, false>(p, stream, num_sms);
    case 128: return launch_bn<128, false>(p, stream, num_sms);
    case 64: return launch_bn<64, false>(p, stream, num_sms);
  }
  set_error("gemm_launch: bad BN %d", p->BN);
  return 1;
}

}  // namespace vitatk

// Fused softmax(Q K^T / sqrt(d)) V attention for ViT (197 tokens, head dim 64), forward and the
// input-gradient backward, one CTA per (image, head) with the whole head resident in shared memory.
//
// Replaces HF eager/sdpa attention (HF modeling_vit.py:185-193,228-249) and its autograd backward.
// Layout: q|k|v packed token-major [B*T, 3*D] exactly as the fused QKV GEMM writes it (head h of q at
// columns h*64, of k at D + h*64, of v at 2D + h*64); the output is token-major [B*T, D] so the proj
// GEMM consumes it without any transpose/copy kernel.
//
// Round-1 implementation uses warp-level mma.sync (m16n8k16 bf16, fp32 accumulate) with a full
// 16 x 208 score row-block held in registers (no online softmax needed at T = 197).
#include "vitatk_internal.h"

namespace vitatk {

static constexpr int HD = 64;          // head dim
static constexpr int TPAD = 208;       // 197 padded to 13 x 16
static constexpr int MT = TPAD / 16;   // 13 row tiles
static constexpr int NT8 = TPAD / 8;   // 26 column tiles of 8
static constexpr int LDS = 72;         // smem row pitch (bf16): 144 B, conflict-free ldmatrix
static constexpr int ATT_WARPS = 7;
static constexpr int ATT_THREADS = ATT_WARPS * 32;

__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], const bf16* p) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(p));
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(a));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack2(uint32_t u) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u));
}
__device__ __forceinline__ void cp_async16(bf16* dst, const bf16* src) {
  const uint32_t a = static_cast<uint32_t>(__cvta_generic_to_shared(dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(a), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// load one [TPAD x 64] head slice (rows >= tokens zero-filled) into smem with pitch LDS
__device__ __forceinline__ void load_head(bf16* dst, const bf16* src, int ld, int tokens, int tid) {
  for (int idx = tid; idx < TPAD * 8; idx += ATT_THREADS) {
    const int row = idx >> 3, ch = idx & 7;
    bf16* d = dst + row * LDS + ch * 8;
    if (row < tokens) cp_async16(d, src + static_cast<size_t>(row) * ld + ch * 8);
    else *reinterpret_cast<uint4*>(d) = make_uint4(0, 0, 0, 0);
  }
}

// acc[nt][.] = A_tile[16 x 64] * Bmat[n][k]^T for all 26 column tiles (Bmat rows are the n index)
__device__ __forceinline__ void rowblock_times_rowsT(float (&acc)[NT8][4], const uint32_t (&a)[4][4], const bf16* Bmat,
                                                     int lane) {
#pragma unroll
  for (int nt = 0; nt < NT8; ++nt) {
    uint32_t b0[4], b1[4];
    const bf16* p = Bmat + (nt * 8 + (lane & 7)) * LDS + (lane >> 3) * 8;
    ldsm_x4(b0, p);       // k 0..31
    ldsm_x4(b1, p + 32);  // k 32..63
    mma16816(acc[nt], a[0], b0[0], b0[1]);
    mma16816(acc[nt], a[1], b0[2], b0[3]);
    mma16816(acc[nt], a[2], b1[0], b1[1]);
    mma16816(acc[nt], a[3], b1[2], b1[3]);
  }
}
// out[dt][.] = P[16 x 208] (A fragments pa) * Bmat[k][n] with Bmat row index = k (transposed ldmatrix)
__device__ __forceinline__ void frag_times_rows(float (&out)[8][4], const uint32_t (&pa)[MT][4], const bf16* Bmat,
                                                int lane) {
#pragma unroll
  for (int kk = 0; kk < MT; ++kk) {
#pragma unroll
    for (int dt = 0; dt < 4; ++dt) {
      uint32_t b[4];
      ldsm_x4_t(b, Bmat + (kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8) * LDS + dt * 16 + (lane >> 4) * 8);
      mma16816(out[2 * dt], pa[kk], b[0], b[1]);
      mma16816(out[2 * dt + 1], pa[kk], b[2], b[3]);
    }
  }
}
__device__ __forceinline__ void load_a_frags(uint32_t (&a)[4][4], const bf16* mat, int row0, int lane) {
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm_x4(a[ks], mat + (row0 + (lane & 15)) * LDS + ks * 16 + (lane >> 4) * 8);
}
// write a 16 x 64 fp32 accumulator tile as bf16 to global rows row0.. (token-major, pitch ld) via smem staging
__device__ __forceinline__ void store_tile(const float (&o)[8][4], bf16* stage, bf16* gdst, int ld, int row0, int tokens,
                                           int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int n = 0; n < 8; ++n) {
    *reinterpret_cast<uint32_t*>(stage + g * LDS + n * 8 + 2 * t) = pack2(o[n][0], o[n][1]);
    *reinterpret_cast<uint32_t*>(stage + (g + 8) * LDS + n * 8 + 2 * t) = pack2(o[n][2], o[n][3]);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int id = lane + 32 * i, r = id >> 3, ch = id & 7;
    if (row0 + r < tokens)
      *reinterpret_cast<uint4*>(gdst + static_cast<size_t>(row0 + r) * ld + ch * 8) =
          *reinterpret_cast<const uint4*>(stage + r * LDS + ch * 8);
  }
  __syncwarp();
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_fwd_kernel(const bf16* __restrict__ qkv, bf16* __restrict__ out, int tokens, int heads, float sl2) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Qs = reinterpret_cast<bf16*>(att_smem);
  bf16* Ks = Qs + TPAD * LDS;
  bf16* Vs = Ks + TPAD * LDS;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * HD, ld = 3 * D;
  const bf16* base = qkv + static_cast<size_t>(b) * tokens * ld + h * HD;
  load_head(Qs, base, ld, tokens, tid);
  load_head(Ks, base + D, ld, tokens, tid);
  load_head(Vs, base + 2 * D, ld, tokens, tid);
  cp_async_wait_all();
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  (void)g;
  for (int mt = warp; mt < MT; mt += ATT_WARPS) {
    uint32_t qa[4][4];
    load_a_frags(qa, Qs, mt * 16, lane);
    float s[NT8][4];
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
    rowblock_times_rowsT(s, qa, Ks, lane);
    float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const int j = nt * 8 + 2 * t;
      if (j >= tokens) s[nt][0] = s[nt][2] = -INFINITY;
      if (j + 1 >= tokens) s[nt][1] = s[nt][3] = -INFINITY;
      mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
      mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
    }
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
    mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
    mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
    const float m0 = mx0 * sl2, m1 = mx1 * sl2;
    float sum0 = 0.f, sum1 = 0.f;
    uint32_t pa[MT][4];
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const float p0 = exp2f(s[nt][0] * sl2 - m0), p1 = exp2f(s[nt][1] * sl2 - m0);
      const float p2 = exp2f(s[nt][2] * sl2 - m1), p3 = exp2f(s[nt][3] * sl2 - m1);
      sum0 += p0 + p1;
      sum1 += p2 + p3;
      pa[nt >> 1][(nt & 1) * 2] = pack2(p0, p1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2, p3);
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    float o[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    frag_times_rows(o, pa, Vs, lane);
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
#pragma unroll
    for (int n = 0; n < 8; ++n) {
      o[n][0] *= inv0;
      o[n][1] *= inv0;
      o[n][2] *= inv1;
      o[n][3] *= inv1;
    }
    // the Q rows of this tile are only read by this warp (already in registers): reuse them as staging
    store_tile(o, Qs + mt * 16 * LDS, out + static_cast<size_t>(b) * tokens * D + h * HD, D, mt * 16, tokens, lane);
  }
}

__global__ void __launch_bounds__(ATT_THREADS, 1)
attn_bwd_kernel(const bf16* __restrict__ qkv, const bf16* __restrict__ dout, bf16* __restrict__ dqkv, int tokens,
                int heads, float scale, float sl2) {
  extern __shared__ __align__(16) uint8_t att_smem[];
  bf16* Qs = reinterpret_cast<bf16*>(att_smem);
  bf16* Ks = Qs + TPAD * LDS;
  bf16* Vs = Ks + TPAD * LDS;
  bf16* Os = Vs + TPAD * LDS;  // dO
  bf16* Stage = Os + TPAD * LDS;  // [ATT_WARPS][16][LDS]
  float* lse2 = reinterpret_cast<float*>(Stage + ATT_WARPS * 16 * LDS);  // [TPAD] log2-domain logsumexp
  float* delta = lse2 + TPAD;                                             // [TPAD]
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int b = blockIdx.x / heads, h = blockIdx.x % heads;
  const int D = heads * HD, ld = 3 * D;
  const bf16* base = qkv + static_cast<size_t>(b) * tokens * ld + h * HD;
  bf16* gbase = dqkv + static_cast<size_t>(b) * tokens * ld + h * HD;
  load_head(Qs, base, ld, tokens, tid);
  load_head(Ks, base + D, ld, tokens, tid);
  load_head(Vs, base + 2 * D, ld, tokens, tid);
  load_head(Os, dout + static_cast<size_t>(b) * tokens * D + h * HD, D, tokens, tid);
  cp_async_wait_all();
  __syncthreads();
  const int g = lane >> 2, t = lane & 3;
  bf16* stage = Stage + warp * 16 * LDS;

  // ---------------- pass A: row blocks -> softmax stats, delta, dQ ----------------
  for (int mt = warp; mt < MT; mt += ATT_WARPS) {
    uint32_t pa[MT][4];
    float l0, l1;
    {
      uint32_t qa[4][4];
      load_a_frags(qa, Qs, mt * 16, lane);
      float s[NT8][4];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
      rowblock_times_rowsT(s, qa, Ks, lane);
      float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const int j = nt * 8 + 2 * t;
        if (j >= tokens) s[nt][0] = s[nt][2] = -INFINITY;
        if (j + 1 >= tokens) s[nt][1] = s[nt][3] = -INFINITY;
        mx0 = fmaxf(mx0, fmaxf(s[nt][0], s[nt][1]));
        mx1 = fmaxf(mx1, fmaxf(s[nt][2], s[nt][3]));
      }
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
      mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
      mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
      const float m0 = mx0 * sl2, m1 = mx1 * sl2;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        sum0 += exp2f(s[nt][0] * sl2 - m0) + exp2f(s[nt][1] * sl2 - m0);
        sum1 += exp2f(s[nt][2] * sl2 - m1) + exp2f(s[nt][3] * sl2 - m1);
      }
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
      sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
      sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
      l0 = m0 + log2f(sum0);
      l1 = m1 + log2f(sum1);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        pa[nt >> 1][(nt & 1) * 2] = pack2(exp2f(s[nt][0] * sl2 - l0), exp2f(s[nt][1] * sl2 - l0));
        pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(exp2f(s[nt][2] * sl2 - l1), exp2f(s[nt][3] * sl2 - l1));
      }
    }
    if (t == 0) {
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      lse2[r0] = r0 < tokens ? l0 : INFINITY;
      lse2[r1] = r1 < tokens ? l1 : INFINITY;
    }
    // dP = dO_tile * V^T
    float dp[NT8][4];
    {
      uint32_t da[4][4];
      load_a_frags(da, Os, mt * 16, lane);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) dp[nt][0] = dp[nt][1] = dp[nt][2] = dp[nt][3] = 0.f;
      rowblock_times_rowsT(dp, da, Vs, lane);
    }
    float d0 = 0.f, d1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const float2 pa01 = unpack2(pa[nt >> 1][(nt & 1) * 2]);
      const float2 pa23 = unpack2(pa[nt >> 1][(nt & 1) * 2 + 1]);
      d0 += pa01.x * dp[nt][0] + pa01.y * dp[nt][1];
      d1 += pa23.x * dp[nt][2] + pa23.y * dp[nt][3];
    }
    d0 += __shfl_xor_sync(0xffffffffu, d0, 1);
    d0 += __shfl_xor_sync(0xffffffffu, d0, 2);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 1);
    d1 += __shfl_xor_sync(0xffffffffu, d1, 2);
    if (t == 0) {
      delta[mt * 16 + g] = d0;
      delta[mt * 16 + g + 8] = d1;
    }
#pragma unroll
    for (int nt = 0; nt < NT8; ++nt) {
      const float2 pa01 = unpack2(pa[nt >> 1][(nt & 1) * 2]);
      const float2 pa23 = unpack2(pa[nt >> 1][(nt & 1) * 2 + 1]);
      pa[nt >> 1][(nt & 1) * 2] = pack2(pa01.x * (dp[nt][0] - d0) * scale, pa01.y * (dp[nt][1] - d0) * scale);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(pa23.x * (dp[nt][2] - d1) * scale, pa23.y * (dp[nt][3] - d1) * scale);
    }
    float dq[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dq[n][0] = dq[n][1] = dq[n][2] = dq[n][3] = 0.f;
    frag_times_rows(dq, pa, Ks, lane);  // dQ = dS * K
    store_tile(dq, stage, gbase, ld, mt * 16, tokens, lane);
  }
  __syncthreads();

  // ---------------- pass B: column blocks -> dV, dK ----------------
  for (int jt = warp; jt < MT; jt += ATT_WARPS) {
    uint32_t pt[MT][4];
    {
      uint32_t ka[4][4];
      load_a_frags(ka, Ks, jt * 16, lane);
      float st[NT8][4];
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) st[nt][0] = st[nt][1] = st[nt][2] = st[nt][3] = 0.f;
      rowblock_times_rowsT(st, ka, Qs, lane);  // S^T[j][i] = K_j . Q_i
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const float2 l = *reinterpret_cast<const float2*>(lse2 + nt * 8 + 2 * t);
        pt[nt >> 1][(nt & 1) * 2] = pack2(exp2f(st[nt][0] * sl2 - l.x), exp2f(st[nt][1] * sl2 - l.y));
        pt[nt >> 1][(nt & 1) * 2 + 1] = pack2(exp2f(st[nt][2] * sl2 - l.x), exp2f(st[nt][3] * sl2 - l.y));
      }
    }
    {
      float dv[8][4];
#pragma unroll
      for (int n = 0; n < 8; ++n) dv[n][0] = dv[n][1] = dv[n][2] = dv[n][3] = 0.f;
      frag_times_rows(dv, pt, Os, lane);  // dV = P^T * dO
      store_tile(dv, stage, gbase + 2 * D, ld, jt * 16, tokens, lane);
    }
    {
      float dpt[NT8][4];
      uint32_t va[4][4];
      load_a_frags(va, Vs, jt * 16, lane);
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) dpt[nt][0] = dpt[nt][1] = dpt[nt][2] = dpt[nt][3] = 0.f;
      rowblock_times_rowsT(dpt, va, Os, lane);  // dP^T[j][i] = V_j . dO_i
#pragma unroll
      for (int nt = 0; nt < NT8; ++nt) {
        const float2 dl = *reinterpret_cast<const float2*>(delta + nt * 8 + 2 * t);
        const float2 p01 = unpack2(pt[nt >> 1][(nt & 1) * 2]);
        const float2 p23 = unpack2(pt[nt >> 1][(nt & 1) * 2 + 1]);
        pt[nt >> 1][(nt & 1) * 2] = pack2(p01.x * (dpt[nt][0] - dl.x) * scale, p01.y * (dpt[nt][1] - dl.y) * scale);
        pt[nt >> 1][(nt & 1) * 2 + 1] = pack2(p23.x * (dpt[nt][2] - dl.x) * scale, p23.y * (dpt[nt][3] - dl.y) * scale);
      }
    }
    float dk[8][4];
#pragma unroll
    for (int n = 0; n < 8; ++n) dk[n][0] = dk[n][1] = dk[n][2] = dk[n][3] = 0.f;
    frag_times_rows(dk, pt, Qs, lane);  // dK = dS^T * Q
    store_tile(dk, stage, gbase + D, ld, jt * 16, tokens, lane);
  }
}

static constexpr int FWD_SMEM = 3 * TPAD * LDS * 2;
static constexpr int BWD_SMEM = 4 * TPAD * LDS * 2 + ATT_WARPS * 16 * LDS * 2 + 2 * TPAD * 4;

int attention_fwd(const bf16* qkv, bf16* out, int batch, int tokens, int heads, cudaStream_t stream) {
  if (tokens > TPAD || tokens < 1) {
    set_error("attention_fwd: tokens=%d unsupported (max %d)", tokens, TPAD);
    return 1;
  }
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    attr = true;
  }
  const float sl2 = 1.4426950408889634f / sqrtf(static_cast<float>(HD));
  attn_fwd_kernel<<<batch * heads, ATT_THREADS, FWD_SMEM, stream>>>(qkv, out, tokens, heads, sl2);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int attention_bwd(const bf16* qkv, const bf16* dout, bf16* dqkv, int batch, int tokens, int heads,
                  cudaStream_t stream) {
  if (tokens > TPAD || tokens < 1) {
    set_error("attention_bwd: tokens=%d unsupported (max %d)", tokens, TPAD);
    return 1;
  }
  static bool attr = false;
  if (!attr) {
    VITATK_CUDA_OK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr = true;
  }
  const float scale = 1.0f / sqrtf(static_cast<float>(HD));
  attn_bwd_kernel<<<batch * heads, ATT_THREADS, BWD_SMEM, stream>>>(qkv, dout, dqkv, tokens, heads, scale,
                                                                    scale * 1.4426950408889634f);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

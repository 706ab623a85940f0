// Internal declarations shared by the kernels and the host-side engine (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vitatk {

typedef __nv_bfloat16 bf16;

// ---------------------------------------------------------------------------------------------
// error plumbing (no C++ exceptions cross the ABI)
// ---------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();
#define VITATK_CUDA_OK(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      ::vitatk::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 1;                                                                           \
    }                                                                                     \
  } while (0)

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch: the hot-path kernels are launched with programmatic stream serialization, call
// pdl_wait() before their first global-memory access (the previous kernel in the stream has then completed and its
// writes are visible) and pdl_launch_dependents() right after, so the NEXT kernel's launch latency and prologue
// (barrier init, TMEM allocation, descriptor prefetch) overlap this kernel's execution / tail.  Measured on B200 it
// changes nothing for the ViT PGD step (its ~2700 launches run 100 us each), so there the launch attribute is opt-in:
// VITATK_PDL=1.  The Swin path is the opposite case -- 560 launches of 10-50 us per iteration, a quarter of which is the
// per-kernel ramp -- and turns it on for its own launches (PdlScope; VITATK_PDL=0 turns it off): 20.5 -> 19.6 ms per
// iteration with only the GEMMs taking part.  Without the attribute griddepcontrol.* are no-ops.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#endif
bool pdl_enabled();
// Engine-scoped default: while a PdlScope(true) is alive on this thread, pdl_enabled() is true unless VITATK_PDL=0.
struct PdlScope {
  explicit PdlScope(bool on);
  ~PdlScope();
  int saved;
};

// cudaFuncSetAttribute (the > 48 KB dynamic shared memory opt-in) applies to the CURRENT device only, so "done once" has
// to be remembered per device: a process may hold engines on several GPUs (Engine(device="cuda:1")).
struct PerDeviceOnce {
  unsigned long long done = 0;
  bool need() {
    int d = 0;
    cudaGetDevice(&d);
    const unsigned long long bit = 1ull << (d & 63);
    if (done & bit) return false;
    done |= bit;
    return true;
  }
};

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                              int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// ---------------------------------------------------------------------------------------------
// tcgen05 GEMM:  out[M,N] = epi( A[M,K] * B[N,K]^T  +  T[M,64*nkb] * LB[N,64*nkb]^T )
// All operands bf16 row-major with the reduction dimension contiguous (K-major), fp32 accumulate
// in TMEM.  The second product is the rank-r LoRA term: T = x*A_lora^T (precomputed, bf16) and
// LB = s*B_lora zero-padded to 64 columns; it is issued as extra k-blocks of the same accumulator
// so the adapter lands in the W*x tile before the epilogue reads it.
// ---------------------------------------------------------------------------------------------
enum EpiMode : int {
  EPI_PLAIN = 0,      // out = acc (+ bias)
  EPI_RESIDUAL = 1,   // out = acc + bias + res
  EPI_GELU_DUAL = 2,  // u = acc + bias ; out = gelu(u) ; out2 = gelu'(u)   (fc1 forward; gelu' is what backward needs)
  EPI_MUL = 3,        // out = acc * res                                  (fc2 backward: dU = dG * gelu'(u))
  EPI_ROWTABLE = 4,   // out = acc + table[m % table_rows][n]            (patch embed: bias+pos / cls+pos)
  EPI_ROWDOT = 5      // out = acc ; rowdot[(m / rows) * (N/64) + n/64][m % rows] = sum over the 64-column slab of
                      // bf16(out) * res   (proj backward: attention's delta = rowsum(dO o O) per head; pair kernel only)
};

struct GemmEpilogue {
  int mode;
  const float* bias;    // [N] fp32 or null
  const bf16* res;      // [M, ld_res] residual (EPI_RESIDUAL) or saved gelu'(u) multiplier (EPI_MUL)
  int ld_res;
  const float* table;   // [table_rows, N] fp32 (EPI_ROWTABLE)
  int table_rows;
  // LayerNorm folded into the GEMM (any mode): the A operand is the RAW LayerNorm input h, B holds gamma o W, and the
  // epilogue first applies acc <- rstd[m] * (acc - mean[m] * c1[n]) (c1[n] = sum_k of the effective folded weight row,
  // LoRA included); the beta / bias part arrives through `bias`.  row_stats = (mean, rstd) per row, null = off.
  const float2* row_stats;
  const float* c1;
  // Single-CTA kernel with one N-tile (the skinny x*A^T GEMMs): while the rows of A stream through shared memory, four
  // otherwise idle epilogue warps also compute the LayerNorm statistics (mean, rstd) of every A row over the full K
  // and write them here -- the folded LayerNorm then needs no separate pass over h at all.  null = off.
  float2* stats_out;
  float stats_eps;
  float* rowdot;        // EPI_ROWDOT side output, [(M / rowdot_rows) * (N / 64), rowdot_pad] fp32
  int rowdot_rows;      // rows per group (tokens per image)
  int rowdot_pad;       // row pitch of the side output (208)
  // Constants through the tensor core.  A consumer GEMM can take its per-column constants (bias, the LayerNorm-fold
  // c1 / c2) as rank-1 updates inside its LoRA k-block instead of loading them in the epilogue: LB carries
  // [c1_hi, c1_lo, c1_hi, c2_hi, c2_lo, c2_hi] in columns stat_col.. of every 64-column group and the producer of T
  // (the skinny GEMM that computes stats_out) writes the matching per-row factors
  // [-mean_hi, -mean_hi, -mean_lo, sigma_hi, sigma_hi, sigma_lo] (sigma = 1 / rstd, hi/lo = bf16 split) into T there.
  // Producer side: stat_col > 0 (with stats_out) enables the write.  Consumer side: row_stats != null && c1 == null
  // means "scale by rstd only".  0 = off.
  int stat_col;
};

// T-tiles (pair kernel only): the GEMM computes its OWN LoRA down-projection T = A * TB^T inside the launch instead of
// reading a T that a separate skinny GEMM produced.  Every 256-row M-block gets one extra work item ("T-tile": the same
// A rows against the n-row adapter operand TB, full K) scheduled right before one of the block's output tiles; its rows
// go to `out` with plain global stores, a gpu-scope release on flags[m] publishes them, and the output tiles of the
// block acquire that flag before their TMA load of the LoRA k-block (which comes last in their k-loop, so they rarely
// wait).  The A block is streamed once from HBM for the T-tile and the block's output tiles together.
struct GemmTT {
  int n;                 // 0 = off; 32 or 64: columns of T produced (>= every column the LoRA k-steps read)
  const bf16* tb;        // [64, K] bf16 row-major (rows >= n unused): the adapter's down-projection, K-major
  int ld_tb;
  bf16* out;             // T [M, ld_out]  (must be the tensor the plan's LoRA k-block reads)
  int ld_out;
  const float* bias;     // [64] fp32 added to T's columns (the "ones" that feed bias columns of LB), or null
  unsigned int* flags;   // [2 * ceil(M / 256)] zero-initialised; the kernel leaves it zeroed again
  // LayerNorm statistics of A's rows from partial (mean, M2) records written by the producer of A (GemmEpilogue::
  // stats_partial_out): the T-tile's epilogue combines them, writes stats_out and the per-row constant factors
  const float2* stats_partial;  // [M, stats_nparts] or null
  int stats_nparts;
  int stats_part_cols;          // columns each partial record covers
  float2* stats_out;            // [M] (mean, rstd)
  float stats_eps;
  int stat_col;                 // first of the six constant-factor columns in T (0 = none)
};

struct GemmPlan {
  // shapes
  int M, N, K;
  int BN;               // 64, 128, 192 or 256
  int reverse_m;        // 1: M-blocks are processed last-to-first (set by the engine on the skinny LoRA GEMMs)
  int two_cta;          // 1: CTA-pair kernel (cta_group::2, 256-row tiles, half of B per CTA); BN = 256 only
  int lora_nkb;         // extra 64-wide k-blocks (0 = no adapter)
  int lora_ksteps;      // UMMA k-steps (16 each) issued per extra k-block = ceil(r/16)
  int lora_group_cols;  // >0: T column offset = (n0 / lora_group_cols) * 64  (fused q|k|v forward)
  // 16-bit storage formats (0 = bf16, 1 = IEEE fp16): the A and B operands of the main k-blocks (tcgen05 kind::f16 needs
  // both in the same format, so the weights that meet an fp16 residual stream are packed as fp16 too; T / LB of the LoRA
  // k-blocks stay bf16), the output, and the residual read by EPI_RESIDUAL.  Set by the engine after gemm_plan_init.
  int a_f16, out_f16, res_f16;
  GemmEpilogue epi;
  GemmTT tt;
  // tensor maps (built once per plan)
  CUtensorMap tmA, tmB, tmLA, tmLB, tmOut, tmOut2;
  CUtensorMap tmTB;     // T-tile B operand: (tt.n / 2)-row x 64-col boxes of tt.tb
  CUtensorMap tmAux;    // residual / multiplier tensor as 64-col x 128-row slabs (pair kernel only)
};

// Build the TMA descriptors of a plan. Pointers may be null when the feature is unused.
int gemm_plan_init(GemmPlan* p, int M, int N, int K, const bf16* A, int lda, const bf16* B, int ldb,
                   bf16* out, int ldo, bf16* out2, int ldo2, const bf16* T, int ldt, const bf16* LB, int ldlb,
                   int lora_nkb, int lora_ksteps, int lora_group_cols, GemmEpilogue epi, const GemmTT* tt = nullptr);
int gemm_launch(const GemmPlan* p, cudaStream_t stream, int num_sms);

int make_tmap_3d(CUtensorMap* tm, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                 uint64_t s2_bytes, uint32_t b0, uint32_t b1, int swizzle_bytes = 128);

// ---------------------------------------------------------------------------------------------
// attention (softmax(QK^T/sqrt(d))V per (image, head)), qkv packed [M, 3*D] token-major
// ---------------------------------------------------------------------------------------------
// tcgen05 forward: S = QK^T and O = PV on the 5th-gen tensor cores, P kept in TMEM as the A operand.
struct AttnFwdPlan {
  int batch, tokens, heads;
  const bf16* qkv;
  bf16* out;
  float* lse2;  // [batch*heads, 208] log2-domain logsumexp per query (for the backward), may be null
  CUtensorMap tmQ, tmKV, tmO;
};
int attention_fwd_plan_init(AttnFwdPlan* p, const bf16* qkv, bf16* out, float* lse2, int batch, int tokens, int heads);
int attention_fwd_tc05(const AttnFwdPlan* p, cudaStream_t stream);
// tcgen05 backward (single fused pass, attention_bwd_fused.cu); lse2 comes from the forward.
struct AttnBwdPlan {
  int batch, tokens, heads;
  const bf16 *dout, *o;
  const float* lse2;
  float* delta;  // [batch*heads, 208] scratch
  bf16* dqkv;
  CUtensorMap tmQKV128, tmQKV208, tmDO208;
  CUtensorMap tmDqkv32;  // 32-column x 32-row store boxes, 64B swizzle
};
int attention_bwd_plan_init(AttnBwdPlan* p, const bf16* qkv, const bf16* dout, const bf16* o, const float* lse2,
                            float* delta, bf16* dqkv, int batch, int tokens, int heads);
// compute_delta = false when delta was already produced by the proj-backward GEMM (EPI_ROWDOT)
int attention_bwd_fused(const AttnBwdPlan* p, cudaStream_t stream, bool compute_delta);
// In-kernel clock64 timelines (timing experiments only): compiled in with -DVITATK_DBG_KERNELS, otherwise they fail.
int gemm_set_trace(long long* dev_buf);
int attention_bwd_set_trace(long long* dev_buf);
int attention_fwd_set_trace(long long* dev_buf);

// ---------------------------------------------------------------------------------------------
// LayerNorm / head / PGD kernels (HBM-bound)
// ---------------------------------------------------------------------------------------------
// (mean, rstd) per row only: the read-only half of LayerNorm, for consumers that fold the normalisation into a GEMM
// x_f16 / g_f16: the LayerNorm input x (forward residual stream) resp. dres / dx_out (backward residual stream) hold IEEE
// fp16 instead of bf16 bit patterns (same 16-bit storage; the pointers stay typed bf16*)
int layernorm_stats(const bf16* x, float2* stats, int rows, int cols, float eps, cudaStream_t stream, int x_f16 = 0);
int layernorm_fwd(const bf16* x, const float* gamma, const float* beta, bf16* y, float2* stats, int rows,
                  int cols, float eps, cudaStream_t stream, int x_f16 = 0);
// dx_out = dres + LN_backward(dy) ; x is the saved LN input, stats = (mean, rstd)
int layernorm_bwd(const bf16* dy, const bf16* x, const float2* stats, const float* gamma, const bf16* dres,
                  bf16* dx_out, int rows, int cols, cudaStream_t stream, int x_f16 = 0, int g_f16 = 0);
constexpr int LN_BT_MAX_KSTEPS = 4;  // ranks <= 64
// layernorm_bwd that also writes T[rows, 16 * ksteps] = dx * LB^T (LB [16 * ksteps, 768] in dx's 16-bit format): the LoRA
// down-projection the next backward GEMM needs, from rows the kernel already holds (cols == 768)
int layernorm_bwd_bt(const bf16* dy, const bf16* x, const float2* stats, const float* gamma, const bf16* dres, bf16* dx_out,
                     int rows, int cols, cudaStream_t stream, int x_f16, int g_f16, const bf16* LB, int ksteps, bf16* T, int ldt);
// final LN on CLS rows + classifier + softmax-CE; writes logits, per-image loss, and (optionally) the
// gradient wrt the final hidden state (non-CLS rows zero-filled).  dlogits == nullptr: the cotangent is
// softmax - onehot (cross-entropy); otherwise the caller's [batch, classes] fp32 cotangent (vector-Jacobian product).
int head_fwd_bwd(const bf16* h, const float* gamma, const float* beta, const float* Wc, const float* bc,
                 const int64_t* labels, float* logits, float* loss, bf16* dh, int batch, int tokens, int dim,
                 int classes, float eps, float grad_scale, cudaStream_t stream, const float* dlogits = nullptr,
                 int h_f16 = 0, int dh_f16 = 0, float* y_out = nullptr, float* dlogits_out = nullptr);

struct PixelNorm {
  float mean[3];
  float inv_std[3];
};
// adv <- clamp(x0 + noise, 0, 1) (noise may be null; or counter-based U(-eps,eps) when seed_enable),
// and writes the normalised im2col rows of adv (bf16) for the patch-embed GEMM.
int pgd_init(const float* x0, const float* noise, float* adv, bf16* cols, int batch, PixelNorm nrm, float eps,
             int use_rng, uint64_t seed, uint64_t image_index0, cudaStream_t stream);
// one fused PGD update: sign step + Linf projection + [0,1] clamp + renormalised im2col for the next step
int pgd_update(const bf16* dcols, const float* x0, float* adv, bf16* cols, int batch, PixelNorm nrm,
               float eps, float alpha, cudaStream_t stream);
// materialise dL/dx (fp32 NCHW) from the im2col-layout gradient
int grad_to_image(const bf16* dcols, float* grad, int batch, PixelNorm nrm, float scale, cudaStream_t stream);
// finalize-time packing of the tensor-core constant columns (GemmEpilogue::stat_col) into a copy of lb [rows, 64]
int lora_const_columns(const bf16* lb, const float* c1, const float* c2, bf16* out, int rows, int col,
                       cudaStream_t stream);
// save_images + reload (Utils.py:106-113): out fp32 NCHW = trunc(clamp(x)*255)/255 and / or the uint8 HWC image itself
int png_roundtrip(const float* images, float* out, uint8_t* u8_hwc, int batch, cudaStream_t stream);
// counts[0] += #(argmax(logits)==label), counts[1] += batch
int count_correct(const float* logits, const int64_t* labels, int batch, int classes, long long* counts,
                  cudaStream_t stream);


// ---------------------------------------------------------------------------------------------
// LoRA training step (train.cu): dropout-aware adapter kernels, weight gradients, Adam, operand re-packing
// ---------------------------------------------------------------------------------------------
struct LoraDxAdapter {
  const float* A;   // fp32 master [r, K]
  float scale;      // alpha / r
  int r, c0;        // rank, first column of this adapter's dY*B inside BT
  uint32_t seed;    // dropout mask key
};
struct LoraDxArgs {
  LoraDxAdapter ad[3];
  int n;
};
// T[m, c0 + j] = sum_k drop(x[m, k]) * A[j, k]  (mask keyed by seed over element index (row0 + m) * K + k; p = 0: no mask)
int lora_down(const bf16* x, int ldx, int K, const float* A, int r, bf16* T, int ldt, int c0, int rows, uint32_t seed,
              float p, long long row0, cudaStream_t stream);
// dX (+)= sum_a drop'_a o (BT_a * s_a A_a), then * mul (optional)
int lora_dx(bf16* dX, int ldx, int K, const bf16* BT, int ldt, const LoraDxArgs& args, const bf16* mul, int ldm, int rows,
            int accumulate, float p, long long row0, cudaStream_t stream);
// G = scale * X^T S over the token rows (G [N, r], or [r, N] with transpose); X optionally masked (p > 0) with element
// index (row0 + m) * mask_ld + n; x_f16: X holds IEEE fp16.  partial: fp16-free fp32 scratch [ceil(rows / 256), N, 16].
int wgrad(const bf16* X, int ldx, int N, const bf16* S, int lds, int c0, int r, int rows, float* partial, float scale,
          int transpose, float* G, uint32_t seed, float p, long long row0, int mask_ld, int x_f16, cudaStream_t stream);
int head_wgrad(const float* y, const float* dlogits, int batch, int dim, int classes, float scale, float* dW, float* db,
               cudaStream_t stream);
int adam_step(float* p, const float* g, float* m, float* v, long long n, float lr, float b1, float b2, float eps, int step,
              cudaStream_t stream);
// fp32 masters of one adapter -> its slots in the site's packed 16-bit operand buffers (vitatk_set_lora layouts);
// fmt bit 0: la_fwd is fp16, bit 1: lb_bwd is fp16 (operands that meet an fp16 residual stream), else bf16
int lora_repack(const float* A, const float* B, int r, int in, int out, float s, bf16* la_fwd, bf16* lb_fwd, bf16* lb_bwd,
                bf16* la_bwd, int row0, int col0, int out0, int ld_lbb, int ld_lab, const float* gamma, int fmt,
                cudaStream_t stream);
uint32_t train_mask_seed(uint64_t seed, uint64_t step, int layer, int adapter);

// ---------------------------------------------------------------------------------------------
// adversarial patch / EOT front end (patch.cu)
// ---------------------------------------------------------------------------------------------
// composite of `samples` = images x T transformed copies of one patch [3, p, p]; tf [samples, 6] = inverse affine (output-
// normalised -> patch-normalised).  Writes the normalised im2col rows (cols) and / or fp32 NCHW images (out_img).
int patch_apply(const float* images, const float* patch, int p, const float* tf, int T, int samples, int circle, PixelNorm nrm,
                bf16* cols, float* out_img, cudaStream_t stream);
// grad [3, p, p] += scale * d loss / d patch from dcols (gradient w.r.t. the normalised im2col rows); fw = forward affine;
// partial: fp32 scratch [samples, 3, p, p]
int patch_grad(const bf16* dcols, const float* tf, const float* fw, int p, int samples, int circle, PixelNorm nrm, float scale,
               float* partial, float* grad, cudaStream_t stream);
int patch_update(float* patch, const float* grad, float* m, float* v, int n, float lr, float dir, int step, float b1, float b2,
                 float eps, cudaStream_t stream);
int repeat_labels(const int64_t* labels, int T, int samples, int64_t* out, cudaStream_t stream);

}  // namespace vitatk

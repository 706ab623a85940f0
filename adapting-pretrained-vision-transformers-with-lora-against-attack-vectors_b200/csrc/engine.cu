// Host-side engine behind the C ABI (include/vitatk.h): weight registry, activation workspace laid out
// for B200 HBM, per-batch GEMM plans (TMA descriptors), and the forward / input-gradient / PGD drivers.
//
// HBM layout (M = batch * 197 tokens, all bf16 unless noted):
//   saved for backward, per layer l : h_in[l] [M,768], h_mid[l] [M,768], qkv[l] [M,2304], u[l] = gelu'(fc1) [M,3072],
//                                     stats1[l], stats2[l] float2[M]          (~13.8 KB / token / layer)
//                                     ao[l] [M,768] (attention out), lse2[l] float[B*12*208]
//   transient scratch               : cols [M,768] (normalised im2col), xn [M,768] (LN out), g [M,3072] (GELU out), T [M,192] (LoRA x*A^T),
//                                     dh_a/dh_b [M,768], du [M,3072], dxn [M,768], dao [M,768], dqkv [M,2304]
// No weight gradients exist on this path (autograd.grad w.r.t. the input only), so GEMM inputs are never
// saved for dW; only what LN / GELU / softmax need for their Jacobians is kept.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <memory>
#include <vector>

#include "../../include/vitatk.h"
#include "vitatk_internal.h"

namespace vitatk {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

static thread_local int g_pdl_scope = 0;
bool pdl_enabled() {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("VITATK_PDL");  // ViT PGD step on B200: no gain (937.8 vs 941.6 adv img/s), so opt-in there
    v = !e ? -1 : (e[0] == '1' ? 1 : 0);
  }
  return v == 1 || (v == -1 && g_pdl_scope > 0);
}
PdlScope::PdlScope(bool on) : saved(g_pdl_scope) { g_pdl_scope = on ? 1 : 0; }
PdlScope::~PdlScope() { g_pdl_scope = saved; }

static constexpr int TOKENS = 197;
static constexpr int PDL_MAX_BATCH = 32;
static constexpr int LORA_PAD = 64;

struct LoraSite {
  int rank = 0;
  const bf16 *la_fwd = nullptr, *lb_fwd = nullptr, *lb_bwd = nullptr, *la_bwd = nullptr;
  // Constants through the tensor core (GemmEpilogue::stat_col): engine-owned copy of lb_fwd whose columns ccol.. carry
  // the site's bias / LayerNorm-fold constants, and (plain-bias sites) the "bias" vector that makes the skinny GEMM
  // write 1.0 into T's matching columns.  ccol == 0: off (the epilogue loads the constants itself).
  int ccol = 0;
  bool cfold = false;
  // QKV site only: the q, k and v adapters share ONE 64-column group (q in columns [0, rq), k in [rq, rq + rk), v after
  // them; la_fwd [64, D], lb_fwd [3D, 64], lb_bwd [64, 3D], la_bwd [D, 64]) instead of one group each.  Host-selected
  // when rq + rk + rv <= 64: one LoRA k-block instead of three, and the site becomes eligible for T-tiles.
  bool packed = false;
  const bf16* lbx = nullptr;
  const float* tones = nullptr;
};

struct LayerWeights {
  const float *ln1_g = nullptr, *ln1_b = nullptr, *ln2_g = nullptr, *ln2_b = nullptr;
  const bf16 *qkv_w = nullptr, *qkv_wt = nullptr, *proj_w = nullptr, *proj_wt = nullptr;
  const bf16 *fc1_w = nullptr, *fc1_wt = nullptr, *fc2_w = nullptr, *fc2_wt = nullptr;
  const float *qkv_b = nullptr, *proj_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  // LayerNorm folded into the qkv / fc1 GEMMs (both set => this layer runs folded): qkv_w / fc1_w then hold gamma o W,
  // qkv_b / fc1_b hold c2 = W beta + LoRA(beta) + b, the QKV / FC1 adapter la_fwd holds gamma o A, and these hold c1
  const float *qkv_c1 = nullptr, *fc1_c1 = nullptr;
  LoraSite lora[4];
};

struct LayerPlans {
  // forward
  GemmPlan t_qkv, qkv, t_proj, proj, t_fc1, fc1, t_fc2, fc2;
  // backward (input gradients)
  GemmPlan bt_fc2, bfc2, bt_fc1, bfc1, bt_proj, bproj, bt_qkv, bqkv;
  // the LayerNorm backward that produces the site's input also writes its dx * B^T (layernorm_bwd_bt): no skinny launch
  bool ln_bt_proj = false, ln_bt_fc2 = false;
};

struct PlanSet {
  int batch = 0;
  GemmPlan patch, bpatch;
  std::vector<LayerPlans> layers;
  std::vector<AttnFwdPlan> attn_fwd;
  std::vector<AttnBwdPlan> attn_bwd;
};

}  // namespace vitatk

using namespace vitatk;

struct vitatk_engine {
  vitatk_config cfg;
  int num_sms = 148;
  bool finalized = false;
  // global weights
  const bf16 *patch_w = nullptr, *patch_wt = nullptr;
  const float *embed_table = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *head_w = nullptr, *head_b = nullptr;
  std::vector<LayerWeights> lw;
  // workspace
  char* ws = nullptr;
  long long ws_bytes = 0;
  std::vector<bf16*> h;      // [layers+1]
  std::vector<bf16*> h_mid;  // [layers]
  std::vector<bf16*> qkv;    // [layers]
  std::vector<bf16*> u;      // [layers]
  std::vector<float2*> st1, st2;
  std::vector<bf16*> ao;     // [layers] attention output (saved: delta = rowsum(dO o O) in the backward)
  std::vector<float*> lse2;  // [layers] log2-domain logsumexp per (image, head, query)
  float* delta = nullptr;
  bf16 *cols = nullptr, *xn = nullptr, *g = nullptr, *T = nullptr;
  bf16 *dh_a = nullptr, *dh_b = nullptr, *du = nullptr, *dxn = nullptr, *dao = nullptr, *dqkv = nullptr;
  float *logits = nullptr, *loss = nullptr, *scratch_img = nullptr;
  int64_t* labels_rep = nullptr;  // [max_batch] per-sample labels of the EOT front end
  std::map<int, PlanSet*> plans;
  long long launches = 0;
  PixelNorm nrm;
  // optional per-launch CUDA-event timing (bench.py's roofline leg; off in the timed region)
  bool zigzag = true;                // skinny LoRA GEMMs walk M last-to-first (VITATK_ZIGZAG=0: first-to-last)
  bool tc_const = true;              // bias / fold constants as tensor-core rank-1 updates (VITATK_TC_CONST=0: epilogue loads)
  char* cbuf = nullptr;              // backing store of the lbx / tones arrays
  bool const_dirty = true;           // weights / adapters changed since the constant columns were packed
  bool fuse_stats = true;            // folded LayerNorm: (mean, rstd) come out of the skinny LoRA GEMM (VITATK_FUSE_STATS=0: stats kernel)
  // T-tiles per site, VITATK_TT_SITES bit mask: 1 proj, 2 fc2, 4 fc2-bwd, 8 fc1-bwd, 16 proj-bwd, 32 qkv-bwd.  Default 0:
  // measured inside the PGD step on B200 (one box, back to back: mask 0 / 17 / 51 / 63 = 245.9 / 246.7 / 247.5 / 249.4 ms)
  // the T-tiles cost as much as the skinny GEMMs they replace -- a T-tile streams its A block at the ring-depth x latency
  // bound (~680 clk per k-block whatever its size) and unbalances the static tile schedule, while the separate skinny
  // GEMM finds its input L2-hot (zig-zag) and uses all 148 SMs.  The mechanism stays available and tested.
  int tt_sites = 0;
  // Residual streams in IEEE fp16: h / h_mid (forward) and dh_a / dh_b (backward) are the only tensors
  // whose 16-bit rounding accumulates over all 24 residual adds; fp16 has three more mantissa bits than bf16 at the same
  // size.  tcgen05 kind::f16 needs A and B in one format, so the host packs the operands that multiply a stream as fp16
  // too (folded qkv / fc1 weights and their adapters' A; fc2^T, proj^T, patch^T and the adapters' B^T for the backward).  The backward stream carries gradients
  // scaled by grad_S (a power of two, undone when the image gradient is materialised; sign() never sees it) so that they
  // sit in fp16's normal range.  VITATK_RES_F16=0: bf16 streams (round-1 behaviour).
  bool res_f16 = true;
  float grad_S = 256.0f;
  bool fuse_tt = true;               // plain LoRA sites: T = x*A^T comes from T-tiles inside the consumer GEMM (VITATK_TT=0: skinny GEMMs)
  unsigned int* tt_flags = nullptr;  // [2 * ceil(max M / 256)] inter-CTA flags of the T-tiles (zero between launches)
  bool fuse_delta = false;           // delta comes out of the proj-backward GEMM epilogue (pair kernel) instead of a kernel
  bool fuse_ln_bt = false;           // VITATK_LN_BT=1: LayerNorm backward also writes the next LoRA site's dx * B^T (no bt_proj / bt_fc2 launch)
  // ---- LoRA training (vitatk_train_*) ----
  struct TrainAdapter {
    int rank = 0;
    float scale = 0.f;
    long long off_a = 0, off_b = 0;  // offsets of A [r, in] and B [out, r] in the flat fp32 parameter / gradient buffers
    int col0 = 0;                    // first column of this adapter inside its site's 64-column group
  };
  struct TrainLayerPlans {
    GemmPlan fc1_t, fc2_t;                                   // forward with the per-layer GELU output buffer
    GemmPlan bt_fc2, bt_fc1, bt_proj, bt_qkv;                // BT = dY * B (skinny)
    GemmPlan bfc2_nl, bfc1_nl, bproj_nl, bqkv_nl;            // frozen-weight input gradients WITHOUT the LoRA k-block
  };
  struct TrainState {
    bool enabled = false;
    float p_drop = 0.f;
    float* params = nullptr;
    float* grads = nullptr;
    long long n = 0, off_cw = -1, off_cb = -1;
    std::vector<std::vector<TrainAdapter>> ad;  // [layer][6]: q, k, v, proj, fc1, fc2
    std::vector<bf16*> g_layer;                 // per-layer GELU output (input of the fc2 adapter, needed for its dA)
    bf16* Td = nullptr;                         // [M, 64] recomputed dropout(x) A^T for dB
    float* partial = nullptr;                   // weight-gradient partial sums
    float *ycls = nullptr, *dlog = nullptr;     // head: classifier input / output cotangent
    char* buf = nullptr;
    std::map<int, std::vector<TrainLayerPlans>*> plans;
  } tr;
  bool prof = false;
  struct ProfRec { int cat; double flops; cudaEvent_t a, b; };
  std::vector<ProfRec> prof_recs;
  std::vector<cudaEvent_t> prof_pool;
};

namespace vitatk {

// (Re)pack the tensor-core constant columns (LoraSite::lbx / tones) from the current weights and adapters.  Runs at the
// first use after finalize and again after any vitatk_set_tensor / vitatk_set_lora (they mark the engine dirty).
static int refresh_const_columns(vitatk_engine* e) {
  const vitatk_config& c = e->cfg;
  e->const_dirty = false;
  for (auto& w : e->lw)
    for (auto& ls : w.lora) ls.ccol = 0, ls.cfold = false, ls.lbx = nullptr, ls.tones = nullptr;
  if (!e->tc_const) return 0;
  const int site_rows[4] = {3 * c.dim, c.dim, c.mlp_dim, c.dim};
  if (!e->cbuf) {
    long long bytes = 0;
    for (int s = 0; s < 4; ++s) bytes += static_cast<long long>(site_rows[s]) * LORA_PAD * 2 + 3 * LORA_PAD * 4;
    bytes *= c.layers;
    VITATK_CUDA_OK(cudaMalloc(&e->cbuf, bytes));
    VITATK_CUDA_OK(cudaMemset(e->cbuf, 0, bytes));
  }
  VITATK_CUDA_OK(cudaDeviceSynchronize());  // nothing may still be reading the previous columns
  char* q = e->cbuf;
  for (int l = 0; l < c.layers; ++l) {
    LayerWeights& w = e->lw[l];
    const bool fold = w.qkv_c1 != nullptr && w.fc1_c1 != nullptr;
    const float* c1s[4] = {fold ? w.qkv_c1 : nullptr, nullptr, fold ? w.fc1_c1 : nullptr, nullptr};
    const float* c2s[4] = {w.qkv_b, w.proj_b, w.fc1_b, w.fc2_b};
    for (int s = 0; s < 4; ++s) {
      LoraSite& ls = w.lora[s];
      bf16* lbx = reinterpret_cast<bf16*>(q);
      q += static_cast<long long>(site_rows[s]) * LORA_PAD * 2;
      float* ones = reinterpret_cast<float*>(q);
      q += 3 * LORA_PAD * 4;
      const bool cf = c1s[s] != nullptr;
      const int need = cf ? 6 : 2;
      // folded sites take their per-row factors from the statistics the skinny GEMM computes in-kernel
      if (ls.rank <= 0 || ls.rank + need > 32 || (cf && !e->fuse_stats)) continue;
      ls.ccol = ls.rank;
      ls.cfold = cf;
      if (lora_const_columns(ls.lb_fwd, c1s[s], c2s[s], lbx, site_rows[s], ls.ccol, nullptr)) return 1;
      ls.lbx = lbx;
      if (!cf) {
        float host_ones[3 * LORA_PAD] = {0};
        for (int g = 0; g < 3; ++g) host_ones[g * LORA_PAD + ls.ccol] = host_ones[g * LORA_PAD + ls.ccol + 1] = 1.0f;
        VITATK_CUDA_OK(cudaMemcpy(ones, host_ones, sizeof(host_ones), cudaMemcpyHostToDevice));
        ls.tones = ones;
      }
    }
  }
  VITATK_CUDA_OK(cudaDeviceSynchronize());
  return 0;
}

static int lora_ksteps(int r) { return (r + 15) / 16; }
// forward k-steps of a site: its rank plus the constant columns the tensor core adds (6 folded, 2 plain bias)
static int site_ksteps(const LoraSite& s) { return lora_ksteps(s.ccol > 0 ? s.ccol + (s.cfold ? 6 : 2) : s.rank); }

static int build_plans(vitatk_engine* e, int batch, PlanSet** out) {
  if (e->const_dirty) {  // (set_tensor / set_lora already dropped the cached plans)
    for (auto& kv : e->plans) delete kv.second;
    e->plans.clear();
    if (refresh_const_columns(e)) return 1;
  }
  auto it = e->plans.find(batch);
  if (it != e->plans.end()) {
    *out = it->second;
    return 0;
  }
  const vitatk_config& c = e->cfg;
  const int M = batch * TOKENS, D = c.dim, F = c.mlp_dim;
  std::unique_ptr<PlanSet> owner(new PlanSet());  // released into e->plans on success, freed on any early return
  PlanSet* ps = owner.get();
  ps->batch = batch;
  ps->layers.resize(c.layers);
  ps->attn_fwd.resize(c.layers);
  ps->attn_bwd.resize(c.layers);
  GemmEpilogue plain = {};
  plain.mode = EPI_PLAIN;
  // patch embedding: h[0] = cols * Wpe^T + table[m % 197]
  {
    GemmEpilogue ep = {EPI_ROWTABLE, nullptr, nullptr, 0, e->embed_table, TOKENS};
    if (gemm_plan_init(&ps->patch, M, D, D, e->cols, D, e->patch_w, D, e->h[0], D, nullptr, 0, nullptr, 0, nullptr, 0, 0,
                       0, 0, ep))
      return 1;
  }
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->lw[l];
    LayerPlans& p = ps->layers[l];
    const LoraSite& sq = w.lora[VITATK_SITE_QKV];
    const LoraSite& sp = w.lora[VITATK_SITE_PROJ];
    const LoraSite& s1 = w.lora[VITATK_SITE_FC1];
    const LoraSite& s2 = w.lora[VITATK_SITE_FC2];
    const bool fold = w.qkv_c1 != nullptr && w.fc1_c1 != nullptr;  // LayerNorm folded into the qkv / fc1 GEMMs
    const bf16* a_ln1 = fold ? e->h[l] : e->xn;
    const bf16* a_ln2 = fold ? e->h_mid[l] : e->xn;
    // q|k|v adapters: one shared 64-column group (packed) or one group each
    const int q_tcols = sq.packed ? LORA_PAD : 3 * LORA_PAD;
    const int q_nkb_bwd = sq.rank > 0 ? (sq.packed ? 1 : 3) : 0;
    const int q_group_cols = (sq.rank > 0 && !sq.packed) ? D : 0;
    // T-tiles: the consumer GEMM computes its own T = x * A^T (GemmTT).  n = 32 when every column the LoRA k-steps read
    // fits, else 64.  Sites whose skinny GEMM also produces LayerNorm statistics (folded qkv / fc1) keep that GEMM.
    auto make_tt = [&](const LoraSite& ls, const bf16* tb, int ld_tb, int ksteps, const float* bias, bool ok) {
      GemmTT t = {};
      if (!e->fuse_tt || !ok || ls.rank <= 0) return t;
      t.n = ksteps * 16 <= 32 ? 32 : 64;
      t.tb = tb;
      t.ld_tb = ld_tb;
      t.out = e->T;
      t.ld_out = 3 * LORA_PAD;
      t.bias = bias;
      t.flags = e->tt_flags;
      return t;
    };
    // ---------------- forward ----------------
    if (attention_fwd_plan_init(&ps->attn_fwd[l], e->qkv[l], e->ao[l], e->lse2[l], batch, TOKENS, c.heads)) return 1;
    if (attention_bwd_plan_init(&ps->attn_bwd[l], e->qkv[l], e->dao, e->ao[l], e->lse2[l], e->delta, e->dqkv, batch,
                                TOKENS, c.heads))
      return 1;
    {
      GemmEpilogue ep = plain;
      if (fold && e->fuse_stats) {  // the skinny GEMM streams h anyway: it also produces LN1's (mean, rstd)
        ep.stats_out = e->st1[l];
        ep.stats_eps = c.ln_eps;
        if (sq.ccol > 0 && sq.cfold) ep.stat_col = sq.ccol;  // ... and the per-row factors of qkv's constants
      }
      if (sq.ccol > 0 && !sq.cfold) ep.bias = sq.tones;
      if (sq.rank > 0 &&
          gemm_plan_init(&p.t_qkv, M, q_tcols, D, a_ln1, D, sq.la_fwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, ep))
        return 1;
    }
    {
      GemmEpilogue ep = plain;
      ep.bias = w.qkv_b;
      if (fold) {
        ep.row_stats = e->st1[l];
        ep.c1 = w.qkv_c1;
      }
      if (sq.ccol > 0) ep.bias = nullptr, ep.c1 = nullptr;  // the constants ride in the LoRA k-block
      if (gemm_plan_init(&p.qkv, M, 3 * D, D, a_ln1, D, w.qkv_w, D, e->qkv[l], 3 * D, nullptr, 0, e->T, 3 * LORA_PAD,
                         sq.ccol > 0 ? sq.lbx : sq.lb_fwd, LORA_PAD, sq.rank > 0 ? 1 : 0, site_ksteps(sq),
                         q_group_cols, ep))
        return 1;
    }
    {
      GemmEpilogue ep = plain;
      if (sp.ccol > 0) ep.bias = sp.tones;  // T[:, ccol..ccol+1] = 1: proj's bias is added by the tensor core
      const GemmTT tt = make_tt(sp, sp.la_fwd, D, site_ksteps(sp), ep.bias, e->tt_sites & 1);
      p.t_proj.M = 0;  // M == 0: not launched (T comes from the consumer's T-tiles)
      if (sp.rank > 0 && tt.n == 0 &&
          gemm_plan_init(&p.t_proj, M, LORA_PAD, D, e->ao[l], D, sp.la_fwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, ep))
        return 1;
      GemmEpilogue ep2 = {EPI_RESIDUAL, sp.ccol > 0 ? nullptr : w.proj_b, e->h[l], D, nullptr, 0};
      if (gemm_plan_init(&p.proj, M, D, D, e->ao[l], D, w.proj_w, D, e->h_mid[l], D, nullptr, 0, e->T, 3 * LORA_PAD,
                         sp.ccol > 0 ? sp.lbx : sp.lb_fwd, LORA_PAD, sp.rank > 0 ? 1 : 0, site_ksteps(sp), 0, ep2, &tt))
        return 1;
    }
    {
      GemmEpilogue ep = plain;
      if (fold && e->fuse_stats) {
        ep.stats_out = e->st2[l];
        ep.stats_eps = c.ln_eps;
        if (s1.ccol > 0 && s1.cfold) ep.stat_col = s1.ccol;
      }
      if (s1.ccol > 0 && !s1.cfold) ep.bias = s1.tones;
      if (s1.rank > 0 &&
          gemm_plan_init(&p.t_fc1, M, LORA_PAD, D, a_ln2, D, s1.la_fwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, ep))
        return 1;
    }
    {
      GemmEpilogue ep = plain;
      ep.mode = EPI_GELU_DUAL;
      ep.bias = w.fc1_b;
      if (fold) {
        ep.row_stats = e->st2[l];
        ep.c1 = w.fc1_c1;
      }
      if (s1.ccol > 0) ep.bias = nullptr, ep.c1 = nullptr;
      if (gemm_plan_init(&p.fc1, M, F, D, a_ln2, D, w.fc1_w, D, e->g, F, e->u[l], F, e->T, 3 * LORA_PAD,
                         s1.ccol > 0 ? s1.lbx : s1.lb_fwd, LORA_PAD, s1.rank > 0 ? 1 : 0, site_ksteps(s1), 0, ep))
        return 1;
    }
    {
      GemmEpilogue ep = plain;
      if (s2.ccol > 0) ep.bias = s2.tones;
      const GemmTT tt = make_tt(s2, s2.la_fwd, F, site_ksteps(s2), ep.bias, e->tt_sites & 2);
      p.t_fc2.M = 0;
      if (s2.rank > 0 && tt.n == 0 &&
          gemm_plan_init(&p.t_fc2, M, LORA_PAD, F, e->g, F, s2.la_fwd, F, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, ep))
        return 1;
      GemmEpilogue ep2 = {EPI_RESIDUAL, s2.ccol > 0 ? nullptr : w.fc2_b, e->h_mid[l], D, nullptr, 0};
      if (gemm_plan_init(&p.fc2, M, D, F, e->g, F, w.fc2_w, F, e->h[l + 1], D, nullptr, 0, e->T, 3 * LORA_PAD,
                         s2.ccol > 0 ? s2.lbx : s2.lb_fwd, LORA_PAD, s2.rank > 0 ? 1 : 0, site_ksteps(s2), 0, ep2, &tt))
        return 1;
    }
    // ---------------- backward ----------------
    // The residual-stream gradient ping-pongs: layer l receives it in dh_in(l) and leaves dh_out(l).
    // dh entering layer l (grad wrt h[l+1]) lives in dh_a; dh_mid in dh_b; result (grad wrt h[l]) in dh_a.
    {
      const GemmTT tt = make_tt(s2, s2.lb_bwd, D, lora_ksteps(s2.rank), nullptr, e->tt_sites & 4);
      p.bt_fc2.M = 0;
      // fused into the LayerNorm backward that writes dh_a (LN1 of layer l + 1); the last layer's dh_a comes from the head
      p.ln_bt_fc2 = e->fuse_ln_bt && s2.rank > 0 && lora_ksteps(s2.rank) <= LN_BT_MAX_KSTEPS && tt.n == 0 && l + 1 < c.layers;
      if (s2.rank > 0 && tt.n == 0 && !p.ln_bt_fc2 &&
          gemm_plan_init(&p.bt_fc2, M, LORA_PAD, D, e->dh_a, D, s2.lb_bwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, plain))
        return 1;
      GemmEpilogue ep = {EPI_MUL, nullptr, e->u[l], F, nullptr, 0};
      if (gemm_plan_init(&p.bfc2, M, F, D, e->dh_a, D, w.fc2_wt, D, e->du, F, nullptr, 0, e->T, 3 * LORA_PAD, s2.la_bwd,
                         LORA_PAD, s2.rank > 0 ? 1 : 0, lora_ksteps(s2.rank), 0, ep, &tt))
        return 1;
    }
    {
      const GemmTT tt = make_tt(s1, s1.lb_bwd, F, lora_ksteps(s1.rank), nullptr, e->tt_sites & 8);
      p.bt_fc1.M = 0;
      if (s1.rank > 0 && tt.n == 0 &&
          gemm_plan_init(&p.bt_fc1, M, LORA_PAD, F, e->du, F, s1.lb_bwd, F, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, plain))
        return 1;
      if (gemm_plan_init(&p.bfc1, M, D, F, e->du, F, w.fc1_wt, F, e->dxn, D, nullptr, 0, e->T, 3 * LORA_PAD, s1.la_bwd,
                         LORA_PAD, s1.rank > 0 ? 1 : 0, lora_ksteps(s1.rank), 0, plain, &tt))
        return 1;
    }
    {
      const GemmTT tt = make_tt(sp, sp.lb_bwd, D, lora_ksteps(sp.rank), nullptr, e->tt_sites & 16);
      p.bt_proj.M = 0;
      p.ln_bt_proj = e->fuse_ln_bt && sp.rank > 0 && lora_ksteps(sp.rank) <= LN_BT_MAX_KSTEPS && tt.n == 0;  // fused into LN2's backward, which writes dh_b
      if (sp.rank > 0 && tt.n == 0 && !p.ln_bt_proj &&
          gemm_plan_init(&p.bt_proj, M, LORA_PAD, D, e->dh_b, D, sp.lb_bwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr, 0,
                         nullptr, 0, 0, 0, 0, plain))
        return 1;
      // proj backward also produces attention's delta = rowsum(dO o O) per (image, head, query): one 64-column slab of
      // the output is one head, so the pair kernel's slab epilogue gets it for free (replaces a 154 MB pass)
      GemmEpilogue ep = plain;
      if (e->fuse_delta) {
        ep.mode = EPI_ROWDOT;
        ep.res = e->ao[l];
        ep.ld_res = D;
        ep.rowdot = e->delta;
        ep.rowdot_rows = TOKENS;
        ep.rowdot_pad = 208;
      }
      if (gemm_plan_init(&p.bproj, M, D, D, e->dh_b, D, w.proj_wt, D, e->dao, D, nullptr, 0, e->T, 3 * LORA_PAD,
                         sp.la_bwd, LORA_PAD, sp.rank > 0 ? 1 : 0, lora_ksteps(sp.rank), 0, ep, &tt))
        return 1;
    }
    {
      const GemmTT tt = make_tt(sq, sq.lb_bwd, 3 * D, lora_ksteps(sq.rank), nullptr, sq.packed && (e->tt_sites & 32));
      p.bt_qkv.M = 0;
      if (sq.rank > 0 && tt.n == 0 &&
          gemm_plan_init(&p.bt_qkv, M, q_tcols, 3 * D, e->dqkv, 3 * D, sq.lb_bwd, 3 * D, e->T, 3 * LORA_PAD, nullptr,
                         0, nullptr, 0, nullptr, 0, 0, 0, 0, plain))
        return 1;
      if (gemm_plan_init(&p.bqkv, M, D, 3 * D, e->dqkv, 3 * D, w.qkv_wt, 3 * D, e->dxn, D, nullptr, 0, e->T,
                         3 * LORA_PAD, sq.la_bwd, q_tcols, q_nkb_bwd, lora_ksteps(sq.rank), 0, plain, &tt))
        return 1;
    }
  }
  // patch-embed input gradient: dcols = dh0 * Wpe   (written over the im2col buffer's twin)
  if (gemm_plan_init(&ps->bpatch, M, D, D, e->dh_a, D, e->patch_wt, D, e->dxn, D, nullptr, 0, nullptr, 0, nullptr, 0, 0,
                     0, 0, plain))
    return 1;
  if (e->res_f16) {  // which operands / outputs live on the fp16 residual streams
    ps->patch.out_f16 = 1;
    ps->bpatch.a_f16 = 1;
    for (int l = 0; l < c.layers; ++l) {
      LayerPlans& p = ps->layers[l];
      const bool fold = e->lw[l].qkv_c1 != nullptr && e->lw[l].fc1_c1 != nullptr;
      if (fold) p.t_qkv.a_f16 = p.qkv.a_f16 = p.t_fc1.a_f16 = p.fc1.a_f16 = 1;  // A = the raw stream (LayerNorm folded)
      p.proj.out_f16 = p.proj.res_f16 = 1;
      p.fc2.out_f16 = p.fc2.res_f16 = 1;
      p.bt_fc2.a_f16 = p.bfc2.a_f16 = 1;    // A = dh_a
      p.bt_proj.a_f16 = p.bproj.a_f16 = 1;  // A = dh_b
    }
  }
  if (e->zigzag) {
    // the skinny x*A^T GEMMs read what the previous kernel has just written front to back: walk it back to front so the
    // tail that is still in L2 is consumed first (the main GEMM that follows reads front to back again)
    for (auto& p : ps->layers)
      for (GemmPlan* g : {&p.t_qkv, &p.t_proj, &p.t_fc1, &p.t_fc2, &p.bt_fc2, &p.bt_fc1, &p.bt_proj, &p.bt_qkv})
        if (g->M > 0) g->reverse_m = 1;
  }
  e->plans[batch] = owner.release();
  *out = ps;
  return 0;
}

// one profiling category per kernel role (include/vitatk.h lists them in this order)
enum ProfCat { CAT_PATCH = 0, CAT_QKV, CAT_PROJ, CAT_FC1, CAT_FC2, CAT_BFC2, CAT_BFC1, CAT_BPROJ, CAT_BQKV, CAT_BPATCH,
               CAT_T_QKV, CAT_T_PROJ, CAT_T_FC1, CAT_T_FC2, CAT_BT_FC2, CAT_BT_FC1, CAT_BT_PROJ, CAT_BT_QKV,
               CAT_ATTN_FWD, CAT_ATTN_BWD, CAT_LN_FWD, CAT_LN_BWD, CAT_HEAD, CAT_PIXEL, CAT_COUNT = 32 };

static cudaEvent_t prof_event(vitatk_engine* e) {
  cudaEvent_t ev;
  if (!e->prof_pool.empty()) {
    ev = e->prof_pool.back();
    e->prof_pool.pop_back();
    return ev;
  }
  cudaEventCreate(&ev);
  return ev;
}
static double plan_flops(const GemmPlan* p) {
  return 2.0 * p->M * p->N * (static_cast<double>(p->K) + 16.0 * p->lora_ksteps * p->lora_nkb);
}

// launch one kernel; with profiling on, bracket it with events on the launching stream
#define RUNC(cat, flops, expr)                                   \
  do {                                                           \
    cudaEvent_t _a = nullptr, _b = nullptr;                      \
    if (e->prof) {                                               \
      _a = prof_event(e);                                        \
      _b = prof_event(e);                                        \
      cudaEventRecord(_a, s);                                    \
    }                                                            \
    if (expr) return 1;                                          \
    ++e->launches;                                               \
    if (e->prof) {                                               \
      cudaEventRecord(_b, s);                                    \
      e->prof_recs.push_back({(cat), (flops), _a, _b});          \
    }                                                            \
  } while (0)
#define RUN_GEMM(cat, plan) RUNC(cat, plan_flops(plan), gemm_launch((plan), s, e->num_sms))

// forward through the encoder from the im2col'd, normalised input in e->cols
static int encoder_forward(vitatk_engine* e, PlanSet* ps, int batch, cudaStream_t s) {
  // small batches are short launches: overlap each kernel's ramp with its predecessor (FGSM batch 8: 2.44 -> 2.29 ms);
  // at batch 256 the launches run ~100 us each and it changes nothing
  PdlScope pdl(batch <= PDL_MAX_BATCH);
  const vitatk_config& c = e->cfg;
  const int M = batch * TOKENS, D = c.dim;
  RUN_GEMM(CAT_PATCH, &ps->patch);
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->lw[l];
    LayerPlans& p = ps->layers[l];
    const int rq = w.lora[VITATK_SITE_QKV].rank, r1 = w.lora[VITATK_SITE_FC1].rank;
    const bool fold = w.qkv_c1 != nullptr && w.fc1_c1 != nullptr;
    if (fold) {  // only (mean, rstd): the normalisation itself happens in the qkv GEMM's epilogue
      if (!(rq > 0 && p.t_qkv.epi.stats_out)) RUNC(CAT_LN_FWD, 0, layernorm_stats(e->h[l], e->st1[l], M, D, c.ln_eps, s, e->res_f16));
    } else
      RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h[l], w.ln1_g, w.ln1_b, e->xn, e->st1[l], M, D, c.ln_eps, s, e->res_f16));
    if (rq > 0) RUN_GEMM(CAT_T_QKV, &p.t_qkv);
    RUN_GEMM(CAT_QKV, &p.qkv);
    RUNC(CAT_ATTN_FWD, 4.0 * batch * c.heads * TOKENS * TOKENS * 64, attention_fwd_tc05(&ps->attn_fwd[l], s));
    if (w.lora[VITATK_SITE_PROJ].rank > 0 && p.t_proj.M > 0) RUN_GEMM(CAT_T_PROJ, &p.t_proj);
    RUN_GEMM(CAT_PROJ, &p.proj);
    if (fold) {
      if (!(r1 > 0 && p.t_fc1.epi.stats_out))
        RUNC(CAT_LN_FWD, 0, layernorm_stats(e->h_mid[l], e->st2[l], M, D, c.ln_eps, s, e->res_f16));
    } else
      RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h_mid[l], w.ln2_g, w.ln2_b, e->xn, e->st2[l], M, D, c.ln_eps, s, e->res_f16));
    if (r1 > 0) RUN_GEMM(CAT_T_FC1, &p.t_fc1);
    RUN_GEMM(CAT_FC1, &p.fc1);
    if (w.lora[VITATK_SITE_FC2].rank > 0 && p.t_fc2.M > 0) RUN_GEMM(CAT_T_FC2, &p.t_fc2);
    RUN_GEMM(CAT_FC2, &p.fc2);
  }
  return 0;
}

// backward from dh_a = dL/dh[layers] down to e->dxn = dL/d(cols)
static int encoder_backward(vitatk_engine* e, PlanSet* ps, int batch, cudaStream_t s) {
  // small batches are short launches: overlap each kernel's ramp with its predecessor (FGSM batch 8: 2.44 -> 2.29 ms);
  // at batch 256 the launches run ~100 us each and it changes nothing
  PdlScope pdl(batch <= PDL_MAX_BATCH);
  const vitatk_config& c = e->cfg;
  const int M = batch * TOKENS, D = c.dim;
  for (int l = c.layers - 1; l >= 0; --l) {
    const LayerWeights& w = e->lw[l];
    LayerPlans& p = ps->layers[l];
    if (w.lora[VITATK_SITE_FC2].rank > 0 && p.bt_fc2.M > 0) RUN_GEMM(CAT_BT_FC2, &p.bt_fc2);
    RUN_GEMM(CAT_BFC2, &p.bfc2);  // du = (dh W2 + lora) * gelu'(u)   (u[l] holds gelu'(u), written by fc1's epilogue)
    if (w.lora[VITATK_SITE_FC1].rank > 0 && p.bt_fc1.M > 0) RUN_GEMM(CAT_BT_FC1, &p.bt_fc1);
    RUN_GEMM(CAT_BFC1, &p.bfc1);  // dxn = du W1 + lora
    if (p.ln_bt_proj)  // dh_mid, and T = dh_mid * B_proj^T for bproj's LoRA k-block
      RUNC(CAT_LN_BWD, 0, layernorm_bwd_bt(e->dxn, e->h_mid[l], e->st2[l], w.ln2_g, e->dh_a, e->dh_b, M, D, s, e->res_f16, e->res_f16,
                                         w.lora[VITATK_SITE_PROJ].lb_bwd, lora_ksteps(w.lora[VITATK_SITE_PROJ].rank), e->T, 3 * LORA_PAD));
    else
      RUNC(CAT_LN_BWD, 0, layernorm_bwd(e->dxn, e->h_mid[l], e->st2[l], w.ln2_g, e->dh_a, e->dh_b, M, D, s, e->res_f16, e->res_f16));  // dh_mid
    if (w.lora[VITATK_SITE_PROJ].rank > 0 && p.bt_proj.M > 0) RUN_GEMM(CAT_BT_PROJ, &p.bt_proj);
    RUN_GEMM(CAT_BPROJ, &p.bproj);  // dao = dh_mid Wp + lora
    RUNC(CAT_ATTN_BWD, 8.0 * batch * c.heads * TOKENS * TOKENS * 64,
         attention_bwd_fused(&ps->attn_bwd[l], s, !e->fuse_delta));
    if (w.lora[VITATK_SITE_QKV].rank > 0 && p.bt_qkv.M > 0) RUN_GEMM(CAT_BT_QKV, &p.bt_qkv);
    RUN_GEMM(CAT_BQKV, &p.bqkv);  // dxn = dqkv Wqkv + lora
    if (l > 0 && ps->layers[l - 1].ln_bt_fc2) {  // dh wrt h[l], and T = dh * B_fc2^T for layer l - 1's bfc2
      const LoraSite& n2 = e->lw[l - 1].lora[VITATK_SITE_FC2];
      RUNC(CAT_LN_BWD, 0, layernorm_bwd_bt(e->dxn, e->h[l], e->st1[l], w.ln1_g, e->dh_b, e->dh_a, M, D, s, e->res_f16, e->res_f16,
                                         n2.lb_bwd, lora_ksteps(n2.rank), e->T, 3 * LORA_PAD));
    } else {
      RUNC(CAT_LN_BWD, 0, layernorm_bwd(e->dxn, e->h[l], e->st1[l], w.ln1_g, e->dh_b, e->dh_a, M, D, s, e->res_f16, e->res_f16));  // dh wrt h[l]
    }
  }
  RUN_GEMM(CAT_BPATCH, &ps->bpatch);  // dxn <- dL/d(cols)
  return 0;
}

static int check_batch(vitatk_engine* e, int batch) {
  if (!e || !e->finalized) {
    set_error("engine not finalized");
    return 1;
  }
  if (batch < 1 || batch > e->cfg.max_batch) {
    set_error("batch %d outside [1, max_batch=%d]", batch, e->cfg.max_batch);
    return 1;
  }
  return 0;
}

}  // namespace vitatk

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* vitatk_last_error(void) { return last_error(); }
int vitatk_version(void) { return 1; }

int vitatk_create(const vitatk_config* cfg, vitatk_engine** out) {
  if (!cfg || !out) {
    set_error("vitatk_create: null argument");
    return 1;
  }
  if (cfg->image_size != 224 || cfg->patch_size != 16 || cfg->dim != 768 || cfg->heads * 64 != cfg->dim ||
      cfg->mlp_dim % 256 != 0 || cfg->layers < 1 || cfg->num_classes < 1 || cfg->max_batch < 1) {
    set_error("vitatk_create: unsupported geometry (need image 224, patch 16, dim 768 = heads*64, mlp %% 256 == 0)");
    return 1;
  }
  int dev = 0;
  VITATK_CUDA_OK(cudaGetDevice(&dev));
  cudaDeviceProp prop;
  VITATK_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
  if (prop.major != 10) {
    set_error("vitatk requires an sm_100a (B200) device; found sm_%d%d. There is no fallback path.", prop.major,
              prop.minor);
    return 1;
  }
  vitatk_engine* e = new vitatk_engine();
  e->cfg = *cfg;
  e->num_sms = prop.multiProcessorCount;
  {
    const char* g2 = getenv("VITATK_GEMM_2CTA");
    const char* fd = getenv("VITATK_FUSE_DELTA");
    const char* zz = getenv("VITATK_ZIGZAG");
    e->zigzag = !(zz && zz[0] == '0');
    const char* tt = getenv("VITATK_TT");
    e->fuse_tt = !(tt && tt[0] == '0') && !(g2 && g2[0] == '0') && cfg->dim % 256 == 0 && cfg->mlp_dim % 256 == 0;
    const char* rf = getenv("VITATK_RES_F16");
    e->res_f16 = !(rf && rf[0] == '0');
    if (!e->res_f16) e->grad_S = 1.0f;
    const char* tts = getenv("VITATK_TT_SITES");
    if (tts) e->tt_sites = atoi(tts);
    const char* fs = getenv("VITATK_FUSE_STATS");
    e->fuse_stats = !(fs && fs[0] == '0');
    const char* tc = getenv("VITATK_TC_CONST");
    e->tc_const = !(tc && tc[0] == '0') && !(g2 && g2[0] == '0') && cfg->dim % 256 == 0 && cfg->mlp_dim % 256 == 0;
    e->fuse_delta = !(g2 && g2[0] == '0') && !(fd && fd[0] == '0') && cfg->dim % 256 == 0;
    // measured neutral on B200 (the CTA-level hand-over costs the LayerNorm kernel what the two skinny launches cost:
    // layernorm_bwd 14.8 -> 19.0 ms per step against 4.7 ms of bt_proj + bt_fc2), so opt-in like the T-tiles
    const char* lb = getenv("VITATK_LN_BT");
    e->fuse_ln_bt = lb && lb[0] == '1' && cfg->dim == 768;
  }
  e->lw.resize(cfg->layers);
  for (int i = 0; i < 3; ++i) {
    e->nrm.mean[i] = cfg->mean[i];
    e->nrm.inv_std[i] = 1.0f / cfg->std[i];
  }
  *out = e;
  return 0;
}

int vitatk_destroy(vitatk_engine* e) {
  if (!e) return 0;
  for (auto& kv : e->plans) delete kv.second;
  if (e->ws) cudaFree(e->ws);
  if (e->cbuf) cudaFree(e->cbuf);
  if (e->tt_flags) cudaFree(e->tt_flags);
  if (e->labels_rep) cudaFree(e->labels_rep);
  if (e->tr.buf) cudaFree(e->tr.buf);
  for (auto& kv : e->tr.plans) delete kv.second;
  for (auto& r : e->prof_recs) {  // profiling events (bench.py's roofline leg)
    cudaEventDestroy(r.a);
    cudaEventDestroy(r.b);
  }
  for (cudaEvent_t ev : e->prof_pool) cudaEventDestroy(ev);
  delete e;
  return 0;
}

int vitatk_set_tensor(vitatk_engine* e, int id, int layer, const void* p, long long nbytes) {
  if (!e || !p) {
    set_error("vitatk_set_tensor: null argument");
    return 1;
  }
  const vitatk_config& c = e->cfg;
  const long long D = c.dim, F = c.mlp_dim, C = c.num_classes;
  long long want = -1;
  LayerWeights* w = nullptr;
  if (id >= 16) {
    if (layer < 0 || layer >= c.layers) {
      set_error("vitatk_set_tensor: layer %d out of range", layer);
      return 1;
    }
    w = &e->lw[layer];
  }
  const float* pf = static_cast<const float*>(p);
  const bf16* pb = static_cast<const bf16*>(p);
  switch (id) {
    case VITATK_PATCH_W: e->patch_w = pb; want = D * D * 2; break;
    case VITATK_PATCH_WT: e->patch_wt = pb; want = D * D * 2; break;
    case VITATK_EMBED_TABLE: e->embed_table = pf; want = TOKENS * D * 4; break;
    case VITATK_LNF_G: e->lnf_g = pf; want = D * 4; break;
    case VITATK_LNF_B: e->lnf_b = pf; want = D * 4; break;
    case VITATK_HEAD_W: e->head_w = pf; want = C * D * 4; break;
    case VITATK_HEAD_B: e->head_b = pf; want = C * 4; break;
    case VITATK_LN1_G: w->ln1_g = pf; want = D * 4; break;
    case VITATK_LN1_B: w->ln1_b = pf; want = D * 4; break;
    case VITATK_LN2_G: w->ln2_g = pf; want = D * 4; break;
    case VITATK_LN2_B: w->ln2_b = pf; want = D * 4; break;
    case VITATK_QKV_W: w->qkv_w = pb; want = 3 * D * D * 2; break;
    case VITATK_QKV_WT: w->qkv_wt = pb; want = 3 * D * D * 2; break;
    case VITATK_QKV_B: w->qkv_b = pf; want = 3 * D * 4; break;
    case VITATK_PROJ_W: w->proj_w = pb; want = D * D * 2; break;
    case VITATK_PROJ_WT: w->proj_wt = pb; want = D * D * 2; break;
    case VITATK_PROJ_B: w->proj_b = pf; want = D * 4; break;
    case VITATK_FC1_W: w->fc1_w = pb; want = F * D * 2; break;
    case VITATK_FC1_WT: w->fc1_wt = pb; want = F * D * 2; break;
    case VITATK_FC1_B: w->fc1_b = pf; want = F * 4; break;
    case VITATK_FC2_W: w->fc2_w = pb; want = F * D * 2; break;
    case VITATK_FC2_WT: w->fc2_wt = pb; want = F * D * 2; break;
    case VITATK_FC2_B: w->fc2_b = pf; want = D * 4; break;
    case VITATK_QKV_C1: w->qkv_c1 = pf; want = 3 * D * 4; break;
    case VITATK_FC1_C1: w->fc1_c1 = pf; want = F * 4; break;
    default: set_error("vitatk_set_tensor: unknown tensor id %d", id); return 1;
  }
  if (nbytes != want) {
    set_error("vitatk_set_tensor: id %d layer %d expects %lld bytes, got %lld", id, layer, want, nbytes);
    return 1;
  }
  // weights changed -> cached TMA plans (and the packed constant columns) are stale
  for (auto& kv : e->plans) delete kv.second;
  e->plans.clear();
  for (auto& kv : e->tr.plans) delete kv.second;
  e->tr.plans.clear();
  e->const_dirty = true;
  return 0;
}

int vitatk_set_lora(vitatk_engine* e, int layer, int site, int rank, const void* la_fwd, const void* lb_fwd,
                    const void* lb_bwd, const void* la_bwd) {
  if (!e || layer < 0 || layer >= e->cfg.layers || site < 0 || site > VITATK_SITE_QKV_PACKED) {
    set_error("vitatk_set_lora: bad layer/site");
    return 1;
  }
  const bool packed = site == VITATK_SITE_QKV_PACKED;
  if (packed) site = VITATK_SITE_QKV;
  if (rank < 0 || rank > LORA_PAD) {
    set_error("vitatk_set_lora: rank %d unsupported (0..%d)", rank, LORA_PAD);
    return 1;
  }
  if (rank > 0 && (!la_fwd || !lb_fwd || !lb_bwd || !la_bwd)) {
    set_error("vitatk_set_lora: null adapter tensor");
    return 1;
  }
  LoraSite& s = e->lw[layer].lora[site];
  s.rank = rank;
  s.packed = packed && rank > 0;
  s.la_fwd = static_cast<const bf16*>(la_fwd);
  s.lb_fwd = static_cast<const bf16*>(lb_fwd);
  s.lb_bwd = static_cast<const bf16*>(lb_bwd);
  s.la_bwd = static_cast<const bf16*>(la_bwd);
  for (auto& kv : e->plans) delete kv.second;
  e->plans.clear();
  e->const_dirty = true;
  return 0;
}

int vitatk_set_normalization(vitatk_engine* e, const float* mean3, const float* std3) {
  if (!e || !mean3 || !std3) {
    set_error("vitatk_set_normalization: null argument");
    return 1;
  }
  for (int i = 0; i < 3; ++i) {
    if (!(std3[i] > 0.f)) {
      set_error("vitatk_set_normalization: std[%d] must be > 0", i);
      return 1;
    }
    e->cfg.mean[i] = mean3[i];
    e->cfg.std[i] = std3[i];
    e->nrm.mean[i] = mean3[i];
    e->nrm.inv_std[i] = 1.0f / std3[i];
  }
  return 0;
}

int vitatk_profile_begin(vitatk_engine* e) {
  if (!e) {
    set_error("vitatk_profile_begin: null engine");
    return 1;
  }
  e->prof = true;
  return 0;
}

int vitatk_profile_end(vitatk_engine* e, double* ms_by_cat, double* flops_by_cat, long long* launches_by_cat) {
  if (!e || !ms_by_cat || !flops_by_cat || !launches_by_cat) {
    set_error("vitatk_profile_end: null argument");
    return 1;
  }
  e->prof = false;
  for (int i = 0; i < CAT_COUNT; ++i) {
    ms_by_cat[i] = 0;
    flops_by_cat[i] = 0;
    launches_by_cat[i] = 0;
  }
  for (auto& r : e->prof_recs) {
    VITATK_CUDA_OK(cudaEventSynchronize(r.b));
    float ms = 0.f;
    VITATK_CUDA_OK(cudaEventElapsedTime(&ms, r.a, r.b));
    ms_by_cat[r.cat] += ms;
    flops_by_cat[r.cat] += r.flops;
    launches_by_cat[r.cat] += 1;
    e->prof_pool.push_back(r.a);
    e->prof_pool.push_back(r.b);
  }
  e->prof_recs.clear();
  return 0;
}

long long vitatk_workspace_bytes(const vitatk_engine* e) { return e ? e->ws_bytes : 0; }
int vitatk_stream_format(const vitatk_engine* e) { return (e && e->res_f16) ? 1 : 0; }
long long vitatk_launch_count(const vitatk_engine* e) { return e ? e->launches : 0; }

int vitatk_finalize(vitatk_engine* e) {
  if (!e) {
    set_error("vitatk_finalize: null engine");
    return 1;
  }
  if (e->finalized) return 0;
  const vitatk_config& c = e->cfg;
  if (!e->patch_w || !e->patch_wt || !e->embed_table || !e->lnf_g || !e->lnf_b || !e->head_w || !e->head_b) {
    set_error("vitatk_finalize: a global tensor is missing");
    return 1;
  }
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->lw[l];
    if (!w.ln1_g || !w.ln1_b || !w.ln2_g || !w.ln2_b || !w.qkv_w || !w.qkv_wt || !w.qkv_b || !w.proj_w || !w.proj_wt ||
        !w.proj_b || !w.fc1_w || !w.fc1_wt || !w.fc1_b || !w.fc2_w || !w.fc2_wt || !w.fc2_b) {
      set_error("vitatk_finalize: layer %d has a missing tensor", l);
      return 1;
    }
  }
  const long long Mmax = static_cast<long long>(c.max_batch) * TOKENS;
  const long long D = c.dim, F = c.mlp_dim;
  auto al = [](long long b) { return (b + 1023) / 1024 * 1024; };
  const long long sz_d = al(Mmax * D * 2), sz_3d = al(Mmax * 3 * D * 2), sz_f = al(Mmax * F * 2);
  const long long sz_st = al(Mmax * 8), sz_t = al(Mmax * 3 * LORA_PAD * 2);
  const long long sz_img = al(static_cast<long long>(c.max_batch) * 3 * 224 * 224 * 4);
  long long total = 0;
  total += (c.layers + 1) * sz_d;                       // h
  total += c.layers * (sz_d + sz_3d + sz_f + 2 * sz_st);  // h_mid, qkv, u, stats
  const long long sz_lse = al(static_cast<long long>(c.max_batch) * c.heads * 208 * 4);
  total += 2 * sz_d + sz_f + sz_t;                      // cols, xn, g, T
  total += c.layers * (sz_d + sz_lse) + sz_lse;         // ao, lse2 per layer; delta
  total += 4 * sz_d + sz_f + sz_3d;                     // dh_a, dh_b, dxn, dao, du, dqkv
  total += al(static_cast<long long>(c.max_batch) * c.num_classes * 4) + al(c.max_batch * 4) + sz_img;
  VITATK_CUDA_OK(cudaMalloc(&e->ws, total));
  VITATK_CUDA_OK(cudaMemset(e->ws, 0, total));
  VITATK_CUDA_OK(cudaMalloc(&e->labels_rep, static_cast<size_t>(c.max_batch) * sizeof(int64_t)));
  {
    const long long nflags = 2 * ((Mmax + 255) / 256);
    VITATK_CUDA_OK(cudaMalloc(&e->tt_flags, nflags * sizeof(unsigned int)));
    VITATK_CUDA_OK(cudaMemset(e->tt_flags, 0, nflags * sizeof(unsigned int)));
  }
  e->ws_bytes = total;
  char* p = e->ws;
  auto take = [&](long long b) {
    char* r = p;
    p += b;
    return r;
  };
  e->h.resize(c.layers + 1);
  e->h_mid.resize(c.layers);
  e->qkv.resize(c.layers);
  e->u.resize(c.layers);
  e->st1.resize(c.layers);
  e->st2.resize(c.layers);
  e->ao.resize(c.layers);
  e->lse2.resize(c.layers);
  for (int l = 0; l <= c.layers; ++l) e->h[l] = reinterpret_cast<bf16*>(take(sz_d));
  for (int l = 0; l < c.layers; ++l) {
    e->h_mid[l] = reinterpret_cast<bf16*>(take(sz_d));
    e->qkv[l] = reinterpret_cast<bf16*>(take(sz_3d));
    e->u[l] = reinterpret_cast<bf16*>(take(sz_f));
    e->st1[l] = reinterpret_cast<float2*>(take(sz_st));
    e->st2[l] = reinterpret_cast<float2*>(take(sz_st));
    e->ao[l] = reinterpret_cast<bf16*>(take(sz_d));
    e->lse2[l] = reinterpret_cast<float*>(take(sz_lse));
  }
  e->delta = reinterpret_cast<float*>(take(sz_lse));
  e->cols = reinterpret_cast<bf16*>(take(sz_d));
  e->xn = reinterpret_cast<bf16*>(take(sz_d));
  e->g = reinterpret_cast<bf16*>(take(sz_f));
  e->T = reinterpret_cast<bf16*>(take(sz_t));
  e->dh_a = reinterpret_cast<bf16*>(take(sz_d));
  e->dh_b = reinterpret_cast<bf16*>(take(sz_d));
  e->dxn = reinterpret_cast<bf16*>(take(sz_d));
  e->dao = reinterpret_cast<bf16*>(take(sz_d));
  e->du = reinterpret_cast<bf16*>(take(sz_f));
  e->dqkv = reinterpret_cast<bf16*>(take(sz_3d));
  e->logits = reinterpret_cast<float*>(take(al(static_cast<long long>(c.max_batch) * c.num_classes * 4)));
  e->loss = reinterpret_cast<float*>(take(al(c.max_batch * 4)));
  e->scratch_img = reinterpret_cast<float*>(take(sz_img));
  if (e->tr.enabled) {  // LoRA training: per-layer GELU outputs + scratch for the weight gradients
    const long long sz_td = al(Mmax * LORA_PAD * 2);
    const long long sz_part = al(((Mmax + 255) / 256) * F * 16 * 4);
    const long long sz_y = al(static_cast<long long>(c.max_batch) * D * 4), sz_dl = al(static_cast<long long>(c.max_batch) * c.num_classes * 4);
    const long long tot = c.layers * sz_f + sz_td + sz_part + sz_y + sz_dl;
    VITATK_CUDA_OK(cudaMalloc(&e->tr.buf, tot));
    VITATK_CUDA_OK(cudaMemset(e->tr.buf, 0, tot));
    char* q = e->tr.buf;
    e->tr.g_layer.resize(c.layers);
    for (int l = 0; l < c.layers; ++l) {
      e->tr.g_layer[l] = reinterpret_cast<bf16*>(q);
      q += sz_f;
    }
    e->tr.Td = reinterpret_cast<bf16*>(q);
    q += sz_td;
    e->tr.partial = reinterpret_cast<float*>(q);
    q += sz_part;
    e->tr.ycls = reinterpret_cast<float*>(q);
    q += sz_y;
    e->tr.dlog = reinterpret_cast<float*>(q);
    e->ws_bytes += tot;
  }
  e->const_dirty = true;  // the tensor-core constant columns are packed on first use (refresh_const_columns)
  e->finalized = true;
  return 0;
}

int vitatk_forward(vitatk_engine* e, const float* images, int batch, float* logits_out, void* stream) {
  if (check_batch(e, batch)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, batch, &ps)) return 1;
  const vitatk_config& c = e->cfg;
  RUNC(CAT_PIXEL, 0, pgd_init(images, nullptr, e->scratch_img, e->cols, batch, e->nrm, 0.f, 0, 0, 0, s));
  if (encoder_forward(e, ps, batch, s)) return 1;
  RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, nullptr, logits_out, nullptr, nullptr,
                   batch, TOKENS, c.dim, c.num_classes, c.ln_eps, 0.f, s, nullptr, e->res_f16, e->res_f16));
  return 0;
}

int vitatk_input_grad(vitatk_engine* e, const float* images, const int64_t* labels, int batch, float* grad_out,
                      float* logits_out, float* loss_out, void* stream) {
  if (check_batch(e, batch)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, batch, &ps)) return 1;
  const vitatk_config& c = e->cfg;
  RUNC(CAT_PIXEL, 0, pgd_init(images, nullptr, e->scratch_img, e->cols, batch, e->nrm, 0.f, 0, 0, 0, s));
  if (encoder_forward(e, ps, batch, s)) return 1;
  // internal gradient is of the SUM of per-image CE (keeps magnitudes independent of batch / sharding);
  // the 1/B of the reference's mean reduction (whitebox_attacks.py:29) is applied when materialising.
  RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, labels, logits_out ? logits_out : e->logits,
                   loss_out ? loss_out : e->loss, e->dh_a, batch, TOKENS, c.dim, c.num_classes, c.ln_eps, e->grad_S, s,
                   nullptr, e->res_f16, e->res_f16));
  if (encoder_backward(e, ps, batch, s)) return 1;
  RUNC(CAT_PIXEL, 0, grad_to_image(e->dxn, grad_out, batch, e->nrm, 1.0f / (batch * e->grad_S), s));
  return 0;
}

int vitatk_vjp(vitatk_engine* e, const float* images, const float* dlogits, int batch, float* grad_out, float* logits_out,
               void* stream) {
  if (check_batch(e, batch)) return 1;
  if (!images || !dlogits || !grad_out) {
    set_error("vitatk_vjp: images, dlogits and grad must be non-null");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, batch, &ps)) return 1;
  const vitatk_config& c = e->cfg;
  RUNC(CAT_PIXEL, 0, pgd_init(images, nullptr, e->scratch_img, e->cols, batch, e->nrm, 0.f, 0, 0, 0, s));
  if (encoder_forward(e, ps, batch, s)) return 1;
  RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, nullptr, logits_out ? logits_out : e->logits,
                   nullptr, e->dh_a, batch, TOKENS, c.dim, c.num_classes, c.ln_eps, e->grad_S, s, dlogits, e->res_f16,
                   e->res_f16));
  if (encoder_backward(e, ps, batch, s)) return 1;
  RUNC(CAT_PIXEL, 0, grad_to_image(e->dxn, grad_out, batch, e->nrm, 1.0f / e->grad_S, s));
  return 0;
}

int vitatk_attack(vitatk_engine* e, const float* images, const int64_t* labels, int batch, float eps, float alpha,
                  int steps, int start, const float* noise, uint64_t seed, uint64_t image_index0, float* adv,
                  void* stream) {
  if (check_batch(e, batch)) return 1;
  if (steps < 1 || !images || !labels || !adv || images == adv) {
    set_error("vitatk_attack: bad arguments (steps>=1, non-null, adv must not alias images)");
    return 1;
  }
  if (start == VITATK_START_NOISE && !noise) {
    set_error("vitatk_attack: start=NOISE needs noise_dev");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, batch, &ps)) return 1;
  const vitatk_config& c = e->cfg;
  RUNC(CAT_PIXEL, 0, pgd_init(images, start == VITATK_START_NOISE ? noise : nullptr, adv, e->cols, batch, e->nrm, eps,
               start == VITATK_START_RNG ? 1 : 0, seed, image_index0, s));
  for (int it = 0; it < steps; ++it) {
    if (encoder_forward(e, ps, batch, s)) return 1;
    RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, labels, e->logits, e->loss, e->dh_a,
                     batch, TOKENS, c.dim, c.num_classes, c.ln_eps, e->grad_S, s, nullptr, e->res_f16, e->res_f16));
    if (encoder_backward(e, ps, batch, s)) return 1;
    RUNC(CAT_PIXEL, 0, pgd_update(e->dxn, images, adv, e->cols, batch, e->nrm, eps, alpha, s));
  }
  return 0;
}

int vitatk_count_correct(vitatk_engine* e, const float* images, const int64_t* labels, int batch, long long* counts,
                         void* stream) {
  if (check_batch(e, batch)) return 1;
  if (!images || !labels || !counts) {
    set_error("vitatk_count_correct: null argument");
    return 1;
  }
  if (vitatk_forward(e, images, batch, e->logits, stream)) return 1;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  RUNC(CAT_HEAD, 0, count_correct(e->logits, labels, batch, e->cfg.num_classes, counts, s));
  return 0;
}

int vitatk_png_roundtrip(const float* images, int batch, float* out, unsigned char* u8_hwc, void* stream) {
  if (!images || batch < 1 || (!out && !u8_hwc)) {
    set_error("vitatk_png_roundtrip: need images, batch >= 1 and at least one output");
    return 1;
  }
  return png_roundtrip(images, out, u8_hwc, batch, static_cast<cudaStream_t>(stream));
}

// ------------------------------- kernel-level entry points -------------------------------
int vitatk_k_gemm(int M, int N, int K, const void* A, int lda, const void* B, int ldb, void* out, int ldo, void* out2,
                  int ldo2, const void* T, int ldt, const void* LB, int ldlb, int lora_nkb, int lora_ksteps_,
                  int lora_group_cols, int epi_mode, const float* bias, const void* res, int ld_res, const float* table,
                  int table_rows, float* rowdot, int rowdot_rows, int rowdot_pad, const float* row_stats, const float* c1,
                  float* stats_out, float stats_eps, const void* tt_tb, int tt_n, const float* tt_bias,
                  unsigned int* tt_flags, int formats, void* stream) {
  GemmPlan p;
  GemmEpilogue ep = {};
  ep.mode = epi_mode;
  ep.bias = bias;
  ep.res = static_cast<const bf16*>(res);
  ep.ld_res = ld_res;
  ep.table = table;
  ep.table_rows = table_rows;
  ep.rowdot = rowdot;
  ep.rowdot_rows = rowdot_rows;
  ep.rowdot_pad = rowdot_pad;
  ep.row_stats = reinterpret_cast<const float2*>(row_stats);
  ep.c1 = c1;
  ep.stats_out = reinterpret_cast<float2*>(stats_out);
  ep.stats_eps = stats_eps;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  GemmTT tt = {};
  if (tt_n > 0) {  // the GEMM computes T = A * tt_tb^T itself (T-tiles) and writes it to T before using it
    tt.n = tt_n;
    tt.tb = static_cast<const bf16*>(tt_tb);
    tt.ld_tb = K;
    tt.out = static_cast<bf16*>(const_cast<void*>(T));
    tt.ld_out = ldt;
    tt.bias = tt_bias;
    tt.flags = tt_flags;
  }
  if (gemm_plan_init(&p, M, N, K, static_cast<const bf16*>(A), lda, static_cast<const bf16*>(B), ldb,
                     static_cast<bf16*>(out), ldo, static_cast<bf16*>(out2), ldo2, static_cast<const bf16*>(T), ldt,
                     static_cast<const bf16*>(LB), ldlb, lora_nkb, lora_ksteps_, lora_group_cols, ep, &tt))
    return 1;
  p.a_f16 = formats & 1;
  p.out_f16 = (formats >> 1) & 1;
  p.res_f16 = (formats >> 2) & 1;
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return gemm_launch(&p, s, sms);
}
int vitatk_k_attention_fwd_tc05(const void* qkv, void* out, float* lse2, int batch, int tokens, int heads, void* stream) {
  AttnFwdPlan p;
  if (attention_fwd_plan_init(&p, static_cast<const bf16*>(qkv), static_cast<bf16*>(out), lse2, batch, tokens, heads))
    return 1;
  return attention_fwd_tc05(&p, static_cast<cudaStream_t>(stream));
}
int vitatk_k_attention_bwd_fused(const void* qkv, const void* dout, const void* o, const float* lse2, float* delta,
                                 void* dqkv, int batch, int tokens, int heads, void* stream) {
  AttnBwdPlan p;
  if (attention_bwd_plan_init(&p, static_cast<const bf16*>(qkv), static_cast<const bf16*>(dout),
                              static_cast<const bf16*>(o), lse2, delta, static_cast<bf16*>(dqkv), batch, tokens, heads))
    return 1;
  return attention_bwd_fused(&p, static_cast<cudaStream_t>(stream), true);
}
int vitatk_k_gemm_trace(long long* dev_buf) { return gemm_set_trace(dev_buf); }
int vitatk_k_attention_bwd_trace(long long* dev_buf) { return attention_bwd_set_trace(dev_buf); }
int vitatk_k_attention_fwd_trace(long long* dev_buf) { return attention_fwd_set_trace(dev_buf); }
int vitatk_k_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stats, int rows,
                           int cols, float eps, int x_f16, void* stream) {
  return layernorm_fwd(static_cast<const bf16*>(x), gamma, beta, static_cast<bf16*>(y),
                       reinterpret_cast<float2*>(stats), rows, cols, eps, static_cast<cudaStream_t>(stream), x_f16);
}
int vitatk_k_layernorm_bwd_bt(const void* dy, const void* x, const float* stats, const float* gamma, const void* dres, void* dx,
                              int rows, int cols, int x_f16, int g_f16, const void* lb, int ksteps, void* T, int ldt, void* stream) {
  return layernorm_bwd_bt(static_cast<const bf16*>(dy), static_cast<const bf16*>(x), reinterpret_cast<const float2*>(stats), gamma,
                          static_cast<const bf16*>(dres), static_cast<bf16*>(dx), rows, cols, static_cast<cudaStream_t>(stream), x_f16,
                          g_f16, static_cast<const bf16*>(lb), ksteps, static_cast<bf16*>(T), ldt);
}
int vitatk_k_layernorm_bwd(const void* dy, const void* x, const float* stats, const float* gamma, const void* dres,
                           void* dx, int rows, int cols, int x_f16, int g_f16, void* stream) {
  return layernorm_bwd(static_cast<const bf16*>(dy), static_cast<const bf16*>(x),
                       reinterpret_cast<const float2*>(stats), gamma, static_cast<const bf16*>(dres),
                       static_cast<bf16*>(dx), rows, cols, static_cast<cudaStream_t>(stream), x_f16, g_f16);
}
int vitatk_k_layernorm_stats(const void* x, float* stats, int rows, int cols, float eps, int x_f16, void* stream) {
  return layernorm_stats(static_cast<const bf16*>(x), reinterpret_cast<float2*>(stats), rows, cols, eps,
                         static_cast<cudaStream_t>(stream), x_f16);
}
static PixelNorm make_norm(const float* mean3, const float* std3) {
  PixelNorm n;
  for (int i = 0; i < 3; ++i) {
    n.mean[i] = mean3[i];
    n.inv_std[i] = 1.0f / std3[i];
  }
  return n;
}
int vitatk_k_pgd_update(const void* dcols, const float* x0, float* adv, void* cols, int batch, const float* mean3,
                        const float* std3, float eps, float alpha, void* stream) {
  return pgd_update(static_cast<const bf16*>(dcols), x0, adv, static_cast<bf16*>(cols), batch, make_norm(mean3, std3),
                    eps, alpha, static_cast<cudaStream_t>(stream));
}
int vitatk_k_pgd_init(const float* x0, const float* noise, float* adv, void* cols, int batch, const float* mean3,
                      const float* std3, float eps, int use_rng, uint64_t seed, uint64_t image_index0, void* stream) {
  return pgd_init(x0, noise, adv, static_cast<bf16*>(cols), batch, make_norm(mean3, std3), eps, use_rng, seed,
                  image_index0, static_cast<cudaStream_t>(stream));
}

}  // extern "C"

// =================================================================================================
// LoRA training step: train_loras.py:295-324 on the engine.  Forward and input-gradient backward are the
// attack path's tcgen05 kernels; train.cu supplies the dropout-aware adapter kernels, weight gradients, Adam and
// re-packing.  Train mode keeps LayerNorm un-folded (the adapters see dropout(LN(h))) and the bias in the epilogue.
// =================================================================================================
namespace vitatk {

static const int AD_SITE[6] = {VITATK_SITE_QKV, VITATK_SITE_QKV, VITATK_SITE_QKV, VITATK_SITE_PROJ, VITATK_SITE_FC1,
                               VITATK_SITE_FC2};

static int build_train_plans(vitatk_engine* e, int batch, std::vector<vitatk_engine::TrainLayerPlans>** out) {
  auto it = e->tr.plans.find(batch);
  if (it != e->tr.plans.end()) {
    *out = it->second;
    return 0;
  }
  const vitatk_config& c = e->cfg;
  const int M = batch * TOKENS, D = c.dim, F = c.mlp_dim;
  std::unique_ptr<std::vector<vitatk_engine::TrainLayerPlans>> owner(new std::vector<vitatk_engine::TrainLayerPlans>(c.layers));
  GemmEpilogue plain = {};
  plain.mode = EPI_PLAIN;
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->lw[l];
    vitatk_engine::TrainLayerPlans& p = (*owner)[l];
    const LoraSite& sq = w.lora[VITATK_SITE_QKV];
    const LoraSite& sp = w.lora[VITATK_SITE_PROJ];
    const LoraSite& s1 = w.lora[VITATK_SITE_FC1];
    const LoraSite& s2 = w.lora[VITATK_SITE_FC2];
    if (sq.rank > 0 && !sq.packed) {
      set_error("training needs the q|k|v adapters packed into one group (rq + rk + rv <= 64)");
      return 1;
    }
    bf16* gl = e->tr.g_layer[l];
    {  // forward fc1 / fc2 through the per-layer GELU buffer (LayerNorm un-folded: A = xn)
      GemmEpilogue ep = plain;
      ep.mode = EPI_GELU_DUAL;
      ep.bias = w.fc1_b;
      if (gemm_plan_init(&p.fc1_t, M, F, D, e->xn, D, w.fc1_w, D, gl, F, e->u[l], F, e->T, 3 * LORA_PAD, s1.lb_fwd, LORA_PAD,
                         s1.rank > 0 ? 1 : 0, lora_ksteps(s1.rank), 0, ep))
        return 1;
      GemmEpilogue ep2 = {EPI_RESIDUAL, w.fc2_b, e->h_mid[l], D, nullptr, 0};
      if (gemm_plan_init(&p.fc2_t, M, D, F, gl, F, w.fc2_w, F, e->h[l + 1], D, nullptr, 0, e->T, 3 * LORA_PAD, s2.lb_fwd,
                         LORA_PAD, s2.rank > 0 ? 1 : 0, lora_ksteps(s2.rank), 0, ep2))
        return 1;
      p.fc2_t.out_f16 = p.fc2_t.res_f16 = e->res_f16 ? 1 : 0;
    }
    // BT = dY * B (un-scaled B^T rows)
    if (s2.rank > 0 && gemm_plan_init(&p.bt_fc2, M, LORA_PAD, D, e->dh_a, D, s2.lb_bwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr,
                                      0, nullptr, 0, 0, 0, 0, plain))
      return 1;
    if (s1.rank > 0 && gemm_plan_init(&p.bt_fc1, M, LORA_PAD, F, e->du, F, s1.lb_bwd, F, e->T, 3 * LORA_PAD, nullptr, 0, nullptr,
                                      0, nullptr, 0, 0, 0, 0, plain))
      return 1;
    if (sp.rank > 0 && gemm_plan_init(&p.bt_proj, M, LORA_PAD, D, e->dh_b, D, sp.lb_bwd, D, e->T, 3 * LORA_PAD, nullptr, 0, nullptr,
                                      0, nullptr, 0, 0, 0, 0, plain))
      return 1;
    if (sq.rank > 0 && gemm_plan_init(&p.bt_qkv, M, LORA_PAD, 3 * D, e->dqkv, 3 * D, sq.lb_bwd, 3 * D, e->T, 3 * LORA_PAD, nullptr,
                                      0, nullptr, 0, nullptr, 0, 0, 0, 0, plain))
      return 1;
    p.bt_fc2.a_f16 = p.bt_proj.a_f16 = e->res_f16 ? 1 : 0;
    // frozen-weight input gradients; the LoRA share is added by lora_dx (its dropout mask does not apply to W's share)
    {
      GemmEpilogue ep = plain;
      if (s2.rank == 0) ep = GemmEpilogue{EPI_MUL, nullptr, e->u[l], F, nullptr, 0};
      if (gemm_plan_init(&p.bfc2_nl, M, F, D, e->dh_a, D, w.fc2_wt, D, e->du, F, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0, 0, ep))
        return 1;
      p.bfc2_nl.a_f16 = e->res_f16 ? 1 : 0;
    }
    if (gemm_plan_init(&p.bfc1_nl, M, D, F, e->du, F, w.fc1_wt, F, e->dxn, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0, 0, plain))
      return 1;
    {
      GemmEpilogue ep = plain;
      if (sp.rank == 0 && e->fuse_delta) {
        ep.mode = EPI_ROWDOT;
        ep.res = e->ao[l];
        ep.ld_res = D;
        ep.rowdot = e->delta;
        ep.rowdot_rows = TOKENS;
        ep.rowdot_pad = 208;
      }
      if (gemm_plan_init(&p.bproj_nl, M, D, D, e->dh_b, D, w.proj_wt, D, e->dao, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0, 0, ep))
        return 1;
      p.bproj_nl.a_f16 = e->res_f16 ? 1 : 0;
    }
    if (gemm_plan_init(&p.bqkv_nl, M, D, 3 * D, e->dqkv, 3 * D, w.qkv_wt, 3 * D, e->dxn, D, nullptr, 0, nullptr, 0, nullptr, 0, 0, 0,
                       0, plain))
      return 1;
  }
  e->tr.plans[batch] = owner.get();
  *out = owner.release();
  return 0;
}

}  // namespace vitatk

extern "C" {

int vitatk_train_enable(vitatk_engine* e, float dropout_p) {
  if (!e || e->finalized || !(dropout_p >= 0.f && dropout_p < 1.f)) {
    set_error("vitatk_train_enable: call before vitatk_finalize with 0 <= dropout < 1");
    return 1;
  }
  e->tr.enabled = true;
  e->tr.p_drop = dropout_p;
  e->tr.ad.assign(e->cfg.layers, std::vector<vitatk_engine::TrainAdapter>(6));
  e->tc_const = false;  // the adapters change every step: constants stay in the epilogue, no packed copies to refresh
  e->const_dirty = true;
  return 0;
}

int vitatk_train_bind(vitatk_engine* e, float* params_dev, float* grads_dev, long long n, long long off_classifier_w,
                      long long off_classifier_b) {
  if (!e || !e->tr.enabled || !params_dev || !grads_dev || n <= 0) {
    set_error("vitatk_train_bind: engine not in training mode or bad arguments");
    return 1;
  }
  const long long C = e->cfg.num_classes, D = e->cfg.dim;
  if (off_classifier_w < 0 || off_classifier_w + C * D > n || off_classifier_b < 0 || off_classifier_b + C > n) {
    set_error("vitatk_train_bind: classifier offsets outside the parameter buffer");
    return 1;
  }
  e->tr.params = params_dev;
  e->tr.grads = grads_dev;
  e->tr.n = n;
  e->tr.off_cw = off_classifier_w;
  e->tr.off_cb = off_classifier_b;
  // the head reads the trainable classifier copy (peft modules_to_save, train_loras.py:84) straight from the masters
  e->head_w = params_dev + off_classifier_w;
  e->head_b = params_dev + off_classifier_b;
  return 0;
}

int vitatk_train_set_adapter(vitatk_engine* e, int layer, int adapter, int rank, float scale, long long off_a, long long off_b) {
  if (!e || !e->tr.enabled || layer < 0 || layer >= e->cfg.layers || adapter < 0 || adapter > 5 || rank < 0 || rank > 64) {
    set_error("vitatk_train_set_adapter: bad arguments");
    return 1;
  }
  vitatk_engine::TrainAdapter& a = e->tr.ad[layer][adapter];
  a.rank = rank;
  a.scale = scale;
  a.off_a = off_a;
  a.off_b = off_b;
  // column of the adapter inside its site's group: q, k, v are packed one after the other
  int col = 0;
  for (int k = 0; k < 3; ++k) {
    e->tr.ad[layer][k].col0 = col;
    col += e->tr.ad[layer][k].rank;
  }
  for (int k = 3; k < 6; ++k) e->tr.ad[layer][k].col0 = 0;
  return 0;
}

// forward (train mode: dropout on the adapters' inputs) + loss + backward + every weight gradient into grads_dev.
// The gradients are those of the MEAN cross-entropy over this call's batch (train_loras.py:309-311).
int vitatk_train_step(vitatk_engine* e, const float* images, const int64_t* labels, int batch, uint64_t seed, uint64_t step,
                      uint64_t image_index0, float* loss_out, float* logits_out, void* stream) {
  if (check_batch(e, batch)) return 1;
  if (!e->tr.enabled || !e->tr.params || !images || !labels) {
    set_error("vitatk_train_step: training mode not set up (vitatk_train_enable / vitatk_train_bind)");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, batch, &ps)) return 1;
  std::vector<vitatk_engine::TrainLayerPlans>* tp = nullptr;
  if (build_train_plans(e, batch, &tp)) return 1;
  const vitatk_config& c = e->cfg;
  const int M = batch * TOKENS, D = c.dim, F = c.mlp_dim, LDT = 3 * LORA_PAD;
  const float p = e->tr.p_drop;
  const long long row0 = static_cast<long long>(image_index0) * TOKENS;
  const int rf = e->res_f16 ? 1 : 0;
  auto A_of = [&](int l, int k) { return e->tr.params + e->tr.ad[l][k].off_a; };
  auto gA = [&](int l, int k) { return e->tr.grads + e->tr.ad[l][k].off_a; };
  auto gB = [&](int l, int k) { return e->tr.grads + e->tr.ad[l][k].off_b; };
  auto mseed = [&](int l, int k) { return train_mask_seed(seed, step, l, k); };
  VITATK_CUDA_OK(cudaMemsetAsync(e->tr.grads, 0, e->tr.n * sizeof(float), s));
  // ------------------------------------------------ forward ------------------------------------------------
  RUNC(CAT_PIXEL, 0, pgd_init(images, nullptr, e->scratch_img, e->cols, batch, e->nrm, 0.f, 0, 0, 0, s));
  RUN_GEMM(CAT_PATCH, &ps->patch);
  for (int l = 0; l < c.layers; ++l) {
    const LayerWeights& w = e->lw[l];
    LayerPlans& pl = ps->layers[l];
    vitatk_engine::TrainLayerPlans& tl = (*tp)[l];
    const auto& ad = e->tr.ad[l];
    if (w.qkv_c1 != nullptr) {
      set_error("vitatk_train_step: the engine was packed with folded LayerNorms; training needs them un-folded");
      return 1;
    }
    RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h[l], w.ln1_g, w.ln1_b, e->xn, e->st1[l], M, D, c.ln_eps, s, rf));
    for (int k = 0; k < 3; ++k)
      if (ad[k].rank > 0)
        RUNC(CAT_T_QKV, 0, lora_down(e->xn, D, D, A_of(l, k), ad[k].rank, e->T, LDT, ad[k].col0, M, mseed(l, k), p, row0, s));
    RUN_GEMM(CAT_QKV, &pl.qkv);
    RUNC(CAT_ATTN_FWD, 0, attention_fwd_tc05(&ps->attn_fwd[l], s));
    if (ad[3].rank > 0)
      RUNC(CAT_T_PROJ, 0, lora_down(e->ao[l], D, D, A_of(l, 3), ad[3].rank, e->T, LDT, 0, M, mseed(l, 3), p, row0, s));
    RUN_GEMM(CAT_PROJ, &pl.proj);
    RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h_mid[l], w.ln2_g, w.ln2_b, e->xn, e->st2[l], M, D, c.ln_eps, s, rf));
    if (ad[4].rank > 0)
      RUNC(CAT_T_FC1, 0, lora_down(e->xn, D, D, A_of(l, 4), ad[4].rank, e->T, LDT, 0, M, mseed(l, 4), p, row0, s));
    RUN_GEMM(CAT_FC1, &tl.fc1_t);
    if (ad[5].rank > 0)
      RUNC(CAT_T_FC2, 0, lora_down(e->tr.g_layer[l], F, F, A_of(l, 5), ad[5].rank, e->T, LDT, 0, M, mseed(l, 5), p, row0, s));
    RUN_GEMM(CAT_FC2, &tl.fc2_t);
  }
  RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, labels, logits_out ? logits_out : e->logits,
                                 loss_out ? loss_out : e->loss, e->dh_a, batch, TOKENS, D, c.num_classes, c.ln_eps, e->grad_S, s,
                                 nullptr, rf, rf, e->tr.ycls, e->tr.dlog));
  const float inv_b = 1.0f / batch;
  RUNC(CAT_HEAD, 0, head_wgrad(e->tr.ycls, e->tr.dlog, batch, D, c.num_classes, inv_b, e->tr.grads + e->tr.off_cw,
                               e->tr.grads + e->tr.off_cb, s));
  // ------------------------------------------------ backward ------------------------------------------------
  const float gs = inv_b / e->grad_S;  // the gradient stream carries grad_S * d(sum CE)
  // one adapter's weight gradients: dB = s dY^T drop(x) A^T (recomputed), dA = s (dY B)^T drop(x)
  auto adapter_wgrad = [&](int l, int k, const bf16* x, int ldx, int in, const bf16* dY, int ldy, int out, int dy_f16) -> int {
    const auto& a = e->tr.ad[l][k];
    if (lora_down(x, ldx, in, A_of(l, k), a.rank, e->tr.Td, LORA_PAD, 0, M, mseed(l, k), p, row0, s)) return 1;
    if (wgrad(dY, ldy, out, e->tr.Td, LORA_PAD, 0, a.rank, M, e->tr.partial, a.scale * gs, 0, gB(l, k), 0u, 0.f, 0, 0, dy_f16, s))
      return 1;
    if (wgrad(x, ldx, in, e->T, LDT, a.col0, a.rank, M, e->tr.partial, a.scale * gs, 1, gA(l, k), mseed(l, k), p, row0, in, 0, s))
      return 1;
    e->launches += 5;  // lora_down + 2 x (wgrad + reduce)
    return 0;
  };
  for (int l = c.layers - 1; l >= 0; --l) {
    const LayerWeights& w = e->lw[l];
    vitatk_engine::TrainLayerPlans& tl = (*tp)[l];
    const auto& ad = e->tr.ad[l];
    // ---- fc2: x = gelu(u) (saved per layer), dY = dh_a ----
    if (ad[5].rank > 0) {
      RUN_GEMM(CAT_BT_FC2, &tl.bt_fc2);
      if (adapter_wgrad(l, 5, e->tr.g_layer[l], F, F, e->dh_a, D, D, rf)) return 1;
    }
    RUN_GEMM(CAT_BFC2, &tl.bfc2_nl);
    if (ad[5].rank > 0) {
      LoraDxArgs a = {};
      a.n = 1;
      a.ad[0] = {A_of(l, 5), ad[5].scale, ad[5].rank, 0, mseed(l, 5)};
      RUNC(CAT_BFC2, 0, lora_dx(e->du, F, F, e->T, LDT, a, e->u[l], F, M, 1, p, row0, s));
    }
    // ---- fc1: x = LN2(h_mid) (recomputed), dY = du ----
    if (ad[4].rank > 0) {
      RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h_mid[l], w.ln2_g, w.ln2_b, e->xn, e->st2[l], M, D, c.ln_eps, s, rf));
      RUN_GEMM(CAT_BT_FC1, &tl.bt_fc1);
      if (adapter_wgrad(l, 4, e->xn, D, D, e->du, F, F, 0)) return 1;
    }
    RUN_GEMM(CAT_BFC1, &tl.bfc1_nl);
    if (ad[4].rank > 0) {
      LoraDxArgs a = {};
      a.n = 1;
      a.ad[0] = {A_of(l, 4), ad[4].scale, ad[4].rank, 0, mseed(l, 4)};
      RUNC(CAT_BFC1, 0, lora_dx(e->dxn, D, D, e->T, LDT, a, nullptr, 0, M, 1, p, row0, s));
    }
    RUNC(CAT_LN_BWD, 0, layernorm_bwd(e->dxn, e->h_mid[l], e->st2[l], w.ln2_g, e->dh_a, e->dh_b, M, D, s, rf, rf));
    // ---- proj: x = attention output, dY = dh_b ----
    if (ad[3].rank > 0) {
      RUN_GEMM(CAT_BT_PROJ, &tl.bt_proj);
      if (adapter_wgrad(l, 3, e->ao[l], D, D, e->dh_b, D, D, rf)) return 1;
    }
    RUN_GEMM(CAT_BPROJ, &tl.bproj_nl);
    if (ad[3].rank > 0) {
      LoraDxArgs a = {};
      a.n = 1;
      a.ad[0] = {A_of(l, 3), ad[3].scale, ad[3].rank, 0, mseed(l, 3)};
      RUNC(CAT_BPROJ, 0, lora_dx(e->dao, D, D, e->T, LDT, a, nullptr, 0, M, 1, p, row0, s));
    }
    // delta = rowsum(dO o O) needs the complete dO: the GEMM epilogue only has it when proj carries no adapter
    RUNC(CAT_ATTN_BWD, 0, attention_bwd_fused(&ps->attn_bwd[l], s, !(ad[3].rank == 0 && e->fuse_delta)));
    // ---- q, k, v: x = LN1(h) (recomputed), dY = the adapter's 768 columns of dqkv ----
    const bool any_qkv = ad[0].rank > 0 || ad[1].rank > 0 || ad[2].rank > 0;
    if (any_qkv) {
      RUNC(CAT_LN_FWD, 0, layernorm_fwd(e->h[l], w.ln1_g, w.ln1_b, e->xn, e->st1[l], M, D, c.ln_eps, s, rf));
      RUN_GEMM(CAT_BT_QKV, &tl.bt_qkv);
      for (int k = 0; k < 3; ++k)
        if (ad[k].rank > 0 && adapter_wgrad(l, k, e->xn, D, D, e->dqkv + k * D, 3 * D, D, 0)) return 1;
    }
    RUN_GEMM(CAT_BQKV, &tl.bqkv_nl);
    if (any_qkv) {
      LoraDxArgs a = {};
      for (int k = 0; k < 3; ++k)
        if (ad[k].rank > 0) a.ad[a.n++] = {A_of(l, k), ad[k].scale, ad[k].rank, ad[k].col0, mseed(l, k)};
      RUNC(CAT_BQKV, 0, lora_dx(e->dxn, D, D, e->T, LDT, a, nullptr, 0, M, 1, p, row0, s));
    }
    RUNC(CAT_LN_BWD, 0, layernorm_bwd(e->dxn, e->h[l], e->st1[l], w.ln1_g, e->dh_b, e->dh_a, M, D, s, rf, rf));
  }
  return 0;
}

// torch.optim.Adam over the bound parameter buffer (train_loras.py:284) + re-packing of every adapter's 16-bit operands.
// m_dev / v_dev: caller-owned fp32 moment buffers of the same length (zero-initialised); step counts from 1.
int vitatk_train_apply(vitatk_engine* e, float* m_dev, float* v_dev, float lr, float beta1, float beta2, float eps, int step,
                       void* stream) {
  if (!e || !e->tr.enabled || !e->tr.params || !m_dev || !v_dev || step < 1) {
    set_error("vitatk_train_apply: bad arguments");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (adam_step(e->tr.params, e->tr.grads, m_dev, v_dev, e->tr.n, lr, beta1, beta2, eps, step, s)) return 1;
  ++e->launches;
  return vitatk_train_repack(e, stream);
}

unsigned int vitatk_train_mask_seed(uint64_t seed, uint64_t step, int layer, int adapter) {
  return train_mask_seed(seed, step, layer, adapter);
}

// masters -> packed operands (also called once after vitatk_train_bind so that the engine computes with the masters)
int vitatk_train_repack(vitatk_engine* e, void* stream) {
  if (!e || !e->tr.enabled || !e->tr.params) {
    set_error("vitatk_train_repack: training mode not set up");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const vitatk_config& c = e->cfg;
  const int D = c.dim, F = c.mlp_dim;
  const int ins[6] = {D, D, D, D, D, F}, outs[6] = {D, D, D, D, F, D};
  for (int l = 0; l < c.layers; ++l) {
    for (int k = 0; k < 6; ++k) {
      const vitatk_engine::TrainAdapter& a = e->tr.ad[l][k];
      if (a.rank <= 0) continue;
      const int site = AD_SITE[k];
      const LoraSite& ls = e->lw[l].lora[site];
      if (!ls.la_fwd) {
        set_error("vitatk_train_repack: layer %d adapter %d has no packed operand buffers (vitatk_set_lora first)", l, k);
        return 1;
      }
      const bool qkv = site == VITATK_SITE_QKV;
      const int out0 = qkv ? k * D : 0;              // row offset of the group in lb_fwd / column offset in lb_bwd
      const int ld_lbb = qkv ? 3 * D : outs[k];      // lb_bwd [64, out_total]
      // operands that meet an fp16 residual stream: lb_bwd of proj / fc2 (la_fwd never does in train mode: LN un-folded)
      const int fmt = (e->res_f16 && (site == VITATK_SITE_PROJ || site == VITATK_SITE_FC2)) ? 2 : 0;
      if (lora_repack(e->tr.params + a.off_a, e->tr.params + a.off_b, a.rank, ins[k], outs[k], a.scale,
                      const_cast<bf16*>(ls.la_fwd), const_cast<bf16*>(ls.lb_fwd), const_cast<bf16*>(ls.lb_bwd),
                      const_cast<bf16*>(ls.la_bwd), a.col0, a.col0, out0, ld_lbb, LORA_PAD, nullptr, fmt, s))
        return 1;
      ++e->launches;
    }
  }
  return 0;
}

}  // extern "C"

// =================================================================================================
// Adversarial patch / EOT front end
// =================================================================================================
extern "C" {

int vitatk_patch_grad(vitatk_engine* e, const float* images, const int64_t* labels, int batch, int T, const float* tf_dev,
                      const float* fw_dev, const float* patch_dev, int p, int circle, float* grad_dev, float* loss_dev,
                      float* logits_dev, void* stream) {
  const int samples = batch * T;
  if (check_batch(e, samples)) return 1;
  if (!images || !labels || !tf_dev || !fw_dev || !patch_dev || !grad_dev || T < 1 || p < 1 || p > 224) {
    set_error("vitatk_patch_grad: bad arguments (batch * T <= max_batch, 1 <= p <= 224)");
    return 1;
  }
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  PlanSet* ps = nullptr;
  if (build_plans(e, samples, &ps)) return 1;
  const vitatk_config& c = e->cfg;
  RUNC(CAT_PIXEL, 0, patch_apply(images, patch_dev, p, tf_dev, T, samples, circle, e->nrm, e->cols, nullptr, s));
  if (encoder_forward(e, ps, samples, s)) return 1;
  RUNC(CAT_HEAD, 0, repeat_labels(labels, T, samples, e->labels_rep, s));
  RUNC(CAT_HEAD, 0, head_fwd_bwd(e->h[c.layers], e->lnf_g, e->lnf_b, e->head_w, e->head_b, e->labels_rep,
                                 logits_dev ? logits_dev : e->logits, loss_dev ? loss_dev : e->loss, e->dh_a, samples, TOKENS, c.dim,
                                 c.num_classes, c.ln_eps, e->grad_S, s, nullptr, e->res_f16, e->res_f16));
  if (encoder_backward(e, ps, samples, s)) return 1;
  // gradient of the MEAN cross-entropy over this call's samples (the caller rescales when it splits a step into chunks)
  RUNC(CAT_PIXEL, 0, patch_grad(e->dxn, tf_dev, fw_dev, p, samples, circle, e->nrm, 1.0f / (samples * e->grad_S), e->scratch_img,
                                grad_dev, s));
  ++e->launches;
  return 0;
}

int vitatk_patch_apply(const float* images, int batch, int T, const float* tf_dev, const float* patch_dev, int p, int circle,
                       float* out_dev, void* stream) {
  if (!images || !tf_dev || !patch_dev || !out_dev || batch < 1 || T < 1 || p < 1 || p > 224) {
    set_error("vitatk_patch_apply: bad arguments");
    return 1;
  }
  PixelNorm n = {{0.f, 0.f, 0.f}, {1.f, 1.f, 1.f}};
  return patch_apply(images, patch_dev, p, tf_dev, T, batch * T, circle, n, nullptr, out_dev, static_cast<cudaStream_t>(stream));
}

int vitatk_patch_update(float* patch_dev, const float* grad_dev, float* m_dev, float* v_dev, int n, float lr, int maximize,
                        int adam_step, float beta1, float beta2, float eps, void* stream) {
  if (!patch_dev || !grad_dev || n < 1 || (adam_step > 0 && (!m_dev || !v_dev))) {
    set_error("vitatk_patch_update: bad arguments");
    return 1;
  }
  return patch_update(patch_dev, grad_dev, m_dev, v_dev, n, lr, maximize ? 1.f : -1.f, adam_step, beta1, beta2, eps,
                      static_cast<cudaStream_t>(stream));
}

}  // extern "C"

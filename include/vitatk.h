/* vitatk — C ABI of the B200-native white-box attack engine (libvitatk.so).
 *
 * The reference has no FFI of its own: its boundary is the Python call surface of
 * whitebox_attacks.py.  Each entry point below names the reference interface it replaces
 * (file:line in rneddojr/Adapting-Pretrained-Vision-Transformers-with-LoRA-against-Attack-Vectors;
 * "HF:" = transformers/models/vit/modeling_vit.py).  INTEGRATION.md shows the ctypes binding a
 * maintainer would add to the reference scripts.
 *
 * Conventions: plain pointers and sizes only (no torch types).  Every pointer named *_dev is a CUDA
 * device pointer owned by the caller; the engine owns only its workspace.  Every call enqueues on the
 * given cudaStream_t (passed as void*) and returns 0 on success, non-zero on failure with a message in
 * vitatk_last_error().  No hidden host synchronisation in the step loop.  One engine per device; an
 * engine is not thread-safe (the reference loop is single-threaded, whitebox_attacks.py:157-173).
 */
#ifndef VITATK_H_
#define VITATK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct vitatk_engine vitatk_engine;

/* Model geometry = Utils.py:84-90 create_vit_model -> HF ViTConfig defaults (ViT-B/16). */
typedef struct vitatk_config {
  int image_size;   /* 224 */
  int patch_size;   /* 16  */
  int dim;          /* 768 */
  int heads;        /* 12  */
  int layers;       /* 12  */
  int mlp_dim;      /* 3072 */
  int num_classes;  /* e.g. 21 */
  int max_batch;    /* largest batch any later call will pass */
  float ln_eps;     /* 1e-12 (HF:layer_norm_eps) */
  float mean[3];    /* Utils.py:92-93 get_normalization */
  float std[3];
} vitatk_config;

/* Tensor ids for vitatk_set_tensor.  bf16 matrices are row-major [out, in] ("W") or its transpose
 * [in, out] ("WT", used by the input-gradient GEMMs); fp32 for vectors/tables. */
enum vitatk_tensor_id {
  /* global (layer argument ignored) */
  VITATK_PATCH_W = 0,    /* bf16 [768, 768]  conv weight flattened [out, c*256+ky*16+kx]   (HF:151) */
  VITATK_PATCH_WT = 1,   /* bf16 [768, 768]  its transpose */
  VITATK_EMBED_TABLE = 2,/* fp32 [197, 768]  row0 = cls+pos[0]; row t = conv bias + pos[t]  (HF:117-124) */
  VITATK_LNF_G = 3,      /* fp32 [768] final LayerNorm (HF:455) */
  VITATK_LNF_B = 4,
  VITATK_HEAD_W = 5,     /* fp32 [C, 768] classifier (HF:613) */
  VITATK_HEAD_B = 6,     /* fp32 [C] */
  /* per layer */
  VITATK_LN1_G = 16, VITATK_LN1_B = 17,      /* fp32 [768]  layernorm_before (HF:325) */
  VITATK_QKV_W = 18,     /* bf16 [2304, 768]  cat(query, key, value).weight (HF:216-218) */
  VITATK_QKV_WT = 19,    /* bf16 [768, 2304] */
  VITATK_QKV_B = 20,     /* fp32 [2304] */
  VITATK_PROJ_W = 21,    /* bf16 [768, 768]  attention.output.dense (HF:262) */
  VITATK_PROJ_WT = 22,
  VITATK_PROJ_B = 23,
  VITATK_LN2_G = 24, VITATK_LN2_B = 25,      /* layernorm_after (HF:326) */
  VITATK_FC1_W = 26,     /* bf16 [3072, 768]  intermediate.dense (HF:290) */
  VITATK_FC1_WT = 27,    /* bf16 [768, 3072] */
  VITATK_FC1_B = 28,
  VITATK_FC2_W = 29,     /* bf16 [768, 3072]  output.dense (HF:305) */
  VITATK_FC2_WT = 30,    /* bf16 [3072, 768] */
  VITATK_FC2_B = 31,
  /* Optional: LayerNorm folded into the qkv / fc1 GEMMs, LN(h) W^T = rstd (h (gamma o W)^T - mean c1) + c2.  When a layer
   * has BOTH c1 vectors set it runs folded and the caller must have passed: QKV_W / FC1_W = gamma o W (the transposes
   * QKV_WT / FC1_WT stay un-folded: the backward needs W), QKV_B / FC1_B = c2 = W beta + s B (A beta) + b, the QKV / FC1
   * adapter la_fwd = gamma o A, and c1[n] = sum_k (gamma o W)[n,k] + sum_j lb_fwd[n,j] sum_k (gamma o A)[j,k]. */
  VITATK_QKV_C1 = 32,    /* fp32 [2304] */
  VITATK_FC1_C1 = 33     /* fp32 [3072] */
};

/* Adapter sites for vitatk_set_lora (train_loras.py:79-95 target_modules; q,k,v share one fused site). */
enum vitatk_lora_site { VITATK_SITE_QKV = 0, VITATK_SITE_PROJ = 1, VITATK_SITE_FC1 = 2, VITATK_SITE_FC2 = 3,
                        /* the QKV site with q, k and v sharing ONE 64-column group (see vitatk_set_lora) */
                        VITATK_SITE_QKV_PACKED = 4 };

const char* vitatk_last_error(void);
int vitatk_version(void);

/* replaces: create_vit_model(...).to(device) (whitebox_attacks.py:92) — allocates nothing big yet */
int vitatk_create(const vitatk_config* cfg, vitatk_engine** out);
int vitatk_destroy(vitatk_engine* e);

/* replaces: model.load_state_dict(torch.load(path)) (whitebox_attacks.py:94).  The pointer is borrowed:
 * the caller keeps the buffer alive for the life of the engine.  nbytes is checked against the shape. */
int vitatk_set_tensor(vitatk_engine* e, int tensor_id, int layer, const void* dev_ptr, long long nbytes);

/* 16-bit format of the engine's two residual streams: 1 = IEEE fp16 (default), 0 = bf16 (VITATK_RES_F16=0).  The tensor
 * cores need both operands of an MMA in ONE 16-bit format, so the "bf16" tensors that multiply a stream must be supplied
 * in the stream's format: PATCH_WT, PROJ_WT, FC2_WT (they meet the gradient stream in the backward), QKV_W / FC1_W when
 * the LayerNorm-fold tensors QKV_C1 / FC1_C1 are supplied (they then meet the raw forward stream), and of the adapters
 * la_fwd of a folded qkv / fc1 site and lb_bwd of the proj / fc2 sites.  Everything else stays bf16. */
int vitatk_stream_format(const vitatk_engine* e);

/* replaces: PeftModel.from_pretrained / get_peft_model (train_loras.py:83-92,419).  rank == 0 removes
 * the adapter.  G = 3 for the fused QKV site (q|k|v groups), 1 otherwise; in/out are the Linear's dims.
 *   la_fwd  bf16 [64*G, in]    rows 64g..64g+r = A_g, rest zero           (T = x A^T)
 *   lb_fwd  bf16 [out, 64]     cols 0..r = (alpha/r) * B (row n of its group), rest zero
 *   lb_bwd  bf16 [64*G, out]   rows 64g..64g+r = B_g^T on group g's columns, zero elsewhere
 *   la_bwd  bf16 [in, 64*G]    cols 64g..64g+r = (alpha/r) * A_g^T, rest zero
 * site == VITATK_SITE_QKV_PACKED (possible when rq + rk + rv <= 64; rank = that sum): the three adapters share one
 * group -- G = 1 above, with A_q in rows [0, rq), A_k in [rq, rq + rk), A_v after them, B_q's columns [0, rq) non-zero only
 * on the q rows of lb_fwd [3*in, 64] (and so on).  One LoRA k-block instead of three, and x*A^T is then computed inside
 * the consumer GEMM (T-tiles) instead of by a separate launch. */
int vitatk_set_lora(vitatk_engine* e, int layer, int site, int rank, const void* la_fwd_dev, const void* lb_fwd_dev,
                    const void* lb_bwd_dev, const void* la_bwd_dev);

/* replaces: attack.set_normalization_used(mean, std) (whitebox_attacks.py:169) and the in-graph
 * (x - mean) / std of whitebox_attacks.py:26 / patch_attack.py:23-25.  Host pointers to 3 floats each. */
int vitatk_set_normalization(vitatk_engine* e, const float* mean3, const float* std3);

/* allocates the activation workspace for max_batch and validates that every tensor is present */
int vitatk_finalize(vitatk_engine* e);
long long vitatk_workspace_bytes(const vitatk_engine* e);

/* replaces: logits = get_model_output(model((x - mean) / std)) (whitebox_attacks.py:26-28).
 * images_dev fp32 [B,3,224,224] in [0,1]; logits_dev fp32 [B, C]. */
int vitatk_forward(vitatk_engine* e, const float* images_dev, int batch, float* logits_dev, void* stream);

/* replaces: F.cross_entropy(logits, labels); loss.backward(); perturbed.grad (whitebox_attacks.py:29-32).
 * grad_dev fp32 [B,3,224,224] = d(mean CE)/d(images); logits_dev / loss_dev (per-image CE) optional. */
int vitatk_input_grad(vitatk_engine* e, const float* images_dev, const int64_t* labels_dev, int batch,
                      float* grad_dev, float* logits_dev, float* loss_dev, void* stream);

/* replaces: the backward pass ART's PyTorchClassifier / autograd drive through NormalizedModel / SignConstrainedModel
 * (patch_attack.py:16-25,50-57, rp2_attack.py:25-30,37-60): grad_dev fp32 [B,3,224,224] = dlogits^T . d logits / d images
 * for a caller-supplied cotangent dlogits_dev fp32 [B, C] (any loss on the logits).  Runs forward + backward; logits_dev
 * optional. */
int vitatk_vjp(vitatk_engine* e, const float* images_dev, const float* dlogits_dev, int batch, float* grad_dev,
               float* logits_dev, void* stream);

/* replaces: batched_fgsm_attack (whitebox_attacks.py:22-38) when steps == 1, alpha == eps, start == NONE,
 * and torchattacks.PGD.__call__ (whitebox_attacks.py:112-113,168-170) otherwise.
 *   start: 0 = none, 1 = counter-based U(-eps,eps) from (seed, image_index0 + b), 2 = caller noise_dev
 *   adv_dev fp32 [B,3,224,224] receives the adversarial images (must not alias images_dev).
 * Every iteration is: im2col'd normalised input -> ViT forward -> CE -> input gradient -> one fused
 * sign-step / projection / clamp / renormalise kernel. */
#define VITATK_START_NONE 0
#define VITATK_START_RNG 1
#define VITATK_START_NOISE 2
int vitatk_attack(vitatk_engine* e, const float* images_dev, const int64_t* labels_dev, int batch, float eps,
                  float alpha, int steps, int start, const float* noise_dev, uint64_t seed, uint64_t image_index0,
                  float* adv_dev, void* stream);

/* replaces: test_model top-1 counting (train_loras.py:56-76). counts_dev int64[2] += {correct, total}. */
int vitatk_count_correct(vitatk_engine* e, const float* images_dev, const int64_t* labels_dev, int batch,
                         long long* counts_dev, void* stream);

/* replaces: save_images (Utils.py:106-113: clamp, *255, truncate to uint8, PNG) + re-loading the file with ToTensor,
 * i.e. the pixel values train_loras.py:56-76 / eval_compose.py:16-59 really evaluate.  out_dev fp32 [B,3,224,224] =
 * trunc(clamp(x,0,1)*255)/255 (may alias images_dev) and / or u8_hwc_dev uint8 [B,224,224,3] (what PIL receives);
 * either may be NULL.  Needs no engine. */
int vitatk_png_roundtrip(const float* images_dev, int batch, float* out_dev, unsigned char* u8_hwc_dev, void* stream);

/* number of kernels the engine enqueued since creation (bench.py reports the per-step delta) */
long long vitatk_launch_count(const vitatk_engine* e);

/* Per-launch CUDA-event timing for bench.py's roofline leg (never on inside a timed region).  Between
 * begin and end every kernel the engine enqueues is bracketed by events on its stream; end() waits for
 * them and returns, per category, the summed duration (ms), algorithmic FLOPs and launch count.
 * Categories (arrays have 32 entries): 0 patch-embed GEMM, 1 qkv, 2 proj, 3 fc1, 4 fc2, 5 fc2-bwd, 6 fc1-bwd,
 * 7 proj-bwd, 8 qkv-bwd, 9 patch-bwd, 10-13 LoRA x*A^T GEMMs (qkv, proj, fc1, fc2), 14-17 LoRA dY*B GEMMs (fc2, fc1,
 * proj, qkv), 18 attention fwd, 19 attention bwd, 20 LayerNorm fwd, 21 LayerNorm bwd, 22 head/CE/count,
 * 23 pixel kernels (PGD init/update, gradient materialisation). */
int vitatk_profile_begin(vitatk_engine* e);
int vitatk_profile_end(vitatk_engine* e, double* ms_by_cat, double* flops_by_cat, long long* launches_by_cat);

/* ---- LoRA training step (replaces the batch body of train_loras.py:295-324) ----
 * peft semantics of train_loras.py:79-95 in train mode: y = W x + b + (alpha / r) B (A dropout_p(x)); trainable = every
 * adapter's A [r, in] and B [out, r] plus the classifier copy (modules_to_save).  All trainable parameters live in ONE
 * caller-owned fp32 device buffer (masters) with a gradient buffer of the same layout; the engine computes with 16-bit
 * operands re-packed from the masters after every update.  Adapter ids: 0 query, 1 key, 2 value, 3 attention output,
 * 4 intermediate (fc1), 5 output (fc2).  Dropout masks are counter-based (seed, step, layer, adapter, element index), so
 * a step is reproducible and independent of how images are sharded over GPUs (image_index0 = global index of image 0).
 *   vitatk_train_enable        before vitatk_finalize: switches the engine to training mode (per-layer GELU buffers,
 *                              bias / LayerNorm un-folded) with dropout probability p on the adapters' inputs
 *   vitatk_train_bind          masters + gradient buffer (n floats) and the classifier's offsets in them
 *   vitatk_train_set_adapter   where adapter (layer, id) lives in the buffers: A at off_a ([r, in]), B at off_b ([out, r])
 *   vitatk_train_repack        masters -> packed operands (call once after binding; vitatk_train_apply does it itself)
 *   vitatk_train_step          forward + mean cross-entropy + backward: every weight gradient written to the gradient
 *                              buffer (an all-reduce over data-parallel ranks goes between step and apply)
 *   vitatk_train_apply         torch.optim.Adam update of the masters (train_loras.py:284) + re-pack */
int vitatk_train_enable(vitatk_engine* e, float dropout_p);
/* key of the dropout mask of adapter (layer, id) at a step; element (row, k) of the adapter's [rows, in] input is kept iff
 * lowbias32((row * in + k) ^ key) >= p * 2^32 (host function, no GPU needed: lets a test pin the oracle's restatement) */
unsigned int vitatk_train_mask_seed(uint64_t seed, uint64_t step, int layer, int adapter);
int vitatk_train_bind(vitatk_engine* e, float* params_dev, float* grads_dev, long long n, long long off_classifier_w,
                      long long off_classifier_b);
int vitatk_train_set_adapter(vitatk_engine* e, int layer, int adapter, int rank, float scale, long long off_a, long long off_b);
int vitatk_train_repack(vitatk_engine* e, void* stream);
int vitatk_train_step(vitatk_engine* e, const float* images_dev, const int64_t* labels_dev, int batch, uint64_t seed,
                      uint64_t step, uint64_t image_index0, float* loss_dev, float* logits_dev, void* stream);
int vitatk_train_apply(vitatk_engine* e, float* m_dev, float* v_dev, float lr, float beta1, float beta2, float eps, int step,
                       void* stream);

/* ---- adversarial patch / EOT front end (replaces what ART's AdversarialPatchPyTorch does around the
 * model in patch_attack.py:47-75,194-204 and rp2_attack.py:33-72) ----
 * One shared patch [3, p, p] (fp32, [0,1]) is pasted into every image under T random transforms per image (scale,
 * rotation, translation; circular or square mask): sample n = image n / T under transform n.  tf_dev [batch*T, 6] is the
 * inverse affine map output-normalised -> patch-normalised coordinates ((x + 0.5) * 2 / 224 - 1 etc.), fw_dev its inverse.
 *   vitatk_patch_grad    composite -> normalised im2col -> ViT forward -> mean CE over the batch*T samples -> backward ->
 *                        grad_dev [3, p, p] += d loss / d patch (deterministic gather, no atomics).  batch*T <= max_batch;
 *                        a larger EOT step is split into several calls by the caller.  loss_dev [batch*T] / logits_dev optional.
 *   vitatk_patch_apply   the composite itself as fp32 images [batch*T, 3, 224, 224] (attack.apply_patch, patch_attack.py:204)
 *   vitatk_patch_update  one optimiser step on the patch: Adam (adam_step >= 1 = step counter; ART's default, lr 5.0) or
 *                        a sign step (adam_step == 0, ART's "pgd"), ascending the loss when maximize != 0, then clip [0,1] */
int vitatk_patch_grad(vitatk_engine* e, const float* images_dev, const int64_t* labels_dev, int batch, int T,
                      const float* tf_dev, const float* fw_dev, const float* patch_dev, int p, int circle, float* grad_dev,
                      float* loss_dev, float* logits_dev, void* stream);
int vitatk_patch_apply(const float* images_dev, int batch, int T, const float* tf_dev, const float* patch_dev, int p,
                       int circle, float* out_dev, void* stream);
int vitatk_patch_update(float* patch_dev, const float* grad_dev, float* m_dev, float* v_dev, int n, float lr, int maximize,
                        int adam_step, float beta1, float beta2, float eps, void* stream);

/* ---- Swin Transformer (shifted-window attention; SURVEY 8(f)-4a, BASELINE configs[2]) ----
 * Same attack surface as the ViT engine for HF SwinForImageClassification (swin-base-patch4-window7-224 geometry: image
 * 224, patch 4, window 7, head dim 32, four stages of width embed_dim * 2^s): forward, input gradient, FGSM / PGD,
 * top-1 counts.  Every Linear runs on the same tcgen05 GEMM with the same LoRA layouts as vitatk_set_lora (q|k|v packed
 * into one 64-column group: VITATK_SITE_QKV here always means the packed layout); window attention, patch merging, the
 * pooled head and the 4x4 patch-embedding pixel kernels are in csrc/swin.cu.  Tensors are bf16 [out, in] (+ transposed
 * copy) for weights and fp32 for everything else, as for the ViT engine; RELBIAS is the relative-position bias already
 * gathered to [heads, 49, 49] (relative_position_bias_table[relative_position_index]); PATCH_W is the conv weight
 * [embed_dim, 3*4*4] zero-padded to 64 columns (PATCH_WT its [64, embed_dim] transpose). */
typedef struct vitatk_swin vitatk_swin;
typedef struct {
  int image_size, patch_size, embed_dim, window;
  int depths[4], heads[4];
  int num_classes, max_batch;
  float ln_eps;
  float mean[3], std[3];
} vitatk_swin_config;
enum vitatk_swin_tensor_id {
  VITATK_SWIN_PATCH_W = 0, VITATK_SWIN_PATCH_WT = 1, VITATK_SWIN_PATCH_B = 2, VITATK_SWIN_EMB_LN_G = 3, VITATK_SWIN_EMB_LN_B = 4,
  VITATK_SWIN_FINAL_LN_G = 5, VITATK_SWIN_FINAL_LN_B = 6, VITATK_SWIN_HEAD_W = 7, VITATK_SWIN_HEAD_B = 8,
  /* per (stage, block) */
  VITATK_SWIN_LN1_G = 16, VITATK_SWIN_LN1_B, VITATK_SWIN_QKV_W, VITATK_SWIN_QKV_WT, VITATK_SWIN_QKV_B, VITATK_SWIN_RELBIAS,
  VITATK_SWIN_PROJ_W, VITATK_SWIN_PROJ_WT, VITATK_SWIN_PROJ_B, VITATK_SWIN_LN2_G, VITATK_SWIN_LN2_B, VITATK_SWIN_FC1_W,
  VITATK_SWIN_FC1_WT, VITATK_SWIN_FC1_B, VITATK_SWIN_FC2_W, VITATK_SWIN_FC2_WT, VITATK_SWIN_FC2_B,
  /* per stage 0..2: patch merging (LayerNorm over 4C, reduction Linear 4C -> 2C without bias) */
  VITATK_SWIN_MERGE_LN_G = 64, VITATK_SWIN_MERGE_LN_B, VITATK_SWIN_MERGE_W, VITATK_SWIN_MERGE_WT
};
int vitatk_swin_create(const vitatk_swin_config* cfg, vitatk_swin** out);
int vitatk_swin_destroy(vitatk_swin* e);
int vitatk_swin_set_tensor(vitatk_swin* e, int tensor_id, int stage, int block, const void* dev_ptr, long long nbytes);
int vitatk_swin_set_lora(vitatk_swin* e, int stage, int block, int site, int rank, const void* la_fwd_dev, const void* lb_fwd_dev,
                         const void* lb_bwd_dev, const void* la_bwd_dev);
int vitatk_swin_set_normalization(vitatk_swin* e, const float* mean3, const float* std3);
int vitatk_swin_finalize(vitatk_swin* e);
long long vitatk_swin_workspace_bytes(const vitatk_swin* e);
long long vitatk_swin_launch_count(const vitatk_swin* e);
int vitatk_swin_forward(vitatk_swin* e, const float* images_dev, int batch, float* logits_dev, void* stream);
int vitatk_swin_input_grad(vitatk_swin* e, const float* images_dev, const int64_t* labels_dev, int batch, float* grad_dev,
                           float* logits_dev, float* loss_dev, void* stream);
int vitatk_swin_attack(vitatk_swin* e, const float* images_dev, const int64_t* labels_dev, int batch, float eps, float alpha,
                       int steps, int start, const float* noise_dev, uint64_t seed, uint64_t image_index0, float* adv_dev,
                       void* stream);
int vitatk_swin_count_correct(vitatk_swin* e, const float* images_dev, const int64_t* labels_dev, int batch, long long* counts_dev,
                              void* stream);

/* ---- kernel-level entry points (used by tests/ and bench.py's roofline leg) ----
 * vitatk_k_gemm with tt_n in {32, 64}: "T-tile" mode of the pair kernel -- the GEMM computes T = A * tt_tb^T (tt_tb bf16
 * [64, K], + tt_bias[64] if given) itself, writes it to T_dev and uses it as its LoRA k-block in the same launch;
 * tt_flags_dev is a zero-initialised uint32 [2 * ceil(M / 256)] scratch that the launch leaves zeroed.
 * formats: bit 0 = A and B hold IEEE fp16 (else bf16; T / LB are always bf16), bit 1 = the output is written as fp16,
 * bit 2 = the EPI_RESIDUAL input is fp16 (the engine keeps its two residual streams -- and the weights that multiply
 * them -- in fp16).  The LayerNorm entry points take
 * x_f16 (their input x) and g_f16 (dres / dx) the same way. */
int vitatk_k_gemm(int M, int N, int K, const void* A_dev, int lda, const void* B_dev, int ldb, void* out_dev,
                  int ldo, void* out2_dev, int ldo2, const void* T_dev, int ldt, const void* LB_dev, int ldlb,
                  int lora_nkb, int lora_ksteps, int lora_group_cols, int epi_mode, const float* bias_dev,
                  const void* res_dev, int ld_res, const float* table_dev, int table_rows, float* rowdot_dev,
                  int rowdot_rows, int rowdot_pad, const float* row_stats_dev, const float* c1_dev, float* stats_out_dev,
                  float stats_eps, const void* tt_tb_dev, int tt_n, const float* tt_bias_dev, unsigned int* tt_flags_dev,
                  int formats, void* stream);
/* tcgen05 forward (the engine's path); lse2_dev (optional) receives [batch*heads, 208] log2-domain logsumexp */
int vitatk_k_attention_fwd_tc05(const void* qkv_dev, void* out_dev, float* lse2_dev, int batch, int tokens, int heads,
                                void* stream);
/* single-pass tcgen05 backward (the engine's path): needs the forward's output o_dev and lse2_dev; delta_dev is a
 * [batch*heads, 208] fp32 scratch that receives rowsum(dO o O); dqkv_dev [batch*tokens, 3*D] receives dq | dk | dv */
int vitatk_k_attention_bwd_fused(const void* qkv_dev, const void* dout_dev, const void* o_dev, const float* lse2_dev,
                                 float* delta_dev, void* dqkv_dev, int batch, int tokens, int heads, void* stream);
/* timing experiments: device buffer of 4096 int64 receiving CTA 0's (event, step, clock64) timeline when
 * VITATK_ATTN_DBG has bit 32 set (scripts/attn_trace.py, scripts/attn_fwd_trace.py); null switches it off.  The
 * instrumented kernels are only compiled with -DVITATK_DBG_KERNELS (VITATK_DBG_BUILD=1 python -c "import vitatk._lib
 * as l; l.build(force=True)"); in the product build these three calls fail with an explanatory error. */
int vitatk_k_gemm_trace(long long* dev_buf);  /* pair GEMM epilogue timeline (VITATK_GEMM_DBG & 512), scripts/gemm_trace.py */
int vitatk_k_attention_bwd_trace(long long* trace_dev);
int vitatk_k_attention_fwd_trace(long long* trace_dev);
int vitatk_k_layernorm_fwd(const void* x_dev, const float* gamma_dev, const float* beta_dev, void* y_dev,
                           float* stats_dev, int rows, int cols, float eps, int x_f16, void* stream);
/* (mean, rstd) per row only (stats_dev fp32 [rows, 2]) */
int vitatk_k_layernorm_stats(const void* x_dev, float* stats_dev, int rows, int cols, float eps, int x_f16, void* stream);
int vitatk_k_layernorm_bwd(const void* dy_dev, const void* x_dev, const float* stats_dev, const float* gamma_dev,
                           const void* dres_dev, void* dx_dev, int rows, int cols, int x_f16, int g_f16, void* stream);
/* LayerNorm backward that also writes T[rows, 16 * ksteps] = dx * lb^T (bf16, row stride ldt): lb [16 * ksteps, 768] in dx's
 * 16-bit format (fp16 when g_f16) -- the down-projection the next backward GEMM's LoRA k-block reads; cols == 768, ksteps <= 4 */
int vitatk_k_layernorm_bwd_bt(const void* dy_dev, const void* x_dev, const float* stats_dev, const float* gamma_dev,
                              const void* dres_dev, void* dx_dev, int rows, int cols, int x_f16, int g_f16,
                              const void* lb_dev, int ksteps, void* T_dev, int ldt, void* stream);
int vitatk_k_pgd_update(const void* dcols_dev, const float* x0_dev, float* adv_dev, void* cols_dev, int batch,
                        const float* mean3, const float* std3, float eps, float alpha, void* stream);
int vitatk_k_pgd_init(const float* x0_dev, const float* noise_dev, float* adv_dev, void* cols_dev, int batch,
                      const float* mean3, const float* std3, float eps, int use_rng, uint64_t seed,
                      uint64_t image_index0, void* stream);
/* 7x7 shifted-window attention of the Swin path on its own (HF modeling_swin.py:410-459, :556-582, :615-636): qkv bf16
 * [batch*R*R, 3C] token-major, head dim 32 (C == 32 * heads), 0 <= shift < 7.  bias is the fp32 [heads, 49, 49]
 * relative-position table (bias_is_table = 0; a fragment-order copy is made per call) or the fragment-order table
 * vitatk_k_win_bias_table wrote for the same (heads, shift) (bias_is_table = 1) -- what the engine keeps per block.
 * vitatk_k_win_bias_table returns the table's size in bytes (and writes it when both pointers are given), -1 on error. */
long long vitatk_k_win_bias_table(const float* bias_dev, float* table_dev, int heads, int shift, void* stream);
int vitatk_k_win_attn_fwd(const void* qkv_dev, const float* bias_dev, int bias_is_table, void* out_dev, int batch, int R,
                          int C, int heads, int shift, void* stream);
int vitatk_k_win_attn_bwd(const void* qkv_dev, const void* dout_dev, const float* bias_dev, int bias_is_table,
                          void* dqkv_dev, int batch, int R, int C, int heads, int shift, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VITATK_H_ */

/* pgd_c_abi.c -- the attack path through the C ABI alone: no Python, no torch.
 *
 * What a non-Python host (or the ctypes stub of INTEGRATION.md) does, in order:
 *   vitatk_create -> vitatk_set_tensor x (7 + 16 per layer) -> vitatk_finalize -> vitatk_attack -> vitatk_count_correct.
 * Weights are random (bf16 bit patterns made on the host); the program checks the invariants that hold for any weights:
 * ||adv - x||_inf <= eps exactly, adv in [0,1], the attack is deterministic, and attacking a sub-batch gives the same
 * images as attacking the whole batch (images are independent).
 *
 *   gcc -O2 -std=c99 -I include -I /usr/local/cuda/include examples/pgd_c_abi.c -o examples/pgd_c_abi \
 *       -L <package dir> -lvitatk -L /usr/local/cuda/lib64 -lcudart -lm -Wl,-rpath,<package dir>
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <cuda_runtime_api.h>

#include "vitatk.h"

#define CK(x) do { if ((x) != 0) { fprintf(stderr, "%s failed: %s\n", #x, vitatk_last_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static uint64_t rng_state = 0x9e3779b97f4a7c15ull;
static float urand(void) {  /* xorshift64*, uniform in [0,1) */
  rng_state ^= rng_state >> 12; rng_state ^= rng_state << 25; rng_state ^= rng_state >> 27;
  return (float)((rng_state * 0x2545F4914F6CDD1Dull) >> 40) / 16777216.0f;
}
static uint16_t bf16_bits(float f) {  /* round to nearest even */
  uint32_t u; memcpy(&u, &f, 4);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
static uint16_t f16_bits(float f) {  /* IEEE half, round to nearest even (|f| here is far inside the normal range) */
  uint32_t u; memcpy(&u, &f, 4);
  uint32_t sign = (u >> 16) & 0x8000u, e = (u >> 23) & 0xffu, m = u & 0x7fffffu;
  if (e < 113) {  /* subnormal half or zero */
    if (e < 102) return (uint16_t)sign;
    m |= 0x800000u;
    uint32_t shift = 126 - e, half = 1u << (shift - 1), r = m >> shift;
    if ((m & (2 * half - 1)) > half || ((m & (2 * half - 1)) == half && (r & 1u))) ++r;
    return (uint16_t)(sign | r);
  }
  uint32_t h = ((e - 112) << 10) | (m >> 13);
  if ((m & 0x1fffu) > 0x1000u || ((m & 0x1fffu) == 0x1000u && (h & 1u))) ++h;
  return (uint16_t)(sign | h);
}
/* device 16-bit matrix [rows, cols] with N(0, sigma)-ish entries, plus its transpose.  The transposed copy multiplies
 * the engine's gradient residual stream in the backward pass and must be in that stream's format (vitatk_stream_format:
 * IEEE fp16 by default, bf16 otherwise); the forward copy is bf16 (no LayerNorm-fold tensors are supplied here). */
static int upload_w(int rows, int cols, float sigma, void** w_dev, void** wt_dev, int wt_f16) {
  size_t n = (size_t)rows * cols;
  uint16_t* w = (uint16_t*)malloc(n * 2), * wt = (uint16_t*)malloc(n * 2);
  for (int r = 0; r < rows; ++r)
    for (int c = 0; c < cols; ++c) {
      float v = sigma * (urand() + urand() + urand() + urand() - 2.0f) * 1.7320508f;
      w[(size_t)r * cols + c] = bf16_bits(v);
      wt[(size_t)c * rows + r] = wt_f16 ? f16_bits(v) : bf16_bits(v);
    }
  CU(cudaMalloc(w_dev, n * 2)); CU(cudaMemcpy(*w_dev, w, n * 2, cudaMemcpyHostToDevice));
  if (wt_dev) { CU(cudaMalloc(wt_dev, n * 2)); CU(cudaMemcpy(*wt_dev, wt, n * 2, cudaMemcpyHostToDevice)); }
  free(w); free(wt);
  return 0;
}
static int upload_f(int n, float base, float jitter, void** dev) {
  float* h = (float*)malloc((size_t)n * 4);
  for (int i = 0; i < n; ++i) h[i] = base + jitter * (urand() - 0.5f);
  CU(cudaMalloc(dev, (size_t)n * 4)); CU(cudaMemcpy(*dev, h, (size_t)n * 4, cudaMemcpyHostToDevice));
  free(h);
  return 0;
}

int main(void) {
  enum { B = 4, C = 21, D = 768, F = 3072, L = 12, PIX = 3 * 224 * 224 };
  const float eps = 8.0f / 255.0f, alpha = 2.0f / 255.0f;
  vitatk_config cfg = {224, 16, D, 12, L, F, C, B, 1e-12f, {0.485f, 0.456f, 0.406f}, {0.229f, 0.224f, 0.225f}};
  vitatk_engine* e = NULL;
  CK(vitatk_create(&cfg, &e));
  void *w, *wt, *v;
  const int sf16 = vitatk_stream_format(e);  /* 1: the residual streams (and the operands that multiply them) are fp16 */
  if (upload_w(D, D, 0.02f, &w, &wt, sf16)) return 1;  /* patch-embed: W^T meets the gradient stream */
  CK(vitatk_set_tensor(e, VITATK_PATCH_W, 0, w, (long long)D * D * 2));
  CK(vitatk_set_tensor(e, VITATK_PATCH_WT, 0, wt, (long long)D * D * 2));
  if (upload_f(197 * D, 0.f, 0.04f, &v)) return 1;
  CK(vitatk_set_tensor(e, VITATK_EMBED_TABLE, 0, v, 197LL * D * 4));
  if (upload_f(D, 1.f, 0.1f, &v)) return 1;
  CK(vitatk_set_tensor(e, VITATK_LNF_G, 0, v, D * 4));
  if (upload_f(D, 0.f, 0.1f, &v)) return 1;
  CK(vitatk_set_tensor(e, VITATK_LNF_B, 0, v, D * 4));
  if (upload_f(C * D, 0.f, 0.08f, &v)) return 1;
  CK(vitatk_set_tensor(e, VITATK_HEAD_W, 0, v, (long long)C * D * 4));
  if (upload_f(C, 0.f, 0.02f, &v)) return 1;
  CK(vitatk_set_tensor(e, VITATK_HEAD_B, 0, v, C * 4));
  for (int l = 0; l < L; ++l) {
    const int ids_g[2] = {VITATK_LN1_G, VITATK_LN2_G}, ids_b[2] = {VITATK_LN1_B, VITATK_LN2_B};
    for (int k = 0; k < 2; ++k) {
      if (upload_f(D, 1.f, 0.1f, &v)) return 1;
      CK(vitatk_set_tensor(e, ids_g[k], l, v, D * 4));
      if (upload_f(D, 0.f, 0.1f, &v)) return 1;
      CK(vitatk_set_tensor(e, ids_b[k], l, v, D * 4));
    }
    const struct { int w, wt, b, out, in; } lin[4] = {{VITATK_QKV_W, VITATK_QKV_WT, VITATK_QKV_B, 3 * D, D},
                                                      {VITATK_PROJ_W, VITATK_PROJ_WT, VITATK_PROJ_B, D, D},
                                                      {VITATK_FC1_W, VITATK_FC1_WT, VITATK_FC1_B, F, D},
                                                      {VITATK_FC2_W, VITATK_FC2_WT, VITATK_FC2_B, D, F}};
    for (int k = 0; k < 4; ++k) {
      /* proj (k = 1) and fc2 (k = 3): their W^T multiplies the gradient stream dh; qkv / fc1 W^T multiply bf16 tensors */
      if (upload_w(lin[k].out, lin[k].in, 0.02f, &w, &wt, sf16 && (k == 1 || k == 3))) return 1;
      CK(vitatk_set_tensor(e, lin[k].w, l, w, (long long)lin[k].out * lin[k].in * 2));
      CK(vitatk_set_tensor(e, lin[k].wt, l, wt, (long long)lin[k].out * lin[k].in * 2));
      if (upload_f(lin[k].out, 0.f, 0.02f, &v)) return 1;
      CK(vitatk_set_tensor(e, lin[k].b, l, v, (long long)lin[k].out * 4));
    }
  }
  CK(vitatk_finalize(e));
  printf("engine ready, workspace %.1f MB\n", vitatk_workspace_bytes(e) / 1e6);

  float* x = (float*)malloc((size_t)B * PIX * 4);
  for (size_t i = 0; i < (size_t)B * PIX; ++i) x[i] = urand();
  int64_t y[B];
  for (int b = 0; b < B; ++b) y[b] = (int64_t)(urand() * C) % C;
  float *x_dev, *adv_dev, *adv2_dev; int64_t* y_dev; long long* counts_dev;
  CU(cudaMalloc((void**)&x_dev, (size_t)B * PIX * 4)); CU(cudaMalloc((void**)&adv_dev, (size_t)B * PIX * 4));
  CU(cudaMalloc((void**)&adv2_dev, (size_t)B * PIX * 4)); CU(cudaMalloc((void**)&y_dev, sizeof(y)));
  CU(cudaMalloc((void**)&counts_dev, 16)); CU(cudaMemset(counts_dev, 0, 16));
  CU(cudaMemcpy(x_dev, x, (size_t)B * PIX * 4, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(y_dev, y, sizeof(y), cudaMemcpyHostToDevice));
  /* PGD-10, counter-based random start keyed by (seed, global image index) */
  CK(vitatk_attack(e, x_dev, y_dev, B, eps, alpha, 10, VITATK_START_RNG, NULL, 7, 0, adv_dev, NULL));
  CK(vitatk_attack(e, x_dev, y_dev, B, eps, alpha, 10, VITATK_START_RNG, NULL, 7, 0, adv2_dev, NULL));
  CK(vitatk_count_correct(e, adv_dev, y_dev, B, counts_dev, NULL));
  CU(cudaDeviceSynchronize());
  float* adv = (float*)malloc((size_t)B * PIX * 4), * adv2 = (float*)malloc((size_t)B * PIX * 4);
  CU(cudaMemcpy(adv, adv_dev, (size_t)B * PIX * 4, cudaMemcpyDeviceToHost));
  CU(cudaMemcpy(adv2, adv2_dev, (size_t)B * PIX * 4, cudaMemcpyDeviceToHost));
  float linf = 0.f, lo = 1.f, hi = 0.f;
  for (size_t i = 0; i < (size_t)B * PIX; ++i) {
    const float d = fabsf(adv[i] - x[i]);
    if (d > linf) linf = d;
    if (adv[i] < lo) lo = adv[i];
    if (adv[i] > hi) hi = adv[i];
  }
  int ok = linf <= eps && lo >= 0.f && hi <= 1.f && linf > 0.5f * eps && memcmp(adv, adv2, (size_t)B * PIX * 4) == 0;
  /* images are independent: the last two images attacked alone (global indices 2, 3) give the same bytes */
  CK(vitatk_attack(e, x_dev + 2 * (size_t)PIX, y_dev + 2, 2, eps, alpha, 10, VITATK_START_RNG, NULL, 7, 2, adv2_dev, NULL));
  CU(cudaDeviceSynchronize());
  CU(cudaMemcpy(adv2, adv2_dev, 2 * (size_t)PIX * 4, cudaMemcpyDeviceToHost));
  ok = ok && memcmp(adv + 2 * (size_t)PIX, adv2, 2 * (size_t)PIX * 4) == 0;
  long long counts[2];
  CU(cudaMemcpy(counts, counts_dev, 16, cudaMemcpyDeviceToHost));
  printf("linf %.9f (eps %.9f)  range [%g, %g]  robust-correct %lld / %lld  launches %lld  -> %s\n", linf, eps, lo, hi,
         counts[0], counts[1], vitatk_launch_count(e), ok ? "OK" : "FAILED");
  CK(vitatk_destroy(e));
  return ok ? 0 : 2;
}

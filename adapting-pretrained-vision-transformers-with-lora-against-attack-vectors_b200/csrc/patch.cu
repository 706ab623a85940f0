// Adversarial-patch / EOT front end: what ART's AdversarialPatchPyTorch does around the model in
// patch_attack.py:47-75,194-204 (and rp2_attack.py:33-72) -- per sample a random scale / rotation / translation of ONE
// shared patch, a circular or square mask, the composite image*(1-mask) + patch*mask, and on the way back the gradient of
// the loss with respect to the shared patch -- as three HBM-bound kernels around the engine's forward / backward:
//
//   patch_apply_kernel   sample n = (image b, transform t): for every output pixel the inverse affine map gives patch
//                        coordinates (U, V) in [-1, 1]^2; the patch is sampled bilinearly (align_corners = false, zero
//                        padding = torch grid_sample), the mask is evaluated analytically at (U, V), and the composite is
//                        written straight into the NORMALISED im2col rows of the patch-embedding GEMM (bf16) -- the
//                        transformed images are never materialised (optionally also as fp32 NCHW, for apply_patch()).
//   patch_grad_kernel    adjoint, as a deterministic GATHER: one thread per (sample, patch pixel) walks the output pixels
//                        whose sampling footprint touches that patch pixel (the forward-mapped 2x2 box) and accumulates
//                        weight * mask * dL/dx_hat / std -- no atomics, partial[n][c][py][px].
//   patch_reduce_kernel  fixed-order sum over the samples (+=), so the patch gradient is bit-reproducible.
//   patch_update_kernel  Adam (ART's default optimizer, lr 5.0) or a sign step ("pgd"), then clip to [0, 1].
// Transform convention (oracle/patch_oracle.py restates it): output pixel (x, y) -> X = (x + 0.5) * 2 / 224 - 1, likewise Y;
// (U, V) = M (X, Y, 1) with the 2x3 matrix M = R(-phi) (. - t) / s supplied per sample by the host.
#include <stdint.h>

#include "vitatk_internal.h"

namespace vitatk {

static constexpr int P_IMG = 224, P_PATCH = 16, P_GRID = 14, P_TOK = 197, P_DIM = 768;

__device__ __forceinline__ float patch_mask(float U, float V, int circle) {
  if (fabsf(U) > 1.f || fabsf(V) > 1.f) return 0.f;
  if (!circle) return 1.f;
  // ART's soft circle (AdversarialPatchPyTorch._get_circular_patch_mask, sharpness 40): 1 - clip((x^2 + y^2)^40, 0, 1)
  const float a = U * U + V * V;
  const float a2 = a * a, a4 = a2 * a2, a8 = a4 * a4, a16 = a8 * a8, a32 = a16 * a16;
  return 1.f - fminf(a32 * a8, 1.f);  // a^40
}

// bilinear sample of channel plane `pc` [p, p] at normalised (U, V), zero outside (grid_sample, align_corners = false)
__device__ __forceinline__ float patch_sample(const float* __restrict__ pc, int p, float U, float V) {
  const float u = (U + 1.f) * 0.5f * p - 0.5f, v = (V + 1.f) * 0.5f * p - 0.5f;
  const float fu = floorf(u), fv = floorf(v);
  const int x0 = static_cast<int>(fu), y0 = static_cast<int>(fv);
  const float ax = u - fu, ay = v - fv;
  float acc = 0.f;
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = x0 + dx, yy = y0 + dy;
      if (xx >= 0 && xx < p && yy >= 0 && yy < p)
        acc = fmaf((dx ? ax : 1.f - ax) * (dy ? ay : 1.f - ay), __ldg(pc + yy * p + xx), acc);
    }
  return acc;
}

// one thread = 8 consecutive pixels of one row of one sample, all three channels
__global__ void __launch_bounds__(256) patch_apply_kernel(const float* __restrict__ images, const float* __restrict__ patch, int p,
                                                          const float* __restrict__ tf, int T, int samples, int circle,
                                                          PixelNorm nrm, bf16* __restrict__ cols, float* __restrict__ out_img) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int per_sample = P_IMG * (P_IMG / 8);
  if (gid >= samples * per_sample) return;
  const int n = gid / per_sample, r = gid % per_sample;
  const int y = r / (P_IMG / 8), x8 = r % (P_IMG / 8);
  const int b = n / T;
  const float* m = tf + static_cast<size_t>(n) * 6;
  const float m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
  const float Y = (y + 0.5f) * (2.f / P_IMG) - 1.f;
  float val[3][8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int x = x8 * 8 + i;
    const float X = (x + 0.5f) * (2.f / P_IMG) - 1.f;
    const float U = fmaf(m0, X, fmaf(m1, Y, m2)), V = fmaf(m3, X, fmaf(m4, Y, m5));
    const float mk = patch_mask(U, V, circle);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float im = __ldg(images + (static_cast<size_t>(b) * 3 + c) * P_IMG * P_IMG + y * P_IMG + x);
      float v = im;
      if (mk > 0.f) v = fmaf(mk, patch_sample(patch + static_cast<size_t>(c) * p * p, p, U, V) - im, im);
      val[c][i] = v;
    }
  }
  const int rowi = n * P_TOK + 1 + (y >> 4) * P_GRID + (x8 >> 1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (out_img != nullptr) {
      float4* o = reinterpret_cast<float4*>(out_img + (static_cast<size_t>(n) * 3 + c) * P_IMG * P_IMG + y * P_IMG + x8 * 8);
      o[0] = make_float4(val[c][0], val[c][1], val[c][2], val[c][3]);
      o[1] = make_float4(val[c][4], val[c][5], val[c][6], val[c][7]);
    }
    if (cols != nullptr) {
      uint4 q;
      __nv_bfloat162* q2 = reinterpret_cast<__nv_bfloat162*>(&q);
#pragma unroll
      for (int i = 0; i < 4; ++i)
        q2[i] = __floats2bfloat162_rn((val[c][2 * i] - nrm.mean[c]) * nrm.inv_std[c], (val[c][2 * i + 1] - nrm.mean[c]) * nrm.inv_std[c]);
      *reinterpret_cast<uint4*>(cols + static_cast<size_t>(rowi) * P_DIM + c * 256 + (y & 15) * 16 + (x8 & 1) * 8) = q;
    }
  }
  // the CLS slot (row n * 197) of the im2col matrix must be zero: written by the threads of image row 0
  if (cols != nullptr && y == 0) {
    uint4 z = make_uint4(0, 0, 0, 0);
    uint4* cz = reinterpret_cast<uint4*>(cols + static_cast<size_t>(n) * P_TOK * P_DIM);
    for (int i = x8; i < P_DIM / 8; i += P_IMG / 8) cz[i] = z;
  }
}

// one thread = (sample n, patch pixel (py, px)): gather over the output pixels whose bilinear footprint touches it
__global__ void __launch_bounds__(256) patch_grad_kernel(const bf16* __restrict__ dcols, const float* __restrict__ tf,
                                                         const float* __restrict__ fw, int p, int samples, int circle,
                                                         PixelNorm nrm, float scale, float* __restrict__ partial) {
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  const int pp = p * p;
  if (gid >= samples * pp) return;
  const int n = gid / pp, e = gid % pp, py = e / p, px = e % p;
  const float* m = tf + static_cast<size_t>(n) * 6;
  const float m0 = m[0], m1 = m[1], m2 = m[2], m3 = m[3], m4 = m[4], m5 = m[5];
  // forward map (patch-normalised -> output-normalised) of the four corners of this pixel's footprint [px-1, px+1] x [py-1, py+1]
  const float* f = fw + static_cast<size_t>(n) * 6;
  float xmin = 1e30f, xmax = -1e30f, ymin = 1e30f, ymax = -1e30f;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const float u = px + ((k & 1) ? 1.f : -1.f), v = py + ((k & 2) ? 1.f : -1.f);
    const float U = (u + 0.5f) * (2.f / p) - 1.f, V = (v + 0.5f) * (2.f / p) - 1.f;
    const float X = fmaf(f[0], U, fmaf(f[1], V, f[2])), Y = fmaf(f[3], U, fmaf(f[4], V, f[5]));
    xmin = fminf(xmin, X); xmax = fmaxf(xmax, X); ymin = fminf(ymin, Y); ymax = fmaxf(ymax, Y);
  }
  const int x_lo = max(0, static_cast<int>(floorf((xmin + 1.f) * (P_IMG * 0.5f) - 0.5f)) - 1);
  const int x_hi = min(P_IMG - 1, static_cast<int>(ceilf((xmax + 1.f) * (P_IMG * 0.5f) - 0.5f)) + 1);
  const int y_lo = max(0, static_cast<int>(floorf((ymin + 1.f) * (P_IMG * 0.5f) - 0.5f)) - 1);
  const int y_hi = min(P_IMG - 1, static_cast<int>(ceilf((ymax + 1.f) * (P_IMG * 0.5f) - 0.5f)) + 1);
  float acc[3] = {0.f, 0.f, 0.f};
  for (int y = y_lo; y <= y_hi; ++y) {
    const float Y = (y + 0.5f) * (2.f / P_IMG) - 1.f;
    for (int x = x_lo; x <= x_hi; ++x) {
      const float X = (x + 0.5f) * (2.f / P_IMG) - 1.f;
      const float U = fmaf(m0, X, fmaf(m1, Y, m2)), V = fmaf(m3, X, fmaf(m4, Y, m5));
      const float mk = patch_mask(U, V, circle);
      if (mk <= 0.f) continue;
      const float u = (U + 1.f) * 0.5f * p - 0.5f, v = (V + 1.f) * 0.5f * p - 0.5f;
      const float wx = 1.f - fabsf(u - px), wy = 1.f - fabsf(v - py);
      if (wx <= 0.f || wy <= 0.f) continue;
      const float w = wx * wy * mk;
      const size_t off = static_cast<size_t>(n * P_TOK + 1 + (y >> 4) * P_GRID + (x >> 4)) * P_DIM + (y & 15) * 16 + (x & 15);
#pragma unroll
      for (int c = 0; c < 3; ++c) acc[c] = fmaf(w, __bfloat162float(dcols[off + c * 256]), acc[c]);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) partial[(static_cast<size_t>(n) * 3 + c) * pp + e] = acc[c] * nrm.inv_std[c] * scale;
}

__global__ void patch_reduce_kernel(const float* __restrict__ partial, int samples, int elems, float* __restrict__ grad) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems) return;
  float acc = 0.f;
  for (int n = 0; n < samples; ++n) acc += partial[static_cast<size_t>(n) * elems + i];
  grad[i] += acc;
}

// step > 0: Adam on the patch (maximising: the caller passes the ASCENT direction sign via `dir`); step == 0: sign step
__global__ void patch_update_kernel(float* __restrict__ patch, const float* __restrict__ grad, float* __restrict__ m,
                                    float* __restrict__ v, int n, float lr, float dir, int step, float b1, float b2, float eps,
                                    float bc1, float bc2_sqrt) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float g = -dir * grad[i];  // the optimiser minimises: descend on -dir * dL/dpatch
  float pnew;
  if (step > 0) {
    const float mi = b1 * m[i] + (1.f - b1) * g, vi = b2 * v[i] + (1.f - b2) * g * g;
    m[i] = mi;
    v[i] = vi;
    pnew = patch[i] - (lr / bc1) * (mi / (sqrtf(vi) / bc2_sqrt + eps));
  } else {
    pnew = patch[i] - lr * (g > 0.f ? 1.f : (g < 0.f ? -1.f : 0.f));
  }
  patch[i] = fminf(fmaxf(pnew, 0.f), 1.f);
}

int patch_apply(const float* images, const float* patch, int p, const float* tf, int T, int samples, int circle, PixelNorm nrm,
                bf16* cols, float* out_img, cudaStream_t stream) {
  const long long total = static_cast<long long>(samples) * P_IMG * (P_IMG / 8);
  patch_apply_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(images, patch, p, tf, T, samples, circle, nrm,
                                                                                     cols, out_img);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int patch_grad(const bf16* dcols, const float* tf, const float* fw, int p, int samples, int circle, PixelNorm nrm, float scale,
               float* partial, float* grad, cudaStream_t stream) {
  const long long total = static_cast<long long>(samples) * p * p;
  patch_grad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, stream>>>(dcols, tf, fw, p, samples, circle, nrm, scale,
                                                                                    partial);
  patch_reduce_kernel<<<(3 * p * p + 255) / 256, 256, 0, stream>>>(partial, samples, 3 * p * p, grad);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

int patch_update(float* patch, const float* grad, float* m, float* v, int n, float lr, float dir, int step, float b1, float b2,
                 float eps, cudaStream_t stream) {
  const float bc1 = step > 0 ? 1.f - powf(b1, static_cast<float>(step)) : 1.f;
  const float bc2 = step > 0 ? sqrtf(1.f - powf(b2, static_cast<float>(step))) : 1.f;
  patch_update_kernel<<<(n + 255) / 256, 256, 0, stream>>>(patch, grad, m, v, n, lr, dir, step, b1, b2, eps, bc1, bc2);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

__global__ void repeat_labels_kernel(const int64_t* __restrict__ labels, int T, int samples, int64_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < samples) out[i] = labels[i / T];
}
int repeat_labels(const int64_t* labels, int T, int samples, int64_t* out, cudaStream_t stream) {
  repeat_labels_kernel<<<(samples + 255) / 256, 256, 0, stream>>>(labels, T, samples, out);
  VITATK_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace vitatk

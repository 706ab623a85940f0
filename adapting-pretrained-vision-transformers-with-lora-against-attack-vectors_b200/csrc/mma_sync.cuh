// Warp-level tensor-core helpers (ldmatrix + mma.sync m16n8k16, fp32 accumulate) for the small products that sit inside
// HBM-bound kernels: the 49 x 49 x 32 window attention of the Swin path (swin.cu) and the rank-r products of the LoRA
// training step (train.cu).  These shapes are far below one tcgen05 tile (a 64 x 64 x 32 product, or N = 16 outputs), so
// the UMMA path's TMEM / mbarrier hand-offs would cost more than the work; the big GEMMs and the ViT attention use
// tcgen05 (gemm_tc05.cu, attention_tc05.cu).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdint.h>

namespace vitatk {
namespace {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void ldsm_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
// Fragment addressing (lane l; i = l >> 3 selects the 8x8 matrix, r = l & 7 its row).
//   A operand, storage [m][k]:              row m0 + r + (i & 1) * 8,  col k0 + (i >> 1) * 8    (ldsm_x4)
//   B operand, storage [n][k] ("col"):      row n0 + r + (i >> 1) * 8, col k0 + (i & 1) * 8     (ldsm_x4: b0 b1 of n-tile n0, then of n0 + 8)
//   B operand, storage [k][n]:              row k0 + r + (i & 1) * 8,  col n0 + (i >> 1) * 8    (ldsm_x4_t: same register order)
//   A operand, storage [k][m] (transposed): row k0 + r + (i >> 1) * 8, col m0 + (i & 1) * 8     (ldsm_x4_t)
__device__ __forceinline__ int frag_r_lo(int lane) { return (lane & 7) + ((lane >> 3) & 1) * 8; }  // r + (i & 1) * 8
__device__ __forceinline__ int frag_c_hi(int lane) { return (lane >> 4) * 8; }                      // (i >> 1) * 8
__device__ __forceinline__ int frag_r_hi(int lane) { return (lane & 7) + (lane >> 4) * 8; }          // r + (i >> 1) * 8
__device__ __forceinline__ int frag_c_lo(int lane) { return ((lane >> 3) & 1) * 8; }                 // (i & 1) * 8

}  // namespace
}  // namespace vitatk

// f16split.cuh — the f16 hi/lo operand split behind every warp-level tensor-pipe contraction of libvrpx.
//
// mma.sync.m16n8k16 (f16 in, f32 accumulate) issues at the rate of the TF32 m16n8k8 but covers twice the k range
// (tools/mma_bench3.cu on B200: 958 vs 479 MAC/clk/SM).  To keep ~fp32 accuracy every fp32 operand x is carried as two
// halves, hi = f16(x) and lo = f16(x - hi): 22 significant bits, like the TF32 hi/lo pair; a product is three MMAs
// (lo·hi + hi·lo + hi·hi, the 2^-22 lo·lo term is dropped).  Two variants of lo:
//   * scaled   lo = f16((x - hi) * 2^11): lo stays a NORMAL f16 whenever hi is; the lo·hi + hi·lo terms need their own
//              accumulator, folded in with 2^-11 at the end (rollout GEMM-B: operands span many magnitudes);
//   * unscaled lo = f16(x - hi): one accumulator; for |x| < 2^-2 the remainder is a subnormal f16, i.e. it is kept to an
//              ABSOLUTE 2^-25 — right for O(1) operands such as BatchNorm'ed embeddings, projections and probabilities.
// Operands must stay below 65504 in magnitude.
#pragma once
#include <cuda_fp16.h>
#include <stdint.h>

namespace vrpx {

__device__ __forceinline__ void mma_f16_16x8x16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// c += A·B for pre-split fragments with UNSCALED lo halves (single accumulator)
__device__ __forceinline__ void mma3_f16(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                         uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma_f16_16x8x16(c, al, bh0, bh1);
  mma_f16_16x8x16(c, ah, bl0, bl1);
  mma_f16_16x8x16(c, ah, bh0, bh1);
}

constexpr float F16_LO_SCALE = 2048.0f;
// {hi(x0) hi(x1), lo(x0) lo(x1)} with lo scaled by 2^11
__device__ __forceinline__ uint2 split_f16x2(float x0, float x1) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn((x0 - hf.x) * F16_LO_SCALE, (x1 - hf.y) * F16_LO_SCALE);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&h), *reinterpret_cast<const uint32_t*>(&l));
}
// {hi(x0) hi(x1), lo(x0) lo(x1)} with unscaled lo
__device__ __forceinline__ uint2 split_f16x2_u(float x0, float x1) {
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&h), *reinterpret_cast<const uint32_t*>(&l));
}

// 2^x, hardware approximation (relative error ~2^-22); 2^-inf = 0
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

}  // namespace vrpx

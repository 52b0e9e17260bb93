// glimpse_mma.cuh — the two per-instance contractions of the glimpse on the warp-level tensor pipe, one warp per instance,
// embeddings streamed from global memory / L2 (first decode steps: rollout_steps.cu; recompute and gradients of the
// REINFORCE backward: decoder_bwd.cu, where the "query" is dc and the "probabilities" are ds).
#pragma once
#include "tile_gemm.cuh"

namespace vrpx {

// S[node][head] = h[node][:] · q[head][:] on the tensor pipe (m16n8k8 TF32 3-term; M = 16 nodes, N = 8 heads), q = 8 x 128
// floats in shared memory.  Fragment coordinates g = lane >> 2, tq = lane & 3; the K (embedding) axis is permuted so that
// every thread streams whole float4 chunks (see the persistent kernel).  store(n, which, v): node n, head 2 tq + which.
template <class Store>
__device__ __forceinline__ void warp_glimpse_scores(const float* q, const float4* __restrict__ hrow, int N, int lane, Store&& store) {
  const int g = lane >> 2, tq = lane & 3;
  float4 qv[8];
#pragma unroll
  for (int c = 0; c < 8; ++c) qv[c] = *reinterpret_cast<const float4*>(q + g * E + 16 * c + 4 * tq);
  __syncwarp();
  for (int n0 = 0; n0 < N; n0 += 16) {
    const int na = n0 + g, nbb = n0 + g + 8;
    float acc6[2][3][4];
#pragma unroll
    for (int a_ = 0; a_ < 2; ++a_)
#pragma unroll
      for (int b_ = 0; b_ < 3; ++b_)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc6[a_][b_][i] = 0.f;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float4 va[4], vb[4];
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = half * 4 + cc;
        va[cc] = (na < N) ? __ldg(hrow + na * (E / 4) + 4 * c + tq) : make_float4(0.f, 0.f, 0.f, 0.f);
        vb[cc] = (nbb < N) ? __ldg(hrow + nbb * (E / 4) + 4 * c + tq) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
#pragma unroll
      for (int cc = 0; cc < 4; ++cc) {
        const int c = half * 4 + cc;
        const float ae[4] = {va[cc].x, va[cc].y, va[cc].z, va[cc].w};
        const float be[4] = {vb[cc].x, vb[cc].y, vb[cc].z, vb[cc].w};
        const float qe[4] = {qv[c].x, qv[c].y, qv[c].z, qv[c].w};
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          uint32_t ah[4], al[4], bh0, bl0, bh1, bl1;
          split_tf32(ae[2 * u], ah[0], al[0]);
          split_tf32(be[2 * u], ah[1], al[1]);
          split_tf32(ae[2 * u + 1], ah[2], al[2]);
          split_tf32(be[2 * u + 1], ah[3], al[3]);
          split_tf32(qe[2 * u], bh0, bl0);
          split_tf32(qe[2 * u + 1], bh1, bl1);
          mma_tf32_16x8x8(acc6[u][0], al, bh0, bh1);
          mma_tf32_16x8x8(acc6[u][1], ah, bl0, bl1);
          mma_tf32_16x8x8(acc6[u][2], ah, bh0, bh1);
        }
      }
    }
    float acc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
      acc[i] = ((acc6[0][0][i] + acc6[1][0][i]) + (acc6[0][1][i] + acc6[1][1][i])) + (acc6[0][2][i] + acc6[1][2][i]);
    if (na < N) { store(na, 0, acc[0]); store(na, 1, acc[1]); }
    if (nbb < N) { store(nbb, 0, acc[2]); store(nbb, 1, acc[3]); }
  }
}

// C[head][dim] = sum_n P[n][head] h_n[dim] (m16n8k8 TF32 3-term; M = 16 dims, N = 8 heads, K = 8 nodes).  P is read from
// slot[n * 8 + head], the result replaces it as slot[head * 128 + dim] (the warp synchronises in between and at the end).
__device__ __forceinline__ void warp_glimpse_values(float* slot, const float4* __restrict__ hrow, int N, int lane) {
  const int g = lane >> 2, tq = lane & 3;
  float cacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) cacc[j][i] = 0.f;
  for (int n0 = 0; n0 < N; n0 += 8) {
    const int na = n0 + tq, nbb = n0 + tq + 4;
    float4 va[4], vb[4];
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      va[cq] = (na < N) ? __ldg(hrow + na * (E / 4) + 8 * cq + g) : make_float4(0.f, 0.f, 0.f, 0.f);
      vb[cq] = (nbb < N) ? __ldg(hrow + nbb * (E / 4) + 8 * cq + g) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32((na < N) ? slot[na * 8 + g] : 0.f, bh0, bl0);
    split_tf32((nbb < N) ? slot[nbb * 8 + g] : 0.f, bh1, bl1);
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      const float ae[4] = {va[cq].x, va[cq].y, va[cq].z, va[cq].w};
      const float be[4] = {vb[cq].x, vb[cq].y, vb[cq].z, vb[cq].w};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t ah[4], al[4];
        split_tf32(ae[2 * u], ah[0], al[0]);
        split_tf32(ae[2 * u + 1], ah[1], al[1]);
        split_tf32(be[2 * u], ah[2], al[2]);
        split_tf32(be[2 * u + 1], ah[3], al[3]);
        mma_tf32_16x8x8(cacc[2 * cq + u], al, bh0, bh1);
        mma_tf32_16x8x8(cacc[2 * cq + u], ah, bl0, bl1);
        mma_tf32_16x8x8(cacc[2 * cq + u], ah, bh0, bh1);
      }
    }
  }
  __syncwarp();   // every lane is done reading P before c is staged in the slot
#pragma unroll
  for (int cq = 0; cq < 4; ++cq)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int d = 32 * cq + 4 * g + 2 * u, j = 2 * cq + u;
      *reinterpret_cast<float2*>(slot + (2 * tq) * E + d) = make_float2(cacc[j][0], cacc[j][2]);
      *reinterpret_cast<float2*>(slot + (2 * tq + 1) * E + d) = make_float2(cacc[j][1], cacc[j][3]);
    }
  __syncwarp();
}

}  // namespace vrpx

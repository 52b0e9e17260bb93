// score_table.cu — per-episode glimpse score table for the fused rollout (table mode of vrpx_rollout).
//
// The decoder's glimpse scores at step t >= 1 are  s[b,head,n] = q~[b,head] · h[b,n]  with
// q~ = A_l · h[b,last] + Q~g[b]  (agents/graph_decoder.py:75-94 after the folding of vrpx/packing.py).  The part that
// depends on the step, (A_l h[b,l])_head · h[b,n], takes only N different values of l per instance, and A_l has rank 48
// per head (A_l,head = W_k,head^T · W_l,head / sqrt(48), the reference's own query/key projections).  So once per
// episode:
//     QK = h · qk_w^T                      (B·N) x 768 on tcgen05 (gemm_tc): q'[b,l] | k'[b,n], 8 heads x 48
//     S1[b][l][head][n] = q'[b,l,head] · k'[b,n,head]        this kernel, mma.sync on f16 hi/lo halves (~fp32)
// and the decode steps read the 8·N scores of row (b, last) instead of running the 128 -> 1024 tile GEMM and the
// 8·N·128 score pass every step.  Algorithmic cost: the table holds B·N·8·N floats (5.2 GB at TSP-50, B = 65536).
#include "gemm.cuh"
#include "tile_gemm.cuh"

namespace vrpx {

constexpr int DQK = 48;            // decoder head dim (384 / 8)
constexpr int QKW = 2 * NH * DQK;  // 768 columns of QK

// One CTA per instance and half of the heads (grid.y = 2), one warp per head: 4-warp CTAs keep the shared-memory
// footprint at 43 KB (N = 50) so that five CTAs share an SM and cover the global-load latency.  Contractions run on
// mma.sync.m16n8k16 with f16 hi/lo halves (f16split.cuh, unscaled lo: q' and k' are O(1) projections), ~fp32 accuracy.
// The warp splits k'[b, :, head] (N x 48) ONCE into fragment order in shared memory —
//   Ks [key][k16 step c][lane t: hi(dims 16c+2t, +1), hi(dims 16c+2t+8, +9), lo(..), lo(..)]    16 B per (key, c, t)
// (one conflict-free LDS.128 per B fragment, hi / lo pairs adjacent as the two-register HMMA operand) — then streams q'
// rows from global memory as A fragments and writes S1 rows.
template <int NT8>
__global__ void __launch_bounds__(128) k_score_table(const float* __restrict__ qk, float* __restrict__ s1, int N) {
  extern __shared__ __align__(16) uint4 ks_all[];
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int warp = blockIdx.y * 4 + (threadIdx.x >> 5);   // = head
  const int64_t b = blockIdx.x;
  uint4* Ks = ks_all + (threadIdx.x >> 5) * (NT8 * 8 * 12);
  const float* qkb = qk + b * N * QKW;
  for (int idx = lane; idx < NT8 * 8 * 12; idx += 32) {
    const int n = idx / 12, c = (idx >> 2) % 3, tt = idx & 3;
    float2 x0 = make_float2(0.f, 0.f), x1 = x0;
    if (n < N) {
      const float* kp = qkb + (size_t)n * QKW + NH * DQK + warp * DQK + 16 * c + 2 * tt;
      x0 = __ldg(reinterpret_cast<const float2*>(kp));
      x1 = __ldg(reinterpret_cast<const float2*>(kp + 8));
    }
    const uint2 p0 = split_f16x2_u(x0.x, x0.y), p1 = split_f16x2_u(x1.x, x1.y);
    Ks[idx] = make_uint4(p0.x, p1.x, p0.y, p1.y);
  }
  __syncwarp();
  const bool vec_ok = (N & 1) == 0;
  // A fragments of the three k16 steps of a 16-row tile of q': a0 (row la, dims 16c+2t, +1), a1 (row lb, same), a2 (row la,
  // dims 16c+2t+8, +9), a3 (row lb, same).  The rows of the NEXT tile are requested before the current one is multiplied.
  auto load_q = [&](int l0, float2 (&qa)[3][2], float2 (&qb)[3][2]) {
    const int la = l0 + g, lb = l0 + g + 8;
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        qa[c][u] = (la < N) ? __ldg(reinterpret_cast<const float2*>(qkb + (size_t)la * QKW + warp * DQK + 16 * c + 2 * t + 8 * u))
                            : make_float2(0.f, 0.f);
        qb[c][u] = (lb < N) ? __ldg(reinterpret_cast<const float2*>(qkb + (size_t)lb * QKW + warp * DQK + 16 * c + 2 * t + 8 * u))
                            : make_float2(0.f, 0.f);
      }
  };
  float2 qna[3][2], qnb[3][2];
  load_q(0, qna, qnb);
  for (int l0 = 0; l0 < N; l0 += 16) {
    const int la = l0 + g, lb = l0 + g + 8;
    float acc[NT8][4];
#pragma unroll
    for (int j = 0; j < NT8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
    float2 qa[3][2], qb[3][2];
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
      for (int u = 0; u < 2; ++u) { qa[c][u] = qna[c][u]; qb[c][u] = qnb[c][u]; }
    load_q(l0 + 16, qna, qnb);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const uint2 s0 = split_f16x2_u(qa[c][0].x, qa[c][0].y), s1_ = split_f16x2_u(qb[c][0].x, qb[c][0].y);
      const uint2 s2 = split_f16x2_u(qa[c][1].x, qa[c][1].y), s3 = split_f16x2_u(qb[c][1].x, qb[c][1].y);
      const uint32_t ah[4] = {s0.x, s1_.x, s2.x, s3.x}, al[4] = {s0.y, s1_.y, s2.y, s3.y};
#pragma unroll
      for (int j = 0; j < NT8; ++j) {
        const uint4 kf = Ks[((8 * j + g) * 3 + c) * 4 + t];   // B: (k = dims 16c+2t.., n = key 8j+g)
        mma3_f16(acc[j], ah, al, kf.x, kf.y, kf.z, kf.w);
      }
    }
    // C fragment: acc[j][0..1] = (l = la, n = 8j + 2t, +1), acc[j][2..3] = (l = lb, ...)
    float* ra = s1 + (((size_t)b * N + la) * NH + warp) * N;
    float* rb = s1 + (((size_t)b * N + lb) * NH + warp) * N;
#pragma unroll
    for (int j = 0; j < NT8; ++j) {
      const int n = 8 * j + 2 * t;
      if (vec_ok) {
        if (n < N) {
          if (la < N) *reinterpret_cast<float2*>(ra + n) = make_float2(acc[j][0], acc[j][1]);
          if (lb < N) *reinterpret_cast<float2*>(rb + n) = make_float2(acc[j][2], acc[j][3]);
        }
      } else {
        if (la < N) {
          if (n < N) ra[n] = acc[j][0];
          if (n + 1 < N) ra[n + 1] = acc[j][1];
        }
        if (lb < N) {
          if (n < N) rb[n] = acc[j][2];
          if (n + 1 < N) rb[n + 1] = acc[j][3];
        }
      }
    }
  }
}

template <int NT8>
static int launch_score_table(const float* qk, float* s1, int64_t nb, int N, cudaStream_t stream) {
  const int smem = (NH / 2) * NT8 * 8 * 12 * (int)sizeof(uint4);
  VRPX_CUDA(cudaFuncSetAttribute(k_score_table<NT8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_score_table<NT8><<<dim3((unsigned)nb, 2), 128, smem, stream>>>(qk, s1, N);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int64_t score_table_slice(int64_t B) { return B < 16384 ? B : 16384; }

// S1[B][N][8][N] from h [B][N][128] and qk_w [768][128]; qk_buf holds score_table_slice(B)·N·768 floats.
int build_score_table(const float* h, const float* qk_w, int64_t B, int N, float* qk_buf, float* s1,
                      cudaStream_t stream) {
  const int64_t slice = score_table_slice(B);
  const int nt = (N + 7) / 8;
  for (int64_t b0 = 0; b0 < B; b0 += slice) {
    const int64_t nb = (B - b0 < slice) ? (B - b0) : slice;
    GemmArgs ga{h + b0 * N * E, nb * N, E, qk_w, QKW, nullptr, 0, nullptr, nullptr, nullptr, qk_buf};
    int rc = gemm_tc(ga, stream);
    if (rc) return rc;
    float* dst = s1 + (size_t)b0 * N * NH * N;
    if (nt <= 3) rc = launch_score_table<3>(qk_buf, dst, nb, N, stream);
    else if (nt <= 7) rc = launch_score_table<7>(qk_buf, dst, nb, N, stream);
    else if (nt <= 13) rc = launch_score_table<13>(qk_buf, dst, nb, N, stream);
    else rc = launch_score_table<16>(qk_buf, dst, nb, N, stream);
    if (rc) return rc;
  }
  return VRPX_OK;
}

}  // namespace vrpx

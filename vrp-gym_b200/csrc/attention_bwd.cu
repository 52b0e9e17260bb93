// attention_bwd.cu — backward of the encoder's per-instance self-attention on the warp-level tensor pipe
// (the autograd of nn.MultiheadAttention's core, graph_encoder.py:74-104; train path of graph_tsp_agent.py:179-186).
//
// For one head with Q, K, V, O, dO of N x 16 and s = 1/4:
//   P = softmax(s Q K^T);  D_i = dO_i · O_i;  dS_ij = P_ij (dO_i · V_j − D_i)
//   dQ_i = s Σ_j dS_ij K_j;   dK_j = s Σ_i dS_ij Q_i;   dV_j = Σ_i P_ij dO_i
// Every product runs on mma.sync.m16n8k16 with f16 hi / lo halves (f16split.cuh, three MMAs per product, ~fp32).  The
// reductions of dQ run over keys and those of dK, dV over queries, so the score matrix is needed in both orientations:
//   pass A (warp = 16 queries):  S = Q K^T, row softmax (its log-sum-exp L_i is kept), dP = dO V^T, dS, dQ = dS K
//   pass B (warp = 16 keys):     S^T = K Q^T, P^T = exp2(S^T − L_i), dP^T = V dO^T, dS^T, dV = P^T dO, dK = dS^T Q
// (a C fragment of one orientation is not an operand fragment of the other, and a transpose through shared memory of the
// split halves costs more than the two extra products).  One CTA = one instance and TWO heads, eight warps: warp w takes
// head w >> 2 and the 16-row tiles (w & 3), (w & 3) + 4.  All operands are split ONCE while they are staged in shared
// memory, in fragment order, in the two layouts of k_enc_attention_f16 (encoder.cu):
//   row layout   X[row][slot t] = {hi(dims 2t, 2t+1), hi(dims 2t+8, 2t+9), lo(..), lo(..)}   Q·c, K, V, dO   (A operand of a
//                row tile, or B operand with n = row)
//   dim layout   XT[dim][k16 step jj][slot t] = {hi(rows 16jj+2t, +1), hi(rows 16jj+2t+8, +9), lo(..), lo(..)}   Q·c, K, dO
//                (B operand with k = rows)
// with c = log2(e) / 4: scores come out in log2 units (one ex2 per element), dK is rescaled by ln 2.  Rows beyond N are
// staged as zeros.  NK8 = ceil(N / 8) is a template parameter (no run-time guards in the unrolled loops).
#include "common.cuh"
#include "f16split.cuh"

namespace vrpx {
namespace ab {

constexpr float QC = 0.25f * 1.4426950408889634f;   // log2(e) / sqrt(16)
constexpr float LN2 = 0.6931471805599453f;
// dO is a gradient: unscaled, the f16 lo half of a 1e-3-sized value is a subnormal with an absolute 2^-25 floor (5e-5
// relative).  Like every f16-split operand of the library it is scaled by 2^8 before the split (|dO| < 256 keeps f16
// finite; gradients down to 1e-3 keep full precision, 1e-5 about 15 bits); everything downstream is linear in dO, the
// three outputs are scaled back.
constexpr float GS = 256.0f, GSI = 1.0f / 256.0f;

template <int NK8>
struct Layout {
  static constexpr int NJJ = (NK8 + 1) / 2;                   // 16-row tiles
  static constexpr int NP = 16 * NJJ;                         // padded rows
  static constexpr int VDS = 4 * (NJJ + (NJJ & 1)) + 4;       // chunks per dim row of the dim layout
  static constexpr int ROW_ARR = NP * 4, DIM_ARR = 16 * VDS;  // uint4 per array
  static constexpr int HEAD_U4 = 4 * ROW_ARR + 3 * DIM_ARR;   // Qr, Kr, Vr, dOr | Qd, Kd, dOd
  static constexpr int HEAD_BYTES = HEAD_U4 * 16 + 2 * NP * 4;   // + L[NP], D[NP]
  static constexpr int SMEM = 2 * HEAD_BYTES;
};

__device__ __forceinline__ void frag_a(const uint4* rows, int ra, int rb, int t, uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  const uint4 fa = rows[ra * 4 + t], fb = rows[rb * 4 + t];
  hi[0] = fa.x; hi[2] = fa.y; lo[0] = fa.z; lo[2] = fa.w;
  hi[1] = fb.x; hi[3] = fb.y; lo[1] = fb.z; lo[3] = fb.w;
}
// A fragment of one k16 step from the C fragments of two neighbouring 8-column tiles
__device__ __forceinline__ void frag_from_c(const float (&c0)[4], const float (&c1)[4], uint32_t (&hi)[4], uint32_t (&lo)[4]) {
  const uint2 p0 = split_f16x2_u(c0[0], c0[1]), p1 = split_f16x2_u(c0[2], c0[3]);
  const uint2 p2 = split_f16x2_u(c1[0], c1[1]), p3 = split_f16x2_u(c1[2], c1[3]);
  hi[0] = p0.x; hi[1] = p1.x; hi[2] = p2.x; hi[3] = p3.x;
  lo[0] = p0.y; lo[1] = p1.y; lo[2] = p2.y; lo[3] = p3.y;
}

template <int NK8>
__global__ void __launch_bounds__(256, (NK8 <= 8) ? 3 : 1) k_enc_attention_bwd_mma(const float* __restrict__ qkv, const float* __restrict__ att,
                                                                const float* __restrict__ datt, float* __restrict__ dqkv, int N) {
  using LT = Layout<NK8>;
  constexpr int NJJ = LT::NJJ, NP = LT::NP, VDS = LT::VDS;
  extern __shared__ __align__(16) unsigned char smraw[];
  const int64_t b = blockIdx.x >> 2;
  const int hp = blockIdx.x & 3;   // head pair: heads 2 hp, 2 hp + 1
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* qrow = qkv + b * N * 384;
  const float* orow = att + b * N * E;
  const float* grow = datt + b * N * E;
  auto head_base = [&](int hl) { return reinterpret_cast<uint4*>(smraw + (size_t)hl * LT::HEAD_BYTES); };

  // ---- staging, row layout: item = (array a: Q·c, K, V, dO; head hl; row n; slot tt)
  for (int i = tid; i < 4 * 2 * NP * 4; i += 256) {
    const int tt = i & 3, n = (i >> 2) % NP, hl = (i / (4 * NP)) & 1, a = i / (8 * NP);
    const int hd = 2 * hp + hl;
    float2 x0 = make_float2(0.f, 0.f), x1 = x0;
    if (n < N) {
      const float* src = (a < 3) ? qrow + (int64_t)n * 384 + a * 128 + hd * 16 + 2 * tt : grow + (int64_t)n * E + hd * 16 + 2 * tt;
      x0 = __ldg(reinterpret_cast<const float2*>(src));
      x1 = __ldg(reinterpret_cast<const float2*>(src + 8));
      const float sc_ = (a == 0) ? QC : (a == 3 ? GS : 1.0f);
      x0.x *= sc_; x0.y *= sc_; x1.x *= sc_; x1.y *= sc_;
    }
    const uint2 p0 = split_f16x2_u(x0.x, x0.y), p1 = split_f16x2_u(x1.x, x1.y);
    head_base(hl)[a * LT::ROW_ARR + n * 4 + tt] = make_uint4(p0.x, p1.x, p0.y, p1.y);
  }
  // ---- staging, dim layout: item = (array a: Q·c, K, dO; k16 step jj; slot tt; 4 consecutive dims c4 of the pair's 32)
  for (int i = tid; i < 3 * NJJ * 4 * 8; i += 256) {
    const int tt = i & 3, c4 = ((i >> 2) & 7) * 4, jj = (i >> 5) % NJJ, a = i / (32 * NJJ);
    const int hl = c4 >> 4, d0 = c4 & 15, hd = 2 * hp + hl;
    float e[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = 16 * jj + 2 * tt + (k & 1) + 8 * (k >> 1);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (n < N) {
        const float* src = (a < 2) ? qrow + (int64_t)n * 384 + a * 128 + hd * 16 + d0 : grow + (int64_t)n * E + hd * 16 + d0;
        v = __ldg(reinterpret_cast<const float4*>(src));
        const float sc_ = (a == 0) ? QC : (a == 2 ? GS : 1.0f);
        v.x *= sc_; v.y *= sc_; v.z *= sc_; v.w *= sc_;
      }
      e[k][0] = v.x; e[k][1] = v.y; e[k][2] = v.z; e[k][3] = v.w;
    }
    uint4* dst = head_base(hl) + 4 * LT::ROW_ARR + a * LT::DIM_ARR;
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) {
      const uint2 p0 = split_f16x2_u(e[0][dd], e[1][dd]), p1 = split_f16x2_u(e[2][dd], e[3][dd]);
      dst[(d0 + dd) * VDS + ((4 * jj + tt) ^ (d0 & 4))] = make_uint4(p0.x, p1.x, p0.y, p1.y);
    }
  }
  // ---- D_i = dO_i · O_i (fp32), one thread per (head, row)
  for (int i = tid; i < 2 * NP; i += 256) {
    const int n = i % NP, hl = i / NP, hd = 2 * hp + hl;
    float D = 0.f;
    if (n < N) {
      const float4* o4 = reinterpret_cast<const float4*>(orow + (int64_t)n * E + hd * 16);
      const float4* g4 = reinterpret_cast<const float4*>(grow + (int64_t)n * E + hd * 16);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float4 o = __ldg(o4 + k), d = __ldg(g4 + k);
        D = fmaf(d.x, o.x, fmaf(d.y, o.y, fmaf(d.z, o.z, fmaf(d.w, o.w, D))));
      }
    }
    reinterpret_cast<float*>(head_base(hl) + LT::HEAD_U4)[NP + n] = D * GS;
  }
  __syncthreads();

  const int hl = warp >> 2, hd = 2 * hp + hl;
  const uint4* Qr = head_base(hl);
  const uint4* Kr = Qr + LT::ROW_ARR;
  const uint4* Vr = Kr + LT::ROW_ARR;
  const uint4* Gr = Vr + LT::ROW_ARR;
  const uint4* Qd = Gr + LT::ROW_ARR;
  const uint4* Kd = Qd + LT::DIM_ARR;
  const uint4* Gd = Kd + LT::DIM_ARR;
  float* Ls = reinterpret_cast<float*>(head_base(hl) + LT::HEAD_U4);
  const float* Ds = Ls + NP;

  // ================= pass A: 16 queries per task =================
  for (int m = warp & 3; m < NJJ; m += 4) {
    const int qa = 16 * m + g, qb = qa + 8;
    uint32_t qh[4], ql[4], gh[4], gl[4];
    frag_a(Qr, qa, qb, t, qh, ql);
    frag_a(Gr, qa, qb, t, gh, gl);
    const float Da = Ds[qa], Db = Ds[qb];
    float sc[2 * NJJ][4];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * NJJ; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      if (j < NK8) {
        const uint4 kf = Kr[(8 * j + g) * 4 + t];
        mma3_f16(sc[j], qh, ql, kf.x, kf.y, kf.z, kf.w);
        if (j == NK8 - 1 && (N & 7)) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (8 * j + 2 * t + e >= N) { sc[j][e] = -INFINITY; sc[j][2 + e] = -INFINITY; }
        }
        ma = fmaxf(ma, fmaxf(sc[j][0], sc[j][1]));
        mb = fmaxf(mb, fmaxf(sc[j][2], sc[j][3]));
      }
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * NJJ; ++j)
      if (j < NK8) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[j][e] = ex2_approx(sc[j][e] - ma);
          sc[j][2 + e] = ex2_approx(sc[j][2 + e] - mb);
          sa += sc[j][e];
          sb += sc[j][2 + e];
        }
      }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    const float ia = 1.0f / sa, ib = 1.0f / sb;
    if (t == 0) {   // log-sum-exp in log2 units, read by pass B
      Ls[qa] = ma + log2f(sa);
      Ls[qb] = mb + log2f(sb);
    }
    // dS_ij = P_ij (dO_i · V_j − D_i), in place of the scores
#pragma unroll
    for (int j = 0; j < 2 * NJJ; ++j)
      if (j < NK8) {
        float dp[4] = {0.f, 0.f, 0.f, 0.f};
        const uint4 vf = Vr[(8 * j + g) * 4 + t];
        mma3_f16(dp, gh, gl, vf.x, vf.y, vf.z, vf.w);
        sc[j][0] = sc[j][0] * ia * (dp[0] - Da);
        sc[j][1] = sc[j][1] * ia * (dp[1] - Da);
        sc[j][2] = sc[j][2] * ib * (dp[2] - Db);
        sc[j][3] = sc[j][3] * ib * (dp[3] - Db);
      }
    // dQ = s · dS K
    float dq[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d) dq[d][0] = dq[d][1] = dq[d][2] = dq[d][3] = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJJ; ++jj) {
      uint32_t ah[4], al[4];
      frag_from_c(sc[2 * jj], sc[2 * jj + 1], ah, al);
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const uint4 kf = Kd[(8 * d + g) * VDS + ((4 * jj + t) ^ (g & 4))];
        mma3_f16(dq[d], ah, al, kf.x, kf.y, kf.z, kf.w);
      }
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (qa < N) *reinterpret_cast<float2*>(dqkv + (b * N + qa) * 384 + hd * 16 + 8 * d + 2 * t) = make_float2(dq[d][0] * (0.25f * GSI), dq[d][1] * (0.25f * GSI));
      if (qb < N) *reinterpret_cast<float2*>(dqkv + (b * N + qb) * 384 + hd * 16 + 8 * d + 2 * t) = make_float2(dq[d][2] * (0.25f * GSI), dq[d][3] * (0.25f * GSI));
    }
  }
  __syncthreads();   // L of every query is in shared memory

  // ================= pass B: 16 keys per task =================
  for (int m = warp & 3; m < NJJ; m += 4) {
    const int ka = 16 * m + g, kb = ka + 8;
    uint32_t kh[4], kl[4], vh[4], vl[4];
    frag_a(Kr, ka, kb, t, kh, kl);
    frag_a(Vr, ka, kb, t, vh, vl);
    float dv[2][4], dk[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      dv[d][0] = dv[d][1] = dv[d][2] = dv[d][3] = 0.f;
      dk[d][0] = dk[d][1] = dk[d][2] = dk[d][3] = 0.f;
    }
#pragma unroll
    for (int ii = 0; ii < NJJ; ++ii) {
      float pt[2][4], ds[2][4];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int q8 = 2 * ii + u;   // query tile of 8 columns
        pt[u][0] = pt[u][1] = pt[u][2] = pt[u][3] = 0.f;
        ds[u][0] = ds[u][1] = ds[u][2] = ds[u][3] = 0.f;
        if (q8 < NK8) {
          const uint4 qf = Qr[(8 * q8 + g) * 4 + t];
          mma3_f16(pt[u], kh, kl, qf.x, qf.y, qf.z, qf.w);                      // S^T in log2 units
          const uint4 gf = Gr[(8 * q8 + g) * 4 + t];
          mma3_f16(ds[u], vh, vl, gf.x, gf.y, gf.z, gf.w);                      // dP^T
          const int c0 = 8 * q8 + 2 * t;
          const float2 L2 = *reinterpret_cast<const float2*>(Ls + c0);
          const float2 D2 = *reinterpret_cast<const float2*>(Ds + c0);
          const bool ok0 = c0 < N, ok1 = c0 + 1 < N;
          pt[u][0] = ok0 ? ex2_approx(pt[u][0] - L2.x) : 0.f;
          pt[u][1] = ok1 ? ex2_approx(pt[u][1] - L2.y) : 0.f;
          pt[u][2] = ok0 ? ex2_approx(pt[u][2] - L2.x) : 0.f;
          pt[u][3] = ok1 ? ex2_approx(pt[u][3] - L2.y) : 0.f;
          ds[u][0] = pt[u][0] * (ds[u][0] - D2.x);
          ds[u][1] = pt[u][1] * (ds[u][1] - D2.y);
          ds[u][2] = pt[u][2] * (ds[u][2] - D2.x);
          ds[u][3] = pt[u][3] * (ds[u][3] - D2.y);
        }
      }
      uint32_t ph[4], pl[4], sh[4], sl[4];
      frag_from_c(pt[0], pt[1], ph, pl);
      frag_from_c(ds[0], ds[1], sh, sl);
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const int off = (8 * d + g) * VDS + ((4 * ii + t) ^ (g & 4));
        const uint4 gf = Gd[off], qf = Qd[off];
        mma3_f16(dv[d], ph, pl, gf.x, gf.y, gf.z, gf.w);   // dV = P^T dO
        mma3_f16(dk[d], sh, sl, qf.x, qf.y, qf.z, qf.w);   // dK = dS^T (Q c)
      }
    }
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      float* oka = dqkv + (b * N + ka) * 384 + 128 + hd * 16 + 8 * d + 2 * t;
      float* okb = dqkv + (b * N + kb) * 384 + 128 + hd * 16 + 8 * d + 2 * t;
      if (ka < N) {
        *reinterpret_cast<float2*>(oka) = make_float2(dk[d][0] * (LN2 * GSI), dk[d][1] * (LN2 * GSI));
        *reinterpret_cast<float2*>(oka + 128) = make_float2(dv[d][0] * GSI, dv[d][1] * GSI);
      }
      if (kb < N) {
        *reinterpret_cast<float2*>(okb) = make_float2(dk[d][2] * (LN2 * GSI), dk[d][3] * (LN2 * GSI));
        *reinterpret_cast<float2*>(okb + 128) = make_float2(dv[d][2] * GSI, dv[d][3] * GSI);
      }
    }
  }
}

template <int NK8>
static int launch(const float* qkv, const float* att, const float* datt, float* dqkv, int64_t B, int N, cudaStream_t stream) {
  constexpr int smem = Layout<NK8>::SMEM;
  VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_bwd_mma<NK8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  // three CTAs per SM at N <= 64 (64.5 KiB each): ask for the largest shared-memory carve-out
  VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_bwd_mma<NK8>, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
  k_enc_attention_bwd_mma<NK8><<<(unsigned)(B * 4), 256, smem, stream>>>(qkv, att, datt, dqkv, N);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace ab

// dqkv [B·N][384] from qkv [B·N][384], att = O [B·N][128], datt = dO [B·N][128]
int attention_backward_mma(const float* qkv, const float* att, const float* datt, float* dqkv, int64_t B, int N, cudaStream_t stream) {
  if (N < 1 || N > VRPX_MAX_NODES || B < 1) {
    set_error("attention_backward_mma: bad argument");
    return VRPX_ERR_ARG;
  }
  switch ((N + 7) / 8) {
#define VRPX_AB_CASE(K) case K: return ab::launch<K>(qkv, att, datt, dqkv, B, N, stream);
    VRPX_AB_CASE(1) VRPX_AB_CASE(2) VRPX_AB_CASE(3) VRPX_AB_CASE(4) VRPX_AB_CASE(5) VRPX_AB_CASE(6) VRPX_AB_CASE(7) VRPX_AB_CASE(8)
    VRPX_AB_CASE(9) VRPX_AB_CASE(10) VRPX_AB_CASE(11) VRPX_AB_CASE(12) VRPX_AB_CASE(13) VRPX_AB_CASE(14) VRPX_AB_CASE(15) VRPX_AB_CASE(16)
#undef VRPX_AB_CASE
  }
  set_error("attention_backward_mma: N=%d out of range", N);
  return VRPX_ERR_ARG;
}

}  // namespace vrpx

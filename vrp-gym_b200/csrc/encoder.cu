// encoder.cu — GraphEncoder / GraphDemandEncoder forward (agents/graph_encoder.py:41-58, :95-138,
// :141-154, :183-198): node/depot embedding, 3 x { MHA + skip + BN, FF + skip + BN }.
//
// Dense contractions go through gemm_tc (tcgen05, f16 hi/lo split) or gemm_simt (fp32 FFMA cross-check);
// the per-instance N x N attention (dh = 16) and the BatchNorm reductions are SIMT kernels here.
#include "f16split.cuh"
#include "gemm.cuh"

namespace vrpx {

// ---------------------------------------------------------------- fp32 SIMT GEMM (cross-check path)
// Block tile 64x64, BK=16, 256 threads, 4x4 micro-tile.  R arbitrary; K % 16 == 0; NOUT % 64 == 0.
constexpr int SG_BM = 64, SG_BN = 64, SG_BK = 16;

__global__ void __launch_bounds__(256) k_gemm_simt(GemmArgs a) {
  __shared__ __align__(16) float Xs[SG_BK][SG_BM + 4];
  __shared__ __align__(16) float Ws[SG_BK][SG_BN + 4];
  const int tid = threadIdx.x;
  const int64_t row0 = (int64_t)blockIdx.x * SG_BM;
  const int col0 = blockIdx.y * SG_BN;
  const int lr = tid >> 2, lk = (tid & 3) * 4;  // loader: row/col lr, k offset lk
  const int ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  for (int k0 = 0; k0 < a.K; k0 += SG_BK) {
    float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + lr < a.R) xv = *reinterpret_cast<const float4*>(a.X + (row0 + lr) * a.K + k0 + lk);
    float4 wv = *reinterpret_cast<const float4*>(a.W + (int64_t)(col0 + lr) * a.K + k0 + lk);
    Xs[lk + 0][lr] = xv.x; Xs[lk + 1][lr] = xv.y; Xs[lk + 2][lr] = xv.z; Xs[lk + 3][lr] = xv.w;
    Ws[lk + 0][lr] = wv.x; Ws[lk + 1][lr] = wv.y; Ws[lk + 2][lr] = wv.z; Ws[lk + 3][lr] = wv.w;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SG_BK; ++k) {
      float4 av = *reinterpret_cast<const float4*>(&Xs[k][ty * 4]);
      float4 bv = *reinterpret_cast<const float4*>(&Ws[k][tx * 4]);
      float ar[4] = {av.x, av.y, av.z, av.w}, br[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(ar[i], br[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int64_t r = row0 + ty * 4 + i;
    if (r >= a.R) continue;
    int c = col0 + tx * 4;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float y = acc[i][j];
      if (a.gate && !(a.gate[r * a.NOUT + c + j] > 0.f)) y = 0.f;
      if (a.bias) y += a.bias[c + j];
      if (a.relu) y = fmaxf(y, 0.f);
      if (a.residual) y += a.residual[r * a.NOUT + c + j];
      if (a.scale) y = fmaf(y, a.scale[c + j], a.shift[c + j]);
      v[j] = y;
    }
    *reinterpret_cast<float4*>(a.Y + r * a.NOUT + c) = make_float4(v[0], v[1], v[2], v[3]);
  }
}

int gemm_simt(const GemmArgs& a, cudaStream_t stream) {
  if (a.K % SG_BK != 0 || a.NOUT % SG_BN != 0 || a.R <= 0) {
    set_error("gemm_simt: unsupported shape R=%lld K=%d NOUT=%d", (long long)a.R, a.K, a.NOUT);
    return VRPX_ERR_ARG;
  }
  dim3 grid((unsigned)((a.R + SG_BM - 1) / SG_BM), (unsigned)(a.NOUT / SG_BN));
  k_gemm_simt<<<grid, 256, 0, stream>>>(a);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

// ---------------------------------------------------------------- embedding (graph_encoder.py:54, :110-132)
// One thread per (row, 4 features of E).  Features are read from the env (f64 -> f32 cast, as
// graph_tsp_agent.py:72 does) or from an explicit x[R][f] array.
// One warp per row, lane = 4 output columns; the lane's slices of the (tiny) embedding weights live in registers and
// the warp strides over the rows, so a row costs one broadcast load of its features, <= 12 FMAs and one 512-byte store.
__global__ void __launch_bounds__(256) k_embed(const vrpx_encoder_weights w, const double* __restrict__ xy,
                                               const double* __restrict__ demand, const float* __restrict__ x,
                                               const int32_t* __restrict__ depot, int64_t R, int N, float* __restrict__ h) {
  const int lane = threadIdx.x & 31, e0 = lane * 4, nf = w.f;
  float nw[4][3], nb[4], dw[4][2], db[4];
  const bool has_depot = depot && w.depot_w;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    nb[j] = w.node_b[e0 + j];
#pragma unroll
    for (int i = 0; i < 3; ++i) nw[j][i] = (i < nf) ? w.node_w[(e0 + j) * nf + i] : 0.f;
    db[j] = has_depot ? w.depot_b[e0 + j] : 0.f;
    dw[j][0] = has_depot ? w.depot_w[(e0 + j) * 2] : 0.f;
    dw[j][1] = has_depot ? w.depot_w[(e0 + j) * 2 + 1] : 0.f;
  }
  const int64_t nwarps = (int64_t)gridDim.x * (blockDim.x >> 5);
  for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < R; r += nwarps) {
    float f[3] = {0.f, 0.f, 0.f};
    if (x) {
      for (int i = 0; i < nf; ++i) f[i] = x[r * nf + i];
    } else {
      const double2 p = *reinterpret_cast<const double2*>(xy + r * 2);
      f[0] = (float)p.x;
      f[1] = (float)p.y;
      if (nf == 3) f[2] = (float)demand[r];
    }
    bool is_depot = false;
    if (has_depot) {
      const int64_t b = r / N;
      is_depot = depot[b] == (int)(r - b * N);
    }
    float out[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      // same operation order as the one-thread-per-element version: bias, then the features in order
      const float yd = fmaf(f[1], dw[j][1], fmaf(f[0], dw[j][0], db[j]));
      float yn = nb[j];
#pragma unroll
      for (int i = 0; i < 3; ++i)
        if (i < nf) yn = fmaf(f[i], nw[j][i], yn);
      out[j] = is_depot ? yd : yn;
    }
    *reinterpret_cast<float4*>(h + r * E + e0) = make_float4(out[0], out[1], out[2], out[3]);
  }
}

// ---------------------------------------------------------------- per-instance self-attention
// nn.MultiheadAttention(128, 8) core (graph_encoder.py:170-172,195): 8 heads x dh 16, scale 1/4,
// no mask, no dropout.  One CTA per instance, one warp per head, lane = query node.
// qkv [R][384] rows = [q | k | v];  att [R][128] = concat_h softmax(q k^T / 4) v.
__global__ void __launch_bounds__(256) k_enc_attention(const float* __restrict__ qkv, float* __restrict__ att,
                                                        int N) {
  extern __shared__ __align__(16) float sm[];
  float* Ks = sm;                 // [8][N][16]
  float* Vs = sm + (size_t)NH * N * 16;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* base = qkv + b * N * 384;
  for (int i = tid; i < N * 64; i += 256) {  // 64 float4 per row of k|v
    int n = i >> 6, c4 = (i & 63) * 4;       // c4 in [0,256): k cols 0..127, v cols 128..255
    float4 v = *reinterpret_cast<const float4*>(base + (int64_t)n * 384 + 128 + c4);
    int c = c4 & 127, hh = c >> 4, d = c & 15;
    float* dst = (c4 < 128 ? Ks : Vs) + ((size_t)hh * N + n) * 16 + d;
    *reinterpret_cast<float4*>(dst) = v;
  }
  __syncthreads();
  const int hh = tid >> 5, lane = tid & 31;
  const float* Kh = Ks + (size_t)hh * N * 16;
  const float* Vh = Vs + (size_t)hh * N * 16;
  for (int n = lane; n < N; n += 32) {
    float q[16];
    const float4* qp = reinterpret_cast<const float4*>(base + (int64_t)n * 384 + hh * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 t = qp[i];
      q[4 * i] = t.x * 0.25f; q[4 * i + 1] = t.y * 0.25f; q[4 * i + 2] = t.z * 0.25f; q[4 * i + 3] = t.w * 0.25f;
    }
    float mx = -INFINITY;
    for (int m = 0; m < N; ++m) {
      const float4* kp = reinterpret_cast<const float4*>(Kh + m * 16);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = kp[i];
        s = fmaf(q[4 * i], t.x, s); s = fmaf(q[4 * i + 1], t.y, s);
        s = fmaf(q[4 * i + 2], t.z, s); s = fmaf(q[4 * i + 3], t.w, s);
      }
      mx = fmaxf(mx, s);
    }
    float sum = 0.f, acc[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = 0.f;
    for (int m = 0; m < N; ++m) {
      const float4* kp = reinterpret_cast<const float4*>(Kh + m * 16);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = kp[i];
        s = fmaf(q[4 * i], t.x, s); s = fmaf(q[4 * i + 1], t.y, s);
        s = fmaf(q[4 * i + 2], t.z, s); s = fmaf(q[4 * i + 3], t.w, s);
      }
      float p = expf(s - mx);
      sum += p;
      const float4* vp = reinterpret_cast<const float4*>(Vh + m * 16);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float4 t = vp[i];
        acc[4 * i] = fmaf(p, t.x, acc[4 * i]); acc[4 * i + 1] = fmaf(p, t.y, acc[4 * i + 1]);
        acc[4 * i + 2] = fmaf(p, t.z, acc[4 * i + 2]); acc[4 * i + 3] = fmaf(p, t.w, acc[4 * i + 3]);
      }
    }
    float inv = 1.0f / sum;
    float4* op = reinterpret_cast<float4*>(att + (b * N + n) * E + hh * 16);
#pragma unroll
    for (int i = 0; i < 4; ++i)
      op[i] = make_float4(acc[4 * i] * inv, acc[4 * i + 1] * inv, acc[4 * i + 2] * inv, acc[4 * i + 3] * inv);
  }
}

// ---------------------------------------------------------------- per-instance self-attention on the tensor pipe
// Same contract as k_enc_attention, computed with warp-level mma.sync.m16n8k16 on f16 hi/lo halves (f16split.cuh,
// unscaled lo: Q/4, K, V and the probabilities are O(1)), ~fp32 accuracy:
//   S = (Q/4) K^T   one k16 step covers the whole head dimension (16): 3 MMAs per 16 queries x 8 keys
//   row softmax in registers
//   O = P V         the score accumulators of two neighbouring key tiles ARE the A fragment of one k16 step over 16 keys
// One CTA per instance, one warp per head.  K and V are split ONCE while they are staged in shared memory, in fragment
// order, so the loops over query tiles read every B fragment with a single LDS.128 and do no conversions:
//   Kf [head][key][lane t: hi(dims 2t, 2t+1), hi(dims 2t+8, 2t+9), lo(..), lo(..)]              16 B per (key, t)
//   Vf [head][dim][k16 step jj][lane t: hi(keys 16jj+2t, +1), hi(keys 16jj+2t+8, +9), lo(..), lo(..)]
//      (hi pair and lo pair adjacent: each is the two-register B operand of one HMMA, no register moves)
//      dim stride padded by 4 chunks (the LDS.128 of a quarter warp — dims g in {2q, 2q+1}, t = 0..3 — is conflict free)
//      and chunk position XOR 4 on every other group of 4 dims (so are the staging stores: dims d and d + 4 per quarter)
// NJJ = ceil(N / 16) k16 steps over the keys (compile-time bound on the registers).
template <int NJJ>
__global__ void __launch_bounds__(256) k_enc_attention_f16(const float* __restrict__ qkv, float* __restrict__ att, int N) {
  extern __shared__ __align__(16) uint4 smf[];
  constexpr int NP = NJJ * 16;                     // keys padded to a multiple of 16
  constexpr int VDS = 4 * (NJJ + (NJJ & 1)) + 4;   // chunks (16 B) per dim row of Vf
  uint4* Kf = smf;                                 // [8][NP][4]
  uint4* Vf = smf + NH * NP * 4;                   // [8][16][VDS]
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* base = qkv + b * N * 384;
  // Staging: ALL global loads of the thread are issued before the first value is converted (the source-level profile of
  // the one-load-one-convert loops showed 38 % of the kernel's stall samples on the first use of a just-loaded value).
  // K: thread -> (key n, head hh, lane slot t): dims 2t, 2t+1 and 2t+8, 2t+9 of the head
  constexpr int KIT = NP * 32 / 256;   // 2 NJJ iterations
  float2 klo[KIT], khi[KIT];
#pragma unroll
  for (int it = 0; it < KIT; ++it) {
    const int i = tid + it * 256;
    const int t = i & 3, hh = (i >> 3) & 7, n = ((i >> 6) << 1) | ((i >> 2) & 1);   // a quarter warp stores keys n, n+1
    klo[it] = khi[it] = make_float2(0.f, 0.f);
    if (n < N) {
      const float* kp = base + (int64_t)n * 384 + 128 + hh * 16 + 2 * t;
      klo[it] = __ldg(reinterpret_cast<const float2*>(kp));
      khi[it] = __ldg(reinterpret_cast<const float2*>(kp + 8));
    }
  }
  // V: thread -> (k16 step jj, lane slot t, 4 consecutive dims c4..c4+3 of the 128): keys 16jj+2t, +1 and 16jj+2t+8, +9
  constexpr int VIT = (NJJ * 4 * 32 + 255) / 256;
  float4 vv[VIT][4];
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int i = tid + it * 256;
    const int t = i & 3, c4 = ((i >> 2) & 31) * 4, jj = i >> 7;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int n = 16 * jj + 2 * t + (k & 1) + 8 * (k >> 1);
      vv[it][k] = (i < NJJ * 4 * 32 && n < N) ? __ldg(reinterpret_cast<const float4*>(base + (int64_t)n * 384 + 256 + c4))
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
#pragma unroll
  for (int it = 0; it < KIT; ++it) {
    const int i = tid + it * 256;
    const int t = i & 3, hh = (i >> 3) & 7, n = ((i >> 6) << 1) | ((i >> 2) & 1);
    const uint2 p0 = split_f16x2_u(klo[it].x, klo[it].y), p1 = split_f16x2_u(khi[it].x, khi[it].y);
    Kf[(hh * NP + n) * 4 + t] = make_uint4(p0.x, p1.x, p0.y, p1.y);
  }
#pragma unroll
  for (int it = 0; it < VIT; ++it) {
    const int i = tid + it * 256;
    if (i >= NJJ * 4 * 32) break;
    const int t = i & 3, c4 = ((i >> 2) & 31) * 4, jj = i >> 7;
    const float4* v = vv[it];
    const float e[4][4] = {{v[0].x, v[0].y, v[0].z, v[0].w}, {v[1].x, v[1].y, v[1].z, v[1].w},
                           {v[2].x, v[2].y, v[2].z, v[2].w}, {v[3].x, v[3].y, v[3].z, v[3].w}};
    const int hh = c4 >> 4, d0 = c4 & 15;
#pragma unroll
    for (int dd = 0; dd < 4; ++dd) {
      const uint2 p0 = split_f16x2_u(e[0][dd], e[1][dd]), p1 = split_f16x2_u(e[2][dd], e[3][dd]);
      Vf[(hh * 16 + d0 + dd) * VDS + ((4 * jj + t) ^ (d0 & 4))] = make_uint4(p0.x, p1.x, p0.y, p1.y);
    }
  }
  __syncthreads();
  const int hh = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const uint4* Kh = Kf + hh * NP * 4;
  const uint4* Vh = Vf + hh * 16 * VDS;
  // A fragment of Q/4: a0 (row g, dims 2t, 2t+1), a1 (row g+8, same), a2 (row g, dims 2t+8, 2t+9), a3 (row g+8, same);
  // the rows of the NEXT query tile are fetched while the current one is processed
  auto load_q = [&](int q, float2& x0, float2& x1) {
    x0 = x1 = make_float2(0.f, 0.f);
    if (q < N) {
      const float* qp = base + (int64_t)q * 384 + hh * 16 + 2 * t;
      x0 = *reinterpret_cast<const float2*>(qp);
      x1 = *reinterpret_cast<const float2*>(qp + 8);
    }
  };
  float2 na0, na1, nb0, nb1;
  load_q(g, na0, na1);
  load_q(g + 8, nb0, nb1);
  for (int q0 = 0; q0 < N; q0 += 16) {
    const int qa = q0 + g, qb = q0 + g + 8;
    const float2 xa0 = na0, xa1 = na1, xb0 = nb0, xb1 = nb1;
    load_q(qa + 16, na0, na1);
    load_q(qb + 16, nb0, nb1);
    uint32_t qh[4], ql[4];
    {
      // scores are kept in log2 units (Q scaled by log2(e) / sqrt(16)): the softmax needs one ex2 per element
      constexpr float QS = 0.25f * 1.4426950408889634f;
      const uint2 s0 = split_f16x2_u(xa0.x * QS, xa0.y * QS), s1 = split_f16x2_u(xb0.x * QS, xb0.y * QS);
      const uint2 s2 = split_f16x2_u(xa1.x * QS, xa1.y * QS), s3 = split_f16x2_u(xb1.x * QS, xb1.y * QS);
      qh[0] = s0.x; qh[1] = s1.x; qh[2] = s2.x; qh[3] = s3.x;
      ql[0] = s0.y; ql[1] = s1.y; ql[2] = s2.y; ql[3] = s3.y;
    }
    // ---- S = (Q/4) K^T: C fragment of key tile j: [0], [1] = row g, keys 8j+2t, 8j+2t+1; [2], [3] = row g+8.
    // Key tiles entirely beyond N are skipped (their probabilities stay 0); only the tile that straddles N is masked.
    float sc[2 * NJJ][4];
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2 * NJJ; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      if (8 * j < N) {
        const uint4 kf = Kh[(8 * j + g) * 4 + t];   // B: (k = dims 2t.., n = key 8j+g)
        mma3_f16(sc[j], qh, ql, kf.x, kf.y, kf.z, kf.w);
        if (8 * j + 8 > N) {
#pragma unroll
          for (int e = 0; e < 2; ++e)
            if (8 * j + 2 * t + e >= N) { sc[j][e] = -INFINITY; sc[j][2 + e] = -INFINITY; }
        }
        ma = fmaxf(ma, fmaxf(sc[j][0], sc[j][1]));
        mb = fmaxf(mb, fmaxf(sc[j][2], sc[j][3]));
      }
    }
    // ---- row softmax (thread holds keys 8j+2t, 8j+2t+1 of rows qa and qb)
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < 2 * NJJ; ++j) {
      if (8 * j < N) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          sc[j][e] = ex2_approx(sc[j][e] - ma);          // 2^(-inf) = 0 for the masked keys
          sc[j][2 + e] = ex2_approx(sc[j][2 + e] - mb);
          sa += sc[j][e];
          sb += sc[j][2 + e];
        }
      }
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    // ---- O = P V: k16 step jj over keys 16jj..16jj+15; A = P from the score registers of key tiles 2jj, 2jj+1
    float o[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int jj = 0; jj < NJJ; ++jj) {
      if (16 * jj >= N) continue;
      const uint2 p0 = split_f16x2_u(sc[2 * jj][0], sc[2 * jj][1]), p1 = split_f16x2_u(sc[2 * jj][2], sc[2 * jj][3]);
      const uint2 p2 = split_f16x2_u(sc[2 * jj + 1][0], sc[2 * jj + 1][1]), p3 = split_f16x2_u(sc[2 * jj + 1][2], sc[2 * jj + 1][3]);
      const uint32_t ph[4] = {p0.x, p1.x, p2.x, p3.x}, pl[4] = {p0.y, p1.y, p2.y, p3.y};
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        const uint4 vf = Vh[(8 * d + g) * VDS + ((4 * jj + t) ^ (g & 4))];   // B: (k = keys 16jj+2t.., n = dim 8d+g)
        mma3_f16(o[d], ph, pl, vf.x, vf.y, vf.z, vf.w);
      }
    }
    const float ia = 1.0f / sa, ib = 1.0f / sb;
#pragma unroll
    for (int d = 0; d < 2; ++d) {
      if (qa < N) *reinterpret_cast<float2*>(att + (b * N + qa) * E + hh * 16 + 8 * d + 2 * t) = make_float2(o[d][0] * ia, o[d][1] * ia);
      if (qb < N) *reinterpret_cast<float2*>(att + (b * N + qb) * E + hh * 16 + 8 * d + 2 * t) = make_float2(o[d][2] * ib, o[d][3] * ib);
    }
  }
}

template <int NJJ>
static int launch_attention_t(const float* qkv, float* att, int64_t Bc, int N, cudaStream_t stream) {
  const int smem = (NH * NJJ * 16 * 4 + NH * 16 * (4 * (NJJ + (NJJ & 1)) + 4)) * (int)sizeof(uint4);
  VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_f16<NJJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  k_enc_attention_f16<NJJ><<<(unsigned)Bc, 256, smem, stream>>>(qkv, att, N);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

static int launch_attention(const float* qkv, float* att, int64_t Bc, int N, cudaStream_t stream) {
  const int njj = (N + 15) / 16;
  if (njj <= 2) return launch_attention_t<2>(qkv, att, Bc, N, stream);
  if (njj <= 4) return launch_attention_t<4>(qkv, att, Bc, N, stream);
  if (njj <= 7) return launch_attention_t<7>(qkv, att, Bc, N, stream);
  return launch_attention_t<8>(qkv, att, Bc, N, stream);
}

// ---------------------------------------------------------------- BatchNorm (graph_encoder.py:141-154)
struct BnSlots {           // per BatchNorm instance, in the small workspace
  double sum[E], sq[E];    // train: column sums of y and y^2
  float scale[E], shift[E];
};

// Column sums over R rows: block = 256 threads = 2 row lanes x 128 features.
__global__ void __launch_bounds__(256) k_bn_stats(const float* __restrict__ y, int64_t R, BnSlots* slot) {
  __shared__ double s1[256], s2[256];
  int c = threadIdx.x & 127, rl = threadIdx.x >> 7;
  double a = 0.0, q = 0.0;
  for (int64_t r = (int64_t)blockIdx.x * 2 + rl; r < R; r += (int64_t)gridDim.x * 2) {
    double v = (double)y[r * E + c];
    a += v;
    q += v * v;
  }
  s1[threadIdx.x] = a;
  s2[threadIdx.x] = q;
  __syncthreads();
  if (rl == 0) {
    atomicAdd(&slot->sum[c], s1[c] + s1[c + 128]);
    atomicAdd(&slot->sq[c], s2[c] + s2[c + 128]);
  }
}

// Fold statistics + affine into y*scale + shift.  train: batch stats (biased var for normalisation,
// unbiased for the running update, momentum 0.1); eval: running stats.  128 threads.
__global__ void k_bn_fold(BnSlots* slot, const float* __restrict__ w, const float* __restrict__ b,
                          float* run_mean, float* run_var, int train, int64_t R, float* stat_out) {
  int c = threadIdx.x;
  float mean, var;
  if (train) {
    double m = slot->sum[c] / (double)R;
    double v = slot->sq[c] / (double)R - m * m;
    if (v < 0.0) v = 0.0;
    mean = (float)m;
    var = (float)v;
    double unb = R > 1 ? v * (double)R / (double)(R - 1) : v;
    run_mean[c] = 0.9f * run_mean[c] + 0.1f * mean;
    run_var[c] = 0.9f * run_var[c] + 0.1f * (float)unb;
    slot->sum[c] = 0.0;
    slot->sq[c] = 0.0;
  } else {
    mean = run_mean[c];
    var = run_var[c];
  }
  float invstd = 1.0f / sqrtf(var + 1e-5f);
  if (stat_out) {  // saved for the backward pass: batch mean and 1/sqrt(var + eps)
    stat_out[c] = mean;
    stat_out[E + c] = invstd;
  }
  float alpha = invstd * w[c];
  slot->scale[c] = alpha;
  slot->shift[c] = b[c] - mean * alpha;
}

__global__ void k_affine(const float* __restrict__ y, const BnSlots* __restrict__ slot, int64_t n4,
                         float* __restrict__ out) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  int c = (int)(i & 31) * 4;
  float4 v = reinterpret_cast<const float4*>(y)[i];
  v.x = fmaf(v.x, slot->scale[c], slot->shift[c]);
  v.y = fmaf(v.y, slot->scale[c + 1], slot->shift[c + 1]);
  v.z = fmaf(v.z, slot->scale[c + 2], slot->shift[c + 2]);
  v.w = fmaf(v.w, slot->scale[c + 3], slot->shift[c + 3]);
  reinterpret_cast<float4*>(out)[i] = v;
}

constexpr int64_t kSmallWs = 32768;          // 6 BnSlots (3 KiB each) rounded up
constexpr int64_t kRowBytes = (512 + 128) * 4;  // qkv|att (or ff hidden) + pre-BN y, per row
// Activations saved for the backward pass (train mode), per row of R = B*N:
//   H[0..2] (inputs of the three layers) | per layer: QKV 384, ATT 128, Y1 128, H1 128, F 512, Y2 128
constexpr int64_t kSavedLayerFloats = 384 + 128 + 128 + 128 + 512 + 128;           // 1408
constexpr int64_t kSavedRowFloats = 3 * 128 + VRPX_LAYERS * kSavedLayerFloats;     // 4608
constexpr int64_t kSavedStatFloats = 2 * VRPX_LAYERS * 2 * 128;                    // 6 x (mean, invstd)

}  // namespace vrpx

using namespace vrpx;

static int g_fuse_ff = 1;     // vrpx_debug_encoder_fuse_ff
static int g_fuse_attn = 1;   // vrpx_debug_encoder_fuse_attention
namespace vrpx { extern int g_ff_dbg; }

extern "C" {

void vrpx_debug_encoder_fuse_ff(int32_t enable) {
  g_fuse_ff = enable & 1;
  vrpx::g_ff_dbg = enable >> 1;   // bits above 0: measurement switches of the kernel (ff_fused.cu, Args::dbg)
}

void vrpx_debug_encoder_fuse_attention(int32_t enable) { g_fuse_attn = enable & 1; }

int vrpx_debug_qkv_attention(const float* X, const float* in_proj_w, const float* in_proj_b, int64_t B, int32_t N, float* att,
                             void* stream) {
  VRPX_CHECK_ARG(X && in_proj_w && att, "NULL argument");
  VRPX_DEVICE_GUARD(X);
  return qkv_attention_fused(X, in_proj_w, in_proj_b, B, N, att, (cudaStream_t)stream);
}

int vrpx_debug_ff_fused(const float* X, int64_t R, const float* W1, const float* b1, const float* W2, const float* b2,
                        const float* residual, const float* scale, const float* shift, float* Y, void* stream) {
  VRPX_CHECK_ARG(X && Y, "NULL argument");
  VRPX_DEVICE_GUARD(X);
  return ff_fused(X, R, W1, b1, W2, b2, residual, scale, shift, Y, (cudaStream_t)stream);
}

int64_t vrpx_encoder_workspace_bytes(int64_t B, int32_t N) { return kSmallWs + B * (int64_t)N * kRowBytes; }

int64_t vrpx_encoder_saved_bytes(int64_t B, int32_t N) {
  return (B * (int64_t)N * kSavedRowFloats + kSavedStatFloats) * (int64_t)sizeof(float);
}

int vrpx_encoder_forward(const vrpx_encoder_weights* w, const vrpx_env* env, const float* x,
                         const int32_t* depot, int64_t B, int32_t N, int32_t train, float* h, void* ws,
                         int64_t ws_bytes, int32_t gemm_path, float* saved, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VRPX_CHECK_ARG(w && h && ws, "weights / h / ws must be non-NULL");
  VRPX_DEVICE_GUARD(h);
  NvtxRange nvtx_range(train ? "vrpx:encoder_forward(train)" : "vrpx:encoder_forward");
  VRPX_CHECK_ARG(!saved || train, "activations are only saved in train mode");
  VRPX_CHECK_ARG(B >= 1 && N >= 1 && N <= VRPX_MAX_NODES, "bad B or N");
  VRPX_CHECK_ARG(w->f == 2 || w->f == 3, "node feature count must be 2 or 3");
  VRPX_CHECK_ARG(x || (env && env->xy && (w->f == 2 || env->demand)), "need x or an env with features");
  VRPX_CHECK_ARG(!x || !env || (env->B == B && env->N == N), "env shape mismatch");
  VRPX_CHECK_ARG(ws_bytes >= kSmallWs + (int64_t)N * kRowBytes, "workspace too small for one instance");
  VRPX_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
  int64_t Bmax = (ws_bytes - kSmallWs) / ((int64_t)N * kRowBytes);
  if (train && Bmax < B) {
    set_error("vrpx_encoder_forward: train-mode BatchNorm needs the whole batch in one pass (ws %lld B < %lld B)",
              (long long)ws_bytes, (long long)vrpx_encoder_workspace_bytes(B, N));
    return VRPX_ERR_ARG;
  }
  auto gemm = [gemm_path](const GemmArgs& ga, cudaStream_t st) { return gemm_dispatch(gemm_path, ga, st); };
  BnSlots* slots = reinterpret_cast<BnSlots*>(ws);
  float* big = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kSmallWs);
  int attn_smem = 2 * NH * N * 16 * (int)sizeof(float);
  VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention, cudaFuncAttributeMaxDynamicSharedMemorySize, attn_smem));
  if (train) VRPX_CUDA(cudaMemsetAsync(ws, 0, kSmallWs, stream));
  if (!train) {
    for (int l = 0; l < VRPX_LAYERS; ++l) {
      const vrpx_encoder_layer& L = w->layer[l];
      k_bn_fold<<<1, E, 0, stream>>>(slots + 2 * l, L.bn1_w, L.bn1_b, L.bn1_mean, L.bn1_var, 0, 0, nullptr);
      VRPX_LAUNCH_CHECK();
      k_bn_fold<<<1, E, 0, stream>>>(slots + 2 * l + 1, L.bn2_w, L.bn2_b, L.bn2_mean, L.bn2_var, 0, 0, nullptr);
      VRPX_LAUNCH_CHECK();
    }
  }
  for (int64_t b0 = 0; b0 < B; b0 += Bmax) {
    int64_t Bc = (B - b0 < Bmax) ? (B - b0) : Bmax;
    int64_t R = Bc * N;
    float* hc = h + b0 * N * E;
    float* qkv = big;                 // [R][384]
    float* att = big + R * 384;       // [R][128]
    float* hid = big;                 // [R][512]  (aliases qkv|att, which are dead by then)
    float* ybuf = big + R * 512;      // [R][128]  pre-BN activations (train)
    float* sv_stats = saved ? saved + R * kSavedRowFloats : nullptr;
    if (saved) hc = saved;            // H[0]: the embedding is the input of layer 0
    {
      const int64_t want = (R + 7) / 8, cap = (int64_t)num_sms() * 8;   // 8 warps (rows in flight) per CTA
      k_embed<<<(unsigned)(want < cap ? want : cap), 256, 0, stream>>>(
          *w, (!x && env) ? env->xy + b0 * N * 2 : nullptr,
          (!x && env && env->demand) ? env->demand + b0 * N : nullptr, x ? x + b0 * N * w->f : nullptr,
          depot ? depot + b0 : nullptr, R, N, hc);
      VRPX_LAUNCH_CHECK();
    }
    for (int l = 0; l < VRPX_LAYERS; ++l) {
      const vrpx_encoder_layer& L = w->layer[l];
      BnSlots* s1 = slots + 2 * l;
      BnSlots* s2 = slots + 2 * l + 1;
      float* h1 = hc;   // output of BN1 (in place unless activations are saved)
      float* hout = hc; // output of BN2 = input of the next layer
      float *y1 = ybuf, *y2 = ybuf;
      if (saved) {
        float* Lb = saved + R * 384 + (int64_t)l * R * kSavedLayerFloats;
        qkv = Lb; att = Lb + R * 384; y1 = Lb + R * 512; h1 = Lb + R * 640; hid = Lb + R * 768; y2 = Lb + R * 1280;
        hout = (l + 1 < VRPX_LAYERS) ? saved + (int64_t)(l + 1) * R * 128 : h + b0 * N * E;
      }
      int rc = 0;
      if (gemm_path == 0 && !saved && g_fuse_attn) {   // Q, K, V stay on chip (attn_fused.cu)
        if ((rc = qkv_attention_fused(hc, L.in_proj_w, L.in_proj_b, Bc, N, att, stream))) return rc;
      } else {
        GemmArgs g1{hc, R, E, L.in_proj_w, 3 * E, L.in_proj_b, 0, nullptr, nullptr, nullptr, qkv};
        if ((rc = gemm(g1, stream))) return rc;
        if (gemm_path == 1) {  // fp32 SIMT cross-check path
          k_enc_attention<<<(unsigned)Bc, 256, attn_smem, stream>>>(qkv, att, N);
          VRPX_LAUNCH_CHECK();
        } else if ((rc = launch_attention(qkv, att, Bc, N, stream))) {
          return rc;
        }
      }
      if (!train) {
        GemmArgs g2{att, R, E, L.out_proj_w, E, L.out_proj_b, 0, hc, s1->scale, s1->shift, hc};
        if ((rc = gemm(g2, stream))) return rc;
      } else {
        GemmArgs g2{att, R, E, L.out_proj_w, E, L.out_proj_b, 0, hc, nullptr, nullptr, y1};
        if ((rc = gemm(g2, stream))) return rc;
        k_bn_stats<<<num_sms() * 4, 256, 0, stream>>>(y1, R, s1);
        VRPX_LAUNCH_CHECK();
        k_bn_fold<<<1, E, 0, stream>>>(s1, L.bn1_w, L.bn1_b, L.bn1_mean, L.bn1_var, 1, R,
                                       sv_stats ? sv_stats + (2 * l) * 256 : nullptr);
        VRPX_LAUNCH_CHECK();
        k_affine<<<(unsigned)((R * 32 + 255) / 256), 256, 0, stream>>>(y1, s1, R * 32, h1);
        VRPX_LAUNCH_CHECK();
      }
      const bool fuse_ff = gemm_path == 0 && !saved && g_fuse_ff;   // the hidden tile stays on chip (ff_fused.cu)
      if (fuse_ff) {
        if (!train) {
          if ((rc = ff_fused(h1, R, L.ff0_w, L.ff0_b, L.ff2_w, L.ff2_b, h1, s2->scale, s2->shift, hc, stream))) return rc;
          continue;
        }
        if ((rc = ff_fused(h1, R, L.ff0_w, L.ff0_b, L.ff2_w, L.ff2_b, h1, nullptr, nullptr, y2, stream))) return rc;
      } else {
        GemmArgs g3{h1, R, E, L.ff0_w, FF, L.ff0_b, 1, nullptr, nullptr, nullptr, hid};
        if ((rc = gemm(g3, stream))) return rc;
      }
      if (!train) {
        GemmArgs g4{hid, R, FF, L.ff2_w, E, L.ff2_b, 0, hc, s2->scale, s2->shift, hc};
        if ((rc = gemm(g4, stream))) return rc;
      } else {
        if (!fuse_ff) {
          GemmArgs g4{hid, R, FF, L.ff2_w, E, L.ff2_b, 0, h1, nullptr, nullptr, y2};
          if ((rc = gemm(g4, stream))) return rc;
        }
        k_bn_stats<<<num_sms() * 4, 256, 0, stream>>>(y2, R, s2);
        VRPX_LAUNCH_CHECK();
        k_bn_fold<<<1, E, 0, stream>>>(s2, L.bn2_w, L.bn2_b, L.bn2_mean, L.bn2_var, 1, R,
                                       sv_stats ? sv_stats + (2 * l + 1) * 256 : nullptr);
        VRPX_LAUNCH_CHECK();
        k_affine<<<(unsigned)((R * 32 + 255) / 256), 256, 0, stream>>>(y2, s2, R * 32, hout);
        VRPX_LAUNCH_CHECK();
      }
      hc = hout;
    }
  }
  return VRPX_OK;
}

}  // extern "C"

// encoder.cu — GraphEncoder / GraphDemandEncoder forward (agents/graph_encoder.py:41-58, :95-138,
// :141-154, :183-198): node/depot embedding, 3 x { MHA + skip + BN, FF + skip + BN }.
//
// Dense contractions go through gemm_tc (tcgen05, f16 hi/lo split) or gemm_simt (fp32 FFMA cross-check);
// the per-instance N x N attention (dh = 16) and the BatchNorm reductions are SIMT kernels here.
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
__global__ void k_embed(const vrpx_encoder_weights w, const double* __restrict__ xy,
                        const double* __restrict__ demand, const float* __restrict__ x,
                        const int32_t* __restrict__ depot, int64_t R, int N, float* __restrict__ h) {
  int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * (E / 4)) return;
  int64_t r = idx / (E / 4);
  int e0 = (int)(idx - r * (E / 4)) * 4;
  float f[3] = {0.f, 0.f, 0.f};
  if (x) {
    for (int i = 0; i < w.f; ++i) f[i] = x[r * w.f + i];
  } else {
    f[0] = (float)xy[r * 2];
    f[1] = (float)xy[r * 2 + 1];
    if (w.f == 3) f[2] = (float)demand[r];
  }
  bool is_depot = false;
  if (depot && w.depot_w) {
    int64_t b = r / N;
    is_depot = depot[b] == (int)(r - b * N);
  }
  float out[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    int e = e0 + j;
    float y;
    if (is_depot) {
      y = w.depot_b[e];
      y = fmaf(f[0], w.depot_w[e * 2], y);
      y = fmaf(f[1], w.depot_w[e * 2 + 1], y);
    } else {
      y = w.node_b[e];
      for (int i = 0; i < w.f; ++i) y = fmaf(f[i], w.node_w[e * w.f + i], y);
    }
    out[j] = y;
  }
  *reinterpret_cast<float4*>(h + r * E + e0) = make_float4(out[0], out[1], out[2], out[3]);
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
// Same contract as k_enc_attention, computed with warp-level mma.sync.m16n8k8 TF32 and the 3-term split (~fp32):
//   S = (Q/4) K^T  (16 queries x 8 keys per mma, 2 k-steps over dh = 16), row softmax in registers,
//   O = P V        (the score accumulators are reused directly as A fragments: key order inside a key tile is
//                   permuted, k-index t <-> key 2t, k-index t+4 <-> key 2t+1, and V is read with the same order).
// One CTA per instance, one warp per head.  NTILES = ceil(N / 8) key tiles (compile-time bound on the registers).
__device__ __forceinline__ void split_tf32_e(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_e(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma3_e(float (&c)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                       uint32_t bl0, uint32_t bh1, uint32_t bl1) {
  mma_tf32_e(c, al, bh0, bh1);
  mma_tf32_e(c, ah, bl0, bl1);
  mma_tf32_e(c, ah, bh0, bh1);
}

constexpr int ATT_VLD = 20;  // padded V row stride (floats): B fragments (key 2t / 2t+1, dim g) hit 32 distinct banks

template <int NTILES>
__global__ void __launch_bounds__(256) k_enc_attention_mma(const float* __restrict__ qkv, float* __restrict__ att, int N) {
  extern __shared__ __align__(16) float sm[];
  const int NP = NTILES * 8;                       // keys padded to a multiple of 8
  float* Ks = sm;                                  // [8 heads][NP][16]
  float* Vs = sm + (size_t)NH * NP * 16;           // [8 heads][NP][ATT_VLD]
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const float* base = qkv + b * N * 384;
  for (int i = tid; i < NP * 64; i += 256) {       // 64 float4 per row of k|v
    const int n = i >> 6, c4 = (i & 63) * 4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (n < N) v = *reinterpret_cast<const float4*>(base + (int64_t)n * 384 + 128 + c4);
    const int c = c4 & 127, hh = c >> 4, d = c & 15;
    if (c4 < 128) *reinterpret_cast<float4*>(Ks + ((size_t)hh * NP + n) * 16 + d) = v;
    else *reinterpret_cast<float4*>(Vs + ((size_t)hh * NP + n) * ATT_VLD + d) = v;
  }
  __syncthreads();
  const int hh = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const float* Kh = Ks + (size_t)hh * NP * 16;
  const float* Vh = Vs + (size_t)hh * NP * ATT_VLD;
  for (int q0 = 0; q0 < N; q0 += 16) {
    const int qa = q0 + g, qb = q0 + g + 8;
    // A fragments of Q/4: thread t owns dims 4t..4t+3; k-step u: k-index t <-> dim 4t+2u, t+4 <-> dim 4t+2u+1
    float4 xa = make_float4(0.f, 0.f, 0.f, 0.f), xb = xa;
    if (qa < N) xa = *reinterpret_cast<const float4*>(base + (int64_t)qa * 384 + hh * 16 + 4 * t);
    if (qb < N) xb = *reinterpret_cast<const float4*>(base + (int64_t)qb * 384 + hh * 16 + 4 * t);
    const float ea[4] = {xa.x * 0.25f, xa.y * 0.25f, xa.z * 0.25f, xa.w * 0.25f};
    const float eb[4] = {xb.x * 0.25f, xb.y * 0.25f, xb.z * 0.25f, xb.w * 0.25f};
    uint32_t qh[2][4], ql[2][4];
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      split_tf32_e(ea[2 * u], qh[u][0], ql[u][0]);
      split_tf32_e(eb[2 * u], qh[u][1], ql[u][1]);
      split_tf32_e(ea[2 * u + 1], qh[u][2], ql[u][2]);
      split_tf32_e(eb[2 * u + 1], qh[u][3], ql[u][3]);
    }
    // ---- S = (Q/4) K^T
    float sc[NTILES][4];
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
      sc[j][0] = sc[j][1] = sc[j][2] = sc[j][3] = 0.f;
      const float4 kv = *reinterpret_cast<const float4*>(Kh + (8 * j + g) * 16 + 4 * t);   // key 8j+g, dims 4t..4t+3
      const float ke[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32_e(ke[2 * u], bh0, bl0);
        split_tf32_e(ke[2 * u + 1], bh1, bl1);
        mma3_e(sc[j], qh[u], ql[u], bh0, bl0, bh1, bl1);
      }
    }
    // ---- row softmax: thread holds keys 8j + 2t, 8j + 2t + 1 of rows qa (sc[j][0..1]) and qb (sc[j][2..3])
    float ma = -INFINITY, mb = -INFINITY;
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = 8 * j + 2 * t + e < N;
        if (!ok) { sc[j][e] = -INFINITY; sc[j][2 + e] = -INFINITY; }
        ma = fmaxf(ma, sc[j][e]);
        mb = fmaxf(mb, sc[j][2 + e]);
      }
    }
    ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
    mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        sc[j][e] = expf(sc[j][e] - ma);          // exp(-inf) = 0 for padded keys
        sc[j][2 + e] = expf(sc[j][2 + e] - mb);
        sa += sc[j][e];
        sb += sc[j][2 + e];
      }
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1); sa += __shfl_xor_sync(0xffffffffu, sa, 2);
    sb += __shfl_xor_sync(0xffffffffu, sb, 1); sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    // ---- O = P V  (A = P from the score registers; k-index t <-> key 8j+2t, t+4 <-> key 8j+2t+1)
    float o[2][4];
#pragma unroll
    for (int d = 0; d < 2; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NTILES; ++j) {
      uint32_t ph[4], pl[4];
      split_tf32_e(sc[j][0], ph[0], pl[0]);   // (row g,   k = t)
      split_tf32_e(sc[j][2], ph[1], pl[1]);   // (row g+8, k = t)
      split_tf32_e(sc[j][1], ph[2], pl[2]);   // (row g,   k = t+4)
      split_tf32_e(sc[j][3], ph[3], pl[3]);   // (row g+8, k = t+4)
      const float* v0 = Vh + (8 * j + 2 * t) * ATT_VLD + g;
#pragma unroll
      for (int d = 0; d < 2; ++d) {
        uint32_t bh0, bl0, bh1, bl1;
        split_tf32_e(v0[8 * d], bh0, bl0);               // (k = t,   n = dim 8d+g) = V[key 8j+2t][8d+g]
        split_tf32_e(v0[ATT_VLD + 8 * d], bh1, bl1);     // (k = t+4, n = dim 8d+g) = V[key 8j+2t+1][8d+g]
        mma3_e(o[d], ph, pl, bh0, bl0, bh1, bl1);
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

static int launch_attention(const float* qkv, float* att, int64_t Bc, int N, cudaStream_t stream) {
  const int nt = (N + 7) / 8;
  const int NP = (nt <= 7 ? 7 : (nt <= 13 ? 13 : 16)) * 8;
  const int smem = NH * NP * (16 + ATT_VLD) * (int)sizeof(float);
  if (nt <= 7) {
    VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_mma<7>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_enc_attention_mma<7><<<(unsigned)Bc, 256, smem, stream>>>(qkv, att, N);
  } else if (nt <= 13) {
    VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_mma<13>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_enc_attention_mma<13><<<(unsigned)Bc, 256, smem, stream>>>(qkv, att, N);
  } else {
    VRPX_CUDA(cudaFuncSetAttribute(k_enc_attention_mma<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_enc_attention_mma<16><<<(unsigned)Bc, 256, smem, stream>>>(qkv, att, N);
  }
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
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

extern "C" {

int64_t vrpx_encoder_workspace_bytes(int64_t B, int32_t N) { return kSmallWs + B * (int64_t)N * kRowBytes; }

int64_t vrpx_encoder_saved_bytes(int64_t B, int32_t N) {
  return (B * (int64_t)N * kSavedRowFloats + kSavedStatFloats) * (int64_t)sizeof(float);
}

int vrpx_encoder_forward(const vrpx_encoder_weights* w, const vrpx_env* env, const float* x,
                         const int32_t* depot, int64_t B, int32_t N, int32_t train, float* h, void* ws,
                         int64_t ws_bytes, int32_t gemm_path, float* saved, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VRPX_CHECK_ARG(w && h && ws, "weights / h / ws must be non-NULL");
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
      int64_t n = R * (E / 4);
      k_embed<<<(unsigned)((n + 255) / 256), 256, 0, stream>>>(
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
      GemmArgs g1{hc, R, E, L.in_proj_w, 3 * E, L.in_proj_b, 0, nullptr, nullptr, nullptr, qkv};
      if ((rc = gemm(g1, stream))) return rc;
      if (gemm_path == 1) {  // fp32 SIMT cross-check path
        k_enc_attention<<<(unsigned)Bc, 256, attn_smem, stream>>>(qkv, att, N);
        VRPX_LAUNCH_CHECK();
      } else if ((rc = launch_attention(qkv, att, Bc, N, stream))) {
        return rc;
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
      GemmArgs g3{h1, R, E, L.ff0_w, FF, L.ff0_b, 1, nullptr, nullptr, nullptr, hid};
      if ((rc = gemm(g3, stream))) return rc;
      if (!train) {
        GemmArgs g4{hid, R, FF, L.ff2_w, E, L.ff2_b, 0, hc, s2->scale, s2->shift, hc};
        if ((rc = gemm(g4, stream))) return rc;
      } else {
        GemmArgs g4{hid, R, FF, L.ff2_w, E, L.ff2_b, 0, h1, nullptr, nullptr, y2};
        if ((rc = gemm(g4, stream))) return rc;
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

// rollout.cu — the fused decoder + environment rollout: ONE persistent cooperative launch loops over
// every decode step of the episode (agents/graph_tsp_agent.py:78-92, graph_vrp_agent.py:69-83,
// graph_irp_agent.py:82-105 with agents/graph_decoder.py:51-115 and gym_vrp/envs/*.py step fused).
//
// Per step and per tile of 32 instances a CTA runs:
//   P0  gather X = h[b, last]                                         (graph_decoder.py:108-109)
//   P1  GEMM-A  q~ = X · A_l^T + Q~g[b]  (+ load · a_load)             context -> per-head folded query
//   P2  per instance (one warp): scores = q~_h · h_n + scrambled additive mask (graph_decoder.py:93-94),
//       softmax over nodes, c_h = sum_n p_hn h_n
//   P3  GEMM-B  q^ = C · M^T + m_c                                     (V-proj, out-proj, _att_output, _kp folded)
//   P4  per instance: u_n = 10 tanh(q^ · h_n), -inf mask, argmax / Philox sample / teacher action,
//       log-prob, then the environment transition (env_rules.cuh) in f64
// followed by a grid-wide barrier: the reference's glimpse mask of instance b reads the masks of
// instances (8b+h) mod G, so all instances advance in lock-step.
//
// Step-invariant work the reference repeats every step (K/V/kp projections, graph mean) is folded into
// host-packed weights (vrpx/packing.py) and the per-episode Q~g table built in the prologue.
#include "env_rules.cuh"

namespace vrpx {

constexpr int TM = 32;        // instances per tile
constexpr int NT = 512;       // threads per CTA (16 warps; <= 128 registers per thread)
constexpr int QW = NH * E;    // 1024: per-instance width of q~ / c
constexpr size_t SMEM_X = (size_t)TM * E * sizeof(float);    // 16 KiB
constexpr size_t SMEM_QC = (size_t)TM * QW * sizeof(float);  // 128 KiB
constexpr int WCHUNK_FLOATS = 16 * 512;                      // one staged weight chunk: 32 KiB
constexpr size_t SMEM_W = 2 * (size_t)WCHUNK_FLOATS * sizeof(float);  // double buffer, 64 KiB
constexpr size_t SMEM_TOTAL = SMEM_X + SMEM_QC + SMEM_W;

struct RolloutParams {
  vrpx_env env;
  vrpx_decoder_weights w;
  const float* h;
  int mode;
  long long G;
  unsigned long long seed, offset;
  uint8_t* tape;
  int t0;
  int Tmax;
  float* logp;
  float* cost;
  int* steps;
  float* logits;
  float* qg;          // [B][1024]
  unsigned* bar;      // grid barrier counter
  int* notdone;       // [Tmax + 1]
};

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_acquire_i(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    while (ld_acquire(bar) < target) __nanosleep(32);
    __threadfence();
  }
  __syncthreads();
}

// Sum 8 per-lane values across the warp; lane l returns the total of v[(l >> 2) & 7].
__device__ __forceinline__ float reduce8(const float v[8], int lane) {
  const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
  float w4[4], w2[2], x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float keep = b4 ? v[i + 4] : v[i], send = b4 ? v[i] : v[i + 4];
    w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float keep = b3 ? w4[i + 2] : w4[i], send = b3 ? w4[i] : w4[i + 2];
    w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
  }
  {
    float keep = b2 ? w2[1] : w2[0], send = b2 ? w2[0] : w2[1];
    x = keep + __shfl_xor_sync(0xffffffffu, send, 4);
  }
  x += __shfl_xor_sync(0xffffffffu, x, 2);
  x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---------------------------------------------------------------- GEMM-A: [TM x 128] · [128 x 1024]
// Xs smem [TM][128]; Wt global [128][1024], streamed through a double-buffered smem stage with cp.async in
// chunks of 16 k-rows x 512 columns (32 KiB).  512 threads = 4 row groups x 128 column threads; each thread owns
// 8 rows x 4 columns {4tx..4tx+3} of the current 512-column half (conflict-free LDS.128).
// EPI 0: qg[b][c] = acc + a_c[c]               (prologue: graph-embedding term + bias)
// EPI 1: qg[b][c] += acc                       (step 1: `first` term, graph_decoder.py:111-113)
// EPI 2: QC[m][c] = acc + qg[b][c] + loadf[m] * a_load[c]
__device__ __forceinline__ void stage_a_chunk(const float* __restrict__ Wt, int chunk, float* __restrict__ dst) {
  const int half = chunk >> 3, k0 = (chunk & 7) * 16;
  const float* src = Wt + (size_t)k0 * QW + half * 512;
#pragma unroll
  for (int i = 0; i < 2048 / NT; ++i) {
    int idx = threadIdx.x + NT * i;       // 2048 float4 per chunk
    int r = idx >> 7, c4 = idx & 127;
    cp_async16(dst + r * 512 + c4 * 4, src + (size_t)r * QW + c4 * 4);
  }
}

template <int EPI>
__device__ __forceinline__ void gemm_a(const float* __restrict__ Xs, const float* __restrict__ Wt,
                                       float* __restrict__ QC, float* __restrict__ Wb, const RolloutParams& p,
                                       int64_t base, int cnt, const float* __restrict__ loadf) {
  const int tid = threadIdx.x, ty = tid >> 7, tx = tid & 127;
  stage_a_chunk(Wt, 0, Wb);
  cp_async_commit();
  float acc[8][4];
  for (int chunk = 0; chunk < 16; ++chunk) {
    const int half = chunk >> 3, k0 = (chunk & 7) * 16;
    if ((chunk & 7) == 0) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    }
    if (chunk + 1 < 16) {
      stage_a_chunk(Wt, chunk + 1, Wb + ((chunk + 1) & 1) * WCHUNK_FLOATS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = Wb + (chunk & 1) * WCHUNK_FLOATS + tx * 4;
#pragma unroll
    for (int kq = 0; kq < 16; kq += 4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(Xs + (ty * 8 + i) * E + k0 + kq);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (kq + kk) * 512);
        const float wv[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = kk == 0 ? xv[i].x : (kk == 1 ? xv[i].y : (kk == 2 ? xv[i].z : xv[i].w));
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x, wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();  // the stage may be refilled by the next iteration's cp.async
    if ((chunk & 7) == 7) {
      const int c = half * 512 + tx * 4;
      float4 ac = make_float4(0.f, 0.f, 0.f, 0.f), al = ac;
      if (EPI == 0) ac = *reinterpret_cast<const float4*>(p.w.a_c + c);
      if (EPI == 2 && p.env.kind == VRPX_IRP) al = *reinterpret_cast<const float4*>(p.w.a_load + c);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int m = ty * 8 + i;
        if (m >= cnt) continue;
        const int64_t b = base + m;
        float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        float4* qgp = reinterpret_cast<float4*>(p.qg + b * QW + c);
        if (EPI == 0) {
          *qgp = make_float4(v.x + ac.x, v.y + ac.y, v.z + ac.z, v.w + ac.w);
        } else if (EPI == 1) {
          float4 q = *qgp;
          *qgp = make_float4(q.x + v.x, q.y + v.y, q.z + v.z, q.w + v.w);
        } else {
          const float4 q = *qgp;
          v = make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
          if (p.env.kind == VRPX_IRP) {
            const float lf = loadf[m];
            v = make_float4(fmaf(lf, al.x, v.x), fmaf(lf, al.y, v.y), fmaf(lf, al.z, v.z), fmaf(lf, al.w, v.w));
          }
          *reinterpret_cast<float4*>(QC + m * QW + c) = v;
        }
      }
    }
  }
}

// ---------------------------------------------------------------- GEMM-B: [TM x 1024] · [1024 x 128]
// C smem [TM][1024]; Mt global [1024][128] staged with cp.async: chunk kc = rows {kg*256 + kc*16 + r} of the four
// k-groups (4 x 16 rows x 128 columns = 32 KiB).  512 threads = 4 k-groups x 4 row groups x 32 column threads,
// 8 rows x 4 columns {4tx..+3} each over a quarter of K; partial sums reduced through smem.
// Result q^[m][e] (+ m_c) is written to Xs[TM][128].
__device__ __forceinline__ void stage_b_chunk(const float* __restrict__ Mt, int kc, float* __restrict__ dst) {
#pragma unroll
  for (int i = 0; i < 2048 / NT; ++i) {
    int idx = threadIdx.x + NT * i;       // 2048 float4 per chunk
    int row = idx >> 5, c4 = idx & 31;    // row in [0,64): kg = row >> 4, r = row & 15
    int k = (row >> 4) * 256 + kc * 16 + (row & 15);
    cp_async16(dst + row * E + c4 * 4, Mt + (size_t)k * E + c4 * 4);
  }
}

__device__ __forceinline__ void gemm_b(float* __restrict__ QC, float* __restrict__ Xs, float* __restrict__ Wb,
                                       const RolloutParams& p) {
  const int tid = threadIdx.x, kg = tid >> 7, ty = (tid >> 5) & 3, tx = tid & 31;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  stage_b_chunk(p.w.m_t, 0, Wb);
  cp_async_commit();
  for (int kc = 0; kc < 16; ++kc) {
    if (kc + 1 < 16) {
      stage_b_chunk(p.w.m_t, kc + 1, Wb + ((kc + 1) & 1) * WCHUNK_FLOATS);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const float* wb = Wb + (kc & 1) * WCHUNK_FLOATS + kg * 16 * E + tx * 4;
    const int k0 = kg * 256 + kc * 16;
#pragma unroll
    for (int kq = 0; kq < 16; kq += 4) {
      float4 xv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(QC + (ty * 8 + i) * QW + k0 + kq);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        const float4 w0 = *reinterpret_cast<const float4*>(wb + (kq + kk) * E);
        const float wv[4] = {w0.x, w0.y, w0.z, w0.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float x = kk == 0 ? xv[i].x : (kk == 1 ? xv[i].y : (kk == 2 ? xv[i].z : xv[i].w));
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(x, wv[j], acc[i][j]);
        }
      }
    }
    __syncthreads();  // also orders the last reads of C before the partials overwrite it
  }
  float* part = QC;  // [4][TM][128]
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(part + (kg * TM + ty * 8 + i) * E + tx * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  for (int o = tid; o < TM * E; o += NT) {
    float s = part[o] + part[TM * E + o] + part[2 * TM * E + o] + part[3 * TM * E + o];
    Xs[o] = s + p.w.m_c[o & (E - 1)];
  }
  __syncthreads();
}

// glimpse-mask source row for attention row (b, hh): mask.repeat(H,1) indexing (graph_decoder.py:93)
__device__ __forceinline__ int64_t quirk_row(int64_t b, int hh, long long G) {
  if (G <= 0) return b;
  int64_t g0 = (b / G) * G;
  return g0 + (((b - g0) * NH + hh) % G);
}

// ---------------------------------------------------------------- the persistent kernel
__global__ void __launch_bounds__(NT, 1) k_rollout(const RolloutParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Xs = reinterpret_cast<float*>(smem_raw);
  float* QC = reinterpret_cast<float*>(smem_raw + SMEM_X);
  float* Wb = reinterpret_cast<float*>(smem_raw + SMEM_X + SMEM_QC);
  __shared__ float s_loadf[TM];
  __shared__ int s_anyleft;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.env.N, kind = p.env.kind;
  const int64_t B = p.env.B;
  const int64_t ntiles = (B + TM - 1) / TM;
  const float* __restrict__ h = p.h;
  unsigned bar_target = 0;

  // ------------------------------------------------ prologue: Q~g[b] = A_g · mean_n h[b,n] + a_c
  for (int64_t tile = blockIdx.x; tile < ntiles && p.t0 == 0; tile += gridDim.x) {
    const int64_t base = tile * TM;
    const int cnt = (int)((B - base < TM) ? (B - base) : TM);
    for (int m = warp; m < TM; m += NT / 32) {
      float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
      if (m < cnt) {
        const float4* hp = reinterpret_cast<const float4*>(h + (base + m) * N * E) + lane;
        for (int n = 0; n < N; ++n) {
          float4 v = __ldg(hp + n * (E / 4));
          g.x += v.x; g.y += v.y; g.z += v.z; g.w += v.w;
        }
        float inv = 1.0f / (float)N;
        g.x *= inv; g.y *= inv; g.z *= inv; g.w *= inv;
      }
      *reinterpret_cast<float4*>(Xs + m * E + lane * 4) = g;
    }
    __syncthreads();
    gemm_a<0>(Xs, p.w.ag_t, QC, Wb, p, base, cnt, s_loadf);
    __syncthreads();
  }

  int t = p.t0;
  for (; t < p.t0 + p.Tmax; ++t) {
    const int trel = t - p.t0;
    bool cta_unfinished = false;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t base = tile * TM;
      const int cnt = (int)((B - base < TM) ? (B - base) : TM);
      if (tid == 0) s_anyleft = 0;
      // ---------------- P0: gather last-node embeddings, vehicle load
      for (int m = warp; m < TM; m += NT / 32) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < cnt && t > 0) {
          int last = p.env.cur[base + m];
          xv = __ldg(reinterpret_cast<const float4*>(h + ((base + m) * N + last) * E) + lane);
        }
        *reinterpret_cast<float4*>(Xs + m * E + lane * 4) = xv;
        if (lane == 0) s_loadf[m] = (m < cnt) ? (float)p.env.load[base + m] : 0.f;
      }
      __syncthreads();
      // ---------------- P1: q~
      if (t == 0) {
        for (int o = tid; o < cnt * QW; o += NT) {
          int m = o >> 10, c = o & (QW - 1);
          float y = p.qg[(base + m) * QW + c] + p.w.a_q0[c];
          if (kind == VRPX_IRP) y = fmaf(s_loadf[m], p.w.a_load[c], y);
          QC[o] = y;
        }
      } else {
        if (t == 1 && kind != VRPX_IRP) {
          gemm_a<1>(Xs, p.w.af_t, QC, Wb, p, base, cnt, s_loadf);
          __syncthreads();  // qg updates are re-read by other threads' epilogue below? (same thread) — keep ordering explicit
        }
        gemm_a<2>(Xs, p.w.al_t, QC, Wb, p, base, cnt, s_loadf);
      }
      __syncthreads();

      // ---------------- P2: glimpse attention, one warp per instance
      for (int m = warp; m < cnt; m += NT / 32) {
        const int64_t b = base + m;
        float* slot = QC + m * QW;
        float4 qt[NH];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) qt[hh] = *reinterpret_cast<const float4*>(slot + hh * E + lane * 4);
        __syncwarp();
        // this lane's head (lane >> 2) & 7 reads the mask of instance quirk_row(b, head)
        const int myh = (lane >> 2) & 7;
        const uint32_t* nbm = p.env.mask + quirk_row(b, myh, p.G) * 4;
        uint32_t nb[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) nb[i] = __ldcg(nbm + i);
        const float4* hp = reinterpret_cast<const float4*>(h + b * N * E) + lane;
        // pass 1: scores[hh][n]; rows are fetched four at a time, one batch ahead of their use
        {
          float4 nxt[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) nxt[i] = (i < N) ? __ldg(hp + i * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int n0 = 0; n0 < N; n0 += 4) {
            float4 hv4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hv4[i] = nxt[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = n0 + 4 + i;
              if (n < N) nxt[i] = __ldg(hp + n * (E / 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = n0 + i;
              if (n < N) {
                const float4 hv = hv4[i];
                float v[NH];
#pragma unroll
                for (int hh = 0; hh < NH; ++hh)
                  v[hh] = fmaf(qt[hh].x, hv.x, fmaf(qt[hh].y, hv.y, fmaf(qt[hh].z, hv.z, qt[hh].w * hv.w)));
                float sc = reduce8(v, lane);
                if ((lane & 3) == 0) slot[myh * E + n] = sc + (float)((nb[n >> 5] >> (n & 31)) & 1u);
              }
            }
          }
        }
        __syncwarp();
        // softmax per head over nodes (lane = node, 4 strides cover N <= 128)
        float pr[NH][4];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int n = lane + 32 * i;
            pr[hh][i] = (n < N) ? slot[hh * E + n] : -INFINITY;
            mx = fmaxf(mx, pr[hh][i]);
          }
          mx = warp_max(mx);
          float sum = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            int n = lane + 32 * i;
            pr[hh][i] = (n < N) ? expf(pr[hh][i] - mx) : 0.f;
            sum += pr[hh][i];
          }
          sum = warp_sum(sum);
          float inv = 1.0f / sum;
#pragma unroll
          for (int i = 0; i < 4; ++i) pr[hh][i] *= inv;
        }
        __syncwarp();
        // probabilities to the slot as P[n][8]
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int n = lane + 32 * i;
          if (n < N) {
            *reinterpret_cast<float4*>(slot + n * 8) = make_float4(pr[0][i], pr[1][i], pr[2][i], pr[3][i]);
            *reinterpret_cast<float4*>(slot + n * 8 + 4) = make_float4(pr[4][i], pr[5][i], pr[6][i], pr[7][i]);
          }
        }
        __syncwarp();
        // pass 2: c[hh][4 dims of this lane]
        float4 c[NH];
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) c[hh] = make_float4(0.f, 0.f, 0.f, 0.f);
        {
          float4 nxt[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) nxt[i] = (i < N) ? __ldg(hp + i * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int n0 = 0; n0 < N; n0 += 4) {
            float4 hv4[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) hv4[i] = nxt[i];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = n0 + 4 + i;
              if (n < N) nxt[i] = __ldg(hp + n * (E / 4));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = n0 + i;
              if (n < N) {
                const float4 hv = hv4[i];
                const float4 p0 = *reinterpret_cast<const float4*>(slot + n * 8);
                const float4 p1 = *reinterpret_cast<const float4*>(slot + n * 8 + 4);
                const float pv[NH] = {p0.x, p0.y, p0.z, p0.w, p1.x, p1.y, p1.z, p1.w};
#pragma unroll
                for (int hh = 0; hh < NH; ++hh) {
                  c[hh].x = fmaf(pv[hh], hv.x, c[hh].x);
                  c[hh].y = fmaf(pv[hh], hv.y, c[hh].y);
                  c[hh].z = fmaf(pv[hh], hv.z, c[hh].z);
                  c[hh].w = fmaf(pv[hh], hv.w, c[hh].w);
                }
              }
            }
          }
        }
        __syncwarp();
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) *reinterpret_cast<float4*>(slot + hh * E + lane * 4) = c[hh];
      }
      // rows >= cnt of C must be finite for GEMM-B (results unused): zero them
      for (int o = cnt * QW + tid; o < TM * QW; o += NT) QC[o] = 0.f;
      __syncthreads();

      // ---------------- P3: q^ = C · M^T + m_c  -> Xs
      gemm_b(QC, Xs, Wb, p);

      // ---------------- P4: logits, action, environment transition
      bool unfinished = false;
      for (int m = warp; m < cnt; m += NT / 32) {
        const int64_t b = base + m;
        float* slot = QC + m * QW;  // free scratch again
        const float4 qh = *reinterpret_cast<const float4*>(Xs + m * E + lane * 4);
        const float4* hp = reinterpret_cast<const float4*>(h + b * N * E) + lane;
        {
          float4 nxt[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) nxt[i] = (i < N) ? __ldg(hp + i * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          for (int n0 = 0; n0 < N; n0 += 8) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 hv = nxt[i];
              v[i] = fmaf(qh.x, hv.x, fmaf(qh.y, hv.y, fmaf(qh.z, hv.z, qh.w * hv.w)));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int n = n0 + 8 + i;
              nxt[i] = (n < N) ? __ldg(hp + n * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            float sc = reduce8(v, lane);
            int n = n0 + ((lane >> 2) & 7);
            if ((lane & 3) == 0 && n < N) slot[n] = 10.0f * tanhf(sc);
          }
        }
        __syncwarp();
        // own mask (graph_decoder.py:98), 4 consecutive nodes per lane
        uint32_t mw[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) mw[i] = __ldcg(p.env.mask + b * 4 + i);
        float u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int n = lane * 4 + i;
          bool ok = n < N && !((mw[n >> 5] >> (n & 31)) & 1u);
          u[i] = ok ? slot[n] : -INFINITY;
        }
        if (p.logits) {
          float* lo = p.logits + ((int64_t)trel * B + b) * N;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            if (lane * 4 + i < N) lo[lane * 4 + i] = u[i];
        }
        // max + first-max index (argmax tie rule: lowest index)
        float mx = u[0];
        int am = lane * 4;
#pragma unroll
        for (int i = 1; i < 4; ++i)
          if (u[i] > mx) { mx = u[i]; am = lane * 4 + i; }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          float omx = __shfl_xor_sync(0xffffffffu, mx, o);
          int oam = __shfl_xor_sync(0xffffffffu, am, o);
          if (omx > mx || (omx == mx && oam < am)) { mx = omx; am = oam; }
        }
        int a = am;
        float lp = 0.f;
        if (p.mode != VRPX_GREEDY) {
          float ex[4], loc = 0.f;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            ex[i] = (u[i] == -INFINITY) ? 0.f : expf(u[i] - mx);
            loc += ex[i];
          }
          float incl = loc;  // inclusive scan of lane totals
#pragma unroll
          for (int o = 1; o < 32; o <<= 1) {
            float y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
          }
          float total = __shfl_sync(0xffffffffu, incl, 31);
          if (p.mode == VRPX_SAMPLE) {
            uint64_t gid = p.offset + (uint64_t)b;
            uint4 r = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t, 0x5eedu),
                                    make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
            float thr = u24(r.x) * total;
            float cum = incl - loc;
            int pick = 1 << 30;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              cum += ex[i];
              if (ex[i] > 0.f && cum > thr && pick == (1 << 30)) pick = lane * 4 + i;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) pick = min(pick, __shfl_xor_sync(0xffffffffu, pick, o));
            a = (pick == (1 << 30)) ? am : pick;  // rounding guard: fall back to the mode
          } else {
            a = (int)p.tape[(int64_t)trel * B + b];
          }
          // log-prob of the taken action (graph_decoder.py:107)
          const int ai = a & 3;
          float usel = ai == 0 ? u[0] : (ai == 1 ? u[1] : (ai == 2 ? u[2] : u[3]));
          float ua = __shfl_sync(0xffffffffu, usel, a >> 2);
          lp = (ua - mx) - logf(total);
        }
        if (lane == 0) {
          if (p.tape && p.mode != VRPX_TEACHER) p.tape[(int64_t)trel * B + b] = (uint8_t)a;
          Bits128 v;
#pragma unroll
          for (int i = 0; i < 4; ++i) v.w[i] = p.env.visited[b * 4 + i];
          int cur = p.env.cur[b];
          double load = p.env.load[b];
          const int depot = p.env.depot[b];
          const double* dem = p.env.demand ? p.env.demand + b * N : nullptr;
          StepResult r = env_transition(kind, N, p.env.xy + b * N * 2, dem, depot, a, v, cur, load);
          p.env.cur[b] = cur;
          p.env.load[b] = load;
#pragma unroll
          for (int i = 0; i < 4; ++i) p.env.visited[b * 4 + i] = v.w[i];
          if (kind == VRPX_IRP) {
            Bits128 x = demand_exceeds(dem, N, load);
#pragma unroll
            for (int i = 0; i < 4; ++i) p.env.mask[b * 4 + i] = v.w[i] | x.w[i];
          }
          // f32 accumulation of f32(reward) in step order (graph_tsp_agent.py:85); cost = -acc_loss
          p.cost[b] = (t == 0 ? 0.f : p.cost[b]) + (float)r.dist;
          if (p.mode != VRPX_GREEDY) p.logp[b] = (t == 0 ? 0.f : p.logp[b]) + lp;
          else if (t == 0) p.logp[b] = 0.f;
          if (!r.all_before) unfinished = true;
        }
      }
      if (unfinished) s_anyleft = 1;  // benign race: all writers store 1
      __syncthreads();
      if (s_anyleft) cta_unfinished = true;
      __syncthreads();
    }
    if (tid == 0 && cta_unfinished) atomicAdd(p.notdone + trel, 1);
    bar_target += gridDim.x;
    grid_barrier(p.bar, bar_target);
    if (ld_acquire_i(p.notdone + trel) == 0) { ++t; break; }
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps = t;
}

constexpr int64_t kRolloutSmall = 4096;  // barrier counter + notdone[<=513]

}  // namespace vrpx

using namespace vrpx;

extern "C" {

int64_t vrpx_rollout_workspace_bytes(int64_t B, int32_t N) {
  (void)N;
  return kRolloutSmall + B * (int64_t)QW * (int64_t)sizeof(float);
}

int vrpx_rollout(const vrpx_env* env, const vrpx_decoder_weights* w, const float* h, int32_t mode,
                 int64_t coupling, uint64_t seed, uint64_t offset, uint8_t* tape, int32_t t_begin, int32_t Tmax,
                 float* logp, float* cost, int32_t* steps, float* logits, void* ws, int64_t ws_bytes,
                 void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VRPX_CHECK_ARG(env && w && h && logp && cost && steps && ws, "NULL argument");
  VRPX_CHECK_ARG(env->kind >= 0 && env->kind <= 2 && env->N >= 2 && env->N <= VRPX_MAX_NODES && env->B >= 1,
                 "bad env header");
  VRPX_CHECK_ARG(env->xy && env->depot && env->visited && env->mask && env->cur && env->load, "env arrays");
  VRPX_CHECK_ARG(env->kind != VRPX_IRP || env->demand, "IRP env needs demand");
  VRPX_CHECK_ARG(mode >= 0 && mode <= 2, "bad mode");
  VRPX_CHECK_ARG(mode != VRPX_TEACHER || tape, "teacher mode needs a tape");
  VRPX_CHECK_ARG(Tmax >= 1 && Tmax <= 1000 && t_begin >= 0, "t_begin / Tmax out of range");
  VRPX_CHECK_ARG(coupling >= 0, "coupling must be >= 0");
  VRPX_CHECK_ARG(ws_bytes >= vrpx_rollout_workspace_bytes(env->B, env->N), "workspace too small");
  VRPX_CHECK_ARG(w->ag_t && w->al_t && w->a_c && w->a_q0 && w->m_t && w->m_c, "decoder weights");
  VRPX_CHECK_ARG(env->kind == VRPX_IRP ? (w->a_load != nullptr) : (w->af_t != nullptr), "decoder weights (kind)");

  RolloutParams p;
  p.env = *env;
  p.w = *w;
  p.h = h;
  p.mode = mode;
  p.G = coupling;
  p.seed = seed;
  p.offset = offset;
  p.tape = tape;
  p.t0 = t_begin;
  p.Tmax = Tmax;
  p.logp = logp;
  p.cost = cost;
  p.steps = steps;
  p.logits = logits;
  p.bar = reinterpret_cast<unsigned*>(ws);
  p.notdone = reinterpret_cast<int*>(ws) + 8;
  p.qg = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kRolloutSmall);
  VRPX_CHECK_ARG((int64_t)(Tmax + 1 + 8) * 4 <= kRolloutSmall, "Tmax too large for workspace header");

  VRPX_CUDA(cudaMemsetAsync(ws, 0, kRolloutSmall, stream));
  VRPX_CUDA(cudaFuncSetAttribute(k_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_TOTAL));
  int64_t ntiles = (env->B + TM - 1) / TM;
  int grid = (int)((ntiles < (int64_t)num_sms()) ? ntiles : (int64_t)num_sms());
  void* args[] = {(void*)&p};
  VRPX_CUDA(cudaLaunchCooperativeKernel((void*)k_rollout, dim3(grid), dim3(NT), args, SMEM_TOTAL, stream));
  count_launch();
  return VRPX_OK;
}

}  // extern "C"

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
#include "tile_gemm.cuh"

namespace vrpx {

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
  float* qg0;         // optional copy of Q~g before the `first` fold (backward)
  uint32_t* mask_hist;  // optional [Tmax][B][4] decoder-visible mask before each step (backward)
  float* load_hist;   // optional [Tmax][B] f32 vehicle load before each step (backward)
  unsigned* bar;      // grid barrier counter
  int* notdone;       // [Tmax + 1]
};

constexpr size_t SMEM_TOTAL = SMEM_X + SMEM_QC + SMEM_W;

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
    tile_gemm_wide(Xs, p.w.ag_t, Wb, [&](int m, int c, float4 v) {   // Q~g = A_g · g + a_c
      if (m >= cnt) return;
      const float4 ac = *reinterpret_cast<const float4*>(p.w.a_c + c);
      v = make_float4(v.x + ac.x, v.y + ac.y, v.z + ac.z, v.w + ac.w);
      *reinterpret_cast<float4*>(p.qg + (base + m) * QW + c) = v;
      if (p.qg0) *reinterpret_cast<float4*>(p.qg0 + (base + m) * QW + c) = v;
    });
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
        const float lf = (m < cnt) ? (float)p.env.load[base + m] : 0.f;
        if (lane == 0) s_loadf[m] = lf;
        if (m < cnt && p.mask_hist && lane < 4)
          p.mask_hist[((int64_t)trel * B + base + m) * 4 + lane] = __ldcg(p.env.mask + (base + m) * 4 + lane);
        if (m < cnt && p.load_hist && lane == 0) p.load_hist[(int64_t)trel * B + base + m] = lf;
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
          tile_gemm_wide(Xs, p.w.af_t, Wb, [&](int m, int c, float4 v) {   // fold `first` (graph_decoder.py:111-113)
            if (m >= cnt) return;
            float4* qgp = reinterpret_cast<float4*>(p.qg + (base + m) * QW + c);
            const float4 q = *qgp;
            *qgp = make_float4(q.x + v.x, q.y + v.y, q.z + v.z, q.w + v.w);
          });
          __syncthreads();  // qg updates are re-read by other threads' epilogue below? (same thread) — keep ordering explicit
        }
        tile_gemm_wide(Xs, p.w.al_t, Wb, [&](int m, int c, float4 v) {   // q~ = A_l · h[last] + Q~g (+ load · a_load)
          if (m >= cnt) return;
          const float4 q = *reinterpret_cast<const float4*>(p.qg + (base + m) * QW + c);
          v = make_float4(v.x + q.x, v.y + q.y, v.z + q.z, v.w + q.w);
          if (kind == VRPX_IRP) {
            const float4 al = *reinterpret_cast<const float4*>(p.w.a_load + c);
            const float lf = s_loadf[m];
            v = make_float4(fmaf(lf, al.x, v.x), fmaf(lf, al.y, v.y), fmaf(lf, al.z, v.z), fmaf(lf, al.w, v.w));
          }
          *reinterpret_cast<float4*>(QC + m * QW + c) = v;
        });
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
      tile_gemm_tall(QC, p.w.m_t, Wb, p.w.m_c, QC, Xs);

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
                 float* logp, float* cost, int32_t* steps, float* logits, const vrpx_rollout_trace* trace,
                 void* ws, int64_t ws_bytes, void* stream_) {
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
  p.qg0 = trace ? trace->qg0 : nullptr;
  p.mask_hist = trace ? trace->mask_hist : nullptr;
  p.load_hist = trace ? trace->load_hist : nullptr;
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

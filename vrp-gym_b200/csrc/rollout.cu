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
#include "rollout.cuh"

namespace vrpx {


int64_t score_table_slice(int64_t B);
int build_score_table(const float* h, const float* qk_w, int64_t B, int N, float* qk_buf, float* s1, cudaStream_t stream);
int build_score_table_fused(const float* h, const float* qk_w, int64_t B, int N, __half* w16, float* s1, cudaStream_t stream);

// Rollout tile: RMT x 16 instances.  With 16 instances per tile the working set that has to survive in L2 between the
// glimpse passes and the pointer-logit pass (all CTAs x tile x 25.6 KB) halves to 60 MB.
constexpr int RMT = 1;
constexpr int RTM = 16 * RMT;
constexpr int RQ_LD = 2 * C16_LD;   // floats per instance slot of QC: the fp16-split c rows of GEMM-B set the stride (1032)
constexpr size_t SMEM_X_MMA = smem_x_mma<RMT>(), SMEM_QC_MMA = (size_t)RTM * RQ_LD * sizeof(float);
constexpr int RNST = 2;   // depth of the per-warp weight rings of the tile GEMMs (3 measured slower: the extra 32 KiB of
                          // shared memory comes out of the L1 that buffers the embedding streams of the other phases)
constexpr size_t SMEM_W_RING = (size_t)(NT / 32) * RNST * 512 * sizeof(float);   // 64 KiB
constexpr size_t SMEM_TOTAL = SMEM_X_MMA + SMEM_QC_MMA + SMEM_W_RING;

// m_t [1024][128] f32 -> M16 [512 k-pairs][128] {hi2, lo2}: the B operand of the fp16-split GEMM-B (tile_gemm.cuh)
__global__ void k_split_m16(const float* __restrict__ m_t, uint2* __restrict__ m16) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // kp * 128 + n
  if (i >= (QW / 2) * E) return;
  const int kp = i >> 7, n = i & (E - 1);
  m16[i] = split_f16x2(m_t[(2 * kp) * E + n], m_t[(2 * kp + 1) * E + n]);
}

// ---------------------------------------------------------------- the persistent kernel
__global__ void __launch_bounds__(NT, 1) k_rollout(const RolloutParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* Xs = reinterpret_cast<float*>(smem_raw);
  float* QC = reinterpret_cast<float*>(smem_raw + SMEM_X_MMA);
  float* Wb = reinterpret_cast<float*>(smem_raw + SMEM_X_MMA + SMEM_QC_MMA);
  float* TB = reinterpret_cast<float*>(smem_raw + SMEM_TOTAL);   // [RTM][tb_segs][8 N] staged table rows (table mode)
  __shared__ float s_loadf[RTM], s_lp[RTM];
  __shared__ int s_anyleft, s_act[RTM];

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.env.N, kind = p.env.kind;
  const int64_t B = p.env.B;
  const int64_t ntiles = (B + RTM - 1) / RTM;
  const float* __restrict__ h = p.h;
  unsigned bar_target = 0;
  long long prof_t = clock64(), prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#define VRPX_PROF(i)                                   \
  if (p.prof && tid == 0) {                            \
    long long now_ = clock64();                        \
    prof_acc[i] += now_ - prof_t;                      \
    prof_t = now_;                                     \
  }

  // ------------------------------------------------ prologue: Q~g[b] = A_g · mean_n h[b,n] + a_c
  for (int64_t tile = blockIdx.x; tile < ntiles && p.t0 == 0; tile += gridDim.x) {
    const int64_t base = tile * RTM;
    const int cnt = (int)((B - base < RTM) ? (B - base) : RTM);
    for (int m = warp; m < RTM; m += NT / 32) {
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
      *reinterpret_cast<float4*>(Xs + m * XS_LD + lane * 4) = g;
    }
    __syncthreads();
    tile_gemm_wide_mma_sw<RMT, RNST>(                                              // Q~g = A_g · g + a_c
        Xs, XS_LD, p.w.ag_t, Wb, [&](int, int c) { return *reinterpret_cast<const float2*>(p.w.a_c + c); },
        [&](int m, int c, float v0, float v1) {
          if (m >= cnt) return;
          const float2 v = make_float2(v0, v1);
          *reinterpret_cast<float2*>(p.qg + (base + m) * QW + c) = v;
          if (p.qg0) *reinterpret_cast<float2*>(p.qg0 + (base + m) * QW + c) = v;
        });
    __syncthreads();
  }

  // Snapshot of the decoder-visible masks for the glimpses of the first step (see RolloutParams::gmask), then a grid
  // barrier: no instance may step before every CTA has taken its snapshot and finished its Q~g rows.
  {
    uint32_t* gm = p.gmask + (size_t)(p.t0 & 1) * B * 4;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t base = tile * RTM;
      const int cnt = (int)((B - base < RTM) ? (B - base) : RTM);
      if (tid < cnt * 4) gm[base * 4 + tid] = __ldcg(p.env.mask + base * 4 + tid);
    }
    bar_target += gridDim.x;
    grid_barrier(p.bar, bar_target);
  }

  // Table mode: the rows S1[b][last], S0[b] (, SL[b]) the NEXT tile will need are copied to shared memory with cp.async
  // while the current tile runs its pointer-logit phase, so the dependent chain cur -> table row -> softmax does not
  // sit exposed at the head of every tile.  One warp per instance, issued and consumed by the same warp.
  const int tb_row = NH * N;                    // floats per staged segment
  const int tb_stride = p.tb_segs * tb_row;     // floats per instance
  bool staged = false;                          // TB holds the rows of the tile about to be processed
  auto stage_tables = [&](int64_t ntile) {
    const int64_t nbase = ntile * RTM;
    const int ncnt = (int)((B - nbase < RTM) ? (B - nbase) : RTM);
    for (int m = warp; m < ncnt; m += NT / 32) {
      const int64_t b = nbase + m;
      const int last = __ldcg(p.env.cur + b);
      float* dst = TB + m * tb_stride;
      const float* src1 = p.s1 + (((size_t)b * N + last) * NH) * N;
      for (int i = lane; i < tb_row / 4; i += 32) cp_async16(dst + 4 * i, src1 + 4 * i);
      if (p.tb_segs >= 2) {
        const float* src0 = p.s0 + (size_t)b * tb_row;
        for (int i = lane; i < tb_row / 4; i += 32) cp_async16(dst + tb_row + 4 * i, src0 + 4 * i);
      }
      if (p.tb_segs >= 3) {
        const float* srcl = p.sl + (size_t)b * tb_row;
        for (int i = lane; i < tb_row / 4; i += 32) cp_async16(dst + 2 * tb_row + 4 * i, srcl + 4 * i);
      }
    }
    cp_async_commit();
  };

  int t = p.t0;
  for (; t < p.t0 + p.Tmax; ++t) {
    const int trel = t - p.t0;
    bool cta_unfinished = false;
    const uint32_t* __restrict__ gm_rd = p.gmask + (size_t)(t & 1) * B * 4;   // masks BEFORE step t, every instance
    uint32_t* __restrict__ gm_wr = p.gmask + (size_t)((t + 1) & 1) * B * 4;   // masks after step t
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int64_t base = tile * RTM;
      const int cnt = (int)((B - base < RTM) ? (B - base) : RTM);
      if (tid == 0) s_anyleft = 0;
      // table mode: from step 2 on the glimpse scores come from the per-episode tables (no GEMM-A, no score pass)
      const bool tbl_step = p.s1 != nullptr && t >= 2;
      // ---------------- P0: gather last-node embeddings, vehicle load; pull this tile's Q~g rows towards L2
      if (!tbl_step)
        for (int i = tid; i < cnt * 32; i += NT)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(p.qg + base * QW) + (size_t)i * 128));
      for (int m = warp; m < RTM; m += NT / 32) {
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < cnt && t > 0 && !tbl_step) {
          int last = p.env.cur[base + m];
          xv = __ldg(reinterpret_cast<const float4*>(h + ((base + m) * N + last) * E) + lane);
        }
        if (!tbl_step) *reinterpret_cast<float4*>(Xs + m * XS_LD + lane * 4) = xv;
        const float lf = (m < cnt) ? (float)p.env.load[base + m] : 0.f;
        if (lane == 0) s_loadf[m] = lf;
        if (m < cnt && p.mask_hist && lane < 4)
          p.mask_hist[((int64_t)trel * B + base + m) * 4 + lane] = __ldcg(p.env.mask + (base + m) * 4 + lane);
        if (m < cnt && p.load_hist && lane == 0) p.load_hist[(int64_t)trel * B + base + m] = lf;
      }
      __syncthreads();
      VRPX_PROF(0)
      // ---- glimpse score pass on the tensor pipe: S[node][head] = H_b[node][:] · q[head][:]  (m16n8k8, M = 16 nodes,
      // N = 8 heads), q = 8 x 128 floats in shared memory.  Fragment coordinates g = lane >> 2, tq = lane & 3.  The K
      // (embedding) axis is permuted so that every thread streams whole float4 chunks: chunk c (0..7) of thread tq
      // covers dims 16c + 4tq + {0,1,2,3}; k-step 2c + u uses mma k-index tq <-> dim 16c+4tq+2u and k-index tq+4 <->
      // dim 16c+4tq+2u+1, for A (node rows) and B (q) alike.  store(n, which, v): node n, head 2tq + which.
      auto glimpse_scores = [&](const float* q, const float4* hrow, auto&& store) {
        const int g = lane >> 2, tq = lane & 3;
        float4 qv[8];   // q[head g][dims of this thread]
#pragma unroll
        for (int c = 0; c < 8; ++c) qv[c] = *reinterpret_cast<const float4*>(q + g * E + 16 * c + 4 * tq);
        __syncwarp();
        for (int n0 = 0; n0 < N; n0 += 16) {
          const int na = n0 + g, nbb = n0 + g + 8;
          // six independent accumulator chains (3 split terms x even/odd k-step): a single chain would serialise the
          // 96 mma of a node tile on the tensor-pipe latency
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
                split_tf32(ae[2 * u], ah[0], al[0]);       // (row g,   k = tq)
                split_tf32(be[2 * u], ah[1], al[1]);       // (row g+8, k = tq)
                split_tf32(ae[2 * u + 1], ah[2], al[2]);   // (row g,   k = tq+4)
                split_tf32(be[2 * u + 1], ah[3], al[3]);   // (row g+8, k = tq+4)
                split_tf32(qe[2 * u], bh0, bl0);           // (k = tq,   n = g)
                split_tf32(qe[2 * u + 1], bh1, bl1);       // (k = tq+4, n = g)
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
          // C fragment: acc[0], acc[1] = node na, heads 2tq, 2tq+1;  acc[2], acc[3] = node nbb
          if (na < N) { store(na, 0, acc[0]); store(na, 1, acc[1]); }
          if (nbb < N) { store(nbb, 0, acc[2]); store(nbb, 1, acc[3]); }
        }
      };
      // ---------------- P1: q~
      if (tbl_step) {
        // nothing: scores are gathered from S1/S0 in P2
      } else if (t == 0) {
        for (int o = tid; o < cnt * QW; o += NT) {
          int m = o >> 10, c = o & (QW - 1);
          float y = p.qg[(base + m) * QW + c] + p.w.a_q0[c];
          if (kind == VRPX_IRP) y = fmaf(s_loadf[m], p.w.a_load[c], y);
          QC[m * RQ_LD + c] = y;
        }
      } else {
        if (t == 1 && kind != VRPX_IRP) {
          tile_gemm_wide_mma_sw<RMT, RNST>(                                        // fold `first` (graph_decoder.py:111-113)
              Xs, XS_LD, p.w.af_t, Wb,
              [&](int m, int c) {
                return (m < cnt) ? *reinterpret_cast<const float2*>(p.qg + (base + m) * QW + c) : make_float2(0.f, 0.f);
              },
              [&](int m, int c, float v0, float v1) {
                if (m < cnt) *reinterpret_cast<float2*>(p.qg + (base + m) * QW + c) = make_float2(v0, v1);
              });
        }
        if (t == 1 && p.s1) {
          // table mode: Q~g is final now -> S0[b][head][n] = Q~g[b]_head · h[b,n]  (IRP: SL from a_load as well)
          __syncthreads();
          for (int m = warp; m < cnt; m += NT / 32) {
            const int64_t b = base + m;
            float* slot = QC + m * RQ_LD;
            const float4* hrow = reinterpret_cast<const float4*>(h + b * N * E);
            const int tq = lane & 3;
            for (int i = lane; i < QW / 4; i += 32)
              *reinterpret_cast<float4*>(slot + 4 * i) = __ldcg(reinterpret_cast<const float4*>(p.qg + b * QW) + i);
            __syncwarp();
            float* s0 = p.s0 + (size_t)b * NH * N;
            glimpse_scores(slot, hrow, [&](int n, int which, float v) { s0[(2 * tq + which) * N + n] = v; });
            if (kind == VRPX_IRP) {
              __syncwarp();
              for (int i = lane; i < QW / 4; i += 32)
                *reinterpret_cast<float4*>(slot + 4 * i) = __ldg(reinterpret_cast<const float4*>(p.w.a_load) + i);
              __syncwarp();
              float* sl = p.sl + (size_t)b * NH * N;
              glimpse_scores(slot, hrow, [&](int n, int which, float v) { sl[(2 * tq + which) * N + n] = v; });
            }
          }
          __syncthreads();
        }
        tile_gemm_wide_mma_sw<RMT, RNST>(                                          // q~ = Q~g (+ load · a_load) + A_l · h[last]
            Xs, XS_LD, p.w.al_t, Wb,
            [&](int m, int c) {
              if (m >= cnt) return make_float2(0.f, 0.f);
              float2 q = *reinterpret_cast<const float2*>(p.qg + (base + m) * QW + c);
              if (kind == VRPX_IRP) {
                const float2 al = *reinterpret_cast<const float2*>(p.w.a_load + c);
                const float lf = s_loadf[m];
                q = make_float2(fmaf(lf, al.x, q.x), fmaf(lf, al.y, q.y));
              }
              return q;
            },
            [&](int m, int c, float v0, float v1) { *reinterpret_cast<float2*>(QC + m * RQ_LD + c) = make_float2(v0, v1); });
      }
      __syncthreads();
      VRPX_PROF(1)

      // ---------------- P2: glimpse attention, one warp per instance
      for (int m = warp; m < cnt; m += NT / 32) {
        const int64_t b = base + m;
        float* slot = QC + m * RQ_LD;
        const int g = lane >> 2, t = lane & 3;
        const float4* hrow = reinterpret_cast<const float4*>(h + b * N * E);
        float pr[NH][4];
        if (!tbl_step) {
          // ---- pass 1: scores = q~_h · h_n + scrambled additive mask
          // masks of the instances whose rows the reference adds to heads 2t and 2t+1 (mask.repeat(H,1), graph_decoder.py:93)
          uint32_t nb0[4], nb1[4];
          {
            const uint32_t* m0 = gm_rd + quirk_row(b, 2 * t, p.G) * 4;
            const uint32_t* m1 = gm_rd + quirk_row(b, 2 * t + 1, p.G) * 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) { nb0[i] = __ldcg(m0 + i); nb1[i] = __ldcg(m1 + i); }
          }
          glimpse_scores(slot, hrow, [&](int n, int which, float v) {
            const uint32_t wsel = which ? nb1[n >> 5] : nb0[n >> 5];
            slot[(2 * t + which) * E + n] = v + (float)((wsel >> (n & 31)) & 1u);
          });
          __syncwarp();
          VRPX_PROF(6)
#pragma unroll
          for (int hh = 0; hh < NH; ++hh)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = lane + 32 * i;
              pr[hh][i] = (n < N) ? slot[hh * E + n] : -INFINITY;
            }
        } else {
          // ---- table mode: scores = S1[b][last] + S0[b] (+ load · SL[b]) + scrambled additive mask
          const float* r1;
          const float* r0 = p.s0 + (size_t)b * NH * N;
          const float* rl = p.sl + (size_t)b * NH * N;
          if (staged) {
            cp_async_wait<0>();
            __syncwarp();
            r1 = TB + m * tb_stride;
            if (p.tb_segs >= 2) r0 = r1 + tb_row;
            if (p.tb_segs >= 3) rl = r1 + 2 * tb_row;
          } else {
            const int last = __ldcg(p.env.cur + b);
            r1 = p.s1 + (((size_t)b * N + last) * NH) * N;
          }
          const float lf = s_loadf[m];
          // lane j holds mask word (j & 3) of the instance whose mask the reference adds to head j >> 2
          const uint32_t mword = __ldcg(gm_rd + quirk_row(b, lane >> 2, p.G) * 4 + (lane & 3));
#pragma unroll
          for (int hh = 0; hh < NH; ++hh)
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const int n = lane + 32 * i;
              const uint32_t wsel = __shfl_sync(0xffffffffu, mword, hh * 4 + i);
              float v = -INFINITY;
              if (n < N) {
                v = r1[hh * N + n] + r0[hh * N + n];   // generic loads: shared (staged) or global (L2-coherent data)
                if (kind == VRPX_IRP) v = fmaf(lf, rl[hh * N + n], v);
                v += (float)((wsel >> lane) & 1u);
              }
              pr[hh][i] = v;
            }
          VRPX_PROF(6)
        }
        // softmax per head over nodes (lane = node, 4 strides cover N <= 128)
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) {
          float mx = -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) mx = fmaxf(mx, pr[hh][i]);
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
        VRPX_PROF(7)
        // ---- pass 2 on the tensor pipe: c[head][dim] = sum_n P[n][head] h_n[dim]  (M = 16 dims, N = 8 heads, K = 8 nodes)
        // Thread g streams the float4 chunks 8c' + g (dims 32c' + 4g + e) of node rows n0+t and n0+t+4; m-tile
        // j = 2c' + u has row g <-> dim 32c'+4g+2u and row g+8 <-> dim 32c'+4g+2u+1.  B = P[n][head] from the slot.
        float cacc[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j)
#pragma unroll
          for (int i = 0; i < 4; ++i) cacc[j][i] = 0.f;
        for (int n0 = 0; n0 < N; n0 += 8) {
          const int na = n0 + t, nbb = n0 + t + 4;
          float4 va[4], vb[4];
#pragma unroll
          for (int cq = 0; cq < 4; ++cq) {
            va[cq] = (na < N) ? __ldg(hrow + na * (E / 4) + 8 * cq + g) : make_float4(0.f, 0.f, 0.f, 0.f);
            vb[cq] = (nbb < N) ? __ldg(hrow + nbb * (E / 4) + 8 * cq + g) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
          uint32_t bh0, bl0, bh1, bl1;
          split_tf32((na < N) ? slot[na * 8 + g] : 0.f, bh0, bl0);      // (k = t,   n = head g)
          split_tf32((nbb < N) ? slot[nbb * 8 + g] : 0.f, bh1, bl1);    // (k = t+4, n = head g)
#pragma unroll
          for (int cq = 0; cq < 4; ++cq) {
            const float ae[4] = {va[cq].x, va[cq].y, va[cq].z, va[cq].w};
            const float be[4] = {vb[cq].x, vb[cq].y, vb[cq].z, vb[cq].w};
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              uint32_t ah[4], al[4];
              split_tf32(ae[2 * u], ah[0], al[0]);       // (row g   = dim 32cq+4g+2u,   k = t   = node na)
              split_tf32(ae[2 * u + 1], ah[1], al[1]);   // (row g+8 = dim 32cq+4g+2u+1, k = t)
              split_tf32(be[2 * u], ah[2], al[2]);       // (row g,   k = t+4 = node nbb)
              split_tf32(be[2 * u + 1], ah[3], al[3]);   // (row g+8, k = t+4)
              mma_tf32_16x8x8(cacc[2 * cq + u], al, bh0, bh1);
              mma_tf32_16x8x8(cacc[2 * cq + u], ah, bl0, bl1);
              mma_tf32_16x8x8(cacc[2 * cq + u], ah, bh0, bh1);
            }
          }
        }
        __syncwarp();  // every lane is done reading P before c overwrites the slot
        // C fragment of m-tile j = 2cq+u: [0] (dim d, head 2t), [1] (dim d, head 2t+1), [2] (dim d+1, head 2t),
        // [3] (dim d+1, head 2t+1) with d = 32cq + 4g + 2u  ->  the A operand of GEMM-B, pre-split for the fp16 tensor
        // path (tile_gemm.cuh): row m of C16, k = head * 128 + dim; per head and cq one 16-byte chunk = the two k-pairs
        // (d0, d0+1), (d0+2, d0+3), d0 = 32cq + 4g, at the swizzled chunk (head*32 + 8cq + g) ^ (t << 1)
        uint4* crow = reinterpret_cast<uint4*>(slot);
#pragma unroll
        for (int cq = 0; cq < 4; ++cq)
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const uint2 p01 = split_f16x2(cacc[2 * cq][e], cacc[2 * cq][e + 2]);           // dims d0, d0+1 of head 2t+e
            const uint2 p23 = split_f16x2(cacc[2 * cq + 1][e], cacc[2 * cq + 1][e + 2]);   // dims d0+2, d0+3
            crow[(((2 * t + e) * 32 + 8 * cq + g) ^ (t << 1))] = make_uint4(p01.x, p01.y, p23.x, p23.y);
          }
      }
      // rows >= cnt of C must be finite for GEMM-B (results unused): zero them
      for (int o = cnt * QW + tid; o < RTM * QW; o += NT) QC[(o >> 10) * RQ_LD + (o & (QW - 1))] = 0.f;
      __syncthreads();
      VRPX_PROF(2)

      // ---------------- P3: q^ = C · M^T + m_c  -> Xs
      tile_gemm_tall_f16<RNST>(reinterpret_cast<const uint2*>(QC), p.m16, Wb, p.w.m_c, QC, Xs, XS_LD);

      // table mode: start copying the next tile's table rows (its `cur` is final: written one step ago, or — when the
      // next tile is this CTA's first tile of the NEXT step — earlier in this step)
      staged = false;
      if (p.s1 && p.tb_segs > 0) {
        int64_t ntile = tile + gridDim.x;
        int tn = t;
        if (ntile >= ntiles) { ntile = blockIdx.x; tn = t + 1; }
        if (tn >= 2 && ntile != tile) {
          stage_tables(ntile);
          staged = true;
        }
      }
      VRPX_PROF(3)
      // ---------------- P4: logits, action, environment transition
      bool unfinished = false;
      for (int m = warp; m < cnt; m += NT / 32) {
        const int64_t b = base + m;
        float* slot = QC + m * RQ_LD;  // free scratch again
        const float4 qh = *reinterpret_cast<const float4*>(Xs + m * XS_LD + lane * 4);
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
        if (lane == 0) { s_act[m] = a; s_lp[m] = lp; }
      }
      __syncthreads();
      // environment transition: one THREAD per instance, so the dependent global loads (state -> coordinates) of the
      // whole tile overlap instead of running back to back on lane 0 of each warp
      if (tid < cnt) {
        const int m = tid;
        const int64_t b = base + m;
        const int a = s_act[m];
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
        Bits128 dm = v;   // decoder-visible mask after this step
        if (kind == VRPX_IRP) {
          Bits128 x = demand_exceeds(dem, N, load);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            dm.w[i] = v.w[i] | x.w[i];
            p.env.mask[b * 4 + i] = dm.w[i];
          }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) gm_wr[b * 4 + i] = dm.w[i];
        // f32 accumulation of f32(reward) in step order (graph_tsp_agent.py:85); cost = -acc_loss
        p.cost[b] = (t == 0 ? 0.f : p.cost[b]) + (float)r.dist;
        if (p.mode != VRPX_GREEDY) p.logp[b] = (t == 0 ? 0.f : p.logp[b]) + s_lp[m];
        else if (t == 0) p.logp[b] = 0.f;
        if (!r.all_before) unfinished = true;
      }
      if (unfinished) s_anyleft = 1;  // benign race: all writers store 1
      __syncthreads();
      if (s_anyleft) cta_unfinished = true;
      __syncthreads();
      VRPX_PROF(4)
    }
    if (tid == 0 && cta_unfinished) atomicAdd(p.notdone + trel, 1);
    bar_target += gridDim.x;
    grid_barrier(p.bar, bar_target);
    VRPX_PROF(5)
    if (ld_acquire_i(p.notdone + trel) == 0) { ++t; break; }
  }
  cp_async_wait<0>();
  if (p.prof && tid == 0)
    for (int i = 0; i < 8; ++i) atomicAdd(reinterpret_cast<unsigned long long*>(p.prof) + i, (unsigned long long)prof_acc[i]);
  if (blockIdx.x == 0 && tid == 0) *p.steps = t;
}

static long long* g_rollout_prof = nullptr;
static int g_split_steps = 1;                        // vrpx_debug_rollout_split (bit 0)
static int g_fused_tables = 1;                       // vrpx_debug_rollout_split (bit 1 clear = fused score-table kernel)
static int g_time_kernel = 0;                        // vrpx_debug_rollout_timing
static cudaEvent_t g_ev0 = nullptr, g_ev1 = nullptr;
static bool g_ev_valid = false;
constexpr int64_t kRolloutSmall = 4096;  // barrier counter + notdone[<=513]
constexpr int64_t kRolloutM16 = (int64_t)(QW / 2) * E * sizeof(uint2);   // 512 KiB: pre-split m_t
constexpr int64_t kRolloutHdr = kRolloutSmall + kRolloutM16;             // Q~g follows

}  // namespace vrpx

using namespace vrpx;

extern "C" {

/* Debug hook: device buffer of 8 x int64 that accumulates per-phase cycles of thread 0 of every CTA (NULL disables). */
void vrpx_debug_rollout_profile(long long* dev_counters) { g_rollout_prof = dev_counters; }

void vrpx_debug_rollout_split(int32_t enable) {
  g_split_steps = enable & 1;
  g_fused_tables = (enable & 2) ? 0 : 1;   // bit 1: build the score table with the two-kernel form (A/B measurements)
}

void vrpx_debug_rollout_timing(int32_t enable) {
  g_time_kernel = enable;
  g_ev_valid = false;
}

float vrpx_debug_rollout_kernel_ms(void) {
  if (!g_ev_valid) return -1.0f;
  float ms = -1.0f;
  if (cudaEventSynchronize(g_ev1) != cudaSuccess || cudaEventElapsedTime(&ms, g_ev0, g_ev1) != cudaSuccess) return -1.0f;
  return ms;
}

// header | Q~g [B][1024] f32 | gmask [2][B][4] u32
static inline int64_t rollout_gmask_offset(int64_t B) { return kRolloutHdr + B * (int64_t)QW * (int64_t)sizeof(float); }
int64_t vrpx_rollout_workspace_bytes(int64_t B, int32_t N) {
  (void)N;
  return rollout_gmask_offset(B) + 2 * B * 4 * (int64_t)sizeof(uint32_t);
}

int64_t vrpx_rollout_workspace_qg_offset(void) { return kRolloutHdr; }

// table-mode workspace: header | Q~g | gmask | S0 | SL (IRP) | S1 | QK slice | c | q^ | m_t^T, every segment 256-byte aligned
namespace {
struct TableLayout {
  int64_t s0, sl, s1, qk, cbuf, qhat, mnt, agn, afn, w16b, qkw16, total;
};
inline int64_t align256(int64_t x) { return (x + 255) & ~(int64_t)255; }
TableLayout table_layout(int kind, int64_t B, int N) {
  TableLayout L;
  const int64_t f = (int64_t)sizeof(float);
  L.s0 = align256(vrpx_rollout_workspace_bytes(B, N));
  L.sl = align256(L.s0 + B * NH * N * f);
  L.s1 = (kind == VRPX_IRP) ? align256(L.sl + B * NH * N * f) : L.sl;
  L.qk = align256(L.s1 + B * N * NH * N * f);
  // split-step mode (rollout_steps.cu): glimpse vectors, folded queries, transposed m_t
  L.cbuf = align256(L.qk + score_table_slice(B) * N * 768 * f);
  L.qhat = align256(L.cbuf + B * QW * f);
  L.mnt = align256(L.qhat + B * E * f);
  L.agn = align256(L.mnt + (int64_t)QW * E * f);
  L.afn = align256(L.agn + (int64_t)QW * E * f);
  L.w16b = align256(L.afn + (int64_t)QW * E * f);
  L.qkw16 = align256(L.w16b + (int64_t)2 * QW * E * 2);
  L.total = align256(L.qkw16 + (int64_t)2 * 768 * E * 2);
  return L;
}
}  // namespace

int64_t vrpx_rollout_table_workspace_bytes(int32_t kind, int64_t B, int32_t N) { return table_layout(kind, B, N).total; }

int vrpx_rollout(const vrpx_env* env, const vrpx_decoder_weights* w, const float* h, int32_t mode,
                 int64_t coupling, uint64_t seed, uint64_t offset, uint8_t* tape, int32_t t_begin, int32_t Tmax,
                 float* logp, float* cost, int32_t* steps, float* logits, const vrpx_rollout_trace* trace,
                 void* ws, int64_t ws_bytes, void* stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  VRPX_CHECK_ARG(env && w && h && logp && cost && steps && ws, "NULL argument");
  VRPX_DEVICE_GUARD(h);
  NvtxRange nvtx_range("vrpx:rollout");
  VRPX_CHECK_ARG(env->kind >= 0 && env->kind <= 2 && env->N >= 2 && env->N <= VRPX_MAX_NODES && env->B >= 1,
                 "bad env header");
  VRPX_CHECK_ARG(env->xy && env->depot && env->visited && env->mask && env->cur && env->load, "env arrays");
  VRPX_CHECK_ARG(env->kind != VRPX_IRP || env->demand, "IRP env needs demand");
  VRPX_CHECK_ARG(mode >= 0 && mode <= 2, "bad mode");
  VRPX_CHECK_ARG(mode != VRPX_TEACHER || tape, "teacher mode needs a tape");
  VRPX_CHECK_ARG(Tmax >= 1 && Tmax <= 1000 && t_begin >= 0, "t_begin / Tmax out of range");
  // attention row (b, head) reads the mask of instance quirk_row(b, head, G): it must stay inside the batch
  VRPX_CHECK_ARG(coupling >= 0 && (coupling == 0 || (coupling <= env->B && env->B % coupling == 0)),
                 "coupling group must be 0 or a divisor of the batch size");
  VRPX_CHECK_ARG(ws_bytes >= vrpx_rollout_workspace_bytes(env->B, env->N), "workspace too small");
  VRPX_CHECK_ARG((reinterpret_cast<uintptr_t>(ws) & 15) == 0, "workspace must be 16-byte aligned");
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
  p.prof = g_rollout_prof;
  p.bar = reinterpret_cast<unsigned*>(ws);
  p.notdone = reinterpret_cast<int*>(ws) + 8;
  p.qg = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + kRolloutHdr);
  p.gmask = reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(ws) + rollout_gmask_offset(env->B));
  p.m16 = reinterpret_cast<const uint2*>(reinterpret_cast<char*>(ws) + kRolloutSmall);
  p.s1 = nullptr;
  p.s0 = p.sl = nullptr;
  p.cbuf = p.qhat = nullptr;
  SplitWorkspace sw{nullptr, nullptr, nullptr, nullptr};
  // table mode: whole-episode call with the large workspace and the rank-48 factors
  if (w->qk_w && t_begin == 0 && Tmax >= 3) {
    const TableLayout L = table_layout(env->kind, env->B, env->N);
    if (ws_bytes >= L.total && (reinterpret_cast<uintptr_t>(ws) & 15) == 0) {
      char* base = reinterpret_cast<char*>(ws);
      float* s1 = reinterpret_cast<float*>(base + L.s1);
      int rc;
      {
        NvtxRange nvtx_tables("vrpx:score_tables");
        if (g_fused_tables)   // projection + table in one kernel (score_table_fused.cu); the two-kernel form is the A/B switch
          rc = build_score_table_fused(h, w->qk_w, env->B, env->N, reinterpret_cast<__half*>(base + L.qkw16), s1, stream);
        else
          rc = build_score_table(h, w->qk_w, env->B, env->N, reinterpret_cast<float*>(base + L.qk), s1, stream);
      }
      if (rc) return rc;
      p.s1 = s1;
      p.s0 = reinterpret_cast<float*>(base + L.s0);
      p.sl = reinterpret_cast<float*>(base + L.sl);
      p.cbuf = reinterpret_cast<float*>(base + L.cbuf);
      p.qhat = reinterpret_cast<float*>(base + L.qhat);
      sw.m_nt = reinterpret_cast<float*>(base + L.mnt);
      sw.ag_n = reinterpret_cast<float*>(base + L.agn);
      sw.af_n = reinterpret_cast<float*>(base + L.afn);
      sw.w16b = reinterpret_cast<__half*>(base + L.w16b);
    }
  }
  // Split-step mode (default for whole-episode table-mode calls): every step is a handful of launches over the whole
  // batch (rollout_steps.cu); the persistent kernel below serves the classic mode, resumed single-step calls and the
  // all-persistent A/B switch
  const bool split = p.s1 != nullptr && g_split_steps;
  GemmPlan plan_b;
  if (split) {
    int rc = prepare_split_weights(p, sw, &plan_b, stream);
    if (rc) return rc;
  }
  // shared-memory staging of the table rows: as many segments as fit beside the GEMM buffers
  p.tb_segs = 0;
  size_t smem_total = SMEM_TOTAL;
  if (p.s1) {
    const size_t avail = 227 * 1024 - SMEM_TOTAL - 1024;   // 1 KiB for the static __shared__ arrays
    const size_t seg = (size_t)RTM * NH * env->N * sizeof(float);
    const int want = (env->kind == VRPX_IRP) ? 3 : 2;
    p.tb_segs = (int)((avail / seg < (size_t)want) ? avail / seg : (size_t)want);
    smem_total += p.tb_segs * seg;
  }
  VRPX_CHECK_ARG((int64_t)(Tmax + 1 + 8) * 4 <= kRolloutSmall, "Tmax too large for workspace header");

  VRPX_CUDA(cudaMemsetAsync(ws, 0, kRolloutSmall, stream));
  if (split) {
    if (g_time_kernel) {
      if (!g_ev0) {
        VRPX_CUDA(cudaEventCreate(&g_ev0));
        VRPX_CUDA(cudaEventCreate(&g_ev1));
      }
      VRPX_CUDA(cudaEventRecord(g_ev0, stream));
    }
    NvtxRange nvtx_decode("vrpx:decode_loop");
    int rc = run_first_steps(p, sw, plan_b, stream);
    if (rc) return rc;
    if ((rc = run_split_steps(p, t_begin + 2, plan_b, stream))) return rc;   // also writes the step count
    if (g_time_kernel) {
      VRPX_CUDA(cudaEventRecord(g_ev1, stream));
      g_ev_valid = true;
    }
    return VRPX_OK;
  }
  k_split_m16<<<(QW / 2) * E / 256, 256, 0, stream>>>(w->m_t, reinterpret_cast<uint2*>(reinterpret_cast<char*>(ws) + kRolloutSmall));
  VRPX_LAUNCH_CHECK();
  VRPX_CUDA(cudaFuncSetAttribute(k_rollout, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_total));
  int64_t ntiles = (env->B + RTM - 1) / RTM;
  int grid = (int)((ntiles < (int64_t)num_sms()) ? ntiles : (int64_t)num_sms());
  void* args[] = {(void*)&p};
  if (g_time_kernel) {
    if (!g_ev0) {
      VRPX_CUDA(cudaEventCreate(&g_ev0));
      VRPX_CUDA(cudaEventCreate(&g_ev1));
    }
    VRPX_CUDA(cudaEventRecord(g_ev0, stream));
  }
  NvtxRange nvtx_decode("vrpx:decode_loop");
  VRPX_CUDA(cudaLaunchCooperativeKernel((void*)k_rollout, dim3(grid), dim3(NT), args, smem_total, stream));
  count_launch();
  if (g_time_kernel) {
    VRPX_CUDA(cudaEventRecord(g_ev1, stream));
    g_ev_valid = true;
  }
  return VRPX_OK;
}

}  // extern "C"

// rollout_steps.cu — split-step form of the table-mode decode steps t >= 2 of vrpx_rollout.
//
// The persistent kernel (rollout.cu) is bound by L2 bandwidth: a 16-instance tile re-streams the 512 KiB of folded
// GEMM-B weights from L2 on every SM (32 KB per instance-step) and the embeddings cross L2 twice per step.  Larger
// tiles are not possible inside one CTA (shared memory, and the L2 working set of the embeddings).  Here the three
// phases of a table step become three launches over the WHOLE batch, with the kernel boundary as the grid barrier the
// mask coupling needs:
//   k_step_glimpse   one warp per instance: scores from the per-episode tables, softmax, glimpse values
//                    c_h = sum_n p_hn h_n (mma.sync TF32 3-term split)  ->  c [B][1024] f32
//   gemm_tc          q^ = c · M^T + m_c as ONE batched tcgen05 GEMM (R = B, K = 1024, NOUT = 128): the weights are read
//                    once per 128 instances instead of once per 16
//   k_step_pointer   one warp per instance: logits 10 tanh(q^ · h_n), mask, argmax / Philox sample / teacher action,
//                    log-prob; then one thread per instance runs the environment transition (env_rules.cuh)
// Same arithmetic as phases P2 and P4 of the persistent kernel (the code is carried over); the results differ only by
// the summation order inside GEMM-B.  Steps 0 and 1 (and the classic mode, resumed calls) stay in the persistent kernel.
// The launches are asynchronous on the caller's stream: no host round trip per step.  An episode that finishes early
// (VRP / IRP) turns the remaining launches into no-ops through the per-step `notdone` counters.
#include "gemm.cuh"
#include "rollout.cuh"

namespace vrpx {

constexpr int SW = 8;   // warps (= instances) per CTA of the pointer kernel
// glimpse kernel: up to 10 warps per CTA, two CTAs per SM.  Every warp streams its instance's embeddings through a
// private cp.async ring of GNS slices of 8 node rows (rows padded to 544 B: the LDS.128 fragment reads are conflict free),
// so the loads in flight cost no registers; the P[n][8] slot behind the ring is sized by N.  Measured at C4: the kernel is
// bound by dependent-instruction latency, i.e. by the number of resident warps (11 warps: 451 us, as much as the
// register-staged version at 15 warps with its exposed load latency).
constexpr int GW_MAX = 10, GNS = 2;
constexpr int G_ROW = 544;                          // bytes per staged node row (512 + 32)
constexpr int G_SLICE = 8 * G_ROW;                  // 4352 B
constexpr int G_CTA_SMEM = 113 * 1024;              // two CTAs per SM
static_assert(GNS * G_SLICE >= QW * 4, "c[1024] is staged in the ring");
__host__ __device__ inline int glimpse_warp_bytes(int N) { return GNS * G_SLICE + ((N * NH * 4 + 15) & ~15); }

// true when the episode was over before step trel: nobody was unfinished at the previous step
__device__ __forceinline__ bool episode_over(const RolloutParams& p, int trel) {
  return trel > 0 && ld_acquire_i(p.notdone + trel - 1) == 0;
}

// ---------------------------------------------------------------- glimpse: tables -> softmax -> c
__global__ void __launch_bounds__(GW_MAX * 32, 2) k_step_glimpse(const RolloutParams p, int t) {
  extern __shared__ __align__(16) unsigned char gsm[];
  const int trel = t - p.t0;
  if (episode_over(p, trel)) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int N = p.env.N, kind = p.env.kind;
  const int64_t B = p.env.B, b = (int64_t)blockIdx.x * (blockDim.x >> 5) + warp;
  if (b >= B) return;
  unsigned char* ring = gsm + warp * glimpse_warp_bytes(N);
  float* slot = reinterpret_cast<float*>(ring + GNS * G_SLICE);
  const int g = lane >> 2, tq = lane & 3;
  const float4* hrow = reinterpret_cast<const float4*>(p.h + b * N * E);
  // slice s = node rows 8s .. 8s+7 -> ring stage s % GNS; lane = 16-byte chunk of the row
  auto issue_slice = [&](int sidx) {
    unsigned char* dst = ring + (sidx % GNS) * G_SLICE + lane * 16;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (8 * sidx + r < N) cp_async16(dst + r * G_ROW, hrow + (8 * sidx + r) * (E / 4) + lane);
    cp_async_commit();
  };
  const int nsl = (N + 7) / 8;
#pragma unroll
  for (int sidx = 0; sidx < GNS; ++sidx) {
    if (sidx < nsl) issue_slice(sidx);
    else cp_async_commit();
  }
  const float lf = (float)p.env.load[b];
  if (p.mask_hist && lane < 4) p.mask_hist[((int64_t)trel * B + b) * 4 + lane] = __ldcg(p.env.mask + b * 4 + lane);
  if (p.load_hist && lane == 0) p.load_hist[(int64_t)trel * B + b] = lf;

  // ---- scores = S1[b][last] + S0[b] (+ load · SL[b]) + scrambled additive mask (graph_decoder.py:93-94)
  float pr[NH][4];
  {
    const int last = __ldcg(p.env.cur + b);
    const float* r1 = p.s1 + (((size_t)b * N + last) * NH) * N;
    const float* r0 = p.s0 + (size_t)b * NH * N;
    const float* rl = p.sl + (size_t)b * NH * N;
    // lane j holds mask word (j & 3) of the instance whose mask the reference adds to head j >> 2
    const uint32_t mword = __ldcg(p.env.mask + quirk_row(b, lane >> 2, p.G) * 4 + (lane & 3));
    // All table loads of a 32-node stride are issued before the first one is used (clamped addresses instead of
    // predicated loads: the source-level profile showed one exposed DRAM round trip per (head, stride) otherwise).
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (32 * i >= N) {   // uniform: no node in this stride
#pragma unroll
        for (int hh = 0; hh < NH; ++hh) pr[hh][i] = -INFINITY;
        continue;
      }
      const int n = lane + 32 * i, nc = (n < N) ? n : N - 1;
      float v1[NH], v0[NH], vl[NH];
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        v1[hh] = __ldg(r1 + hh * N + nc);
        v0[hh] = __ldcg(r0 + hh * N + nc);
        vl[hh] = (kind == VRPX_IRP) ? __ldcg(rl + hh * N + nc) : 0.f;
      }
#pragma unroll
      for (int hh = 0; hh < NH; ++hh) {
        const uint32_t wsel = __shfl_sync(0xffffffffu, mword, hh * 4 + i);
        float v = v1[hh] + v0[hh];
        if (kind == VRPX_IRP) v = fmaf(lf, vl[hh], v);
        v += (float)((wsel >> lane) & 1u);
        pr[hh][i] = (n < N) ? v : -INFINITY;
      }
    }
  }
  // ---- softmax per head over nodes (lane = node; only the ceil(N / 32) strides that hold nodes are touched)
#pragma unroll
  for (int hh = 0; hh < NH; ++hh) {
    float mx = pr[hh][0];
#pragma unroll
    for (int i = 1; i < 4; ++i)
      if (32 * i < N) mx = fmaxf(mx, pr[hh][i]);
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (32 * i < N) {
        pr[hh][i] = expf(pr[hh][i] - mx);   // exp(-inf) = 0 for the lanes beyond N (ex2.approx was measured: no time gained,
                                            // mean logit error 4x larger)
        sum += pr[hh][i];
      }
    sum = warp_sum(sum);
    const float inv = 1.0f / sum;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (32 * i < N) pr[hh][i] *= inv;
  }
  // probabilities to the slot as P[n][8]
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int n = lane + 32 * i;
    if (n < N) {
      *reinterpret_cast<float4*>(slot + n * 8) = make_float4(pr[0][i], pr[1][i], pr[2][i], pr[3][i]);
      *reinterpret_cast<float4*>(slot + n * 8 + 4) = make_float4(pr[4][i], pr[5][i], pr[6][i], pr[7][i]);
    }
  }
  __syncwarp();
  // ---- c[head][dim] = sum_n P[n][head] h_n[dim] on the tensor pipe (M = 16 dims, N = 8 heads, K = 8 nodes).
  // Thread g reads the float4 chunks 8c' + g (dims 32c' + 4g + e) of node rows n0+tq and n0+tq+4; m-tile
  // j = 2c' + u has row g <-> dim 32c'+4g+2u and row g+8 <-> dim 32c'+4g+2u+1.  B = P[n][head] from the slot.
  float cacc[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) cacc[j][i] = 0.f;
  for (int sidx = 0; sidx < nsl; ++sidx) {
    cp_async_wait<GNS - 1>();   // slice sidx has landed (one group per slice, possibly empty, keeps the count uniform)
    __syncwarp();
    const int n0 = 8 * sidx, na = n0 + tq, nbb = n0 + tq + 4;
    const unsigned char* st = ring + (sidx % GNS) * G_SLICE;
    float4 va[4], vb[4];
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      va[cq] = *reinterpret_cast<const float4*>(st + tq * G_ROW + (8 * cq + g) * 16);
      vb[cq] = *reinterpret_cast<const float4*>(st + (tq + 4) * G_ROW + (8 * cq + g) * 16);
    }
    if (n0 + 8 > N) {   // uniform: only the last slice has rows beyond N (stale ring contents)
#pragma unroll
      for (int cq = 0; cq < 4; ++cq) {
        if (na >= N) va[cq] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (nbb >= N) vb[cq] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
    }
    __syncwarp();   // every lane has read the stage before it is refilled
    if (sidx + GNS < nsl) issue_slice(sidx + GNS);
    else cp_async_commit();
    uint32_t bh0, bl0, bh1, bl1;
    split_tf32((na < N) ? slot[na * 8 + g] : 0.f, bh0, bl0);      // (k = tq,   n = head g)
    split_tf32((nbb < N) ? slot[nbb * 8 + g] : 0.f, bh1, bl1);    // (k = tq+4, n = head g)
#pragma unroll
    for (int cq = 0; cq < 4; ++cq) {
      const float ae[4] = {va[cq].x, va[cq].y, va[cq].z, va[cq].w};
      const float be[4] = {vb[cq].x, vb[cq].y, vb[cq].z, vb[cq].w};
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        uint32_t ah[4], al[4];
        split_tf32(ae[2 * u], ah[0], al[0]);       // (row g   = dim 32cq+4g+2u,   k = tq   = node na)
        split_tf32(ae[2 * u + 1], ah[1], al[1]);   // (row g+8 = dim 32cq+4g+2u+1, k = tq)
        split_tf32(be[2 * u], ah[2], al[2]);       // (row g,   k = tq+4 = node nbb)
        split_tf32(be[2 * u + 1], ah[3], al[3]);   // (row g+8, k = tq+4)
        mma_tf32_16x8x8(cacc[2 * cq + u], al, bh0, bh1);
        mma_tf32_16x8x8(cacc[2 * cq + u], ah, bl0, bl1);
        mma_tf32_16x8x8(cacc[2 * cq + u], ah, bh0, bh1);
      }
    }
  }
  cp_async_wait<0>();
  __syncwarp();
  // C fragment of m-tile j = 2cq+u: [0] (dim d, head 2tq), [1] (dim d, head 2tq+1), [2] (dim d+1, head 2tq),
  // [3] (dim d+1, head 2tq+1) with d = 32cq + 4g + 2u  ->  c[head][dim] staged in the (now idle) ring, then one
  // coalesced copy
  float* cst = reinterpret_cast<float*>(ring);
#pragma unroll
  for (int cq = 0; cq < 4; ++cq)
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int d = 32 * cq + 4 * g + 2 * u, j = 2 * cq + u;
      *reinterpret_cast<float2*>(cst + (2 * tq) * E + d) = make_float2(cacc[j][0], cacc[j][2]);
      *reinterpret_cast<float2*>(cst + (2 * tq + 1) * E + d) = make_float2(cacc[j][1], cacc[j][3]);
    }
  __syncwarp();
  float4* dst = reinterpret_cast<float4*>(p.cbuf + b * QW);
  for (int i = lane; i < QW / 4; i += 32) dst[i] = *reinterpret_cast<const float4*>(cst + 4 * i);
}

// ---------------------------------------------------------------- pointer: logits -> action -> environment
__global__ void __launch_bounds__(SW * 32) k_step_pointer(const RolloutParams p, int t) {
  __shared__ float s_slot[SW][VRPX_MAX_NODES];
  __shared__ float s_lp[SW];
  __shared__ int s_act[SW], s_anyleft;
  const int trel = t - p.t0;
  if (episode_over(p, trel)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.env.N, kind = p.env.kind;
  // the CTAs walk the batch in the opposite direction to the glimpse kernel: the embeddings that kernel read last are
  // still in L2 when this one starts, and the ones read last here are there for the next step's glimpse kernel
  const int64_t B = p.env.B, base = (int64_t)(gridDim.x - 1 - blockIdx.x) * SW;
  const int cnt = (int)((B - base < SW) ? (B - base) : SW);
  if (tid == 0) s_anyleft = 0;
  if (warp < cnt) {
    const int64_t b = base + warp;
    float* slot = s_slot[warp];
    const float4 qh = __ldcg(reinterpret_cast<const float4*>(p.qhat + b * E) + lane);
    const float4* hp = reinterpret_cast<const float4*>(p.h + b * N * E) + lane;
    // own mask (graph_decoder.py:98).  The logits of masked nodes are never used (they become -inf below), so only the
    // embedding rows of the CANDIDATE nodes are read: on average half of the instance over an episode.
    uint32_t mw[4], cand[4];
    int cum[5];
    cum[0] = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      mw[i] = __ldcg(p.env.mask + b * 4 + i);
      const int rem = N - 32 * i;   // nodes of this word
      const uint32_t range = rem >= 32 ? 0xffffffffu : (rem > 0 ? ((1u << rem) - 1u) : 0u);
      cand[i] = ~mw[i] & range;
      cum[i + 1] = cum[i] + __popc(cand[i]);
    }
    const int total = cum[4];
    // lanes 0..7 resolve the node index of candidate j0 + lane (-1 beyond the list)
    auto resolve = [&](int j0) {
      int idx = -1;
      const int j = j0 + lane;
      if (lane < 8 && j < total) {
        const int w = (j >= cum[1]) + (j >= cum[2]) + (j >= cum[3]);
        const uint32_t word = w == 0 ? cand[0] : (w == 1 ? cand[1] : (w == 2 ? cand[2] : cand[3]));
        const int base_cnt = w == 0 ? 0 : (w == 1 ? cum[1] : (w == 2 ? cum[2] : cum[3]));
        idx = 32 * w + (int)__fns(word, 0, j - base_cnt + 1);
      }
      return idx;
    };
    {
      int myidx = resolve(0);
      float4 nxt[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int n = __shfl_sync(0xffffffffu, myidx, i);
        nxt[i] = (n >= 0) ? __ldg(hp + n * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
      }
      for (int j0 = 0; j0 < total; j0 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 hv = nxt[i];
          v[i] = fmaf(qh.x, hv.x, fmaf(qh.y, hv.y, fmaf(qh.z, hv.z, qh.w * hv.w)));
        }
        const int nsel = __shfl_sync(0xffffffffu, myidx, (lane >> 2) & 7);   // node of the sum this lane group reduces
        myidx = resolve(j0 + 8);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int n = __shfl_sync(0xffffffffu, myidx, i);
          nxt[i] = (n >= 0) ? __ldg(hp + n * (E / 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const float sc = reduce8(v, lane);
        if ((lane & 3) == 0 && nsel >= 0) slot[nsel] = 10.0f * tanhf(sc);
      }
    }
    __syncwarp();
    float u[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int n = lane * 4 + i;
      const bool ok = n < N && !((mw[n >> 5] >> (n & 31)) & 1u);
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
      const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oam = __shfl_xor_sync(0xffffffffu, am, o);
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
        const float y = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += y;
      }
      const float total = __shfl_sync(0xffffffffu, incl, 31);
      if (p.mode == VRPX_SAMPLE) {
        const uint64_t gid = p.offset + (uint64_t)b;
        const uint4 r = philox4x32_10(make_uint4((uint32_t)gid, (uint32_t)(gid >> 32), (uint32_t)t, 0x5eedu),
                                      make_uint2((uint32_t)p.seed, (uint32_t)(p.seed >> 32)));
        const float thr = u24(r.x) * total;
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
      const float usel = ai == 0 ? u[0] : (ai == 1 ? u[1] : (ai == 2 ? u[2] : u[3]));
      const float ua = __shfl_sync(0xffffffffu, usel, a >> 2);
      lp = (ua - mx) - logf(total);
    }
    if (lane == 0) { s_act[warp] = a; s_lp[warp] = lp; }
  }
  __syncthreads();
  // environment transition: one THREAD per instance (the dependent global loads of the CTA's instances overlap)
  bool unfinished = false;
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
    if (kind == VRPX_IRP) {
      Bits128 x = demand_exceeds(dem, N, load);
#pragma unroll
      for (int i = 0; i < 4; ++i) p.env.mask[b * 4 + i] = v.w[i] | x.w[i];
    }
    // f32 accumulation of f32(reward) in step order (graph_tsp_agent.py:85); cost = -acc_loss
    p.cost[b] = (t == 0 ? 0.f : p.cost[b]) + (float)r.dist;
    if (p.mode != VRPX_GREEDY) p.logp[b] = (t == 0 ? 0.f : p.logp[b]) + s_lp[m];
    else if (t == 0) p.logp[b] = 0.f;
    if (!r.all_before) unfinished = true;
  }
  if (unfinished) s_anyleft = 1;  // benign race: all writers store 1
  __syncthreads();
  if (tid == 0 && s_anyleft) atomicAdd(p.notdone + trel, 1);
}

// steps executed = first step at whose start nobody was unfinished (it still runs, like the reference's loop), else Tmax
__global__ void k_rollout_finish(const int* __restrict__ notdone, int t0, int Tmax, int* __restrict__ steps) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int t = t0 + Tmax;
  for (int i = 0; i < Tmax; ++i)
    if (notdone[i] == 0) { t = t0 + i + 1; break; }
  *steps = t;
}

__global__ void k_transpose_m(const float* __restrict__ m_t, float* __restrict__ m_nt) {   // [1024][128] -> [128][1024]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= QW * E) return;
  const int k = i >> 7, n = i & (E - 1);
  m_nt[n * QW + k] = m_t[i];
}

int prepare_split_weights(const float* m_t, float* m_nt, cudaStream_t stream) {
  k_transpose_m<<<QW * E / 256, 256, 0, stream>>>(m_t, m_nt);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

int run_split_steps(const RolloutParams& p, int t_first, const float* m_nt, cudaStream_t stream) {
  const int64_t B = p.env.B;
  int gw = G_CTA_SMEM / glimpse_warp_bytes(p.env.N);
  if (gw > GW_MAX) gw = GW_MAX;
  const unsigned grid = (unsigned)((B + SW - 1) / SW), ggrid = (unsigned)((B + gw - 1) / gw);
  const size_t smem = (size_t)gw * glimpse_warp_bytes(p.env.N);
  VRPX_CUDA(cudaFuncSetAttribute(k_step_glimpse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  for (int t = t_first; t < p.t0 + p.Tmax; ++t) {
    k_step_glimpse<<<ggrid, gw * 32, smem, stream>>>(p, t);
    VRPX_LAUNCH_CHECK();
    GemmArgs ga{p.cbuf, B, QW, m_nt, E, p.w.m_c, 0, nullptr, nullptr, nullptr, p.qhat};
    int rc = gemm_tc(ga, stream);
    if (rc) return rc;
    k_step_pointer<<<grid, SW * 32, 0, stream>>>(p, t);
    VRPX_LAUNCH_CHECK();
  }
  k_rollout_finish<<<1, 32, 0, stream>>>(p.notdone, p.t0, p.Tmax, p.steps);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace vrpx

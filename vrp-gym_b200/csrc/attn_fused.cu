// attn_fused.cu — encoder QKV projection + per-instance self-attention in ONE kernel (graph_encoder.py:74-104: the
// nn.MultiheadAttention call of a layer: in_proj, 8 heads of dim 16, softmax(QK^T / 4) V; the out-projection stays a GEMM).
//
// The two-kernel form (k_gemm_tc4 writes QKV [R][384] to HBM, k_enc_attention_f16 reads it back) moves 5 GB each way per
// layer at C4 (65,536 x TSP-50) around 1.7 GB of embeddings in and 1.7 GB of attention out.  Here Q, K and V never leave
// the SM.  Structure = score_table_fused.cu:
//   * a tile = TI = floor(128 / N) WHOLE instances (TI·N <= 128 rows of h);
//   * CTAs are specialised by head group: CTA type g = blockIdx % 2 keeps the 96 KiB of split in_proj weights of heads
//     4g..4g+3 resident in shared memory for its whole life and walks over all tiles; the two types of a tile run side
//     by side, so h comes from DRAM once;
//   * per (tile, head): D[128 x 48] = XA · W_head^T on tcgen05 (A = the tile as f16 hi/lo in tensor memory, cross terms
//     first, gemm_tc4.cu; W_head = the head's q | k | v rows), then the epilogue warps add the bias and stage D in shared
//     memory already split and in mma.sync fragment order (the layouts of k_enc_attention_f16: Q and K [row][lane slot],
//     V transposed [dim][k16 step][lane slot]) and run that kernel's query-tile loop: one task = (instance, 16 queries).
// Roles (24 warps): 0-3 X converters (raw fp32 smem tile -> XA in TMEM), 4 TMA producer, 5 MMA issuer (whole warp,
// elected lane), 8-15 and 16-23 two epilogue + attention TEAMS (two warps per TMEM lane quarter: one stages q | k, the
// other v): team A owns the even (tile, head) steps, accumulator 0 and staging buffer 0, team B the odd ones — a step is
// a serial chain (TMEM read-out -> split -> barrier -> attention) of 8 warps, two of them in flight fill the issue slots.
// TMEM: XA 128 columns | D 2 x 64 (48 used).  Shared memory: weights 96 KiB | raw X 64 KiB | 2 x (Q 9 KiB | K 8.5 KiB | V).
#include "f16split.cuh"
#include "gemm.cuh"
#include "tc_common.cuh"

namespace vrpx {
namespace qa {
using namespace tc4;

constexpr int DH = 16;                  // encoder head dim
constexpr int HN = 3 * DH;              // 48 projection columns per head: q | k | v
constexpr int HPC = 4;                  // heads per CTA type
constexpr int NTYPES = NH / HPC;        // 2
constexpr int NTHREADS = 768;
constexpr int W_TMA = 4, W_MMA = 5, W_EPI0 = 8;
constexpr int XBOX = 16 * 1024;         // raw X box: 128 rows x 32 floats
constexpr int WBOX = HN * 128;          // weight box: 48 rows x 128 bytes (64 halves) = 6 KiB
constexpr int W_HEAD = 4 * WBOX;        // hi k0-63 | hi k64-127 | lo k0-63 | lo k64-127 = 24 KiB
constexpr int QROWS = 144, KROWS = 136; // tile rows + the overhang of the last query / key tile of the last instance
constexpr int SM_W = 0;
constexpr int SM_X = HPC * W_HEAD;              // 96 KiB
constexpr int SM_BIAS = SM_X + 4 * XBOX;        // [HPC][48] f32 (q part pre-scaled)
constexpr int SM_STAGE = SM_BIAS + HPC * HN * 4;
// one staging buffer per team: Qf [QROWS][4] uint4 | Kf [KROWS][4] uint4 | Vf [TI][16][VDS] uint4 (sized by the host)
constexpr int ST_KF = QROWS * 64, ST_VF = ST_KF + KROWS * 64;
constexpr int SMEM_MAX = 227 * 1024 - 512;
constexpr int VF_MAX = (SMEM_MAX - 1024 - SM_STAGE) / 2 - ST_VF;   // bytes of Vf per team that still fit
static_assert(VF_MAX >= 16 * 36 * 16, "one instance of the largest bucket must fit");
constexpr uint32_t TM_XA = 0, TM_D = 128, TMEM_COLS = 256;
constexpr uint32_t IDESC = make_idesc(128, HN);
// scores are kept in log2 units (Q scaled by log2(e) / sqrt(16)): the softmax needs one ex2 per element
constexpr float QS = 0.25f * 1.4426950408889634f;

// named barrier of one epilogue team (8 warps).  Immediate ids (a register id makes ptxas reserve all 16 barriers), and
// every use is ONE instruction that all warps of the team reach: compute-sanitizer's synccheck rejects a named barrier
// reached through different instructions, and its racecheck only understands bar.sync ordering (an mbarrier in this
// place is reported as 264 shared-memory hazards).
__device__ __forceinline__ void team_barrier(int team) {
  __syncwarp();
  if (team == 0) asm volatile("bar.sync 1, 256;" ::: "memory");
  else asm volatile("bar.sync 2, 256;" ::: "memory");
}

#define VRPX_QA_LD16(v, taddr)                                                                                      \
  asm volatile(                                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                     \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                              \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                    \
      : "r"(taddr)                                                                                                  \
      : "memory")

// in_proj_w [384][128] f32 (rows 0..127 = W_q, 128..255 = W_k, 256..383 = W_v; head h = rows 16h..16h+15 of each) ->
// head-major split halves whi | wlo [8][48][128] f16 of W * 2^8: rows 0..15 of head h = its q rows, 16..31 k, 32..47 v
__global__ void k_prepare_inproj(const float* __restrict__ w, __half* __restrict__ whi, __half* __restrict__ wlo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= NH * HN * E) return;
  const int k = i & (E - 1), r = (i >> 7) % HN, hh = i / (HN * E);
  const int src = (r / DH) * E + hh * DH + (r % DH);
  const float x = w[src * E + k] * W_SCALE;
  const __half hgh = __float2half_rn(x);
  whi[i] = hgh;
  wlo[i] = __float2half_rn(x - __half2float(hgh));
}

// NK8 = ceil(N / 8) key tiles of the instance: compile-time, so that the unrolled score / softmax / PV loops carry no
// run-time guards (the only run-time mask is on the tile that straddles N)
template <int NK8>
__global__ void __launch_bounds__(NTHREADS, 1)
k_qkv_attention(const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapWh,
                const __grid_constant__ CUtensorMap mapWl, const float* __restrict__ bias, float* __restrict__ att, int64_t B,
                int N, int TI, int stage_bytes) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_w_full, s_xr_full, s_xr_free, s_xa_full, s_xa_free, s_d_full[2], s_d_free[2];
  __shared__ uint32_t s_tmem;
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);
  constexpr int NJJ = (NK8 + 1) / 2;               // k16 steps over the keys
  constexpr int VDS = 4 * (NJJ + (NJJ & 1)) + 4;   // chunks (16 B) per dim row of Vf (k_enc_attention_f16)

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int type = blockIdx.x % NTYPES, stream = blockIdx.x / NTYPES, nstreams = gridDim.x / NTYPES;
  const int64_t ntiles = (B + TI - 1) / TI;
  const int head0 = type * HPC;

  if (tid == 0) {
    mbar_init(smem_u32(&s_w_full), 1);
    mbar_init(smem_u32(&s_xr_full), 1);
    mbar_init(smem_u32(&s_xr_free), 4);
    mbar_init(smem_u32(&s_xa_full), 4);
    mbar_init(smem_u32(&s_xa_free), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_d_full[i]), 1);
      mbar_init(smem_u32(&s_d_free[i]), 8);   // the 8 warps of the team that owns the accumulator
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  // Q / K overhang rows and the padded keys of V stay zero (finite) for the whole kernel
  for (int i = tid; i < 2 * stage_bytes / 16; i += NTHREADS) reinterpret_cast<uint4*>(smem + SM_STAGE)[i] = make_uint4(0u, 0u, 0u, 0u);
  for (int i = tid; i < HPC * HN; i += NTHREADS) {
    const int hs = i / HN, c = i % HN;
    const float bv = bias ? bias[(c / DH) * E + (head0 + hs) * DH + (c % DH)] : 0.f;
    reinterpret_cast<float*>(smem + SM_BIAS)[i] = (c < DH) ? bv * QS : bv;
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  if (warp == W_TMA) {
    // ===================== TMA producer: the resident weights once, then the raw X tile of every visit =====================
    if (lane == 0) {
      {
        const uint32_t bar = smem_u32(&s_w_full);
        mbar_expect_tx(bar, HPC * W_HEAD);
        for (int hs = 0; hs < HPC; ++hs) {
          unsigned char* dst = smem + SM_W + hs * W_HEAD;
          const int row0 = (head0 + hs) * HN;
          tma_load_2d(smem_u32(dst), &mapWh, 0, row0, bar);
          tma_load_2d(smem_u32(dst + WBOX), &mapWh, 64, row0, bar);
          tma_load_2d(smem_u32(dst + 2 * WBOX), &mapWl, 0, row0, bar);
          tma_load_2d(smem_u32(dst + 3 * WBOX), &mapWl, 64, row0, bar);
        }
      }
      uint32_t cx = 0;
      for (int64_t tile = stream; tile < ntiles; tile += nstreams, ++cx) {
        mbar_wait_sleep(smem_u32(&s_xr_free), (cx & 1) ^ 1, 200);
        const uint32_t bar = smem_u32(&s_xr_full);
        mbar_expect_tx(bar, 4 * XBOX);
        const int row0 = (int)(tile * TI * N);
#pragma unroll
        for (int kq = 0; kq < 4; ++kq) tma_load_2d(smem_u32(smem + SM_X + kq * XBOX), &mapX, kq * 32, row0, bar);
      }
    }
  } else if (warp < 4) {
    // ===================== X converters: raw fp32 row -> f16 hi / lo packed words -> XA (thread = tile row) =====================
    const int r = tid;
    const uint32_t xa = tmem + ((uint32_t)(warp * 32) << 16) + TM_XA;
    uint32_t xi = 0;
    for (int64_t tile = stream; tile < ntiles; tile += nstreams, ++xi) {
      mbar_wait_sleep(smem_u32(&s_xr_full), xi & 1, 100);
#pragma unroll 1
      for (int kh = 0; kh < 2; ++kh) {
        uint32_t hi[32], lo[32];
#pragma unroll
        for (int half = 0; half < 2; ++half)
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 xv = *reinterpret_cast<const float4*>(smem + SM_X + (2 * kh + half) * XBOX + r * 128 + ((c ^ (r & 7)) << 4));
            split_pair(xv.x, xv.y, hi[half * 16 + 2 * c], lo[half * 16 + 2 * c]);
            split_pair(xv.z, xv.w, hi[half * 16 + 2 * c + 1], lo[half * 16 + 2 * c + 1]);
          }
        if (kh == 0) {
          mbar_wait_sleep(smem_u32(&s_xa_free), (xi & 1) ^ 1, 100);   // the previous visit's products are complete
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        }
        tmem_st32(xa + kh * 32, hi);
        tmem_st32(xa + 64 + kh * 32, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&s_xr_free));                       // the raw tile may be refilled
        mbar_arrive(smem_u32(&s_xa_full));
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer (whole warp converged, elected lane issues) =====================
    mbar_wait(smem_u32(&s_w_full), 0);
    uint32_t ti = 0, dc = 0;   // tile visits, accumulator uses
    for (int64_t tile = stream; tile < ntiles; tile += nstreams, ++ti) {
      mbar_wait_sleep(smem_u32(&s_xa_full), ti & 1, 50);
      for (int hs = 0; hs < HPC; ++hs, ++dc) {
        const uint32_t s = dc & 1, ph = (dc >> 1) & 1;
        mbar_wait_sleep(smem_u32(&s_d_free[s]), ph ^ 1, 50);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + TM_D + s * 64, xh = tmem + TM_XA, xl = tmem + TM_XA + 64;
        unsigned char* wb = smem + SM_W + hs * W_HEAD;
        const uint64_t wh[2] = {make_desc(smem_u32(wb)), make_desc(smem_u32(wb + WBOX))};
        const uint64_t wl[2] = {make_desc(smem_u32(wb + 2 * WBOX)), make_desc(smem_u32(wb + 3 * WBOX))};
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint64_t o = (uint64_t)(2 * jj);
            const uint32_t ka = 8 * (4 * kh + jj);
            mma_f16_ts_w(d, xl + ka, wh[kh] + o, (kh | jj) ? 1u : 0u, IDESC);
            mma_f16_ts_w(d, xh + ka, wl[kh] + o, 1u, IDESC);
          }
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) mma_f16_ts_w(d, xh + 8 * (4 * kh + jj), wh[kh] + (uint64_t)(2 * jj), 1u, IDESC);
        mma_commit_w(smem_u32(&s_d_full[s]));
      }
      mma_commit_w(smem_u32(&s_xa_free));
    }
  } else if (warp >= W_EPI0) {
    // ===================== epilogue + attention =====================
    const int team = (warp - W_EPI0) >> 3, ew = (warp - W_EPI0) & 7;
    const int q = warp & 3, half = ew >> 2;   // TMEM lane quarter; 0: q | k columns, 1: v columns
    const int g = lane >> 2, t = lane & 3;
    unsigned char* stage = smem + SM_STAGE + team * stage_bytes;
    uint4* Qf = reinterpret_cast<uint4*>(stage);
    uint4* Kf = reinterpret_cast<uint4*>(stage + ST_KF);
    uint4* Vf = reinterpret_cast<uint4*>(stage + ST_VF);
    const float* sbias = reinterpret_cast<const float*>(smem + SM_BIAS);
    const int MT = (N + 15) / 16;
    const int r = q * 32 + lane, ri = r / N, rn = r - ri * N;   // this thread's tile row: instance and node
    uint32_t dc = 0;
    for (int64_t tile = stream; tile < ntiles; tile += nstreams) {
      const int64_t b0 = tile * TI;
      const int ninst = (int)((B - b0 < TI) ? (B - b0) : TI);
      for (int hs = 0; hs < HPC; ++hs, ++dc) {
        const uint32_t s = dc & 1, ph = (dc >> 1) & 1;
        if ((int)s != team) continue;
        const int head = head0 + hs;
        const float* bs = sbias + hs * HN;
        team_barrier(team);   // every warp of the team is done with the previous step's Q / K / V
        mbar_wait(smem_u32(&s_d_full[s]), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + TM_D + s * 64;
        if (half == 0) {
          // ---- q | k of tile row r -> fragment order [row][slot tt] = {hi(dims 2tt, +1), hi(dims 2tt+8, +9), lo(..), lo(..)}
          uint32_t v[32];
          VRPX_QA_LD16(v, taddr);
          VRPX_QA_LD16((v + 16), taddr + 16);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_d_free[s]));
          uint4 fq[4], fk[4];
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
            const float q0 = fmaf(__uint_as_float(v[2 * tt]), OUT_SCALE * QS, bs[2 * tt]);
            const float q1 = fmaf(__uint_as_float(v[2 * tt + 1]), OUT_SCALE * QS, bs[2 * tt + 1]);
            const float q8 = fmaf(__uint_as_float(v[2 * tt + 8]), OUT_SCALE * QS, bs[2 * tt + 8]);
            const float q9 = fmaf(__uint_as_float(v[2 * tt + 9]), OUT_SCALE * QS, bs[2 * tt + 9]);
            const uint2 p0 = split_f16x2_u(q0, q1), p1 = split_f16x2_u(q8, q9);
            fq[tt] = make_uint4(p0.x, p1.x, p0.y, p1.y);
          }
#pragma unroll
          for (int tt = 0; tt < 4; ++tt) {
            const float k0 = fmaf(__uint_as_float(v[16 + 2 * tt]), OUT_SCALE, bs[16 + 2 * tt]);
            const float k1 = fmaf(__uint_as_float(v[16 + 2 * tt + 1]), OUT_SCALE, bs[16 + 2 * tt + 1]);
            const float k8 = fmaf(__uint_as_float(v[16 + 2 * tt + 8]), OUT_SCALE, bs[16 + 2 * tt + 8]);
            const float k9 = fmaf(__uint_as_float(v[16 + 2 * tt + 9]), OUT_SCALE, bs[16 + 2 * tt + 9]);
            const uint2 p0 = split_f16x2_u(k0, k1), p1 = split_f16x2_u(k8, k9);
            fk[tt] = make_uint4(p0.x, p1.x, p0.y, p1.y);
          }
          if (ri < ninst) {
#pragma unroll
            for (int tt = 0; tt < 4; ++tt) {
              Qf[r * 4 + tt] = fq[tt];
              Kf[r * 4 + tt] = fk[tt];
            }
          }
        } else {
          // ---- v of tile row r (key rn of instance ri) -> Vf[ri][dim][k16 step jj][slot tt]: this key's half-word of
          // {hi(keys 16jj+2tt, +1), hi(keys 16jj+2tt+8, +9), lo(..), lo(..)}
          uint32_t v[16];
          VRPX_QA_LD16(v, taddr + 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_d_free[s]));
          __half hv[16], lv[16];
#pragma unroll
          for (int d = 0; d < 16; ++d) {
            const float x = fmaf(__uint_as_float(v[d]), OUT_SCALE, bs[32 + d]);
            hv[d] = __float2half_rn(x);
            lv[d] = __float2half_rn(x - __half2float(hv[d]));
          }
          if (ri < ninst) {
            const int jj = rn >> 4, kk = rn & 15, tt = (kk & 7) >> 1;
            unsigned char* vb = reinterpret_cast<unsigned char*>(Vf + (size_t)ri * 16 * VDS) + (kk >> 3) * 4 + (kk & 1) * 2;
#pragma unroll
            for (int d = 0; d < 16; ++d) {
              unsigned char* p = vb + (d * VDS + ((4 * jj + tt) ^ (d & 4))) * 16;
              *reinterpret_cast<__half*>(p) = hv[d];
              *reinterpret_cast<__half*>(p + 8) = lv[d];
            }
          }
        }
        team_barrier(team);   // Q, K, V of the (tile, head) are in shared memory
        // ---- attention: task = (instance i, 16-row query tile m), round robin over the 8 warps
        for (int task = ew; task < ninst * MT; task += 8) {
          const int i = task / MT, m = task - i * MT;
          const int qa = 16 * m + g, qb = qa + 8;
          const uint4* Kh = Kf + (size_t)i * N * 4;
          const uint4* Vh = Vf + (size_t)i * 16 * VDS;
          uint32_t qh[4], ql[4];
          {
            const uint4 fa = Qf[(i * N + qa) * 4 + t], fb = Qf[(i * N + qb) * 4 + t];
            qh[0] = fa.x; qh[2] = fa.y; ql[0] = fa.z; ql[2] = fa.w;
            qh[1] = fb.x; qh[3] = fb.y; ql[1] = fb.z; ql[3] = fb.w;
          }
          // S = Q K^T: C fragment of key tile j: [0], [1] = row g, keys 8j+2t, +1; [2], [3] = row g+8.  Key tiles entirely
          // beyond N are skipped (their probabilities stay 0); only the tile that straddles N is masked.
          float sc[2 * NJJ][4];
          float ma = -INFINITY, mb = -INFINITY;
          // two key tiles at a time, their three-MMA chains interleaved (a chain alone waits out the HMMA latency twice)
#pragma unroll
          for (int j = 0; j < 2 * NJJ; j += 2) {
#pragma unroll
            for (int e = 0; e < 4; ++e) sc[j][e] = sc[j + 1][e] = 0.f;
            if (j < NK8) {
              const bool two = j + 1 < NK8;   // compile-time after unrolling
              const uint4 k0 = Kh[(8 * j + g) * 4 + t];
              const uint4 k1 = two ? Kh[(8 * j + 8 + g) * 4 + t] : make_uint4(0u, 0u, 0u, 0u);
              mma_f16_16x8x16(sc[j], ql, k0.x, k0.y);
              if (two) mma_f16_16x8x16(sc[j + 1], ql, k1.x, k1.y);
              mma_f16_16x8x16(sc[j], qh, k0.z, k0.w);
              if (two) mma_f16_16x8x16(sc[j + 1], qh, k1.z, k1.w);
              mma_f16_16x8x16(sc[j], qh, k0.x, k0.y);
              if (two) mma_f16_16x8x16(sc[j + 1], qh, k1.x, k1.y);
#pragma unroll
              for (int jx = j; jx < j + 2; ++jx) {
                if (jx >= NK8) continue;
                if (jx == NK8 - 1 && (N & 7)) {
#pragma unroll
                  for (int e = 0; e < 2; ++e)
                    if (8 * jx + 2 * t + e >= N) { sc[jx][e] = -INFINITY; sc[jx][2 + e] = -INFINITY; }
                }
                ma = fmaxf(ma, fmaxf(sc[jx][0], sc[jx][1]));
                mb = fmaxf(mb, fmaxf(sc[jx][2], sc[jx][3]));
              }
            }
          }
          ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 1)); ma = fmaxf(ma, __shfl_xor_sync(0xffffffffu, ma, 2));
          mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 1)); mb = fmaxf(mb, __shfl_xor_sync(0xffffffffu, mb, 2));
          float sa = 0.f, sb = 0.f;
#pragma unroll
          for (int j = 0; j < 2 * NJJ; ++j) {
            if (j < NK8) {
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
          // O = P V: k16 step jj over keys 16jj..16jj+15; A = P from the score registers of key tiles 2jj, 2jj+1
          float o[2][4];
#pragma unroll
          for (int d = 0; d < 2; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
          for (int jj = 0; jj < NJJ; ++jj) {
            const uint2 p0 = split_f16x2_u(sc[2 * jj][0], sc[2 * jj][1]), p1 = split_f16x2_u(sc[2 * jj][2], sc[2 * jj][3]);
            const uint2 p2 = split_f16x2_u(sc[2 * jj + 1][0], sc[2 * jj + 1][1]), p3 = split_f16x2_u(sc[2 * jj + 1][2], sc[2 * jj + 1][3]);
            const uint32_t ph_[4] = {p0.x, p1.x, p2.x, p3.x}, pl_[4] = {p0.y, p1.y, p2.y, p3.y};
            const uint4 v0 = Vh[g * VDS + ((4 * jj + t) ^ (g & 4))], v1 = Vh[(8 + g) * VDS + ((4 * jj + t) ^ (g & 4))];
            mma_f16_16x8x16(o[0], pl_, v0.x, v0.y);
            mma_f16_16x8x16(o[1], pl_, v1.x, v1.y);
            mma_f16_16x8x16(o[0], ph_, v0.z, v0.w);
            mma_f16_16x8x16(o[1], ph_, v1.z, v1.w);
            mma_f16_16x8x16(o[0], ph_, v0.x, v0.y);
            mma_f16_16x8x16(o[1], ph_, v1.x, v1.y);
          }
          const float ia = 1.0f / sa, ib = 1.0f / sb;
          float* oa = att + ((b0 + i) * N + qa) * E + head * DH + 2 * t;
          float* ob = att + ((b0 + i) * N + qb) * E + head * DH + 2 * t;
#pragma unroll
          for (int d = 0; d < 2; ++d) {
            if (qa < N) *reinterpret_cast<float2*>(oa + 8 * d) = make_float2(o[d][0] * ia, o[d][1] * ia);
            if (qb < N) *reinterpret_cast<float2*>(ob + 8 * d) = make_float2(o[d][2] * ib, o[d][3] * ib);
          }
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

template <int NK8>
static int launch(const CUtensorMap& mx, const CUtensorMap& mwh, const CUtensorMap& mwl, const float* bias, float* att, int64_t B,
                  int N, cudaStream_t stream) {
  constexpr int NJJ = (NK8 + 1) / 2;
  constexpr int VDS = 4 * (NJJ + (NJJ & 1)) + 4;
  int TI = 128 / N;
  if (TI > VF_MAX / (16 * VDS * 16)) TI = VF_MAX / (16 * VDS * 16);
  const int stage_bytes = ST_VF + TI * 16 * VDS * 16;
  const int SMEM_BYTES = SM_STAGE + 2 * stage_bytes + 1024;
  VRPX_CUDA(cudaFuncSetAttribute(k_qkv_attention<NK8>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t ntiles = (B + TI - 1) / TI;
  int64_t streams = num_sms() / NTYPES;
  if (streams > ntiles) streams = ntiles;
  if (streams < 1) streams = 1;
  k_qkv_attention<NK8><<<(unsigned)(streams * NTYPES), NTHREADS, SMEM_BYTES, stream>>>(mx, mwh, mwl, bias, att, B, N, TI, stage_bytes);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace qa

// att [B·N][128] = concat_h softmax(Q_h K_h^T / 4) V_h with [Q | K | V] = X · in_proj_w^T + in_proj_b, per instance of N
// rows of X [B·N][128]
int qkv_attention_fused(const float* X, const float* in_proj_w, const float* in_proj_b, int64_t B, int N, float* att,
                        cudaStream_t stream) {
  using namespace qa;
  if (N < 1 || N > VRPX_MAX_NODES || B < 1) {
    set_error("qkv_attention_fused: bad argument");
    return VRPX_ERR_ARG;
  }
  __half* w16 = split_scratch(stream);
  if (!w16) return VRPX_ERR_CUDA;
  constexpr int NW = NH * HN * E;
  k_prepare_inproj<<<(NW + 255) / 256, 256, 0, stream>>>(in_proj_w, w16, w16 + NW);
  VRPX_LAUNCH_CHECK();
  CUtensorMap mx, mwh, mwl;
  int rc;
  if ((rc = make_map(&mx, X, B * N, E, false))) return rc;
  if ((rc = make_map(&mwh, w16, NH * HN, E, true, HN))) return rc;
  if ((rc = make_map(&mwl, w16 + NW, NH * HN, E, true, HN))) return rc;
  switch ((N + 7) / 8) {
#define VRPX_QA_CASE(K) case K: return launch<K>(mx, mwh, mwl, in_proj_b, att, B, N, stream);
    VRPX_QA_CASE(1) VRPX_QA_CASE(2) VRPX_QA_CASE(3) VRPX_QA_CASE(4) VRPX_QA_CASE(5) VRPX_QA_CASE(6) VRPX_QA_CASE(7) VRPX_QA_CASE(8)
    VRPX_QA_CASE(9) VRPX_QA_CASE(10) VRPX_QA_CASE(11) VRPX_QA_CASE(12) VRPX_QA_CASE(13) VRPX_QA_CASE(14) VRPX_QA_CASE(15) VRPX_QA_CASE(16)
#undef VRPX_QA_CASE
  }
  set_error("qkv_attention_fused: N=%d out of range", N);
  return VRPX_ERR_ARG;
}

}  // namespace vrpx

// ff_fused.cu — the encoder's feed-forward block as ONE tcgen05 kernel (agents/graph_encoder.py:177-181,196):
//
//     Y = ( X + relu(X · W1^T + b1) · W2^T + b2 ) * scale + shift          X, Y [R][128] f32, hidden width 512
//
// (scale / shift = the eval-mode BatchNorm folded affine of `bn2`, or none in train mode).  The 512-wide hidden activation
// never leaves the SM: per 128-row tile it is produced in eight 64-column chunks into tensor memory, converted in place to
// the f16 hi/lo A operand of the second product, and consumed from tensor memory.  Unfused, the hidden tile costs a 6.7 GB
// write and a 6.7 GB read per layer at R = 3.28 M rows — the two GEMMs it replaces were HBM-bound on exactly that traffic.
//
// Precision: the f16 hi/lo split of gemm_tc4.cu (three MMAs per product, ~fp32).  The tensor core truncates when it adds a
// k16 step into its fp32 accumulator, so the order of the MMAs matters (tools/gemm_error_probe.py):
//   FF1 chunk (K = 128, all of it resident): the 16 cross-term MMAs are issued before the 8 main ones;
//   FF2 (K = 512, accumulated over the chunks): main and cross terms have their own accumulators, added in fp32 by the
//   epilogue.
//
// Both products take their A operand from TENSOR MEMORY (the X tile as well: with A in shared memory an M128 x N64 x K16
// MMA reads 6 KB of operands per 32 clocks, more than the 128 B/clk a shared memory delivers — the first version of this
// kernel ran at 45 % of the tensor rate for that reason); shared memory carries only the raw X tile and the weights.
//
// Roles (18 warps, one CTA per SM, persistent over the row tiles):
//   warp  8      TMA producer (one thread polling three rings): the raw fp32 X tile of the NEXT row tile (four SWIZZLE_128B
//                boxes), per chunk W1c hi/lo (64 hidden rows x K 128) and W2c hi/lo (128 outputs x K 64), two slots each
//   warp  9      MMA issuer (one thread): FF1(j+1) is issued BEFORE FF2(j), so the conversion of chunk j overlaps FF1(j+1)
//   warps 0-7    chunk converters (thread = tile row = TMEM lane x half of the chunk's 64 columns): D1 -> +b1, ReLU,
//                split -> HA (tcgen05.ld / st).  They are the serial resource the tensor pipe waits for (a chunk is
//                ~12 instructions per element, a third of them quarter-rate conversions): with four warps the kernel ran
//                at half the tensor rate
//   warps 10-17  X converters + epilogue.  First the NEXT tile's X: thread = (row, k half), raw smem row -> split -> XA in
//                TMEM (as soon as the current tile's FF1 products have released it); then the current tile's epilogue:
//                D2 main + cross -> +b2, +residual, affine -> coalesced stores through a swizzled smem transpose patch
// TMEM (512 columns): XA 64 hi + 64 lo | D1 64 | HA 32 hi + 32 lo | D2 main 128 | D2 cross 128.
// Shared memory (224 KiB): raw X 64 KiB | W1 ring 2 x 32 KiB | W2 ring 2 x 32 KiB | 8 transpose patches 32 KiB.
#include "gemm.cuh"
#include "tc_common.cuh"

namespace vrpx {
namespace ff {
using namespace tc4;

constexpr int BM = 128;                 // rows per tile
constexpr int HC = 64;                  // hidden columns per chunk
constexpr int NCH = FF / HC;            // 8 chunks
constexpr int W_TMA = 8, W_MMA = 9, W_EPI0 = 10;
constexpr int NWARPS = 18, NTHREADS = NWARPS * 32;
constexpr int XBOX = 16 * 1024;         // one SWIZZLE_128B box: 128 rows x 128 bytes (32 floats)
constexpr int SM_X = 0;                 // raw fp32 X tile: k 0-31 | 32-63 | 64-95 | 96-127
constexpr int SM_W1 = 4 * XBOX;         // slot: hi k 0-63 (64 rows x 128 B = 8 KiB) | hi k 64-127 | lo | lo
constexpr int W1_SLOT = 32 * 1024;
constexpr int SM_W2 = SM_W1 + 2 * W1_SLOT;   // slot: hi (128 rows x 128 B) | lo
constexpr int W2_SLOT = 32 * 1024;
constexpr int SM_PATCH = SM_W2 + 2 * W2_SLOT;
constexpr int SMEM_BYTES = SM_PATCH + 8 * 4096 + 1024;   // + alignment slack
constexpr uint32_t TM_XA = 0, TM_D1 = 128, TM_HA = 192, TM_D2M = 256, TM_D2X = 384, TMEM_COLS = 512;
constexpr uint32_t IDESC1 = make_idesc(BM, HC), IDESC2 = make_idesc(BM, E);

struct Args {
  const float* X;         // [R][128] input rows
  int64_t R;
  const float* b1;        // [512]
  const float* b2;        // [128]
  const float* residual;  // [R][128] (may alias X and Y)
  const float* scale;     // [128] or nullptr
  const float* shift;     // [128] or nullptr
  float* Y;               // [R][128]
  int dbg;                // measurement switches (vrpx_debug_ff_fused_flags): bit 0 = chunk converters skip their arithmetic, bit 1 = no weight TMA after the first fill of the rings, bit 2 = converters skip their TMEM loads / stores,
                          // bit 3 = the MMA thread does not wait for the converters (wrong results, shows the issue-bound time)
};

#define VRPX_TMEM_LD32(v, taddr)                                                                                          \
  asm volatile(                                                                                                           \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                           \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                           \
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                           \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),       \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),            \
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),           \
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])                         \
      : "r"(taddr)                                                                                                        \
      : "memory")

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* v) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_ff_fused(const Args a, const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW1h, const __grid_constant__ CUtensorMap mapW1l,
           const __grid_constant__ CUtensorMap mapW2h, const __grid_constant__ CUtensorMap mapW2l) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_xr_full, s_xr_free, s_xa_full, s_xa_free, s_w1_full[2], s_w1_free[2], s_w2_full[2],
      s_w2_free[2], s_d1_full, s_d1_free, s_ha_full, s_ha_free, s_d2_full, s_d2_free;
  __shared__ uint32_t s_tmem;
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t ntiles = (a.R + BM - 1) / BM;

  if (tid == 0) {
    mbar_init(smem_u32(&s_xr_full), 1);
    mbar_init(smem_u32(&s_xr_free), 8);
    mbar_init(smem_u32(&s_xa_full), 8);
    mbar_init(smem_u32(&s_xa_free), 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_w1_full[i]), 1);
      mbar_init(smem_u32(&s_w1_free[i]), 1);
      mbar_init(smem_u32(&s_w2_full[i]), 1);
      mbar_init(smem_u32(&s_w2_free[i]), 1);
    }
    mbar_init(smem_u32(&s_d1_full), 1);
    mbar_init(smem_u32(&s_d1_free), 8);
    mbar_init(smem_u32(&s_ha_full), 8);
    mbar_init(smem_u32(&s_ha_free), 1);
    mbar_init(smem_u32(&s_d2_full), 1);
    mbar_init(smem_u32(&s_d2_free), 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
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
    // ===================== TMA producer: two independent rings =====================
    if (lane == 0) {
      int64_t my_tiles = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) ++my_tiles;
      const uint32_t total = (uint32_t)(my_tiles * NCH);
      uint32_t c1 = 0, c2 = 0, cx = 0;
      while (c1 < total || c2 < total || cx < (uint32_t)my_tiles) {
        if (cx < (uint32_t)my_tiles && mbar_try(smem_u32(&s_xr_free), (cx & 1) ^ 1)) {
          const int row0 = (int)((blockIdx.x + (int64_t)cx * gridDim.x) * BM);
          const uint32_t bar = smem_u32(&s_xr_full);
          mbar_expect_tx(bar, 4 * XBOX);
#pragma unroll
          for (int kq = 0; kq < 4; ++kq) tma_load_2d(smem_u32(smem + SM_X + kq * XBOX), &mapX, kq * 32, row0, bar);
          ++cx;
        }
        if (c1 < total) {
          const uint32_t s = c1 & 1, ph = (c1 >> 1) & 1;
          if (mbar_try(smem_u32(&s_w1_free[s]), ph ^ 1)) {
            const int j = (int)(c1 % NCH);
            unsigned char* dst = smem + SM_W1 + s * W1_SLOT;
            const uint32_t bar = smem_u32(&s_w1_full[s]);
            if ((a.dbg & 2) && c1 >= 2) {
              mbar_arrive(bar);   // measurement: no weight traffic after the first fill (results are wrong)
            } else {
              mbar_expect_tx(bar, W1_SLOT);
              tma_load_2d(smem_u32(dst), &mapW1h, 0, j * HC, bar);
              tma_load_2d(smem_u32(dst + 8192), &mapW1h, 64, j * HC, bar);
              tma_load_2d(smem_u32(dst + 16384), &mapW1l, 0, j * HC, bar);
              tma_load_2d(smem_u32(dst + 24576), &mapW1l, 64, j * HC, bar);
            }
            ++c1;
          }
        }
        if (c2 < total) {
          const uint32_t s = c2 & 1, ph = (c2 >> 1) & 1;
          if (mbar_try(smem_u32(&s_w2_free[s]), ph ^ 1)) {
            const int j = (int)(c2 % NCH);
            unsigned char* dst = smem + SM_W2 + s * W2_SLOT;
            const uint32_t bar = smem_u32(&s_w2_full[s]);
            if ((a.dbg & 2) && c2 >= 2) {
              mbar_arrive(bar);
            } else {
              mbar_expect_tx(bar, W2_SLOT);
              tma_load_2d(smem_u32(dst), &mapW2h, j * HC, 0, bar);
              tma_load_2d(smem_u32(dst + 16384), &mapW2l, j * HC, 0, bar);
            }
            ++c2;
          }
        }
      }
    }
  } else if (warp == W_MMA) {
    // ===================== MMA issuer =====================
    {   // the WHOLE warp runs the issue loop (converged); one elected lane issues each instruction (tc_common.cuh)
      uint32_t cc = 0, ti = 0;   // chunk counter (both products advance it in lock step), tile counter
      auto ff1 = [&](uint32_t c) {   // D1 = X · W1c^T, cross terms first
        const uint32_t s = c & 1, ph = (c >> 1) & 1;
        mbar_wait(smem_u32(&s_w1_full[s]), ph);
        if (!(a.dbg & 8)) mbar_wait(smem_u32(&s_d1_free), (c & 1) ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + TM_D1, xh = tmem + TM_XA, xl = tmem + TM_XA + 64;
        unsigned char* wb = smem + SM_W1 + s * W1_SLOT;
        const uint64_t wh[2] = {make_desc(smem_u32(wb)), make_desc(smem_u32(wb + 8192))};
        const uint64_t wl[2] = {make_desc(smem_u32(wb + 16384)), make_desc(smem_u32(wb + 24576))};
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint64_t o = (uint64_t)(2 * jj);
            const uint32_t ka = 8 * (4 * kh + jj);          // 8 packed TMEM columns per k16 step
            mma_f16_ts_w(d, xl + ka, wh[kh] + o, (kh | jj) ? 1u : 0u, IDESC1);
            mma_f16_ts_w(d, xh + ka, wl[kh] + o, 1u, IDESC1);
          }
#pragma unroll
        for (int kh = 0; kh < 2; ++kh)
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) mma_f16_ts_w(d, xh + 8 * (4 * kh + jj), wh[kh] + (uint64_t)(2 * jj), 1u, IDESC1);
        mma_commit_w(smem_u32(&s_w1_free[s]));
        mma_commit_w(smem_u32(&s_d1_full));
      };
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        mbar_wait(smem_u32(&s_xa_full), ti & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        ff1(cc);
        for (int j = 0; j < NCH; ++j, ++cc) {
          if (j + 1 < NCH) {
            ff1(cc + 1);
            if (j + 2 == NCH) mma_commit_w(smem_u32(&s_xa_free));   // every FF1 product of the tile has been issued
          }
          // FF2(j): D2 += H_j · W2c^T, main and cross terms in separate accumulators
          const uint32_t s = cc & 1, ph = (cc >> 1) & 1;
          mbar_wait(smem_u32(&s_w2_full[s]), ph);
          if (!(a.dbg & 8)) mbar_wait(smem_u32(&s_ha_full), cc & 1);
          if (j == 0) mbar_wait(smem_u32(&s_d2_free), (ti & 1) ^ 1);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = tmem + TM_HA, al = ah + 32;
          unsigned char* wb = smem + SM_W2 + s * W2_SLOT;
          const uint64_t wh = make_desc(smem_u32(wb)), wl = make_desc(smem_u32(wb + 16384));
#pragma unroll
          for (int jj = 0; jj < 4; ++jj) {
            const uint64_t o = (uint64_t)(2 * jj);
            const uint32_t accum = (j | jj) ? 1u : 0u;
            mma_f16_ts_w(tmem + TM_D2X, al + 8 * jj, wh + o, accum, IDESC2);
            mma_f16_ts_w(tmem + TM_D2X, ah + 8 * jj, wl + o, 1u, IDESC2);
            mma_f16_ts_w(tmem + TM_D2M, ah + 8 * jj, wh + o, accum, IDESC2);
          }
          mma_commit_w(smem_u32(&s_w2_free[s]));
          mma_commit_w(smem_u32(&s_ha_free));
        }
        mma_commit_w(smem_u32(&s_d2_full));
      }
    }
  } else if (warp < 8) {
    // ===================== chunk converters: D1 -> relu(. + b1) -> f16 hi/lo A operand in TMEM =====================
    const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
    const int hf = warp >> 2;            // this warp's 32 of the chunk's 64 columns
    uint32_t cc = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int j = 0; j < NCH; ++j, ++cc) {
        mbar_wait(smem_u32(&s_d1_full), cc & 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        uint32_t v0[32];
        if (!(a.dbg & 4)) {
          VRPX_TMEM_LD32(v0, tmem + lane_base + TM_D1 + hf * 32);
          asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c) v0[c] = 0x3c003c00u;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_d1_free));         // the next FF1 may overwrite D1
        uint32_t hi[16], lo[16];
        // the 32 biases of this half chunk: the same addresses for every thread (L1 broadcast; 224 KiB of the 227 KiB of
        // shared memory hold operands, there is no room to stage b1 there)
        const float2* bj = reinterpret_cast<const float2*>(a.b1 + j * HC + hf * 32);
        if (!(a.dbg & 1)) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            const float2 ba = __ldg(bj + c);
            const float x0 = fmaxf(fmaf(__uint_as_float(v0[2 * c]), OUT_SCALE, ba.x), 0.f);
            const float x1 = fmaxf(fmaf(__uint_as_float(v0[2 * c + 1]), OUT_SCALE, ba.y), 0.f);
            split_pair(x0, x1, hi[c], lo[c]);
          }
        } else {
#pragma unroll
          for (int c = 0; c < 16; ++c) hi[c] = lo[c] = v0[c] & 0x3c003c00u;
        }
        mbar_wait(smem_u32(&s_ha_free), (cc & 1) ^ 1);            // FF2 of the previous chunk has consumed HA
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t ha = tmem + lane_base + TM_HA + hf * 16;   // packed: 16 columns per 32 hidden values
        if (!(a.dbg & 4)) {
          tmem_st16(ha, hi);
          tmem_st16(ha + 32, lo);
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_ha_full));
      }
    }
  } else if (warp >= W_EPI0) {
    // ===================== epilogue (the transpose-patch scheme of gemm_tc4.cu) =====================
    const int ew = warp - W_EPI0;
    const int q = warp & 3;              // TMEM lane quarter this warp may touch
    const int chalf = ew >> 2;           // this warp's pair of 32-column chunks
    unsigned char* patch = smem + SM_PATCH + ew * 4096;
    const int lr = lane >> 3, lc = lane & 7;
    // Next tile's X: raw fp32 smem row (two SWIZZLE_128B boxes = 64 floats of this thread's k half) -> f16 hi / lo packed
    // words -> XA in tensor memory (column c holds k = 2c, 2c + 1).  `xi` counts the tiles converted so far.
    const int kh = ew >> 2;
    uint32_t xi = 0;
    auto convert_x = [&]() {
      const int r = q * 32 + lane;
      mbar_wait(smem_u32(&s_xr_full), xi & 1);
      uint32_t hi[32], lo[32];
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 xv = *reinterpret_cast<const float4*>(smem + SM_X + (2 * kh + half) * XBOX + r * 128 + ((c ^ (r & 7)) << 4));
          split_pair(xv.x, xv.y, hi[half * 16 + 2 * c], lo[half * 16 + 2 * c]);
          split_pair(xv.z, xv.w, hi[half * 16 + 2 * c + 1], lo[half * 16 + 2 * c + 1]);
        }
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_xr_free));          // the raw tile may be refilled (next-next tile)
      mbar_wait(smem_u32(&s_xa_free), (xi & 1) ^ 1);              // the previous tile's FF1 products are complete
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t xa = tmem + ((uint32_t)(q * 32) << 16) + TM_XA + kh * 32;
      tmem_st32(xa, hi);
      tmem_st32(xa + 64, lo);
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_xa_full));
      ++xi;
    };
    if ((int64_t)blockIdx.x < ntiles) convert_x();
    uint32_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      if (tile + gridDim.x < ntiles) convert_x();
      const int64_t row_base = tile * BM + q * 32;
      mbar_wait(smem_u32(&s_d2_full), ti & 1);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int cc = 2 * chalf + k;
        uint32_t v[32], w[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + cc * 32;
        VRPX_TMEM_LD32(v, taddr + TM_D2M);
        VRPX_TMEM_LD32(w, taddr + TM_D2X);
        const int c = cc * 32 + lc * 4;   // this lane's 4 columns after the transpose
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(w[i]));
        if (k == 1) {   // both chunks are in registers: the next tile's FF2 may overwrite the accumulators
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&s_d2_free));
        }
        // the (coalesced, L2-resident) residual rows are fetched while the chunk goes through the transpose patch
        float4 res[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = row_base + 4 * i + lr;
          res[i] = (r < a.R) ? *reinterpret_cast<const float4*>(a.residual + r * E + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        __syncwarp();
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
          *reinterpret_cast<uint4*>(patch + lane * 128 + ((g4 ^ (lane & 7)) << 4)) = make_uint4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
        __syncwarp();
        float4 x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + lr;
          x[i] = *reinterpret_cast<const float4*>(patch + rr * 128 + ((lc ^ (rr & 7)) << 4));
        }
        // per-column constants of this lane's 4 columns (L1-resident; kept out of the registers between tiles)
        const float4 bias4 = __ldg(reinterpret_cast<const float4*>(a.b2 + c));
        const float4 sc4 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sh4 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.shift + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float bb[4] = {bias4.x, bias4.y, bias4.z, bias4.w}, ss[4] = {sc4.x, sc4.y, sc4.z, sc4.w},
                    hh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = row_base + 4 * i + lr;
          float y[4] = {x[i].x * OUT_SCALE, x[i].y * OUT_SCALE, x[i].z * OUT_SCALE, x[i].w * OUT_SCALE};
          const float rs[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
#pragma unroll
          for (int e = 0; e < 4; ++e) y[e] = fmaf((y[e] + bb[e]) + rs[e], ss[e], hh[e]);
          if (r < a.R) *reinterpret_cast<float4*>(a.Y + r * E + c) = make_float4(y[0], y[1], y[2], y[3]);
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

}  // namespace ff

int g_ff_dbg = 0;

// Y = (residual + relu(X · W1^T + b1) · W2^T + b2) * scale + shift;  W1 [512][128], W2 [128][512] (torch Linear layouts)
int ff_fused(const float* X, int64_t R, const float* W1, const float* b1, const float* W2, const float* b2,
             const float* residual, const float* scale, const float* shift, float* Y, cudaStream_t stream) {
  using namespace ff;
  if (R <= 0 || !X || !W1 || !b1 || !W2 || !b2 || !residual || !Y || (scale && !shift)) {
    set_error("ff_fused: bad argument");
    return VRPX_ERR_ARG;
  }
  __half* w16 = split_scratch(stream);
  if (!w16) return VRPX_ERR_CUDA;
  constexpr int NW = FF * E;   // 65536 weights per matrix
  int rc;
  if ((rc = split_weights(W1, w16, NW, stream))) return rc;             // W1 hi | W1 lo
  if ((rc = split_weights(W2, w16 + 2 * NW, NW, stream))) return rc;    // W2 hi | W2 lo
  CUtensorMap mx, m1h, m1l, m2h, m2l;
  if ((rc = make_map(&mx, X, R, E, false))) return rc;
  if ((rc = make_map(&m1h, w16, FF, E, true, HC))) return rc;
  if ((rc = make_map(&m1l, w16 + NW, FF, E, true, HC))) return rc;
  if ((rc = make_map(&m2h, w16 + 2 * NW, E, FF, true, E))) return rc;
  if ((rc = make_map(&m2l, w16 + 3 * NW, E, FF, true, E))) return rc;
  static_assert(SMEM_BYTES + 512 <= 227 * 1024, "dynamic + static shared memory must fit the 227 KiB of an SM");
  VRPX_CUDA(cudaFuncSetAttribute(k_ff_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const int64_t ntiles = (R + BM - 1) / BM;
  const int grid = (int)((ntiles < (int64_t)num_sms()) ? ntiles : (int64_t)num_sms());
  Args a{X, R, b1, b2, residual, scale, shift, Y, g_ff_dbg};
  k_ff_fused<<<grid, NTHREADS, SMEM_BYTES, stream>>>(a, mx, m1h, m1l, m2h, m2l);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace vrpx

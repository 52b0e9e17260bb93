// gemm_tn_tc.cu — weight-gradient GEMM  C[M][N] += A^T · B  (A [R][M], B [R][N] row-major fp32, R = B·N rows of the batch)
// on tcgen05: the contraction runs over the ROWS, so both operands arrive "MN-major".  They are transposed on the way
// into shared memory instead: TMA brings raw fp32 boxes of 64 rows, eight converter warps split every value into f16
// hi / lo halves (f16split.cuh arithmetic) and write them — eight consecutive rows of one column per 16-byte store —
// into the K-major SWIZZLE_128B operand tiles the MMA descriptors of gemm_tc4.cu already use.  Per 64-row block and
// 128 x 128 output tile: 12 MMAs (lo·hi + hi·lo + hi·hi over four k16 steps), accumulator in tensor memory.
//   * split-K: the CTAs of one output tile take interleaved 64-row blocks; every CTA drains its accumulator into fp32
//     registers every DRAIN blocks (the tensor core truncates when it accumulates: −1.7e-8 relative per accumulating
//     MMA, gemm_tc4.cu) and adds its partial sum to C with one pass of atomics at the end;
//   * the CTAs of the different output tiles walk the row blocks in step, so a row block comes from DRAM once and from
//     L2 for the other tiles.
// The kernel is HBM-bound by construction (64 KiB of operands per 786 clocks of MMA), which is the point: the mma.sync
// kernel it replaces (k_gemm_tn_mma) ran at 45 TFLOP/s, 87 ms of a 611 ms train step at 65,536 x TSP-50.
// Operand scales: 2^8 on both operands like every f16-split operand of the library (|A|, |B| < 256 keep f16 finite), result * 2^-16.
// Roles (20 warps): 0 TMA producer, 1 MMA issuer, 4-19 converters + accumulator drain (16 warps: the conversion of a block
// is a chain of shared-memory reads, conversions and 16-byte stores that eight warps could not keep busy).
#include "gemm.cuh"
#include "tc_common.cuh"

namespace vrpx {
namespace tn {
using namespace tc4;

constexpr int KB = 64;                      // rows per block
constexpr int NTHREADS = 640;
constexpr int NCONV = 16;                   // converter warps
constexpr int W_TMA = 0, W_MMA = 1, W_CONV0 = 4;
constexpr int RAW_BOX = KB * 128;           // 64 rows x 32 floats = 8 KiB
constexpr int RAW_OPER = 4 * RAW_BOX;       // 128 columns of one operand = 32 KiB
constexpr int RAW_BLOCK = 2 * RAW_OPER;     // A | B = 64 KiB
constexpr int NRAW = 2;                     // raw ring depth (blocks)
constexpr int OPT = 128 * 128;              // one K-major operand tile: 128 rows x 64 halves = 16 KiB
constexpr int SM_RAW = 0;
constexpr int SM_AH = NRAW * RAW_BLOCK, SM_AL = SM_AH + OPT, SM_BH = SM_AL + OPT, SM_BL = SM_BH + OPT;
constexpr int SMEM_BYTES = SM_BL + OPT + 1024;
static_assert(SMEM_BYTES + 512 <= 227 * 1024, "shared memory budget");
constexpr int DRAIN = 8;                    // blocks per accumulator drain: 96 accumulating MMAs
constexpr float A_SCALE = 256.0f, B_SCALE = 256.0f, C_SCALE = 1.0f / (A_SCALE * B_SCALE);
constexpr uint32_t IDESC = make_idesc(128, 128);
constexpr uint32_t TMEM_COLS = 128;

__device__ __forceinline__ void mma_f16_ss_w(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

#define VRPX_TN_LD16(v, taddr)                                                                                      \
  asm volatile(                                                                                                     \
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                                     \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                              \
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), \
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])                    \
      : "r"(taddr)                                                                                                  \
      : "memory")

// one operand of one block: raw [64 rows][128 columns] fp32 (4 SWIZZLE_128B boxes of 32 columns) -> hi / lo tiles
// [128 columns][64 rows] f16, K-major SWIZZLE_128B.  Thread = (column c, row groups rg0, rg0 + 4);
// the lanes of a warp hold consecutive columns: the raw reads and the 16-byte tile stores are conflict free.
// `csum` (optional) collects the sum of the thread's scaled values: the column sums of A are the bias gradient.
__device__ __forceinline__ void convert_operand(const unsigned char* raw, unsigned char* hi, unsigned char* lo, int ct, float scale,
                                                float* csum = nullptr) {
  const int c = ct & 127, rg0 = ct >> 7;
  const unsigned char* rb = raw + (c >> 5) * RAW_BOX + (c & 3) * 4;
  const int ch = (c & 31) >> 2;
  const int trow = (c >> 3) * 1024 + (c & 7) * 128;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int rg = rg0 + 4 * j;
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int row = 8 * rg + i;
      x[i] = *reinterpret_cast<const float*>(rb + row * 128 + ((ch ^ (row & 7)) << 4)) * scale;
    }
    if (csum) *csum += ((x[0] + x[1]) + (x[2] + x[3])) + ((x[4] + x[5]) + (x[6] + x[7]));
    uint32_t h[4], l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 hh = __floats2half2_rn(x[2 * i], x[2 * i + 1]);
      const float2 hf = __half22float2(hh);
      const __half2 ll = __floats2half2_rn(x[2 * i] - hf.x, x[2 * i + 1] - hf.y);
      h[i] = *reinterpret_cast<const uint32_t*>(&hh);
      l[i] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    const int off = trow + ((rg ^ (c & 7)) << 4);
    *reinterpret_cast<uint4*>(hi + off) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(lo + off) = make_uint4(l[0], l[1], l[2], l[3]);
  }
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tn_tc(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, float* __restrict__ C,
             float* __restrict__ colsum, int64_t R, int M, int N) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_raw_full[NRAW], s_raw_free[NRAW], s_conv_full, s_conv_free;
  __shared__ uint32_t s_tmem;
  unsigned char* smem = smem_dyn + ((1024u - (smem_u32(smem_dyn) & 1023u)) & 1023u);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ntn = N / 128, P = (M / 128) * ntn;
  const int pair = blockIdx.x % P, split = blockIdx.x / P, nsplit = gridDim.x / P;
  const int mt = pair / ntn, nt = pair % ntn;
  const int64_t nblk = (R + KB - 1) / KB;
  const int64_t myblk = (nblk > split) ? (nblk - split + nsplit - 1) / nsplit : 0;

  if (tid == 0) {
    for (int i = 0; i < NRAW; ++i) {
      mbar_init(smem_u32(&s_raw_full[i]), 1);
      mbar_init(smem_u32(&s_raw_free[i]), NCONV);
    }
    mbar_init(smem_u32(&s_conv_full), NCONV);
    mbar_init(smem_u32(&s_conv_free), 1);
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
    if (lane == 0) {
      for (int64_t i = 0; i < myblk; ++i) {
        const int rs = (int)(i % NRAW);
        mbar_wait_sleep(smem_u32(&s_raw_free[rs]), (uint32_t)((i / NRAW) & 1) ^ 1, 100);
        const uint32_t bar = smem_u32(&s_raw_full[rs]);
        mbar_expect_tx(bar, RAW_BLOCK);
        const int row0 = (int)((split + i * nsplit) * KB);
        unsigned char* dst = smem + SM_RAW + rs * RAW_BLOCK;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          tma_load_2d(smem_u32(dst + q * RAW_BOX), &mapA, mt * 128 + q * 32, row0, bar);
          tma_load_2d(smem_u32(dst + RAW_OPER + q * RAW_BOX), &mapB, nt * 128 + q * 32, row0, bar);
        }
      }
    }
  } else if (warp == W_MMA) {
    const uint64_t ah = make_desc(smem_u32(smem + SM_AH)), al = make_desc(smem_u32(smem + SM_AL));
    const uint64_t bh = make_desc(smem_u32(smem + SM_BH)), bl = make_desc(smem_u32(smem + SM_BL));
    for (int64_t i = 0; i < myblk; ++i) {
      mbar_wait_sleep(smem_u32(&s_conv_full), (uint32_t)(i & 1), 50);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const bool fresh = (i % DRAIN) == 0;   // the converters drained the accumulator before they converted this block
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint64_t o = (uint64_t)(2 * j);
        mma_f16_ss_w(tmem, al + o, bh + o, (fresh && j == 0) ? 0u : 1u, IDESC);
        mma_f16_ss_w(tmem, ah + o, bl + o, 1u, IDESC);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) mma_f16_ss_w(tmem, ah + (uint64_t)(2 * j), bh + (uint64_t)(2 * j), 1u, IDESC);
      mma_commit_w(smem_u32(&s_conv_free));
    }
  } else if (warp >= W_CONV0) {
    const int cw = warp - W_CONV0, q = warp & 3, cq = cw >> 2;   // TMEM lane quarter, 32-column quarter of the accumulator
    const int ct = tid - W_CONV0 * 32;   // 0..511
    float acc[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) acc[j] = 0.f;
    // column sums of A (= the bias gradient when A is dY): the CTAs of the first column tile see every row of A once
    const bool want_sum = colsum != nullptr && nt == 0;
    float csum = 0.f;
    auto drain = [&]() {
      const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + cq * 32;
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        uint32_t v[16];
        VRPX_TN_LD16(v, taddr + 16 * b);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[16 * b + j] += __uint_as_float(v[j]);
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    };
    for (int64_t i = 0; i < myblk; ++i) {
      const int rs = (int)(i % NRAW);
      mbar_wait(smem_u32(&s_raw_full[rs]), (uint32_t)((i / NRAW) & 1));
      if (i > 0) {
        mbar_wait(smem_u32(&s_conv_free), (uint32_t)((i - 1) & 1));   // the previous block's products are complete
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if ((i % DRAIN) == 0) drain();
      }
      const unsigned char* raw = smem + SM_RAW + rs * RAW_BLOCK;
      convert_operand(raw, smem + SM_AH, smem + SM_AL, ct, A_SCALE, want_sum ? &csum : nullptr);
      convert_operand(raw + RAW_OPER, smem + SM_BH, smem + SM_BL, ct, B_SCALE);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(smem_u32(&s_raw_free[rs]));
        mbar_arrive(smem_u32(&s_conv_full));
      }
    }
    if (myblk > 0) {
      mbar_wait(smem_u32(&s_conv_free), (uint32_t)((myblk - 1) & 1));
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      drain();
      float* crow = C + (size_t)(mt * 128 + q * 32 + lane) * N + nt * 128 + cq * 32;
#pragma unroll
      for (int j = 0; j < 32; ++j) atomicAdd(crow + j, acc[j] * C_SCALE);
    }
    if (want_sum) atomicAdd(colsum + mt * 128 + (ct & 127), csum * (1.0f / A_SCALE));
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

}  // namespace tn

// C [M][N] += A^T · B on tcgen05; M and N multiples of 128, A and B 16-byte aligned row-major [R][M], [R][N];
// colsum_A (optional) [M] += column sums of A
int gemm_tn_tc(const float* A, const float* Bm, float* C, float* colsum_A, int64_t R, int M, int N, cudaStream_t stream) {
  using namespace tn;
  if (M % 128 || N % 128 || R < 1 || R > (int64_t)INT32_MAX) {
    set_error("gemm_tn_tc: bad shape");
    return VRPX_ERR_ARG;
  }
  CUtensorMap ma, mb;
  int rc;
  if ((rc = make_map(&ma, A, R, M, false, KB))) return rc;
  if ((rc = make_map(&mb, Bm, R, N, false, KB))) return rc;
  const int P = (M / 128) * (N / 128);
  const int64_t nblk = (R + KB - 1) / KB;
  int64_t nsplit = num_sms() / P;
  if (nsplit > nblk) nsplit = nblk;
  if (nsplit < 1) nsplit = 1;
  VRPX_CUDA(cudaFuncSetAttribute(k_gemm_tn_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  k_gemm_tn_tc<<<(unsigned)(nsplit * P), NTHREADS, SMEM_BYTES, stream>>>(ma, mb, C, colsum_A, R, M, N);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace vrpx

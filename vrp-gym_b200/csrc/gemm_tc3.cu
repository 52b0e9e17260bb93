// gemm_tc3.cu — persistent, warp-specialised tcgen05 GEMM (3xTF32): TMA-staged tiles, A operand fed from TMEM.
//
// Superseded as the production path by gemm_tc4.cu (f16-split, half the tensor time); kept as the 3xTF32 cross-check.
// The A operand lives in TMEM because with both operands in shared memory the port saturates (measured on the
// generation before: l1tex 60-75 % busy, tensor pipe 12-30 %): a 128x128xK=8 tf32 MMA reads 4 KiB of A and 4 KiB of B from smem every
// 64 cycles, which by itself saturates the 128 B/clk shared-memory port once the converters' traffic is added.  Here
// the converter warps write the hi / lo halves of the X tile straight into TENSOR MEMORY (tcgen05.st, lane = tile row,
// one 32-bit column per k element) and the MMAs take A from TMEM, so shared memory only carries the raw TMA tiles
// and the B (weight) operand.
//
//   Y[R][NOUT] = epilogue( X[R][K] · W[NOUT][K]^T ),  fp32 in / fp32 out, ~fp32 accuracy (precision
//   policy: hi = x & 0xffffe000, lo = x - hi; D = Xlo·Whi + Xhi·Wlo + Xhi·Whi on tcgen05.mma kind::tf32, fp32 in TMEM).
//
// One CTA per SM loops over 128x128 output tiles (column tile fastest, so the CTAs that share an X row tile run
// side by side and hit L2).  Roles (12 warps):
//   warp 4      TMA producer: cp.async.bulk.tensor.2d (SWIZZLE_128B boxes of 32 floats x 128 rows) for X and W
//   warps 0-3   converters: split the raw fp32 tiles in place into hi (tf32-exact) + a second lo tile
//   warp 5      MMA issuer: 12 tcgen05.mma per 32-wide k-block, tcgen05.commit releases the smem stage
//   warps 8-15  epilogue (two warps per TMEM lane quarter, two 32-column chunks each): tcgen05.ld from one of TWO
//               TMEM accumulators while the other is being filled
// Hand-offs are mbarriers: raw_full (TMA tx bytes) -> conv_done -> stage_free (commit), acc_full / acc_free.
#include <cuda.h>

#include "gemm.cuh"

namespace vrpx {
namespace tc3 {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int STAGES = 4;
constexpr int NTHREADS = 512;
constexpr int TILE_BYTES = BM * BK * 4;        // 16 KiB
constexpr int STAGE_BYTES = 3 * TILE_BYTES;    // X raw | W raw/hi | W lo   (X hi/lo live in TMEM)
constexpr uint32_t TMEM_A0 = 2 * BN;            // TMEM columns: 2 accumulators, then STAGES x (32 hi + 32 lo) A columns
constexpr uint32_t TMEM_COLS = 512;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 8 * 4096 + 1024;   // stages | 8 epilogue transpose patches | alignment slack
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, SBO 1024 B
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// hi/lo split of one 16 KiB tile by 128 threads; element-wise, so the TMA swizzle is preserved.
__device__ __forceinline__ void split_tile(unsigned char* raw_hi, unsigned char* lo, int t128) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int off = (t128 + 128 * i) * 16;
    float4 v = *reinterpret_cast<const float4*>(raw_hi + off);
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    *reinterpret_cast<float4*>(raw_hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

// Row r (= converter thread) of the raw X tile: 8 swizzled 16-byte chunks -> 32 hi + 32 lo words in TMEM lane r.
__device__ __forceinline__ void x_tile_to_tmem(const unsigned char* raw, int r, uint32_t taddr_hi) {
  uint32_t hi[32], lo[32];
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const float4 v = *reinterpret_cast<const float4*>(raw + r * 128 + ((c ^ (r & 7)) << 4));
    const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t h = __float_as_uint(e[i]) & 0xffffe000u;
      hi[4 * c + i] = h;
      lo[4 * c + i] = __float_as_uint(e[i] - __uint_as_float(h));
    }
  }
  tmem_st32(taddr_hi, hi);
  tmem_st32(taddr_hi + 32, lo);
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(NTHREADS, 1)
k_gemm_tc3(GemmArgs a, const __grid_constant__ CUtensorMap mapX, const __grid_constant__ CUtensorMap mapW) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_raw_full[STAGES], s_conv_done[STAGES], s_stage_free[STAGES], s_acc_full[2], s_acc_free[2];
  __shared__ uint32_t s_tmem;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nct = a.NOUT / BN;
  const int64_t nrt = (a.R + BM - 1) / BM;
  const int64_t ntiles = nrt * nct;
  const int nkb = a.K / BK;

  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(smem_u32(&s_raw_full[i]), 1);
      mbar_init(smem_u32(&s_conv_done[i]), 4);
      mbar_init(smem_u32(&s_stage_free[i]), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(smem_u32(&s_acc_full[i]), 1);
      mbar_init(smem_u32(&s_acc_free[i]), 8);
    }
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

  if (warp == 4) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t kbc = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int row0 = (int)(tile / nct) * BM, col0 = (int)(tile % nct) * BN;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const uint32_t s = kbc % STAGES, ph = (kbc / STAGES) & 1;
          mbar_wait(smem_u32(&s_stage_free[s]), ph ^ 1);
          unsigned char* st = smem + s * STAGE_BYTES;
          const uint32_t bar = smem_u32(&s_raw_full[s]);
          mbar_expect_tx(bar, 2 * TILE_BYTES);
          tma_load_2d(smem_u32(st), &mapX, kb * BK, row0, bar);
          tma_load_2d(smem_u32(st + TILE_BYTES), &mapW, kb * BK, col0, bar);
        }
      }
    }
  } else if (warp < 4) {
    // ===================== converters =====================
    uint32_t kbc = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int kb = 0; kb < nkb; ++kb, ++kbc) {
        const uint32_t s = kbc % STAGES, ph = (kbc / STAGES) & 1;
        unsigned char* st = smem + s * STAGE_BYTES;
        mbar_wait(smem_u32(&s_raw_full[s]), ph);
        x_tile_to_tmem(st, tid, tmem + ((uint32_t)(warp * 32) << 16) + TMEM_A0 + s * 64);
        split_tile(st + TILE_BYTES, st + 2 * TILE_BYTES, tid);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&s_conv_done[s]));
      }
    }
  } else if (warp == 5) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t kbc = 0, ti = 0;
      for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
        mbar_wait(smem_u32(&s_acc_free[acc]), aph ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t d = tmem + acc * BN;
        for (int kb = 0; kb < nkb; ++kb, ++kbc) {
          const uint32_t s = kbc % STAGES, ph = (kbc / STAGES) & 1;
          unsigned char* st = smem + s * STAGE_BYTES;
          mbar_wait(smem_u32(&s_conv_done[s]), ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t ah = tmem + TMEM_A0 + s * 64, alo = ah + 32;
          const uint64_t wh = make_desc(smem_u32(st + TILE_BYTES)), wl = make_desc(smem_u32(st + 2 * TILE_BYTES));
#pragma unroll
          for (int j = 0; j < BK / 8; ++j) {
            const uint64_t o = (uint64_t)(2 * j);
            mma_tf32_ts(d, alo + 8 * j, wh + o, (kb | j) ? 1u : 0u);
            mma_tf32_ts(d, ah + 8 * j, wl + o, 1u);
            mma_tf32_ts(d, ah + 8 * j, wh + o, 1u);
          }
          mma_commit(smem_u32(&s_stage_free[s]));
        }
        mma_commit(smem_u32(&s_acc_full[acc]));
      }
    }
  } else if (warp >= 8) {
    // ===================== epilogue =====================
    // tcgen05.ld hands every thread one tile ROW (32 consecutive columns).  Storing that directly would touch 32
    // different 128-byte lines per instruction, so each 32x32 chunk is transposed through a 4 KiB swizzled smem patch:
    // afterwards lane l owns columns 4(l&7)..+3 of rows 4i + (l>>3), i = 0..7, and every global load / store
    // instruction of the warp covers 4 rows x 128 contiguous bytes.
    const int q = warp & 3;  // TMEM lane quarter this warp may touch
    const int chalf = (warp - 8) >> 2;   // this warp's pair of 32-column chunks
    unsigned char* patch = smem + STAGES * STAGE_BYTES + (warp - 8) * 4096;
    const int lr = lane >> 3, lc = lane & 7;
    uint32_t ti = 0;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++ti) {
      const uint32_t acc = ti & 1, aph = (ti >> 1) & 1;
      const int64_t row_base = (tile / nct) * BM + q * 32;
      const int col0 = (int)(tile % nct) * BN;
      mbar_wait(smem_u32(&s_acc_full[acc]), aph);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cc = 2 * chalf; cc < 2 * chalf + 2; ++cc) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + acc * BN + cc * 32;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
              "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
              "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
              "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
            : "r"(taddr)
            : "memory");
        const int c = col0 + cc * 32 + lc * 4;   // this lane's 4 columns after the transpose
        const float4 bias4 = a.bias ? __ldg(reinterpret_cast<const float4*>(a.bias + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 sc4 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.scale + c)) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float4 sh4 = a.scale ? __ldg(reinterpret_cast<const float4*>(a.shift + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
        // the (coalesced) residual rows are fetched while the TMEM load is in flight
        float4 res[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int64_t r = row_base + 4 * i + lr;
          res[i] = (a.residual && r < a.R) ? *reinterpret_cast<const float4*>(a.residual + r * a.NOUT + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        __syncwarp();  // the previous chunk's reads of the patch are complete
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4)
          *reinterpret_cast<uint4*>(patch + lane * 128 + ((g4 ^ (lane & 7)) << 4)) = make_uint4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + lr;
          const int64_t r = row_base + rr;
          const float4 x = *reinterpret_cast<const float4*>(patch + rr * 128 + ((lc ^ (rr & 7)) << 4));
          float y[4] = {x.x, x.y, x.z, x.w};
          const float4 gt = (a.gate && r < a.R) ? *reinterpret_cast<const float4*>(a.gate + r * a.NOUT + c) : make_float4(1.f, 1.f, 1.f, 1.f);
          const float gg[4] = {gt.x, gt.y, gt.z, gt.w}, rs[4] = {res[i].x, res[i].y, res[i].z, res[i].w};
          const float bb[4] = {bias4.x, bias4.y, bias4.z, bias4.w}, ss[4] = {sc4.x, sc4.y, sc4.z, sc4.w},
                      hh[4] = {sh4.x, sh4.y, sh4.z, sh4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            if (!(gg[e] > 0.f)) y[e] = 0.f;
            y[e] += bb[e];
            if (a.relu) y[e] = fmaxf(y[e], 0.f);
            y[e] += rs[e];
            if (a.scale) y[e] = fmaf(y[e], ss[e], hh[e]);
          }
          if (r < a.R) *reinterpret_cast<float4*>(a.Y + r * a.NOUT + c) = make_float4(y[0], y[1], y[2], y[3]);
        }
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(&s_acc_free[acc]));
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS) : "memory");
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D row-major fp32 matrix [rows][cols] -> boxes of 32 columns x 128 rows, SWIZZLE_128B, zero fill out of bounds
static int make_map(CUtensorMap* m, const float* base, int64_t rows, int cols) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available");
    return VRPX_ERR_CUDA;
  }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)BM};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d", (int)r, (long long)rows, cols);
    return VRPX_ERR_CUDA;
  }
  return VRPX_OK;
}

}  // namespace tc3

int gemm_tc_tf32(const GemmArgs& a, cudaStream_t stream) {  // previous production path, kept as a cross-check
  using namespace tc3;
  if (a.K % BK != 0 || a.NOUT % BN != 0 || a.R <= 0) {
    set_error("gemm_tc: unsupported shape R=%lld K=%d NOUT=%d", (long long)a.R, a.K, a.NOUT);
    return VRPX_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(a.X) & 15) || (reinterpret_cast<uintptr_t>(a.W) & 15)) {
    set_error("gemm_tc: operands must be 16-byte aligned");
    return VRPX_ERR_ARG;
  }
  CUtensorMap mx, mw;
  int rc;
  if ((rc = make_map(&mx, a.X, a.R, a.K))) return rc;
  if ((rc = make_map(&mw, a.W, a.NOUT, a.K))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    VRPX_CUDA(cudaFuncSetAttribute(k_gemm_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    attr_set = true;
  }
  const int64_t ntiles = ((a.R + BM - 1) / BM) * (a.NOUT / BN);
  const int grid = (int)((ntiles < (int64_t)num_sms()) ? ntiles : (int64_t)num_sms());
  k_gemm_tc3<<<grid, NTHREADS, SMEM_BYTES, stream>>>(a, mx, mw);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace vrpx

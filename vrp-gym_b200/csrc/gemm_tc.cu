// gemm_tc.cu — first-generation (non-persistent, thread-staged) tcgen05 GEMM, kept as cross-check path 2; the
// production path is gemm_tc2.cu.  fp32-accurate GEMM on the 5th-gen tensor cores (tcgen05.mma kind::tf32, accumulators in
// TMEM) for the encoder's dense layers (agents/graph_encoder.py:170-181: in_proj / out_proj / FF).
//
// Precision policy: the spec asks <= 1e-5 relative on logits, single-pass TF32 gives ~1e-3.  Each fp32
// operand is split in-kernel into hi = x & 0xffffe000 (exactly a TF32 number) and lo = x - hi, and
//   D = Xlo·Whi + Xhi·Wlo + Xhi·Whi          ("3xTF32", error ~2^-21 per product)
// is accumulated in fp32 in TMEM.
//
// Tile: 128 rows x 128 columns per CTA, K streamed in blocks of 32 floats (one 128-byte swizzle atom),
// two smem stages.  All 256 threads stage + split the operands into the canonical K-major SWIZZLE_128B
// layout, one elected thread issues the MMAs, completion is tracked with tcgen05.commit -> mbarrier.
// Epilogue: tcgen05.ld 32x32b (thread = tile row), bias / ReLU / residual / folded BatchNorm, fp32 store.
#include "gemm.cuh"

namespace vrpx {
namespace tc {

constexpr int BM = 128, BN = 128, BK = 32;
constexpr int NTHREADS = 256;
constexpr int STAGES = 2;
constexpr int TILE_BYTES = BM * BK * 4;             // 16 KiB: one operand tile (128 rows x 128 B)
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // Xhi, Xlo, Whi, Wlo
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024;  // + alignment slack

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor):
//   [0,14) start address >> 4, [16,30) LBO >> 4 (=1, unused for swizzled K-major), [32,46) SBO >> 4
//   (= 1024 B between 8-row groups), [46,48) version = 1, [61,64) layout type = 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format TF32 (2) @7/@10,
// both K-major, N >> 3 @17, M >> 4 @24.
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(IDESC), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// Stage one 128 x 32 fp32 operand tile: coalesced float4 loads, hi/lo split, swizzled stores.
// 16-byte chunk c of row r lands at r*128 + ((c ^ (r & 7)) << 4)   (Swizzle<3,4,3>).
__device__ __forceinline__ void stage_tile(const float* __restrict__ src, int64_t row0, int64_t nrows, int ld,
                                           int k0, unsigned char* hi, unsigned char* lo, int tid) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int idx = tid + NTHREADS * i;
    int r = idx >> 3, c = idx & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (row0 + r < nrows) v = __ldg(reinterpret_cast<const float4*>(src + (row0 + r) * ld + k0 + c * 4));
    float4 h, l;
    h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
    h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
    h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
    h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
    l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
    int off = r * 128 + ((c ^ (r & 7)) << 4);
    *reinterpret_cast<float4*>(hi + off) = h;
    *reinterpret_cast<float4*>(lo + off) = l;
  }
}

__global__ void __launch_bounds__(NTHREADS, 1) k_gemm_tc(GemmArgs a) {
  extern __shared__ unsigned char smem_dyn[];
  __shared__ __align__(8) uint64_t s_bar[STAGES + 1];
  __shared__ uint32_t s_tmem;
  unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nct = a.NOUT / BN;                       // column tiles vary fastest so the CTAs sharing an
  const int64_t row0 = (int64_t)(blockIdx.x / nct) * BM;  // X row tile run back to back (L2 reuse)
  const int col0 = (int)(blockIdx.x % nct) * BN;

  if (tid == 0) {
    for (int i = 0; i <= STAGES; ++i) mbar_init(smem_u32(&s_bar[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = s_tmem;

  const int nkb = a.K / BK;
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb % STAGES;
    unsigned char* st = smem + s * STAGE_BYTES;
    if (kb >= STAGES) mbar_wait(smem_u32(&s_bar[s]), ((kb / STAGES) - 1) & 1);  // MMAs of k-block kb-2 drained
    stage_tile(a.X, row0, a.R, a.K, kb * BK, st, st + TILE_BYTES, tid);
    stage_tile(a.W, col0, a.NOUT, a.K, kb * BK, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES, tid);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy stores -> async proxy (UMMA)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint64_t xh = make_desc(smem_u32(st)), xl = make_desc(smem_u32(st + TILE_BYTES));
      const uint64_t wh = make_desc(smem_u32(st + 2 * TILE_BYTES)), wl = make_desc(smem_u32(st + 3 * TILE_BYTES));
#pragma unroll
      for (int j = 0; j < BK / 8; ++j) {       // UMMA_K = 8 tf32 = 32 bytes -> start address += 2 (16-byte units)
        const uint64_t o = (uint64_t)(2 * j);
        mma_tf32(tmem, xl + o, wh + o, (kb | j) ? 1u : 0u);
        mma_tf32(tmem, xh + o, wl + o, 1u);
        mma_tf32(tmem, xh + o, wh + o, 1u);
      }
      mma_commit(smem_u32(&s_bar[s]));
      if (kb == nkb - 1) mma_commit(smem_u32(&s_bar[STAGES]));
    }
  }
  // ---- epilogue: all MMAs complete -> TMEM -> registers -> global
  mbar_wait(smem_u32(&s_bar[STAGES]), 0);
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int r_in_tile = (warp & 3) * 32 + lane;   // a warp may only touch TMEM lanes 32*(warp%4)..+31
  const int64_t r = row0 + r_in_tile;
#pragma unroll 1
  for (int cc = 0; cc < 2; ++cc) {
    const int cbase = (warp >> 2) * 64 + cc * 32;
    uint32_t v[32];
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)cbase;
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (r < a.R) {
      const int c0 = col0 + cbase;
      float* yrow = a.Y + r * a.NOUT + c0;
      const float* rrow = a.residual ? a.residual + r * a.NOUT + c0 : nullptr;
      const float* grow = a.gate ? a.gate + r * a.NOUT + c0 : nullptr;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float o[4];
        float4 res = rrow ? *reinterpret_cast<const float4*>(rrow + 4 * q) : make_float4(0.f, 0.f, 0.f, 0.f);
        const float rr[4] = {res.x, res.y, res.z, res.w};
        float4 gt = grow ? *reinterpret_cast<const float4*>(grow + 4 * q) : make_float4(1.f, 1.f, 1.f, 1.f);
        const float gg[4] = {gt.x, gt.y, gt.z, gt.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c0 + 4 * q + e;
          float y = __uint_as_float(v[4 * q + e]);
          if (!(gg[e] > 0.f)) y = 0.f;
          if (a.bias) y += __ldg(a.bias + c);
          if (a.relu) y = fmaxf(y, 0.f);
          y += rr[e];
          if (a.scale) y = fmaf(y, __ldg(a.scale + c), __ldg(a.shift + c));
          o[e] = y;
        }
        *reinterpret_cast<float4*>(yrow + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN) : "memory");
  }
}

}  // namespace tc

int gemm_tc_v1(const GemmArgs& a, cudaStream_t stream) {
  if (a.K % tc::BK != 0 || a.NOUT % tc::BN != 0 || a.R <= 0) {
    set_error("gemm_tc: unsupported shape R=%lld K=%d NOUT=%d", (long long)a.R, a.K, a.NOUT);
    return VRPX_ERR_ARG;
  }
  static bool attr_set = false;
  if (!attr_set) {
    VRPX_CUDA(cudaFuncSetAttribute(tc::k_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    attr_set = true;
  }
  dim3 grid((unsigned)(((a.R + tc::BM - 1) / tc::BM) * (a.NOUT / tc::BN)));
  tc::k_gemm_tc<<<grid, tc::NTHREADS, tc::SMEM_BYTES, stream>>>(a);
  VRPX_LAUNCH_CHECK();
  return VRPX_OK;
}

}  // namespace vrpx

// Test hook: Y = X · W^T through either path (tests/test_gemm_gpu.py).
extern "C" int vrpx_debug_gemm(const float* X, int64_t R, int32_t K, const float* W, int32_t NOUT,
                               const float* bias, int32_t relu, const float* residual, const float* scale,
                               const float* shift, float* Y, int32_t path, void* stream) {
  vrpx::GemmArgs g{X, R, K, W, NOUT, bias, relu, residual, scale, shift, Y};
  if (path == 0) return vrpx::gemm_tc(g, (cudaStream_t)stream);
  if (path == 2) return vrpx::gemm_tc_v1(g, (cudaStream_t)stream);
  if (path == 3) return vrpx::gemm_tc_v2(g, (cudaStream_t)stream);
  return vrpx::gemm_simt(g, (cudaStream_t)stream);
}

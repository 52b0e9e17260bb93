// tc_common.cuh — tcgen05 / TMA / mbarrier building blocks shared by the tensor-core kernels (gemm_tc4.cu, ff_fused.cu):
// PTX wrappers, the f16 hi/lo operand split, tensor-map encoding and the per-stream scratch for split weights.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

#include <map>
#include <mutex>
#include <utility>

#include "common.cuh"

namespace vrpx {
namespace tc4 {

// operand scaling of the f16 hi/lo split: W and X are multiplied by 2^8 before the split so that their lo halves stay
// normal f16 numbers; the epilogue multiplies the accumulator by 2^-16
constexpr float W_SCALE = 256.0f, X_SCALE = 256.0f, OUT_SCALE = 1.0f / (W_SCALE * X_SCALE);

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
// the same for a role that waits much longer than it works (MMA issuer, producers): sleep between polls so that the spin
// does not take issue slots from the warps that share its scheduler
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity, unsigned ns) {
  uint32_t ok;
  for (;;) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) break;
    __nanosleep(ns);
  }
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
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {  // K-major SWIZZLE_128B, SBO 1024 B (gemm_tc.cu)
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3fff);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// D[tmem] (+)= A[tmem] · B[smem]^T   (A operand from tensor memory)
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] · B[smem]^T   (both operands through shared-memory descriptors)
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Warp-collective variants: EVERY lane of a converged warp executes the call, one elected lane issues the instruction.
// Issued from inside an `if (lane == 0)` region the compiler cannot prove the operands warp-uniform and wraps each
// UTCHMMA in an ELECT / R2UR / BRA.U.ANY loop (~8 dependent instructions, 40-60 clocks per MMA): a kernel whose MMAs take
// 32-64 clocks each then runs at half the tensor rate, bound by the issuing thread (measured: ff_fused.cu, 49 % tensor
// active).  With the whole warp converged the descriptor arithmetic stays in uniform registers and the MMA is one
// predicated instruction.
__device__ __forceinline__ void mma_f16_ts_w(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t accumulate, uint32_t idesc) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma_commit_w(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}\n" ::"r"(bar)
      : "memory");
}

// kind::f16 instruction descriptor: D = f32 (bit 4), A = B = f16 (format 0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
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

// {f16(x0) f16(x1)} and the f16 pair of the remainders
__device__ __forceinline__ void split_pair(float x0, float x1, uint32_t& hi, uint32_t& lo) {
  x0 *= X_SCALE;
  x1 *= X_SCALE;
  const __half2 h = __floats2half2_rn(x0, x1);
  const float2 hf = __half22float2(h);
  const __half2 l = __floats2half2_rn(x0 - hf.x, x1 - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}


// W [n] f32 -> f16 hi | lo of W * 2^8 at w16 and w16 + n (gemm_tc4.cu)
int split_weights(const float* W, __half* w16, int n, cudaStream_t stream);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
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

// 2-D row-major matrix [rows][cols] -> boxes of 128 bytes x 128 rows (32 floats or 64 halves), SWIZZLE_128B, zero fill
inline int make_map(CUtensorMap* m, const void* base, int64_t rows, int cols, bool f16, int box_rows = 128) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("gemm_tc: cuTensorMapEncodeTiled entry point not available");
    return VRPX_ERR_CUDA;
  }
  const size_t es = f16 ? 2 : 4;
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)cols * es};
  cuuint32_t box[2] = {(cuuint32_t)(128 / es), (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims,
                  strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm_tc: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%d", (int)r, (long long)rows, cols);
    return VRPX_ERR_CUDA;
  }
  return VRPX_OK;
}

// Scratch for the split weights: one 4 MiB buffer per (device, stream), allocated on first use and kept for the life of
// the process.  Calls on one stream are ordered (the split kernel of call n+1 runs after the GEMM of call n), calls on
// different streams get different buffers.  (cudaMallocAsync per call was measured 10x slower end to end: the default
// pool returns its memory at every synchronisation.)
constexpr int kMaxSplitWeights = 1 << 20;   // halves per half (hi | lo): room for two 512 x 1024 matrices
inline __half* split_scratch(cudaStream_t stream) {
  static std::mutex mu;
  static std::map<std::pair<int, cudaStream_t>, __half*> cache;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(dev, stream);
  auto it = cache.find(key);
  if (it != cache.end()) return it->second;
  void* p = nullptr;
  if (cudaMalloc(&p, (size_t)2 * kMaxSplitWeights * sizeof(__half)) != cudaSuccess) {
    set_error("gemm_tc: cudaMalloc of the weight-split scratch failed");
    return nullptr;
  }
  cache[key] = static_cast<__half*>(p);
  return static_cast<__half*>(p);
}


}  // namespace tc4
}  // namespace vrpx

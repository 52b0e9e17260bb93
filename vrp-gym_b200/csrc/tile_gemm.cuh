// tile_gemm.cuh — CTA-level building blocks shared by the persistent rollout kernel (rollout.cu) and the
// recompute-based decoder backward (decoder_bwd.cu): a tile of TM = 32 instances, 512 threads, fp32 FFMA
// register tiles, shared weights streamed from L2 through a cp.async double buffer.
#pragma once
#include "env_rules.cuh"

namespace vrpx {

constexpr int TM = 32;        // instances per tile
constexpr int NT = 512;       // threads per CTA (16 warps; <= 128 registers per thread)
constexpr int QW = NH * E;    // 1024: per-instance width of q~ / c
constexpr size_t SMEM_X = (size_t)TM * E * sizeof(float);    // 16 KiB
constexpr size_t SMEM_QC = (size_t)TM * QW * sizeof(float);  // 128 KiB
constexpr int WCHUNK_FLOATS = 16 * 512;                      // one staged weight chunk: 32 KiB
constexpr size_t SMEM_W = 2 * (size_t)WCHUNK_FLOATS * sizeof(float);  // double buffer, 64 KiB

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

// glimpse-mask source row for attention row (b, hh): mask.repeat(H,1) indexing (graph_decoder.py:93)
__device__ __forceinline__ int64_t quirk_row(int64_t b, int hh, long long G) {
  if (G <= 0) return b;
  int64_t g0 = (b / G) * G;
  return g0 + (((b - g0) * NH + hh) % G);
}

// ---------------------------------------------------------------- wide GEMM: [TM x 128] · [128 x 1024]
// Xs smem [TM][128]; Wt global [128][1024], streamed in chunks of 16 k-rows x 512 columns (32 KiB).
// 512 threads = 4 row groups x 128 column threads; each thread owns 8 rows x 4 columns {4tx..4tx+3} of the current
// 512-column half (conflict-free LDS.128).  epi(m, c, v): row m of the tile, columns c..c+3.
__device__ __forceinline__ void stage_wide_chunk(const float* __restrict__ Wt, int chunk, float* __restrict__ dst) {
  const int half = chunk >> 3, k0 = (chunk & 7) * 16;
  const float* src = Wt + (size_t)k0 * QW + half * 512;
#pragma unroll
  for (int i = 0; i < 2048 / NT; ++i) {
    int idx = threadIdx.x + NT * i;       // 2048 float4 per chunk
    int r = idx >> 7, c4 = idx & 127;
    cp_async16(dst + r * 512 + c4 * 4, src + (size_t)r * QW + c4 * 4);
  }
}

template <class Epi>
__device__ __forceinline__ void tile_gemm_wide(const float* __restrict__ Xs, const float* __restrict__ Wt,
                                               float* __restrict__ Wb, Epi epi) {
  const int tid = threadIdx.x, ty = tid >> 7, tx = tid & 127;
  stage_wide_chunk(Wt, 0, Wb);
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
      stage_wide_chunk(Wt, chunk + 1, Wb + ((chunk + 1) & 1) * WCHUNK_FLOATS);
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
#pragma unroll
      for (int i = 0; i < 8; ++i) epi(ty * 8 + i, c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]));
    }
  }
}

// ---------------------------------------------------------------- tall GEMM: [TM x 1024] · [1024 x 128]
// Cs smem [TM][1024]; Mt global [1024][128] staged with cp.async: chunk kc = rows {kg*256 + kc*16 + r} of the four
// k-groups (4 x 16 rows x 128 columns = 32 KiB).  512 threads = 4 k-groups x 4 row groups x 32 column threads,
// 8 rows x 4 columns {4tx..+3} each over a quarter of K; partial sums reduced through `part` (smem, 4*TM*128 floats,
// may alias Cs).  out[m][e] = sum + bias[e] (bias may be null) is written to smem `out` [TM][128].
__device__ __forceinline__ void stage_tall_chunk(const float* __restrict__ Mt, int kc, float* __restrict__ dst) {
#pragma unroll
  for (int i = 0; i < 2048 / NT; ++i) {
    int idx = threadIdx.x + NT * i;       // 2048 float4 per chunk
    int row = idx >> 5, c4 = idx & 31;    // row in [0,64): kg = row >> 4, r = row & 15
    int k = (row >> 4) * 256 + kc * 16 + (row & 15);
    cp_async16(dst + row * E + c4 * 4, Mt + (size_t)k * E + c4 * 4);
  }
}

__device__ __forceinline__ void tile_gemm_tall(float* __restrict__ Cs, const float* __restrict__ Mt,
                                               float* __restrict__ Wb, const float* __restrict__ bias,
                                               float* __restrict__ part, float* __restrict__ out) {
  const int tid = threadIdx.x, kg = tid >> 7, ty = (tid >> 5) & 3, tx = tid & 31;
  float acc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  stage_tall_chunk(Mt, 0, Wb);
  cp_async_commit();
  for (int kc = 0; kc < 16; ++kc) {
    if (kc + 1 < 16) {
      stage_tall_chunk(Mt, kc + 1, Wb + ((kc + 1) & 1) * WCHUNK_FLOATS);
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
      for (int i = 0; i < 8; ++i) xv[i] = *reinterpret_cast<const float4*>(Cs + (ty * 8 + i) * QW + k0 + kq);
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
    __syncthreads();  // also orders the last reads of Cs before `part` (which may alias it) is written
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(part + (kg * TM + ty * 8 + i) * E + tx * 4) =
        make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  __syncthreads();
  for (int o = tid; o < TM * E; o += NT) {
    float s = part[o] + part[TM * E + o] + part[2 * TM * E + o] + part[3 * TM * E + o];
    out[o] = s + (bias ? bias[o & (E - 1)] : 0.f);
  }
  __syncthreads();
}

// ================================================================ tensor-pipe variants (rollout kernel)
// The same two tile GEMMs on the legacy warp-level tensor path: mma.sync.m16n8k8 TF32 with the 3-term split
// (acc += Alo·Bhi + Ahi·Blo + Ahi·Bhi, ~fp32 accuracy).  Measured on B200 (tools/mma_bench.cu): 480 MAC/clk/SM for
// mma.sync TF32 vs 123 MAC/clk/SM for FFMA, and the tensor pipe runs beside the FMA pipe.  tcgen05 needs M >= 64
// rows per instruction and operands in the UMMA smem layout; a 32-instance tile whose 1024-wide result must stay in
// shared memory for the per-instance phases does not fit that shape, so these tile GEMMs use register fragments.
// Shared-memory leading dimensions are padded so that fragment loads are bank-conflict free:
constexpr int XS_LD = E + 4;        // 132: A fragments (row g, col t): bank (4g + t) % 32
constexpr int QC_LD = QW + 4;       // 1028
constexpr int WARP_LD = 64 + 8;     // 72: per-warp weight stage [8 k-rows][64 columns], B fragments (k = t, n = g): bank (8t + g) % 32
constexpr int WARP_STAGE_FLOATS = 8 * WARP_LD;            // 576 floats = 2304 B: one mma k-step of a warp's 64 columns
template <int MT> constexpr size_t smem_x_mma() { return (size_t)(16 * MT) * XS_LD * sizeof(float); }
template <int MT> constexpr size_t smem_qc_mma() { return (size_t)(16 * MT) * QC_LD * sizeof(float); }
constexpr size_t SMEM_W_MMA = (size_t)(NT / 32) * 2 * WARP_STAGE_FLOATS * sizeof(float);   // 16 warps x 2 stages = 72 KiB

__device__ __forceinline__ void split_tf32(float x, uint32_t& hi, uint32_t& lo) {
  hi = __float_as_uint(x) & 0xffffe000u;
  lo = __float_as_uint(x - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragments (hi and lo) of the two 16-row tiles for one k-step, from a row-major smem matrix with leading dim ld.
template <int MT>
__device__ __forceinline__ void load_a_frags(const float* __restrict__ A, int ld, int k0, int g, int t,
                                             uint32_t (&ah)[MT][4], uint32_t (&al)[MT][4]) {
#pragma unroll
  for (int mt = 0; mt < MT; ++mt) {
    const float* r0 = A + (mt * 16 + g) * ld + k0 + t;
    const float* r1 = r0 + 8 * ld;
    split_tf32(r0[0], ah[mt][0], al[mt][0]);
    split_tf32(r1[0], ah[mt][1], al[mt][1]);
    split_tf32(r0[4], ah[mt][2], al[mt][2]);
    split_tf32(r1[4], ah[mt][3], al[mt][3]);
  }
}

// Each warp streams ITS OWN 64 weight columns (8 k-rows = one mma k-step per stage, double buffered with cp.async), so
// the k loop needs no CTA-wide barrier — only __syncwarp.  src points at element (k = 0, first column of the warp).
__device__ __forceinline__ void stage_warp_kstep(const float* __restrict__ src, int ld, int kstep, float* __restrict__ dst,
                                                 int lane) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + 32 * i;        // 8 rows x 16 float4
    const int r = idx >> 4, c4 = idx & 15;
    cp_async16(dst + r * WARP_LD + c4 * 4, src + (size_t)(kstep * 8 + r) * ld + c4 * 4);
  }
}

// acc[MT][8][4] += A[16 MT][K = 8*nks] (smem, leading dim lda, k offset k_base) · W[k][64 columns of this warp]
template <int MT>
__device__ __forceinline__ void warp_gemm_mma(float (&acc)[MT][8][4], const float* __restrict__ A, int lda, int k_base,
                                              const float* __restrict__ wsrc, int wld, int nks, float* __restrict__ wbuf,
                                              int lane) {
  const int g = lane >> 2, t = lane & 3;
  stage_warp_kstep(wsrc, wld, 0, wbuf, lane);
  cp_async_commit();
  for (int ks = 0; ks < nks; ++ks) {
    if (ks + 1 < nks) {
      stage_warp_kstep(wsrc, wld, ks + 1, wbuf + ((ks + 1) & 1) * WARP_STAGE_FLOATS, lane);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    uint32_t ah[MT][4], al[MT][4];
    load_a_frags<MT>(A, lda, k_base + ks * 8, g, t, ah, al);
    const float* wb = wbuf + (ks & 1) * WARP_STAGE_FLOATS + g;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(wb[t * WARP_LD + 8 * j], bh0, bl0);
      split_tf32(wb[(t + 4) * WARP_LD + 8 * j], bh1, bl1);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma_tf32_16x8x8(acc[mt][j], al[mt], bh0, bh1);
        mma_tf32_16x8x8(acc[mt][j], ah[mt], bl0, bl1);
        mma_tf32_16x8x8(acc[mt][j], ah[mt], bh0, bh1);
      }
    }
    __syncwarp();  // all lanes are done with this stage before the next iteration refills it
  }
}

// out[32][1024] = init + Xs[32][128] (ld XS_LD) · Wt[128][1024].  Warp w owns columns [64w, 64w + 64).
// init(m, c) -> float2 start value for (row m, columns c, c+1), fetched BEFORE the k loop so that its global-memory
// latency hides behind the weight pipeline; epi(m, c, v0, v1) stores the result.
template <int MT, class Init, class Epi>
__device__ __forceinline__ void tile_gemm_wide_mma(const float* __restrict__ Xs, const float* __restrict__ Wt,
                                                   float* __restrict__ Wb, Init init, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = warp * 64 + 8 * j + 2 * t;
      const float2 i0 = init(mt * 16 + g, c), i1 = init(mt * 16 + g + 8, c);
      acc[mt][j][0] = i0.x; acc[mt][j][1] = i0.y; acc[mt][j][2] = i1.x; acc[mt][j][3] = i1.y;
    }
  warp_gemm_mma<MT>(acc, Xs, XS_LD, 0, Wt + warp * 64, QW, 16, Wb + warp * 2 * WARP_STAGE_FLOATS, lane);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = warp * 64 + 8 * j + 2 * t;
      epi(mt * 16 + g, c, acc[mt][j][0], acc[mt][j][1]);
      epi(mt * 16 + g + 8, c, acc[mt][j][2], acc[mt][j][3]);
    }
}

// out[32][128] (ld out_ld) = Cs[32][1024] (ld QC_LD) · Mt[1024][128] + bias.  Warp (kg = w >> 1, ng = w & 1) owns
// columns [64 ng, 64 ng + 64) over the k range [128 kg, 128 kg + 128); the 8 partial sums go through `part`
// ([8][32][128] floats, may alias Cs).  The caller must have synchronised the CTA after writing Cs.
template <int MT>
__device__ __forceinline__ void tile_gemm_tall_mma(float* __restrict__ Cs, const float* __restrict__ Mt,
                                                   float* __restrict__ Wb, const float* __restrict__ bias,
                                                   float* __restrict__ part, float* __restrict__ out, int out_ld) {
  constexpr int TMm = 16 * MT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kg = warp >> 1, ng = warp & 1;
  float acc[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
  warp_gemm_mma<MT>(acc, Cs, QC_LD, kg * 128, Mt + (size_t)kg * 128 * E + ng * 64, E, 16,
                    Wb + warp * 2 * WARP_STAGE_FLOATS, lane);
  __syncthreads();  // every warp is done reading Cs before `part` (which may alias it) is written
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = ng * 64 + 8 * j + 2 * t;
      *reinterpret_cast<float2*>(part + (kg * TMm + mt * 16 + g) * E + c) = make_float2(acc[mt][j][0], acc[mt][j][1]);
      *reinterpret_cast<float2*>(part + (kg * TMm + mt * 16 + g + 8) * E + c) = make_float2(acc[mt][j][2], acc[mt][j][3]);
    }
  __syncthreads();
  for (int o = tid; o < TMm * E; o += NT) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += part[k * TMm * E + o];
    out[(o >> 7) * out_ld + (o & (E - 1))] = s + (bias ? bias[o & (E - 1)] : 0.f);
  }
  __syncthreads();
}

// ---------------------------------------------------------------- swizzled-stage variants (decoder backward)
// Same tensor-pipe tile GEMMs with an UNPADDED per-warp weight stage [8 k-rows][NC = 8 NJ columns]: the 16-byte chunks of
// row r are stored at column c ^ ((r & 3) << 3), which makes the B-fragment reads (k = t, n = g) conflict free without
// the 8-float row padding — the stages of 16 warps then fit the backward kernel's 64 KiB weight buffer.
template <int NJ>
__device__ __forceinline__ void stage_warp_kstep_sw(const float* __restrict__ src, int ld, int kstep,
                                                    float* __restrict__ dst, int lane) {
  constexpr int NC = 8 * NJ, F4 = NC / 4;
#pragma unroll
  for (int i = 0; i < (8 * F4) / 32; ++i) {
    const int idx = lane + 32 * i;
    const int r = idx / F4, c4 = idx % F4;
    cp_async16(dst + r * NC + ((c4 * 4) ^ ((r & 3) << 3)), src + (size_t)(kstep * 8 + r) * ld + c4 * 4);
  }
}

template <int MT, int NJ>
__device__ __forceinline__ void warp_gemm_mma_sw(float (&acc)[MT][NJ][4], const float* __restrict__ A, int lda, int k_base,
                                                 const float* __restrict__ wsrc, int wld, int nks,
                                                 float* __restrict__ wbuf, int lane) {
  constexpr int NC = 8 * NJ, STAGE = 8 * NC;
  const int g = lane >> 2, t = lane & 3;
  stage_warp_kstep_sw<NJ>(wsrc, wld, 0, wbuf, lane);
  cp_async_commit();
  for (int ks = 0; ks < nks; ++ks) {
    if (ks + 1 < nks) {
      stage_warp_kstep_sw<NJ>(wsrc, wld, ks + 1, wbuf + ((ks + 1) & 1) * STAGE, lane);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncwarp();
    uint32_t ah[MT][4], al[MT][4];
    load_a_frags<MT>(A, lda, k_base + ks * 8, g, t, ah, al);
    const float* wb = wbuf + (ks & 1) * STAGE;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int c = (8 * j + g) ^ (t << 3);
      uint32_t bh0, bl0, bh1, bl1;
      split_tf32(wb[t * NC + c], bh0, bl0);
      split_tf32(wb[(t + 4) * NC + c], bh1, bl1);
#pragma unroll
      for (int mt = 0; mt < MT; ++mt) {
        mma_tf32_16x8x8(acc[mt][j], al[mt], bh0, bh1);
        mma_tf32_16x8x8(acc[mt][j], ah[mt], bl0, bl1);
        mma_tf32_16x8x8(acc[mt][j], ah[mt], bh0, bh1);
      }
    }
    __syncwarp();
  }
}

// out[16 MT][1024] = init + Xs[16 MT][128] (ld lda) · Wt[128][1024]; warp w owns columns [64w, 64w + 64).
// Wb: 16 warps x 2 stages x 512 floats = 64 KiB.
template <int MT, class Init, class Epi>
__device__ __forceinline__ void tile_gemm_wide_mma_sw(const float* __restrict__ Xs, int lda, const float* __restrict__ Wt,
                                                      float* __restrict__ Wb, Init init, Epi epi) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  float acc[MT][8][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = warp * 64 + 8 * j + 2 * t;
      const float2 i0 = init(mt * 16 + g, c), i1 = init(mt * 16 + g + 8, c);
      acc[mt][j][0] = i0.x; acc[mt][j][1] = i0.y; acc[mt][j][2] = i1.x; acc[mt][j][3] = i1.y;
    }
  warp_gemm_mma_sw<MT, 8>(acc, Xs, lda, 0, Wt + warp * 64, QW, 16, Wb + warp * 2 * 512, lane);
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = warp * 64 + 8 * j + 2 * t;
      epi(mt * 16 + g, c, acc[mt][j][0], acc[mt][j][1]);
      epi(mt * 16 + g + 8, c, acc[mt][j][2], acc[mt][j][3]);
    }
}

// out[16 MT][128] (ld out_ld) = Cs[16 MT][1024] (ld ldc) · Mt[1024][128] + bias.  Warp (kg = w >> 2, ng = w & 3) owns
// columns [32 ng, 32 ng + 32) over the k range [256 kg, 256 kg + 256); the 4 partial sums go through `part`
// ([4][16 MT][128] floats = 64 KiB at MT = 2; may alias Wb — it is written only after every warp has left the k loop).
template <int MT>
__device__ __forceinline__ void tile_gemm_tall_mma_sw(const float* __restrict__ Cs, int ldc, const float* __restrict__ Mt,
                                                      float* __restrict__ Wb, const float* __restrict__ bias,
                                                      float* __restrict__ part, float* __restrict__ out, int out_ld) {
  constexpr int TMm = 16 * MT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
  const int kg = warp >> 2, ng = warp & 3;
  float acc[MT][4][4];
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[mt][j][i] = 0.f;
  warp_gemm_mma_sw<MT, 4>(acc, Cs, ldc, kg * 256, Mt + (size_t)kg * 256 * E + ng * 32, E, 32, Wb + warp * 2 * 256, lane);
  __syncthreads();
#pragma unroll
  for (int mt = 0; mt < MT; ++mt)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = ng * 32 + 8 * j + 2 * t;
      *reinterpret_cast<float2*>(part + (kg * TMm + mt * 16 + g) * E + c) = make_float2(acc[mt][j][0], acc[mt][j][1]);
      *reinterpret_cast<float2*>(part + (kg * TMm + mt * 16 + g + 8) * E + c) = make_float2(acc[mt][j][2], acc[mt][j][3]);
    }
  __syncthreads();
  for (int o = tid; o < TMm * E; o += NT) {
    const float s = (part[o] + part[TMm * E + o]) + (part[2 * TMm * E + o] + part[3 * TMm * E + o]);
    out[(o >> 7) * out_ld + (o & (E - 1))] = s + (bias ? bias[o & (E - 1)] : 0.f);
  }
  __syncthreads();
}

}  // namespace vrpx

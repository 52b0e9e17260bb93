// mma.sync TF32 m16n8k8 on B200: (a) throughput with distinct operand registers per instruction, (b) latency of a
// dependent accumulator chain, (c) throughput with the hi/lo split ALU work interleaved as in the rollout kernel.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
template <int MODE>
__global__ void k(float* out, const float* in, int iters) {
  float c[8][4];
  for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  unsigned a[8][4], b[8][2];
  for (int j = 0; j < 8; ++j) { for (int i = 0; i < 4; ++i) a[j][i] = __float_as_uint(in[threadIdx.x + 32 * (j * 4 + i)]); b[j][0] = a[j][1] ^ 3; b[j][1] = a[j][2] ^ 5; }
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {        // 8 independent chains, distinct operands
#pragma unroll
      for (int j = 0; j < 8; ++j) mma(c[j], a[j][0], a[j][1], a[j][2], a[j][3], b[j][0], b[j][1]);
    } else if (MODE == 1) { // one dependent chain
#pragma unroll
      for (int j = 0; j < 8; ++j) mma(c[0], a[j][0], a[j][1], a[j][2], a[j][3], b[j][0], b[j][1]);
    } else {                // 8 chains + split ALU (6 values x 3 ops) per mma triple
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        unsigned h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) { float x = __uint_as_float(a[j][i]) + c[j][i]; h[i] = __float_as_uint(x) & 0xffffe000u; l[i] = __float_as_uint(x - __uint_as_float(h[i])); }
        mma(c[j], l[0], l[1], l[2], l[3], b[j][0], b[j][1]);
        mma(c[(j + 1) & 7], h[0], h[1], h[2], h[3], b[j][1], b[j][0]);
        mma(c[(j + 2) & 7], h[0], h[1], h[2], h[3], b[j][0], b[j][1]);
      }
    }
  }
  float s = 0; for (int j = 0; j < 8; ++j) for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float *out, *in; cudaMalloc(&out, sms * 1024 * 4); cudaMalloc(&in, 4096 * 4); cudaMemset(in, 0x3c, 4096 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int warps : {1, 4, 16}) for (int mode = 0; mode < 3; ++mode) {
    int iters = 20000; float ms = 0;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      if (mode == 0) k<0><<<sms, warps * 32>>>(out, in, iters); else if (mode == 1) k<1><<<sms, warps * 32>>>(out, in, iters); else k<2><<<sms, warps * 32>>>(out, in, iters);
      cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
    }
    double n = (double)warps * iters * (mode == 2 ? 24 : 8);
    double cyc = ms * 1e-3 * clk * 1e3;
    printf("warps/SM=%2d mode=%d: %.3f ms, %.3f mma/clk/SM, %.1f clk per mma per warp\n", warps, mode, ms, n / cyc, cyc / (n / warps));
  }
  return 0;
}

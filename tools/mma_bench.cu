// Microbenchmark: issue rate of FFMA and legacy mma.sync.m16n8k8 TF32 on B200 (per SM), to size the SIMT phases of
// the rollout kernel.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench mma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ffma(float* out, int iters) {
  float a[16];
  for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
  float x = 1.0001f, y = 0.9999f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
  }
  float s = 0;
  for (int i = 0; i < 16; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  unsigned a0 = 0x3f800000u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x3f000000u, b1 = 0x3e800000u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  }
  float s = 0;
  for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk = 0;
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sms * 8 * 1024 * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 32; warps *= 2) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32, blocks = sms * (warps * 32 / threads);
    int iters = 20000;
    for (int which = 0; which < 2; ++which) {
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        if (which == 0) k_ffma<<<blocks, threads>>>(out, iters); else k_mma<<<blocks, threads>>>(out, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      double instr = (double)blocks * (threads / 32) * iters * (which == 0 ? 16 : 8);
      double mac_per_instr = which == 0 ? 32 : 16 * 8 * 8;
      double macs = instr * mac_per_instr;
      printf("%s warps/SM=%2d: %.3f ms  %.1f TMAC/s  => %.0f MAC/clk/SM (at %.0f MHz), %.2f warp-instr/clk/SM\n",
             which == 0 ? "FFMA    " : "mma.tf32", warps, ms, macs / ms / 1e9, macs / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3,
             instr / (ms * 1e-3) / sms / (clk * 1e3));
    }
  }
  return 0;
}

// Microbenchmark: mma.sync rates on B200 per SM — TF32 m16n8k8 vs F16 / BF16 m16n8k16 (fp32 accumulate), with 8
// independent accumulators per warp (DEP=1) or 8 accumulators x 3 back-to-back dependent MMAs (DEP=3, the 3-term split
// pattern).  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_bench3 mma_bench3.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int DEP>
__global__ void k_mma(float* out, int iters) {
  float c[8][4];
  for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 4; ++i) c[j][i] = 0.f;
  unsigned a0 = 0x3c003c00u + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = 0x38003800u, b1 = 0x34003400u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int d = 0; d < DEP; ++d) {
        if (KIND == 0)
          asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        else if (KIND == 1)
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        else
          asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                       : "+f"(c[j][0]), "+f"(c[j][1]), "+f"(c[j][2]), "+f"(c[j][3])
                       : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
      }
  }
  float s = 0;
  for (int j = 0; j < 8; ++j)
    for (int i = 0; i < 4; ++i) s += c[j][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int DEP>
void run(const char* name, int sms, int clk, float* out) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int warps = 4; warps <= 16; warps *= 2) {
    int threads = warps * 32, blocks = sms, iters = 20000;
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      k_mma<KIND, DEP><<<blocks, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
    }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double instr = (double)blocks * warps * iters * 8 * DEP;
    double macs = instr * 16 * 8 * (KIND == 0 ? 8 : 16);
    printf("%s dep=%d warps/SM=%2d: %.3f ms  %.0f MAC/clk/SM, %.3f mma/clk/SM (at %.0f MHz nominal)\n", name, DEP, warps, ms,
           macs / (ms * 1e-3) / sms / (clk * 1e3), instr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1e3);
  }
}

int main() {
  int sms = 0, clk = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out;
  cudaMalloc(&out, sms * 1024 * sizeof(float));
  run<0, 1>("tf32 m16n8k8 ", sms, clk, out);
  run<0, 3>("tf32 m16n8k8 ", sms, clk, out);
  run<1, 1>("f16  m16n8k16", sms, clk, out);
  run<1, 3>("f16  m16n8k16", sms, clk, out);
  run<2, 1>("bf16 m16n8k16", sms, clk, out);
  run<2, 3>("bf16 m16n8k16", sms, clk, out);
  return 0;
}

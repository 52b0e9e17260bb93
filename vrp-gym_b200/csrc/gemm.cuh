// gemm.cuh — internal GEMM interface shared by the SIMT cross-check path (encoder.cu) and the
// tcgen05 3xTF32 tensor-core path (gemm_tc.cu).
#pragma once
#include "common.cuh"

namespace vrpx {

// Y[R][NOUT] = epilogue( X[R][K] · W[NOUT][K]^T )      (W in torch nn.Linear layout)
// epilogue order: acc * (gate > 0) -> + bias -> relu -> + residual -> * scale + shift
struct GemmArgs {
  const float* X;
  int64_t R;
  int K;
  const float* W;
  int NOUT;
  const float* bias;      // [NOUT] or nullptr
  int relu;
  const float* residual;  // [R][NOUT] or nullptr
  const float* scale;     // [NOUT] or nullptr  (eval-mode BatchNorm folded affine)
  const float* shift;     // [NOUT] or nullptr
  float* Y;
  const float* gate = nullptr;  // [R][NOUT] or nullptr: ReLU-backward gate (acc is zeroed where gate <= 0)
};

int gemm_simt(const GemmArgs& a, cudaStream_t stream);  // fp32 FFMA, smem tiled
int gemm_tc(const GemmArgs& a, cudaStream_t stream);    // tcgen05.mma kind::tf32, 3-term split (≈fp32); persistent, TMA-staged, A operand in TMEM (gemm_tc3.cu)
int gemm_tc_v1(const GemmArgs& a, cudaStream_t stream); // first-generation tcgen05 kernel (gemm_tc.cu), cross-check
int gemm_tc_v2(const GemmArgs& a, cudaStream_t stream); // persistent TMA kernel with both operands in smem (gemm_tc2.cu)

inline int gemm_dispatch(int path, const GemmArgs& a, cudaStream_t stream) {
  return path == 0 ? gemm_tc(a, stream) : (path == 2 ? gemm_tc_v1(a, stream) : (path == 3 ? gemm_tc_v2(a, stream) : gemm_simt(a, stream)));
}

}  // namespace vrpx

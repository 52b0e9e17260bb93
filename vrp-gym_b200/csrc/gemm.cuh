// gemm.cuh — internal GEMM interface shared by the SIMT cross-check path (encoder.cu) and the
// tcgen05 tensor-core path (gemm_tc4.cu).
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>

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

// a prepared tcgen05 GEMM (gemm_tc4.cu): tensor maps + launch shape, reusable while the buffers stay where they are
struct GemmPlan {
  GemmArgs a;
  CUtensorMap mx, mwh, mwl;
  int grid, variant;
};
int gemm_tc_plan(const GemmArgs& a, __half* w16, GemmPlan* plan, cudaStream_t stream);   // w16: 2 * NOUT * K halves
int gemm_tc_launch(const GemmPlan& plan, cudaStream_t stream);

int gemm_simt(const GemmArgs& a, cudaStream_t stream);     // fp32 FFMA, smem tiled (encoder.cu)
int gemm_tc(const GemmArgs& a, cudaStream_t stream);       // production: tcgen05.mma kind::f16 on f16 hi/lo halves (≈fp32), persistent,
                                                           // TMA-staged, A operand in TMEM (gemm_tc4.cu)

// the encoder feed-forward block in one kernel (ff_fused.cu): Y = (residual + relu(X·W1^T + b1)·W2^T + b2) * scale + shift
int ff_fused(const float* X, int64_t R, const float* W1, const float* b1, const float* W2, const float* b2,
             const float* residual, const float* scale, const float* shift, float* Y, cudaStream_t stream);

// QKV projection + per-instance self-attention of an encoder layer in one kernel (attn_fused.cu)
int qkv_attention_fused(const float* X, const float* in_proj_w, const float* in_proj_b, int64_t B, int N, float* att,
                        cudaStream_t stream);

// weight-gradient GEMM C [M][N] += A^T · B over the R rows of A [R][M], B [R][N] on tcgen05 (gemm_tn_tc.cu); M, N % 128 == 0
int gemm_tn_tc(const float* A, const float* Bm, float* C, float* colsum_A, int64_t R, int M, int N, cudaStream_t stream);
// C += A^T · B and (optional) colsum_A [M] += column sums of A, through gemm_tn_tc when the shape allows, else the warp-level
// kernels of decoder_bwd.cu (the bias gradient of a linear layer rides on its weight-gradient GEMM)
int gemm_tn_accumulate(const float* A, const float* Bm, float* C, float* colsum_A, int64_t R, int M, int N, cudaStream_t stream);

// backward of the per-instance self-attention on mma.sync (attention_bwd.cu): dqkv [B·N][384] from qkv, att = O, datt = dO
int attention_backward_mma(const float* qkv, const float* att, const float* datt, float* dqkv, int64_t B, int N, cudaStream_t stream);

// path: 0 tcgen05 f16-split (production), 1 fp32 SIMT (cross-check: separates tensor-core error from algorithmic error)
inline int gemm_dispatch(int path, const GemmArgs& a, cudaStream_t stream) {
  if (path != 0 && path != 1) {
    set_error("gemm path %d does not exist (0: tcgen05 f16-split, 1: fp32 SIMT)", path);
    return VRPX_ERR_ARG;
  }
  return path == 0 ? gemm_tc(a, stream) : gemm_simt(a, stream);
}

}  // namespace vrpx

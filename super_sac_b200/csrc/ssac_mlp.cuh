// Shared declarations of the ensemble-MLP GEMM paths (fp32 FFMA tiles and tcgen05 3xTF32 tiles).
#pragma once
#include "ssac_common.cuh"

namespace ssac {

enum { L_NT = 0, L_NN = 1, L_TN = 2 };

// C[g] (M x N, ldc) = epilogue( opA(A[g]) * opB(B[g]) )
//   L_NT: A [M x K] row-major, B = W [N x K] row-major (nn.Linear weight)      -> forward layers
//   L_NN: A [M x K] row-major, B = W [K x N] row-major                         -> backward data path
//   L_TN: A given as [K x M] row-major, B [K x N] row-major                    -> backward weight path
struct GemmP {
  const float* A; int64_t lda, a_gs;
  const float* Bm; int64_t ldb, b_gs;
  const int32_t* b_index;  // group -> weight block (REDQ subset); NULL = identity
  float* C; int64_t ldc, c_gs;
  const float* bias; int64_t bias_gs;                                  // + bias[n]
  const float* mask; int64_t ldmask, mask_gs;                          // .* (mask[m][n] > 0)
  const float* extra; int64_t ldextra, extra_gs; float extra_scale;   // + s * extra[m][n] (before the mask)
  float* colsum; int64_t colsum_gs;                                    // L_TN: colsum[m] = sum_k A[k][m]  (bias grads)
  const float* a_kscale; int64_t a_kscale_gs;                          // L_TN (TMA-fed A only): A[k][m] *= a_kscale[g][k] while staging
  // TMA-fed A only: the operand is GENERATED from the loaded tile while it is staged, A' = A > 0 ? a_gate[g][j] : 0 with j
  // the index along A's contiguous dimension (k for L_NN / L_NT, m for L_TN).  With A = h2 and a_gate = W3 this is
  // v = W3 .* (h2 > 0) of the split critic backward, which therefore never exists in memory.
  const float* a_gate; int64_t a_gate_gs;
  int M, N, K;
  int relu, accumulate;
  int pdl;   // host side: launch with programmatic stream serialization (the predecessor in the stream is PDL-aware)
};

int launch_gemm_tc(int layout, const GemmP& p, int G, cudaStream_t s, const char* what);

// Work fused into the output-layer kernel of a forward pass (narrow heads, O <= 32): what the reference does with a
// dozen element-wise launches after the last nn.Linear.
struct HeadEpi {
  int kind;  // 0 = none, 1 = tanh-Normal sample + log-prob, 2 = deterministic head (+TD3 noise), 3 = critic loss seed
  // kinds 1 / 2: nets/distributions.py:9-15,64-114, learning_utils.py:48-59
  const float* eps; const float* noise; float sigma, clip, lo, hi;
  float* a; int64_t lda; float* logp; float* tanh_out; int A;
  // kind 3: learning.py:90-98,112
  const float* y; const float* w; const float* imp; const float* popart; int pop; float inv_count; float* dq; float* loss;
  // kind 3 with the TD target computed in place (no PopArt; learning_utils.py:319-353): y[b] = r + gamma (1-d)
  // (min_m qt[m][b] - alpha logp[b]); y_out and the logged statistics are written by the blocks of net 0
  const float* qt; int M; const float* logp_t; const float* log_alpha; const float* r; const float* d; float gamma;
  float* y_out; float* td_logs;   // td_logs[0..3] += {sum (y - c), sum (y - c)^2, sum alpha logp}, td_logs[3] = c = y[0]
};

}  // namespace ssac

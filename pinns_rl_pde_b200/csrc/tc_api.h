// tc_api.h -- entry points of the tcgen05 3xTF32 GEMM kernels (tc_gemm.cuh), one translation unit per kernel family
// (tc_rows_fwd.cu, tc_rows_bwd.cu, tc_wgrad.cu) so that the library builds in parallel.
// Every function returns 0 when launched, TC_UNSUPPORTED when the shape is not covered (the caller then uses the
// exact-fp32 CUDA-core GEMM), < 0 on a launch error.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "jet_math.cuh"

namespace pinnk {
constexpr int TC_UNSUPPORTED = 1;

// Loss fusion of the last hidden layer (tc_linear_act_fwd with w_out and this struct): the epilogue completes the output
// jets U through shared memory, evaluates the PDE residual, the loss sum and its seeds dL/dU, runs the adjoint of the
// output layer and of the tanh on the spot and writes dL/dZ of the last hidden layer: the activation output, U, the
// loss epilogue kernel and the last-layer reverse kernel all disappear from the step.
struct TcLossFuse {
  PdeDesc pde;
  JetSpec js;
  int loss_kind;
  float huber_delta;
  float weight;        // multiplies rho(e) in the loss sum
  float grad_weight;   // multiplies rho'(e) in the seeds
  double* loss_slot;   // += sum rho * weight
  float* dz_out;       // [M, N] dL/d(pre-activation jets of the last hidden layer)
  float* gw_out;       // [N] += dL/dw_out, or null
  float* gb_out;       // [1] += dL/db_out, or null
  const float* b_out;  // [1] output-layer bias, or null
};

// Z[M,N] = X[M,K] W[N,K]^T (+ bias on value-column rows)
// ring / ring_floats (here and below): device scratch for the K-split launch of 256-wide contractions (one launch of CTA
// pairs instead of two K-half passes; needs (sm_count / 2) * 32768 floats); null -> two passes
int tc_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int K, int N, int jet_cols,
                  int sm_count, cudaStream_t st, float* ring = nullptr, int64_t ring_floats = 0);
// Forward Linear + activation jets in one kernel: Z = X W^T + b (stash, may be null for K = 128), Yact = act(Z).
// act: 1 tanh, 2 sin(omega z); (k0, k1) = jet orders of the (at most two) directions.
// w_out != null folds the network's output layer nn.Linear(N, 1) in: u_part[(N/128) * 4][M] receives per-warp partial
// output jets (sum them in order + bias: output_combine_kernel); Yact may then be null (forward-only callers).
int tc_linear_act_fwd(const float* X, const float* W, const float* bias, float* Z, float* Yact, int64_t M, int K, int N,
                      int k0, int k1, int act, float omega, int sm_count, cudaStream_t st,
                      const float* w_out = nullptr, float* u_part = nullptr, const TcLossFuse* loss = nullptr,
                      float* ring = nullptr, int64_t ring_floats = 0);
// dX[M,in] = dZ[M,out] W[out,in]   (W row-major [out,in])
int tc_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int in_dim, int out_dim, int sm_count,
                    cudaStream_t st, float* ring = nullptr, int64_t ring_floats = 0);
// dgrad + activation adjoint in one kernel: dZprev = act'(Zprev)^T (dZ W).  from_y = 1 (tanh only): Zprev holds the
// activation OUTPUT jets instead of the pre-activation jets (the forward pass then does not stash the latter)
int tc_linear_dgrad_actbwd(const float* dZ, const float* W, const float* Zprev, float* dZprev, int64_t M, int in_dim,
                           int out_dim, int k0, int k1, int act, float omega, int sm_count, cudaStream_t st, int from_y,
                           float* ring = nullptr, int64_t ring_floats = 0);
// internal: the K-split launches (own translation units tc_ksplit_fwd.cu / tc_ksplit_bwd.cu); TC_UNSUPPORTED when the
// shape, the row count or the scratch size rules the launch out
int tc_ks_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int N, int jet_cols, int sm_count,
                     cudaStream_t st, float* ring, int64_t ring_floats);
int tc_ks_linear_act_fwd(const float* X, const float* W, const float* bias, float* Z, float* Yact, int64_t M, int N, int k0,
                         int k1, int act, float omega, int sm_count, cudaStream_t st, const float* w_out, float* u_part,
                         float* ring, int64_t ring_floats);
int tc_ks_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int in_dim, int sm_count, cudaStream_t st,
                       float* ring, int64_t ring_floats);
int tc_ks_linear_dgrad_actbwd(const float* dZ, const float* W, const float* Zprev, float* dZprev, int64_t M, int in_dim, int k0,
                              int k1, int act, float omega, int sm_count, cudaStream_t st, float* ring, int64_t ring_floats);
// jet layouts (orders of the at most two directions) the fused epilogues are instantiated for
inline bool tc_jets_supported(int k0, int k1) {
  return (k0 == 0 && k1 == 0) || (k0 == 1 && k1 == 0) || (k0 == 2 && k1 == 1) || (k0 == 3 && k1 == 0) ||       // 1, 2, 4, 4
         (k0 == 1 && k1 == 1) || (k0 == 3 && k1 == 1) || (k0 == 2 && k1 == 2) || (k0 == 4 && k1 == 1);         // 3, 5, 5, 6 columns
}
// dgrad of the first hidden layer fused with the complete reverse of the network's input layer
// (nn.Linear(net_in_dim <= 4, in_dim) + activation): accumulates dW0 / db0, writes nothing else.  vec0 / vec1: direction
// vectors (4 floats) of the two jet directions.
int tc_linear_dgrad_firstbwd(const float* dZ, const float* W, int64_t M, int in_dim, int out_dim, int k0, int k1, int act,
                             float omega, const float* x, const float* t, int net_in_dim, const float* vec0,
                             const float* vec1, const float* W0, const float* b0, float* gW0, float* gb0, int sm_count,
                             cudaStream_t st);
// internal: EPI_ACTBWD_Y dispatch (lives in its own translation unit)
int tc_dgrad_actbwd_y(int k0, int k1, const float* dZ, const float* W, int in_dim, float* dZprev, int64_t M,
                      const float* Yprev, int sm_count, cudaStream_t st, int out_dim, int accum);
// dW[out,in] += dZ[M,out]^T X[M,in] ;  db[out] += sum over value-column rows of dZ
// det_scratch != null: deterministic reduction -- per-CTA partial slabs in det_scratch (det_floats floats available,
// >= 16640 per 128 x 128 block), added in CTA order by a second kernel instead of atomics in arrival order
int tc_linear_wgrad(const float* dZ, const float* X, float* dW, float* db, int64_t M, int in_dim, int out_dim,
                    int jet_cols, int sm_count, cudaStream_t st, float* det_scratch = nullptr, int64_t det_floats = 0);
// dgrad + tanh adjoint (from the previous layer's OUTPUT jets) and wgrad of one Linear(128, 128) in ONE launch of CTA pairs
// sharing a tile stream (bwd_pair_kernel): 1536 B of HBM traffic per row instead of 2560
int tc_bwd_pair(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db, int64_t M, int in_dim,
                int out_dim, int k0, int k1, int sm_count, cudaStream_t st);
// stage timers of the rows kernels (zeros unless built with -DPINNK_STAGE_TIMERS); 0 ok, 1 not compiled in
int tc_stage_timers_fwd(unsigned long long* out16, int reset);
int tc_stage_timers_bwd(unsigned long long* out16, int reset);
int tc_stage_timers_wgrad(unsigned long long* out16, int reset);
int tc_stage_timers_bwd_y(unsigned long long* out16, int reset);
}  // namespace pinnk

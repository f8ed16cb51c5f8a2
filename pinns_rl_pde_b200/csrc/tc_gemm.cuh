// tc_gemm.cuh -- tcgen05 (5th-gen tensor core) 3xTF32 GEMM path for the hidden Linear layers.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinnk {
constexpr int TC_UNSUPPORTED = 1;

// Z[M,N] = X[M,K] W[N,K]^T (+ bias on value-column rows).  Returns 0 when launched,
// TC_UNSUPPORTED when the shape is not covered (caller uses the exact-fp32 CUDA-core GEMM).
static inline int tc_linear_fwd(const float*, const float*, const float*, float*, int64_t, int, int, int, int, cudaStream_t) {
  return TC_UNSUPPORTED;
}
}  // namespace pinnk

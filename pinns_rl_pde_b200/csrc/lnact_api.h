// lnact_api.h -- LayerNorm + tanh jets as one kernel each way, one THREAD per feature pair (lnact_feat.cu).
// Y = tanh(LayerNorm(Z) (+ S)) on derivative jets: resnet.py:45-65 (Linear -> LayerNorm -> tanh and
// Linear -> LayerNorm -> (+ x) -> tanh).  Same contract as lnact_fwd_kernel / lnact_bwd_kernel (kernels_ew.cuh).
// Every function returns 0 when launched, 1 when the shape is not covered (the caller then runs the separate LayerNorm
// and activation kernels), -1 on a launch error.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinnk {

// width in {128, 256, 512}; jet layout [u | k0 columns of direction 0 | k1 columns of direction 1]
bool lnact_feat_supported(int width, int k0, int k1);

int lnact_feat_fwd(const float* Z, const float* S, float* Y, int64_t n, int width, int k0, int k1, const float* gamma,
                   const float* beta, float eps, int sm_count, cudaStream_t st);

// Gin (+ Gin2, nullable) = dL/dY; Gout = dL/dZ (distinct from Gin); Gz (nullable, may alias Gin) = dL/d(pre-activation)
int lnact_feat_bwd(const float* Z, const float* S, const float* Gin, const float* Gin2, float* Gz, float* Gout, int64_t n,
                   int width, int k0, int k1, const float* gamma, const float* beta, float eps, float* dgamma, float* dbeta,
                   int sm_count, cudaStream_t st);

}  // namespace pinnk

// sgemm.cuh -- exact-fp32 CUDA-core GEMM used for layer shapes the tcgen05 path does not cover
// (odd widths, the 64-wide Fourier layer, tiny boundary/initial row sets) and as the in-library
// cross-check of the 3xTF32 tensor-core kernels.
//
//   C[M,N] (+)= A(m,k) * B(k,n)
//     A_KMAJOR : A[m*lda + k]   else A[k*lda + m]
//     B_KMAJOR : B[n*ldb + k]   else B[k*ldb + n]
//   EPI_STORE      C = acc
//   EPI_BIAS_C0    C = acc + bias[n] on rows with (m % jet_cols == 0)   (Linear bias hits the value column only)
//   EPI_ATOMIC     atomicAdd(C, acc)  with split-K over gridDim.z        (weight gradients)
//
// 128x128x8 tiles, 256 threads, 8x8 register micro-tile, register-prefetch double buffering.
// Feature dimensions must be multiples of 4 (float4 global accesses); row counts are arbitrary.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pinnk {

enum { EPI_STORE = 0, EPI_BIAS_C0 = 1, EPI_ATOMIC = 2 };

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 8, SG_THREADS = 256;

template <bool KMAJOR>
__device__ __forceinline__ float4 sg_load(const float* __restrict__ P, int64_t ld, int64_t mn, int64_t k,
                                          int64_t MN, int64_t K) {
  // KMAJOR: float4 along k at row mn.  else: float4 along mn at row k.
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (KMAJOR) {
    if (mn < MN && k < K) v = *reinterpret_cast<const float4*>(P + mn * ld + k);
  } else {
    if (k < K && mn < MN) v = *reinterpret_cast<const float4*>(P + k * ld + mn);
  }
  return v;
}

template <bool A_KMAJOR, bool B_KMAJOR, int EPI>
__global__ void __launch_bounds__(SG_THREADS)
sgemm_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C,
             int64_t M, int64_t N, int64_t K, int64_t lda, int64_t ldb, int64_t ldc,
             const float* __restrict__ bias, int jet_cols, int64_t k_chunk) {
  __shared__ __align__(16) float As[2][SG_BK][SG_BM];
  __shared__ __align__(16) float Bs[2][SG_BK][SG_BN];
  const int tid = threadIdx.x;
  const int64_t m0 = (int64_t)blockIdx.x * SG_BM;
  const int64_t n0 = (int64_t)blockIdx.y * SG_BN;
  const int64_t k_begin = (int64_t)blockIdx.z * k_chunk;
  const int64_t k_end = (k_begin + k_chunk < K) ? (k_begin + k_chunk) : K;
  if (k_begin >= k_end) return;

  // load coordinates
  const int a_mn = A_KMAJOR ? (tid >> 1) : ((tid & 31) << 2);
  const int a_k = A_KMAJOR ? ((tid & 1) << 2) : (tid >> 5);
  const int b_mn = B_KMAJOR ? (tid >> 1) : ((tid & 31) << 2);
  const int b_k = B_KMAJOR ? ((tid & 1) << 2) : (tid >> 5);

  const int ty = tid >> 4, tx = tid & 15;
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto stage = [&](int buf, const float4& va, const float4& vb) {
    if (A_KMAJOR) {
      As[buf][a_k + 0][a_mn] = va.x; As[buf][a_k + 1][a_mn] = va.y;
      As[buf][a_k + 2][a_mn] = va.z; As[buf][a_k + 3][a_mn] = va.w;
    } else {
      *reinterpret_cast<float4*>(&As[buf][a_k][a_mn]) = va;
    }
    if (B_KMAJOR) {
      Bs[buf][b_k + 0][b_mn] = vb.x; Bs[buf][b_k + 1][b_mn] = vb.y;
      Bs[buf][b_k + 2][b_mn] = vb.z; Bs[buf][b_k + 3][b_mn] = vb.w;
    } else {
      *reinterpret_cast<float4*>(&Bs[buf][b_k][b_mn]) = vb;
    }
  };

  float4 va = sg_load<A_KMAJOR>(A, lda, m0 + a_mn, k_begin + a_k, M, k_end);
  float4 vb = sg_load<B_KMAJOR>(B, ldb, n0 + b_mn, k_begin + b_k, N, k_end);
  stage(0, va, vb);
  __syncthreads();

  int buf = 0;
  for (int64_t kt = k_begin; kt < k_end; kt += SG_BK) {
    const bool has_next = (kt + SG_BK) < k_end;
    if (has_next) {
      va = sg_load<A_KMAJOR>(A, lda, m0 + a_mn, kt + SG_BK + a_k, M, k_end);
      vb = sg_load<B_KMAJOR>(B, ldb, n0 + b_mn, kt + SG_BK + b_k, N, k_end);
    }
#pragma unroll
    for (int kk = 0; kk < SG_BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (has_next) {
      stage(buf ^ 1, va, vb);
      __syncthreads();
      buf ^= 1;
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t m = m0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int64_t n = n0 + jh * 64 + tx * 4;
      if (n >= N) continue;   // N % 4 == 0 so a float4 never straddles the edge
      float4 v = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]);
      float* dst = C + m * ldc + n;
      if (EPI == EPI_ATOMIC) {
        atomicAdd(dst + 0, v.x); atomicAdd(dst + 1, v.y); atomicAdd(dst + 2, v.z); atomicAdd(dst + 3, v.w);
      } else {
        if (EPI == EPI_BIAS_C0 && bias != nullptr && (m % jet_cols) == 0) {
          const float4 bv = *reinterpret_cast<const float4*>(bias + n);
          v.x += bv.x; v.y += bv.y; v.z += bv.z; v.w += bv.w;
        }
        *reinterpret_cast<float4*>(dst) = v;
      }
    }
  }
}

}  // namespace pinnk

// kernels_ew.cuh -- CUDA-core kernels of the jet pipeline: first/last Linear (K<=4 / N=1),
// activation jets and their adjoints, LayerNorm jets and their adjoints, the PDE/loss epilogue
// and the scoring reduction.  Jet tensors are row-major [points, C, width]: the C jet columns
// of a point are adjacent rows, so a tensor is also the [points*C, width] operand of a GEMM.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "jet_math.cuh"

namespace pinnk {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ float load_in(const float* __restrict__ x, const float* __restrict__ t,
                                         int64_t p, int i, int in_dim) {
  if (t == nullptr) return x[p * in_dim + i];
  return (i < in_dim - 1) ? x[p * (in_dim - 1) + i] : t[p];
}

// ------------------------------------------------------------------ first Linear (K = in_dim <= 4)
// Z[p,0,:] = W xt_p + b ; Z[p,col0[d],:] = W vec_d ; higher orders 0.     (first nn.Linear of every net)
__global__ void first_linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                                        const float* __restrict__ W, const float* __restrict__ b,
                                        int w_transposed, int out_dim, JetSpec js, float* __restrict__ Z) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * out_dim) return;
  const int64_t p = idx / out_dim;
  const int f = (int)(idx - p * out_dim);
  float w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = (i < js.in_dim) ? (w_transposed ? W[i * out_dim + f] : W[f * js.in_dim + i]) : 0.f;
  float z0 = b ? b[f] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < js.in_dim) z0 = fmaf(w[i], load_in(x, t, p, i, js.in_dim), z0);
  float* zp = Z + p * js.ncols * out_dim + f;
  zp[0] = z0;
  for (int d = 0; d < js.ndirs; ++d) {
    float z1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) z1 = fmaf(w[i], js.vec[d][i], z1);
    zp[(int64_t)js.col0[d] * out_dim] = z1;
    for (int k = 2; k <= js.order[d]; ++k) zp[(int64_t)(js.col0[d] + k - 1) * out_dim] = 0.f;
  }
}

// dW[f,i] += sum_p Zb[p,0,f] xt[p,i] + sum_p sum_d Zb[p,col0[d],f] vec_d[i] ;  db[f] += sum_p Zb[p,0,f]
__global__ void first_linear_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                                        int out_dim, JetSpec js, const float* __restrict__ Zb,
                                        float* __restrict__ gW, float* __restrict__ gb) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= out_dim) return;
  float aw[4] = {0.f, 0.f, 0.f, 0.f};
  float ab = 0.f;
  for (int64_t p = blockIdx.y; p < n; p += gridDim.y) {
    const float* zp = Zb + p * js.ncols * out_dim + f;
    const float g0 = zp[0];
    ab += g0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < js.in_dim) aw[i] = fmaf(g0, load_in(x, t, p, i, js.in_dim), aw[i]);
    for (int d = 0; d < js.ndirs; ++d) {
      const float g1 = zp[(int64_t)js.col0[d] * out_dim];
#pragma unroll
      for (int i = 0; i < 4; ++i) aw[i] = fmaf(g1, js.vec[d][i], aw[i]);
    }
  }
  if (gW) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < js.in_dim) atomicAdd(gW + (int64_t)f * js.in_dim + i, aw[i]);
  }
  if (gb) atomicAdd(gb + f, ab);
}

// ------------------------------------------------------------------ activation jets
template <int ACT> struct ActTag {};

// Y = act(Z (+ S)).  One thread per (point, feature); directions processed one at a time.
template <int ACT, int MAXK>
__global__ void act_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ Y,
                               int64_t n, int width, JetSpec js, float omega) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * width) return;
  const int64_t p = idx / width;
  const int f = (int)(idx - p * width);
  const int64_t base = p * js.ncols * width + f;
  float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
  z[0] = Z[base] + (S ? S[base] : 0.f);
  if (ACT == 1) {  // tanh
    y[0] = tanhf(z[0]);
    w[0] = 1.f - y[0] * y[0];
  } else {         // sin(omega z): y = sin, w = cos
    z[0] *= omega;
    sincosf(z[0], &y[0], &w[0]);
  }
  Y[base] = y[0];
  // Every jet column of the element is requested before the first one is used: with the loads of a direction issued only when
  // the previous direction had been stored, a thread had at most MAXK requests in flight and the kernel ran latency-bound
  // (C4, 18 columns: 3.6 TB/s at 95 % occupancy with long-scoreboard stalls on top, profiles/r02_final.md).  The direction loop is
  // unrolled over kMaxDirs so that zin[][] stays in registers.
  float zin[kMaxDirs][MAXK];
#pragma unroll
  for (int d = 0; d < kMaxDirs; ++d) {
    if (d < js.ndirs) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * width;
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
          const int64_t o = b1 + (int64_t)(k - 1) * width;
          zin[d][k - 1] = Z[o] + (S ? S[o] : 0.f);
        }
    }
  }
#pragma unroll
  for (int d = 0; d < kMaxDirs; ++d) {
    if (d < js.ndirs) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * width;
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) z[k] = (ACT == 2) ? zin[d][k - 1] * omega : zin[d][k - 1];
      if (ACT == 1) tanh_dir_fwd<MAXK, float>(K, z, y, w);
      else sincos_dir_fwd<MAXK, float>(K, z, y, w);
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) Y[b1 + (int64_t)(k - 1) * width] = y[k];
    }
  }
}

// G (in: dL/dY, out: dL/dZ) in place; Z (+S) is the stashed pre-activation.
template <int ACT, int MAXK>
__global__ void act_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ G,
                               int64_t n, int width, JetSpec js, float omega, const float* __restrict__ G2 = nullptr) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * width) return;
  const int64_t p = idx / width;
  const int f = (int)(idx - p * width);
  const int64_t base = p * js.ncols * width + f;
  float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
  z[0] = Z[base] + (S ? S[base] : 0.f);
  if (ACT == 1) {
    y[0] = tanhf(z[0]);
    w[0] = 1.f - y[0] * y[0];
  } else {
    z[0] *= omega;
    sincosf(z[0], &y[0], &w[0]);
  }
  yb[0] = G[base] + (G2 ? G2[base] : 0.f);      // G2: a second adjoint of the same output (arriving over a skip connection)
  float wb0 = 0.f;   // tanh: adjoint of w0 ; sin: adjoint of c0
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    const int64_t b1 = base + (int64_t)js.col0[d] * width;
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
      if (k <= K) {
        const int64_t o = b1 + (int64_t)(k - 1) * width;
        z[k] = Z[o] + (S ? S[o] : 0.f);
        if (ACT == 2) z[k] *= omega;
        yb[k] = G[o] + (G2 ? G2[o] : 0.f);
      } else {
        yb[k] = 0.f;
      }
    if (ACT == 1) {
      tanh_dir_fwd<MAXK, float>(K, z, y, w);
      tanh_dir_bwd<MAXK, float>(K, z, y, w, yb, zb, wb0);
    } else {
#pragma unroll
      for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
      wb[0] = wb0;
      sincos_dir_fwd<MAXK, float>(K, z, y, w);
      sincos_dir_bwd<MAXK, float>(K, z, y, w, yb, wb, zb);
      wb0 = wb[0];
    }
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
      if (k <= K) G[b1 + (int64_t)(k - 1) * width] = (ACT == 2) ? zb[k] * omega : zb[k];
  }
  if (ACT == 1) G[base] = tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0);
  else G[base] = (yb[0] * w[0] - wb0 * y[0]) * omega;
}

// The same two kernels with PPT points per thread (same feature, consecutive points): every load of a phase is issued for
// all PPT points before the recurrences run.  A/B variant (PINNK_ACT_PPT=2), parity-green but slower than one point per
// thread on a B200 (64 - 70 registers against 40 - 48: the occupancy lost outweighs the batched loads), so not the default.
// A ragged last group clamps its loads to the group's first point and skips the stores.
template <int ACT, int MAXK, int PPT>
__global__ void act_fwd_multi_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ Y,
                                     int64_t n, int width, JetSpec js, float omega) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t groups = (n + PPT - 1) / PPT;
  if (idx >= groups * width) return;
  const int64_t q = idx / width;
  const int f = (int)(idx - q * width);
  int64_t base[PPT];
  bool ok[PPT];
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    const int64_t p = q * PPT + u;
    ok[u] = p < n;
    base[u] = (ok[u] ? p : q * PPT) * js.ncols * width + f;
  }
  float z[PPT][MAXK + 1], y[PPT][MAXK + 1], w[PPT][MAXK + 1];
#pragma unroll
  for (int u = 0; u < PPT; ++u) z[u][0] = Z[base[u]] + (S ? S[base[u]] : 0.f);
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    if (ACT == 1) {
      y[u][0] = tanhf(z[u][0]);
      w[u][0] = 1.f - y[u][0] * y[u][0];
    } else {
      z[u][0] *= omega;
      sincosf(z[u][0], &y[u][0], &w[u][0]);
    }
    if (ok[u]) Y[base[u]] = y[u][0];
  }
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    const int64_t c0 = (int64_t)js.col0[d] * width;
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
          const int64_t o = base[u] + c0 + (int64_t)(k - 1) * width;
          z[u][k] = Z[o] + (S ? S[o] : 0.f);
          if (ACT == 2) z[u][k] *= omega;
        }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      if (ACT == 1) tanh_dir_fwd<MAXK, float>(K, z[u], y[u], w[u]);
      else sincos_dir_fwd<MAXK, float>(K, z[u], y[u], w[u]);
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K && ok[u]) Y[base[u] + c0 + (int64_t)(k - 1) * width] = y[u][k];
  }
}

template <int ACT, int MAXK, int PPT>
__global__ void act_bwd_multi_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ G,
                                     int64_t n, int width, JetSpec js, float omega) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t groups = (n + PPT - 1) / PPT;
  if (idx >= groups * width) return;
  const int64_t q = idx / width;
  const int f = (int)(idx - q * width);
  int64_t base[PPT];
  bool ok[PPT];
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    const int64_t p = q * PPT + u;
    ok[u] = p < n;
    base[u] = (ok[u] ? p : q * PPT) * js.ncols * width + f;
  }
  float z[PPT][MAXK + 1], y[PPT][MAXK + 1], w[PPT][MAXK + 1], yb[PPT][MAXK + 1], zb[PPT][MAXK + 1], wb[MAXK + 1];
  float wb0[PPT];
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    z[u][0] = Z[base[u]] + (S ? S[base[u]] : 0.f);
    yb[u][0] = G[base[u]];
    wb0[u] = 0.f;
  }
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    if (ACT == 1) {
      y[u][0] = tanhf(z[u][0]);
      w[u][0] = 1.f - y[u][0] * y[u][0];
    } else {
      z[u][0] *= omega;
      sincosf(z[u][0], &y[u][0], &w[u][0]);
    }
  }
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    const int64_t c0 = (int64_t)js.col0[d] * width;
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
          const int64_t o = base[u] + c0 + (int64_t)(k - 1) * width;
          z[u][k] = Z[o] + (S ? S[o] : 0.f);
          if (ACT == 2) z[u][k] *= omega;
          yb[u][k] = G[o];
        } else {
          yb[u][k] = 0.f;
        }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      if (ACT == 1) {
        tanh_dir_fwd<MAXK, float>(K, z[u], y[u], w[u]);
        tanh_dir_bwd<MAXK, float>(K, z[u], y[u], w[u], yb[u], zb[u], wb0[u]);
      } else {
#pragma unroll
        for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
        wb[0] = wb0[u];
        sincos_dir_fwd<MAXK, float>(K, z[u], y[u], w[u]);
        sincos_dir_bwd<MAXK, float>(K, z[u], y[u], w[u], yb[u], wb, zb[u]);
        wb0[u] = wb[0];
      }
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K && ok[u]) G[base[u] + c0 + (int64_t)(k - 1) * width] = (ACT == 2) ? zb[u][k] * omega : zb[u][k];
  }
#pragma unroll
  for (int u = 0; u < PPT; ++u) {
    if (!ok[u]) continue;
    if (ACT == 1) G[base[u]] = tanh_finish_bwd<float>(y[u][0], w[u][0], yb[u][0], wb0[u]);
    else G[base[u]] = (yb[u][0] * w[u][0] - wb0[u] * y[u][0]) * omega;
  }
}

// Fourier features: Y[..., 0:m] = sin(Z), Y[..., m:2m] = cos(Z)   (fourier.py:12-16); no adjoint (B is a buffer).
template <int MAXK>
__global__ void sincos_fwd_kernel(const float* __restrict__ Z, float* __restrict__ Y, int64_t n, int m, JetSpec js) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * m) return;
  const int64_t p = idx / m;
  const int f = (int)(idx - p * m);
  const int64_t zb = p * js.ncols * m + f;
  const int64_t yb = p * js.ncols * 2 * m + f;
  float z[MAXK + 1], s[MAXK + 1], c[MAXK + 1];
  z[0] = Z[zb];
  sincosf(z[0], &s[0], &c[0]);
  Y[yb] = s[0];
  Y[yb + m] = c[0];
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
      if (k <= K) z[k] = Z[zb + (int64_t)(js.col0[d] + k - 1) * m];
    sincos_dir_fwd<MAXK, float>(K, z, s, c);
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
      if (k <= K) {
        const int64_t o = yb + (int64_t)(js.col0[d] + k - 1) * 2 * m;
        Y[o] = s[k];
        Y[o + m] = c[k];
      }
  }
}

// ------------------------------------------------------------------ fused first / last layer kernels
// First Linear (K = in_dim <= 4) + activation in one pass: Y = act(W xt + b) jets; the pre-activation is cheap to
// recompute from (x, t), so it is only written when a caller still wants the stash (Zout != nullptr).
template <int ACT, int MAXK>
__global__ void first_act_fwd_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                                     const float* __restrict__ W, const float* __restrict__ b, int out_dim, JetSpec js,
                                     float* __restrict__ Zout, float* __restrict__ Y, float omega) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * out_dim) return;
  const int64_t p = idx / out_dim;
  const int f = (int)(idx - p * out_dim);
  float wr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wr[i] = (i < js.in_dim) ? W[f * js.in_dim + i] : 0.f;
  float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
  z[0] = b ? b[f] : 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    if (i < js.in_dim) z[0] = fmaf(wr[i], load_in(x, t, p, i, js.in_dim), z[0]);
  const int64_t base = p * js.ncols * out_dim + f;
  if (Zout) Zout[base] = z[0];
  if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
  else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
  Y[base] = y[0];
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    float z1 = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) z1 = fmaf(wr[i], js.vec[d][i], z1);
    const int64_t b1 = base + (int64_t)js.col0[d] * out_dim;
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      z[k] = (k == 1) ? z1 : 0.f;
      if (Zout && k <= K) Zout[b1 + (int64_t)(k - 1) * out_dim] = z[k];
      if (ACT == 2) z[k] *= omega;
    }
    if (ACT == 1) tanh_dir_fwd<MAXK, float>(K, z, y, w);
    else sincos_dir_fwd<MAXK, float>(K, z, y, w);
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
      if (k <= K) Y[b1 + (int64_t)(k - 1) * out_dim] = y[k];
  }
}

// Reverse of the above in one pass: G = dL/dY jets -> dW[f,i], db[f] of the first Linear (nothing is written back).
template <int ACT, int MAXK>
__global__ void first_act_bwd_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                                     const float* __restrict__ W, const float* __restrict__ b, int out_dim, JetSpec js,
                                     const float* __restrict__ G, float* __restrict__ gW, float* __restrict__ gb, float omega) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= out_dim) return;
  float wr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wr[i] = (i < js.in_dim) ? W[f * js.in_dim + i] : 0.f;
  const float bf = b ? b[f] : 0.f;
  float z1d[kMaxDirs];
  for (int d = 0; d < js.ndirs; ++d) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) a = fmaf(wr[i], js.vec[d][i], a);
    z1d[d] = a;
  }
  float aw[4] = {0.f, 0.f, 0.f, 0.f};
  float ab = 0.f;
  for (int64_t p = blockIdx.y; p < n; p += gridDim.y) {
    float xin[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) xin[i] = (i < js.in_dim) ? load_in(x, t, p, i, js.in_dim) : 0.f;
    float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
    z[0] = bf;
#pragma unroll
    for (int i = 0; i < 4; ++i) z[0] = fmaf(wr[i], xin[i], z[0]);
    if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
    else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
    const int64_t base = p * js.ncols * out_dim + f;
    yb[0] = G[base];
    float wb0 = 0.f;
    for (int d = 0; d < js.ndirs; ++d) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * out_dim;
#pragma unroll
      for (int k = 1; k <= MAXK; ++k) {
        z[k] = (k == 1) ? z1d[d] : 0.f;
        if (ACT == 2) z[k] *= omega;
        yb[k] = (k <= K) ? G[b1 + (int64_t)(k - 1) * out_dim] : 0.f;
      }
      if (ACT == 1) {
        tanh_dir_fwd<MAXK, float>(K, z, y, w);
        tanh_dir_bwd<MAXK, float>(K, z, y, w, yb, zb, wb0);
      } else {
#pragma unroll
        for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
        wb[0] = wb0;
        sincos_dir_fwd<MAXK, float>(K, z, y, w);
        sincos_dir_bwd<MAXK, float>(K, z, y, w, yb, wb, zb);
        wb0 = wb[0];
      }
      const float g1 = (ACT == 2) ? zb[1] * omega : zb[1];       // only the first-order pre-activation depends on W
#pragma unroll
      for (int i = 0; i < 4; ++i) aw[i] = fmaf(g1, js.vec[d][i], aw[i]);
    }
    const float g0 = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0) : (yb[0] * w[0] - wb0 * y[0]) * omega;
    ab += g0;
#pragma unroll
    for (int i = 0; i < 4; ++i) aw[i] = fmaf(g0, xin[i], aw[i]);
  }
  if (gW) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < js.in_dim) atomicAdd(gW + (int64_t)f * js.in_dim + i, aw[i]);
  }
  if (gb) atomicAdd(gb + f, ab);
}

// Last Linear (N = 1) reverse + adjoint of the activation feeding it, in one pass:
//   dL/dY[p,c,f] = Ub[p,c] W[f];  dW[f] += sum Ub[p,c] Y[p,c,f] (Y recomputed from the stashed Z);  db += sum_p Ub[p,0]
//   Gout = dL/dZ jets.
template <int ACT, int MAXK>
__global__ void last_act_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ Ub, int64_t n, int width,
                                    JetSpec js, const float* __restrict__ W, float* __restrict__ Gout,
                                    float* __restrict__ gW, float* __restrict__ gb, float omega) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = f < width;
  const float wf = active ? W[f] : 0.f;
  float acc = 0.f, accb = 0.f;
  for (int64_t p = blockIdx.y; p < n; p += gridDim.y) {
    const float* ub = Ub + p * js.ncols;
    if (f == 0) accb += ub[0];
    if (!active) continue;
    const int64_t base = p * js.ncols * width + f;
    float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
    z[0] = Z[base];
    if (ACT == 1) { y[0] = tanhf(z[0]); w[0] = 1.f - y[0] * y[0]; }
    else { z[0] *= omega; sincosf(z[0], &y[0], &w[0]); }
    yb[0] = ub[0] * wf;
    acc = fmaf(ub[0], y[0], acc);
    float wb0 = 0.f;
    for (int d = 0; d < js.ndirs; ++d) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * width;
#pragma unroll
      for (int k = 1; k <= MAXK; ++k) {
        if (k <= K) {
          z[k] = Z[b1 + (int64_t)(k - 1) * width];
          if (ACT == 2) z[k] *= omega;
          yb[k] = ub[js.col0[d] + k - 1] * wf;
        } else {
          yb[k] = 0.f;
        }
      }
      if (ACT == 1) tanh_dir_fwd<MAXK, float>(K, z, y, w);
      else sincos_dir_fwd<MAXK, float>(K, z, y, w);
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) acc = fmaf(ub[js.col0[d] + k - 1], y[k], acc);
      if (ACT == 1) {
        tanh_dir_bwd<MAXK, float>(K, z, y, w, yb, zb, wb0);
      } else {
#pragma unroll
        for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
        wb[0] = wb0;
        sincos_dir_bwd<MAXK, float>(K, z, y, w, yb, wb, zb);
        wb0 = wb[0];
      }
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) Gout[b1 + (int64_t)(k - 1) * width] = (ACT == 2) ? zb[k] * omega : zb[k];
    }
    Gout[base] = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0) : (yb[0] * w[0] - wb0 * y[0]) * omega;
  }
  if (active && gW) atomicAdd(gW + f, acc);
  if (f == 0 && gb) atomicAdd(gb, accb);
}

// ------------------------------------------------------------------ LayerNorm jets (one warp per point)
// NPER = ceil(width/32) features per lane.
template <int MAXK, int NPER>
__global__ void layernorm_fwd_kernel(const float* __restrict__ Z, float* __restrict__ Y, int64_t n, int width,
                                     JetSpec js, const float* __restrict__ gamma, const float* __restrict__ beta,
                                     float eps) {
  const int lane = threadIdx.x & 31;
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= n) return;
  const float inv_w = 1.f / (float)width;
  const int64_t base = p * js.ncols * width;
  float c0[NPER], g[NPER];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    c0[i] = (f < width) ? Z[base + f] : 0.f;
    g[i] = (f < width) ? gamma[f] : 0.f;
    sum += c0[i];
  }
  const float mean0 = warp_sum(sum) * inv_w;
  float vs = 0.f;
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    c0[i] = (f < width) ? (c0[i] - mean0) : 0.f;
    vs += c0[i] * c0[i];
  }
  float v[MAXK + 1], s[MAXK + 1];
  v[0] = warp_sum(vs) * inv_w + eps;
  s[0] = 1.f / sqrtf(v[0]);
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    if (f < width) Y[base + f] = g[i] * c0[i] * s[0] + beta[f];
  }
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    const int64_t b1 = base + (int64_t)js.col0[d] * width;
    float c[MAXK + 1][NPER];
#pragma unroll
    for (int i = 0; i < NPER; ++i) c[0][i] = c0[i];
    // every load of the direction is issued before the first reduction needs one (a warp walks its point alone: a load
    // issued behind a shuffle chain is a round trip on its serial path)
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int f = lane + 32 * i;
        c[k][i] = (k <= K && f < width) ? Z[b1 + (int64_t)(k - 1) * width + f] : 0.f;
      }
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      if (k <= K) {
        float sm = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) sm += c[k][i];
        const float mk = warp_sum(sm) * inv_w;
        float vk = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          const int f = lane + 32 * i;
          c[k][i] = (f < width) ? (c[k][i] - mk) : 0.f;
#pragma unroll
          for (int j = 0; j <= k; ++j) vk += c[j][i] * c[k - j][i];
        }
        v[k] = warp_sum(vk) * inv_w;
      }
    }
    rsqrt_dir_fwd<MAXK, float>(K, v, s);
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      if (k <= K) {
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          const int f = lane + 32 * i;
          float a = 0.f;
#pragma unroll
          for (int j = 0; j <= k; ++j) a += c[j][i] * s[k - j];
          if (f < width) Y[b1 + (int64_t)(k - 1) * width + f] = g[i] * a;
        }
      }
    }
  }
}

// Reverse of the above.  G_in = dL/dY, G_out = dL/dZ (distinct buffers); dgamma/dbeta accumulated per block.
// Four resident blocks per SM (128 registers, a few hundred bytes of spills at 8 features per lane) beat the three that the
// kernel's natural 168 registers allow: C3's LayerNorm reverse 16.7 -> 15.1 ms; five blocks (96 registers) lose again (19.5 ms).
#ifndef PINNK_LNB_MINBLOCKS
#define PINNK_LNB_MINBLOCKS 4
#endif
template <int MAXK, int NPER>
__global__ void __launch_bounds__(128, (NPER <= 8) ? PINNK_LNB_MINBLOCKS : 1) layernorm_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ Gin,
                                     float* __restrict__ Gout, int64_t n, int width, JetSpec js,
                                     const float* __restrict__ gamma, float eps,
                                     float* __restrict__ dgamma, float* __restrict__ dbeta) {
  extern __shared__ float red[];   // [2][width]
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_w = 1.f / (float)width;
  for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float g[NPER], dg[NPER], db[NPER];
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    g[i] = (f < width) ? gamma[f] : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  for (int64_t p = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < n;
       p += (int64_t)gridDim.x * warps_per_block) {
    const int64_t base = p * js.ncols * width;
    // ---- recompute order 0
    float c0[NPER], yb0[NPER], cb0[NPER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      c0[i] = (f < width) ? Z[base + f] : 0.f;
      yb0[i] = (f < width) ? Gin[base + f] : 0.f;
      sum += c0[i];
    }
    const float mean0 = warp_sum(sum) * inv_w;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      c0[i] = (f < width) ? (c0[i] - mean0) : 0.f;
      vs += c0[i] * c0[i];
    }
    float v[MAXK + 1], s[MAXK + 1];
    v[0] = warp_sum(vs) * inv_w + eps;
    s[0] = 1.f / sqrtf(v[0]);
    // order-0 direct terms
    float sb0 = 0.f, vb0 = 0.f;
    {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        dg[i] += yb0[i] * c0[i] * s[0];
        db[i] += yb0[i];
        cb0[i] = g[i] * yb0[i] * s[0];
        part += g[i] * yb0[i] * c0[i];
      }
      sb0 = warp_sum(part);
    }
    for (int d = 0; d < js.ndirs; ++d) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * width;
      float c[MAXK + 1][NPER], yb[MAXK + 1][NPER];
#pragma unroll
      for (int i = 0; i < NPER; ++i) { c[0][i] = c0[i]; yb[0][i] = 0.f; }   // yb[0] direct term handled above
      // (all loads of the direction first: see layernorm_fwd_kernel)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          const int f = lane + 32 * i;
          const bool in = k <= K && f < width;
          c[k][i] = in ? Z[b1 + (int64_t)(k - 1) * width + f] : 0.f;
          yb[k][i] = in ? Gin[b1 + (int64_t)(k - 1) * width + f] : 0.f;
        }
#pragma unroll
      for (int k = 1; k <= MAXK; ++k) {
        if (k <= K) {
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) sm += c[k][i];
          const float mk = warp_sum(sm) * inv_w;
          float vk = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            const int f = lane + 32 * i;
            c[k][i] = (f < width) ? (c[k][i] - mk) : 0.f;
#pragma unroll
            for (int j = 0; j <= k; ++j) vk += c[j][i] * c[k - j][i];
          }
          v[k] = warp_sum(vk) * inv_w;
        } else {
          v[k] = 0.f;
        }
      }
      rsqrt_dir_fwd<MAXK, float>(K, v, s);
      // y_k = gamma * sum_i c_i s_{k-i}:   dgamma, cb (direct), sb
      float sb[MAXK + 1], cb[MAXK + 1][NPER];
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) {
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          float a = 0.f;
#pragma unroll
          for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
            if (k <= K) {
              a += yb[k][i] * s[k - m];          // -> cb[m]
              part += g[i] * yb[k][i] * c[k - m][i];  // -> sb[m]
            }
          cb[m][i] = g[i] * a;
        }
        sb[m] = warp_sum(part);
      }
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j <= k; ++j) a += c[j][i] * s[k - j];
            dg[i] += yb[k][i] * a;
          }
        }
      // s recurrence
      float vb[MAXK + 1];
      float sb_in0 = sb[0];
      rsqrt_dir_bwd<MAXK, float>(K, v, s, sb, vb, vb0);
      sb0 += sb[0];
      (void)sb_in0;
      // v_k = mean_f sum_i c_i c_{k-i}  (k >= 1) -> cb[i] += (2/W) sum_{k>=max(i,1)} vb_k c_{k-i}
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) {
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          float a = 0.f;
#pragma unroll
          for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
            if (k <= K) a += vb[k] * c[k - m][i];
          cb[m][i] += 2.f * inv_w * a;
        }
      }
#pragma unroll
      for (int i = 0; i < NPER; ++i) cb0[i] += cb[0][i];
      // z_k = c_k + mean: zb_k = cb_k - mean_f cb_k
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) sm += cb[k][i];
          const float mk = warp_sum(sm) * inv_w;
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            const int f = lane + 32 * i;
            if (f < width) Gout[b1 + (int64_t)(k - 1) * width + f] = cb[k][i] - mk;
          }
        }
    }
    // order-0 closure: s0 = v0^-1/2 ; v0 = mean c0^2 + eps ; c0 = z0 - mean
    vb0 += sb0 * (-0.5f) * s[0] / v[0];
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      cb0[i] += 2.f * inv_w * vb0 * c0[i];
      sm += cb0[i];
    }
    const float m0 = warp_sum(sm) * inv_w;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      if (f < width) Gout[base + f] = cb0[i] - m0;
    }
  }
  // block-level accumulation of dgamma / dbeta
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    if (f < width) {
      atomicAdd(&red[f], dg[i]);
      atomicAdd(&red[width + f], db[i]);
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < width; f += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + f, red[f]);
    if (dbeta) atomicAdd(dbeta + f, red[width + f]);
  }
}

// ------------------------------------------------------------------ LayerNorm + activation in one sweep
// Y = act(LayerNorm(Z) (+ S)): resnet.py:45-65 runs Linear -> LayerNorm -> tanh and Linear -> LayerNorm -> (+ x) -> tanh; as
// separate kernels the LayerNorm output makes a round trip through HBM in the forward pass and is read twice more in the
// reverse pass.  One warp per point as above; the normalised jets stay in registers and go straight into the activation
// recurrences of each of the lane's NPER features.
template <int ACT, int MAXK, int NPER>
__global__ void lnact_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ Y, int64_t n,
                                 int width, JetSpec js, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 float eps, float omega) {
  const int lane = threadIdx.x & 31;
  const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (p >= n) return;
  const float inv_w = 1.f / (float)width;
  const int64_t base = p * js.ncols * width;
  float c0[NPER], g[NPER], y0[NPER], w0[NPER];
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    c0[i] = (f < width) ? Z[base + f] : 0.f;
    g[i] = (f < width) ? gamma[f] : 0.f;
    sum += c0[i];
  }
  const float mean0 = warp_sum(sum) * inv_w;
  float vs = 0.f;
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    c0[i] = (f < width) ? (c0[i] - mean0) : 0.f;
    vs += c0[i] * c0[i];
  }
  float v[MAXK + 1], s[MAXK + 1];
  v[0] = warp_sum(vs) * inv_w + eps;
  s[0] = 1.f / sqrtf(v[0]);
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    if (f < width) {
      float z0 = g[i] * c0[i] * s[0] + beta[f];
      if (S) z0 += S[base + f];
      if (ACT == 1) { y0[i] = tanhf(z0); w0[i] = 1.f - y0[i] * y0[i]; }
      else { z0 *= omega; sincosf(z0, &y0[i], &w0[i]); }
      Y[base + f] = y0[i];
    } else {
      y0[i] = 0.f; w0[i] = 0.f;
    }
  }
  for (int d = 0; d < js.ndirs; ++d) {
    const int K = js.order[d];
    const int64_t b1 = base + (int64_t)js.col0[d] * width;
    float c[MAXK + 1][NPER], sj[MAXK + 1][NPER];
#pragma unroll
    for (int i = 0; i < NPER; ++i) c[0][i] = c0[i];
    // every load of the direction is issued before the first reduction needs one
#pragma unroll
    for (int k = 1; k <= MAXK; ++k)
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int f = lane + 32 * i;
        const bool in = k <= K && f < width;
        c[k][i] = in ? Z[b1 + (int64_t)(k - 1) * width + f] : 0.f;
        sj[k][i] = (in && S) ? S[b1 + (int64_t)(k - 1) * width + f] : 0.f;
      }
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      if (k <= K) {
        float sm = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) sm += c[k][i];
        const float mk = warp_sum(sm) * inv_w;
        float vk = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          const int f = lane + 32 * i;
          c[k][i] = (f < width) ? (c[k][i] - mk) : 0.f;
#pragma unroll
          for (int j = 0; j <= k; ++j) vk += c[j][i] * c[k - j][i];
        }
        v[k] = warp_sum(vk) * inv_w;
      }
    }
    rsqrt_dir_fwd<MAXK, float>(K, v, s);
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      if (f < width) {
        float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
        z[0] = 0.f; y[0] = y0[i]; w[0] = w0[i];
#pragma unroll
        for (int k = 1; k <= MAXK; ++k) {
          if (k <= K) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j <= k; ++j) a += c[j][i] * s[k - j];
            z[k] = g[i] * a + sj[k][i];
            if (ACT == 2) z[k] *= omega;
          } else {
            z[k] = 0.f;
          }
        }
        if (ACT == 1) tanh_dir_fwd<MAXK, float>(K, z, y, w);
        else sincos_dir_fwd<MAXK, float>(K, z, y, w);
#pragma unroll
        for (int k = 1; k <= MAXK; ++k)
          if (k <= K) Y[b1 + (int64_t)(k - 1) * width + f] = y[k];
      }
    }
  }
}

// Reverse of lnact_fwd_kernel in one sweep: Gin (+ Gin2, the adjoint arriving over a skip connection, nullable) = dL/dY;
// Gout = dL/dZ (a buffer distinct from Gin); Gz (nullable, may alias Gin) receives dL/d(pre-activation), which is the adjoint
// of the skip source S.  LayerNorm and activation jets are recomputed from Z (+ S); per direction the activation adjoint
// produces the adjoint of the normalised jets in registers and the LayerNorm reverse consumes it on the spot.  The order-0
// terms, which need the activation's order-0 adjoint of ALL directions, are linear accumulations and run after the loop.
template <int ACT, int MAXK, int NPER>
__global__ void __launch_bounds__(128)
lnact_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, const float* Gin, const float* __restrict__ Gin2,
                 float* Gz, float* __restrict__ Gout, int64_t n, int width, JetSpec js, const float* __restrict__ gamma,
                 const float* __restrict__ beta, float eps, float omega, float* __restrict__ dgamma,
                 float* __restrict__ dbeta) {
  extern __shared__ float red[];   // [2][width]
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const float inv_w = 1.f / (float)width;
  for (int i = threadIdx.x; i < 2 * width; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  float g[NPER], dg[NPER], db[NPER];
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    g[i] = (f < width) ? gamma[f] : 0.f;
    dg[i] = 0.f;
    db[i] = 0.f;
  }
  for (int64_t p = (int64_t)blockIdx.x * warps_per_block + (threadIdx.x >> 5); p < n;
       p += (int64_t)gridDim.x * warps_per_block) {
    const int64_t base = p * js.ncols * width;
    // ---- order 0 of the forward: centred input, 1/sigma, activation value
    float c0[NPER], y0[NPER], w0[NPER], yb0[NPER], wb0[NPER], cb0[NPER];
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      c0[i] = (f < width) ? Z[base + f] : 0.f;
      sum += c0[i];
    }
    const float mean0 = warp_sum(sum) * inv_w;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      c0[i] = (f < width) ? (c0[i] - mean0) : 0.f;
      vs += c0[i] * c0[i];
    }
    float v[MAXK + 1], s[MAXK + 1];
    v[0] = warp_sum(vs) * inv_w + eps;
    s[0] = 1.f / sqrtf(v[0]);
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      if (f < width) {
        float z0 = g[i] * c0[i] * s[0] + beta[f];
        if (S) z0 += S[base + f];
        if (ACT == 1) { y0[i] = tanhf(z0); w0[i] = 1.f - y0[i] * y0[i]; }
        else { z0 *= omega; sincosf(z0, &y0[i], &w0[i]); }
        yb0[i] = Gin[base + f] + (Gin2 ? Gin2[base + f] : 0.f);
      } else {
        y0[i] = 0.f; w0[i] = 0.f; yb0[i] = 0.f;
      }
      wb0[i] = 0.f;
      cb0[i] = 0.f;
    }
    float sb0 = 0.f, vb0 = 0.f;
    for (int d = 0; d < js.ndirs; ++d) {
      const int K = js.order[d];
      const int64_t b1 = base + (int64_t)js.col0[d] * width;
      float c[MAXK + 1][NPER], yb[MAXK + 1][NPER], sj[MAXK + 1][NPER];
#pragma unroll
      for (int i = 0; i < NPER; ++i) { c[0][i] = c0[i]; yb[0][i] = 0.f; }
      // every load of the direction (input jets, skip jets, output adjoints) is issued before the first reduction needs one
      // and before any store of this direction (Gz may alias Gin)
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          const int f = lane + 32 * i;
          const bool in = k <= K && f < width;
          const int64_t o = b1 + (int64_t)(k - 1) * width + f;
          c[k][i] = in ? Z[o] : 0.f;
          sj[k][i] = (in && S) ? S[o] : 0.f;
          yb[k][i] = in ? Gin[o] + (Gin2 ? Gin2[o] : 0.f) : 0.f;
        }
#pragma unroll
      for (int k = 1; k <= MAXK; ++k) {
        if (k <= K) {
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) sm += c[k][i];
          const float mk = warp_sum(sm) * inv_w;
          float vk = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            const int f = lane + 32 * i;
            c[k][i] = (f < width) ? (c[k][i] - mk) : 0.f;
#pragma unroll
            for (int j = 0; j <= k; ++j) vk += c[j][i] * c[k - j][i];
          }
          v[k] = warp_sum(vk) * inv_w;
        } else {
          v[k] = 0.f;
        }
      }
      rsqrt_dir_fwd<MAXK, float>(K, v, s);
      // ---- activation: forward jets from the normalised jets, then its adjoint -> yb[k][i] = dL/d(LayerNorm output jets)
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int f = lane + 32 * i;
        if (f < width) {
          float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], ab[MAXK + 1], zb[MAXK + 1];
          z[0] = 0.f; y[0] = y0[i]; w[0] = w0[i]; ab[0] = yb0[i];
#pragma unroll
          for (int k = 1; k <= MAXK; ++k) {
            if (k <= K) {
              float a = 0.f;
#pragma unroll
              for (int j = 0; j <= k; ++j) a += c[j][i] * s[k - j];
              z[k] = g[i] * a + sj[k][i];
              if (ACT == 2) z[k] *= omega;
              ab[k] = yb[k][i];
            } else {
              z[k] = 0.f; ab[k] = 0.f;
            }
          }
          if (ACT == 1) {
            tanh_dir_fwd<MAXK, float>(K, z, y, w);
            tanh_dir_bwd<MAXK, float>(K, z, y, w, ab, zb, wb0[i]);
          } else {
            float wb[MAXK + 1];
#pragma unroll
            for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
            wb[0] = wb0[i];
            sincos_dir_fwd<MAXK, float>(K, z, y, w);
            sincos_dir_bwd<MAXK, float>(K, z, y, w, ab, wb, zb);
            wb0[i] = wb[0];
          }
          yb0[i] = ab[0];
#pragma unroll
          for (int k = 1; k <= MAXK; ++k)
            yb[k][i] = (k <= K) ? ((ACT == 2) ? zb[k] * omega : zb[k]) : 0.f;
        }
      }
      if (Gz) {
#pragma unroll
        for (int k = 1; k <= MAXK; ++k)
          if (k <= K) {
#pragma unroll
            for (int i = 0; i < NPER; ++i) {
              const int f = lane + 32 * i;
              if (f < width) Gz[b1 + (int64_t)(k - 1) * width + f] = yb[k][i];
            }
          }
      }
      // ---- LayerNorm reverse for this direction (as layernorm_bwd_kernel)
      float sb[MAXK + 1], cb[MAXK + 1][NPER];
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) {
        float part = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          float a = 0.f;
#pragma unroll
          for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
            if (k <= K) {
              a += yb[k][i] * s[k - m];
              part += g[i] * yb[k][i] * c[k - m][i];
            }
          cb[m][i] = g[i] * a;
        }
        sb[m] = warp_sum(part);
      }
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            float a = 0.f;
#pragma unroll
            for (int j = 0; j <= k; ++j) a += c[j][i] * s[k - j];
            dg[i] += yb[k][i] * a;
          }
        }
      float vb[MAXK + 1];
      rsqrt_dir_bwd<MAXK, float>(K, v, s, sb, vb, vb0);
      sb0 += sb[0];
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) {
#pragma unroll
        for (int i = 0; i < NPER; ++i) {
          float a = 0.f;
#pragma unroll
          for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
            if (k <= K) a += vb[k] * c[k - m][i];
          cb[m][i] += 2.f * inv_w * a;
        }
      }
#pragma unroll
      for (int i = 0; i < NPER; ++i) cb0[i] += cb[0][i];
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= K) {
          float sm = 0.f;
#pragma unroll
          for (int i = 0; i < NPER; ++i) sm += cb[k][i];
          const float mk = warp_sum(sm) * inv_w;
#pragma unroll
          for (int i = 0; i < NPER; ++i) {
            const int f = lane + 32 * i;
            if (f < width) Gout[b1 + (int64_t)(k - 1) * width + f] = cb[k][i] - mk;
          }
        }
    }
    // ---- order 0: close the activation adjoint, then the LayerNorm's direct terms and its closure
    {
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int f = lane + 32 * i;
        const float a0 = (ACT == 1) ? tanh_finish_bwd<float>(y0[i], w0[i], yb0[i], wb0[i])
                                    : (yb0[i] * w0[i] - wb0[i] * y0[i]) * omega;
        if (Gz && f < width) Gz[base + f] = a0;
        dg[i] += a0 * c0[i] * s[0];
        db[i] += a0;
        cb0[i] += g[i] * a0 * s[0];
        part += g[i] * a0 * c0[i];
      }
      sb0 += warp_sum(part);
    }
    vb0 += sb0 * (-0.5f) * s[0] / v[0];
    float sm = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      cb0[i] += 2.f * inv_w * vb0 * c0[i];
      sm += cb0[i];
    }
    const float m0 = warp_sum(sm) * inv_w;
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      if (f < width) Gout[base + f] = cb0[i] - m0;
    }
  }
#pragma unroll
  for (int i = 0; i < NPER; ++i) {
    const int f = lane + 32 * i;
    if (f < width) {
      atomicAdd(&red[f], dg[i]);
      atomicAdd(&red[width + f], db[i]);
    }
  }
  __syncthreads();
  for (int f = threadIdx.x; f < width; f += blockDim.x) {
    if (dgamma) atomicAdd(dgamma + f, red[f]);
    if (dbeta) atomicAdd(dbeta + f, red[width + f]);
  }
}

// ------------------------------------------------------------------ last Linear (N = 1)
// U[row] = sum_f W[f] X[row,f] (+ b on the value column).  One warp per row.
__global__ void last_linear_fwd_kernel(const float* __restrict__ X, int64_t rows, int width, int ncols,
                                       const float* __restrict__ W, const float* __restrict__ b,
                                       float* __restrict__ U) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  float a = 0.f;
  for (int f = lane; f < width; f += 32) a = fmaf(W[f], X[row * width + f], a);
  a = warp_sum(a);
  if (lane == 0) U[row] = a + ((b && (row % ncols) == 0) ? b[0] : 0.f);
}

// Xb[row,f] = Ub[row] W[f] ; dW[f] += sum_row Ub[row] X[row,f] ; db += sum_p Ub[p,0]
__global__ void last_linear_bwd_kernel(const float* __restrict__ X, const float* __restrict__ Ub, int64_t rows,
                                       int width, int ncols, const float* __restrict__ W,
                                       float* __restrict__ Xb, float* __restrict__ gW, float* __restrict__ gb) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = f < width;
  const float wf = active ? W[f] : 0.f;
  float acc = 0.f, accb = 0.f;
  for (int64_t row = blockIdx.y; row < rows; row += gridDim.y) {
    const float ub = Ub[row];
    if (active) {
      acc = fmaf(ub, X[row * width + f], acc);
      Xb[row * width + f] = ub * wf;
    }
    if (f == 0 && (row % ncols) == 0) accb += ub;
  }
  if (active && gW) atomicAdd(gW + f, acc);
  if (f == 0 && gb) atomicAdd(gb, accb);
}

// db[f] += sum over value-column rows of G[row,f]    (bias gradient of a hidden Linear)
__global__ void bias_grad_kernel(const float* __restrict__ G, int64_t n, int width, int ncols, float* __restrict__ gb) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= width) return;
  float acc = 0.f;
  for (int64_t p = blockIdx.y; p < n; p += gridDim.y) acc += G[p * ncols * width + f];
  atomicAdd(gb + f, acc);
}

__global__ void add_inplace_kernel(float* __restrict__ A, const float* __restrict__ B, int64_t count) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) A[i] += B[i];
}

// ------------------------------------------------------------------ PDE / loss epilogue
struct SegmentDev {
  PdeDesc pde;
  int loss_kind;
  float huber_delta;
  float weight;        // multiplies rho(e) in the loss sum
  float grad_weight;   // multiplies rho'(e) in the seed (grad_scale[component] * weight); 0 = no seeds
  int64_t row_start, row_count, pair_offset;
  const float* target;
  float* error_out;
  const float* error_grad;   // per-row upstream dL/de (overrides rho' * grad_weight)
  double* loss_slot;
};

// rows are relative to the chunk: U / Ub hold [chunk_rows, C]; `row_lo`.. is the part of the segment in this chunk.
// xs / xs_stride: first spatial coordinate of the chunk's points (only the Black-Scholes residual reads it)
__global__ void epilogue_kernel(const float* __restrict__ U, float* __restrict__ Ub, JetSpec js, SegmentDev sg,
                                int64_t chunk_row0, int64_t lo, int64_t hi, const float* __restrict__ xs, int xs_stride) {
  const int64_t i = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // absolute row of the call
  double part = 0.0;
  if (i < hi) {
    const int64_t r = i - chunk_row0;
    float u[kMaxCols], du[kMaxCols];
    for (int c = 0; c < js.ncols; ++c) u[c] = U[r * js.ncols + c];
    const bool need_x = sg.pde.kind == PDE_BLACK_SCHOLES;
    float e = pde_residual<float>(sg.pde, js, u, du, need_x ? xs[r * xs_stride] : 0.f);
    float du2[kMaxCols];
    if (sg.pair_offset != 0) {
      const int64_t r2 = r + sg.pair_offset;
      for (int c = 0; c < js.ncols; ++c) u[c] = U[r2 * js.ncols + c];
      e -= pde_residual<float>(sg.pde, js, u, du2, need_x ? xs[r2 * xs_stride] : 0.f);
    }
    const int64_t k = i - sg.row_start;
    if (sg.target) e -= sg.target[k];
    if (sg.error_out) sg.error_out[k] = e;
    float drho;
    const float rho = loss_rho<float>(sg.loss_kind, sg.huber_delta, e, &drho);
    part = (double)rho * (double)sg.weight;
    if (Ub != nullptr && (sg.grad_weight != 0.f || sg.error_grad != nullptr)) {
      const float gsc = sg.error_grad ? sg.error_grad[k] : drho * sg.grad_weight;
      for (int c = 0; c < js.ncols; ++c) Ub[r * js.ncols + c] += gsc * du[c];
      if (sg.pair_offset != 0) {
        const int64_t r2 = r + sg.pair_offset;
        for (int c = 0; c < js.ncols; ++c) Ub[r2 * js.ncols + c] -= gsc * du2[c];
      }
    }
  }
  part = warp_sum_d(part);
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = part;
  __syncthreads();
  if (wid == 0) {
    double v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    v = warp_sum_d(v);
    if (lane == 0 && sg.loss_slot) atomicAdd(sg.loss_slot, v);
  }
}

// |r| and (sum|r|, sum r^2, max|r|, count) for the adaptive samplers
__global__ void score_kernel(const float* __restrict__ U, JetSpec js, PdeDesc pde, int64_t rows,
                             float* __restrict__ abs_out, double* __restrict__ stats, const float* __restrict__ xs,
                             int xs_stride) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double s1 = 0.0, s2 = 0.0;
  float mx = 0.f;
  if (r < rows) {
    float u[kMaxCols];
    for (int c = 0; c < js.ncols; ++c) u[c] = U[r * js.ncols + c];
    const float a = fabsf(pde_residual<float>(pde, js, u, nullptr, pde.kind == PDE_BLACK_SCHOLES ? xs[r * xs_stride] : 0.f));
    if (abs_out) abs_out[r] = a;
    s1 = a; s2 = (double)a * a; mx = a;
  }
  s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  __shared__ double sh1[32], sh2[32];
  __shared__ float shm[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) { sh1[wid] = s1; sh2[wid] = s2; shm[wid] = mx; }
  __syncthreads();
  if (wid == 0) {
    const int nw = blockDim.x >> 5;
    s1 = lane < nw ? sh1[lane] : 0.0; s2 = lane < nw ? sh2[lane] : 0.0; mx = lane < nw ? shm[lane] : 0.f;
    s1 = warp_sum_d(s1); s2 = warp_sum_d(s2);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) {
      atomicAdd(stats + 0, s1);
      atomicAdd(stats + 1, s2);
      // max of non-negative doubles == max of their bit patterns
      atomicMax(reinterpret_cast<unsigned long long*>(stats + 2), (unsigned long long)__double_as_longlong((double)mx));
      const int64_t base = (int64_t)blockIdx.x * blockDim.x;
      const int64_t cnt = (rows - base) < (int64_t)blockDim.x ? (rows - base) : (int64_t)blockDim.x;
      atomicAdd(stats + 3, (double)cnt);
    }
  }
}

// ------------------------------------------------------------------ fused optimizer tail (trainer.py:690-694,292-297)
// clip_grad_norm_(max_norm) + Adam with L2 weight decay on the flat gradient, writing the parameters in place.
struct ParamTable {
  float* ptr[64];
  int64_t off[65];      // flat offsets, off[n] = total
  int n;
};

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t count, double* __restrict__ out) {
  double acc = 0.0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (int64_t)gridDim.x * blockDim.x) {
    const double v = g[i];
    acc += v * v;
  }
  acc = warp_sum_d(acc);
  __shared__ double sh[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) sh[wid] = acc;
  __syncthreads();
  if (wid == 0) {
    double v = (lane < (blockDim.x >> 5)) ? sh[lane] : 0.0;
    v = warp_sum_d(v);
    if (lane == 0) atomicAdd(out, v);
  }
}

__global__ void adam_kernel(ParamTable tab, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            const double* __restrict__ sumsq, float max_norm, float lr, float beta1, float beta2, float eps,
                            float weight_decay, float bc1, float bc2_sqrt, const double* __restrict__ dyn) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tab.off[tab.n]) return;
  if (dyn) {                                       // graph-replayable step: step count and lr live on the device
    // bias corrections in double, as torch.optim.Adam computes them (1 - beta ** step in Python floats): in fp32,
    // 1 - 0.999^step is off by ~1e-5 relative for the first steps
    const double step = dyn[0];
    lr = (float)dyn[1];
    bc1 = (float)(1.0 - pow((double)beta1, step));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, step));
  }
  int lo = 0, hi = tab.n - 1;                      // tensor holding flat index i
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tab.off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  float* p = tab.ptr[lo] + (i - tab.off[lo]);
  float clip = 1.f;
  if (max_norm > 0.f) {                            // torch: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
    const float total = (float)sqrt(*sumsq);
    clip = fminf(max_norm / (total + 1e-6f), 1.f);
  }
  float grad = g[i] * clip;
  const float pv = *p;
  if (weight_decay != 0.f) grad = fmaf(weight_decay, pv, grad);
  const float mi = beta1 * m[i] + (1.f - beta1) * grad;        // torch.lerp(m, g, 1 - beta1)
  const float vi = beta2 * v[i] + (1.f - beta2) * grad * grad;
  m[i] = mi;
  v[i] = vi;
  const float denom = sqrtf(vi) / bc2_sqrt + eps;
  *p = pv - (lr / bc1) * (mi / denom);
}

}  // namespace pinnk

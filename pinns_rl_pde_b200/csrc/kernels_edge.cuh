// kernels_edge.cuh -- the two "edge" layers of the network (first Linear: K = in_dim <= 4; last Linear: N = 1) fused
// with their neighbouring activation, specialised for the jet layouts of the BASELINE configs:
//   two directions (x order K0, t order K1), compile-time (K0, K1), so every loop over jet columns is unrolled, every
//   load of an iteration is issued up front, and PPT points are in flight per thread.
// The generic kernels in kernels_ew.cuh (run-time JetSpec loops) stay as the fallback for other layouts.
//   reference: the first/last nn.Linear of FeedForwardNetwork / SIREN / FourierNetwork (feedforward.py:41-54) under
//   torch autograd; here forward jets and the hand-written adjoint.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "jet_math.cuh"
#include "kernels_ew.cuh"

namespace pinnk {

template <int ACT>
__device__ __forceinline__ void act0(float& z0, float omega, float& y0, float& w0) {
  if (ACT == 1) { y0 = tanhf(z0); w0 = 1.f - y0 * y0; }
  else { z0 *= omega; sincosf(z0, &y0, &w0); }
}

// Y[p, :, f] = act(W0 xt_p + b0) jets.  grid.x covers the features (blockDim.x = 128), grid.y strides over points;
// a thread keeps its weight row in registers and walks its points PPT at a time (stores coalesced over f).
template <int ACT, int K0, int K1, int PPT>
__global__ void __launch_bounds__(128)
first_act_fwd_fast_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                          const float* __restrict__ W, const float* __restrict__ b, int out_dim, JetSpec js,
                          float* __restrict__ Y, float omega) {
  constexpr int C = 1 + K0 + K1, MAXK = (K0 > K1 ? K0 : K1);
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= out_dim) return;
  float wr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wr[i] = (i < js.in_dim) ? W[f * js.in_dim + i] : 0.f;
  const float bf = b ? b[f] : 0.f;
  float z1d[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) a = fmaf(wr[i], js.vec[d][i], a);
    z1d[d] = (ACT == 2) ? a * omega : a;
  }
  for (int64_t p0 = (int64_t)blockIdx.y * PPT; p0 < n; p0 += (int64_t)gridDim.y * PPT) {
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const int64_t p = p0 + u;
      if (p < n) {
        float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
        z[0] = bf;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (i < js.in_dim) z[0] = fmaf(wr[i], load_in(x, t, p, i, js.in_dim), z[0]);
        act0<ACT>(z[0], omega, y[0], w[0]);
        float* yp = Y + p * C * out_dim + f;
        yp[0] = y[0];
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          constexpr int dummy = 0; (void)dummy;
          const int KD = d ? K1 : K0, cb = d ? K0 : 0;
          if (KD > 0) {
#pragma unroll
            for (int k = 1; k <= MAXK; ++k) z[k] = (k == 1) ? z1d[d] : 0.f;
            if (ACT == 1) tanh_dir_fwd<MAXK, float>(KD, z, y, w);
            else sincos_dir_fwd<MAXK, float>(KD, z, y, w);
#pragma unroll
            for (int k = 1; k <= MAXK; ++k)
              if (k <= KD) yp[(cb + k) * out_dim] = y[k];
          }
        }
      }
    }
  }
}

// Reverse of the above: G = dL/dY jets -> dW0[f, i], db0[f]; pre-activation recomputed from (x, t).
template <int ACT, int K0, int K1, int PPT>
__global__ void __launch_bounds__(128)
first_act_bwd_fast_kernel(const float* __restrict__ x, const float* __restrict__ t, int64_t n,
                          const float* __restrict__ W, const float* __restrict__ b, int out_dim, JetSpec js,
                          const float* __restrict__ G, float* __restrict__ gW, float* __restrict__ gb, float omega) {
  constexpr int C = 1 + K0 + K1, MAXK = (K0 > K1 ? K0 : K1);
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= out_dim) return;
  float wr[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) wr[i] = (i < js.in_dim) ? W[f * js.in_dim + i] : 0.f;
  const float bf = b ? b[f] : 0.f;
  float z1d[2];
#pragma unroll
  for (int d = 0; d < 2; ++d) {
    float a = 0.f;
#pragma unroll
    for (int i = 0; i < 4; ++i) a = fmaf(wr[i], js.vec[d][i], a);
    z1d[d] = (ACT == 2) ? a * omega : a;
  }
  float aw[4] = {0.f, 0.f, 0.f, 0.f};
  float ab = 0.f;
  for (int64_t p0 = (int64_t)blockIdx.y * PPT; p0 < n; p0 += (int64_t)gridDim.y * PPT) {
    float g[PPT][C];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const float* gp = G + (p0 + u) * C * out_dim + f;
#pragma unroll
      for (int c = 0; c < C; ++c) g[u][c] = (p0 + u < n) ? __ldg(gp + c * out_dim) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const int64_t p = p0 + u;
      if (p < n) {
        float xin[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) xin[i] = (i < js.in_dim) ? load_in(x, t, p, i, js.in_dim) : 0.f;
        float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
        z[0] = bf;
#pragma unroll
        for (int i = 0; i < 4; ++i) z[0] = fmaf(wr[i], xin[i], z[0]);
        act0<ACT>(z[0], omega, y[0], w[0]);
        yb[0] = g[u][0];
        float wb0 = 0.f;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int KD = d ? K1 : K0, cb = d ? K0 : 0;
          if (KD > 0) {
#pragma unroll
            for (int k = 1; k <= MAXK; ++k) {
              z[k] = (k == 1) ? z1d[d] : 0.f;
              yb[k] = (k <= KD) ? g[u][cb + k] : 0.f;
            }
            if (ACT == 1) {
              tanh_dir_fwd<MAXK, float>(KD, z, y, w);
              tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
            } else {
#pragma unroll
              for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
              wb[0] = wb0;
              sincos_dir_fwd<MAXK, float>(KD, z, y, w);
              sincos_dir_bwd<MAXK, float>(KD, z, y, w, yb, wb, zb);
              wb0 = wb[0];
            }
            const float g1 = (ACT == 2) ? zb[1] * omega : zb[1];   // only the first-order pre-activation depends on W
#pragma unroll
            for (int i = 0; i < 4; ++i) aw[i] = fmaf(g1, js.vec[d][i], aw[i]);
          }
        }
        const float g0 = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0) : (yb[0] * w[0] - wb0 * y[0]) * omega;
        ab += g0;
#pragma unroll
        for (int i = 0; i < 4; ++i) aw[i] = fmaf(g0, xin[i], aw[i]);
      }
    }
  }
  if (gW) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
      if (i < js.in_dim) atomicAdd(gW + (int64_t)f * js.in_dim + i, aw[i]);
  }
  if (gb) atomicAdd(gb + f, ab);
}

// Last Linear (N = 1) reverse + adjoint of the activation feeding it:
//   dL/dY[p,c,f] = Ub[p,c] W[f];  dW[f] += sum Ub[p,c] Y[p,c,f] (Y recomputed from the stashed Z);  db += sum_p Ub[p,0];
//   Gout = dL/dZ jets.
// FROMY (tanh only): Z holds the activation OUTPUT jets (no pre-activation stash); z jets via tanh_dir_recover.
template <int ACT, int K0, int K1, int PPT, bool FROMY>
__global__ void __launch_bounds__(128)
last_act_bwd_fast_kernel(const float* __restrict__ Z, const float* __restrict__ Ub, int64_t n, int width,
                         const float* __restrict__ W, float* __restrict__ Gout, float* __restrict__ gW,
                         float* __restrict__ gb, float omega) {
  constexpr int C = 1 + K0 + K1, MAXK = (K0 > K1 ? K0 : K1);
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= width) return;                                   // width % 128 == 0 on this path: whole blocks only
  const float wf = W[f];
  float acc = 0.f, accb = 0.f;
  for (int64_t p0 = (int64_t)blockIdx.y * PPT; p0 < n; p0 += (int64_t)gridDim.y * PPT) {
    float zr[PPT][C], ub[PPT][C];
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const float* zp = Z + (p0 + u) * C * width + f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        zr[u][c] = (p0 + u < n) ? __ldg(zp + c * width) : 0.f;
        ub[u][c] = (p0 + u < n) ? __ldg(Ub + (p0 + u) * C + c) : 0.f;
      }
    }
#pragma unroll
    for (int u = 0; u < PPT; ++u) {
      const int64_t p = p0 + u;
      if (p < n) {
        if (f == 0) accb += ub[u][0];
        float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1], wb[MAXK + 1];
        float inv_w0 = 0.f;
        if constexpr (FROMY) {
          y[0] = zr[u][0];
          w[0] = 1.f - y[0] * y[0];
          inv_w0 = (w[0] > 1e-30f) ? __fdividef(1.f, w[0]) : 0.f;
        } else {
          z[0] = zr[u][0];
          act0<ACT>(z[0], omega, y[0], w[0]);
        }
        (void)inv_w0;
        yb[0] = ub[u][0] * wf;
        acc = fmaf(ub[u][0], y[0], acc);
        float wb0 = 0.f;
        float* gp = Gout + p * C * width + f;
#pragma unroll
        for (int d = 0; d < 2; ++d) {
          const int KD = d ? K1 : K0, cb = d ? K0 : 0;
          if (KD > 0) {
#pragma unroll
            for (int k = 1; k <= MAXK; ++k) {
              if constexpr (FROMY) y[k] = (k <= KD) ? zr[u][cb + k] : 0.f;
              else z[k] = (k <= KD) ? ((ACT == 2) ? zr[u][cb + k] * omega : zr[u][cb + k]) : 0.f;
              yb[k] = (k <= KD) ? ub[u][cb + k] * wf : 0.f;
            }
            if constexpr (FROMY) tanh_dir_recover<MAXK, float>(KD, y, w, z, inv_w0);
            else if (ACT == 1) tanh_dir_fwd<MAXK, float>(KD, z, y, w);
            else sincos_dir_fwd<MAXK, float>(KD, z, y, w);
#pragma unroll
            for (int k = 1; k <= MAXK; ++k)
              if (k <= KD) acc = fmaf(ub[u][cb + k], y[k], acc);
            if (ACT == 1) {
              tanh_dir_bwd<MAXK, float>(KD, z, y, w, yb, zb, wb0);
            } else {
#pragma unroll
              for (int k = 0; k <= MAXK; ++k) wb[k] = 0.f;
              wb[0] = wb0;
              sincos_dir_bwd<MAXK, float>(KD, z, y, w, yb, wb, zb);
              wb0 = wb[0];
            }
#pragma unroll
            for (int k = 1; k <= MAXK; ++k)
              if (k <= KD) gp[(cb + k) * width] = (ACT == 2) ? zb[k] * omega : zb[k];
          }
        }
        gp[0] = (ACT == 1) ? tanh_finish_bwd<float>(y[0], w[0], yb[0], wb0) : (yb[0] * w[0] - wb0 * y[0]) * omega;
      }
    }
  }
  if (gW) atomicAdd(gW + f, acc);
  if (f == 0 && gb) atomicAdd(gb, accb);
}

// U[row] = sum_f W[f] X[row, f] (+ b on value rows), width == 128: a lane owns 4 features (one LDG.128 per row), a warp
// walks RPW rows per step with all loads in flight.
template <int RPW>
__global__ void __launch_bounds__(256)
last_linear_fwd_w128_kernel(const float* __restrict__ X, int64_t rows, int ncols, const float* __restrict__ W,
                            const float* __restrict__ b, float* __restrict__ U) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float4 wv = *reinterpret_cast<const float4*>(W + lane * 4);
  const float bv = b ? b[0] : 0.f;
  for (int64_t r0 = warp * RPW; r0 < rows; r0 += nwarps * RPW) {
    float4 v[RPW];
#pragma unroll
    for (int u = 0; u < RPW; ++u)
      v[u] = (r0 + u < rows) ? __ldg(reinterpret_cast<const float4*>(X + (r0 + u) * 128) + lane) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < RPW; ++u) {
      float a = fmaf(wv.x, v[u].x, fmaf(wv.y, v[u].y, fmaf(wv.z, v[u].z, wv.w * v[u].w)));
      a = warp_sum(a);
      if (lane == 0 && r0 + u < rows) U[r0 + u] = a + (((r0 + u) % ncols) == 0 ? bv : 0.f);
    }
  }
}

// U[row] = sum_{p < parts} u_part[p][row] (+ b on value rows): the fixed-order sum of the per-warp partial output jets the
// fused last hidden layer wrote (tc_linear_act_fwd with w_out).  20 B per row.
__global__ void output_combine_kernel(const float* __restrict__ u_part, int parts, int64_t rows, int ncols,
                                      const float* __restrict__ b, float* __restrict__ U) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float a = 0.f;
  for (int p = 0; p < parts; ++p) a += u_part[(int64_t)p * rows + r];
  U[r] = a + ((b && (r % ncols) == 0) ? b[0] : 0.f);
}

}  // namespace pinnk

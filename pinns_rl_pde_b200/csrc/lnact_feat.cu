// lnact_feat.cu -- LayerNorm + tanh on derivative jets, forward and reverse, one THREAD per feature pair.
//
// resnet.py:45-65 runs Linear -> LayerNorm -> tanh and Linear -> LayerNorm -> (+ x) -> tanh.  The warp-per-point kernels of
// kernels_ew.cuh (layernorm_*_kernel, lnact_*_kernel) give a lane 8 features of a 256-wide layer: ~170 registers, 16 warps per
// SM, and a point's ~20 cross-feature sums are 20 shuffle chains on one warp's serial path.  Here a block of width / 2
// threads walks points; a thread owns two adjacent features (one float2 per jet column), so every per-feature recurrence is
// a handful of registers, and ALL cross-feature sums a phase needs are reduced together (block_multi_sum: one transposing
// shuffle reduction per warp, one shared-memory exchange, one barrier):
//   forward  2 rounds: column sums (means) | centred second moments v_k
//   reverse  3 rounds: the same two, then every sum the LayerNorm adjoint needs (T_c = sum gamma yb_c, sb0, sb_d[m]);
//            the means of the adjoint jets follow from T_c analytically (the centred jets have zero mean), so no further
//            reduction is needed after the scalar 1/sigma recurrence has been reversed.
// The scalar recurrences (rsqrt_dir_fwd / rsqrt_dir_bwd) run redundantly in every thread.
#include "lnact_api.h"

#include <type_traits>

#include "jet_math.cuh"

namespace pinnk {
namespace {

// a[v], v < NVP: per-lane partial sums.  Returns the warp total of value (lane / (32 / NVP)).  Each level sends half of the
// values a lane still carries to the partner lane and keeps the other half: NVP - 1 + log2(32 / NVP) shuffles instead of 5 NVP.
template <int NVP>
__device__ __forceinline__ float warp_multi_sum(float (&a)[NVP], const int lane) {
  static_assert(NVP == 8 || NVP == 16, "values per reduction");
  constexpr int LV = (NVP == 16) ? 4 : 3;
#pragma unroll
  for (int lv = 0; lv < LV; ++lv) {
    const int S = 16 >> lv, n = NVP >> lv;
    const bool upper = (lane & S) != 0;
#pragma unroll
    for (int i = 0; i < NVP / 2; ++i)
      if (i < n / 2) {
        const float lo = a[i], hi = a[i + n / 2];
        const float send = upper ? lo : hi, keep = upper ? hi : lo;
        a[i] = keep + __shfl_xor_sync(0xffffffffu, send, S);
      }
  }
#pragma unroll
  for (int S = (16 >> LV); S >= 1; S >>= 1) a[0] += __shfl_xor_sync(0xffffffffu, a[0], S);
  return a[0];
}

// a[v] (per-thread partials) -> a[v] = block totals, in every thread.  `part` holds NVP * NW floats and must not be the buffer
// of the previous call (the callers alternate between two): one barrier per call is then enough.  Fixed summation order.
template <int NVP, int NW>
__device__ __forceinline__ void block_multi_sum(float (&a)[NVP], float* __restrict__ part, const int lane, const int warp) {
  constexpr int LPV = 32 / NVP;
  const float w = warp_multi_sum<NVP>(a, lane);
  if constexpr (NW == 1) {
#pragma unroll
    for (int v = 0; v < NVP; ++v) a[v] = __shfl_sync(0xffffffffu, w, v * LPV);
  } else {
    if ((lane % LPV) == 0) part[(lane / LPV) * NW + warp] = w;
    __syncthreads();
    constexpr int NL = (NVP * NW + 31) / 32;
    float t[NL];
#pragma unroll
    for (int j = 0; j < NL; ++j) {
      const int idx = lane + 32 * j;
      t[j] = (idx < NVP * NW) ? part[idx] : 0.f;
#pragma unroll
      for (int S = NW / 2; S >= 1; S >>= 1) t[j] += __shfl_xor_sync(0xffffffffu, t[j], S);
    }
#pragma unroll
    for (int v = 0; v < NVP; ++v) a[v] = __shfl_sync(0xffffffffu, t[(v * NW) / 32], (v * NW) % 32);
  }
}

template <int K0, int K1>
struct Jets {
  static constexpr int C = 1 + K0 + K1;
  static constexpr int KM = (K0 > K1) ? K0 : K1;
  static constexpr int MAXK = (KM > 0) ? KM : 1;
};
__device__ __forceinline__ float comp(const float2 v, const int e) { return e ? v.y : v.x; }
__device__ __forceinline__ void set_comp(float2& v, const int e, const float x) { if (e) v.y = x; else v.x = x; }

// centred jets of the point and the sums of their second moments: on return c[] is centred and a[0] = sum c0^2,
// a[CB_d + k - 1] = sum_f sum_j c_j c_{k-j} (block totals)
template <int K0, int K1, int NW>
__device__ __forceinline__ void centre_and_moments(float2 (&c)[1 + K0 + K1], float (&a)[8], float* part0, float* part1,
                                                   const int lane, const int warp, const float inv_w) {
  constexpr int C = 1 + K0 + K1, MAXK = Jets<K0, K1>::MAXK;
  static_assert(C <= 8, "jet columns");
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = (i < C) ? c[i < C ? i : 0].x + c[i < C ? i : 0].y : 0.f;
  block_multi_sum<8, NW>(a, part0, lane, warp);
#pragma unroll
  for (int i = 0; i < C; ++i) {
    const float m = a[i] * inv_w;
    c[i].x -= m;
    c[i].y -= m;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = 0.f;
  a[0] = c[0].x * c[0].x + c[0].y * c[0].y;
  auto moments = [&](auto dtag) {
    constexpr int D = decltype(dtag)::value;
    constexpr int KD = D ? K1 : K0, CB = D ? 1 + K0 : 1;
    if constexpr (KD > 0) {
      float2 cj[MAXK + 1];
      cj[0] = c[0];
#pragma unroll
      for (int k = 1; k <= MAXK; ++k) cj[k] = (k <= KD) ? c[(CB + k - 1 < C) ? CB + k - 1 : 0] : make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 1; k <= MAXK; ++k)
        if (k <= KD) {
          float vx = 0.f, vy = 0.f;
#pragma unroll
          for (int j = 0; j <= k; ++j) {
            vx = fmaf(cj[j].x, cj[k - j].x, vx);
            vy = fmaf(cj[j].y, cj[k - j].y, vy);
          }
          a[(CB + k - 1 < 8) ? CB + k - 1 : 0] = vx + vy;
        }
    }
  };
  moments(std::integral_constant<int, 0>());
  moments(std::integral_constant<int, 1>());
  block_multi_sum<8, NW>(a, part1, lane, warp);
}

template <int K0, int K1, int NW>
__global__ void __launch_bounds__(NW * 32)
lnact_feat_fwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, float* __restrict__ Y, int64_t n,
                      const float* __restrict__ gamma, const float* __restrict__ beta, float eps) {
  constexpr int C = 1 + K0 + K1, W2 = NW * 32, MAXK = Jets<K0, K1>::MAXK;
  __shared__ float part[2][8 * NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (scalar loads: a parameter tensor may be a view at any 4-byte offset of a caller's flat buffer)
  const float2 g = make_float2(__ldg(gamma + 2 * tid), __ldg(gamma + 2 * tid + 1));
  const float2 bt = make_float2(__ldg(beta + 2 * tid), __ldg(beta + 2 * tid + 1));
  const float inv_w = 1.f / (float)(2 * W2);
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t o = p * (int64_t)(C * W2) + tid;
    const float2* const zp = reinterpret_cast<const float2*>(Z) + o;
    float2* const yp = reinterpret_cast<float2*>(Y) + o;
    float2 c[C], sj[C];
#pragma unroll
    for (int i = 0; i < C; ++i) c[i] = zp[i * W2];
    if (S != nullptr) {
      const float2* const sp = reinterpret_cast<const float2*>(S) + o;
#pragma unroll
      for (int i = 0; i < C; ++i) sj[i] = sp[i * W2];
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i) sj[i] = make_float2(0.f, 0.f);
    }
    float a[8];
    centre_and_moments<K0, K1, NW>(c, a, part[0], part[1], lane, warp, inv_w);
    const float v0 = a[0] * inv_w + eps;
    const float s0 = 1.f / sqrtf(v0);
    float2 y0, w0;
    {
      const float zx = g.x * c[0].x * s0 + bt.x + sj[0].x, zy = g.y * c[0].y * s0 + bt.y + sj[0].y;
      y0.x = tanhf(zx); y0.y = tanhf(zy);
      w0.x = 1.f - y0.x * y0.x; w0.y = 1.f - y0.y * y0.y;
      yp[0] = y0;
    }
    auto direction = [&](auto dtag) {
      constexpr int D = decltype(dtag)::value;
      constexpr int KD = D ? K1 : K0, CB = D ? 1 + K0 : 1;
      if constexpr (KD > 0) {
        float v[MAXK + 1], s[MAXK + 1];
        v[0] = v0; s[0] = s0;
#pragma unroll
        for (int k = 1; k <= MAXK; ++k) v[k] = (k <= KD) ? a[(CB + k - 1 < 8) ? CB + k - 1 : 0] * inv_w : 0.f;
        rsqrt_dir_fwd<MAXK, float>(KD, v, s);
        float2 out[MAXK + 1];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1];
          z[0] = 0.f; y[0] = comp(y0, e); w[0] = comp(w0, e);
          const float c0e = comp(c[0], e), ge = comp(g, e);
#pragma unroll
          for (int k = 1; k <= MAXK; ++k) {
            if (k <= KD) {
              float acc = c0e * s[k];
#pragma unroll
              for (int j = 1; j <= k; ++j) acc = fmaf(comp(c[(CB + j - 1 < C) ? CB + j - 1 : 0], e), s[k - j], acc);
              z[k] = ge * acc + comp(sj[(CB + k - 1 < C) ? CB + k - 1 : 0], e);
            } else {
              z[k] = 0.f;
            }
          }
          tanh_dir_fwd<MAXK, float>(KD, z, y, w);
#pragma unroll
          for (int k = 1; k <= MAXK; ++k)
            if (k <= KD) set_comp(out[k], e, y[k]);
        }
#pragma unroll
        for (int k = 1; k <= MAXK; ++k)
          if (k <= KD) yp[(CB + k - 1) * W2] = out[k];
      }
    };
    direction(std::integral_constant<int, 0>());
    direction(std::integral_constant<int, 1>());
  }
}

#ifndef PINNK_LNF_BWD_THREADS_PER_SM
#define PINNK_LNF_BWD_THREADS_PER_SM 1024
#endif
template <int K0, int K1, int NW>
__global__ void __launch_bounds__(NW * 32, PINNK_LNF_BWD_THREADS_PER_SM / (NW * 32))
lnact_feat_bwd_kernel(const float* __restrict__ Z, const float* __restrict__ S, const float* Gin, const float* __restrict__ Gin2,
                      float* Gz, float* __restrict__ Gout, int64_t n, const float* __restrict__ gamma,
                      const float* __restrict__ beta, float eps, float* __restrict__ dgamma, float* __restrict__ dbeta) {
  constexpr int C = 1 + K0 + K1, W2 = NW * 32, MAXK = Jets<K0, K1>::MAXK;
  // third round: T_c (C) | sb0 direct term (1) | sb of direction 0 (K0 + 1) | sb of direction 1 (K1 + 1, if any)
  constexpr int I_SB0 = C, I_SBX = C + 1, I_SBT = C + 1 + (K0 + 1), NV3 = I_SBT + ((K1 > 0) ? K1 + 1 : 0);
  static_assert(NV3 <= 16, "sums of the third round");
  __shared__ float part[2][16 * NW];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // (scalar loads: a parameter tensor may be a view at any 4-byte offset of a caller's flat buffer)
  const float2 g = make_float2(__ldg(gamma + 2 * tid), __ldg(gamma + 2 * tid + 1));
  const float2 bt = make_float2(__ldg(beta + 2 * tid), __ldg(beta + 2 * tid + 1));
  const float inv_w = 1.f / (float)(2 * W2);
  float2 dg = make_float2(0.f, 0.f), db = make_float2(0.f, 0.f);
  int tog = 0;                   // three rounds per point: the two buffers keep alternating across points
  for (int64_t p = blockIdx.x; p < n; p += gridDim.x) {
    const int64_t o = p * (int64_t)(C * W2) + tid;
    const float2* const zp = reinterpret_cast<const float2*>(Z) + o;
    float2 c[C], yb[C];          // yb: dL/dY on the way in, dL/d(LayerNorm output jets) after the activation adjoint
    // every load of the point precedes its first store (Gz may alias Gin; a thread reads and writes the same elements)
#pragma unroll
    for (int i = 0; i < C; ++i) c[i] = zp[i * W2];
    {
      const float2* const gp = reinterpret_cast<const float2*>(Gin) + o;
#pragma unroll
      for (int i = 0; i < C; ++i) yb[i] = gp[i * W2];
      if (Gin2 != nullptr) {
        const float2* const gp2 = reinterpret_cast<const float2*>(Gin2) + o;
#pragma unroll
        for (int i = 0; i < C; ++i) {
          const float2 t = gp2[i * W2];
          yb[i].x += t.x;
          yb[i].y += t.y;
        }
      }
    }
    float2 sj[C];
    if (S != nullptr) {
      const float2* const sp = reinterpret_cast<const float2*>(S) + o;
#pragma unroll
      for (int i = 0; i < C; ++i) sj[i] = sp[i * W2];
    } else {
#pragma unroll
      for (int i = 0; i < C; ++i) sj[i] = make_float2(0.f, 0.f);
    }
    float a8[8];
    centre_and_moments<K0, K1, NW>(c, a8, part[tog], part[tog ^ 1], lane, warp, inv_w);
    const float v0 = a8[0] * inv_w + eps;
    const float s0 = 1.f / sqrtf(v0);
    // 1/sigma jets of both directions (kept for the reverse)
    float vx[MAXK + 1], sx[MAXK + 1], vt[MAXK + 1], st[MAXK + 1];
    vx[0] = v0; sx[0] = s0; vt[0] = v0; st[0] = s0;
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      vx[k] = (k <= K0) ? a8[(k < 8) ? k : 0] * inv_w : 0.f;
      vt[k] = (k <= K1) ? a8[(K0 + k < 8) ? K0 + k : 0] * inv_w : 0.f;
      sx[k] = 0.f; st[k] = 0.f;
    }
    if constexpr (K0 > 0) rsqrt_dir_fwd<MAXK, float>(K0, vx, sx);
    if constexpr (K1 > 0) rsqrt_dir_fwd<MAXK, float>(K1, vt, st);
    // ---- activation: forward jets from the normalised jets, then its adjoint: yb <- dL/d(LayerNorm output jets)
    float a16[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a16[i] = 0.f;
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float c0e = comp(c[0], e), ge = comp(g, e);
      const float z0 = ge * c0e * s0 + comp(bt, e) + comp(sj[0], e);
      const float y0 = tanhf(z0), w0 = 1.f - y0 * y0;
      float yb0 = comp(yb[0], e), wb0 = 0.f;
      auto act_dir = [&](auto dtag) {
        constexpr int D = decltype(dtag)::value;
        constexpr int KD = D ? K1 : K0, CB = D ? 1 + K0 : 1;
        if constexpr (KD > 0) {
          float z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], ab[MAXK + 1], zb[MAXK + 1];
          z[0] = 0.f; y[0] = y0; w[0] = w0; ab[0] = yb0;
#pragma unroll
          for (int k = 1; k <= MAXK; ++k) {
            if (k <= KD) {
              const float* const s = D ? st : sx;
              float acc = c0e * s[k];
#pragma unroll
              for (int j = 1; j <= k; ++j) acc = fmaf(comp(c[(CB + j - 1 < C) ? CB + j - 1 : 0], e), s[k - j], acc);
              z[k] = ge * acc + comp(sj[(CB + k - 1 < C) ? CB + k - 1 : 0], e);
              ab[k] = comp(yb[(CB + k - 1 < C) ? CB + k - 1 : 0], e);
            } else {
              z[k] = 0.f; ab[k] = 0.f;
            }
          }
          tanh_dir_fwd<MAXK, float>(KD, z, y, w);
          tanh_dir_bwd<MAXK, float>(KD, z, y, w, ab, zb, wb0);
          yb0 = ab[0];
#pragma unroll
          for (int k = 1; k <= MAXK; ++k)
            if (k <= KD) set_comp(yb[(CB + k - 1 < C) ? CB + k - 1 : 0], e, zb[k]);
        }
      };
      act_dir(std::integral_constant<int, 0>());
      act_dir(std::integral_constant<int, 1>());
      const float a0 = tanh_finish_bwd<float>(y0, w0, yb0, wb0);
      set_comp(yb[0], e, a0);
      // ---- partial sums of the LayerNorm adjoint
#pragma unroll
      for (int i = 0; i < C; ++i) a16[i] = fmaf(ge, comp(yb[i], e), a16[i]);
      a16[I_SB0] = fmaf(ge * a0, c0e, a16[I_SB0]);
      auto sb_dir = [&](auto dtag) {
        constexpr int D = decltype(dtag)::value;
        constexpr int KD = D ? K1 : K0, CB = D ? 1 + K0 : 1, IB = D ? I_SBT : I_SBX;
        if constexpr (KD > 0) {
#pragma unroll
          for (int m = 0; m <= MAXK; ++m)
            if (m <= KD) {
              float acc = 0.f;
#pragma unroll
              for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
                if (k <= KD) {
                  const float ckm = (k - m == 0) ? c0e : comp(c[(CB + k - m - 1 < C && k - m >= 1) ? CB + k - m - 1 : 0], e);
                  acc = fmaf(comp(yb[(CB + k - 1 < C) ? CB + k - 1 : 0], e), ckm, acc);
                }
              a16[(IB + m < 16) ? IB + m : 0] = fmaf(ge, acc, a16[(IB + m < 16) ? IB + m : 0]);
            }
        }
      };
      sb_dir(std::integral_constant<int, 0>());
      sb_dir(std::integral_constant<int, 1>());
    }
    if (Gz != nullptr) {
      float2* const zq = reinterpret_cast<float2*>(Gz) + o;
#pragma unroll
      for (int i = 0; i < C; ++i) zq[i * W2] = yb[i];
    }
    block_multi_sum<16, NW>(a16, part[tog], lane, warp);
    tog ^= 1;
    // ---- scalar reverse of the 1/sigma recurrences (every thread)
    float sb0 = a16[I_SB0], vb0 = 0.f;
    float vbx[MAXK + 1], vbt[MAXK + 1];
#pragma unroll
    for (int k = 0; k <= MAXK; ++k) { vbx[k] = 0.f; vbt[k] = 0.f; }
    if constexpr (K0 > 0) {
      float sb[MAXK + 1];
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) sb[m] = (m <= K0) ? a16[(I_SBX + m < 16) ? I_SBX + m : 0] : 0.f;
      rsqrt_dir_bwd<MAXK, float>(K0, vx, sx, sb, vbx, vb0);
      sb0 += sb[0];
    }
    if constexpr (K1 > 0) {
      float sb[MAXK + 1];
#pragma unroll
      for (int m = 0; m <= MAXK; ++m) sb[m] = (m <= K1) ? a16[(I_SBT + m < 16) ? I_SBT + m : 0] : 0.f;
      rsqrt_dir_bwd<MAXK, float>(K1, vt, st, sb, vbt, vb0);
      sb0 += sb[0];
    }
    vb0 += sb0 * (-0.5f) * s0 / v0;
    // means of the adjoint jets: mean_f cb_k = (1/W) sum_{k' >= k} s_{k'-k} T_{k'} (the c terms have zero mean)
    float mean0 = s0 * a16[0];
    float meanx[MAXK + 1], meant[MAXK + 1];
#pragma unroll
    for (int k = 1; k <= MAXK; ++k) {
      float mx = 0.f, mt = 0.f;
#pragma unroll
      for (int kk = k; kk <= MAXK; ++kk) {
        if (kk <= K0) mx = fmaf(sx[kk - k], a16[(kk < 16) ? kk : 0], mx);
        if (kk <= K1) mt = fmaf(st[kk - k], a16[(K0 + kk < 16) ? K0 + kk : 0], mt);
      }
      meanx[k] = mx * inv_w;
      meant[k] = mt * inv_w;
      if (k <= K0) mean0 = fmaf(sx[k], a16[(k < 16) ? k : 0], mean0);
      if (k <= K1) mean0 = fmaf(st[k], a16[(K0 + k < 16) ? K0 + k : 0], mean0);
    }
    meanx[0] = 0.f; meant[0] = 0.f;
    mean0 *= inv_w;
    // ---- per-feature outputs
    float2* const op = reinterpret_cast<float2*>(Gout) + o;
    float2 out[C];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const float c0e = comp(c[0], e), ge = comp(g, e), a0 = comp(yb[0], e);
      float cb0 = ge * a0 * s0;
      float dge = a0 * c0e * s0;
      auto out_dir = [&](auto dtag) {
        constexpr int D = decltype(dtag)::value;
        constexpr int KD = D ? K1 : K0, CB = D ? 1 + K0 : 1;
        if constexpr (KD > 0) {
          const float* const s = D ? st : sx;
          const float* const vb = D ? vbt : vbx;
          const float* const mean = D ? meant : meanx;
          float cj[MAXK + 1], ybk[MAXK + 1];
          cj[0] = c0e; ybk[0] = 0.f;
#pragma unroll
          for (int k = 1; k <= MAXK; ++k) {
            cj[k] = (k <= KD) ? comp(c[(CB + k - 1 < C) ? CB + k - 1 : 0], e) : 0.f;
            ybk[k] = (k <= KD) ? comp(yb[(CB + k - 1 < C) ? CB + k - 1 : 0], e) : 0.f;
          }
#pragma unroll
          for (int m = 0; m <= MAXK; ++m)
            if (m <= KD) {
              float ay = 0.f, av = 0.f;
#pragma unroll
              for (int k = (m > 1 ? m : 1); k <= MAXK; ++k)
                if (k <= KD) {
                  ay = fmaf(ybk[k], s[k - m], ay);
                  av = fmaf(vb[k], cj[k - m], av);
                }
              const float cbm = ge * ay + 2.f * inv_w * av;
              if (m == 0) cb0 += cbm;
              else set_comp(out[(CB + m - 1 < C) ? CB + m - 1 : 0], e, cbm - mean[m]);
            }
#pragma unroll
          for (int k = 1; k <= MAXK; ++k)
            if (k <= KD) {
              float acc = 0.f;
#pragma unroll
              for (int j = 0; j <= k; ++j) acc = fmaf(cj[j], s[k - j], acc);
              dge = fmaf(ybk[k], acc, dge);
            }
        }
      };
      out_dir(std::integral_constant<int, 0>());
      out_dir(std::integral_constant<int, 1>());
      cb0 += 2.f * inv_w * vb0 * c0e;
      set_comp(out[0], e, cb0 - mean0);
      if (e) { dg.y += dge; db.y += a0; } else { dg.x += dge; db.x += a0; }
    }
#pragma unroll
    for (int i = 0; i < C; ++i) op[i * W2] = out[i];
  }
  if (dgamma != nullptr) { atomicAdd(dgamma + 2 * tid, dg.x); atomicAdd(dgamma + 2 * tid + 1, dg.y); }
  if (dbeta != nullptr) { atomicAdd(dbeta + 2 * tid, db.x); atomicAdd(dbeta + 2 * tid + 1, db.y); }
}

template <typename F>
bool dispatch_jets(int k0, int k1, F&& f) {
#define PK_LN_CASE(A, B) if (k0 == A && k1 == B) { f(std::integral_constant<int, A>(), std::integral_constant<int, B>()); return true; }
  PK_LN_CASE(1, 0) PK_LN_CASE(2, 0) PK_LN_CASE(1, 1) PK_LN_CASE(2, 1) PK_LN_CASE(3, 1) PK_LN_CASE(4, 1) PK_LN_CASE(2, 2)
#undef PK_LN_CASE
  return false;
}

template <typename Kern>
int grid_for(Kern kern, int threads, int64_t n, int sm_count, int* cache) {
  if (*cache <= 0) {
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, 0) != cudaSuccess || occ < 1) { cudaGetLastError(); occ = 2; }
    *cache = occ;
  }
  int64_t blocks = (int64_t)sm_count * (*cache);
  if (blocks > n) blocks = n;
  return (int)(blocks < 1 ? 1 : blocks);
}

template <int K0, int K1, int NW>
int launch_fwd(const float* Z, const float* S, float* Y, int64_t n, const float* gamma, const float* beta, float eps,
               int sm_count, cudaStream_t st) {
  static int occ = 0;
  auto kern = lnact_feat_fwd_kernel<K0, K1, NW>;
  const int grid = grid_for(kern, NW * 32, n, sm_count, &occ);
  kern<<<grid, NW * 32, 0, st>>>(Z, S, Y, n, gamma, beta, eps);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
template <int K0, int K1, int NW>
int launch_bwd(const float* Z, const float* S, const float* Gin, const float* Gin2, float* Gz, float* Gout, int64_t n,
               const float* gamma, const float* beta, float eps, float* dgamma, float* dbeta, int sm_count, cudaStream_t st) {
  static int occ = 0;
  auto kern = lnact_feat_bwd_kernel<K0, K1, NW>;
  const int grid = grid_for(kern, NW * 32, n, sm_count, &occ);
  kern<<<grid, NW * 32, 0, st>>>(Z, S, Gin, Gin2, Gz, Gout, n, gamma, beta, eps, dgamma, dbeta);
  return cudaGetLastError() == cudaSuccess ? 0 : -1;
}

}  // namespace

bool lnact_feat_supported(int width, int k0, int k1) {
  if (width != 128 && width != 256 && width != 512) return false;
  return dispatch_jets(k0, k1, [](auto, auto) {});
}

int lnact_feat_fwd(const float* Z, const float* S, float* Y, int64_t n, int width, int k0, int k1, const float* gamma,
                   const float* beta, float eps, int sm_count, cudaStream_t st) {
  if (n < 1 || !lnact_feat_supported(width, k0, k1)) return 1;
  if (((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(Y)) & 7) != 0) return -1;   // float2 rows
  int rc = -1;
  dispatch_jets(k0, k1, [&](auto a, auto b) {
    constexpr int A = decltype(a)::value, B = decltype(b)::value;
    if (width == 128) rc = launch_fwd<A, B, 2>(Z, S, Y, n, gamma, beta, eps, sm_count, st);
    else if (width == 256) rc = launch_fwd<A, B, 4>(Z, S, Y, n, gamma, beta, eps, sm_count, st);
    else rc = launch_fwd<A, B, 8>(Z, S, Y, n, gamma, beta, eps, sm_count, st);
  });
  return rc;
}

int lnact_feat_bwd(const float* Z, const float* S, const float* Gin, const float* Gin2, float* Gz, float* Gout, int64_t n,
                   int width, int k0, int k1, const float* gamma, const float* beta, float eps, float* dgamma, float* dbeta,
                   int sm_count, cudaStream_t st) {
  if (n < 1 || !lnact_feat_supported(width, k0, k1)) return 1;
  if (((reinterpret_cast<uintptr_t>(Z) | reinterpret_cast<uintptr_t>(S) | reinterpret_cast<uintptr_t>(Gin) | reinterpret_cast<uintptr_t>(Gin2) |
        reinterpret_cast<uintptr_t>(Gz) | reinterpret_cast<uintptr_t>(Gout)) & 7) != 0) return -1;                                    // float2 rows
  int rc = -1;
  dispatch_jets(k0, k1, [&](auto a, auto b) {
    constexpr int A = decltype(a)::value, B = decltype(b)::value;
    if (width == 128) rc = launch_bwd<A, B, 2>(Z, S, Gin, Gin2, Gz, Gout, n, gamma, beta, eps, dgamma, dbeta, sm_count, st);
    else if (width == 256) rc = launch_bwd<A, B, 4>(Z, S, Gin, Gin2, Gz, Gout, n, gamma, beta, eps, dgamma, dbeta, sm_count, st);
    else rc = launch_bwd<A, B, 8>(Z, S, Gin, Gin2, Gz, Gout, n, gamma, beta, eps, dgamma, dbeta, sm_count, st);
  });
  return rc;
}

}  // namespace pinnk

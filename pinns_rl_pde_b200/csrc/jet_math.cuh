// jet_math.cuh -- per-element Taylor-mode ("jet") recurrences and their adjoints.
//
// A jet of one scalar feature at one collocation point is
//     a_0                      value
//     a_{d,1..K_d}             normalised Taylor coefficients along direction d
// (d^k/ds^k = k! a_k).  All directions share a_0.  The functions below process ONE
// direction at a time against the shared order-0 quantities so that every array is a
// compile-time-indexed register array (MAXK = largest order in the spec, <= 4).
//
// Replaces, for the hot path, what torch autograd-of-autograd computes in the reference:
//   tanh            pinnrl/neural_networks/feedforward.py:41-54 + pde_base.py:661-732
//   sin(omega z)    pinnrl/neural_networks/siren.py:36-46
//   [sin,cos](xB)   pinnrl/neural_networks/fourier.py:12-16
//   LayerNorm       pinnrl/neural_networks/resnet.py:45-53 (scalar s = v^-1/2 recurrence here;
//                   the cross-feature reductions live in the kernel)
//   PDE residuals   pinnrl/pdes/{heat,burgers,kdv,allen_cahn,cahn_hilliard}*.py compute_residual
//
// Everything is __host__ __device__ and templated on the scalar type so tests/ can
// compile this header with g++ and check it in fp64 against the oracle on CPU.
#pragma once

#ifdef __CUDACC__
#define PK_HD __host__ __device__ __forceinline__
#else
#define PK_HD inline
#endif

#include <math.h>

namespace pinnk {

constexpr int kMaxOrder = 4;
constexpr int kMaxDirs = 5;
constexpr int kMaxCols = 1 + kMaxDirs * kMaxOrder;  // 21

// Jet layout shared by host and device: column 0 is the value, then K_d columns per direction.
struct JetSpec {
  int ndirs;
  int ncols;                 // 1 + sum(order)
  int in_dim;                // network input dimension (x..., t)
  int order[kMaxDirs];
  int col0[kMaxDirs];        // column of a_{d,1}
  float vec[kMaxDirs][4];    // direction vector in input space
};

// ----------------------------------------------------------------------------- tanh
// forward along one direction: given z[1..K] and y[0], w[0] = 1 - y0^2, fill y[1..K], w[1..K].
template <int MAXK, typename T>
PK_HD void tanh_dir_fwd(int K, const T (&z)[MAXK + 1], T (&y)[MAXK + 1], T (&w)[MAXK + 1]) {
#pragma unroll
  for (int k = 1; k <= MAXK; ++k) {
    if (k <= K) {
      T acc = T(0);
#pragma unroll
      for (int j = 1; j <= k; ++j) acc += T(j) * z[j] * w[k - j];
      y[k] = acc * (T(1) / T(k));
      T wk = T(0);
#pragma unroll
      for (int i = 0; i <= k; ++i) wk -= y[i] * y[k - i];
      w[k] = wk;
    }
  }
}

// reverse along one direction.  In: z[1..K], y[0..K], w[0..K], yb[1..K] (output adjoints).
// In/out: yb[0], wb0 (shared order-0 accumulators).  Out: zb[1..K].
template <int MAXK, typename T>
PK_HD void tanh_dir_bwd(int K, const T (&z)[MAXK + 1], const T (&y)[MAXK + 1], const T (&w)[MAXK + 1],
                        T (&yb)[MAXK + 1], T (&zb)[MAXK + 1], T& wb0) {
  T wb[MAXK + 1];
#pragma unroll
  for (int k = 0; k <= MAXK; ++k) { wb[k] = T(0); zb[k] = T(0); }
#pragma unroll
  for (int k = MAXK; k >= 1; --k) {
    if (k <= K) {
      // w_k = -sum_i y_i y_{k-i}
#pragma unroll
      for (int i = 0; i <= k; ++i) yb[i] -= T(2) * wb[k] * y[k - i];
      // y_k = (1/k) sum_j j z_j w_{k-j}
      const T g = yb[k] * (T(1) / T(k));
#pragma unroll
      for (int j = 1; j <= k; ++j) {
        zb[j] += g * T(j) * w[k - j];
        wb[k - j] += g * T(j) * z[j];
      }
    }
  }
  wb0 += wb[0];
}

// Pre-activation jets from the OUTPUT jets (reverse pass without a pre-activation stash): given y[0..K], w[0] = 1 - y0^2
// and inv_w0 = 1 / w[0] (0 when the unit is saturated to w0 == 0: every coefficient of such a unit is exactly 0),
// fill z[1..K] and w[1..K].  Inverts y_k = (1/k) sum_j j z_j w_{k-j} for z_k; consistent with the forward pass because
// the forward computed y_k with the same w_0.
template <int MAXK, typename T>
PK_HD void tanh_dir_recover(int K, const T (&y)[MAXK + 1], T (&w)[MAXK + 1], T (&z)[MAXK + 1], T inv_w0) {
#pragma unroll
  for (int k = 1; k <= MAXK; ++k) {
    if (k <= K) {
      T acc = T(0);
#pragma unroll
      for (int j = 1; j < k; ++j) acc += T(j) * z[j] * w[k - j];
      z[k] = (y[k] - acc * (T(1) / T(k))) * inv_w0;
      T wk = T(0);
#pragma unroll
      for (int i = 0; i <= k; ++i) wk -= y[i] * y[k - i];
      w[k] = wk;
    } else {
      z[k] = T(0);
    }
  }
}

// order-0 closure: yb0 holds the value adjoint plus all direction contributions.
template <typename T>
PK_HD T tanh_finish_bwd(T y0, T w0, T yb0, T wb0) {
  return (yb0 - T(2) * wb0 * y0) * w0;
}

// ----------------------------------------------------------------------------- sin / cos
// zeta = omega * z is passed in already scaled.  s[0], c[0] given.
template <int MAXK, typename T>
PK_HD void sincos_dir_fwd(int K, const T (&zeta)[MAXK + 1], T (&s)[MAXK + 1], T (&c)[MAXK + 1]) {
#pragma unroll
  for (int k = 1; k <= MAXK; ++k) {
    if (k <= K) {
      T as = T(0), ac = T(0);
#pragma unroll
      for (int j = 1; j <= k; ++j) {
        as += T(j) * zeta[j] * c[k - j];
        ac -= T(j) * zeta[j] * s[k - j];
      }
      s[k] = as * (T(1) / T(k));
      c[k] = ac * (T(1) / T(k));
    }
  }
}

// In: zeta[1..K], s[0..K], c[0..K], sb[1..K], cb[1..K].  In/out: sb[0], cb[0].  Out: zetab[1..K].
template <int MAXK, typename T>
PK_HD void sincos_dir_bwd(int K, const T (&zeta)[MAXK + 1], const T (&s)[MAXK + 1], const T (&c)[MAXK + 1],
                          T (&sb)[MAXK + 1], T (&cb)[MAXK + 1], T (&zetab)[MAXK + 1]) {
#pragma unroll
  for (int k = 0; k <= MAXK; ++k) zetab[k] = T(0);
#pragma unroll
  for (int k = MAXK; k >= 1; --k) {
    if (k <= K) {
      const T gs = sb[k] * (T(1) / T(k));
      const T gc = cb[k] * (T(1) / T(k));
#pragma unroll
      for (int j = 1; j <= k; ++j) {
        zetab[j] += T(j) * (gs * c[k - j] - gc * s[k - j]);
        cb[k - j] += gs * T(j) * zeta[j];
        sb[k - j] -= gc * T(j) * zeta[j];
      }
    }
  }
}

// ----------------------------------------------------------------------------- s = v^(-1/2)
// v[0] (incl. eps), s[0] given; v[1..K] given; fill s[1..K].
template <int MAXK, typename T>
PK_HD void rsqrt_dir_fwd(int K, const T (&v)[MAXK + 1], T (&s)[MAXK + 1]) {
  const T inv_v0 = T(1) / v[0];
#pragma unroll
  for (int k = 1; k <= MAXK; ++k) {
    if (k <= K) {
      T acc = T(0);
#pragma unroll
      for (int j = 1; j <= k; ++j) acc += (T(-0.5) * T(j) - T(k - j)) * v[j] * s[k - j];
      s[k] = acc * inv_v0 * (T(1) / T(k));
    }
  }
}

// In: v[0..K], s[0..K], sb[1..K].  In/out: sb[0], vb0.  Out: vb[1..K].
template <int MAXK, typename T>
PK_HD void rsqrt_dir_bwd(int K, const T (&v)[MAXK + 1], const T (&s)[MAXK + 1], T (&sb)[MAXK + 1],
                         T (&vb)[MAXK + 1], T& vb0) {
  const T inv_v0 = T(1) / v[0];
#pragma unroll
  for (int k = 0; k <= MAXK; ++k) vb[k] = T(0);
#pragma unroll
  for (int k = MAXK; k >= 1; --k) {
    if (k <= K) {
      vb0 -= sb[k] * s[k] * inv_v0;
      const T g = sb[k] * inv_v0 * (T(1) / T(k));
#pragma unroll
      for (int j = 1; j <= k; ++j) {
        const T cf = T(-0.5) * T(j) - T(k - j);
        vb[j] += g * cf * s[k - j];
        sb[k - j] += g * cf * v[j];
      }
    }
  }
}

// ----------------------------------------------------------------------------- PDE epilogues
enum PdeKind : int {
  PDE_HEAT = 0,            // r = u_t - alpha * u_x (compat reference) or u_t - alpha u_xx (math)
  PDE_BURGERS = 1,         // r = u_t + u u_x - nu u_xx
  PDE_KDV = 2,             // r = u_t + 6 u u_x + u_xxx
  PDE_ALLEN_CAHN = 3,      // r = u_t - eps^2 u_xx - u + u^3
  PDE_CAHN_HILLIARD = 4,   // r = u_t + eps^2 u_xxxx - 1{|u|<=10}[(3u^2-1)u_xx + 6 u u_x^2]
  PDE_UT_ONLY = 5,         // multi-dim as-written: r = u_t
  PDE_UT_ALLEN_CAHN_ND = 6,// multi-dim Allen-Cahn as written: r = u_t - u + u^3
  PDE_CAHN_HILLIARD_2D = 7,// intended 2-D operator from x, y, x+y, x-y order-4 jets (+t)
  PDE_VALUE = 8,           // r = u              (BC/IC rows: error is u - target)
  PDE_DX = 9,              // r = u_x            (Heat periodic BC derivative match)
  PDE_WAVE = 10,           // r = u_tt - c^2 u_xx          (1-D; spec [x order 2, t order 2])   wave_equation.py:38-119
  PDE_CONVECTION = 11,     // r = u_t + v u_x              (1-D; spec [x order 1, t order 1])   convection_equation.py:43-78
  PDE_BLACK_SCHOLES = 12,  // r = V_t + sigma^2/2 S^2 V_SS + r S V_S - r V   (1-D; spec [S order 2, t order 1]; needs the
                           //     coordinate S = x of the point: `xs`)   black_scholes.py:44-94
  PDE_PENDULUM = 13,       // r = u_tt + (g/L) sin u       (spec [t order 2], one direction)   pendulum_equation.py:60-94
};

struct PdeDesc {
  int kind;
  int compat_math;   // Heat only: 0 = as written, 1 = intended operator
  float p0;          // alpha | nu | epsilon
  float p1;
};

// U holds the jet columns of the network output at one point (layout of JetSpec).
// For 1-D PDEs the spec is [x (order K), t (order 1)]: U = {u, a_x1..a_xK, a_t1}.
// Returns r and, if dU != nullptr, dr/dU[col] for every column.
template <typename T>
PK_HD T pde_residual(const PdeDesc& pd, const JetSpec& js, const T* U, T* dU, T xs = T(0)) {
  const int nc = js.ncols;
  if (dU) for (int c = 0; c < nc; ++c) dU[c] = T(0);
  const T u = U[0];
  const T p0 = T(pd.p0);
  switch (pd.kind) {
    case PDE_VALUE: if (dU) dU[0] = T(1); return u;
    case PDE_DX: if (dU) dU[js.col0[0]] = T(1); return U[js.col0[0]];
    case PDE_UT_ONLY: { const int ct = js.col0[js.ndirs - 1]; if (dU) dU[ct] = T(1); return U[ct]; }
    case PDE_UT_ALLEN_CAHN_ND: {
      const int ct = js.col0[js.ndirs - 1];
      if (dU) { dU[ct] = T(1); dU[0] = T(3) * u * u - T(1); }
      return U[ct] - u + u * u * u;
    }
    default: break;
  }
  const int ct = js.col0[js.ndirs - 1];     // time direction is last
  const T u_t = U[ct];
  const int cx = js.col0[0];
  if (pd.kind == PDE_CAHN_HILLIARD_2D) {
    // directions: 0:x 1:y 2:x+y 3:x-y (order 4 each), 4:t
    const T e2 = p0 * p0;
    const int c0 = js.col0[0], c1 = js.col0[1], c2 = js.col0[2], c3 = js.col0[3];
    const T ux = U[c0], uy = U[c1];
    const T uxx = T(2) * U[c0 + 1], uyy = T(2) * U[c1 + 1];
    const T bih = T(24) * ((T(2) / T(3)) * (U[c0 + 3] + U[c1 + 3]) + (U[c2 + 3] + U[c3 + 3]) / T(6));
    const T lap = uxx + uyy;
    const T inside = (u <= T(10) && u >= T(-10)) ? T(1) : T(0);
    const T f1 = T(3) * u * u - T(1);
    const T g2 = ux * ux + uy * uy;
    const T r = u_t + e2 * bih - inside * (f1 * lap + T(6) * u * g2);
    if (dU) {
      dU[ct] = T(1);
      dU[c0 + 3] = e2 * T(16); dU[c1 + 3] = e2 * T(16);
      dU[c2 + 3] = e2 * T(4);  dU[c3 + 3] = e2 * T(4);
      dU[c0 + 1] = -inside * f1 * T(2); dU[c1 + 1] = -inside * f1 * T(2);
      dU[c0] = -inside * T(12) * u * ux; dU[c1] = -inside * T(12) * u * uy;
      dU[0] = -inside * (T(6) * u * lap + T(6) * g2);
    }
    return r;
  }
  switch (pd.kind) {
    case PDE_HEAT: {
      if (pd.compat_math) { if (dU) { dU[ct] = T(1); dU[cx + 1] = -T(2) * p0; } return u_t - p0 * T(2) * U[cx + 1]; }
      if (dU) { dU[ct] = T(1); dU[cx] = -p0; }
      return u_t - p0 * U[cx];
    }
    case PDE_WAVE: {       // spec [x order 2, t order 2]: u_tt = 2 a_t2, u_xx = 2 a_x2
      const T c2 = p0 * p0;
      if (dU) { dU[ct + 1] = T(2); dU[cx + 1] = -T(2) * c2; }
      return T(2) * U[ct + 1] - c2 * T(2) * U[cx + 1];
    }
    case PDE_CONVECTION: {
      if (dU) { dU[ct] = T(1); dU[cx] = p0; }
      return u_t + p0 * U[cx];
    }
    case PDE_BLACK_SCHOLES: {   // p0 = sigma, p1 = risk-free rate; xs = S
      const T rf = T(pd.p1), hs2 = T(0.5) * p0 * p0 * xs * xs;
      if (dU) { dU[ct] = T(1); dU[cx + 1] = T(2) * hs2; dU[cx] = rf * xs; dU[0] = -rf; }
      return u_t + hs2 * T(2) * U[cx + 1] + rf * xs * U[cx] - rf * u;
    }
    case PDE_PENDULUM: {        // one direction (t, order 2): ct = 1, u_tt = 2 a_t2; p0 = g / L
      if (dU) { dU[ct + 1] = T(2); dU[0] = p0 * cos(u); }
      return T(2) * U[ct + 1] + p0 * sin(u);
    }
    case PDE_BURGERS: {
      const T ux = U[cx], uxx = T(2) * U[cx + 1];
      if (dU) { dU[ct] = T(1); dU[0] = ux; dU[cx] = u; dU[cx + 1] = -T(2) * p0; }
      return u_t + u * ux - p0 * uxx;
    }
    case PDE_KDV: {
      const T ux = U[cx], uxxx = T(6) * U[cx + 2];
      if (dU) { dU[ct] = T(1); dU[0] = T(6) * ux; dU[cx] = T(6) * u; dU[cx + 2] = T(6); }
      return u_t + T(6) * u * ux + uxxx;
    }
    case PDE_ALLEN_CAHN: {
      const T e2 = p0 * p0, uxx = T(2) * U[cx + 1];
      if (dU) { dU[ct] = T(1); dU[cx + 1] = -T(2) * e2; dU[0] = T(3) * u * u - T(1); }
      return u_t - e2 * uxx - u + u * u * u;
    }
    case PDE_CAHN_HILLIARD: {
      const T e2 = p0 * p0;
      const T ux = U[cx], uxx = T(2) * U[cx + 1], uxxxx = T(24) * U[cx + 3];
      const T inside = (u <= T(10) && u >= T(-10)) ? T(1) : T(0);
      const T f1 = T(3) * u * u - T(1);
      if (dU) {
        dU[ct] = T(1); dU[cx + 3] = T(24) * e2; dU[cx + 1] = -inside * f1 * T(2);
        dU[cx] = -inside * T(12) * u * ux; dU[0] = -inside * (T(6) * u * uxx + T(6) * ux * ux);
      }
      return u_t + e2 * uxxxx - inside * (f1 * uxx + T(6) * u * ux * ux);
    }
    default: return T(0);
  }
}

// per-sample loss rho(e) and rho'(e): mse | mae | huber(delta)   (pde_base.py:309-326)
enum LossKind : int { LOSS_MSE = 0, LOSS_MAE = 1, LOSS_HUBER = 2 };
template <typename T>
PK_HD T loss_rho(int kind, T delta, T e, T* drho) {
  if (kind == LOSS_MAE) { *drho = (e > T(0)) ? T(1) : ((e < T(0)) ? T(-1) : T(0)); return e < T(0) ? -e : e; }
  if (kind == LOSS_HUBER) {
    const T a = e < T(0) ? -e : e;
    if (a < delta) { *drho = e; return T(0.5) * e * e; }
    *drho = (e > T(0)) ? delta : -delta; return delta * (a - T(0.5) * delta);
  }
  *drho = T(2) * e; return e * e;
}

}  // namespace pinnk

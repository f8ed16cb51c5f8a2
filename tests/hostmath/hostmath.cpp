// Test shim: compiles the product's __host__ __device__ jet math (csrc/jet_math.cuh) with g++ so the
// recurrences and their hand-written adjoints can be checked in fp64 against the oracle on a CPU-only box.
// Built and used by tests/test_hostmath.py only; never part of the product.
#include "../../pinns_rl_pde_b200/csrc/jet_math.cuh"
using namespace pinnk;

template <int MAXK>
static void tanh_one(int K, const double* zin, const double* ybin, double* yout, double* zbout) {
  double z[MAXK + 1], y[MAXK + 1], w[MAXK + 1], yb[MAXK + 1], zb[MAXK + 1];
  for (int k = 0; k <= MAXK; ++k) { z[k] = k <= K ? zin[k] : 0; yb[k] = k <= K ? ybin[k] : 0; }
  y[0] = tanh(z[0]); w[0] = 1 - y[0] * y[0];
  tanh_dir_fwd<MAXK, double>(K, z, y, w);
  double wb0 = 0;
  tanh_dir_bwd<MAXK, double>(K, z, y, w, yb, zb, wb0);
  zb[0] = tanh_finish_bwd<double>(y[0], w[0], yb[0], wb0);
  for (int k = 0; k <= K; ++k) { yout[k] = y[k]; zbout[k] = zb[k]; }
}
template <int MAXK>
static void sin_one(int K, double omega, const double* zin, const double* sbin, const double* cbin, double* sout, double* cout,
                    double* zbout) {
  double z[MAXK + 1], s[MAXK + 1], c[MAXK + 1], sb[MAXK + 1], cb[MAXK + 1], zb[MAXK + 1];
  for (int k = 0; k <= MAXK; ++k) { z[k] = k <= K ? omega * zin[k] : 0; sb[k] = k <= K ? sbin[k] : 0; cb[k] = k <= K ? cbin[k] : 0; }
  s[0] = sin(z[0]); c[0] = cos(z[0]);
  sincos_dir_fwd<MAXK, double>(K, z, s, c);
  sincos_dir_bwd<MAXK, double>(K, z, s, c, sb, cb, zb);
  zb[0] = sb[0] * c[0] - cb[0] * s[0];
  for (int k = 0; k <= K; ++k) { sout[k] = s[k]; cout[k] = c[k]; zbout[k] = omega * zb[k]; }
}
template <int MAXK>
static void rsqrt_one(int K, const double* vin, const double* sbin, double* sout, double* vbout) {
  double v[MAXK + 1], s[MAXK + 1], sb[MAXK + 1], vb[MAXK + 1];
  for (int k = 0; k <= MAXK; ++k) { v[k] = k <= K ? vin[k] : 0; sb[k] = k <= K ? sbin[k] : 0; }
  s[0] = 1.0 / sqrt(v[0]);
  rsqrt_dir_fwd<MAXK, double>(K, v, s);
  double vb0 = 0;
  rsqrt_dir_bwd<MAXK, double>(K, v, s, sb, vb, vb0);
  vb0 += sb[0] * (-0.5) * s[0] / v[0];
  vb[0] = vb0;
  for (int k = 0; k <= K; ++k) { sout[k] = s[k]; vbout[k] = vb[k]; }
}

extern "C" {
void hm_tanh(int K, const double* z, const double* yb, double* y, double* zb) { tanh_one<4>(K, z, yb, y, zb); }
void hm_tanh_k(int maxk, int K, const double* z, const double* yb, double* y, double* zb) {
  if (maxk == 1) tanh_one<1>(K, z, yb, y, zb); else if (maxk == 2) tanh_one<2>(K, z, yb, y, zb);
  else if (maxk == 3) tanh_one<3>(K, z, yb, y, zb); else tanh_one<4>(K, z, yb, y, zb);
}
// z jets recovered from the output jets (reverse pass without a pre-activation stash): y in, z[1..K] and w out
void hm_tanh_recover(int K, const double* y, double* z, double* w) {
  double yy[5] = {0, 0, 0, 0, 0}, ww[5] = {0, 0, 0, 0, 0}, zz[5] = {0, 0, 0, 0, 0};
  for (int k = 0; k <= K; ++k) yy[k] = y[k];
  ww[0] = 1.0 - yy[0] * yy[0];
  tanh_dir_recover<4, double>(K, yy, ww, zz, ww[0] > 0 ? 1.0 / ww[0] : 0.0);
  for (int k = 0; k <= K; ++k) { z[k] = zz[k]; w[k] = ww[k]; }
}
void hm_sin(int K, double omega, const double* z, const double* sb, const double* cb, double* s, double* c, double* zb) {
  sin_one<4>(K, omega, z, sb, cb, s, c, zb);
}
void hm_rsqrt(int K, const double* v, const double* sb, double* s, double* vb) { rsqrt_one<4>(K, v, sb, s, vb); }
double hm_pde(int kind, int compat_math, double p0, int ndirs, const int* orders, int in_dim, const double* U, double* dU) {
  JetSpec js{}; js.ndirs = ndirs; js.in_dim = in_dim; int col = 1;
  for (int d = 0; d < ndirs; ++d) { js.order[d] = orders[d]; js.col0[d] = col; col += orders[d]; }
  js.ncols = col;
  PdeDesc pd{kind, compat_math, (float)p0, 0.f};
  return pde_residual<double>(pd, js, U, dU);
}
double hm_pde_x(int kind, double p0, double p1, int ndirs, const int* orders, int in_dim, const double* U, double* dU, double xs) {
  JetSpec js{}; js.ndirs = ndirs; js.in_dim = in_dim; int col = 1;
  for (int d = 0; d < ndirs; ++d) { js.order[d] = orders[d]; js.col0[d] = col; col += orders[d]; }
  js.ncols = col;
  PdeDesc pd{kind, 0, (float)p0, (float)p1};
  return pde_residual<double>(pd, js, U, dU, xs);
}
double hm_rho(int kind, double delta, double e, double* drho) { return loss_rho<double>(kind, delta, e, drho); }
}

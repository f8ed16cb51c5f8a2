// On-device samplers of the adaptive collocation strategies:
//   pde_base.py:895-935  RAR: probs = |r| + 1e-8, normalised; torch.multinomial(probs, n, replacement=True)
//                        (torch.multinomial refuses more than 2^24 categories)
//   pde_base.py:806-835  jittered grid: linspace x linspace meshgrid + randn jitter, clamped to the domain
// The draw is an inverse-CDF lookup in two levels: fp64 sums of 1024-candidate blocks, one inclusive scan of the block sums
// (a single CTA: 64 Ki blocks for 64 Mi candidates), then one warp per sample -- binary search of the block, then the
// position inside the block from a warp prefix sum.  Exact in distribution, any number of candidates, no host round trip.
#pragma once
#include <cstdint>

namespace pinnk {

constexpr int SAMPLE_BLOCK = 1024;          // candidates per first-level block

__global__ void sample_block_sums_kernel(const float* __restrict__ w, int64_t n, float eps, double* __restrict__ sums) {
  const int64_t b = blockIdx.x;
  const int64_t i0 = b * SAMPLE_BLOCK;
  double acc = 0.0;
  for (int k = threadIdx.x; k < SAMPLE_BLOCK; k += blockDim.x) {
    const int64_t i = i0 + k;
    if (i < n) acc += (double)(w[i] + eps);              // w[i] + eps in fp32 first: the reference's `residual_mag + 1e-8`
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double part[8];
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) s += part[k];      // fixed order
    sums[b] = s;
  }
}

// inclusive scan of sums[0..nb) in place, one CTA of 1024 threads; cdf[nb] = total
__global__ void sample_scan_kernel(double* __restrict__ sums, int64_t nb) {
  __shared__ double carry_s;
  __shared__ double warp_tot[32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x == 0) carry_s = 0.0;
  __syncthreads();
  for (int64_t base = 0; base < nb; base += blockDim.x) {
    const int64_t i = base + threadIdx.x;
    double v = (i < nb) ? sums[i] : 0.0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double u = __shfl_up_sync(0xffffffffu, v, o);
      if (lane >= o) v += u;
    }
    if (lane == 31) warp_tot[wid] = v;
    __syncthreads();
    if (wid == 0) {
      double t = (lane < (int)(blockDim.x >> 5)) ? warp_tot[lane] : 0.0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const double u = __shfl_up_sync(0xffffffffu, t, o);
        if (lane >= o) t += u;
      }
      warp_tot[lane] = t;                                  // inclusive totals of the warps
    }
    __syncthreads();
    const double prefix = carry_s + (wid > 0 ? warp_tot[wid - 1] : 0.0);
    if (i < nb) sums[i] = v + prefix;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = v + prefix;
    __syncthreads();
  }
  if (threadIdx.x == 0) sums[nb] = carry_s;
}

// one warp per sample: u[j] in [0, 1) (fp64: 2^24 fp32 values could not even address 64 Mi candidates)
__global__ void sample_draw_kernel(const float* __restrict__ w, int64_t n, float eps, const double* __restrict__ cdf,
                                   int64_t nb, const double* __restrict__ u, int64_t m, int64_t* __restrict__ idx) {
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const double total = cdf[nb];
  for (int64_t j = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; j < m; j += warps) {
    double target = u[j] * total;
    // first block whose inclusive sum exceeds the target
    int64_t lo = 0, hi = nb - 1;
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      if (cdf[mid] > target) hi = mid; else lo = mid + 1;
    }
    const int64_t b = lo;
    target -= (b > 0 ? cdf[b - 1] : 0.0);
    const int64_t i0 = b * SAMPLE_BLOCK + (int64_t)lane * 32;
    float wv[32];
    double mine = 0.0;
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const int64_t i = i0 + k;
      wv[k] = (i < n) ? (w[i] + eps) : 0.f;
      mine += (double)wv[k];
    }
    double incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    // the first lane whose inclusive sum exceeds the target owns the sample (the last non-empty lane when rounding pushed the
    // target past the block's sum)
    const unsigned hit = __ballot_sync(0xffffffffu, incl > target);
    const unsigned nonempty = __ballot_sync(0xffffffffu, mine > 0.0);
    int owner = hit ? (__ffs(hit) - 1) : (nonempty ? (31 - __clz(nonempty)) : 0);
    if (lane == owner) {
      double run = incl - mine;
      int64_t pick = -1;
#pragma unroll
      for (int k = 0; k < 32; ++k) {
        run += (double)wv[k];
        if (pick < 0 && run > target && wv[k] > 0.f) pick = i0 + k;
      }
      if (pick < 0) {                                       // rounding: last element of this lane with mass
#pragma unroll
        for (int k = 0; k < 32; ++k) if (wv[k] > 0.f) pick = i0 + k;
      }
      if (pick < 0) pick = (i0 < n) ? i0 : n - 1;
      idx[j] = pick;
    }
  }
}

// jittered n_side x n_side grid (pde_base.py:808-829): x[i * n_side + j] = clamp(xs[i] + nx * x_noise), t = clamp(ts[j] + nt *
// t_noise).  xs / ts are the linspace vectors and nx / nt the randn draws the caller made with torch (same values, same RNG
// consumption as the reference); the multiply and the add round separately, as torch's `x + randn * noise` does.
__global__ void jittered_grid_kernel(const float* __restrict__ xs, const float* __restrict__ ts, int n_side,
                                     const float* __restrict__ nx, const float* __restrict__ nt, float x_noise, float t_noise,
                                     float x_lo, float x_hi, float t_lo, float t_hi, float* __restrict__ x, float* __restrict__ t) {
  const int64_t n = (int64_t)n_side * n_side;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(p / n_side), j = (int)(p - (int64_t)i * n_side);
    const float xv = __fadd_rn(xs[i], __fmul_rn(nx[p], x_noise));
    const float tv = __fadd_rn(ts[j], __fmul_rn(nt[p], t_noise));
    x[p] = fminf(fmaxf(xv, x_lo), x_hi);
    t[p] = fminf(fmaxf(tv, t_lo), t_hi);
  }
}

}  // namespace pinnk

// Value-only forward of the RL sampler's Q-network over a candidate grid:
//   rl_agent.py:15-88   DQNNetwork = [Linear -> LayerNorm -> ReLU -> Dropout] x (num_layers - 1), Linear(hidden -> action_dim)
//   rl_agent.py:214-229 RLAgent.select_action: policy_net(points).view(1, -1) on <= 100^(d+1) grid points
// One kernel for the whole network: a block owns R candidate rows, thread f owns hidden unit f; the activations of the R
// rows stay in shared memory from the state vector to the Q-values; weight matrices are staged 32 K-columns at a time,
// transposed, through shared memory (coalesced global reads, conflict-free per-unit reads) and each staged weight is used
// for R rows.  The work is ~0.3 GFLOP per 100 x 100 grid: the point of the kernel is one FFMA-bound launch instead of ~12
// launch-bound ones with [N, hidden] round trips, not tensor cores.
#pragma once
#include <cstdint>

namespace pinnk {

constexpr int DQN_MAX_LAYERS = 8;
constexpr int DQN_ROWS = 16;

struct DqnNet {
  const float* W[DQN_MAX_LAYERS];       // [hidden, in] row-major (nn.Linear.weight)
  const float* b[DQN_MAX_LAYERS];       // [hidden] or null
  const float* gamma[DQN_MAX_LAYERS];   // LayerNorm weight / bias (null = no affine)
  const float* beta[DQN_MAX_LAYERS];
  const float* mask[DQN_MAX_LAYERS];    // [n, hidden] dropout mask already scaled by 1 / (1 - p), or null
  float eps[DQN_MAX_LAYERS];
  int n_hidden;                         // number of Linear-LayerNorm-ReLU groups
  int state_dim, hidden, out_dim;
  const float* W_out;                   // [out_dim, hidden]
  const float* b_out;                   // [out_dim] or null
};

__device__ __forceinline__ float dqn_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sum over the block's `hidden` threads of v[r], r < R; result broadcast to every thread.  red: [R][32] floats.
template <int R>
__device__ __forceinline__ void dqn_block_sum(float (&v)[R], float* red, int n_warps) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const float s = dqn_warp_sum(v[r]);
    if (lane == 0) red[r * 32 + wid] = s;
  }
  __syncthreads();
#pragma unroll
  for (int r = 0; r < R; ++r) {
    float s = (lane < n_warps) ? red[r * 32 + lane] : 0.f;
    v[r] = dqn_warp_sum(s);
  }
  __syncthreads();
}

// shared memory: h0 | h1 [R][ld] activations (ld = widest layer rounded up to 4, zero padded), wt [KC][pitch] the current
// K-chunk of a weight matrix TRANSPOSED (pitch odd: the transposing store and the per-unit read are both conflict-free),
// red [R][32] block-reduction scratch.
constexpr int DQN_KC = 32;
__host__ __device__ inline int dqn_ld(int hidden, int state_dim) { const int m = hidden > state_dim ? hidden : state_dim; return (m + 3) & ~3; }
__host__ __device__ inline int dqn_pitch(int hidden) { return hidden | 1; }
__host__ __device__ inline size_t dqn_smem_bytes(int rows, int hidden, int state_dim) {
  return sizeof(float) * ((size_t)2 * rows * dqn_ld(hidden, state_dim) + (size_t)DQN_KC * dqn_pitch(hidden) + (size_t)rows * 32);
}

template <int R>
__global__ void __launch_bounds__(1024) dqn_forward_kernel(DqnNet net, const float* __restrict__ states, int64_t n,
                                                           float* __restrict__ q_out) {
  extern __shared__ __align__(16) float sm[];
  const int H = net.hidden;
  const int ld = dqn_ld(H, net.state_dim), pitch = dqn_pitch(H);
  float* h0 = sm;
  float* h1 = h0 + R * ld;
  float* wt = h1 + R * ld;
  float* red = wt + DQN_KC * pitch;
  const int f = threadIdx.x;
  const bool live = f < H;
  const int n_warps = (blockDim.x + 31) >> 5;
  const float inv_h = 1.f / (float)H;
  for (int64_t row0 = (int64_t)blockIdx.x * R; row0 < n; row0 += (int64_t)gridDim.x * R) {
    const int rows = (int)((n - row0 < R) ? (n - row0) : R);
    for (int i = f; i < R * ld; i += blockDim.x) {          // state vectors, zero padded to ld
      const int r = i / ld, k = i - r * ld;
      h0[i] = (r < rows && k < net.state_dim) ? states[(row0 + r) * net.state_dim + k] : 0.f;
    }
    float* hin = h0;
    float* hout = h1;
    int in_dim = net.state_dim;
    for (int l = 0; l < net.n_hidden; ++l) {
      float acc[R];
      const float bias = (live && net.b[l]) ? net.b[l][f] : 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) acc[r] = bias;
      const float* __restrict__ W = net.W[l];
      for (int k0 = 0; k0 < in_dim; k0 += DQN_KC) {
        __syncthreads();                                      // previous chunk consumed / hin complete
        for (int i = f; i < H * DQN_KC; i += blockDim.x) {    // coalesced rows of W -> transposed chunk
          const int u = i / DQN_KC, kk = i - u * DQN_KC;
          wt[kk * pitch + u] = (k0 + kk < in_dim) ? W[(int64_t)u * in_dim + k0 + kk] : 0.f;
        }
        __syncthreads();
        const int kc = (in_dim - k0 < DQN_KC) ? ((in_dim - k0 + 3) & ~3) : DQN_KC;   // h is zero padded to a multiple of 4
        if (live) {
          for (int kk = 0; kk < kc; kk += 4) {
            const float w0 = wt[(kk + 0) * pitch + f], w1 = wt[(kk + 1) * pitch + f];
            const float w2 = wt[(kk + 2) * pitch + f], w3 = wt[(kk + 3) * pitch + f];
#pragma unroll
            for (int r = 0; r < R; ++r) {
              const float4 hv = *reinterpret_cast<const float4*>(&hin[r * ld + k0 + kk]);   // broadcast read
              acc[r] = fmaf(w0, hv.x, fmaf(w1, hv.y, fmaf(w2, hv.z, fmaf(w3, hv.w, acc[r]))));
            }
          }
        }
      }
      // LayerNorm over the hidden units of each row (biased variance, two passes like at::native_layer_norm)
      float s[R];
#pragma unroll
      for (int r = 0; r < R; ++r) s[r] = live ? acc[r] : 0.f;
      dqn_block_sum<R>(s, red, n_warps);
      float c[R];
#pragma unroll
      for (int r = 0; r < R; ++r) { c[r] = acc[r] - s[r] * inv_h; s[r] = live ? c[r] * c[r] : 0.f; }
      dqn_block_sum<R>(s, red, n_warps);
      if (f < ld) {
        const float g = (live && net.gamma[l]) ? net.gamma[l][f] : 1.f;
        const float be = (live && net.beta[l]) ? net.beta[l][f] : 0.f;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          float y = fmaf(c[r] * rsqrtf(s[r] * inv_h + net.eps[l]), g, be);
          y = fmaxf(y, 0.f);
          if (live && net.mask[l] && r < rows) y *= net.mask[l][(row0 + r) * H + f];
          hout[r * ld + f] = live ? y : 0.f;                 // padding columns stay zero
        }
      }
      float* tmp = hin; hin = hout; hout = tmp;
      in_dim = H;
    }
    __syncthreads();
    for (int i = f; i < rows * net.out_dim; i += blockDim.x) {
      const int r = i / net.out_dim, a = i - r * net.out_dim;
      const float* w = net.W_out + (int64_t)a * in_dim;
      float q = net.b_out ? net.b_out[a] : 0.f;
      for (int k = 0; k < in_dim; ++k) q = fmaf(w[k], hin[r * ld + k], q);
      q_out[(row0 + r) * net.out_dim + a] = q;
    }
    __syncthreads();
  }
}

// ---- wide Q-networks (hidden a multiple of 128 up to 1024; config.yaml:363 ships hidden_dim 512): the Linear layers between
// hidden groups run on the tcgen05 3xTF32 rows kernel (tc_linear_fwd), and this kernel does everything around them, one
// warp per candidate row with the row's hidden units in registers:
//   FIRST  z = W_0 s + b_0 from the state vector (state_dim <= 8), else z is read from Z (the GEMM's output);
//   LayerNorm (two passes, biased variance) -> ReLU -> dropout mask;
//   LAST   the output layer Linear(hidden, out_dim) as a warp reduction, else the activations go to Y for the next GEMM.
template <int NPER, bool FIRST, bool LAST>
__global__ void dqn_ln_relu_rows_kernel(const float* __restrict__ Z, const float* __restrict__ states, int state_dim,
                                        const float* __restrict__ W0, const float* __restrict__ b0,
                                        const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                        const float* __restrict__ mask, int64_t n, float* __restrict__ Y,
                                        const float* __restrict__ W_out, const float* __restrict__ b_out, int out_dim,
                                        float* __restrict__ q_out) {
  constexpr int H = NPER * 32;
  const int lane = threadIdx.x & 31;
  const int64_t warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const float inv_h = 1.f / (float)H;
  for (int64_t row = (((int64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; row < n; row += warps) {
    float z[NPER];
    if constexpr (FIRST) {
      float s[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = (k < state_dim) ? states[row * state_dim + k] : 0.f;
#pragma unroll
      for (int i = 0; i < NPER; ++i) {
        const int f = lane + 32 * i;
        float a = b0 ? b0[f] : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < state_dim) a = fmaf(W0[(int64_t)f * state_dim + k], s[k], a);
        z[i] = a;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NPER; ++i) z[i] = Z[row * H + lane + 32 * i];
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) sum += z[i];
    const float mean = dqn_warp_sum(sum) * inv_h;
    float vs = 0.f;
#pragma unroll
    for (int i = 0; i < NPER; ++i) { z[i] -= mean; vs += z[i] * z[i]; }
    const float rstd = rsqrtf(dqn_warp_sum(vs) * inv_h + eps);
#pragma unroll
    for (int i = 0; i < NPER; ++i) {
      const int f = lane + 32 * i;
      float y = fmaf(z[i] * rstd, gamma ? gamma[f] : 1.f, beta ? beta[f] : 0.f);
      y = fmaxf(y, 0.f);
      if (mask) y *= mask[row * H + f];
      z[i] = y;
      if constexpr (!LAST) Y[row * H + f] = y;
    }
    if constexpr (LAST) {
      for (int a = 0; a < out_dim; ++a) {
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < NPER; ++i) q = fmaf(W_out[(int64_t)a * H + lane + 32 * i], z[i], q);
        q = dqn_warp_sum(q);
        if (lane == 0) q_out[row * out_dim + a] = q + (b_out ? b_out[a] : 0.f);
      }
    }
  }
}

}  // namespace pinnk

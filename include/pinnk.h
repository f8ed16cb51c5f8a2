/* pinnk.h -- C ABI of the B200 PINN hot path (libpinnk.so).
 *
 * The reference (josegarciav/PINNs-RL-PDE, package `pinnrl`) is pure Python; it has no FFI.
 * Its boundary for this path is the Python plugin API
 *     PINNModel.forward(x)                          pinnrl/neural_networks/__init__.py:144-154
 *     PDEBase.compute_residual(model, x, t)         pinnrl/pdes/pde_base.py:577-588 (+ five overrides)
 *     PDEBase.compute_loss(model, x, t) -> dict     pinnrl/pdes/pde_base.py:1086-1235,
 *                                                   pinnrl/pdes/heat_equation.py:375-623
 *     loss["total"].backward()                      pinnrl/training/trainer.py:689
 *     RAR / RL residual scoring                     pinnrl/pdes/pde_base.py:895-935,1364-1377
 * Each entry point below names the reference call it stands in for.  The Python host
 * (pinns_rl_pde_b200/) binds these with ctypes; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every pointer marked "device" is a CUDA device pointer owned by the caller (torch);
 *     the library allocates nothing on the device: the caller supplies the workspace
 *   - all floating-point data is fp32; loss sums are fp64
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream)
 *   - every function returns 0 on success or a negative PINNK_E_* code and never throws;
 *     pinnk_last_error() returns a thread-local message for the last failure
 *   - a plan is immutable; calls using one workspace must be serialised on one stream
 */
#ifndef PINNK_H_
#define PINNK_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PINNK_ABI_VERSION 3

#define PINNK_E_INVALID   (-1)   /* bad argument / unsupported program            */
#define PINNK_E_CUDA      (-2)   /* CUDA runtime error (message has the string)   */
#define PINNK_E_WORKSPACE (-3)   /* workspace too small                           */

/* network program: what PINNModel.forward does, op by op (neural_networks/{feedforward,resnet,siren,fourier}.py) */
enum { PINNK_OP_LINEAR = 1, PINNK_OP_ACT = 2, PINNK_OP_LAYERNORM = 3,
       PINNK_OP_SKIP_SAVE = 4, PINNK_OP_SKIP_ADD = 5, PINNK_OP_SINCOS = 6 };
enum { PINNK_ACT_TANH = 1, PINNK_ACT_SIN = 2 };

typedef struct {
  int32_t kind;          /* PINNK_OP_*                                                        */
  int32_t in_dim;        /* feature width entering the op                                     */
  int32_t out_dim;       /* feature width leaving it (SINCOS: 2*in_dim)                       */
  int32_t act;           /* ACT: PINNK_ACT_*                                                  */
  float   scale;         /* ACT sin: omega_0 (siren.py:46); otherwise 1                       */
  float   eps;           /* LAYERNORM eps                                                     */
  int32_t w_index;       /* index into params[]: LINEAR weight [out,in] | LAYERNORM gamma; -1 */
  int32_t b_index;       /* index into params[]: bias / beta; -1 = none                       */
  int32_t w_transposed;  /* LINEAR: weight stored [in,out] (Fourier B, fourier.py:45)         */
  int32_t reserved;
  int64_t gw_offset;     /* float offset of dL/dW in the flat gradient buffer; -1 = frozen    */
  int64_t gb_offset;     /* same for bias / beta                                              */
} PinnkOp;

/* derivative jets carried through the network: column 0 = value, then order[d] normalised
 * Taylor coefficients along vec[d] (input space: x..., t).  Replaces the nested autograd.grad
 * chains of pde_base.py:661-732, allen_cahn.py:61-87, cahn_hilliard.py:61-136. */
typedef struct {
  int32_t ndirs;             /* 0..5                                     */
  int32_t order[5];          /* 1..4 each                                */
  float   vec[5][4];         /* direction in input space, in_dim entries */
} PinnkJetSpec;

/* per-row error functional e = f(U[row]) [- f(U[row+pair_offset])] [- target[i]] */
enum { PINNK_PDE_HEAT = 0, PINNK_PDE_BURGERS = 1, PINNK_PDE_KDV = 2, PINNK_PDE_ALLEN_CAHN = 3,
       PINNK_PDE_CAHN_HILLIARD = 4, PINNK_PDE_UT_ONLY = 5, PINNK_PDE_UT_ALLEN_CAHN_ND = 6,
       PINNK_PDE_CAHN_HILLIARD_2D = 7, PINNK_PDE_VALUE = 8, PINNK_PDE_DX = 9,
       PINNK_PDE_WAVE = 10,            /* u_tt - c^2 u_xx, jets [x:2, t:2]; p0 = c              (wave_equation.py:38-119)        */
       PINNK_PDE_CONVECTION = 11,      /* u_t + v u_x, jets [x:1, t:1]; p0 = v                   (convection_equation.py:43-78)   */
       PINNK_PDE_BLACK_SCHOLES = 12,   /* V_t + p0^2/2 S^2 V_SS + p1 S V_S - p1 V, jets [S:2, t:1] (black_scholes.py:44-94)        */
       PINNK_PDE_PENDULUM = 13 };      /* u_tt + p0 sin u, jets [t:2]; p0 = g / L                (pendulum_equation.py:60-94)     */
enum { PINNK_LOSS_MSE = 0, PINNK_LOSS_MAE = 1, PINNK_LOSS_HUBER = 2 };

typedef struct {
  int32_t kind;          /* PINNK_PDE_*                                                        */
  int32_t compat_math;   /* HEAT: 0 = operator as written (u_t - a u_x), 1 = u_t - a u_xx */
  float   p0;            /* alpha | nu | epsilon | c | v | sigma | g/L                         */
  float   p1;            /* Black-Scholes: risk-free rate                                      */
} PinnkPde;

typedef struct {
  PinnkPde pde;          /* functional applied to the rows                                     */
  int32_t component;     /* loss_sums[] slot this segment adds to                              */
  int32_t loss_kind;     /* PINNK_LOSS_* (pde_base.py:309-326)                                 */
  float   huber_delta;
  float   weight;        /* multiplies rho(e): 1/count for a mean                              */
  int64_t row_start;     /* first row (point index within this call)                           */
  int64_t row_count;
  int64_t pair_offset;   /* != 0: e = f(U[row]) - f(U[row+pair_offset]) (periodic match)       */
  const float* target;   /* device, nullable, [row_count]: subtracted from e                   */
  float* error_out;      /* device, nullable, [row_count]: receives e (compute_residual)       */
  const float* error_grad; /* device, nullable, [row_count]: upstream dL/de per row; when set the reverse
                            pass is seeded with error_grad * de/dU instead of weight * rho'(e)
                            (autograd through a compute_residual() result, CONTRIBUTING.md:241) */
} PinnkSegment;

typedef struct pinnk_plan_s* pinnk_plan_t;

/* Build a plan for (network program, jet spec).  `chunk_points` bounds how many points are
 * in flight at once (the activation stash lives in the workspace and is reused per chunk). */
int pinnk_plan_create(const PinnkOp* ops, int32_t n_ops, int32_t in_dim, const PinnkJetSpec* jets,
                      int64_t chunk_points, int32_t device, pinnk_plan_t* out);
void pinnk_plan_destroy(pinnk_plan_t plan);
int64_t pinnk_plan_workspace_bytes(pinnk_plan_t plan);   /* bytes the caller must provide */
int32_t pinnk_plan_ncols(pinnk_plan_t plan);             /* jet columns C                 */
int64_t pinnk_plan_grad_floats(pinnk_plan_t plan);       /* length of the flat gradient   */

/* PINNModel.forward / jets of the network output.  x: device [n, in_dim-1] (or [n, in_dim]
 * when t == NULL), t: device [n, 1]; out_jets: device [n, C].  Forward only. */
int pinnk_jets_forward(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                       int64_t n, float* out_jets, void* workspace, int64_t workspace_bytes, void* stream);

/* Vector-Jacobian product of pinnk_jets_forward w.r.t. the parameters: flat_grad += J^T adj_jets.
 * Stands in for autograd through compute_residual / model(x) when the caller builds its own loss
 * (CONTRIBUTING.md:241).  Recomputes the forward chunk by chunk; nothing is kept between calls. */
int pinnk_jets_vjp(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                   int64_t n, const float* adj_jets, float* flat_grad, void* workspace,
                   int64_t workspace_bytes, void* stream);

/* compute_loss (+ its backward) over one set of rows: for every segment adds
 * weight * sum_rows rho(e) to loss_sums[component] (device, fp64) and, when flat_grad != NULL,
 * adds d/dtheta of sum_seg grad_scale[component] * weight * sum rho(e) to flat_grad (device).
 * grad_scale: host array indexed by component (NULL = all ones).
 * Stands in for pde.compute_loss(model,x,t) followed by loss.backward()
 * (trainer.py:578,689; benchmarks/sampling.py:190-203). */
int pinnk_loss_step(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                    int64_t n, const PinnkSegment* segments, int32_t n_segments,
                    const float* grad_scale, double* loss_sums, float* flat_grad,
                    void* workspace, int64_t workspace_bytes, void* stream);

/* pinnk_loss_step with flags, for callers that split compute_residual and backward the way autograd does
 * (r = pde.compute_residual(model, x, t); loss(r).backward() -- CONTRIBUTING.md:241, notebook 05):
 *   PINNK_STEP_KEEP_STASH   a forward-only call (flat_grad == NULL) leaves the stash and the output jets in the workspace;
 *   PINNK_STEP_REUSE_STASH  the reverse-pass call finds them there and does not recompute the forward.  The caller vouches
 *                           that nothing else used the workspace and that rows and parameters are unchanged.
 * Both need n <= chunk_points. */
#define PINNK_STEP_KEEP_STASH 1
#define PINNK_STEP_REUSE_STASH 2
int pinnk_loss_step_flags(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                          int64_t n, const PinnkSegment* segments, int32_t n_segments,
                          const float* grad_scale, double* loss_sums, float* flat_grad,
                          void* workspace, int64_t workspace_bytes, void* stream, int32_t flags);

/* Forward-only residual scoring for the adaptive samplers (pde_base.py:909-921,1364-1377):
 * abs_residual_out (device, nullable, [n]) receives |r|; stats (device, fp64 [4]) accumulates
 * sum|r|, sum r^2, max|r|, count. */
int pinnk_score(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                int64_t n, const PinnkPde* pde, float* abs_residual_out, double* stats,
                void* workspace, int64_t workspace_bytes, void* stream);

/* Fused optimizer tail of the trainer step (trainer.py:690-694 clip_grad_norm_, :292-297 Adam with L2 weight decay):
 * params[i] (device, numels[i] floats, written in place) are laid out in flat_grad / exp_avg / exp_avg_sq in order.
 * max_norm <= 0 disables clipping.  scratch: device, 1 double.  step counts from 1. */
int pinnk_adam_step(float* const* params, const int64_t* numels, int32_t n_tensors, const float* flat_grad,
                    float* exp_avg, float* exp_avg_sq, double* scratch, int64_t step, float lr, float beta1,
                    float beta2, float eps, float weight_decay, float max_norm, void* stream);
/* The same step with the step count and the learning rate in DEVICE memory (dyn[0] = step >= 1, dyn[1] = lr, doubles):
 * nothing step-dependent is baked into the launch, so a captured CUDA graph of the trainer step can be replayed
 * (the caller advances dyn[0] on the stream before this call). */
int pinnk_adam_step_dev(float* const* params, const int64_t* numels, int32_t n_tensors, const float* flat_grad,
                        float* exp_avg, float* exp_avg_sq, double* scratch, const double* dyn, float beta1,
                        float beta2, float eps, float weight_decay, float max_norm, void* stream);

/* Q-network of the RL sampler, value-only forward over a candidate grid in ONE launch.  Replaces
 * DQNNetwork.forward (rl/rl_agent.py:15-88: [Linear -> LayerNorm -> ReLU -> Dropout] x n_hidden, Linear(hidden -> out_dim))
 * as called by RLAgent.select_action (rl_agent.py:214-229) from PDEBase.generate_collocation_points("adaptive")
 * (pdes/pde_base.py:961-1018).  All pointers are device pointers; bias / ln_weight / ln_bias / dropout_mask may be null.
 * dropout_mask: [n, hidden] floats, 0 or 1/(1-p), drawn by the caller (the reference's policy net is left in train mode,
 * so its dropout is live during scoring); null = eval mode.  states [n, in_dim of layer 0], q_out [n, out_dim]. */
typedef struct PinnkDqnLayer {
  const float* weight;       /* [out_dim, in_dim] row-major (nn.Linear.weight) */
  const float* bias;         /* [out_dim] */
  const float* ln_weight;    /* [out_dim] */
  const float* ln_bias;      /* [out_dim] */
  const float* dropout_mask; /* [n, out_dim] or null */
  float eps;                 /* LayerNorm eps */
  int32_t in_dim, out_dim;
} PinnkDqnLayer;
int pinnk_dqn_forward(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                      int32_t out_dim, const float* states, int64_t n, float* q_out, void* stream);
/* The same forward for hidden widths that are multiples of 128, up to 1024 (config.yaml:363 ships hidden_dim 512;
 * train.py:348-351): the hidden Linear layers run on the tcgen05 3xTF32 rows kernel, LayerNorm / ReLU / dropout mask / the
 * first and the output layer in one row kernel around them.  ws: device scratch of 2 * n * hidden floats. */
int pinnk_dqn_forward_wide(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                           int32_t out_dim, const float* states, int64_t n, float* q_out, float* ws, int64_t ws_floats,
                           void* stream);

const char* pinnk_last_error(void);
int32_t pinnk_abi_version(void);
/* number of kernel launches issued by this library since load (bench.py's gpu_launches) */
int64_t pinnk_launch_count(void);


/* Optional per-kernel-class device timing (CUDA event pairs on the launching stream), used by bench.py
 * for the roofline line.  Off by default; enabling it adds two event records per launch. */
void pinnk_prof_enable(int32_t on);
int32_t pinnk_prof_classes(void);
const char* pinnk_prof_class_name(int32_t cls);
int pinnk_prof_collect(double* ms_per_class, int64_t* launches_per_class, int32_t n_classes);

/* Debug / micro-benchmark: one hidden nn.Linear forward over stacked jet rows,
 * Z[M,N] = X[M,K] W[N,K]^T (+ bias on rows with row % jet_cols == 0).
 * mode 0 = exact-fp32 CUDA-core GEMM, 1 = tcgen05 3xTF32 GEMM (error if the shape is not covered). */
int pinnk_debug_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int32_t K,
                           int32_t N, int32_t jet_cols, int32_t mode, void* stream);
/* dX[M,K] = dZ[M,N] W[N,K] and dW[N,K] += dZ[M,N]^T X[M,K], db[N] += value-column rows of dZ: the two reverse GEMMs. */
int pinnk_debug_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int32_t K, int32_t N,
                             int32_t mode, void* stream);
int pinnk_debug_linear_wgrad(const float* dZ, const float* X, float* dW, float* db, int64_t M, int32_t K, int32_t N,
                             int32_t jet_cols, int32_t mode, void* stream);
/* The tcgen05 forward (trans = 0: Z[M,N] = X[M,K] W[N,K]^T + bias) or dgrad (trans = 1: Z[M,N] = X[M,K] W[K,N]) GEMM with a
 * scratch buffer for the K-split launch of K = 256 contractions (ring: device, ring_floats floats; null = two K-half passes). */
int pinnk_debug_linear_ks(const float* X, const float* W, const float* bias, float* Z, int64_t M, int32_t K, int32_t N,
                          int32_t jet_cols, int32_t trans, float* ring, int64_t ring_floats, void* stream);
/* Reverse pass of ONE hidden Linear(128, 128) fed by a tanh layer, on raw tensors (the per-layer step of
 * loss["total"].backward(), trainer.py:689): dZprev[M,128] = tanh'(.)^T (dZ[M,128] W[128,128]) with the tanh adjoint taken
 * from the previous layer's OUTPUT jets Yprev[M,128]; dW[128,128] += dZ^T Yprev; db[128] += value-column rows of dZ.
 * (k0, k1): jet orders of the (at most two) directions, 1 + k0 + k1 in {1, 2, 4}.  _pair: one launch of CTA pairs sharing the
 * tile stream (bwd_pair_kernel); _split: the dgrad + adjoint kernel and the wgrad kernel as two launches. */
int pinnk_debug_bwd_pair(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db,
                         int64_t M, int32_t k0, int32_t k1, void* stream);
int pinnk_debug_bwd_split(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db,
                          int64_t M, int32_t k0, int32_t k1, void* stream);

/* On-device samplers of the adaptive collocation strategies.
 * pinnk_sample_weighted: m draws WITH replacement from {0..n-1} with probability proportional to weights[i] + eps -- the RAR
 * draw `torch.multinomial(|r| + 1e-8, m, replacement=True)` of pde_base.py:924-931, fed directly by pinnk_score's abs_out, for
 * any n (torch.multinomial stops at 2^24 categories).  u: device [m] fp64 uniforms in [0, 1) (drawn by the caller: torch owns
 * the RNG); idx_out: device [m] int64; ws: device scratch of pinnk_sample_workspace_doubles(n) doubles.  Three launches, no
 * host synchronisation.
 * pinnk_jittered_grid: the jittered n_side x n_side collocation grid of _sample_uniform (pde_base.py:806-829) in one launch:
 * x[i*n_side+j] = clamp(xs[i] + noise_x * x_noise), t = clamp(ts[j] + noise_t * t_noise); xs / ts = the linspace vectors,
 * noise_* = the randn draws (made by the caller with the reference's RNG calls; the result is bit-identical to the torch ops). */
int64_t pinnk_sample_workspace_doubles(int64_t n);
int pinnk_sample_weighted(const float* weights, int64_t n, float eps, const double* u, int64_t m, int64_t* idx_out,
                          double* ws, int64_t ws_doubles, void* stream);
int pinnk_jittered_grid(const float* xs, const float* ts, int32_t n_side, const float* noise_x, const float* noise_t,
                        float x_noise, float t_noise, float x_lo, float x_hi, float t_lo, float t_hi, float* x_out,
                        float* t_out, void* stream);

/* Builder tool: cycles block 0 of the tcgen05 rows kernels spent waiting on each pipeline barrier, accumulated since the
 * last reset (which: 0 = forward rows kernels, 1 = reverse rows kernels, 2 = wgrad; out16: 16 counters, see csrc/tc_gemm.cuh).
 * Returns 1 and zeros unless the library was built with -DPINNK_STAGE_TIMERS. */
int pinnk_debug_stage_timers(int32_t which, uint64_t* out16, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* PINNK_H_ */

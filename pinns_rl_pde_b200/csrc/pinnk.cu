// pinnk.cu -- plan, orchestration and the C ABI (include/pinnk.h) of the B200 PINN hot path.
//
// One call = the reference's compute_loss(+backward) / compute_residual / scoring over a set of
// collocation rows (see include/pinnk.h for the reference call each entry point replaces).
// Rows are processed in chunks of plan->chunk points: forward jets through the network program
// (stashing pre-activations in the caller's workspace), PDE/loss epilogue, then the hand-written
// reverse pass that accumulates into the flat gradient buffer.  Nothing survives between chunks
// except the gradient and the loss sums, so memory is bounded by the chunk size.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <string.h>
#include <algorithm>
#include <atomic>
#include <type_traits>
#include <string>
#include <vector>

#include "../../include/pinnk.h"
#include "jet_math.cuh"
#include "kernels_ew.cuh"
#include "kernels_edge.cuh"
#include "kernels_dqn.cuh"
#include "kernels_sample.cuh"
#include "sgemm.cuh"
#include "tc_api.h"
#include "lnact_api.h"

using namespace pinnk;

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

// ---- optional per-kernel-class timing (bench.py roofline): CUDA event pairs on the launching stream
enum ProfClass { PC_FIRST_FWD = 0, PC_GEMM_FWD, PC_ACT_FWD, PC_LN_FWD, PC_LAST_FWD, PC_EPILOGUE, PC_LAST_BWD,
                 PC_ACT_BWD, PC_LN_BWD, PC_GEMM_DGRAD, PC_GEMM_WGRAD, PC_FIRST_BWD, PC_MISC, PC_FWD_LOSS, PC_BWD_PAIR, PC_COUNT };
static const char* kProfNames[PC_COUNT] = {"first_linear_fwd", "gemm_fwd", "act_fwd", "layernorm_fwd", "last_linear_fwd",
                                           "epilogue", "last_linear_bwd", "act_bwd", "layernorm_bwd", "gemm_dgrad",
                                           "gemm_wgrad", "first_linear_bwd", "misc", "fwd_loss_fused", "bwd_pair"};
struct ProfRec { int cls; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
struct ProfScope {
  bool on; ProfRec r; cudaStream_t st;
  ProfScope(int cls, cudaStream_t s) : on(g_prof_on), st(s) {
    if (on) { r.cls = cls; cudaEventCreate(&r.a); cudaEventCreate(&r.b); cudaEventRecord(r.a, st); }
  }
  ~ProfScope() { if (on) { cudaEventRecord(r.b, st); g_prof.push_back(r); } }
};

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define PK_CHECK_CUDA(expr)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return fail(PINNK_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));         \
  } while (0)

#define PK_LAUNCH_OK()                                                                       \
  do {                                                                                       \
    g_launches.fetch_add(1, std::memory_order_relaxed);                                      \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return fail(PINNK_E_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e));    \
  } while (0)

static inline unsigned blocks_for(int64_t work, int threads) { return (unsigned)((work + threads - 1) / threads); }

struct OpRt {
  PinnkOp op;
  int64_t out_off;     // float offset (per point -> multiplied by chunk) of the op's output jets; -1 = none
  int in_op;           // index of the op whose output feeds this one (-1 = network input)
  int skip_src;        // ACT with fused skip: op index whose output is added (-2 = none)
};

struct pinnk_plan_s {
  std::vector<OpRt> ops;
  JetSpec js;
  int maxk;
  int64_t chunk;
  int device;
  int max_width;
  int64_t stash_floats_per_point;   // sum of C*width over stashed outputs
  int64_t grad_floats;
  int64_t ws_bytes;
  int sm_count;
  bool fuse_first;   // ops[0] LINEAR (trainable, not transposed) + ops[1] plain ACT: one kernel each way, no Z stash
  bool fuse_last;    // last LINEAR preceded by a plain ACT: reverse of both in one kernel
  // workspace layout (float offsets)
  int64_t off_stash, off_U, off_Ub, off_adj[3];
};

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

// PINNK_SM_COUNT=n sizes every persistent grid for n SMs (experiments: kernels sharing the GPU on two streams)
static int sm_count_of(int device) {
  int smc = 148;
  cudaDeviceGetAttribute(&smc, cudaDevAttrMultiProcessorCount, device);   // stays 148 when no device is visible
  cudaGetLastError();
  if (smc <= 0) smc = 148;
  if (const char* e = getenv("PINNK_SM_COUNT")) { const int v = atoi(e); if (v > 0 && v < smc) smc = v; }
  return smc;
}

extern "C" const char* pinnk_last_error(void) { return g_err.c_str(); }
extern "C" int32_t pinnk_abi_version(void) { return PINNK_ABI_VERSION; }
extern "C" int64_t pinnk_launch_count(void) { return g_launches.load(); }
extern "C" void pinnk_prof_enable(int32_t on) { g_prof_on = on != 0; }
extern "C" int32_t pinnk_prof_classes(void) { return PC_COUNT; }
extern "C" const char* pinnk_prof_class_name(int32_t c) { return (c >= 0 && c < PC_COUNT) ? kProfNames[c] : ""; }
extern "C" int pinnk_prof_collect(double* ms, int64_t* launches, int32_t n) {
  for (int i = 0; i < n; ++i) { ms[i] = 0.0; launches[i] = 0; }
  for (auto& r : g_prof) {
    float t = 0.f;
    cudaEventSynchronize(r.b);
    cudaEventElapsedTime(&t, r.a, r.b);
    if (r.cls < n) { ms[r.cls] += t; launches[r.cls] += 1; }
    cudaEventDestroy(r.a); cudaEventDestroy(r.b);
  }
  g_prof.clear();
  return 0;
}

extern "C" int pinnk_plan_create(const PinnkOp* ops, int32_t n_ops, int32_t in_dim, const PinnkJetSpec* jets,
                                 int64_t chunk_points, int32_t device, pinnk_plan_t* out) {
  if (!ops || !jets || !out || n_ops < 2) return fail(PINNK_E_INVALID, "plan_create: null argument or fewer than 2 ops");
  if (in_dim < 1 || in_dim > 4) return fail(PINNK_E_INVALID, "plan_create: in_dim must be 1..4");
  if (jets->ndirs < 0 || jets->ndirs > kMaxDirs) return fail(PINNK_E_INVALID, "plan_create: ndirs must be 0..5");
  if (chunk_points < 1) return fail(PINNK_E_INVALID, "plan_create: chunk_points must be >= 1");
  auto* pl = new pinnk_plan_s();
  pl->device = device;
  pl->chunk = chunk_points;
  JetSpec& js = pl->js;
  memset(&js, 0, sizeof(js));
  js.ndirs = jets->ndirs;
  js.in_dim = in_dim;
  int col = 1, maxk = 0;
  for (int d = 0; d < jets->ndirs; ++d) {
    if (jets->order[d] < 1 || jets->order[d] > kMaxOrder) {
      delete pl;
      return fail(PINNK_E_INVALID, "plan_create: jet order must be 1..4");
    }
    js.order[d] = jets->order[d];
    js.col0[d] = col;
    col += jets->order[d];
    if (jets->order[d] > maxk) maxk = jets->order[d];
    for (int i = 0; i < 4; ++i) js.vec[d][i] = (i < in_dim) ? jets->vec[d][i] : 0.f;
  }
  js.ncols = col;
  pl->maxk = maxk;

  // validate the program and lay out the stash
  int cur_width = in_dim, prev = -1, skip_saved = -2, pending_skip = -2;
  int64_t off = 0, grad_end = 0;
  int max_width = 1;
  for (int i = 0; i < n_ops; ++i) {
    OpRt r;
    r.op = ops[i];
    r.out_off = -1;
    r.in_op = prev;
    r.skip_src = -2;
    const PinnkOp& o = ops[i];
    std::string where = "plan_create: op " + std::to_string(i) + ": ";
    auto bad = [&](const std::string& m) { delete pl; return fail(PINNK_E_INVALID, where + m); };
    switch (o.kind) {
      case PINNK_OP_LINEAR:
        if (o.in_dim != cur_width) return bad("LINEAR in_dim does not match the incoming width");
        if (o.w_index < 0) return bad("LINEAR needs a weight");
        if (i == 0) {
          if (o.in_dim != in_dim) return bad("first op must consume the network input");
        } else if (i == n_ops - 1) {
          if (o.out_dim != 1) return bad("last LINEAR must have out_dim == 1 (PINN output u)");
          if (o.w_transposed) return bad("transposed weight only supported on the first LINEAR");
        } else {
          if (o.w_transposed) return bad("transposed weight only supported on the first LINEAR");
          if ((o.in_dim % 4) || (o.out_dim % 4)) return bad("hidden LINEAR widths must be multiples of 4");
        }
        if (pending_skip != -2) return bad("SKIP_ADD must be followed by ACT");
        cur_width = o.out_dim;
        break;
      case PINNK_OP_ACT:
        if (o.act != PINNK_ACT_TANH && o.act != PINNK_ACT_SIN) return bad("unsupported activation (tanh | sin)");
        if (o.in_dim != cur_width || o.out_dim != cur_width) return bad("ACT width mismatch");
        if (prev < 0) return bad("ACT cannot be the first op");
        r.skip_src = pending_skip;
        pending_skip = -2;
        break;
      case PINNK_OP_LAYERNORM:
        if (o.in_dim != cur_width || o.out_dim != cur_width) return bad("LAYERNORM width mismatch");
        if (o.w_index < 0 || o.b_index < 0) return bad("LAYERNORM needs gamma and beta");
        if (cur_width > 512) return bad("LAYERNORM width > 512 not supported");
        if (prev < 0) return bad("LAYERNORM cannot be the first op");
        if (pending_skip != -2) return bad("SKIP_ADD must be followed by ACT");
        break;
      case PINNK_OP_SINCOS:
        if (o.in_dim != cur_width || o.out_dim != 2 * cur_width) return bad("SINCOS width mismatch");
        if (prev < 0) return bad("SINCOS cannot be the first op");
        if (prev != 0 || ops[0].gw_offset >= 0) return bad("SINCOS is only supported right after a frozen first LINEAR (Fourier features)");
        cur_width = o.out_dim;
        break;
      case PINNK_OP_SKIP_SAVE:
        if (prev < 0) return bad("SKIP_SAVE cannot be the first op");
        if (skip_saved != -2) return bad("nested skips are not supported");
        skip_saved = prev;
        break;
      case PINNK_OP_SKIP_ADD:
        if (skip_saved == -2) return bad("SKIP_ADD without SKIP_SAVE");
        if (pl->ops[skip_saved].op.out_dim != cur_width) return bad("SKIP_ADD width mismatch");
        pending_skip = skip_saved;
        skip_saved = -2;
        break;
      default:
        return bad("unknown op kind");
    }
    const bool produces = (o.kind == PINNK_OP_LINEAR || o.kind == PINNK_OP_ACT || o.kind == PINNK_OP_LAYERNORM ||
                           o.kind == PINNK_OP_SINCOS);
    if (produces) {
      if (i != n_ops - 1) {
        r.out_off = off;
        off += (int64_t)js.ncols * cur_width;
      }
      prev = i;
      if (cur_width > max_width) max_width = cur_width;
    }
    if (o.gw_offset >= 0) {
      const int64_t cnt = (o.kind == PINNK_OP_LINEAR) ? (int64_t)o.in_dim * o.out_dim : o.out_dim;
      if (o.gw_offset + cnt > grad_end) grad_end = o.gw_offset + cnt;
    }
    if (o.gb_offset >= 0 && o.gb_offset + o.out_dim > grad_end) grad_end = o.gb_offset + o.out_dim;
    pl->ops.push_back(r);
  }
  if (ops[n_ops - 1].kind != PINNK_OP_LINEAR || ops[0].kind != PINNK_OP_LINEAR) {
    delete pl;
    return fail(PINNK_E_INVALID, "plan_create: program must start and end with LINEAR");
  }
  if (pending_skip != -2 || skip_saved != -2) {
    delete pl;
    return fail(PINNK_E_INVALID, "plan_create: dangling skip connection");
  }
  pl->fuse_first = n_ops >= 3 && ops[1].kind == PINNK_OP_ACT && pl->ops[1].skip_src < 0 && !ops[0].w_transposed &&
                   (ops[0].gw_offset >= 0 || ops[0].gb_offset >= 0);
  pl->fuse_last = n_ops >= 3 && ops[n_ops - 2].kind == PINNK_OP_ACT && pl->ops[n_ops - 2].skip_src < 0 &&
                  pl->ops[n_ops - 2].in_op >= 1;
  pl->max_width = max_width;
  pl->stash_floats_per_point = off;
  pl->grad_floats = grad_end;
  // workspace: stash | U | Ub | adj0 | adj1 | adj2   (all sized for one chunk)
  int64_t o = 0;
  const int64_t C = js.ncols, n = chunk_points;
  pl->off_stash = o; o = align_up(o + off * n, 64);
  pl->off_U = o;     o = align_up(o + C * n, 64);
  pl->off_Ub = o;    o = align_up(o + C * n, 64);
  for (int k = 0; k < 3; ++k) { pl->off_adj[k] = o; o = align_up(o + C * n * max_width, 64); }
  pl->ws_bytes = o * (int64_t)sizeof(float);
  pl->sm_count = sm_count_of(device);
  *out = pl;
  return 0;
}

extern "C" void pinnk_plan_destroy(pinnk_plan_t plan) { delete plan; }
extern "C" int64_t pinnk_plan_workspace_bytes(pinnk_plan_t plan) { return plan ? plan->ws_bytes : 0; }
extern "C" int32_t pinnk_plan_ncols(pinnk_plan_t plan) { return plan ? plan->js.ncols : 0; }
extern "C" int64_t pinnk_plan_grad_floats(pinnk_plan_t plan) { return plan ? plan->grad_floats : 0; }

// ------------------------------------------------------------------------------------------------
template <typename F>
static int dispatch_maxk(int maxk, F&& f) {
  switch (maxk) {
    case 0: case 1: return f(std::integral_constant<int, 1>());
    case 2: return f(std::integral_constant<int, 2>());
    case 3: return f(std::integral_constant<int, 3>());
    default: return f(std::integral_constant<int, 4>());
  }
}


struct ChunkCtx {
  pinnk_plan_t pl;
  const float* const* params;
  float* ws;
  cudaStream_t st;
  int64_t n;          // points in this chunk
  const float* x;     // chunk-local input pointers
  const float* t;
  const TcLossFuse* loss_fuse = nullptr;   // loss_step: fold residual + loss + output-layer reverse into the last hidden layer
  mutable bool loss_done = false;          // set by forward_chunk when the fused kernel ran
  float* stash(int op) const { return ws + pl->off_stash + pl->ops[op].out_off * pl->chunk; }
  float* U() const { return ws + pl->off_U; }
  float* Ub() const { return ws + pl->off_Ub; }
  float* adj(int k) const { return ws + pl->off_adj[k]; }
};

// PINNK_DISABLE_TC=1 routes every GEMM to the exact-fp32 CUDA-core kernel (A/B checks of the tensor-core path)
static bool tc_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_DISABLE_TC"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

static int linear_fwd_any(const float* X, const float* W, const float* b, float* Z, int64_t M, int in_dim, int out_dim,
                          int ncols, int sm_count, cudaStream_t st, bool allow_tc, float* ring = nullptr, int64_t ring_floats = 0) {
  if (allow_tc) {
    int rc = tc_linear_fwd(X, W, b, Z, M, in_dim, out_dim, ncols, sm_count, st, ring, ring_floats);
    if (rc == 0) { g_launches.fetch_add(1); return 0; }
    if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_fwd launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  }
  dim3 grid(blocks_for(M, SG_BM), blocks_for(out_dim, SG_BN), 1);
  sgemm_kernel<true, true, EPI_BIAS_C0><<<grid, SG_THREADS, 0, st>>>(X, W, Z, M, out_dim, in_dim, in_dim, in_dim,
                                                                      out_dim, b, ncols, in_dim);
  PK_LAUNCH_OK();
  return 0;
}

// Scratch of the K-split launches of 256-wide layers (ring of partial products, tc_api.h).  Forward: an adjoint buffer (idle
// until the reverse pass; adj(0) may hold the output layer's partials, so adj(1)).  Reverse: the stash slot of the Linear being
// differentiated -- its output (the pre-activations) was consumed by the adjoint kernels that ran just before.
static inline int64_t adj_floats(const ChunkCtx& c) { return (int64_t)c.pl->js.ncols * c.pl->chunk * c.pl->max_width; }
static inline int64_t stash_floats(const ChunkCtx& c, int op) { return (int64_t)c.pl->js.ncols * c.pl->chunk * c.pl->ops[op].op.out_dim; }

static int gemm_fwd(const ChunkCtx& c, const float* X, const float* W, const float* b, float* Z, int in_dim, int out_dim) {
  ProfScope ps(PC_GEMM_FWD, c.st);
  return linear_fwd_any(X, W, b, Z, c.n * c.pl->js.ncols, in_dim, out_dim, c.pl->js.ncols, c.pl->sm_count, c.st, tc_enabled(),
                        c.adj(1), adj_floats(c));
}

static int gemm_dgrad(const ChunkCtx& c, const float* Zb, const float* W, float* Xb, int in_dim, int out_dim,
                      float* ring = nullptr, int64_t ring_floats = 0) {
  ProfScope ps(PC_GEMM_DGRAD, c.st);
  const int64_t M = c.n * c.pl->js.ncols;
  if (tc_enabled()) {
    int rc = tc_linear_dgrad(Zb, W, Xb, M, in_dim, out_dim, c.pl->sm_count, c.st, ring, ring_floats);
    if (rc == 0) { g_launches.fetch_add(1); return 0; }
    if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_dgrad launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  }
  dim3 grid(blocks_for(M, SG_BM), blocks_for(in_dim, SG_BN), 1);
  sgemm_kernel<true, false, EPI_STORE><<<grid, SG_THREADS, 0, c.st>>>(Zb, W, Xb, M, in_dim, out_dim, out_dim, in_dim,
                                                                       in_dim, nullptr, 1, out_dim);
  PK_LAUNCH_OK();
  return 0;
}

// PINNK_DETERMINISTIC=1: the hidden layers' weight gradients are reduced in a fixed order (per-CTA partial slabs in an
// idle adjoint buffer + wgrad_det_reduce_kernel) instead of with atomics in CTA arrival order: bit-identical dW / db of
// every 128-multiple hidden Linear from run to run (the arrival-order reduction moved them by 2e-6 .. 4e-6 relative).
// The few hundred per-thread atomics of the edge layers (input / output layer gradients) keep their arrival order.
static bool deterministic_enabled() {
  const char* e = getenv("PINNK_DETERMINISTIC");         // read per call: tests switch it on and off
  return e && e[0] == '1';
}

static int gemm_wgrad(const ChunkCtx& c, const float* Zb, const float* X, float* gW, float* gb, int in_dim, int out_dim,
                      float* scratch = nullptr) {
  ProfScope ps(PC_GEMM_WGRAD, c.st);
  const int64_t M = c.n * c.pl->js.ncols;   // contraction length
  if (tc_enabled() && gW) {
    const bool det = scratch != nullptr && deterministic_enabled();
    int rc = tc_linear_wgrad(Zb, X, gW, gb, M, in_dim, out_dim, c.pl->js.ncols, c.pl->sm_count, c.st, det ? scratch : nullptr,
                             det ? (int64_t)c.pl->js.ncols * c.pl->chunk * c.pl->max_width : 0);
    if (rc == 0) { g_launches.fetch_add(det ? 2 : 1); return 0; }
    if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_wgrad launch failed: ") + cudaGetErrorString(cudaGetLastError()));
  }
  if (gW) {
    const unsigned tiles = blocks_for(out_dim, SG_BM) * blocks_for(in_dim, SG_BN);
    int64_t splits = (2 * (int64_t)c.pl->sm_count + tiles - 1) / tiles;
    int64_t k_chunk = align_up((M + splits - 1) / splits, 64);
    if (k_chunk < 256) k_chunk = 256;
    splits = (M + k_chunk - 1) / k_chunk;
    dim3 grid(blocks_for(out_dim, SG_BM), blocks_for(in_dim, SG_BN), (unsigned)splits);
    sgemm_kernel<false, false, EPI_ATOMIC><<<grid, SG_THREADS, 0, c.st>>>(Zb, X, gW, out_dim, in_dim, M, out_dim, in_dim,
                                                                           in_dim, nullptr, 1, k_chunk);
    PK_LAUNCH_OK();
  }
  if (gb) {
    dim3 grid(blocks_for(out_dim, 128), (unsigned)std::min<int64_t>(c.n, 256));
    bias_grad_kernel<<<grid, 128, 0, c.st>>>(Zb, c.n, out_dim, c.pl->js.ncols, gb);
    PK_LAUNCH_OK();
  }
  return 0;
}

template <int MAXK>
static int ln_fwd(const ChunkCtx& c, const float* Z, float* Y, int width, const float* g, const float* b, float eps) {
  ProfScope ps(PC_LN_FWD, c.st);
  const int threads = 256;
  const unsigned blocks = blocks_for(c.n * 32, threads);
  const int nper = (width + 31) / 32;
#define LN_F(NP) layernorm_fwd_kernel<MAXK, NP><<<blocks, threads, 0, c.st>>>(Z, Y, c.n, width, c.pl->js, g, b, eps)
  // (NPER = 16: the 512-wide residual network the reference's YAML ships, config.yaml:15-19)
  if (nper <= 1) LN_F(1); else if (nper <= 2) LN_F(2); else if (nper <= 4) LN_F(4); else if (nper <= 8) LN_F(8); else LN_F(16);
#undef LN_F
  PK_LAUNCH_OK();
  return 0;
}

template <int MAXK>
static int ln_bwd(const ChunkCtx& c, const float* Z, const float* Gin, float* Gout, int width, const float* g, float eps,
                  float* dg, float* db) {
  ProfScope ps(PC_LN_BWD, c.st);
  const int threads = 128;
  const int wpb = threads / 32;
  int64_t blocks = (c.n + wpb - 1) / wpb;
  const int64_t cap = (int64_t)c.pl->sm_count * 16;
  if (blocks > cap) blocks = cap;
  const size_t sh = 2 * (size_t)width * sizeof(float);
  const int nper = (width + 31) / 32;
#define LN_B(NP) layernorm_bwd_kernel<MAXK, NP><<<(unsigned)blocks, threads, sh, c.st>>>(Z, Gin, Gout, c.n, width, c.pl->js, g, eps, dg, db)
  if (nper <= 1) LN_B(1); else if (nper <= 2) LN_B(2); else if (nper <= 4) LN_B(4); else if (nper <= 8) LN_B(8); else LN_B(16);
#undef LN_B
  PK_LAUNCH_OK();
  return 0;
}

// LayerNorm + activation as ONE kernel each way (lnact_fwd_kernel / lnact_bwd_kernel): index of the ACT op that consumes
// LAYERNORM op `ln` directly (possibly across a SKIP_ADD marker), or -1.  Forward and reverse pass both ask this function:
// for a fused pair the LayerNorm output is never written, and the reverse pass recomputes it from the LayerNorm input.
// PINNK_ENABLE_LNACT=1 switches it on (read per call so that tests can switch it).  Measured and NOT the default: with one warp
// per point and 8 features per lane the activation recurrences of a point run serially at 8 - 16 warps per SM, and the pair
// costs more than the two bandwidth-bound kernels it replaces (C3, 262 144 points: forward 22.2 ms vs 6.8 + 9.5, reverse
// 75.9 ms vs 16.8 + 10.7; profiles/r02r_c3_lnact.log).  The separate kernels already run at 4.3 - 5.7 TB/s.
//
// Round 2, last: the feature-per-thread kernels of lnact_feat.cu (a block per point, two features per thread, all
// cross-feature sums of a phase in one block reduction) ARE the default where they apply -- tanh, width 128 / 256 / 512, the
// two-direction jet layouts of the 1-D PDEs.  PINNK_ENABLE_LNACT=0: always the separate kernels; =1: the warp-per-point pair
// above for every LayerNorm -> activation pair (comparison runs); unset: lnact_feat.cu where it applies, separate kernels else.
static int lnact_mode() {
  const char* e = getenv("PINNK_ENABLE_LNACT");
  if (!e || !e[0]) return 2;
  return e[0] == '1' ? 1 : (e[0] == '0' ? 0 : 2);
}
static bool jet_orders(const JetSpec& js, int& k0, int& k1);
static bool lnact_feat_ok(const pinnk_plan_t pl, int width, int act) {
  int k0 = 0, k1 = 0;
  if (act != PINNK_ACT_TANH || !jet_orders(pl->js, k0, k1)) return false;
  const JetSpec& js = pl->js;
  if (js.ndirs < 1 || js.ncols != 1 + k0 + k1 || js.col0[0] != 1 || (js.ndirs > 1 && js.col0[1] != 1 + k0)) return false;
  return lnact_feat_supported(width, k0, k1);
}
static int lnact_partner(const pinnk_plan_t pl, int ln) {
  const int mode = lnact_mode();
  if (mode == 0) return -1;
  const int n_ops = (int)pl->ops.size();
  if (ln < 1 || ln >= n_ops - 1 || pl->ops[ln].op.kind != PINNK_OP_LAYERNORM) return -1;
  int j = ln + 1;
  if (j < n_ops - 1 && pl->ops[j].op.kind == PINNK_OP_SKIP_ADD) ++j;
  if (j >= n_ops - 1 || pl->ops[j].op.kind != PINNK_OP_ACT || pl->ops[j].in_op != ln) return -1;
  if (mode == 1) return pl->ops[ln].op.in_dim <= 256 ? j : -1;
  return lnact_feat_ok(pl, pl->ops[ln].op.in_dim, pl->ops[j].op.act) ? j : -1;
}
static int lnact_of_act(const pinnk_plan_t pl, int act) {
  if (act < 1 || pl->ops[act].op.kind != PINNK_OP_ACT) return -1;
  const int ln = pl->ops[act].in_op;
  return (ln >= 1 && lnact_partner(pl, ln) == act) ? ln : -1;
}

template <int MAXK>
static int lnact_fwd(const ChunkCtx& c, const float* Z, const float* S, float* Y, int width, const float* g, const float* b,
                     float eps, int act, float omega) {
  ProfScope ps(PC_LN_FWD, c.st);
  if (lnact_mode() == 2 && lnact_feat_ok(c.pl, width, act)) {
    int k0 = 0, k1 = 0;
    jet_orders(c.pl->js, k0, k1);
    const int rc = lnact_feat_fwd(Z, S, Y, c.n, width, k0, k1, g, b, eps, c.pl->sm_count, c.st);
    if (rc == 0) return 0;
    return fail(PINNK_E_CUDA, "lnact_feat_fwd launch failed");
  }
  const int threads = 256;
  const unsigned blocks = blocks_for(c.n * 32, threads);
  const int nper = (width + 31) / 32;
#define LNA_F(A, NP) lnact_fwd_kernel<A, MAXK, NP><<<blocks, threads, 0, c.st>>>(Z, S, Y, c.n, width, c.pl->js, g, b, eps, omega)
#define LNA_FA(NP) do { if (act == PINNK_ACT_TANH) LNA_F(1, NP); else LNA_F(2, NP); } while (0)
  if (nper <= 1) LNA_FA(1); else if (nper <= 2) LNA_FA(2); else if (nper <= 4) LNA_FA(4); else LNA_FA(8);
#undef LNA_FA
#undef LNA_F
  PK_LAUNCH_OK();
  return 0;
}

template <int MAXK>
static int lnact_bwd(const ChunkCtx& c, const float* Z, const float* S, const float* Gin, const float* Gin2, float* Gz,
                     float* Gout, int width, const float* g, const float* b, float eps, int act, float omega, float* dg,
                     float* db) {
  ProfScope ps(PC_LN_BWD, c.st);
  if (lnact_mode() == 2 && lnact_feat_ok(c.pl, width, act)) {
    int k0 = 0, k1 = 0;
    jet_orders(c.pl->js, k0, k1);
    const int rc = lnact_feat_bwd(Z, S, Gin, Gin2, Gz, Gout, c.n, width, k0, k1, g, b, eps, dg, db, c.pl->sm_count, c.st);
    if (rc == 0) return 0;
    return fail(PINNK_E_CUDA, "lnact_feat_bwd launch failed");
  }
  const int threads = 128;
  const int wpb = threads / 32;
  int64_t blocks = (c.n + wpb - 1) / wpb;
  const int64_t cap = (int64_t)c.pl->sm_count * 16;
  if (blocks > cap) blocks = cap;
  const size_t sh = 2 * (size_t)width * sizeof(float);
  const int nper = (width + 31) / 32;
#define LNA_B(A, NP) lnact_bwd_kernel<A, MAXK, NP><<<(unsigned)blocks, threads, sh, c.st>>>(Z, S, Gin, Gin2, Gz, Gout, c.n, width, c.pl->js, g, b, eps, omega, dg, db)
#define LNA_BA(NP) do { if (act == PINNK_ACT_TANH) LNA_B(1, NP); else LNA_B(2, NP); } while (0)
  if (nper <= 1) LNA_BA(1); else if (nper <= 2) LNA_BA(2); else if (nper <= 4) LNA_BA(4); else LNA_BA(8);
#undef LNA_BA
#undef LNA_B
  PK_LAUNCH_OK();
  return 0;
}

// (K0, K1) of the fused tcgen05 epilogues, or false when the jet spec has more than two directions
static bool jet_orders(const JetSpec& js, int& k0, int& k1) {
  if (js.ndirs > 2) return false;
  k0 = js.ndirs > 0 ? js.order[0] : 0;
  k1 = js.ndirs > 1 ? js.order[1] : 0;
  return true;
}

// compile-time jet layouts of the specialised edge-layer kernels (kernels_edge.cuh)
template <typename F>
static bool dispatch_edge_jets(int k0, int k1, F&& f) {
#define PK_EDGE_CASE(A, B) if (k0 == A && k1 == B) { f(std::integral_constant<int, A>(), std::integral_constant<int, B>()); return true; }
  PK_EDGE_CASE(0, 0) PK_EDGE_CASE(1, 0) PK_EDGE_CASE(2, 0) PK_EDGE_CASE(3, 0) PK_EDGE_CASE(4, 0) PK_EDGE_CASE(1, 1) PK_EDGE_CASE(2, 1) PK_EDGE_CASE(3, 1) PK_EDGE_CASE(4, 1) PK_EDGE_CASE(2, 2)
#undef PK_EDGE_CASE
  return false;
}
// PINNK_ACT_PPT=2: stand-alone activation kernels with two points per thread.  Measured and NOT the default: the extra
// registers cost more occupancy than the batched loads gain (C3 act_fwd 18.4 -> 20.9 ms, C4-math act_bwd 5.6 -> 7.5 ms;
// profiles/r01m_act_ppt_ab.log) -- these kernels are not latency-bound per thread.
static bool act_ppt2_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_ACT_PPT"); v = (e && e[0] == '2') ? 1 : 0; }
  return v == 1;
}

static bool edge_fast_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_DISABLE_EDGE_FAST"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1;
}

// PINNK_DISABLE_OUT_FUSE=1: output layer as its own pass over the last activation (A/B checks)
static bool out_fuse_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_DISABLE_OUT_FUSE"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1 && tc_enabled();
}

// PINNK_DISABLE_FIRST_FUSE=1: reverse of the input layer as its own kernel after a plain dgrad (A/B checks)
static bool first_fuse_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_DISABLE_FIRST_FUSE"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1 && tc_enabled();
}

// PINNK_ENABLE_PAIR=1: dgrad + tanh adjoint and wgrad of a hidden Linear(128, 128) in ONE launch of CTA pairs sharing the tile
// stream (bwd_pair_kernel).  Measured and NOT the default: each role then runs on 74 SMs, and a single SM cannot pull its
// tiles fast enough to make up for the halved parallelism (one 4 Mi-row layer: 2.41 ms paired vs 1.90 ms as two launches;
// the roles alone on 74 SMs take 1.28 / 1.49 ms -- profiles/r02c_pair_probe.log).  Kept for A/B runs.
static bool pair_enabled() {
  const char* e = getenv("PINNK_ENABLE_PAIR");            // read per call: tests switch it
  return (e && e[0] == '1') && tc_enabled();
}

static int first_trainable_op(const pinnk_plan_t pl) {
  const int n_ops = (int)pl->ops.size();
  for (int i = 0; i < n_ops; ++i)
    if (pl->ops[i].op.gw_offset >= 0 || pl->ops[i].op.gb_offset >= 0) return i;
  return n_ops;
}

// Is the pre-activation stash of hidden LINEAR op `lin` elided?  True when (a) the forward runs it as the fused
// Linear + tanh kernel without a partial-sum buffer and (b) the only reader of that stash in the reverse pass -- the
// adjoint of the tanh -- runs in a kernel that can work from the activation OUTPUT jets instead (EPI_ACTBWD_Y /
// last_act_bwd_fast_kernel<FROMY>).  Forward and reverse pass both ask this one function, so they cannot disagree.
// PINNK_KEEP_Z=1 switches the elision off (A/B checks).
static bool z_elided(const pinnk_plan_t pl, int lin) {
  static int keep = -1;
  if (keep < 0) { const char* e = getenv("PINNK_KEEP_Z"); keep = (e && e[0] == '1') ? 1 : 0; }
  if (keep || !tc_enabled() || !edge_fast_enabled()) return false;
  const int n_ops = (int)pl->ops.size();
  if (lin <= 0 || lin + 2 > n_ops - 1) return false;
  const PinnkOp& L = pl->ops[lin].op;
  const OpRt& A = pl->ops[lin + 1];
  const OpRt& N = pl->ops[lin + 2];
  if (L.kind != PINNK_OP_LINEAR || A.op.kind != PINNK_OP_ACT || A.skip_src >= 0 || A.op.act != PINNK_ACT_TANH) return false;
  if (A.in_op != lin || N.op.kind != PINNK_OP_LINEAR || N.in_op != lin + 1) return false;
  int k0 = 0, k1 = 0;
  if (!jet_orders(pl->js, k0, k1) || !tc_jets_supported(k0, k1)) return false;
  // producer: tc_linear_act_fwd in one pass (K = 128, or a narrower first hidden layer such as the Fourier network's K = 64)
  if (L.in_dim > 128 || L.in_dim < 16 || (L.in_dim % 4) != 0 || (L.out_dim % 128) != 0) return false;
  if (lin + 2 == n_ops - 1)                                                         // consumer: last_act_bwd_fast_kernel
    return pl->fuse_last && (N.op.in_dim % 128) == 0 && first_trainable_op(pl) <= n_ops - 2;
  return (N.op.out_dim == 128 || N.op.out_dim == 256) && (N.op.in_dim % 128) == 0; // consumer: tc_linear_dgrad_actbwd
}

// forward jets of one chunk; fills the stash and U[n, C]
// keep_stash = false (forward-only callers: jets_forward, scoring): fused Linear+activation kernels skip the
// pre-activation store, which nothing reads without a reverse pass
template <int MAXK>
static int forward_chunk(const ChunkCtx& c, bool keep_stash = true) {
  const pinnk_plan_t pl = c.pl;
  const JetSpec& js = pl->js;
  const int n_ops = (int)pl->ops.size();
  const int threads = 256;
  for (int i = 0; i < n_ops; ++i) {
    const OpRt& r = pl->ops[i];
    const PinnkOp& o = r.op;
    const float* in = (r.in_op >= 0) ? c.stash(r.in_op) : nullptr;
    switch (o.kind) {
      case PINNK_OP_LINEAR: {
        const float* W = c.params[o.w_index];
        const float* b = (o.b_index >= 0) ? c.params[o.b_index] : nullptr;
        if (i == 0 && pl->fuse_first && tc_enabled()) {
          ProfScope ps(PC_FIRST_FWD, c.st);
          const PinnkOp& a = pl->ops[1].op;
          const unsigned blocks = blocks_for(c.n * o.out_dim, threads);
          int ek0 = 0, ek1 = 0;
          if (edge_fast_enabled() && (o.out_dim % 128) == 0 && jet_orders(js, ek0, ek1)) {
            constexpr int PPT = 4;
            dim3 grid((unsigned)(o.out_dim / 128), (unsigned)std::min<int64_t>((c.n + PPT - 1) / PPT, 16 * (int64_t)pl->sm_count));
            const bool tanh_act = a.act == PINNK_ACT_TANH;
            const bool ok = dispatch_edge_jets(ek0, ek1, [&](auto ka, auto kb) {
              constexpr int KA = decltype(ka)::value, KB = decltype(kb)::value;
              if (tanh_act) first_act_fwd_fast_kernel<1, KA, KB, PPT><<<grid, 128, 0, c.st>>>(c.x, c.t, c.n, W, b, o.out_dim, js, c.stash(1), 1.f);
              else first_act_fwd_fast_kernel<2, KA, KB, PPT><<<grid, 128, 0, c.st>>>(c.x, c.t, c.n, W, b, o.out_dim, js, c.stash(1), a.scale);
            });
            if (ok) { PK_LAUNCH_OK(); ++i; break; }
          }
          if (a.act == PINNK_ACT_TANH)
            first_act_fwd_kernel<1, MAXK><<<blocks, threads, 0, c.st>>>(c.x, c.t, c.n, W, b, o.out_dim, js, nullptr, c.stash(1), 1.f);
          else
            first_act_fwd_kernel<2, MAXK><<<blocks, threads, 0, c.st>>>(c.x, c.t, c.n, W, b, o.out_dim, js, nullptr, c.stash(1), a.scale);
          PK_LAUNCH_OK();
          ++i;
        } else if (i == 0) {
          ProfScope ps(PC_FIRST_FWD, c.st);
          first_linear_fwd_kernel<<<blocks_for(c.n * o.out_dim, threads), threads, 0, c.st>>>(
              c.x, c.t, c.n, W, b, o.w_transposed, o.out_dim, js, c.stash(i));
          PK_LAUNCH_OK();
        } else if (i == n_ops - 1) {
          ProfScope ps(PC_LAST_FWD, c.st);
          const int64_t rows = c.n * js.ncols;
          if (edge_fast_enabled() && o.in_dim == 128) {
            constexpr int RPW = 4;
            const int64_t warps = (rows + RPW - 1) / RPW;
            const unsigned blocks = (unsigned)std::min<int64_t>((warps + 7) / 8, 8 * (int64_t)pl->sm_count);
            last_linear_fwd_w128_kernel<RPW><<<blocks, 256, 0, c.st>>>(in, rows, js.ncols, W, b, c.U());
          } else
          last_linear_fwd_kernel<<<blocks_for(rows * 32, threads), threads, 0, c.st>>>(in, rows, o.in_dim, js.ncols, W, b, c.U());
          PK_LAUNCH_OK();
        } else {
          // Linear + activation in one tcgen05 kernel when the next op is a plain activation
          int k0 = 0, k1 = 0;
          if (tc_enabled() && i + 1 < n_ops - 1 && pl->ops[i + 1].op.kind == PINNK_OP_ACT && pl->ops[i + 1].skip_src < 0 &&
              jet_orders(js, k0, k1)) {
            const PinnkOp& a = pl->ops[i + 1].op;
            const bool want_z = (keep_stash && !z_elided(pl, i)) || o.in_dim > 128;      // (K = 256: the partial-sum buffer)
            // last hidden layer: fold the output layer nn.Linear(width, 1) into the epilogue (partials in adj(0), which
            // is idle during the forward); forward-only callers then do not store the activation output either
            const bool fuse_out = out_fuse_enabled() && i + 2 == n_ops - 1 && pl->ops[i + 2].op.kind == PINNK_OP_LINEAR &&
                                  pl->ops[i + 2].in_op == i + 1 && !pl->ops[i + 2].op.w_transposed;
            const PinnkOp& lo = pl->ops[n_ops - 1].op;
            const float* w_out = fuse_out ? c.params[lo.w_index] : nullptr;
            if (fuse_out && c.loss_fuse != nullptr && keep_stash && !want_z) {
              // residual + loss + seeds + reverse of the output layer and of this tanh inside the epilogue: no activation
              // store, no U, no loss kernel, no last-layer reverse kernel
              ProfScope psl(PC_FWD_LOSS, c.st);
              int rcl = tc_linear_act_fwd(in, W, b, nullptr, nullptr, c.n * js.ncols, o.in_dim, o.out_dim, k0, k1,
                                          a.act == PINNK_ACT_TANH ? 1 : 2, a.scale, pl->sm_count, c.st, w_out, nullptr, c.loss_fuse);
              if (rcl == 0) { g_launches.fetch_add(1); c.loss_done = true; i = n_ops - 1; break; }
              if (rcl != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_act_fwd (loss fusion) launch failed: ") + cudaGetErrorString(cudaGetLastError()));
            }
            ProfScope ps(PC_GEMM_FWD, c.st);          // (opened here: the loss-fused launch above is its own class)
            float* y_out = (fuse_out && !keep_stash) ? nullptr : c.stash(i + 1);
            int rc = tc_linear_act_fwd(in, W, b, want_z ? c.stash(i) : nullptr, y_out, c.n * js.ncols, o.in_dim, o.out_dim, k0, k1,
                                       a.act == PINNK_ACT_TANH ? 1 : 2, a.scale, pl->sm_count, c.st, w_out, fuse_out ? c.adj(0) : nullptr,
                                       nullptr, c.adj(1), adj_floats(c));
            if (rc == 0 && fuse_out) {
              g_launches.fetch_add(1);
              ProfScope ps2(PC_LAST_FWD, c.st);
              const int64_t rows = c.n * js.ncols;
              output_combine_kernel<<<blocks_for(rows, threads), threads, 0, c.st>>>(
                  c.adj(0), 4 * (o.out_dim / 128), rows, js.ncols, lo.b_index >= 0 ? c.params[lo.b_index] : nullptr, c.U());
              PK_LAUNCH_OK();
              i = n_ops - 1;        // activation and output layer are done
              break;
            }
            if (rc == 0) { g_launches.fetch_add(1); ++i; break; }
            if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_act_fwd launch failed: ") + cudaGetErrorString(cudaGetLastError()));
          }
          int rc = gemm_fwd(c, in, W, b, c.stash(i), o.in_dim, o.out_dim);
          if (rc) return rc;
        }
        break;
      }
      case PINNK_OP_ACT: {
        ProfScope ps(PC_ACT_FWD, c.st);
        const float* S = (r.skip_src >= 0) ? c.stash(r.skip_src) : nullptr;
        const unsigned blocks = blocks_for(c.n * o.in_dim, threads);
        if (act_ppt2_enabled()) {
          const unsigned b2 = blocks_for(((c.n + 1) / 2) * o.in_dim, threads);
          if (o.act == PINNK_ACT_TANH)
            act_fwd_multi_kernel<1, MAXK, 2><<<b2, threads, 0, c.st>>>(in, S, c.stash(i), c.n, o.in_dim, js, 1.f);
          else
            act_fwd_multi_kernel<2, MAXK, 2><<<b2, threads, 0, c.st>>>(in, S, c.stash(i), c.n, o.in_dim, js, o.scale);
        } else if (o.act == PINNK_ACT_TANH)
          act_fwd_kernel<1, MAXK><<<blocks, threads, 0, c.st>>>(in, S, c.stash(i), c.n, o.in_dim, js, 1.f);
        else
          act_fwd_kernel<2, MAXK><<<blocks, threads, 0, c.st>>>(in, S, c.stash(i), c.n, o.in_dim, js, o.scale);
        PK_LAUNCH_OK();
        break;
      }
      case PINNK_OP_LAYERNORM: {
        const int ja = lnact_partner(pl, i);
        if (ja >= 0) {      // LayerNorm + the activation behind it in one sweep; the LayerNorm output is never written
          const OpRt& ar = pl->ops[ja];
          int rc = lnact_fwd<MAXK>(c, in, ar.skip_src >= 0 ? c.stash(ar.skip_src) : nullptr, c.stash(ja), o.in_dim,
                                   c.params[o.w_index], c.params[o.b_index], o.eps, ar.op.act, ar.op.scale);
          if (rc) return rc;
          i = ja;
          break;
        }
        int rc = ln_fwd<MAXK>(c, in, c.stash(i), o.in_dim, c.params[o.w_index], c.params[o.b_index], o.eps);
        if (rc) return rc;
        break;
      }
      case PINNK_OP_SINCOS: {
        ProfScope ps(PC_MISC, c.st);
        sincos_fwd_kernel<MAXK><<<blocks_for(c.n * o.in_dim, threads), threads, 0, c.st>>>(in, c.stash(i), c.n, o.in_dim, js);
        PK_LAUNCH_OK();
        break;
      }
      default: break;   // skip markers
    }
  }
  return 0;
}

// reverse pass of one chunk: Ub[n, C] -> flat_grad (accumulated)
template <int MAXK>
// last_done: the loss-fused forward already produced dL/dZ of the last hidden layer in adj(0) and the output layer's gradient
static int backward_chunk(const ChunkCtx& c, float* flat_grad, bool last_done = false) {
  const pinnk_plan_t pl = c.pl;
  const JetSpec& js = pl->js;
  const int n_ops = (int)pl->ops.size();
  const int threads = 256;
  int cur = 0;            // adjoint buffer holding dL/d(output of the op being processed)
  int held = -1;          // buffer holding the skip-branch adjoint
  int pend = -1;          // buffer whose adjoint the next fused LayerNorm + activation reverse kernel adds to its input
  auto other = [&](int a, int b) { for (int k = 0; k < 3; ++k) if (k != a && k != b) return k; return -1; };
  auto G = [&](int64_t off) -> float* { return off >= 0 ? flat_grad + off : nullptr; };
  // earliest op that still needs an input adjoint: stop propagating below the last trainable op
  int first_trainable = n_ops;
  for (int i = 0; i < n_ops; ++i)
    if (pl->ops[i].op.gw_offset >= 0 || pl->ops[i].op.gb_offset >= 0) { first_trainable = i; break; }
  for (int i = last_done ? n_ops - 3 : n_ops - 1; i >= first_trainable; --i) {
    const OpRt& r = pl->ops[i];
    const PinnkOp& o = r.op;
    const float* in = (r.in_op >= 0) ? c.stash(r.in_op) : nullptr;
    switch (o.kind) {
      case PINNK_OP_LINEAR: {
        const float* W = c.params[o.w_index];
        if (i == n_ops - 1 && pl->fuse_last && tc_enabled() && first_trainable <= n_ops - 2) {
          ProfScope ps(PC_LAST_BWD, c.st);
          const OpRt& pa = pl->ops[n_ops - 2];
          dim3 grid(blocks_for(o.in_dim, 128), (unsigned)std::min<int64_t>(c.n, 32 * (int64_t)pl->sm_count));
          int ek0 = 0, ek1 = 0;
          if (edge_fast_enabled() && (o.in_dim % 128) == 0 && jet_orders(js, ek0, ek1)) {
            constexpr int PPT = 2;
            dim3 g2((unsigned)(o.in_dim / 128), (unsigned)std::min<int64_t>((c.n + PPT - 1) / PPT, 16 * (int64_t)pl->sm_count));
            const bool tanh_act = pa.op.act == PINNK_ACT_TANH;
            const bool from_y = z_elided(pl, pa.in_op);
            const bool ok = dispatch_edge_jets(ek0, ek1, [&](auto ka, auto kb) {
              constexpr int KA = decltype(ka)::value, KB = decltype(kb)::value;
              if (from_y) last_act_bwd_fast_kernel<1, KA, KB, PPT, true><<<g2, 128, 0, c.st>>>(c.stash(n_ops - 2), c.Ub(), c.n, o.in_dim, W, c.adj(cur), G(o.gw_offset), G(o.gb_offset), 1.f);
              else if (tanh_act) last_act_bwd_fast_kernel<1, KA, KB, PPT, false><<<g2, 128, 0, c.st>>>(c.stash(pa.in_op), c.Ub(), c.n, o.in_dim, W, c.adj(cur), G(o.gw_offset), G(o.gb_offset), 1.f);
              else last_act_bwd_fast_kernel<2, KA, KB, PPT, false><<<g2, 128, 0, c.st>>>(c.stash(pa.in_op), c.Ub(), c.n, o.in_dim, W, c.adj(cur), G(o.gw_offset), G(o.gb_offset), pa.op.scale);
            });
            if (ok) { PK_LAUNCH_OK(); --i; break; }
            if (from_y) return fail(PINNK_E_INVALID, "backward: pre-activation stash elided but no edge kernel for this jet layout");
          }
          if (pa.op.act == PINNK_ACT_TANH)
            last_act_bwd_kernel<1, MAXK><<<grid, 128, 0, c.st>>>(c.stash(pa.in_op), c.Ub(), c.n, o.in_dim, js, W, c.adj(cur),
                                                                  G(o.gw_offset), G(o.gb_offset), 1.f);
          else
            last_act_bwd_kernel<2, MAXK><<<grid, 128, 0, c.st>>>(c.stash(pa.in_op), c.Ub(), c.n, o.in_dim, js, W, c.adj(cur),
                                                                  G(o.gw_offset), G(o.gb_offset), pa.op.scale);
          PK_LAUNCH_OK();
          --i;     // the activation's adjoint is done
        } else if (i == n_ops - 1) {
          ProfScope ps(PC_LAST_BWD, c.st);
          const int64_t rows = c.n * js.ncols;
          dim3 grid(blocks_for(o.in_dim, 128), (unsigned)std::min<int64_t>(rows, 32 * (int64_t)pl->sm_count));
          last_linear_bwd_kernel<<<grid, 128, 0, c.st>>>(in, c.Ub(), rows, o.in_dim, js.ncols, W, c.adj(cur),
                                                         G(o.gw_offset), G(o.gb_offset));
          PK_LAUNCH_OK();
        } else if (i == 0) {
          ProfScope ps(PC_FIRST_BWD, c.st);
          dim3 grid(blocks_for(o.out_dim, 128), (unsigned)std::min<int64_t>(c.n, 32 * (int64_t)pl->sm_count));
          first_linear_bwd_kernel<<<grid, 128, 0, c.st>>>(c.x, c.t, c.n, o.out_dim, js, c.adj(cur), G(o.gw_offset), G(o.gb_offset));
          PK_LAUNCH_OK();
        } else {
          // Linear(128, 128) fed by a plain tanh whose output jets are stashed: dgrad + tanh adjoint and the weight gradient in
          // ONE launch of CTA pairs sharing the tile stream (bwd_pair_kernel): dZ and Y come out of HBM once, not twice
          if (i > first_trainable && pair_enabled() && !deterministic_enabled() && o.in_dim == 128 && o.out_dim == 128 &&
              o.gw_offset >= 0 && r.in_op == i - 1 && pl->ops[i - 1].op.kind == PINNK_OP_ACT && pl->ops[i - 1].skip_src < 0 &&
              pl->ops[i - 1].in_op >= 0 && !(pl->fuse_first && i - 1 == 1 && first_trainable == 0) &&
              z_elided(pl, pl->ops[i - 1].in_op)) {
            int pk0 = 0, pk1 = 0;
            if (jet_orders(js, pk0, pk1)) {
              ProfScope ps(PC_BWD_PAIR, c.st);
              const int nxt = other(cur, held);
              int prc = tc_bwd_pair(c.adj(cur), W, c.stash(i - 1), c.adj(nxt), G(o.gw_offset), G(o.gb_offset), c.n * js.ncols,
                                    o.in_dim, o.out_dim, pk0, pk1, pl->sm_count, c.st);
              if (prc == 0) { g_launches.fetch_add(1); cur = nxt; --i; break; }
              if (prc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_bwd_pair launch failed: ") + cudaGetErrorString(cudaGetLastError()));
            }
          }
          // (the adjoint buffer the dgrad below will write is idle during the wgrad: scratch of the deterministic reduction)
          int rc = gemm_wgrad(c, c.adj(cur), in, G(o.gw_offset), G(o.gb_offset), o.in_dim, o.out_dim, c.adj(other(cur, held)));
          if (rc) return rc;
          if (i > first_trainable) {
            const int nxt = other(cur, held);
            // dgrad + adjoint of the activation feeding this Linear in one tcgen05 kernel
            int k0 = 0, k1 = 0;
            const OpRt& pa = pl->ops[i - 1];
            // (the first activation's adjoint is fused with the first Linear instead, and has no stash to read)
            const bool first_pair = pl->fuse_first && i - 1 == 1 && first_trainable == 0;
            // (an activation fused with the LayerNorm in front of it has no stashed input: its adjoint runs in lnact_bwd)
            if (tc_enabled() && r.in_op == i - 1 && pa.op.kind == PINNK_OP_ACT && pa.skip_src < 0 && pa.in_op >= 0 &&
                !first_pair && lnact_of_act(pl, i - 1) < 0 && jet_orders(js, k0, k1)) {
              ProfScope ps(PC_GEMM_DGRAD, c.st);
              const bool from_y = z_elided(pl, pa.in_op);      // the forward did not stash this pre-activation
              rc = tc_linear_dgrad_actbwd(c.adj(cur), W, from_y ? c.stash(i - 1) : c.stash(pa.in_op), c.adj(nxt), c.n * js.ncols,
                                          o.in_dim, o.out_dim, k0, k1, pa.op.act == PINNK_ACT_TANH ? 1 : 2, pa.op.scale,
                                          pl->sm_count, c.st, from_y ? 1 : 0, r.out_off >= 0 ? c.stash(i) : nullptr,
                                          r.out_off >= 0 ? stash_floats(c, i) : 0);
              if (rc == 0) { g_launches.fetch_add(1); cur = nxt; --i; break; }
              if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_dgrad_actbwd launch failed: ") + cudaGetErrorString(cudaGetLastError()));
              if (from_y) return fail(PINNK_E_INVALID, "backward: pre-activation stash elided but the fused adjoint kernel refused the shape");
            }
            if (first_pair && first_fuse_enabled() && jet_orders(js, k0, k1)) {
              // dgrad of the first hidden layer + the complete reverse of the input layer in one tcgen05 kernel:
              // dL/dY0 is never written, dW0 / db0 accumulate in the epilogue's registers
              ProfScope ps(PC_FIRST_BWD, c.st);
              const PinnkOp& l0 = pl->ops[0].op;
              rc = tc_linear_dgrad_firstbwd(c.adj(cur), W, c.n * js.ncols, o.in_dim, o.out_dim, k0, k1,
                                            pa.op.act == PINNK_ACT_TANH ? 1 : 2, pa.op.scale, c.x, c.t, js.in_dim,
                                            js.ndirs > 0 ? js.vec[0] : nullptr, js.ndirs > 1 ? js.vec[1] : nullptr,
                                            c.params[l0.w_index], l0.b_index >= 0 ? c.params[l0.b_index] : nullptr,
                                            G(l0.gw_offset), G(l0.gb_offset), pl->sm_count, c.st);
              if (rc == 0) { g_launches.fetch_add(1); i = 0; break; }      // ops 1 (activation) and 0 (input layer) are done
              if (rc != TC_UNSUPPORTED) return fail(PINNK_E_CUDA, std::string("tc_linear_dgrad_firstbwd launch failed: ") + cudaGetErrorString(cudaGetLastError()));
            }
            rc = gemm_dgrad(c, c.adj(cur), W, c.adj(nxt), o.in_dim, o.out_dim, r.out_off >= 0 ? c.stash(i) : nullptr,
                            r.out_off >= 0 ? stash_floats(c, i) : 0);
            if (rc) return rc;
            cur = nxt;
          }
        }
        break;
      }
      case PINNK_OP_ACT: {
        if (i == 1 && pl->fuse_first && tc_enabled() && first_trainable == 0) {
          // adjoint of the first activation + weight gradient of the first Linear, pre-activation recomputed from (x, t)
          ProfScope ps(PC_FIRST_BWD, c.st);
          const PinnkOp& l0 = pl->ops[0].op;
          const float* W0 = c.params[l0.w_index];
          const float* b0 = (l0.b_index >= 0) ? c.params[l0.b_index] : nullptr;
          dim3 grid(blocks_for(l0.out_dim, 128), (unsigned)std::min<int64_t>(c.n, 32 * (int64_t)pl->sm_count));
          int ek0 = 0, ek1 = 0;
          if (edge_fast_enabled() && (l0.out_dim % 128) == 0 && jet_orders(js, ek0, ek1)) {
            constexpr int PPT = 2;
            dim3 g2((unsigned)(l0.out_dim / 128), (unsigned)std::min<int64_t>((c.n + PPT - 1) / PPT, 16 * (int64_t)pl->sm_count));
            const bool tanh_act = o.act == PINNK_ACT_TANH;
            const bool ok = dispatch_edge_jets(ek0, ek1, [&](auto ka, auto kb) {
              constexpr int KA = decltype(ka)::value, KB = decltype(kb)::value;
              if (tanh_act) first_act_bwd_fast_kernel<1, KA, KB, PPT><<<g2, 128, 0, c.st>>>(c.x, c.t, c.n, W0, b0, l0.out_dim, js, c.adj(cur), G(l0.gw_offset), G(l0.gb_offset), 1.f);
              else first_act_bwd_fast_kernel<2, KA, KB, PPT><<<g2, 128, 0, c.st>>>(c.x, c.t, c.n, W0, b0, l0.out_dim, js, c.adj(cur), G(l0.gw_offset), G(l0.gb_offset), o.scale);
            });
            if (ok) { PK_LAUNCH_OK(); --i; break; }
          }
          if (o.act == PINNK_ACT_TANH)
            first_act_bwd_kernel<1, MAXK><<<grid, 128, 0, c.st>>>(c.x, c.t, c.n, W0, b0, l0.out_dim, js, c.adj(cur),
                                                                   G(l0.gw_offset), G(l0.gb_offset), 1.f);
          else
            first_act_bwd_kernel<2, MAXK><<<grid, 128, 0, c.st>>>(c.x, c.t, c.n, W0, b0, l0.out_dim, js, c.adj(cur),
                                                                   G(l0.gw_offset), G(l0.gb_offset), o.scale);
          PK_LAUNCH_OK();
          --i;     // the first Linear is done
          break;
        }
        const float* S = (r.skip_src >= 0) ? c.stash(r.skip_src) : nullptr;
        const unsigned blocks = blocks_for(c.n * o.in_dim, threads);
        const int ln = lnact_of_act(pl, i);
        if (ln >= 0) {
          // activation adjoint + LayerNorm reverse in one sweep: dL/dY in adj(cur) (+ adj(pend), the skip adjoint deferred at
          // SKIP_SAVE) -> dL/d(LayerNorm input) in a free buffer; with a skip the pre-activation adjoint is written in place
          // and becomes the held skip adjoint (no copy)
          const OpRt& lr = pl->ops[ln];
          const int nxt = other(cur, held >= 0 ? held : pend);
          int rc = lnact_bwd<MAXK>(c, c.stash(lr.in_op), S, c.adj(cur), pend >= 0 ? c.adj(pend) : nullptr,
                                   r.skip_src >= 0 ? c.adj(cur) : nullptr, c.adj(nxt), o.in_dim, c.params[lr.op.w_index],
                                   c.params[lr.op.b_index], lr.op.eps, o.act, o.scale, G(lr.op.gw_offset), G(lr.op.gb_offset));
          if (rc) return rc;
          pend = -1;
          if (r.skip_src >= 0) held = cur;
          cur = nxt;
          i = ln;          // the LayerNorm (and a SKIP_ADD marker in between) are done
          break;
        }
        if (z_elided(pl, r.in_op)) return fail(PINNK_E_INVALID, "backward: generic activation adjoint reached for an elided pre-activation stash");
        ProfScope ps(PC_ACT_BWD, c.st);
        // G2: the skip adjoint deferred at SKIP_SAVE is added while the output adjoint is read (no separate add kernel)
        const float* G2 = pend >= 0 ? c.adj(pend) : nullptr;
        if (act_ppt2_enabled() && G2 == nullptr) {
          const unsigned b2 = blocks_for(((c.n + 1) / 2) * o.in_dim, threads);
          if (o.act == PINNK_ACT_TANH)
            act_bwd_multi_kernel<1, MAXK, 2><<<b2, threads, 0, c.st>>>(in, S, c.adj(cur), c.n, o.in_dim, js, 1.f);
          else
            act_bwd_multi_kernel<2, MAXK, 2><<<b2, threads, 0, c.st>>>(in, S, c.adj(cur), c.n, o.in_dim, js, o.scale);
        } else if (o.act == PINNK_ACT_TANH)
          act_bwd_kernel<1, MAXK><<<blocks, threads, 0, c.st>>>(in, S, c.adj(cur), c.n, o.in_dim, js, 1.f, G2);
        else
          act_bwd_kernel<2, MAXK><<<blocks, threads, 0, c.st>>>(in, S, c.adj(cur), c.n, o.in_dim, js, o.scale, G2);
        PK_LAUNCH_OK();
        pend = -1;
        // dL/dz feeds both the LayerNorm branch and the skip: this buffer stays as the held skip adjoint (the LayerNorm
        // reverse below reads it and writes another buffer, so no copy is needed)
        if (r.skip_src >= 0) held = cur;
        break;
      }
      case PINNK_OP_LAYERNORM: {
        const int nxt = other(cur, held);
        int rc = ln_bwd<MAXK>(c, in, c.adj(cur), c.adj(nxt), o.in_dim, c.params[o.w_index], o.eps, G(o.gw_offset), G(o.gb_offset));
        if (rc) return rc;
        cur = nxt;
        break;
      }
      case PINNK_OP_SKIP_SAVE: {
        if (held < 0) return fail(PINNK_E_INVALID, "backward: SKIP_SAVE without a held adjoint");
        if (i - 1 >= first_trainable && pl->ops[i - 1].op.kind == PINNK_OP_ACT && !(i - 1 == 1 && pl->fuse_first && tc_enabled()) &&
            !z_elided(pl, pl->ops[i - 1].in_op)) {
          // the op below is an activation whose reverse kernel (act_bwd_kernel / lnact_bwd_kernel) adds the two adjoints
          // while it reads them: no separate add pass
          pend = held;
          held = -1;
          break;
        }
        const int64_t cnt = c.n * js.ncols * pl->ops[r.in_op].op.out_dim;
        add_inplace_kernel<<<blocks_for(cnt, threads), threads, 0, c.st>>>(c.adj(cur), c.adj(held), cnt);
        PK_LAUNCH_OK();
        held = -1;
        break;
      }
      default: break;   // SKIP_ADD marker, SINCOS (never reached: frozen)
    }
  }
  return 0;
}

static int check_common(pinnk_plan_t plan, const float* const* params, const float* x, int64_t n, void* ws, int64_t ws_bytes) {
  if (!plan) return fail(PINNK_E_INVALID, "null plan");
  if (!params || !x) return fail(PINNK_E_INVALID, "null params / x");
  if (n < 0) return fail(PINNK_E_INVALID, "negative row count");
  if (!ws || ws_bytes < plan->ws_bytes) return fail(PINNK_E_WORKSPACE, "workspace smaller than pinnk_plan_workspace_bytes()");
  return 0;
}

static ChunkCtx make_ctx(pinnk_plan_t plan, const float* const* params, const float* x, const float* t, int64_t p0,
                         int64_t cn, void* ws, void* stream) {
  ChunkCtx c;
  c.pl = plan; c.params = params; c.ws = (float*)ws; c.st = (cudaStream_t)stream; c.n = cn;
  const int d = plan->js.in_dim;
  c.x = t ? x + p0 * (d - 1) : x + p0 * d;
  c.t = t ? t + p0 : nullptr;
  return c;
}

static inline const float* params_host_ptr(const float* const* params, int idx) { return params[idx]; }

// PINNK_DISABLE_LOSS_FUSE=1: residual / loss / output-layer reverse as separate kernels (A/B checks)
static bool loss_fuse_enabled() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("PINNK_DISABLE_LOSS_FUSE"); v = (e && e[0] == '1') ? 0 : 1; }
  return v == 1 && tc_enabled() && out_fuse_enabled();
}

// Index of the segment whose loss can be folded into the last hidden layer's forward epilogue for the chunk
// [p0, p0 + cn), or -1: exactly one segment touches the chunk, it covers it completely and is a plain error functional
// (no target, no paired rows, no per-row error output / upstream gradient); the network ends in
// Linear(128, 128) + tanh + Linear(128, 1) whose pre-activation stash is elided (so the reverse pass would work from
// the output jets anyway).
static int loss_fusable_segment(pinnk_plan_t pl, const PinnkSegment* segs, int n_segs, int64_t p0, int64_t cn) {
  if (!loss_fuse_enabled()) return -1;
  const int n_ops = (int)pl->ops.size();
  if (n_ops < 4) return -1;
  const PinnkOp& L = pl->ops[n_ops - 3].op;
  if (!z_elided(pl, n_ops - 3) || L.in_dim != 128 || L.out_dim != 128) return -1;
  if (pl->ops[n_ops - 1].op.w_transposed || first_trainable_op(pl) > n_ops - 3) return -1;
  int found = -1;
  for (int s = 0; s < n_segs; ++s) {
    const PinnkSegment& g = segs[s];
    const int64_t lo = std::max(g.row_start, p0), hi = std::min(g.row_start + g.row_count, p0 + cn);
    if (lo >= hi) continue;
    if (found >= 0) return -1;
    if (g.row_start > p0 || g.row_start + g.row_count < p0 + cn) return -1;
    if (g.target || g.error_out || g.error_grad || g.pair_offset != 0) return -1;
    if (g.pde.kind == PDE_BLACK_SCHOLES) return -1;            // needs the point coordinate, which the GEMM epilogue does not have
    found = s;
  }
  return found;
}

extern "C" int pinnk_jets_forward(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                                  int64_t n, float* out_jets, void* ws, int64_t ws_bytes, void* stream) {
  int rc = check_common(plan, params, x, n, ws, ws_bytes);
  if (rc) return rc;
  if (!out_jets) return fail(PINNK_E_INVALID, "null out_jets");
  const int C = plan->js.ncols;
  for (int64_t p0 = 0; p0 < n; p0 += plan->chunk) {
    const int64_t cn = std::min(plan->chunk, n - p0);
    ChunkCtx c = make_ctx(plan, params, x, t, p0, cn, ws, stream);
    rc = dispatch_maxk(plan->maxk, [&](auto mk) { return forward_chunk<decltype(mk)::value>(c, false); });
    if (rc) return rc;
    PK_CHECK_CUDA(cudaMemcpyAsync(out_jets + p0 * C, c.U(), sizeof(float) * cn * C, cudaMemcpyDeviceToDevice, c.st));
  }
  return 0;
}

extern "C" int pinnk_jets_vjp(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                              int64_t n, const float* adj_jets, float* flat_grad, void* ws, int64_t ws_bytes,
                              void* stream) {
  int rc = check_common(plan, params, x, n, ws, ws_bytes);
  if (rc) return rc;
  if (!adj_jets || !flat_grad) return fail(PINNK_E_INVALID, "null adj_jets / flat_grad");
  const int C = plan->js.ncols;
  for (int64_t p0 = 0; p0 < n; p0 += plan->chunk) {
    const int64_t cn = std::min(plan->chunk, n - p0);
    ChunkCtx c = make_ctx(plan, params, x, t, p0, cn, ws, stream);
    rc = dispatch_maxk(plan->maxk, [&](auto mk) { return forward_chunk<decltype(mk)::value>(c); });
    if (rc) return rc;
    PK_CHECK_CUDA(cudaMemcpyAsync(c.Ub(), adj_jets + p0 * C, sizeof(float) * cn * C, cudaMemcpyDeviceToDevice, c.st));
    rc = dispatch_maxk(plan->maxk, [&](auto mk) { return backward_chunk<decltype(mk)::value>(c, flat_grad); });
    if (rc) return rc;
  }
  return 0;
}

extern "C" int pinnk_loss_step(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                               int64_t n, const PinnkSegment* segs, int32_t n_segs, const float* grad_scale,
                               double* loss_sums, float* flat_grad, void* ws, int64_t ws_bytes, void* stream) {
  return pinnk_loss_step_flags(plan, params, x, t, n, segs, n_segs, grad_scale, loss_sums, flat_grad, ws, ws_bytes, stream, 0);
}

// flags: PINNK_STEP_KEEP_STASH  -- a forward-only call (flat_grad == NULL) leaves in the workspace everything a reverse
//                                  pass needs (n <= chunk_points: one chunk);
//        PINNK_STEP_REUSE_STASH -- the workspace still holds the stash and the output jets of a KEEP_STASH call on the
//                                  same rows and parameters (the caller vouches for that): the forward is not recomputed.
// Together they turn "r = compute_residual(...); loss(r).backward()" from two forward passes + one reverse into one + one.
extern "C" int pinnk_loss_step_flags(pinnk_plan_t plan, const float* const* params, const float* x, const float* t,
                                     int64_t n, const PinnkSegment* segs, int32_t n_segs, const float* grad_scale,
                                     double* loss_sums, float* flat_grad, void* ws, int64_t ws_bytes, void* stream,
                                     int32_t flags) {
  int rc = check_common(plan, params, x, n, ws, ws_bytes);
  if (rc) return rc;
  if (!segs || n_segs < 1) return fail(PINNK_E_INVALID, "loss_step: no segments");
  const bool keep_stash = (flags & PINNK_STEP_KEEP_STASH) != 0, reuse_stash = (flags & PINNK_STEP_REUSE_STASH) != 0;
  if ((keep_stash || reuse_stash) && n > plan->chunk)
    return fail(PINNK_E_INVALID, "loss_step: KEEP_STASH / REUSE_STASH need n <= chunk_points (one chunk)");
  if (reuse_stash && !flat_grad) return fail(PINNK_E_INVALID, "loss_step: REUSE_STASH without a gradient buffer");
  const int C = plan->js.ncols;
  for (int s = 0; s < n_segs; ++s) {
    const PinnkSegment& g = segs[s];
    if (g.row_start < 0 || g.row_count < 0 || g.row_start + g.row_count > n)
      return fail(PINNK_E_INVALID, "loss_step: segment rows out of range");
    if (g.pair_offset != 0) {
      if (n > plan->chunk) return fail(PINNK_E_INVALID, "loss_step: paired segments need n <= chunk_points");
      if (g.pair_offset < g.row_count || g.row_start + g.pair_offset + g.row_count > n)
        return fail(PINNK_E_INVALID, "loss_step: pair_offset must address a disjoint row range inside the call");
    }
  }
  const int threads = 256;
  for (int64_t p0 = 0; p0 < n; p0 += plan->chunk) {
    const int64_t cn = std::min(plan->chunk, n - p0);
    ChunkCtx c = make_ctx(plan, params, x, t, p0, cn, ws, stream);
    const bool keep = flat_grad != nullptr || keep_stash;
    // loss fusion: one plain segment covers the whole chunk and the network ends in Linear(128,128) + tanh + Linear(.,1)
    TcLossFuse lf;
    const int fs = (flat_grad != nullptr && !reuse_stash) ? loss_fusable_segment(plan, segs, n_segs, p0, cn) : -1;
    if (fs >= 0) {
      const PinnkSegment& g = segs[fs];
      const PinnkOp& lo = plan->ops.back().op;
      memset(&lf, 0, sizeof(lf));
      lf.pde.kind = g.pde.kind; lf.pde.compat_math = g.pde.compat_math; lf.pde.p0 = g.pde.p0; lf.pde.p1 = g.pde.p1;
      lf.js = plan->js;
      lf.loss_kind = g.loss_kind; lf.huber_delta = g.huber_delta; lf.weight = g.weight;
      lf.grad_weight = g.weight * (grad_scale ? grad_scale[g.component] : 1.f);
      lf.loss_slot = loss_sums ? loss_sums + g.component : nullptr;
      lf.dz_out = c.adj(0);
      lf.gw_out = lo.gw_offset >= 0 ? flat_grad + lo.gw_offset : nullptr;
      lf.gb_out = lo.gb_offset >= 0 ? flat_grad + lo.gb_offset : nullptr;
      lf.b_out = lo.b_index >= 0 ? params_host_ptr(params, lo.b_index) : nullptr;
      c.loss_fuse = &lf;
    }
    if (!reuse_stash) {
      rc = dispatch_maxk(plan->maxk, [&](auto mk) { return forward_chunk<decltype(mk)::value>(c, keep); });
      if (rc) return rc;
    }
    if (c.loss_done) {
      rc = dispatch_maxk(plan->maxk, [&](auto mk) { return backward_chunk<decltype(mk)::value>(c, flat_grad, true); });
      if (rc) return rc;
      continue;
    }
    if (flat_grad) PK_CHECK_CUDA(cudaMemsetAsync(c.Ub(), 0, sizeof(float) * cn * C, c.st));
    for (int s = 0; s < n_segs; ++s) {
      const PinnkSegment& g = segs[s];
      const int64_t lo = std::max(g.row_start, p0), hi = std::min(g.row_start + g.row_count, p0 + cn);
      if (lo >= hi) continue;
      SegmentDev sd;
      sd.pde.kind = g.pde.kind; sd.pde.compat_math = g.pde.compat_math; sd.pde.p0 = g.pde.p0; sd.pde.p1 = g.pde.p1;
      sd.loss_kind = g.loss_kind; sd.huber_delta = g.huber_delta; sd.weight = g.weight;
      sd.grad_weight = flat_grad ? g.weight * (grad_scale ? grad_scale[g.component] : 1.f) : 0.f;
      sd.row_start = g.row_start; sd.row_count = g.row_count; sd.pair_offset = g.pair_offset;
      sd.target = g.target; sd.error_out = g.error_out; sd.error_grad = flat_grad ? g.error_grad : nullptr;
      sd.loss_slot = loss_sums ? loss_sums + g.component : nullptr;
      ProfScope ps(PC_EPILOGUE, c.st);
      epilogue_kernel<<<blocks_for(hi - lo, threads), threads, 0, c.st>>>(c.U(), flat_grad ? c.Ub() : nullptr, plan->js, sd, p0, lo, hi,
                                                                            c.x, c.t ? plan->js.in_dim - 1 : plan->js.in_dim);
      PK_LAUNCH_OK();
    }
    if (flat_grad) {
      rc = dispatch_maxk(plan->maxk, [&](auto mk) { return backward_chunk<decltype(mk)::value>(c, flat_grad); });
      if (rc) return rc;
    }
  }
  return 0;
}

extern "C" int pinnk_score(pinnk_plan_t plan, const float* const* params, const float* x, const float* t, int64_t n,
                           const PinnkPde* pde, float* abs_out, double* stats, void* ws, int64_t ws_bytes, void* stream) {
  int rc = check_common(plan, params, x, n, ws, ws_bytes);
  if (rc) return rc;
  if (!pde || !stats) return fail(PINNK_E_INVALID, "score: null pde / stats");
  PdeDesc pd;
  pd.kind = pde->kind; pd.compat_math = pde->compat_math; pd.p0 = pde->p0; pd.p1 = pde->p1;
  const int threads = 256;
  for (int64_t p0 = 0; p0 < n; p0 += plan->chunk) {
    const int64_t cn = std::min(plan->chunk, n - p0);
    ChunkCtx c = make_ctx(plan, params, x, t, p0, cn, ws, stream);
    rc = dispatch_maxk(plan->maxk, [&](auto mk) { return forward_chunk<decltype(mk)::value>(c, false); });
    if (rc) return rc;
    score_kernel<<<blocks_for(cn, threads), threads, 0, c.st>>>(c.U(), plan->js, pd, cn, abs_out ? abs_out + p0 : nullptr, stats,
                                                                  c.x, c.t ? plan->js.in_dim - 1 : plan->js.in_dim);
    PK_LAUNCH_OK();
  }
  return 0;
}

// ---- debug / micro-benchmark entry: one hidden Linear forward, Z[M,N] = X[M,K] W[N,K]^T (+bias on value rows).
// mode 0 = exact-fp32 CUDA-core GEMM, 1 = tcgen05 3xTF32 (fails if the shape is unsupported).
extern "C" int pinnk_debug_linear_fwd(const float* X, const float* W, const float* bias, float* Z, int64_t M, int32_t K,
                                      int32_t N, int32_t jet_cols, int32_t mode, void* stream) {
  if (!X || !W || !Z || M < 1 || (K % 4) || (N % 4) || jet_cols < 1) return fail(PINNK_E_INVALID, "debug_linear_fwd: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const int smc = sm_count_of(dev);
  if (mode == 1) {
    int rc = tc_linear_fwd(X, W, bias, Z, M, K, N, jet_cols, smc, (cudaStream_t)stream);
    if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_linear_fwd: shape not covered by the tcgen05 path");
    if (rc != 0) return fail(PINNK_E_CUDA, std::string("tc_linear_fwd: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    return 0;
  }
  return linear_fwd_any(X, W, bias, Z, M, K, N, jet_cols, smc, (cudaStream_t)stream, false);
}

extern "C" int pinnk_debug_linear_ks(const float* X, const float* W, const float* bias, float* Z, int64_t M, int32_t K, int32_t N,
                                     int32_t jet_cols, int32_t trans, float* ring, int64_t ring_floats, void* stream) {
  if (!X || !W || !Z || M < 1 || (K % 4) || (N % 4) || jet_cols < 1) return fail(PINNK_E_INVALID, "debug_linear_ks: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const int smc = sm_count_of(dev);
  const int rc = trans ? tc_linear_dgrad(X, W, Z, M, N, K, smc, (cudaStream_t)stream, ring, ring_floats)
                       : tc_linear_fwd(X, W, bias, Z, M, K, N, jet_cols, smc, (cudaStream_t)stream, ring, ring_floats);
  if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_linear_ks: shape not covered by the tcgen05 path");
  if (rc != 0) return fail(PINNK_E_CUDA, std::string("debug_linear_ks: ") + cudaGetErrorString(cudaGetLastError()));
  g_launches.fetch_add(1);
  return 0;
}

// dX[M,K_in] = dZ[M,N] W[N,K_in]  (mode as above)
extern "C" int pinnk_debug_linear_dgrad(const float* dZ, const float* W, float* dX, int64_t M, int32_t K, int32_t N,
                                        int32_t mode, void* stream) {
  if (!dZ || !W || !dX || M < 1 || (K % 4) || (N % 4)) return fail(PINNK_E_INVALID, "debug_linear_dgrad: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const int smc = sm_count_of(dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1) {
    int rc = tc_linear_dgrad(dZ, W, dX, M, K, N, smc, st);
    if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_linear_dgrad: shape not covered by the tcgen05 path");
    if (rc != 0) return fail(PINNK_E_CUDA, std::string("tc_linear_dgrad: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    return 0;
  }
  dim3 grid(blocks_for(M, SG_BM), blocks_for(K, SG_BN), 1);
  sgemm_kernel<true, false, EPI_STORE><<<grid, SG_THREADS, 0, st>>>(dZ, W, dX, M, K, N, N, K, K, nullptr, 1, N);
  PK_LAUNCH_OK();
  return 0;
}

// dW[N,K_in] += dZ[M,N]^T X[M,K_in] ; db[N] += sum of value-column rows of dZ  (mode as above)
extern "C" int pinnk_debug_linear_wgrad(const float* dZ, const float* X, float* dW, float* db, int64_t M, int32_t K,
                                        int32_t N, int32_t jet_cols, int32_t mode, void* stream) {
  if (!dZ || !X || !dW || M < 1 || (K % 4) || (N % 4) || jet_cols < 1 || (M % jet_cols)) return fail(PINNK_E_INVALID, "debug_linear_wgrad: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const int smc = sm_count_of(dev);
  cudaStream_t st = (cudaStream_t)stream;
  if (mode == 1) {
    int rc = tc_linear_wgrad(dZ, X, dW, db, M, K, N, jet_cols, smc, st);
    if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_linear_wgrad: shape not covered by the tcgen05 path");
    if (rc != 0) return fail(PINNK_E_CUDA, std::string("tc_linear_wgrad: ") + cudaGetErrorString(cudaGetLastError()));
    g_launches.fetch_add(1);
    return 0;
  }
  const unsigned tiles = blocks_for(N, SG_BM) * blocks_for(K, SG_BN);
  int64_t splits = (2 * (int64_t)smc + tiles - 1) / tiles;
  int64_t k_chunk = align_up((M + splits - 1) / splits, 64);
  if (k_chunk < 256) k_chunk = 256;
  splits = (M + k_chunk - 1) / k_chunk;
  dim3 grid(blocks_for(N, SG_BM), blocks_for(K, SG_BN), (unsigned)splits);
  sgemm_kernel<false, false, EPI_ATOMIC><<<grid, SG_THREADS, 0, st>>>(dZ, X, dW, N, K, M, N, K, K, nullptr, 1, k_chunk);
  PK_LAUNCH_OK();
  if (db) {
    dim3 g2(blocks_for(N, 128), (unsigned)std::min<int64_t>(M / jet_cols, 256));
    bias_grad_kernel<<<g2, 128, 0, st>>>(dZ, M / jet_cols, N, jet_cols, db);
    PK_LAUNCH_OK();
  }
  return 0;
}

// ---- debug / micro-benchmark entry: the paired reverse kernel of one Linear(128, 128) + tanh on raw tensors:
// dZprev = tanh'(from Yprev)^T (dZ W), dW += dZ^T Yprev, db += value rows of dZ  (k0, k1 = jet orders; 1 + k0 + k1 in {1, 2, 4})
extern "C" int pinnk_debug_bwd_pair(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db,
                                    int64_t M, int32_t k0, int32_t k1, void* stream) {
  if (!dZ || !W || !Yprev || !dZprev || !dW || M < 1) return fail(PINNK_E_INVALID, "debug_bwd_pair: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  int rc = tc_bwd_pair(dZ, W, Yprev, dZprev, dW, db, M, 128, 128, k0, k1, sm_count_of(dev), (cudaStream_t)stream);
  if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_bwd_pair: shape / jet layout not covered by the paired kernel");
  if (rc != 0) return fail(PINNK_E_CUDA, std::string("tc_bwd_pair: ") + cudaGetErrorString(cudaGetLastError()));
  g_launches.fetch_add(1);
  return 0;
}

// the two-launch route on the same raw tensors (reference for the paired kernel): tc_linear_dgrad_actbwd(from_y) + tc_linear_wgrad
extern "C" int pinnk_debug_bwd_split(const float* dZ, const float* W, const float* Yprev, float* dZprev, float* dW, float* db,
                                     int64_t M, int32_t k0, int32_t k1, void* stream) {
  if (!dZ || !W || !Yprev || !dZprev || !dW || M < 1) return fail(PINNK_E_INVALID, "debug_bwd_split: bad argument");
  int dev = 0;
  cudaGetDevice(&dev);
  const int smc = sm_count_of(dev);
  const char* skip = getenv("PINNK_DEBUG_SKIP_WGRAD");          // probes: the dgrad + adjoint kernel alone
  int rc = (skip && skip[0] == '1') ? 0 : tc_linear_wgrad(dZ, Yprev, dW, db, M, 128, 128, 1 + k0 + k1, smc, (cudaStream_t)stream);
  if (rc == 0) rc = tc_linear_dgrad_actbwd(dZ, W, Yprev, dZprev, M, 128, 128, k0, k1, 1, 1.f, smc, (cudaStream_t)stream, 1);
  if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "debug_bwd_split: shape / jet layout not covered");
  if (rc != 0) return fail(PINNK_E_CUDA, std::string("debug_bwd_split: ") + cudaGetErrorString(cudaGetLastError()));
  g_launches.fetch_add(2);
  return 0;
}

// ---- builder tool: per-role barrier wait cycles of the rows kernels (which = 0 forward TU, 1 backward TU)
extern "C" int pinnk_debug_stage_timers(int32_t which, uint64_t* out16, int32_t reset) {
  if (!out16) return fail(PINNK_E_INVALID, "debug_stage_timers: null output");
  cudaDeviceSynchronize();
  if (which == 0) return tc_stage_timers_fwd((unsigned long long*)out16, reset);
  if (which == 1) {          // plain / stashed-z reverse kernels + the output-jet variant (own translation unit)
    unsigned long long a[16], b[16];
    const int r1 = tc_stage_timers_bwd(a, reset), r2 = tc_stage_timers_bwd_y(b, reset);
    for (int i = 0; i < 16; ++i) out16[i] = (i == 15) ? (a[i] > b[i] ? a[i] : b[i]) : a[i] + b[i];
    return r1 ? r1 : r2;
  }
  return tc_stage_timers_wgrad((unsigned long long*)out16, reset);
}

// ---- fused optimizer tail: clip_grad_norm_ + Adam(L2) on the flat gradient (trainer.py:690-694,292-297)
static int adam_step_impl(float* const* params, const int64_t* numels, int32_t n_tensors, const float* flat_grad,
                          float* exp_avg, float* exp_avg_sq, double* scratch, int64_t step, float lr, const double* dyn,
                          float beta1, float beta2, float eps, float weight_decay, float max_norm, void* stream) {
  if (!params || !numels || !flat_grad || !exp_avg || !exp_avg_sq || !scratch || n_tensors < 1 || n_tensors > 64 ||
      (!dyn && step < 1))
    return fail(PINNK_E_INVALID, "adam_step: bad argument (1..64 tensors, step >= 1)");
  ParamTable tab;
  tab.n = n_tensors;
  int64_t off = 0;
  for (int i = 0; i < n_tensors; ++i) { tab.ptr[i] = params[i]; tab.off[i] = off; off += numels[i]; }
  tab.off[n_tensors] = off;
  cudaStream_t st = (cudaStream_t)stream;
  if (max_norm > 0.f) {
    PK_CHECK_CUDA(cudaMemsetAsync(scratch, 0, sizeof(double), st));
    sumsq_kernel<<<(unsigned)std::min<int64_t>((off + 255) / 256, 592), 256, 0, st>>>(flat_grad, off, scratch);
    PK_LAUNCH_OK();
  }
  // (double, like torch.optim.Adam's Python-float bias corrections; the device-side variant does the same in the kernel)
  const float bc1 = dyn ? 1.f : (float)(1.0 - pow((double)beta1, (double)step));
  const float bc2_sqrt = dyn ? 1.f : (float)sqrt(1.0 - pow((double)beta2, (double)step));
  adam_kernel<<<blocks_for(off, 256), 256, 0, st>>>(tab, flat_grad, exp_avg, exp_avg_sq, scratch, max_norm, lr, beta1, beta2, eps,
                                                    weight_decay, bc1, bc2_sqrt, dyn);
  PK_LAUNCH_OK();
  return 0;
}

extern "C" int pinnk_adam_step(float* const* params, const int64_t* numels, int32_t n_tensors, const float* flat_grad,
                               float* exp_avg, float* exp_avg_sq, double* scratch, int64_t step, float lr, float beta1,
                               float beta2, float eps, float weight_decay, float max_norm, void* stream) {
  return adam_step_impl(params, numels, n_tensors, flat_grad, exp_avg, exp_avg_sq, scratch, step, lr, nullptr, beta1, beta2,
                        eps, weight_decay, max_norm, stream);
}

// the same step with the step count and the learning rate read from device memory (dyn[0] = step >= 1, dyn[1] = lr), so
// that a CUDA graph holding the launch stays valid from one optimizer step to the next
extern "C" int pinnk_adam_step_dev(float* const* params, const int64_t* numels, int32_t n_tensors, const float* flat_grad,
                                   float* exp_avg, float* exp_avg_sq, double* scratch, const double* dyn, float beta1,
                                   float beta2, float eps, float weight_decay, float max_norm, void* stream) {
  if (!dyn) return fail(PINNK_E_INVALID, "adam_step_dev: null dyn");
  return adam_step_impl(params, numels, n_tensors, flat_grad, exp_avg, exp_avg_sq, scratch, 0, 0.f, dyn, beta1, beta2, eps,
                        weight_decay, max_norm, stream);
}

// ---- on-device samplers (pde_base.py:806-935): weighted draw with replacement, jittered grid
extern "C" int64_t pinnk_sample_workspace_doubles(int64_t n) { return n > 0 ? (n + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK + 1 : 1; }

extern "C" int pinnk_sample_weighted(const float* weights, int64_t n, float eps, const double* u, int64_t m, int64_t* idx_out,
                                     double* ws, int64_t ws_doubles, void* stream) {
  if (!weights || !u || !idx_out || !ws || n < 1 || m < 0) return fail(PINNK_E_INVALID, "sample_weighted: bad argument");
  const int64_t nb = (n + SAMPLE_BLOCK - 1) / SAMPLE_BLOCK;
  if (ws_doubles < nb + 1) return fail(PINNK_E_WORKSPACE, "sample_weighted: workspace smaller than pinnk_sample_workspace_doubles()");
  if (m == 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  sample_block_sums_kernel<<<(unsigned)nb, 256, 0, st>>>(weights, n, eps, ws);
  PK_LAUNCH_OK();
  sample_scan_kernel<<<1, 1024, 0, st>>>(ws, nb);
  PK_LAUNCH_OK();
  int dev = 0;
  cudaGetDevice(&dev);
  const int64_t warps_needed = m;
  const unsigned blocks = (unsigned)std::min<int64_t>((warps_needed + 7) / 8, 16 * (int64_t)sm_count_of(dev));
  sample_draw_kernel<<<blocks, 256, 0, st>>>(weights, n, eps, ws, nb, u, m, idx_out);
  PK_LAUNCH_OK();
  return 0;
}

extern "C" int pinnk_jittered_grid(const float* xs, const float* ts, int32_t n_side, const float* noise_x, const float* noise_t,
                                   float x_noise, float t_noise, float x_lo, float x_hi, float t_lo, float t_hi, float* x_out,
                                   float* t_out, void* stream) {
  if (!xs || !ts || !noise_x || !noise_t || !x_out || !t_out || n_side < 1) return fail(PINNK_E_INVALID, "jittered_grid: bad argument");
  const int64_t n = (int64_t)n_side * n_side;
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned blocks = (unsigned)std::min<int64_t>((n + 255) / 256, 32 * (int64_t)sm_count_of(dev));
  jittered_grid_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(xs, ts, n_side, noise_x, noise_t, x_noise, t_noise, x_lo, x_hi,
                                                                  t_lo, t_hi, x_out, t_out);
  PK_LAUNCH_OK();
  return 0;
}

// ---- RL sampler: Q-network forward over the candidate grid (rl_agent.py:15-88,214-229), one launch
static int dqn_forward_impl(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                            int32_t out_dim, const float* states, int64_t n, float* q_out, float* ws, int64_t ws_floats,
                            void* stream);

extern "C" int pinnk_dqn_forward(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                                 int32_t out_dim, const float* states, int64_t n, float* q_out, void* stream) {
  return dqn_forward_impl(layers, n_hidden, w_out, b_out, out_dim, states, n, q_out, nullptr, 0, stream);
}

// the same network for hidden widths that are multiples of 128 (up to 1024): the hidden Linear layers run on the tcgen05
// 3xTF32 rows kernel; ws = 2 * n * hidden floats of device scratch
extern "C" int pinnk_dqn_forward_wide(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                                      int32_t out_dim, const float* states, int64_t n, float* q_out, float* ws,
                                      int64_t ws_floats, void* stream) {
  if (!ws) return fail(PINNK_E_WORKSPACE, "dqn_forward_wide: null workspace");
  return dqn_forward_impl(layers, n_hidden, w_out, b_out, out_dim, states, n, q_out, ws, ws_floats, stream);
}

static int dqn_forward_impl(const PinnkDqnLayer* layers, int32_t n_hidden, const float* w_out, const float* b_out,
                            int32_t out_dim, const float* states, int64_t n, float* q_out, float* ws, int64_t ws_floats,
                            void* stream) {
  if (!layers || n_hidden < 1 || n_hidden > DQN_MAX_LAYERS || !w_out || out_dim < 1 || !states || !q_out || n < 0)
    return fail(PINNK_E_INVALID, "dqn_forward: bad argument (1..8 hidden layers, out_dim >= 1)");
  if (n == 0) return 0;
  DqnNet net;
  memset(&net, 0, sizeof(net));
  net.n_hidden = n_hidden;
  net.state_dim = layers[0].in_dim;
  net.hidden = layers[0].out_dim;
  net.out_dim = out_dim;
  net.W_out = w_out;
  net.b_out = b_out;
  if (net.state_dim < 1 || net.hidden < 1 || net.hidden > 1024) return fail(PINNK_E_INVALID, "dqn_forward: hidden width must be 1..1024");
  for (int l = 0; l < n_hidden; ++l) {
    const PinnkDqnLayer& L = layers[l];
    if (!L.weight || L.out_dim != net.hidden || L.in_dim != (l == 0 ? net.state_dim : net.hidden) || !(L.eps >= 0.f))
      return fail(PINNK_E_INVALID, "dqn_forward: layer shapes must chain state_dim -> hidden -> ... -> hidden");
    net.W[l] = L.weight; net.b[l] = L.bias; net.gamma[l] = L.ln_weight; net.beta[l] = L.ln_bias; net.mask[l] = L.dropout_mask;
    net.eps[l] = L.eps;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  if (ws != nullptr) {
    // wide route: hidden Linear layers on the tcgen05 rows kernel, everything else in dqn_ln_relu_rows_kernel
    const int H = net.hidden;
    if ((H % 128) != 0 || H > 1024 || net.state_dim > 8) return fail(PINNK_E_INVALID, "dqn_forward_wide: hidden must be a multiple of 128 up to 1024, state_dim <= 8");
    if (ws_floats < 2 * n * H) return fail(PINNK_E_WORKSPACE, "dqn_forward_wide: workspace smaller than 2 * n * hidden floats");
    cudaStream_t st = (cudaStream_t)stream;
    const int smc = sm_count_of(dev);
    float* A = ws;                       // activations of the previous group
    float* Zb = ws + n * H;              // GEMM output
    const unsigned blocks = (unsigned)std::min<int64_t>((n + 7) / 8, 16 * (int64_t)smc);
#define PK_DQN_ROWS(NP, FIRSTv, LASTv, Zin, l)                                                                          \
    dqn_ln_relu_rows_kernel<NP, FIRSTv, LASTv><<<blocks, 256, 0, st>>>(Zin, states, net.state_dim, net.W[0], net.b[0],  \
        net.gamma[l], net.beta[l], net.eps[l], net.mask[l], n, A, net.W_out, net.b_out, net.out_dim, q_out)
#define PK_DQN_DISPATCH(FIRSTv, LASTv, Zin, l)                                                                          \
    do { switch (H / 32) { case 4: PK_DQN_ROWS(4, FIRSTv, LASTv, Zin, l); break; case 8: PK_DQN_ROWS(8, FIRSTv, LASTv, Zin, l); break;  \
                           case 12: PK_DQN_ROWS(12, FIRSTv, LASTv, Zin, l); break; case 16: PK_DQN_ROWS(16, FIRSTv, LASTv, Zin, l); break; \
                           case 20: PK_DQN_ROWS(20, FIRSTv, LASTv, Zin, l); break; case 24: PK_DQN_ROWS(24, FIRSTv, LASTv, Zin, l); break; \
                           case 28: PK_DQN_ROWS(28, FIRSTv, LASTv, Zin, l); break; default: PK_DQN_ROWS(32, FIRSTv, LASTv, Zin, l); break; } } while (0)
    for (int l = 0; l < n_hidden; ++l) {
      const bool last = l == n_hidden - 1;
      if (l == 0) {
        if (last) PK_DQN_DISPATCH(true, true, nullptr, 0); else PK_DQN_DISPATCH(true, false, nullptr, 0);
      } else {
        int rc = tc_linear_fwd(A, net.W[l], net.b[l], Zb, n, H, H, 1, smc, st);
        if (rc == TC_UNSUPPORTED) return fail(PINNK_E_INVALID, "dqn_forward_wide: hidden width not covered by the tcgen05 rows kernel");
        if (rc != 0) return fail(PINNK_E_CUDA, std::string("dqn_forward_wide: tc_linear_fwd: ") + cudaGetErrorString(cudaGetLastError()));
        g_launches.fetch_add(H / 128);
        if (last) PK_DQN_DISPATCH(false, true, Zb, l); else PK_DQN_DISPATCH(false, false, Zb, l);
      }
      PK_LAUNCH_OK();
    }
#undef PK_DQN_DISPATCH
#undef PK_DQN_ROWS
    return 0;
  }
  const int threads = (net.hidden + 31) / 32 * 32;
  const size_t smem = dqn_smem_bytes(DQN_ROWS, net.hidden, net.state_dim);
  if (smem > 227 * 1024) return fail(PINNK_E_INVALID, "dqn_forward: hidden width needs more than 227 KB of shared memory");
  if (smem > 48 * 1024 &&
      cudaFuncSetAttribute(dqn_forward_kernel<DQN_ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
    return fail(PINNK_E_CUDA, "dqn_forward: shared memory request refused");
  const int64_t groups = (n + DQN_ROWS - 1) / DQN_ROWS;
  const unsigned blocks = (unsigned)std::min<int64_t>(groups, 16 * (int64_t)sm_count_of(dev));
  dqn_forward_kernel<DQN_ROWS><<<blocks, threads, smem, (cudaStream_t)stream>>>(net, states, n, q_out);
  PK_LAUNCH_OK();
  return 0;
}

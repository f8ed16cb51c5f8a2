"""ctypes binding of libpinnk.so (include/pinnk.h).  There is no fallback: if the CUDA library
is missing or a call fails, the caller gets an exception."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PINNK_LIB", os.path.join(HERE, "libpinnk.so"))     # PINNK_LIB: A/B builds of the same library

OP_LINEAR, OP_ACT, OP_LAYERNORM, OP_SKIP_SAVE, OP_SKIP_ADD, OP_SINCOS = 1, 2, 3, 4, 5, 6
ACT_TANH, ACT_SIN = 1, 2
(PDE_HEAT, PDE_BURGERS, PDE_KDV, PDE_ALLEN_CAHN, PDE_CAHN_HILLIARD, PDE_UT_ONLY, PDE_UT_ALLEN_CAHN_ND,
 PDE_CAHN_HILLIARD_2D, PDE_VALUE, PDE_DX, PDE_WAVE, PDE_CONVECTION, PDE_BLACK_SCHOLES, PDE_PENDULUM) = range(14)
LOSS_MSE, LOSS_MAE, LOSS_HUBER = 0, 1, 2
ABI_VERSION = 3
STEP_KEEP_STASH, STEP_REUSE_STASH = 1, 2      # pinnk_loss_step_flags (include/pinnk.h)


class PinnkOp(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in_dim", C.c_int32), ("out_dim", C.c_int32), ("act", C.c_int32),
                ("scale", C.c_float), ("eps", C.c_float), ("w_index", C.c_int32), ("b_index", C.c_int32),
                ("w_transposed", C.c_int32), ("reserved", C.c_int32), ("gw_offset", C.c_int64),
                ("gb_offset", C.c_int64)]


class PinnkJetSpec(C.Structure):
    _fields_ = [("ndirs", C.c_int32), ("order", C.c_int32 * 5), ("vec", (C.c_float * 4) * 5)]


class PinnkPde(C.Structure):
    _fields_ = [("kind", C.c_int32), ("compat_math", C.c_int32), ("p0", C.c_float), ("p1", C.c_float)]


class PinnkSegment(C.Structure):
    _fields_ = [("pde", PinnkPde), ("component", C.c_int32), ("loss_kind", C.c_int32),
                ("huber_delta", C.c_float), ("weight", C.c_float), ("row_start", C.c_int64),
                ("row_count", C.c_int64), ("pair_offset", C.c_int64), ("target", C.c_void_p),
                ("error_out", C.c_void_p), ("error_grad", C.c_void_p)]


class PinnkDqnLayer(C.Structure):
    _fields_ = [("weight", C.c_void_p), ("bias", C.c_void_p), ("ln_weight", C.c_void_p), ("ln_bias", C.c_void_p),
                ("dropout_mask", C.c_void_p), ("eps", C.c_float), ("in_dim", C.c_int32), ("out_dim", C.c_int32)]


EXPORTS = ["pinnk_plan_create", "pinnk_plan_destroy", "pinnk_plan_workspace_bytes", "pinnk_plan_ncols",
           "pinnk_plan_grad_floats", "pinnk_jets_forward", "pinnk_jets_vjp", "pinnk_loss_step", "pinnk_loss_step_flags", "pinnk_score",
           "pinnk_last_error", "pinnk_abi_version", "pinnk_launch_count", "pinnk_prof_enable", "pinnk_prof_classes",
           "pinnk_prof_class_name", "pinnk_prof_collect", "pinnk_debug_linear_fwd",
           "pinnk_debug_linear_dgrad", "pinnk_debug_linear_wgrad", "pinnk_adam_step", "pinnk_adam_step_dev", "pinnk_dqn_forward",
           "pinnk_debug_stage_timers", "pinnk_debug_bwd_pair", "pinnk_debug_bwd_split",
           "pinnk_dqn_forward_wide", "pinnk_sample_workspace_doubles", "pinnk_sample_weighted", "pinnk_jittered_grid",
           "pinnk_debug_linear_ks"]

_lib = None


class PinnkError(RuntimeError):
    pass


def load():
    """Load libpinnk.so or raise.  Never substitutes another implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PinnkError(f"{LIB_PATH} not found: build it with `python -m pinns_rl_pde_b200.build` "
                         "(there is no CPU or PyTorch fallback for this path)")
    lib = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int32
    lib.pinnk_plan_create.argtypes = [C.POINTER(PinnkOp), i32, i32, C.POINTER(PinnkJetSpec), i64, i32, C.POINTER(vp)]
    lib.pinnk_plan_create.restype = C.c_int
    lib.pinnk_plan_destroy.argtypes = [vp]
    lib.pinnk_plan_destroy.restype = None
    lib.pinnk_plan_workspace_bytes.argtypes = [vp]
    lib.pinnk_plan_workspace_bytes.restype = i64
    lib.pinnk_plan_ncols.argtypes = [vp]
    lib.pinnk_plan_ncols.restype = i32
    lib.pinnk_plan_grad_floats.argtypes = [vp]
    lib.pinnk_plan_grad_floats.restype = i64
    lib.pinnk_jets_forward.argtypes = [vp, vp, vp, vp, i64, vp, vp, i64, vp]
    lib.pinnk_jets_forward.restype = C.c_int
    lib.pinnk_jets_vjp.argtypes = [vp, vp, vp, vp, i64, vp, vp, vp, i64, vp]
    lib.pinnk_jets_vjp.restype = C.c_int
    lib.pinnk_loss_step.argtypes = [vp, vp, vp, vp, i64, C.POINTER(PinnkSegment), i32, vp, vp, vp, vp, i64, vp]
    lib.pinnk_loss_step.restype = C.c_int
    lib.pinnk_loss_step_flags.argtypes = [vp, vp, vp, vp, i64, C.POINTER(PinnkSegment), i32, vp, vp, vp, vp, i64, vp, i32]
    lib.pinnk_loss_step_flags.restype = C.c_int
    lib.pinnk_score.argtypes = [vp, vp, vp, vp, i64, C.POINTER(PinnkPde), vp, vp, vp, i64, vp]
    lib.pinnk_score.restype = C.c_int
    lib.pinnk_last_error.argtypes = []
    lib.pinnk_last_error.restype = C.c_char_p
    lib.pinnk_abi_version.argtypes = []
    lib.pinnk_abi_version.restype = i32
    lib.pinnk_launch_count.argtypes = []
    lib.pinnk_launch_count.restype = i64
    lib.pinnk_prof_enable.argtypes = [i32]
    lib.pinnk_prof_enable.restype = None
    lib.pinnk_prof_classes.argtypes = []
    lib.pinnk_prof_classes.restype = i32
    lib.pinnk_prof_class_name.argtypes = [i32]
    lib.pinnk_prof_class_name.restype = C.c_char_p
    lib.pinnk_prof_collect.argtypes = [C.POINTER(C.c_double), C.POINTER(i64), i32]
    lib.pinnk_prof_collect.restype = C.c_int
    lib.pinnk_debug_linear_fwd.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.pinnk_debug_linear_fwd.restype = C.c_int
    lib.pinnk_adam_step.argtypes = [vp, vp, i32, vp, vp, vp, vp, i64, C.c_float, C.c_float, C.c_float, C.c_float,
                                    C.c_float, C.c_float, vp]
    lib.pinnk_adam_step.restype = C.c_int
    lib.pinnk_adam_step_dev.argtypes = [vp, vp, i32, vp, vp, vp, vp, vp, C.c_float, C.c_float, C.c_float, C.c_float,
                                        C.c_float, vp]
    lib.pinnk_adam_step_dev.restype = C.c_int
    lib.pinnk_dqn_forward.argtypes = [C.POINTER(PinnkDqnLayer), i32, vp, vp, i32, vp, i64, vp, vp]
    lib.pinnk_dqn_forward.restype = C.c_int
    lib.pinnk_dqn_forward_wide.argtypes = [C.POINTER(PinnkDqnLayer), i32, vp, vp, i32, vp, i64, vp, vp, i64, vp]
    lib.pinnk_dqn_forward_wide.restype = C.c_int
    lib.pinnk_debug_linear_ks.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp, i64, vp]
    lib.pinnk_debug_linear_ks.restype = C.c_int
    lib.pinnk_debug_linear_dgrad.argtypes = [vp, vp, vp, i64, i32, i32, i32, vp]
    lib.pinnk_debug_linear_dgrad.restype = C.c_int
    lib.pinnk_debug_linear_wgrad.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, vp]
    lib.pinnk_debug_linear_wgrad.restype = C.c_int
    for fn in (lib.pinnk_debug_bwd_pair, lib.pinnk_debug_bwd_split):
        fn.argtypes = [vp, vp, vp, vp, vp, vp, i64, i32, i32, vp]
        fn.restype = C.c_int
    lib.pinnk_sample_workspace_doubles.argtypes = [i64]
    lib.pinnk_sample_workspace_doubles.restype = i64
    lib.pinnk_sample_weighted.argtypes = [vp, i64, C.c_float, vp, i64, vp, vp, i64, vp]
    lib.pinnk_sample_weighted.restype = C.c_int
    lib.pinnk_jittered_grid.argtypes = [vp, vp, i32, vp, vp] + [C.c_float] * 6 + [vp, vp, vp]
    lib.pinnk_jittered_grid.restype = C.c_int
    lib.pinnk_debug_stage_timers.argtypes = [i32, C.POINTER(C.c_uint64), i32]
    lib.pinnk_debug_stage_timers.restype = C.c_int
    if lib.pinnk_abi_version() != ABI_VERSION:
        raise PinnkError(f"libpinnk.so ABI {lib.pinnk_abi_version()} != binding {ABI_VERSION}: rebuild")
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().pinnk_last_error()
        raise PinnkError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def launch_count() -> int:
    return int(load().pinnk_launch_count())


def prof_enable(on: bool):
    load().pinnk_prof_enable(1 if on else 0)


def prof_collect():
    """{class name: (total ms, launches)} since the last collect (synchronises the recorded events)."""
    lib = load()
    n = lib.pinnk_prof_classes()
    ms = (C.c_double * n)()
    cnt = (C.c_int64 * n)()
    lib.pinnk_prof_collect(ms, cnt, n)
    return {lib.pinnk_prof_class_name(i).decode(): (ms[i], int(cnt[i])) for i in range(n)}


def debug_linear_fwd(X, W, bias, jet_cols: int, mode: int):
    """Z = X W^T (+bias on value rows) through one GEMM kernel of the library (tests / micro-benchmarks)."""
    import torch
    lib = load()
    M, K = X.shape
    N = W.shape[0]
    Z = torch.empty(M, N, dtype=torch.float32, device=X.device)
    check(lib.pinnk_debug_linear_fwd(X.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None,
                                     Z.data_ptr(), M, K, N, jet_cols, mode,
                                     C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_debug_linear_fwd")
    return Z


def debug_linear_ks(X, W, bias, jet_cols: int, trans: bool, ring=None, out=None):
    """tcgen05 forward (Z = X W^T + bias) or dgrad (Z = X W) GEMM with the K-split scratch (``ring``: fp32 device tensor or
    None for the two K-half passes)."""
    import torch
    lib = load()
    M, K = X.shape
    N = W.shape[1] if trans else W.shape[0]
    Z = out if out is not None else torch.empty(M, N, dtype=torch.float32, device=X.device)
    check(lib.pinnk_debug_linear_ks(X.data_ptr(), W.data_ptr(), bias.data_ptr() if bias is not None else None, Z.data_ptr(),
                                    M, K, N, jet_cols, 1 if trans else 0, ring.data_ptr() if ring is not None else None,
                                    ring.numel() if ring is not None else 0,
                                    C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_debug_linear_ks")
    return Z


def debug_linear_dgrad(dZ, W, mode: int):
    import torch
    lib = load()
    M, N = dZ.shape
    K = W.shape[1]
    dX = torch.empty(M, K, dtype=torch.float32, device=dZ.device)
    check(lib.pinnk_debug_linear_dgrad(dZ.data_ptr(), W.data_ptr(), dX.data_ptr(), M, K, N, mode,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_debug_linear_dgrad")
    return dX


def debug_linear_wgrad(dZ, X, jet_cols: int, mode: int):
    import torch
    lib = load()
    M, N = dZ.shape
    K = X.shape[1]
    dW = torch.zeros(N, K, dtype=torch.float32, device=dZ.device)
    db = torch.zeros(N, dtype=torch.float32, device=dZ.device)
    check(lib.pinnk_debug_linear_wgrad(dZ.data_ptr(), X.data_ptr(), dW.data_ptr(), db.data_ptr(), M, K, N, jet_cols, mode,
                                       C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_debug_linear_wgrad")
    return dW, db


def debug_bwd_layer(dZ, W, Yprev, k0: int, k1: int, pair: bool):
    """(dZprev, dW, db) of one hidden Linear(128, 128) + tanh reverse step: the paired kernel or the two-launch route."""
    import torch
    lib = load()
    M = dZ.shape[0]
    dZp = torch.empty(M, 128, dtype=torch.float32, device=dZ.device)
    dW = torch.zeros(128, 128, dtype=torch.float32, device=dZ.device)
    db = torch.zeros(128, dtype=torch.float32, device=dZ.device)
    fn = lib.pinnk_debug_bwd_pair if pair else lib.pinnk_debug_bwd_split
    check(fn(dZ.data_ptr(), W.data_ptr(), Yprev.data_ptr(), dZp.data_ptr(), dW.data_ptr(), db.data_ptr(), M, k0, k1,
             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_debug_bwd_pair" if pair else "pinnk_debug_bwd_split")
    return dZp, dW, db


def stage_timers(which: int, reset: bool = True):
    """Per-role barrier wait cycles of the rows kernels (zeros unless built with -DPINNK_STAGE_TIMERS)."""
    out = (C.c_uint64 * 16)()
    rc = load().pinnk_debug_stage_timers(which, out, 1 if reset else 0)
    names = ["tma:raw_empty", "cvt:raw_full", "cvt:empty", "cvt:work", "mma:tempty", "mma:full", "mma:issue",
             "epi:tfull", "epi:work", "tiles", "total", "total_ns", "prologue_ns", "kernel_ns", "launches", "max_loop_ns"]
    return rc, {n: int(out[i]) for i, n in enumerate(names)}

"""Drop-in implementations of the reference's hot-path calls on top of libpinnk.

    compute_residual(pde, model, x, t)   ->  PDEBase.compute_residual      (pde_base.py:577-588 + overrides)
    compute_loss(pde, model, x, t)       ->  PDEBase.compute_loss          (pde_base.py:1086-1235)
                                             HeatEquation.compute_loss     (heat_equation.py:375-623)
    model_forward(model, xt)             ->  PINNModel.forward             (neural_networks/__init__.py:144-154)
    score_residual(pde, model, x, t)     ->  |compute_residual| for RAR / RL sampling (pde_base.py:909-921,1364-1377)

``pde`` is duck-typed: the reference's own PDE objects and this package's mirrors both work
(class name, ``dimension``, ``domain``, ``time_domain``, ``config``, ``boundary_conditions``).
Every returned tensor carries an autograd graph to the model parameters whose backward is the
hand-written reverse pass; torch autograd never differentiates through the network itself.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .engine import Segment, get_engine, get_program

_PDE_NAMES = {
    "HeatEquation": "heat", "BurgersEquation": "burgers", "KdVEquation": "kdv",
    "AllenCahnEquation": "allen_cahn", "CahnHilliardEquation": "cahn_hilliard",
    "WaveEquation": "wave", "ConvectionEquation": "convection",
    "BlackScholesEquation": "black_scholes", "PendulumEquation": "pendulum",
}
_ORDER_1D = {"burgers": 2, "kdv": 3, "allen_cahn": 2, "cahn_hilliard": 4, "wave": 2, "convection": 1, "black_scholes": 2}
_KIND_1D = {"heat": L.PDE_HEAT, "burgers": L.PDE_BURGERS, "kdv": L.PDE_KDV,
            "allen_cahn": L.PDE_ALLEN_CAHN, "cahn_hilliard": L.PDE_CAHN_HILLIARD,
            "wave": L.PDE_WAVE, "convection": L.PDE_CONVECTION, "black_scholes": L.PDE_BLACK_SCHOLES}


class UnsupportedPDE(NotImplementedError):
    pass


def pde_name(pde) -> str:
    for klass in type(pde).__mro__:
        if klass.__name__ in _PDE_NAMES:
            return _PDE_NAMES[klass.__name__]
    raise UnsupportedPDE(f"{type(pde).__name__} is not on the B200 hot path "
                         f"(supported: {sorted(_PDE_NAMES)})")


def _float_param(pde, name: str, default=None) -> float:
    if hasattr(pde, "get_parameter"):
        v = pde.get_parameter(name, default=default)
    else:
        v = getattr(pde, name, default)
    if isinstance(v, torch.Tensor):
        if v.requires_grad:
            raise UnsupportedPDE("trainable PDE parameters (inverse mode) are not supported by the fused path")
        v = float(v.detach().cpu())
    if v is None:
        raise ValueError(f"Required parameter '{name}' not found in config")
    return float(v)


def _velocity_1d(pde) -> float:
    """convection_equation.py:34-41: scalar or per-dimension list under ``velocity``."""
    v = pde.get_parameter("velocity", default=1.0) if hasattr(pde, "get_parameter") else getattr(pde, "velocity", 1.0)
    if isinstance(v, (list, tuple)):
        v = v[0]
    if isinstance(v, torch.Tensor):
        if v.requires_grad:
            raise UnsupportedPDE("trainable PDE parameters (inverse mode) are not supported by the fused path")
        v = float(v.detach().reshape(-1)[0].cpu())
    return float(v)


def residual_p1(pde) -> float:
    """Second scalar parameter of the residual (PinnkPde.p1): the Black-Scholes risk-free rate."""
    return _float_param(pde, "r", 0.05) if pde_name(pde) == "black_scholes" else 0.0


def residual_spec(pde) -> Tuple[List, int, float, int]:
    """(directions, PINNK_PDE kind, p0, compat_math) for ``pde`` -- the jets its residual needs."""
    name = pde_name(pde)
    d = int(pde.dimension)
    compat = getattr(pde, "compat", "reference")
    if compat not in ("reference", "math"):
        raise ValueError("pde.compat must be 'reference' or 'math'")
    p0 = {"heat": lambda: _float_param(pde, "alpha"), "burgers": lambda: _float_param(pde, "nu", 0.01),
          "kdv": lambda: 0.0, "allen_cahn": lambda: _float_param(pde, "epsilon", 0.1),
          "cahn_hilliard": lambda: _float_param(pde, "epsilon", 0.1),
          "wave": lambda: _float_param(pde, "c", 1.0), "convection": lambda: _velocity_1d(pde),
          "black_scholes": lambda: _float_param(pde, "sigma", 0.2),
          "pendulum": lambda: _float_param(pde, "g", 9.81) / _float_param(pde, "L", 1.0)}[name]()
    unit = lambda i: tuple(1.0 if k == i else 0.0 for k in range(d + 1))
    t_dir = (unit(d), 1)
    if name in ("wave", "convection", "black_scholes", "pendulum") and d != 1:
        # wave_equation.py:78-107 degenerates to u_tt, convection_equation.py:67-76 raises in autograd.grad
        # (no allow_unused): neither is on the hot path in more than one space dimension
        raise UnsupportedPDE(f"{name}: only the 1-D operator is implemented on the B200 path")
    if name == "wave":
        return [(unit(0), 2), (unit(1), 2)], L.PDE_WAVE, p0, 0           # second-order time jets: u_tt = 2 a_t2
    if name == "pendulum":
        return [(unit(1), 2)], L.PDE_PENDULUM, p0, 0                      # an ODE in t: no spatial jets at all
    if d == 1:
        if name == "heat":
            math = compat == "math"
            return [(unit(0), 2 if math else 1), t_dir], L.PDE_HEAT, p0, int(math)
        return [(unit(0), _ORDER_1D[name]), t_dir], _KIND_1D[name], p0, 0
    if compat == "reference":
        # as written, every spatial derivative of a multi-dim residual is identically zero
        kind = L.PDE_UT_ALLEN_CAHN_ND if name == "allen_cahn" else L.PDE_UT_ONLY
        return [t_dir], kind, p0, 0
    if name == "cahn_hilliard" and d == 2:
        dirs = [((1.0, 0.0, 0.0), 4), ((0.0, 1.0, 0.0), 4), ((1.0, 1.0, 0.0), 4), ((1.0, -1.0, 0.0), 4), t_dir]
        return dirs, L.PDE_CAHN_HILLIARD_2D, p0, 0
    raise UnsupportedPDE(f"compat='math' is not implemented for {name} in {d} dimensions")


def _loss_kind(pde) -> Tuple[int, float]:
    name = pde._loss_function_name() if hasattr(pde, "_loss_function_name") else "mse"
    delta = pde._huber_delta() if hasattr(pde, "_huber_delta") else 1.0
    return {"mae": L.LOSS_MAE, "huber": L.LOSS_HUBER}.get(name, L.LOSS_MSE), float(delta)


def _trainable(model: nn.Module) -> List[nn.Parameter]:
    return get_program(model).grad_params


def _prep(model: nn.Module, x: torch.Tensor, t: Optional[torch.Tensor]):
    dev = next(model.parameters()).device
    x = x.detach().to(device=dev, dtype=torch.float32)
    if t is not None:
        t = t.detach().to(device=dev, dtype=torch.float32)
    return x, t


# ------------------------------------------------------------------ autograd bridges
class _ErrorFn(torch.autograd.Function):
    """e[rows] of one segment prototype; backward = reverse pass seeded with dL/de."""

    @staticmethod
    def forward(ctx, engine, proto: Segment, x, t, keep_stash, *params):
        e = torch.empty(proto.row_count, dtype=torch.float32, device=x.device)
        seg = Segment(**{**proto.__dict__, "error_out": e})
        # when a reverse pass will follow and the rows fit one chunk, the forward leaves its stash in the workspace and
        # backward() starts from it instead of recomputing the forward (one forward + one reverse instead of two + one)
        engine.loss_step(x, t, [seg], 1, want_grad=False, keep_stash=bool(keep_stash))
        ctx.engine, ctx.proto, ctx.x, ctx.t, ctx.token = engine, proto, x, t, engine._stash_token
        return e.view(-1, 1)

    @staticmethod
    def backward(ctx, ge):
        ge = ge.reshape(-1).to(torch.float32).contiguous()
        seg = Segment(**{**ctx.proto.__dict__, "error_grad": ge})
        eng, token = ctx.engine, ctx.token
        if token is not None and (token[1:4] != (ctx.x.data_ptr(), eng._ptr(ctx.t), ctx.x.shape[0]) or token[4] != eng.stash_signature()):
            token = None            # rows or parameters changed since the forward: recompute
        _, flat = eng.loss_step(ctx.x, ctx.t, [seg], 1, want_grad=True, reuse_token=token)
        grads = ctx.engine.program.split_flat(flat)
        return (None, None, None, None, None, *grads)


class _JetsFn(torch.autograd.Function):
    """Output jets U[n, C]; backward = reverse pass seeded with dL/dU."""

    @staticmethod
    def forward(ctx, engine, x, t, *params):
        ctx.engine, ctx.x, ctx.t = engine, x, t
        return engine.jets_forward(x, t)

    @staticmethod
    def backward(ctx, gU):
        flat = ctx.engine.jets_vjp(ctx.x, ctx.t, gU.to(torch.float32).contiguous())
        return (None, None, None, *ctx.engine.program.split_flat(flat))


class _LossFn(torch.autograd.Function):
    """Component losses [n_components] with the per-component gradients computed eagerly."""

    @staticmethod
    def forward(ctx, calls, n_components, program, *params):
        dev = calls[0][1].device
        want_grad = any(ctx.needs_input_grad[3:])
        sums = torch.zeros(n_components, dtype=torch.float64, device=dev)
        G = torch.zeros(n_components, program.grad_floats, dtype=torch.float32, device=dev) if want_grad else None
        for engine, x, t, segments in calls:
            comp = segments[0].component
            assert all(s.component == comp for s in segments)
            engine.loss_step(x, t, segments, n_components, want_grad, None,
                             G[comp] if want_grad else None, sums)
        ctx.G, ctx.program = G, program
        return sums.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        flat = (g.to(torch.float32).unsqueeze(0) @ ctx.G).squeeze(0)
        return (None, None, None, *ctx.program.split_flat(flat))


class _LazyLossFn(torch.autograd.Function):
    """Component losses [n_components] of ONE merged call (all row sets behind each other, small batches): the forward
    is a forward-only sweep that leaves its stash in the workspace; backward(g) is ONE reverse sweep whose seeds carry
    g[component], so whatever weights autograd applies to the components (fixed, adaptive, one component at a time with
    retain_graph) the gradient is exact.  A second backward, or one after something else used the engine, recomputes the
    forward first.  ~27 launches per loss + backward instead of ~70 for three eager per-component passes."""

    @staticmethod
    def forward(ctx, call, n_components, program, *params):
        engine, x, t, segments = call
        want_grad = any(ctx.needs_input_grad[3:])
        sums = torch.zeros(n_components, dtype=torch.float64, device=x.device)
        engine.loss_step(x, t, segments, n_components, False, None, None, sums, keep_stash=want_grad)
        ctx.call, ctx.n, ctx.program = call, n_components, program
        ctx.token = engine._stash_token if want_grad else None
        ctx.versions = [p._version for p in params]
        ctx.params = params
        return sums.to(torch.float32)

    @staticmethod
    def backward(ctx, g):
        engine, x, t, segments = ctx.call
        if [p._version for p in ctx.params] != ctx.versions:
            raise RuntimeError("a model parameter was modified in place between compute_loss and backward")
        scale = [float(v) for v in g.detach().to(torch.float32).reshape(-1).tolist()]
        token = ctx.token
        if token is not None and (token[1:4] != (x.data_ptr(), engine._ptr(t), x.shape[0])
                                  or token[4] != engine.stash_signature()):
            token = None
        ctx.token = None                    # the reverse sweep consumes the stash
        _, flat = engine.loss_step(x, t, segments, ctx.n, True, scale, None, None, reuse_token=token)
        return (None, None, None, *ctx.program.split_flat(flat))


# ------------------------------------------------------------------ public functions
def jets(model: nn.Module, xt: torch.Tensor, directions: Sequence) -> torch.Tensor:
    """Differentiable (w.r.t. parameters) output jets of the network: column 0 = u, then the
    normalised Taylor coefficients along each direction."""
    xt, _ = _prep(model, xt, None)
    eng = get_engine(model, directions, xt.shape[0])
    return _JetsFn.apply(eng, xt, None, *_trainable(model))


def model_forward(model: nn.Module, xt: torch.Tensor) -> torch.Tensor:
    """``model(xt)`` -> [n, 1] through the CUDA path (value column only)."""
    return jets(model, xt, [])


def compute_residual(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
    dirs, kind, p0, cm = residual_spec(pde)
    if pde_name(pde) in ("black_scholes", "pendulum"):
        for p in model.parameters():            # black_scholes.py:61-63, pendulum_equation.py:79-81: side effect of the reference
            p.requires_grad_(True)
    x, t = _prep(model, x, t)
    n = x.shape[0]
    eng = get_engine(model, dirs, n)
    proto = Segment(kind=kind, row_start=0, row_count=n, p0=p0, p1=residual_p1(pde), compat_math=cm)
    if not model.training:
        model.train()   # pde_base.py:638 -- the reference flips the model into training mode here
    keep = torch.is_grad_enabled() and len(eng.program.grad_params) > 0     # (needs_input_grad ignores no_grad())
    return _ErrorFn.apply(eng, proto, x, t, keep, *eng.program.grad_params)


def compute_derivatives(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, temporal_derivatives=None,
                        spatial_derivatives=None) -> Dict[str, torch.Tensor]:
    """``PDEBase.compute_derivatives`` (pde_base.py:590-794), the documented building block of plugin PDEs
    (CONTRIBUTING.md:152-244), from ONE forward jet pass (``pinnk_jets_forward``); every returned ``[N, 1]`` tensor is
    differentiable w.r.t. the model parameters through ``pinnk_jets_vjp`` (not w.r.t. ``x`` / ``t``: derivatives of
    derivatives are requested as higher orders here, not by differentiating the results again).

    Same keys as the reference: ``dt`` / ``dt2``; 1-D ``dx`` / ``dx{i}`` and ``laplacian``; multi-dim ``dx1``, ``dx1x1``,
    ... per axis and ``laplacian``.  ``pde.compat == "reference"`` (default) reproduces the reference's bookkeeping:
      * 1-D: ONE more differentiation per LISTED order, starting from ``u`` (pde_base.py:694-732), so the k-th listed
        non-zero order holds the k-th derivative whatever its key says -- ``spatial_derivatives=[2]`` puts u_x under
        ``"dx2"`` and ``"laplacian"``; the same for the temporal list (:655-692);
      * dimension >= 2: every spatial entry is zero (differentiation w.r.t. a fresh slice of x).
    ``compat == "math"`` returns what the keys say (true i-th derivatives along each axis, true Laplacian)."""
    temporal = sorted(set(int(i) for i in (temporal_derivatives or [])))
    spatial = sorted(set(int(i) for i in (spatial_derivatives or [])))
    if temporal and max(temporal) > 2:
        raise ValueError(f"Temporal derivative order {max(temporal)} is not supported. Maximum order is 2.")
    if spatial and max(spatial) > 4:
        raise ValueError(f"Spatial derivative order {max(spatial)} is not supported. Maximum order is 4.")
    compat = getattr(pde, "compat", "reference")
    if compat not in ("reference", "math"):
        raise ValueError("pde.compat must be 'reference' or 'math'")
    as_written = compat == "reference"
    d = int(pde.dimension)
    for p in model.parameters():                 # pde_base.py:634-638: side effects of the reference
        p.requires_grad_(True)
    if not model.training:
        model.train()
    t_listed = [i for i in temporal if i > 0]
    s_listed = [i for i in spatial if i > 0]
    # derivative order actually stored under each listed key
    t_real = {i: (k + 1 if as_written else i) for k, i in enumerate(t_listed)}
    s_real = {i: (k + 1 if as_written else i) for k, i in enumerate(s_listed)}
    if as_written and t_listed and t_listed[0] != 1:
        pass                                      # (order 2 listed alone: one differentiation of u, like the spatial quirk)
    unit = lambda i: tuple(1.0 if k == i else 0.0 for k in range(d + 1))
    dirs, col = [], {}
    c = 1
    kt = max(t_real.values()) if t_real else 0
    if kt:
        dirs.append((unit(d), kt))
        col["t"] = c
        c += kt
    spatial_live = bool(s_listed) and (d == 1 or not as_written)
    ks = max(s_real.values()) if s_real else 0
    if spatial_live:
        for a in range(d):
            dirs.append((unit(a), ks))
            col[a] = c
            c += ks
    x, t = _prep(model, x, t)
    U = jets(model, torch.cat([x, t], dim=1), dirs)
    fact = (1.0, 1.0, 2.0, 6.0, 24.0)
    out: Dict[str, torch.Tensor] = {}
    for i in t_listed:
        k = t_real[i]
        out["dt" if i == 1 else f"dt{i}"] = U[:, col["t"] + k - 1:col["t"] + k] * fact[k]
    if s_listed:
        if d == 1:
            for i in s_listed:
                k = s_real[i]
                out["dx" if i == 1 else f"dx{i}"] = U[:, col[0] + k - 1:col[0] + k] * fact[k]
            if 2 in spatial:
                out["laplacian"] = out["dx2"]
        else:
            zero = torch.zeros(x.shape[0], 1, dtype=torch.float32, device=x.device)
            for a in range(d):
                name = f"x{a + 1}"
                for order in s_listed:
                    for i in range(1, order + 1):          # pde_base.py:741-778: keys d<name>, d<name><name>, ...
                        key = "d" + name * i
                        out[key] = (U[:, col[a] + i - 1:col[a] + i] * fact[i]) if spatial_live else zero
            if 2 in spatial:
                lap = out["dx1x1"]
                for a in range(1, d):
                    lap = lap + out["d" + f"x{a + 1}" * 2]
                out["laplacian"] = lap
    return out


def score_residual(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, want_abs: bool = True,
                   stats: Optional[torch.Tensor] = None):
    """Forward-only |r| [n] and stats [sum|r|, sum r^2, max|r|, count] (fp64 on device)."""
    dirs, kind, p0, cm = residual_spec(pde)
    x, t = _prep(model, x, t)
    eng = get_engine(model, dirs, x.shape[0])
    return eng.score(x, t, kind, p0, cm, want_abs, stats, residual_p1(pde))


def _weights(pde, heat: bool) -> Tuple[float, float, float, float, bool]:
    """(w_res, w_bc, w_ic, w_smooth, adaptive) exactly as the reference resolves them."""
    training = getattr(pde.config, "training", None)
    if heat:
        if training is None:
            return 1.0, 10.0, 10.0, 0.0, False
        if isinstance(training, dict):
            lw = training.get("loss_weights", {})
            aw = training.get("adaptive_weights", {})
            adaptive = aw.get("enabled", False) if isinstance(aw, dict) else False
        else:
            lw = getattr(training, "loss_weights", {}) or {}
            aw = getattr(training, "adaptive_weights", None)
            adaptive = bool(getattr(aw, "enabled", False)) if aw is not None else False
        if isinstance(lw, dict):
            return (lw.get("pde", lw.get("residual", 1.0)), lw.get("boundary", 10.0), lw.get("initial", 10.0),
                    lw.get("smoothness", 0.0), adaptive)
        return (getattr(lw, "pde", getattr(lw, "residual", 1.0)), getattr(lw, "boundary", 10.0),
                getattr(lw, "initial", 10.0), getattr(lw, "smoothness", 0.0), adaptive)
    # base class (pde_base.py:1171-1224): hasattr probes fail on plain dicts
    smooth = 0.0
    if hasattr(training, "loss_weights") and training.loss_weights:
        smooth = training.loss_weights.get("smoothness", 0.0)
    adaptive = (training is not None and hasattr(training, "adaptive_weights")
                and training.adaptive_weights.enabled)
    if training is not None and hasattr(training, "loss_weights") and training.loss_weights:
        lw = training.loss_weights
        return lw.get("pde", lw.get("residual", 1.0)), lw.get("boundary", 10.0), lw.get("initial", 10.0), smooth, adaptive
    return 1.0, 10.0, 10.0, smooth, adaptive


def _data_loss(pde, model) -> torch.Tensor:
    obs = getattr(pde, "observation_data", None)
    dev = next(model.parameters()).device
    if not obs:
        return torch.zeros((), device=dev)            # (device-side fill: no host copy on the per-step path)
    u = model_forward(model, torch.cat([obs["x"], obs["t"]], dim=1))
    return pde._apply_loss_fn(u - obs["u"].to(dev))


def _cached_rows(pde, key, build):
    """Boundary / initial rows and targets that are pure functions of the PDE object (linspace grids, no RNG): built once
    per (device, domain, boundary functions, initial condition, sizes) instead of ~20 tiny torch launches per step -- at the
    reference's batch sizes those launches were a quarter of a step.  Nothing is stored while a CUDA graph is being captured
    (the tensors would live in the graph's private pool)."""
    ic = getattr(getattr(pde, "config", None), "initial_condition", None)
    if isinstance(ic, dict) and ic.get("type") == "random" and key and key[0] != "loss_weights":
        # the `random` initial condition is amp * (2 * rand_like - 1) (pde_base.py:533-538): the reference redraws it on
        # every compute_loss call (once per boundary_conditions entry, once for the initial term), so neither the targets
        # nor the torch RNG stream may be frozen by a cache
        return build()
    cache = pde.__dict__.setdefault("_pinnk_rows_cache", {})
    hit = cache.get(key)
    if hit is not None:
        return hit
    val = build()
    if not (torch.cuda.is_available() and torch.cuda.is_current_stream_capturing()):
        if len(cache) >= 8:
            cache.clear()
        cache[key] = val
    return val


def _rows_key(pde, dev, *extra):
    ic = getattr(pde.config, "initial_condition", None)
    return (str(dev), repr(pde.domain), repr(pde.time_domain), tuple((k, id(f)) for k, f in pde.boundary_conditions.items()),
            repr(ic), getattr(pde, "compat", "reference")) + tuple(extra)


def data_term_active(pde) -> bool:
    """True when compute_loss has a data term or a non-forward training mode (pde_base.py:1150-1233): observation data
    attached, or mode in {inverse, data_only, data_augmented}.  The fused / sharded / adaptive steps build the objective
    from the three physics components only, so they must not be used then."""
    mode = pde._training_mode() if hasattr(pde, "_training_mode") else "forward"
    return bool(getattr(pde, "observation_data", None)) or mode != "forward"


def _require_physics_only(pde, what: str):
    if data_term_active(pde):
        raise NotImplementedError(
            f"{what} covers the residual / boundary / initial objective only; this PDE has observation data or a "
            "non-forward training mode (data term, data_only gating): use compute_loss(...)['total'].backward()")


MERGE_ALL_MAX_POINTS = 32768    # fused step: up to this many collocation rows, ALL row sets go through one pass


def _build_calls(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, n_global: Optional[int] = None,
                 merge_value_rows: bool = False, merge_all_rows: bool = False):
    """The libpinnk calls (engine, rows, segments) that make up compute_loss: residual rows (component 0), boundary
    rows (1), initial rows (2), exactly the point sets and targets the reference builds."""
    name = pde_name(pde)
    heat = name == "heat"
    dim = int(pde.dimension)
    if not heat and dim != 1:
        # the reference builds 2-column boundary points for a (dim+1)-input model and dies in mm
        raise RuntimeError(f"compute_loss: boundary points have 2 columns but the model takes {dim + 1} "
                           "(the reference's PDEBase.compute_loss fails the same way for dimension >= 2)")
    dirs, kind, p0, cm = residual_spec(pde)
    x, t = _prep(model, x, t)
    dev = x.device
    n = x.shape[0]
    ng = n if n_global is None else int(n_global)
    lk, delta = _loss_kind(pde)
    mk = dict(loss_kind=lk, huber_delta=delta)
    if not model.training:
        model.train()
    program = get_program(model)
    calls = []
    eng_r = get_engine(model, dirs, n, program=program)
    calls.append((eng_r, x, t, [Segment(kind=kind, row_start=0, row_count=n, component=0, weight=1.0 / max(n, 1),
                                        p0=p0, p1=residual_p1(pde), compat_math=cm, **mk)]))
    dom, td = pde.domain, pde.time_domain
    if heat:
        training = getattr(pde.config, "training", None)
        if training is not None:
            if isinstance(training, dict):
                nb = training.get("num_boundary_points", training.get("num_collocation_points", ng) // 10)
                ni = training.get("num_initial_points", training.get("num_collocation_points", ng) // 5)
            else:
                nb = getattr(training, "num_boundary_points", training.num_collocation_points // 10)
                ni = getattr(training, "num_initial_points", training.num_collocation_points // 5)
        else:
            nb, ni = max(ng // 10, 10), max(ng // 5, 10)
        t_max = pde.config.time_domain[1]
        t_early = t_max * 0.01
        n_early = max(nb // 4, 1)

        def heat_time_rows():
            return torch.cat([torch.linspace(0, t_early, n_early, device=dev),
                              torch.linspace(t_early, t_max, nb - n_early, device=dev)]).reshape(-1, 1)

        if dim == 1:
            x_min, x_max = pde.config.domain[0]

            def build_heat_rows():
                tb = heat_time_rows()
                pts = torch.cat([torch.cat([torch.full((nb, 1), x_min, device=dev), tb], dim=1),
                                 torch.cat([torch.full((nb, 1), x_max, device=dev), tb], dim=1)], dim=0)
                xb10 = (x_max - x_min) * 0.1
                xi = torch.cat([torch.linspace(x_min, x_min + xb10, ni // 4, device=dev),
                                torch.linspace(x_min + xb10, x_max - xb10, ni // 2, device=dev),
                                torch.linspace(x_max - xb10, x_max, ni // 4, device=dev)]).reshape(-1, 1)
                ti = torch.zeros_like(xi)
                if "initial" in pde.boundary_conditions:
                    target = pde.boundary_conditions["initial"](xi, ti)
                else:
                    target = torch.sin(pde.config.initial_condition.get("frequency", 2.0) * torch.pi * xi)
                return pts, xi, ti, target.detach()

            pts, xi, ti, target = _cached_rows(pde, _rows_key(pde, dev, "heat1d", nb, ni, repr(pde.config.domain),
                                                              repr(pde.config.time_domain)), build_heat_rows)
            eng_b = get_engine(model, [((1.0, 0.0), 1)], 2 * nb, whole=True, program=program)
            calls.append((eng_b, pts, None, [
                Segment(kind=L.PDE_VALUE, row_start=0, row_count=nb, component=1, weight=1.0 / nb, pair_offset=nb, **mk),
                Segment(kind=L.PDE_DX, row_start=0, row_count=nb, component=1, weight=1.0 / nb, pair_offset=nb, **mk)]))
        else:
            per_axis = max(nb // (2 * dim), 1)
            lo_pts, hi_pts = [], []
            for axis in range(dim):
                free = torch.empty(per_axis, dim, device=dev)
                for d in range(dim):
                    lo, hi = pde.config.domain[d]
                    free[:, d] = torch.rand(per_axis, device=dev) * (hi - lo) + lo
                t_axis = torch.rand(per_axis, 1, device=dev) * (td[1] - td[0]) + td[0]
                cmin, cmax = free.clone(), free.clone()
                cmin[:, axis], cmax[:, axis] = pde.config.domain[axis]
                lo_pts.append(torch.cat([cmin, t_axis], dim=1))
                hi_pts.append(torch.cat([cmax, t_axis], dim=1))
            pts = torch.cat(lo_pts + hi_pts, dim=0)
            eng_b = get_engine(model, [], pts.shape[0], whole=True, program=program)
            calls.append((eng_b, pts, None, [
                Segment(kind=L.PDE_VALUE, row_start=a * per_axis, row_count=per_axis, component=1,
                        weight=1.0 / per_axis, pair_offset=dim * per_axis, **mk) for a in range(dim)]))
            xi = torch.empty(ni, dim, device=dev)
            for d in range(dim):
                lo, hi = pde.config.domain[d]
                xi[:, d] = torch.rand(ni, device=dev) * (hi - lo) + lo
            ti = torch.zeros(ni, 1, device=dev)
            if "initial" in pde.boundary_conditions:
                target = pde.boundary_conditions["initial"](xi, ti)
            else:
                k = pde.config.initial_condition.get("frequency", 2.0)
                target = torch.ones(ni, 1, device=dev)
                for d in range(dim):
                    target = target * torch.sin(k * torch.pi * xi[:, d:d + 1])
        n_i = xi.shape[0]
        eng_i = get_engine(model, [], n_i, program=program)
        calls.append((eng_i, torch.cat([xi, ti], dim=1), None, [
            Segment(kind=L.PDE_VALUE, row_start=0, row_count=n_i, component=2, weight=1.0 / max(n_i, 1),
                    target=target.detach().to(torch.float32).reshape(-1).contiguous(), **mk)]))
    else:
        def build_rows():
            # built on the device (no host copy: the call may be inside a CUDA-graph capture); same float32 values as
            # torch.tensor([x_min, x_max])
            xb = torch.cat([torch.full((1, 1), float(dom[0][0]), dtype=torch.float32, device=dev),
                            torch.full((1, 1), float(dom[0][1]), dtype=torch.float32, device=dev)], dim=0)
            tb = torch.linspace(td[0], td[1], 100, device=dev).reshape(-1, 1)
            xb = xb.repeat_interleave(len(tb), dim=0)
            tb = tb.repeat(len(xb) // len(tb), 1)
            bc_targets = [fn(xb, tb).detach().to(torch.float32).reshape(-1).contiguous()
                          for fn in pde.boundary_conditions.values()]
            xi = torch.linspace(dom[0][0], dom[0][1], 100, device=dev).reshape(-1, 1)
            ti = torch.zeros_like(xi)
            if "initial" in pde.boundary_conditions:
                target = pde.boundary_conditions["initial"](xi, ti)
            else:
                ic = pde.config.initial_condition
                if ic.get("type", "sine") == "sine" and "amplitude" in ic and "frequency" in ic:
                    target = ic["amplitude"] * torch.sin(ic["frequency"] * torch.pi * xi)
                else:
                    target = pde._create_boundary_condition("initial", ic)(xi, ti)
            ic_target = target.detach().to(torch.float32).reshape(-1).contiguous()
            return xb, tb, bc_targets, xi, ti, ic_target, torch.cat([xb, xi], dim=0), torch.cat([tb, ti], dim=0)

        xb, tb, bc_targets, xi, ti, ic_target, x_all, t_all = _cached_rows(pde, _rows_key(pde, dev, "1d"), build_rows)
        nbp = xb.shape[0]
        segs = [Segment(kind=L.PDE_VALUE, row_start=0, row_count=nbp, component=1, weight=1.0 / nbp, target=tgt, **mk)
                for tgt in bc_targets]
        if merge_value_rows and segs:
            # boundary and initial rows are both value-only rows (one jet column): ONE pass over [boundary rows; initial
            # rows] with one segment per error functional instead of two passes of ~30 tiny launches each.  Only valid
            # when all components accumulate into one gradient buffer (the fused trainer step).
            segs.append(Segment(kind=L.PDE_VALUE, row_start=nbp, row_count=100, component=2, weight=0.01,
                                target=ic_target, **mk))
            calls.append((get_engine(model, [], nbp + 100, program=program), x_all, t_all, segs))
        else:
            if segs:
                calls.append((get_engine(model, [], nbp, program=program), xb, tb, segs))
            calls.append((get_engine(model, [], 100, program=program), xi, ti, [
                Segment(kind=L.PDE_VALUE, row_start=0, row_count=100, component=2, weight=0.01, target=ic_target, **mk)]))

    if merge_all_rows and len(calls) > 1:
        # Small batches are bound by the number of launches, not by the kernels: put boundary / initial rows BEHIND the
        # collocation rows of the same call (they ride through the jet kernels with unused derivative columns -- a few
        # hundred rows) so that a step is one forward and one reverse sweep (~27 launches) instead of two or three.  The
        # value / d/dx error functionals read the same columns in the residual's jet layout (direction 0 is x).  Like
        # merge_value_rows this is only valid when every component accumulates into one gradient buffer.
        import dataclasses
        rows, segs, off = [], [], 0
        for _, xx, tt, ss in calls:
            pts = xx if tt is None else torch.cat([xx, tt], dim=1)
            rows.append(pts)
            segs.extend(dataclasses.replace(sg, row_start=sg.row_start + off) for sg in ss)
            off += pts.shape[0]
        eng = get_engine(model, dirs, off, whole=True, program=program)
        if eng.chunk >= off:                       # paired rows (periodic BC) must share a chunk with their partners
            calls = [(eng, torch.cat(rows, dim=0), None, segs)]
    return calls, _weights(pde, heat)


def loss_components_and_grads(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, n_global: Optional[int] = None):
    """Component losses [3] (residual, boundary, initial) AND the gradient of each one as a row of ``G [3, P]``
    (``model.parameters()`` order), no autograd graph: what adaptive re-weighting needs (trainer.py:580-634).  The reference
    gets the per-component gradient norms of its LRW strategy from three extra ``backward(retain_graph=True)`` passes and then
    differentiates the weighted total a fourth time; here one reverse pass per row set fills ``G`` and any weighting is
    ``w @ G``."""
    _require_physics_only(pde, "loss_components_and_grads (adaptive re-weighting step)")
    calls, weights = _build_calls(pde, model, x, t, n_global)
    program = calls[0][0].program
    dev = calls[0][1].device
    sums = torch.zeros(3, dtype=torch.float64, device=dev)
    G = torch.zeros(3, program.grad_floats, dtype=torch.float32, device=dev)
    for engine, xx, tt, segments in calls:
        comp = segments[0].component
        assert all(sg.component == comp for sg in segments)
        engine.loss_step(xx, tt, segments, 3, True, None, G[comp], sums)
    return sums.to(torch.float32), G, weights


def _eager_loss_grad() -> bool:
    """PINNK_EAGER_LOSS_GRAD=1: per-component gradients computed inside compute_loss at every batch size (A/B)."""
    import os
    return os.environ.get("PINNK_EAGER_LOSS_GRAD", "0") == "1"


def loss_components(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, n_global: Optional[int] = None):
    """The three physics components [residual, boundary, initial] as one differentiable tensor, plus the
    weights the reference would combine them with.  ``n_global``: number of collocation rows of the whole
    (possibly sharded) batch -- the reference derives default BC/IC point counts from it."""
    lazy = x.shape[0] <= MERGE_ALL_MAX_POINTS and not _eager_loss_grad()
    calls, weights = _build_calls(pde, model, x, t, n_global, merge_value_rows=lazy, merge_all_rows=lazy)
    program = calls[0][0].program
    if lazy and len(calls) == 1:
        comp = _LazyLossFn.apply(calls[0], 3, program, *program.grad_params)
    elif lazy:                                    # could not merge (one chunk does not hold all rows): separate row sets
        calls, weights = _build_calls(pde, model, x, t, n_global)
        comp = _LossFn.apply(calls, 3, program, *program.grad_params)
    else:
        comp = _LossFn.apply(calls, 3, program, *program.grad_params)
    return comp, weights


def _smoothness_calls(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, program):
    """The finite-difference smoothness regulariser of HeatEquation.compute_loss (heat_equation.py:625-650) as libpinnk
    calls: for every spatial axis the value-only rows [x + eps e_d ; x ; x - eps e_d] (clamped to the domain as the
    reference clamps them) go through one forward sweep, and two paired mean-|e| segments -- e = u(x + eps) - u(x) and
    e = u(x) - u(x - eps), weight 1 / (N eps) -- give mean|du_forward| + mean|du_backward| in component 3 and seed the
    reverse pass with sign(e) / (N eps).  Groups of at most chunk / 3 points keep partners inside one chunk."""
    eps = 1e-4
    x, t = _prep(model, x, t)
    n = x.shape[0]
    calls = []
    if n == 0:
        return calls
    eng = get_engine(model, [], 3 * n, whole=True, program=program)
    group = max(1, min(n, eng.chunk // 3))
    for d in range(int(pde.dimension)):
        lo, hi = float(pde.domain[d][0]), float(pde.domain[d][1])
        for g0 in range(0, n, group):
            xs, ts = x[g0:g0 + group], t[g0:g0 + group]
            b = xs.shape[0]
            xp, xm = xs.clone(), xs.clone()
            xp[:, d:d + 1] = torch.clamp(xs[:, d:d + 1] + eps, lo, hi)
            xm[:, d:d + 1] = torch.clamp(xs[:, d:d + 1] - eps, lo, hi)
            w = 1.0 / (n * eps)
            segs = [Segment(kind=L.PDE_VALUE, row_start=0, row_count=b, component=3, weight=w, pair_offset=b,
                            loss_kind=L.LOSS_MAE),
                    Segment(kind=L.PDE_VALUE, row_start=b, row_count=b, component=3, weight=w, pair_offset=b,
                            loss_kind=L.LOSS_MAE)]
            calls.append((eng, torch.cat([xp, xs, xm], dim=0), torch.cat([ts, ts, ts], dim=0), segs))
    return calls


def loss_step_flat(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, n_global: Optional[int] = None,
                   res_scale: float = 1.0, rest_scale: float = 1.0, flat: Optional[torch.Tensor] = None):
    """compute_loss and the gradient of its weighted total in ONE pass per row set, with the final weights folded into
    the seeds of the reverse pass -- the trainer's inner step when the weights are known up front
    (trainer.py:578,689 with fixed ``loss_weights``).  No per-component gradient buffers, no autograd graph.

    Returns (components fp32 [4] = residual, boundary, initial means and the Heat smoothness term (0 when its weight is 0);
    weights (w_res, w_bc, w_ic, w_smooth); flat gradient of ``res_scale * (w_res * residual + w_smooth * smoothness) +
    rest_scale * (w_bc * boundary + w_ic * initial)`` in ``model.parameters()`` order).  ``res_scale`` / ``rest_scale`` are
    the shard weights of the data-parallel step (parallel.py): residual and smoothness are means over this rank's
    collocation rows, boundary / initial rows are replicated."""
    _require_physics_only(pde, "loss_step_flat (fused trainer step)")
    calls, weights = _build_calls(pde, model, x, t, n_global, merge_value_rows=True,
                                  merge_all_rows=x.shape[0] <= MERGE_ALL_MAX_POINTS)
    w_res, w_bc, w_ic, w_smooth, adaptive = weights
    if adaptive:
        w_res = w_bc = w_ic = 1.0
    program = calls[0][0].program
    dev = calls[0][1].device
    smooth_on = bool(w_smooth) and pde_name(pde) == "heat"          # only HeatEquation.compute_loss has the term
    if smooth_on:
        calls = calls + _smoothness_calls(pde, model, x, t, program)
    if flat is None:
        flat = torch.zeros(program.grad_floats, dtype=torch.float32, device=dev)
    else:
        flat.zero_()
    sums = torch.zeros(4, dtype=torch.float64, device=dev)
    scale = [res_scale * w_res, rest_scale * w_bc, rest_scale * w_ic, res_scale * float(w_smooth)]
    for engine, xx, tt, segments in calls:
        engine.loss_step(xx, tt, segments, 4, True, scale, flat, sums)
    return sums.to(torch.float32), (w_res, w_bc, w_ic, float(w_smooth) if smooth_on else 0.0), flat


def compute_loss(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor) -> Dict[str, torch.Tensor]:
    comp, (w_res, w_bc, w_ic, w_smooth, adaptive) = loss_components(pde, model, x, t)
    dev = comp.device
    smooth = torch.zeros((), device=dev)
    if pde_name(pde) == "heat" and w_smooth > 0:
        xs, ts = _prep(model, x, t)
        smooth = _heat_smoothness(pde, model, xs, ts)
    has_data = bool(getattr(pde, "observation_data", None))
    data = _data_loss(pde, model)
    data_w = pde._data_loss_weight(1.0) if hasattr(pde, "_data_loss_weight") else 1.0
    mode = pde._training_mode() if hasattr(pde, "_training_mode") else "forward"
    active = 0.0 if mode == "data_only" else 1.0
    if mode in ("inverse", "data_only", "data_augmented") and data_w <= 0.0:
        data_w = 1.0
    losses = {"residual": comp[0], "boundary": comp[1], "initial": comp[2], "smoothness": smooth, "data": data}
    # total = sum_c w_c comp_c (+ smoothness, + data) as ONE dot product with a cached weight vector: at small batches the
    # eight scalar torch ops (and their autograd nodes) of the term-by-term sum were a tenth of the step
    wts = (active, active, active) if adaptive else (active * w_res, active * w_bc, active * w_ic)
    wv = _cached_rows(pde, ("loss_weights", str(dev), wts),
                      lambda: torch.cat([torch.full((1,), float(w), dtype=torch.float32, device=dev) for w in wts]))
    total = torch.dot(comp, wv)
    if w_smooth:
        total = total + w_smooth * smooth
    if has_data:
        total = total + data_w * data
    losses["total"] = total
    return losses


def _heat_smoothness(pde, model, x, t):
    """heat_equation.py:625-650: finite-difference smoothness regulariser on model values."""
    eps = 1e-4
    u_c = model_forward(model, torch.cat([x, t], dim=1))
    out = torch.tensor(0.0, device=x.device)
    for d in range(int(pde.dimension)):
        xp, xm = x.clone(), x.clone()
        xp[:, d:d + 1] = torch.clamp(x[:, d:d + 1] + eps, pde.domain[d][0], pde.domain[d][1])
        xm[:, d:d + 1] = torch.clamp(x[:, d:d + 1] - eps, pde.domain[d][0], pde.domain[d][1])
        u_p = model_forward(model, torch.cat([xp, t], dim=1))
        u_m = model_forward(model, torch.cat([xm, t], dim=1))
        out = out + torch.mean(torch.abs((u_p - u_c) / eps)) + torch.mean(torch.abs((u_c - u_m) / eps))
    return out


def loss_and_flat_grad(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor):
    """compute_loss followed by the gradient of losses['total'] as ONE flat buffer in
    ``model.parameters()`` order -- the unit that data-parallel ranks all-reduce
    (trainer.py:578,689 collapsed into one call)."""
    program = get_program(model)
    with torch.enable_grad():
        losses = compute_loss(pde, model, x, t)
        grads = torch.autograd.grad(losses["total"], program.grad_params, allow_unused=True)
    flat = torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1)
                      for p, g in zip(program.grad_params, grads)])
    return {k: v.detach() for k, v in losses.items()}, flat

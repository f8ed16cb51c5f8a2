"""Host-side mirror of ``pinnrl.pdes`` for the five hot-path PDEs.

Same class names, constructor (``PDEConfig``), attributes and method signatures as the reference
(pinnrl/pdes/pde_base.py and the five equation modules); ``compute_residual`` / ``compute_loss`` /
residual-based sampling run through libpinnk (see functional.py).  Collocation samplers and the
BC/IC target closures are host-side torch code kept semantically identical to the reference so that
seeded runs see the same points and targets.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, List, Optional, Tuple, Union

import numpy as np
import torch
import torch.nn as nn

from . import functional as F
from . import rl as _rl


@dataclass
class PDEConfig:
    """pinnrl/pdes/pde_base.py:22-48."""
    name: str
    domain: Union[Tuple[float, float], List[Tuple[float, float]]]
    time_domain: Tuple[float, float]
    parameters: Dict[str, float]
    boundary_conditions: Dict[str, Dict[str, Any]]
    initial_condition: Dict[str, Any]
    exact_solution: Dict[str, Any]
    dimension: int = 1
    input_dim: Optional[int] = None
    output_dim: Optional[int] = None
    architecture: Optional[str] = None
    device: Optional[torch.device] = None
    training: Optional[Dict[str, Any]] = None
    trainable_parameters: List[str] = field(default_factory=list)
    parameter_initial_guesses: Dict[str, float] = field(default_factory=dict)
    observation_data: Optional[Dict[str, Any]] = None


class PDEBase:
    """pinnrl/pdes/pde_base.py:51-1561, the parts on or next to the hot path."""

    compat = "reference"   # "reference": operators exactly as the reference computes them;
                           # "math": the intended operators

    def __init__(self, config: PDEConfig, rl_agent=None):
        self.config = config
        self.rl_agent = rl_agent
        dom = config.domain
        if isinstance(dom, list) and len(dom) > 0:
            if isinstance(dom[0], (list, tuple)):
                dom = [(float(d[0]), float(d[1])) for d in dom]
            else:
                dom = [(float(dom[0]), float(dom[1]))]
        elif not isinstance(dom, list):
            dom = [(0.0, 1.0)]
        self.domain = dom
        self.config.domain = dom
        td = getattr(config, "time_domain", [0.0, 1.0])
        self.time_domain = tuple(td) if isinstance(td, list) else td
        dev = getattr(config, "device", None)
        self.device = dev if isinstance(dev, torch.device) else torch.device(str(dev)) if dev is not None else torch.device("cpu")
        self.config.device = self.device
        self.dimension = config.dimension
        if getattr(config, "parameters", None) is None:
            config.parameters = {}
        if list(getattr(config, "trainable_parameters", []) or []):
            raise F.UnsupportedPDE("trainable PDE parameters (inverse mode) are outside the B200 hot path")
        self._trainable_params = nn.ParameterDict()
        self.observation_data = self._load_observation_data(getattr(config, "observation_data", None))
        self._setup_boundary_conditions()
        self.validation_points = None
        self.collocation_history: List[np.ndarray] = []
        if self.config.input_dim is None:
            self.config.input_dim = self.dimension + 1
        if self.config.output_dim is None:
            self.config.output_dim = 1

    # ---- configuration helpers (pde_base.py:246-357)
    def get_parameter(self, name: str, default=None, required: bool = False):
        params = getattr(self.config, "parameters", None)
        if params is None:
            if required:
                raise ValueError(f"Required parameter '{name}' not found in config")
            return default
        value = params.get(name, default)
        if value is None and required:
            raise ValueError(f"Required parameter '{name}' not found in config")
        return value

    def _tr(self, key, default):
        tr = getattr(self.config, "training", None)
        if tr is None:
            return default
        return tr.get(key, default) if isinstance(tr, dict) else getattr(tr, key, default)

    def _loss_function_name(self) -> str:
        return self._tr("loss_function", "mse")

    def _huber_delta(self) -> float:
        return float(self._tr("huber_delta", 1.0))

    def _training_mode(self) -> str:
        return str(self._tr("mode", "forward"))

    def _data_loss_weight(self, default: float = 1.0) -> float:
        try:
            lw = self.config.training.loss_weights
            return float(lw.get("data", default)) if isinstance(lw, dict) else float(getattr(lw, "data", default))
        except AttributeError:
            return default

    def _apply_loss_fn(self, error: torch.Tensor) -> torch.Tensor:
        name = self._loss_function_name()
        if name == "mae":
            return torch.mean(torch.abs(error))
        if name == "huber":
            return torch.nn.functional.huber_loss(error, torch.zeros_like(error), reduction="mean",
                                                  delta=self._huber_delta())
        return torch.mean(error ** 2)

    def trainable_parameters_iter(self):
        return iter(())

    def _load_observation_data(self, obs):
        if not obs:
            return None
        if all(k in obs for k in ("x", "t", "u")):
            out = {}
            for k in ("x", "t", "u"):
                v = obs[k]
                v = v if isinstance(v, torch.Tensor) else torch.tensor(np.asarray(v, dtype=np.float32))
                out[k] = (v.reshape(-1, 1) if v.dim() == 1 else v).to(self.device)
            return out
        raise F.UnsupportedPDE("only in-memory observation data {x,t,u} is supported")

    # ---- BC / IC target closures (pde_base.py:474-575)
    def _setup_boundary_conditions(self):
        self.boundary_conditions: Dict[str, Callable] = {}
        bcs = getattr(self.config, "boundary_conditions", None)
        if bcs:
            for kind, params in bcs.items():
                self.boundary_conditions[kind] = self._create_boundary_condition(kind, params)
        if "initial" not in self.boundary_conditions and hasattr(self.config, "initial_condition"):
            self.boundary_conditions["initial"] = self._create_boundary_condition("initial", self.config.initial_condition)

    def _create_boundary_condition(self, bc_type: str, params: Dict[str, Any]) -> Callable:
        if bc_type in ("left", "right"):
            bc_type = "dirichlet"
        if bc_type in ("dirichlet", "neumann"):
            value = params.get("value", 0.0)
            return lambda x, t: torch.full_like(x[:, 0:1], value)
        if bc_type == "periodic":
            if self.dimension == 1:
                return lambda x, t: torch.sin(2 * torch.pi * x[:, 0:1])
            return lambda x, t: torch.sin(2 * torch.pi * torch.sum(x, dim=1, keepdim=True))
        if bc_type == "initial":
            kind = params.get("type", "sine")
            if kind in ("sine", "sin_exp_decay"):
                a, f = params.get("amplitude", 1.0), params.get("frequency", 1.0)
                return lambda x, t: a * torch.sin(f * torch.pi * x[:, 0:1])
            if kind == "tanh":
                e = params.get("epsilon", 0.1)
                return lambda x, t: torch.tanh(x[:, 0:1] / e)
            if kind == "gaussian":
                m, s = params.get("mean", 0.0), params.get("std", 0.1)
                return lambda x, t: torch.exp(-((x[:, 0:1] - m) ** 2) / (2 * s ** 2))
            if kind == "fixed":
                v = params.get("value", 0.0)
                return lambda x, t: torch.full_like(x[:, 0:1], v)
            if kind == "random":
                amp = params.get("amplitude", 0.1)
                return lambda x, t: amp * (2 * torch.rand_like(x[:, 0:1]) - 1)
            print(f"Warning: Unrecognized initial condition type '{kind}'. Defaulting to zero.")
            return lambda x, t: torch.zeros_like(x[:, 0:1])
        print(f"Warning: Unsupported boundary condition type '{bc_type}'. Defaulting to zero.")
        return lambda x, t: torch.zeros_like(x[:, 0:1])

    # ---- the hot path
    def compute_residual(self, model: nn.Module, x: torch.Tensor, t: torch.Tensor) -> torch.Tensor:
        return F.compute_residual(self, model, x, t)

    def compute_loss(self, model: nn.Module, x: torch.Tensor, t: torch.Tensor) -> Dict[str, torch.Tensor]:
        return F.compute_loss(self, model, x, t)

    def compute_derivatives(self, model, x, t, temporal_derivatives=None, spatial_derivatives=None):
        """pde_base.py:590-794 -- the plugin-PDE building block (CONTRIBUTING.md:152-244): same keys, one jet pass."""
        return F.compute_derivatives(self, model, x, t, temporal_derivatives, spatial_derivatives)

    def score_residual(self, model, x, t, want_abs=True):
        return F.score_residual(self, model, x, t, want_abs)

    def exact_solution(self, x, t):
        raise NotImplementedError("Subclasses must implement exact_solution")

    # ---- samplers (pde_base.py:806-1084)
    def _sample_uniform(self, num_points: int):
        if self.dimension == 1:
            n_side = int(np.sqrt(num_points))
            x = torch.linspace(self.domain[0][0], self.domain[0][1], n_side, device=self.device).reshape(-1, 1)
            t = torch.linspace(self.time_domain[0], self.time_domain[1], n_side, device=self.device).reshape(-1, 1)
            if x.is_cuda and n_side >= 1:
                # meshgrid + jitter + clamp in ONE libpinnk launch (pinnk_jittered_grid): same linspace vectors, same two
                # randn draws in the same order as the reference, bit-identical points, 1 launch instead of ~10
                return jittered_grid_device(x.reshape(-1), t.reshape(-1), self.domain[0], self.time_domain)
            X, T = torch.meshgrid(x.squeeze(), t.squeeze(), indexing="ij")
            x, t = X.reshape(-1, 1), T.reshape(-1, 1)
            x = x + torch.randn_like(x) * (self.domain[0][1] - self.domain[0][0]) * 0.01
            t = t + torch.randn_like(t) * (self.time_domain[1] - self.time_domain[0]) * 0.01
            x = torch.clamp(x, self.domain[0][0], self.domain[0][1])
            t = torch.clamp(t, self.time_domain[0], self.time_domain[1])
            return x, t
        ppd = max(2, int(num_points ** (1 / (self.dimension + 1))) + 1)
        grids = [torch.linspace(self.domain[d][0], self.domain[d][1], ppd) for d in range(self.dimension)]
        grids.append(torch.linspace(self.time_domain[0], self.time_domain[1], ppd))
        pts = torch.stack([g.reshape(-1) for g in torch.meshgrid(*grids, indexing="ij")], dim=1)
        if len(pts) > num_points:
            pts = pts[torch.randperm(len(pts))[:num_points]]
        elif len(pts) < num_points:
            pts = torch.cat([pts, pts[torch.randint(0, len(pts), (num_points - len(pts),))]], dim=0)
        pts = pts + torch.randn_like(pts) * 0.01
        for d in range(self.dimension):
            pts[:, d] = torch.clamp(pts[:, d], self.domain[d][0], self.domain[d][1])
        pts[:, -1] = torch.clamp(pts[:, -1], self.time_domain[0], self.time_domain[1])
        return pts[:, :self.dimension], pts[:, -1].reshape(-1, 1)

    def _sample_stratified(self, num_points: int):
        dims = self.dimension + 1
        lows = [self.domain[d][0] for d in range(self.dimension)] + [self.time_domain[0]]
        ups = [self.domain[d][1] for d in range(self.dimension)] + [self.time_domain[1]]
        s = torch.zeros(num_points, dims, device=self.device)
        for d in range(dims):
            bin_size = (ups[d] - lows[d]) / num_points
            off = torch.rand(num_points, device=self.device)
            idx = torch.arange(num_points, dtype=torch.float32, device=self.device)
            s[:, d] = lows[d] + (idx + off) * bin_size
            s[:, d] = s[torch.randperm(num_points, device=self.device), d]
        return s[:, :self.dimension], s[:, -1].reshape(-1, 1)

    def _sample_residual_based(self, num_points: int, model: Optional[nn.Module] = None):
        """RAR (pde_base.py:895-935): candidates are scored forward-only by the CUDA path; the draw is a
        two-level multinomial (block, then point) so pools beyond torch.multinomial's 2^24 limit work."""
        if model is None:
            return self._sample_uniform(num_points)
        x_pool, t_pool = self._sample_uniform(num_points * 4)
        x_pool, t_pool = x_pool.to(self.device), t_pool.to(self.device)
        mag, _ = F.score_residual(self, model, x_pool, t_pool, want_abs=True)
        # probs = |r| + 1e-8, multinomial with replacement (pde_base.py:924-931).  Up to torch.multinomial's 2^24 categories
        # the draw IS torch.multinomial on the reference's normalised probabilities: the same generator calls, hence the same
        # random stream as the reference for the rest of the run (trajectory parity).  Beyond that, where the reference fails
        # (SURVEY F7), or with PINNK_DEVICE_SAMPLER=1: inverse-CDF draw on the device straight from the scoring kernel's |r|
        # (no normalised copy, no host round trip).
        if use_device_sampler(mag.numel()):
            sel = weighted_sample_device(mag, 1e-8, num_points)
        else:
            probs = mag.reshape(-1) + 1e-8
            sel = torch.multinomial(probs / probs.sum(), num_points, replacement=True)
        return x_pool[sel].detach(), t_pool[sel].detach()

    def generate_collocation_points(self, num_points: int, strategy: str = "uniform", **kwargs):
        if strategy == "uniform":
            x, t = self._sample_uniform(num_points)
        elif strategy == "stratified":
            x, t = self._sample_stratified(num_points)
        elif strategy == "residual_based":
            x, t = self._sample_residual_based(num_points, kwargs.get("model", None))
        elif strategy == "adaptive":
            if self.rl_agent is None:
                return self.generate_collocation_points(num_points, strategy="uniform")
            x, t = self._sample_adaptive(num_points)
        else:
            raise ValueError(f"Unknown sampling strategy: {strategy}")
        return x.to(self.device), t.to(self.device)

    def _sample_adaptive(self, num_points: int):
        """RL-agent-driven sampling over a <=100-per-axis candidate grid (pde_base.py:961-1072)."""
        gs = min(100, max(10, int(np.sqrt(num_points))))
        axes = [torch.linspace(self.domain[d][0], self.domain[d][1], gs, device=self.device) for d in range(self.dimension)]
        axes.append(torch.linspace(self.time_domain[0], self.time_domain[1], gs, device=self.device))
        pts = torch.stack([g.flatten() for g in torch.meshgrid(*axes, indexing="ij")], dim=1)
        with torch.no_grad():
            probs = _rl.grid_scores(self.rl_agent, pts)      # DQNNetwork-shaped policy nets: one libpinnk launch
        sel = multinomial_large(probs.flatten(), min(num_points, len(pts)))
        chosen = pts[sel]
        if len(chosen) < num_points:
            extra = torch.randint(0, len(chosen), (num_points - len(chosen),), device=self.device)
            chosen = torch.cat([chosen, chosen[extra]], dim=0)
        scale = min(0.01, min((self.domain[d][1] - self.domain[d][0]) / gs for d in range(self.dimension)),
                    (self.time_domain[1] - self.time_domain[0]) / gs)
        chosen = chosen + torch.randn_like(chosen) * scale
        for d in range(self.dimension):
            chosen[:, d] = torch.clamp(chosen[:, d], self.domain[d][0], self.domain[d][1])
        chosen[:, -1] = torch.clamp(chosen[:, -1], self.time_domain[0], self.time_domain[1])
        self.collocation_history.append(chosen.cpu().numpy())
        if len(self.collocation_history) > 1 and hasattr(self.rl_agent, "update_epsilon"):
            self.rl_agent.update_epsilon(len(self.collocation_history))
        x = chosen[:, 0].reshape(-1, 1) if self.dimension == 1 else chosen[:, :self.dimension]
        return x, chosen[:, -1].reshape(-1, 1)

    def update_sampling_strategy(self, x, t, residual):
        reward = torch.mean(torch.abs(residual))
        self.rl_agent.update(torch.cat([x, t], dim=1), reward)

    def validate(self, model, num_points: int = 1000):
        x, t = self.generate_collocation_points(num_points)
        err = torch.abs(model(torch.cat([x, t], dim=1)) - self.exact_solution(x, t))
        return {"l2_error": torch.mean(err ** 2).item(), "max_error": torch.max(err).item(),
                "mean_error": torch.mean(err).item()}


def use_device_sampler(n: int) -> bool:
    """torch.multinomial serves up to 2^24 categories with the reference's own generator calls; pinnk_sample_weighted takes
    over beyond that, or everywhere with PINNK_DEVICE_SAMPLER=1 (a different, equally distributed draw)."""
    return n > (1 << 24) or os.environ.get("PINNK_DEVICE_SAMPLER", "0") == "1"


def weighted_sample_device(weights: torch.Tensor, eps: float, num_samples: int) -> torch.Tensor:
    """``torch.multinomial((w + eps) / sum, num_samples, replacement=True)`` on the device for ANY number of categories
    (pinnk_sample_weighted: fp64 block sums, one scan, one warp per sample; torch.multinomial stops at 2^24).
    The uniforms come from torch's CUDA generator (one ``torch.rand`` call of ``num_samples`` doubles).  int64 indices."""
    import ctypes as C
    from . import _lib as L
    w = weights.detach().reshape(-1)
    if not (w.is_cuda and w.dtype == torch.float32):
        raise L.PinnkError("weighted_sample_device needs float32 CUDA weights (there is no CPU path)")
    w = w.contiguous()
    n, m = w.numel(), int(num_samples)
    idx = torch.empty(m, dtype=torch.int64, device=w.device)
    if n == 0 or m == 0:
        if m:
            raise ValueError("cannot sample from an empty candidate set")
        return idx
    lib = L.load()
    ws_n = int(lib.pinnk_sample_workspace_doubles(n))
    ws = torch.empty(ws_n, dtype=torch.float64, device=w.device)
    u = torch.rand(m, dtype=torch.float64, device=w.device)
    L.check(lib.pinnk_sample_weighted(w.data_ptr(), n, float(eps), u.data_ptr(), m, idx.data_ptr(), ws.data_ptr(), ws_n,
                                      C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_sample_weighted")
    return idx


def jittered_grid_device(xs: torch.Tensor, ts: torch.Tensor, x_dom, t_dom):
    """The jittered ``len(xs) x len(ts)`` grid of ``_sample_uniform`` (pde_base.py:806-829) in one launch; draws the jitter
    with the reference's two ``randn`` calls ([n, 1] for x, then [n, 1] for t)."""
    import ctypes as C
    from . import _lib as L
    n_side = xs.numel()
    n = n_side * n_side
    nx = torch.randn(n, 1, dtype=torch.float32, device=xs.device)
    nt = torch.randn(n, 1, dtype=torch.float32, device=xs.device)
    x = torch.empty(n, 1, dtype=torch.float32, device=xs.device)
    t = torch.empty(n, 1, dtype=torch.float32, device=xs.device)
    if n:
        L.check(L.load().pinnk_jittered_grid(xs.contiguous().data_ptr(), ts.contiguous().data_ptr(), n_side, nx.data_ptr(),
                                             nt.data_ptr(), (x_dom[1] - x_dom[0]) * 0.01, (t_dom[1] - t_dom[0]) * 0.01,
                                             float(x_dom[0]), float(x_dom[1]), float(t_dom[0]), float(t_dom[1]),
                                             x.data_ptr(), t.data_ptr(),
                                             C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_jittered_grid")
    return x, t


def multinomial_large(weights: torch.Tensor, num_samples: int, block: int = 1 << 20) -> torch.Tensor:
    """``torch.multinomial(w / w.sum(), num_samples, replacement=True)`` without the 2^24 category limit
    (SURVEY F7): draw a block proportionally to block mass, then a point inside the block."""
    w = weights.reshape(-1).to(torch.float32)
    n = w.numel()
    if w.is_cuda and use_device_sampler(n):
        return weighted_sample_device(w, 0.0, num_samples)
    if n <= (1 << 24):
        return torch.multinomial(w / w.sum(), num_samples, replacement=True)
    nb = -(-n // block)
    pad = torch.zeros(nb * block, dtype=w.dtype, device=w.device)
    pad[:n] = w
    pad = pad.view(nb, block)
    mass = pad.sum(dim=1, dtype=torch.float64)
    which = torch.multinomial((mass / mass.sum()).to(torch.float32), num_samples, replacement=True)
    counts = torch.bincount(which, minlength=nb)
    out = []
    for b in torch.nonzero(counts).flatten().tolist():
        inner = torch.multinomial(pad[b] / pad[b].sum(), int(counts[b]), replacement=True)
        out.append(inner + b * block)
    sel = torch.cat(out)
    return sel[torch.randperm(sel.numel(), device=sel.device)]


# ---------------------------------------------------------------------------------- the five PDEs
class HeatEquation(PDEBase):
    """heat_equation.py: residual u_t - alpha*u_x as written (compat='reference') or
    u_t - alpha*u_xx (compat='math'); periodic-BC loss of heat_equation.py:375-623."""

    @property
    def alpha(self):
        return self.get_parameter("alpha", required=True)

    def _calculate_decay_rate(self, k: float) -> float:
        L_ = self.config.domain[0][1] - self.config.domain[0][0]
        return self.alpha * (2 * torch.pi * k / L_) ** 2

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "sine")
            A, k = params.get("amplitude", 1.0), params.get("frequency", 2.0)
            L_ = self.config.domain[0][1] - self.config.domain[0][0]
            wave = 2 * torch.pi * k / L_
            if kind == "sin_exp_decay":
                decay = self._calculate_decay_rate(k)
                if self.dimension == 1:
                    return lambda x, t: A * torch.sin(wave * x) * torch.exp(-decay * t)

                def ic(x, t):
                    sol = torch.ones_like(x[:, 0:1])
                    for d in range(self.dimension):
                        Ld = self.config.domain[d][1] - self.config.domain[d][0]
                        sol = sol * torch.sin(2 * torch.pi * k / Ld * x[:, d:d + 1])
                    return A * sol * torch.exp(-decay * t)
                return ic
            if kind == "sine":
                if self.dimension == 1:
                    return lambda x, t: A * torch.sin(wave * x)
                return lambda x, t: A * torch.prod(torch.sin(wave * x), dim=1, keepdim=True)
            if kind == "sine_2d":
                kx, ky = params.get("frequency_x", 2.0), params.get("frequency_y", 2.0)
                return lambda x, t: A * torch.sin(kx * torch.pi * x[:, 0:1]) * torch.sin(ky * torch.pi * x[:, 1:2])
            return super()._create_boundary_condition(bc_type, params)
        ex = getattr(self.config, "exact_solution", None) or {}
        if bc_type == "dirichlet" and ex.get("type") == "sin_exp_decay":
            A, k = ex.get("amplitude", 1.0), ex.get("frequency", 2.0)
            wave = 2 * torch.pi * k / (self.config.domain[0][1] - self.config.domain[0][0])
            decay = self._calculate_decay_rate(k)
            return lambda x, t: A * torch.sin(wave * x) * torch.exp(-decay * t)
        return super()._create_boundary_condition(bc_type, params)

    def exact_solution(self, x, t):
        ex = self.config.exact_solution or {}
        A, k = ex.get("amplitude", 1.0), ex.get("frequency", 2.0)
        L_ = self.config.domain[0][1] - self.config.domain[0][0]
        wave = 2 * torch.pi * k / L_
        return A * torch.sin(wave * x[:, 0:1]) * torch.exp(-self.alpha * wave ** 2 * t)


class BurgersEquation(PDEBase):
    """burgers_equation.py: r = u_t + u u_x - nu u_xx (key ``nu``, default 0.01)."""

    @property
    def nu(self):
        return self.get_parameter("nu", default=0.01)

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "sine")
            if kind == "sine":
                A, k = params.get("amplitude", -1.0), params.get("frequency", 1.0)
                if self.dimension == 1:
                    return lambda x, t: A * torch.sin(k * torch.pi * x)
                return lambda x, t: A * torch.prod(torch.sin(k * torch.pi * x), dim=1, keepdim=True)
            if kind == "tanh":
                e = params.get("epsilon", 0.1)
                if self.dimension == 1:
                    return lambda x, t: torch.tanh((x - 0.5) / e)
                return lambda x, t: torch.prod(torch.tanh((x - 0.5) / e), dim=1, keepdim=True)
            raise ValueError(f"Unsupported initial condition type: {kind}")
        return super()._create_boundary_condition(bc_type, params)


class KdVEquation(PDEBase):
    """kdv_equation.py: r = u_t + 6 u u_x + u_xxx."""

    @property
    def speed(self):
        return self.get_parameter("speed", default=1.0)

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "soliton")
            if kind != "soliton":
                raise ValueError(f"Unsupported initial condition type: {kind}")
            c = torch.tensor(params.get("speed", self.speed), dtype=torch.float32, device=self.device)
            if self.dimension == 1:
                return lambda x, t: 2 * c * (1 / torch.cosh(torch.sqrt(c) * x)) ** 2
            return lambda x, t: 2 * c * (1 / torch.cosh(torch.sqrt(c) * torch.sum(x, dim=1, keepdim=True))) ** 2
        return super()._create_boundary_condition(bc_type, params)


class _PhaseField(PDEBase):
    @property
    def epsilon(self):
        return self.get_parameter("epsilon", default=0.1)

    _allow_random_ic = False

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "tanh")
            if kind == "tanh":
                if self.dimension == 1:
                    return lambda x, t: torch.tanh(x / (2 * self.epsilon))
                return lambda x, t: torch.tanh(torch.sum(x, dim=1, keepdim=True) / (2 * self.epsilon))
            if kind == "random" and self._allow_random_ic:
                amp = params.get("amplitude", 0.1)
                return lambda x, t: amp * (2 * torch.rand_like(x[:, 0:1]) - 1)
            raise ValueError(f"Unsupported initial condition type: {kind}")
        return super()._create_boundary_condition(bc_type, params)


class AllenCahnEquation(_PhaseField):
    """allen_cahn.py: r = u_t - eps^2 u_xx - u + u^3."""


class CahnHilliardEquation(_PhaseField):
    """cahn_hilliard.py: r = u_t - Lap(-eps^2 Lap u + clamp(u)^3 - clamp(u))."""
    _allow_random_ic = True


class _SineIC(PDEBase):
    """Shared initial-condition table of the wave and convection equations (wave_equation.py:138-169,
    convection_equation.py:97-119): ``A sin(k pi x)``, defaults A = 1, k = 2."""
    _ic_types = ("sine",)

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "sine")
            if kind in self._ic_types:
                A, k = params.get("amplitude", 1.0), params.get("frequency", 2.0)
                if self.dimension == 1:
                    return lambda x, t: A * torch.sin(k * torch.pi * x)
                return lambda x, t: A * torch.sin(k * torch.pi * torch.sum(x, dim=1, keepdim=True))
            raise ValueError(f"Unsupported initial condition type: {kind}")
        return super()._create_boundary_condition(bc_type, params)


class WaveEquation(_SineIC):
    """wave_equation.py: r = u_tt - c^2 u_xx (key ``c``, default 1.0); second-order jets in t as well as x."""

    @property
    def c(self):
        return self.get_parameter("c", default=1.0)

    def exact_solution(self, x, t):
        if self.dimension == 1:
            return torch.sin(2 * torch.pi * (x - self.c * t))
        sol = torch.ones_like(x[:, 0:1])
        for d in range(self.dimension):
            sol = sol * torch.sin(2 * torch.pi * (x[:, d:d + 1] - self.c * t))
        return sol


class ConvectionEquation(_SineIC):
    """convection_equation.py: r = u_t + v u_x (key ``velocity``: scalar or per-dimension list, default 1.0)."""
    _ic_types = ("sine", "sin")

    @property
    def velocity(self):
        v = self.get_parameter("velocity", default=1.0)
        return [v] * self.dimension if isinstance(v, (int, float)) else v

    def exact_solution(self, x, t):
        if self.dimension == 1:
            return torch.sin(2 * torch.pi * (x - self.velocity[0] * t))
        sol = torch.ones_like(x[:, 0:1])
        for d in range(self.dimension):
            sol = sol * torch.sin(2 * torch.pi * (x[:, d:d + 1] - self.velocity[d] * t))
        return sol


class BlackScholesEquation(PDEBase):
    """black_scholes.py: r = V_t + sigma^2/2 S^2 V_SS + r S V_S - r V (keys ``sigma`` 0.2, ``r`` 0.05); the residual's
    coefficients depend on the point's coordinate S."""

    @property
    def sigma(self):
        return self.get_parameter("sigma", default=0.2)

    @property
    def r(self):
        return self.get_parameter("r", default=0.05)

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":
            kind = params.get("type", "call_option")
            if kind in ("call_option", "option"):
                K = params.get("strike_price", params.get("strike", 1.0))
                if self.dimension == 1:
                    return lambda x, t: torch.maximum(x - K, torch.zeros_like(x))
                return lambda x, t: torch.maximum(torch.sum(x, dim=1, keepdim=True) - K, torch.zeros_like(x[:, 0:1]))
            raise ValueError(f"Unsupported initial condition type: {kind}")
        return super()._create_boundary_condition(bc_type, params)


class PendulumEquation(PDEBase):
    """pendulum_equation.py: r = u_tt + (g/L) sin u (keys ``g`` 9.81, ``L`` 1.0): an ODE in t, second-order time jets only."""

    @property
    def g(self):
        return self.get_parameter("g", default=9.81)

    @property
    def L(self):
        return self.get_parameter("L", default=1.0)

    def _create_boundary_condition(self, bc_type, params):
        if bc_type == "initial":                                  # pendulum_equation.py:125-156
            kind = params.get("type", "small_angle")
            if kind == "small_angle":
                th0 = params.get("initial_angle", 0.1)
                return lambda x, t: torch.full_like(x, th0)
            if kind == "sine":
                a, f = params.get("amplitude", 1.0), params.get("frequency", 1.0)
                return lambda x, t: a * torch.sin(f * x)
            if kind == "gaussian":
                a, c, sg = params.get("amplitude", 1.0), params.get("center", 0.0), params.get("sigma", 0.1)
                return lambda x, t: a * torch.exp(-((x - c) ** 2) / (2 * sg ** 2))
            raise ValueError(f"Unknown initial condition type: {kind}")
        return super()._create_boundary_condition(bc_type, params)


_FACTORY = {"heat": HeatEquation, "burgers": BurgersEquation, "kdv": KdVEquation,
            "allen_cahn": AllenCahnEquation, "cahn_hilliard": CahnHilliardEquation,
            "wave": WaveEquation, "convection": ConvectionEquation,
            "black_scholes": BlackScholesEquation, "pendulum": PendulumEquation}


def create_pde(name: str, config: PDEConfig) -> PDEBase:
    key = name.lower().replace(" ", "_").replace("-", "_").replace("_equation", "")
    if key not in _FACTORY:
        raise ValueError(f"PDE '{name}' is outside the B200 hot path (supported: {sorted(_FACTORY)})")
    return _FACTORY[key](config=config)

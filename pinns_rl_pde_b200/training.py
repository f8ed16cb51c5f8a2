"""Host-side mirror of the inner training step of ``pinnrl.training.trainer.PDETrainer``
(trainer.py:292-332 optimiser/scheduler set-up, :539-698 the step, :373-389 the L-BFGS closure step, :140-162 validation
loss, :171-279 live snapshot, :580-634 adaptive re-weighting).  Plots, dashboards and metadata files are outside the hot
path and not mirrored; for those, patch the reference (``patch_reference``) and use its own trainer.

``fused=True`` runs the whole Adam step inside libpinnk; ``graph=True`` replays it as a CUDA graph per batch shape.
With ``world_size > 1`` (one process per GPU, torch.distributed/NCCL) collocation rows are sharded
across ranks and the flat gradient + loss sums are all-reduced once per step (parallel.py).
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import functional as F
from . import parallel


@dataclass
class LBFGSConfig:
    """config/__init__.py:47-63: settings handed to torch.optim.LBFGS."""
    history_size: int = 50
    max_iter: int = 20
    line_search_fn: Optional[str] = "strong_wolfe"
    tolerance_grad: float = 1e-7
    tolerance_change: float = 1e-9


@dataclass
class AdaptiveWeightsConfig:
    """config/__init__.py:67-87."""
    enabled: bool = False
    strategy: str = "rbw"              # "lrw" | "rbw"
    alpha: float = 0.9
    eps: float = 1e-5
    initial_weights: Optional[List[float]] = None

    def __post_init__(self):
        if self.initial_weights is None:
            self.initial_weights = [0.5, 0.3, 0.2]


class AdaptiveLossWeights:
    """components/adaptive_weights.py:6-134 on device 3-vectors.  ``update(losses=...)`` (RBW: weights follow the running
    losses, smoothed with the previous weights) or ``update(gradients=...)`` (LRW: weights inversely proportional to the
    running per-component gradient norms); the first call returns the initial weights and seeds the running average."""

    def __init__(self, strategy="rbw", alpha=0.9, eps=1e-5, initial_weights=None):
        self.strategy, self.alpha, self.eps = str(strategy).lower(), float(alpha), float(eps)
        if self.strategy not in ("lrw", "rbw"):
            raise ValueError(f"unknown adaptive-weights strategy {strategy!r}")
        self.initial_weights = None if initial_weights is None else torch.tensor(initial_weights)
        self.weights = self.running = self.prev_weights = None

    def update(self, losses: Optional[torch.Tensor] = None, gradients: Optional[torch.Tensor] = None) -> torch.Tensor:
        v = gradients if self.strategy == "lrw" else losses
        if v is None:
            raise ValueError(f"Invalid combination of strategy ({self.strategy}) and inputs")
        if self.running is None:
            self.running = v
            self.weights = (self.initial_weights.to(v.device) if self.initial_weights is not None else torch.ones_like(v))
            return self.weights
        self.running = self.alpha * self.running + (1 - self.alpha) * v
        eps = torch.tensor(self.eps, device=v.device)
        if self.strategy == "lrw":
            inv = 1.0 / (self.running + eps)
            self.weights = inv / torch.sum(inv)
            return self.weights
        self.weights = self.running / (self.running.sum() + eps)
        if self.prev_weights is not None:
            self.weights = self.alpha * self.prev_weights + (1 - self.alpha) * self.weights
        self.prev_weights = self.weights.clone()
        return self.weights

    def get_weights(self) -> torch.Tensor:
        if self.weights is not None:
            return self.weights
        return self.initial_weights if self.initial_weights is not None else torch.ones(3) / 3.0


@dataclass
class TrainingConfig:
    """The fields of pinnrl.config.TrainingConfig the step reads (config/__init__.py:90-169)."""
    num_epochs: int = 100
    batch_size: int = 2048
    num_collocation_points: int = 5000
    num_boundary_points: int = 500
    num_initial_points: int = 500
    learning_rate: float = 1e-3
    weight_decay: float = 0.0
    gradient_clipping: float = 1.0
    collocation_distribution: str = "uniform"
    loss_weights: Optional[Dict[str, float]] = None
    scheduler: str = "cosine"          # "cosine" | "none"
    min_lr: float = 1e-6
    mode: str = "forward"
    loss_function: str = "mse"
    huber_delta: float = 1.0
    optimizer: str = "adam"            # "adam" | "lbfgs" (config/__init__.py:123; the adam_lbfgs hand-over: switch_to_lbfgs())
    lbfgs: Optional[LBFGSConfig] = None
    adaptive_weights: Optional[AdaptiveWeightsConfig] = None

    def __post_init__(self):
        if self.lbfgs is None:
            self.lbfgs = LBFGSConfig()
        if self.adaptive_weights is None:
            self.adaptive_weights = AdaptiveWeightsConfig()
        if self.optimizer not in ("adam", "lbfgs"):
            raise ValueError(f"unknown optimizer {self.optimizer!r}: choose 'adam' or 'lbfgs'")
        if self.loss_weights is None:
            self.loss_weights = {"residual": 1.0, "boundary": 1.0, "initial": 1.0}
        self.loss_weights.setdefault("data", 1.0)

    def get(self, key, default=None):
        return getattr(self, key, default)

    def __getitem__(self, key):
        return getattr(self, key)


class FusedAdam:
    """clip_grad_norm_ + Adam(L2 weight decay) on the flat gradient in two libpinnk launches
    (trainer.py:690-694,292-297); state and update rule of ``torch.optim.Adam`` (bias corrections in double like
    torch's Python-float ones; the moments and the update in fp32)."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0,
                 capturable: bool = False):
        import ctypes as C
        from . import _lib
        self.params = [p for p in params if p.requires_grad]
        if not 1 <= len(self.params) <= 64:
            raise ValueError("FusedAdam handles 1..64 parameter tensors")
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.exp_avg = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(n, dtype=torch.float32, device=dev)
        self._scratch = torch.zeros(1, dtype=torch.float64, device=dev)
        self.lr, self.betas, self.eps, self.weight_decay, self.max_norm = lr, betas, eps, weight_decay, max_norm
        self.step_count = 0
        self._C, self._lib = C, _lib
        self._ptrs = (C.c_void_p * len(self.params))()
        self._numels = (C.c_int64 * len(self.params))(*[p.numel() for p in self.params])
        self.param_groups = [{"lr": lr}]           # so torch LR schedulers can drive it
        # capturable: step count and lr live in device memory ([step, lr] doubles) so a captured CUDA graph of the step
        # stays valid from one call to the next (pinnk_adam_step_dev)
        self.capturable = bool(capturable)
        self._dyn = torch.tensor([0.0, lr], dtype=torch.float64, device=dev) if self.capturable else None
        self._dyn_lr = float(lr)

    def step(self, flat_grad: torch.Tensor):
        C, L = self._C, self._lib
        n = self.exp_avg.numel()
        if not (flat_grad.is_cuda and flat_grad.dtype == torch.float32 and flat_grad.is_contiguous()
                and flat_grad.numel() == n and flat_grad.device == self.exp_avg.device):
            # (requires_grad flags changed after construction would otherwise make the kernel read out of bounds)
            raise L.PinnkError(f"FusedAdam.step: flat gradient must be a contiguous float32 CUDA tensor of {n} elements "
                               f"on {self.exp_avg.device} (got {tuple(flat_grad.shape)}, {flat_grad.dtype}, {flat_grad.device})")
        self.step_count += 1
        for i, p in enumerate(self.params):
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                raise L.PinnkError("FusedAdam needs contiguous float32 CUDA parameters")
            self._ptrs[i] = p.data_ptr()
        lr = float(self.param_groups[0]["lr"])
        if self.capturable:
            capturing = torch.cuda.is_current_stream_capturing()
            if lr != self._dyn_lr:
                if capturing:
                    raise L.PinnkError("FusedAdam: set the learning rate outside the captured region (sync_lr())")
                self.sync_lr()
            self._dyn[0:1].add_(1.0)               # on the stream: replayed with the graph
            L.check(L.load().pinnk_adam_step_dev(C.cast(self._ptrs, C.c_void_p), C.cast(self._numels, C.c_void_p),
                                                 len(self.params), flat_grad.data_ptr(), self.exp_avg.data_ptr(),
                                                 self.exp_avg_sq.data_ptr(), self._scratch.data_ptr(), self._dyn.data_ptr(),
                                                 self.betas[0], self.betas[1], self.eps, self.weight_decay, self.max_norm,
                                                 C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_adam_step_dev")
            return
        L.check(L.load().pinnk_adam_step(C.cast(self._ptrs, C.c_void_p), C.cast(self._numels, C.c_void_p), len(self.params),
                                         flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                         self._scratch.data_ptr(), self.step_count, lr, self.betas[0], self.betas[1],
                                         self.eps, self.weight_decay, self.max_norm,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_adam_step")


    def sync_lr(self):
        """Push ``param_groups[0]['lr']`` to the device copy a captured step reads."""
        lr = float(self.param_groups[0]["lr"])
        if self.capturable and lr != self._dyn_lr:
            self._dyn[1:2].fill_(lr)
            self._dyn_lr = lr


class PDETrainer:
    def __init__(self, model: nn.Module, pde, optimizer_config=None, config=None, device=None, rl_agent=None,
                 fused: bool = False, graph: bool = False):
        """``fused``: whole step inside libpinnk (loss + weighted gradient in one pass per row set, clip + Adam).
        ``graph`` (needs ``fused``, single process): the fused step is captured into a CUDA graph per batch shape after
        ``GRAPH_WARMUP`` eager steps and replayed afterwards -- small batches (the reference's 2 048-point default) are
        bound by the ~100 launches of a step, not by the kernels."""
        self.model, self.pde, self.rl_agent = model, pde, rl_agent
        self.training: TrainingConfig = (getattr(config, "training", None) or config
                                         or getattr(pde.config, "training", None) or TrainingConfig())
        if not isinstance(self.training, TrainingConfig):
            raise TypeError("PDETrainer needs a TrainingConfig")
        self.device = device or next(model.parameters()).device
        oc = optimizer_config or {}
        self.fused = bool(fused)
        self.graph = bool(graph)
        if self.graph and not self.fused:
            raise ValueError("PDETrainer(graph=True) captures the fused step: pass fused=True")
        self._graphs: Dict[tuple, dict] = {}
        lr, wd = oc.get("learning_rate", self.training.learning_rate), oc.get("weight_decay", self.training.weight_decay)
        self._base_lr = float(lr)          # what the optimiser was really built with (optimizer_config wins, trainer.py:292-297)
        aw = self.training.adaptive_weights
        self.use_adaptive_weights = bool(aw.enabled)
        self.adaptive_weights = (AdaptiveLossWeights(aw.strategy, aw.alpha, aw.eps, aw.initial_weights)
                                 if self.use_adaptive_weights else None)
        if self.use_adaptive_weights and self.graph:
            raise ValueError("adaptive re-weighting changes the weights every step on the host: not with graph=True")
        self._is_lbfgs = self.training.optimizer == "lbfgs"
        if self._is_lbfgs and self.fused:
            raise ValueError("fused=True is the Adam step; L-BFGS drives compute_loss + backward through its closure")
        if self._is_lbfgs:
            self.optimizer = self._build_lbfgs(lr)
        elif self.fused:
            # whole step in libpinnk: loss + weighted gradient in one pass per row set, then clip + Adam in two launches
            self.optimizer = FusedAdam(model.parameters(), lr=lr, weight_decay=wd, max_norm=self.training.gradient_clipping,
                                       capturable=self.graph)
            self._flat = None
        else:
            self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        self.scheduler = None
        self._epoch = 0
        if self._is_lbfgs:
            self.scheduler = self._plateau_scheduler()
        elif self.training.scheduler == "cosine" and not self.fused:
            self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
                self.optimizer, T_max=self.training.num_epochs, eta_min=self.training.min_lr)
        self.history: Dict[str, List[float]] = {"train_loss": [], "residual_loss": [], "boundary_loss": [],
                                                "initial_loss": []}

    def train_step(self, x: torch.Tensor, t: torch.Tensor, n_global: Optional[int] = None) -> Dict[str, torch.Tensor]:
        """zero_grad -> compute_loss -> backward -> clip -> Adam (trainer.py:577-578,689-694).
        Multi-GPU: with ``n_global=None`` ``x, t`` are the GLOBAL rows (same on every rank) and each rank takes its
        shard; with ``n_global`` given they already are this rank's shard of an ``n_global``-row batch."""
        if parallel.world_size() > 1 and n_global is None:
            n_global = x.shape[0]
            lo, hi = parallel.shard_bounds(n_global)
            x, t = x[lo:hi], t[lo:hi]
        if self._is_lbfgs:
            return self._lbfgs_step(x, t)              # (adaptive weights are disabled under L-BFGS, trainer.py:464-468)
        if self.use_adaptive_weights and self.training.mode != "data_only":
            return self._adaptive_step(x, t)
        if self.fused and F.data_term_active(self.pde):
            # data term / data_only gating (pde_base.py:1150-1233): the fused objective is physics-only, so these modes take
            # compute_loss + the autograd bridge for the gradient and keep the fused clip + Adam
            if self.graph or parallel.world_size() > 1:
                raise NotImplementedError("observation data / non-forward training modes: use fused=True without graph, single process")
            losses, flat = F.loss_and_flat_grad(self.pde, self.model, x, t)
            self.optimizer.step(flat.contiguous())
            return losses
        if self.fused:
            if self.graph and parallel.world_size() == 1:
                return self._graph_step(x, t)
            return self._fused_step(x, t, n_global)
        self.optimizer.zero_grad(set_to_none=True)
        if parallel.world_size() > 1:
            losses = parallel.sharded_loss_backward(self.pde, self.model, x, t, n_global=n_global)
        else:
            losses = self.pde.compute_loss(self.model, x, t)
            losses["total"].backward()
        if self.training.gradient_clipping > 0:
            nn.utils.clip_grad_norm_(self.model.parameters(), self.training.gradient_clipping)
        self.optimizer.step()
        return losses

    def _fused_step(self, x, t, n_global=None):
        w = parallel.world_size()
        n_loc = x.shape[0]
        n_all = n_loc if n_global is None else int(n_global)
        frac = n_loc / max(n_all, 1) if w > 1 else 1.0
        # one step buffer [flat gradient || residual, boundary, initial sums || pad]: the reverse pass accumulates into
        # its head, the loss sums land in its tail, and the data-parallel step all-reduces it in place (no concatenation)
        P = F.get_program(self.model).grad_floats
        if self._flat is None or self._flat.numel() != P + 4:
            self._flat = torch.zeros(P + 4, dtype=torch.float32, device=x.device if x.is_cuda else self.device)
        buf = self._flat
        comp, (w_res, w_bc, w_ic, w_sm), flat = F.loss_step_flat(self.pde, self.model, x, t, n_global=n_all, res_scale=frac,
                                                                 rest_scale=1.0 / w, flat=buf[:P])
        if w > 1:
            buf[P:P + 4].copy_(comp * comp.new_tensor([frac, 1.0 / w, 1.0 / w, frac]))
            parallel.reduce_inplace(buf)
            sums = buf[P:P + 4]
        else:
            sums = comp
        self.optimizer.step(flat)
        zero = torch.zeros((), device=flat.device)
        total = w_res * sums[0] + w_bc * sums[1] + w_ic * sums[2]
        if w_sm:
            total = total + w_sm * sums[3]
        return {"residual": sums[0], "boundary": sums[1], "initial": sums[2], "smoothness": sums[3], "data": zero,
                "total": total}

    def _adaptive_step(self, x, t):
        """trainer.py:580-694 with adaptive re-weighting: per-component losses and gradients in one reverse pass per row set
        (``G [3, P]``), LRW weights from the rows' norms / RBW weights from the losses, weighted gradient ``w @ G``, clip, Adam.
        ``pde.compat == "reference"`` keeps a quirk of the reference's LRW branch: its last per-component ``backward`` is not
        followed by a ``zero_grad`` (trainer.py:611-622,689), so the gradient it steps with is ``w @ G`` PLUS the gradient
        of the initial-condition component; ``compat == "math"`` steps with ``w @ G``."""
        if parallel.world_size() > 1:
            raise NotImplementedError("adaptive re-weighting is single-process in this mirror")
        comp, G, weights = F.loss_components_and_grads(self.pde, self.model, x, t)
        if weights[3]:
            raise NotImplementedError("adaptive re-weighting with a smoothness term (a fourth component) is not mirrored")
        lrw = self.adaptive_weights.strategy == "lrw"
        if lrw:
            w = self.adaptive_weights.update(gradients=torch.linalg.vector_norm(G, dim=1))
        else:
            w = self.adaptive_weights.update(losses=comp.detach().clone())
        w = w.to(torch.float32)
        flat = w @ G
        if lrw and getattr(self.pde, "compat", "reference") == "reference":
            flat = flat + G[2]
        total = torch.dot(w, comp)
        self.history.setdefault("loss_weights", []).append(w.detach())
        if self.fused:
            self.optimizer.step(flat.contiguous())
        else:
            self.optimizer.zero_grad(set_to_none=True)
            program = F.get_program(self.model)
            for p, g in zip(program.grad_params, program.split_flat(flat)):
                p.grad = g.clone()
            if self.training.gradient_clipping > 0:
                nn.utils.clip_grad_norm_(self.model.parameters(), self.training.gradient_clipping)
            self.optimizer.step()
        zero = torch.zeros((), device=flat.device)
        return {"residual": comp[0], "boundary": comp[1], "initial": comp[2], "smoothness": zero, "data": zero.clone(),
                "total": total}

    def _build_lbfgs(self, lr):
        """trainer.py:299-309."""
        cfg = self.training.lbfgs
        return torch.optim.LBFGS(self.model.parameters(), lr=lr, history_size=cfg.history_size, max_iter=cfg.max_iter,
                                 line_search_fn=cfg.line_search_fn, tolerance_grad=cfg.tolerance_grad,
                                 tolerance_change=cfg.tolerance_change)

    def switch_to_lbfgs(self):
        """trainer.py:366-371: second phase of ``adam_lbfgs`` (the fused / graph Adam step hands over to the closure route)."""
        self.optimizer = self._build_lbfgs(self._base_lr)
        self._is_lbfgs, self.fused, self.graph = True, False, False
        self.scheduler = self._plateau_scheduler()

    def _plateau_scheduler(self):
        """trainer.py:311-325: L-BFGS has its own line search, so the reference overrides whatever scheduler is configured
        with ReduceLROnPlateau (factor / patience / min_lr of the scheduler config), stepped on the epoch's mean loss."""
        t = self.training
        return torch.optim.lr_scheduler.ReduceLROnPlateau(self.optimizer, mode="min", factor=getattr(t, "lr_factor", 0.5),
                                                          patience=getattr(t, "lr_patience", 10), min_lr=t.min_lr)

    def _lbfgs_step(self, x, t):
        """trainer.py:373-389: one L-BFGS step; the closure re-evaluates compute_loss + backward on libpinnk as often as the
        line search asks (up to ``max_iter`` times, ~1.25 evaluations per iteration with strong Wolfe)."""
        if parallel.world_size() > 1:
            raise NotImplementedError("L-BFGS is a full-batch single-process optimiser in the reference (trainer.py:456-462)")
        captured: Dict[str, Dict[str, torch.Tensor]] = {}

        def closure():
            self.optimizer.zero_grad()
            losses = self.pde.compute_loss(self.model, x, t)
            losses["total"].backward()
            captured["losses"] = losses
            return losses["total"]

        self.optimizer.step(closure)
        if "losses" not in captured:               # no closure call happened (tolerance already met)
            captured["losses"] = self.pde.compute_loss(self.model, x, t)
        return captured["losses"]

    GRAPH_WARMUP = 2

    def _graph_step(self, x, t):
        """The fused step through a CUDA graph: the first GRAPH_WARMUP calls for a batch shape run eagerly (they size the
        workspaces and are real optimizer steps), the next one is captured, every later one copies the batch into the
        graph's static rows and replays it.  Boundary / initial rows that the reference draws with torch's RNG stay
        random under replay (torch registers the generator with the graph).  The returned tensors are overwritten by
        the next step of the same shape."""
        key = (tuple(x.shape), tuple(t.shape), x.dtype, t.dtype)
        g = self._graphs.get(key)
        if g is None:
            g = self._graphs[key] = {"calls": 0, "graph": None}
        if g["graph"] is None:
            g["calls"] += 1
            if g["calls"] <= self.GRAPH_WARMUP:
                return self._fused_step(x, t)
            self.optimizer.sync_lr()
            g["x"], g["t"] = x.detach().clone(), t.detach().clone()
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g["out"] = self._fused_step(g["x"], g["t"])
            g["graph"] = graph
            # the graph holds raw pointers into the engines' workspaces: keep those engines alive even if a later, larger
            # batch makes the cache replace them
            from . import engine as _engine
            g["keep"] = list(_engine._CACHE.get(self.model, {}).values())
        else:
            g["x"].copy_(x)
            g["t"].copy_(t)
            self.optimizer.step_count += 1
        self.optimizer.sync_lr()
        g["graph"].replay()
        return g["out"]

    def _compute_validation_loss(self, num_points: int = 1000) -> Dict[str, float]:
        """trainer.py:140-162: compute_loss on freshly generated points, returned as floats.  The reference builds the full
        autograd graph for it (its derivatives ARE autograd); here the jets are forward-mode, so the call runs under
        no_grad: forward kernels only, no stash, no gradient buffers.  Same RNG consumption as the reference
        (one generate_collocation_points call); as there, eval() is undone by the derivative evaluation (pde_base.py:638)."""
        self.model.eval()
        x_val, t_val = self.pde.generate_collocation_points(num_points)
        with torch.no_grad():
            losses = self.pde.compute_loss(self.model, x_val.to(self.device), t_val.to(self.device))
        return {"total_loss": losses["total"].item(), "residual_loss": losses["residual"].item(),
                "boundary_loss": losses["boundary"].item(), "initial_loss": losses["initial"].item()}

    def live_snapshot(self, epoch: int, grid_size: int = 60) -> Dict[str, object]:
        """The arrays of trainer.py:171-279 (``live_snapshot.npz``): predicted u and the signed residual on a fixed
        grid_size x grid_size grid -- (x, t) for 1-D problems, (x1, x2) at the middle of the time interval otherwise.
        Both fields come from forward-only libpinnk passes (no autograd graph, no stash)."""
        import numpy as np
        pde, dev = self.pde, self.device
        dim = int(getattr(pde, "dimension", 1))
        t_lo, t_hi = float(pde.time_domain[0]), float(pde.time_domain[1])
        if dim <= 1:
            ax, ay = (np.linspace(float(pde.domain[0][0]), float(pde.domain[0][1]), grid_size, dtype=np.float32),
                      np.linspace(t_lo, t_hi, grid_size, dtype=np.float32))
            xx, tt = np.meshgrid(ax, ay, indexing="xy")
            x_flat = torch.tensor(xx.reshape(-1, 1), device=dev)
            t_flat = torch.tensor(tt.reshape(-1, 1), device=dev)
            meta = dict(dimension=1, x_label="x", y_label="t", fixed_t=float("nan"))
        else:
            fixed_t = 0.5 * (t_lo + t_hi)
            ax = np.linspace(float(pde.domain[0][0]), float(pde.domain[0][1]), grid_size, dtype=np.float32)
            ay = np.linspace(float(pde.domain[1][0]), float(pde.domain[1][1]), grid_size, dtype=np.float32)
            x1, x2 = np.meshgrid(ax, ay, indexing="xy")
            x_flat = torch.tensor(np.stack([x1.reshape(-1), x2.reshape(-1)], axis=1), device=dev, dtype=torch.float32)
            t_flat = torch.full((x_flat.shape[0], 1), fixed_t, dtype=torch.float32, device=dev)
            meta = dict(dimension=2, x_label="x1", y_label="x2", fixed_t=float(fixed_t))
        was_training = self.model.training
        try:
            with torch.no_grad():
                u = F.model_forward(self.model, torch.cat([x_flat, t_flat], dim=1))
                u_np = u.detach().cpu().numpy()
                if u_np.ndim == 2 and u_np.shape[-1] > 1:
                    u_np = u_np[..., 0]
                u_np = u_np.reshape(grid_size, grid_size)
                try:
                    r = self.pde.compute_residual(self.model, x_flat, t_flat)
                    r_np = r.detach().cpu().numpy().reshape(grid_size, grid_size)
                except Exception:
                    r_np = np.zeros_like(u_np)           # the reference swallows residual failures here too
        finally:
            self.model.train(was_training)
        return dict(axis_x=ax, axis_y=ay, u_pred=u_np, residual=r_np, epoch=int(epoch), **meta)

    def _save_live_snapshot(self, experiment_dir: str, epoch: int, grid_size: int = 60) -> None:
        """trainer.py:171-279: write ``live_snapshot.npz`` for the dashboard's monitor tab (same keys)."""
        import numpy as np
        if not experiment_dir:
            return
        np.savez(os.path.join(experiment_dir, "live_snapshot.npz"), **self.live_snapshot(epoch, grid_size))

    def cosine_lr(self, epoch: int) -> float:
        t = self.training
        return t.min_lr + 0.5 * (self._base_lr - t.min_lr) * (1 + __import__("math").cos(__import__("math").pi * epoch / max(t.num_epochs, 1)))

    def train(self, num_epochs: int, batch_size: int, num_points: int, experiment_dir: str = None):
        self.model.train()
        if self._is_lbfgs:
            batch_size = num_points                # trainer.py:456-462: L-BFGS needs the deterministic full-batch loss
        for _ in range(num_epochs):
            epoch = []
            for _ in range(num_points // batch_size):
                strategy = "adaptive" if self.rl_agent is not None else self.training.collocation_distribution
                kw = {"model": self.model} if strategy == "residual_based" else {}
                x, t = self.pde.generate_collocation_points(batch_size, strategy=strategy, **kw)
                losses = self.train_step(x.to(self.device), t.to(self.device))
                epoch.append({k: float(v.item()) for k, v in losses.items() if k in ("total", "residual", "boundary", "initial")})
            if isinstance(self.scheduler, torch.optim.lr_scheduler.ReduceLROnPlateau):
                if epoch:
                    self.scheduler.step(sum(e["total"] for e in epoch) / len(epoch))      # trainer.py:164-169
            elif self.scheduler is not None:
                self.scheduler.step()
            elif self.fused and self.training.scheduler == "cosine":
                self._epoch += 1
                self.optimizer.param_groups[0]["lr"] = self.cosine_lr(self._epoch)
            if epoch:
                for key, name in (("total", "train_loss"), ("residual", "residual_loss"),
                                  ("boundary", "boundary_loss"), ("initial", "initial_loss")):
                    self.history[name].append(sum(e[key] for e in epoch) / len(epoch))
        return self.history

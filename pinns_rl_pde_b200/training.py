"""Host-side mirror of the inner training step of ``pinnrl.training.trainer.PDETrainer``
(trainer.py:292-332 optimiser/scheduler set-up, :539-698 the step).  Plots, dashboards, metadata
files, adaptive loss re-weighting and L-BFGS are outside the hot path and not mirrored; for those,
patch the reference (``patch_reference``) and use its own trainer.

With ``world_size > 1`` (one process per GPU, torch.distributed/NCCL) collocation rows are sharded
across ranks and the flat gradient + loss sums are all-reduced once per step (parallel.py).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch
import torch.nn as nn

from . import functional as F
from . import parallel


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

    def __post_init__(self):
        if self.loss_weights is None:
            self.loss_weights = {"residual": 1.0, "boundary": 1.0, "initial": 1.0}
        self.loss_weights.setdefault("data", 1.0)

    def get(self, key, default=None):
        return getattr(self, key, default)

    def __getitem__(self, key):
        return getattr(self, key)


class FusedAdam:
    """clip_grad_norm_ + Adam(L2 weight decay) on the flat gradient in two libpinnk launches
    (trainer.py:690-694,292-297); state and update rule identical to ``torch.optim.Adam``."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, max_norm=0.0):
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

    def step(self, flat_grad: torch.Tensor):
        C, L = self._C, self._lib
        self.step_count += 1
        for i, p in enumerate(self.params):
            if not (p.is_cuda and p.is_contiguous() and p.dtype == torch.float32):
                raise L.PinnkError("FusedAdam needs contiguous float32 CUDA parameters")
            self._ptrs[i] = p.data_ptr()
        lr = float(self.param_groups[0]["lr"])
        L.check(L.load().pinnk_adam_step(C.cast(self._ptrs, C.c_void_p), C.cast(self._numels, C.c_void_p), len(self.params),
                                         flat_grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                         self._scratch.data_ptr(), self.step_count, lr, self.betas[0], self.betas[1],
                                         self.eps, self.weight_decay, self.max_norm,
                                         C.c_void_p(torch.cuda.current_stream().cuda_stream)), "pinnk_adam_step")


class PDETrainer:
    def __init__(self, model: nn.Module, pde, optimizer_config=None, config=None, device=None, rl_agent=None,
                 fused: bool = False):
        self.model, self.pde, self.rl_agent = model, pde, rl_agent
        self.training: TrainingConfig = (getattr(config, "training", None) or config
                                         or getattr(pde.config, "training", None) or TrainingConfig())
        if not isinstance(self.training, TrainingConfig):
            raise TypeError("PDETrainer needs a TrainingConfig")
        self.device = device or next(model.parameters()).device
        oc = optimizer_config or {}
        self.fused = bool(fused)
        lr, wd = oc.get("learning_rate", self.training.learning_rate), oc.get("weight_decay", self.training.weight_decay)
        if self.fused:
            # whole step in libpinnk: loss + weighted gradient in one pass per row set, then clip + Adam in two launches
            self.optimizer = FusedAdam(model.parameters(), lr=lr, weight_decay=wd, max_norm=self.training.gradient_clipping)
            self._flat = None
        else:
            self.optimizer = torch.optim.Adam(model.parameters(), lr=lr, weight_decay=wd)
        self.scheduler = None
        self._epoch = 0
        if self.training.scheduler == "cosine" and not self.fused:
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
        if self.fused:
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
        comp, (w_res, w_bc, w_ic), flat = F.loss_step_flat(self.pde, self.model, x, t, n_global=n_all, res_scale=frac,
                                                           rest_scale=1.0 / w, flat=self._flat)
        self._flat = flat
        sums = torch.stack([frac * comp[0], comp[1] / w, comp[2] / w])
        if w > 1:
            flat, sums = parallel.reduce_flat(flat, sums)
        self.optimizer.step(flat)
        zero = torch.zeros((), device=flat.device)
        return {"residual": sums[0], "boundary": sums[1], "initial": sums[2], "smoothness": zero, "data": zero.clone(),
                "total": w_res * sums[0] + w_bc * sums[1] + w_ic * sums[2]}

    def cosine_lr(self, epoch: int) -> float:
        t = self.training
        return t.min_lr + 0.5 * (t.learning_rate - t.min_lr) * (1 + __import__("math").cos(__import__("math").pi * epoch / max(t.num_epochs, 1)))

    def train(self, num_epochs: int, batch_size: int, num_points: int, experiment_dir: str = None):
        self.model.train()
        for _ in range(num_epochs):
            epoch = []
            for _ in range(num_points // batch_size):
                strategy = "adaptive" if self.rl_agent is not None else self.training.collocation_distribution
                kw = {"model": self.model} if strategy == "residual_based" else {}
                x, t = self.pde.generate_collocation_points(batch_size, strategy=strategy, **kw)
                losses = self.train_step(x.to(self.device), t.to(self.device))
                epoch.append({k: float(v.item()) for k, v in losses.items() if k in ("total", "residual", "boundary", "initial")})
            if self.scheduler is not None:
                self.scheduler.step()
            elif self.fused and self.training.scheduler == "cosine":
                self._epoch += 1
                self.optimizer.param_groups[0]["lr"] = self.cosine_lr(self._epoch)
            if epoch:
                for key, name in (("total", "train_loss"), ("residual", "residual_loss"),
                                  ("boundary", "boundary_loss"), ("initial", "initial_loss")):
                    self.history[name].append(sum(e[key] for e in epoch) / len(epoch))
        return self.history

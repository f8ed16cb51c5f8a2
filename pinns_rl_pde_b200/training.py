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


class PDETrainer:
    def __init__(self, model: nn.Module, pde, optimizer_config=None, config=None, device=None, rl_agent=None):
        self.model, self.pde, self.rl_agent = model, pde, rl_agent
        self.training: TrainingConfig = (getattr(config, "training", None) or config
                                         or getattr(pde.config, "training", None) or TrainingConfig())
        if not isinstance(self.training, TrainingConfig):
            raise TypeError("PDETrainer needs a TrainingConfig")
        self.device = device or next(model.parameters()).device
        oc = optimizer_config or {}
        self.optimizer = torch.optim.Adam(model.parameters(), lr=oc.get("learning_rate", self.training.learning_rate),
                                          weight_decay=oc.get("weight_decay", self.training.weight_decay))
        self.scheduler = None
        if self.training.scheduler == "cosine":
            self.scheduler = torch.optim.lr_scheduler.CosineAnnealingLR(
                self.optimizer, T_max=self.training.num_epochs, eta_min=self.training.min_lr)
        self.history: Dict[str, List[float]] = {"train_loss": [], "residual_loss": [], "boundary_loss": [],
                                                "initial_loss": []}

    def train_step(self, x: torch.Tensor, t: torch.Tensor) -> Dict[str, torch.Tensor]:
        """zero_grad -> compute_loss -> backward -> clip -> Adam (trainer.py:577-578,689-694)."""
        self.optimizer.zero_grad(set_to_none=True)
        if parallel.world_size() > 1:
            losses = parallel.sharded_loss_backward(self.pde, self.model, x, t)
        else:
            losses = self.pde.compute_loss(self.model, x, t)
            losses["total"].backward()
        if self.training.gradient_clipping > 0:
            nn.utils.clip_grad_norm_(self.model.parameters(), self.training.gradient_clipping)
        self.optimizer.step()
        return losses

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
            if epoch:
                for key, name in (("total", "train_loss"), ("residual", "residual_loss"),
                                  ("boundary", "boundary_loss"), ("initial", "initial_loss")):
                    self.history[name].append(sum(e[key] for e in epoch) / len(epoch))
        return self.history

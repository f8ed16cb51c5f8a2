"""Runs the UNMODIFIED reference trainer (pinnrl.training.trainer.PDETrainer, plotting deps stubbed) for one LRW step on the
CPU and compares the gradient it leaves in ``.grad`` with (a) clip(w @ G) and (b) clip(w @ G + g_initial): the reference's
LRW branch does not zero the gradients after its last per-component backward (trainer.py:611-622,689), so (b) is what it
steps with.  Build container only (needs /root/reference):  python tests/golden/lrw_quirk_probe.py
Recorded result (2026-10-18): rel diff to w @ G 0.659, to w @ G + g_initial 7.4e-08 -- this is what
PDETrainer._adaptive_step reproduces under compat="reference"."""
import sys, copy, os, tempfile
from unittest.mock import MagicMock
sys.path.insert(0, "/root/reference")
for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "plotly", "plotly.graph_objects", "plotly.subplots", "plotly.express"):
    mm = MagicMock(); mm.__path__ = []; sys.modules[_m] = mm
import torch
from pinnrl.config import (AdaptiveWeightsConfig, Config, EarlyStoppingConfig, LBFGSConfig, LearningRateSchedulerConfig, ModelConfig, TrainingConfig)
from pinnrl.neural_networks import PINNModel
from pinnrl.pdes.burgers_equation import BurgersEquation
from pinnrl.pdes.pde_base import PDEConfig
from pinnrl.training.trainer import PDETrainer
DEV = torch.device("cpu")
tc = TrainingConfig(num_epochs=1, batch_size=64, num_collocation_points=64, num_boundary_points=16, num_initial_points=16,
    learning_rate=1e-2, weight_decay=0.0, gradient_clipping=1.0,
    early_stopping=EarlyStoppingConfig(enabled=False, patience=999, min_delta=1e-7),
    learning_rate_scheduler=LearningRateSchedulerConfig(type="cosine", warmup_epochs=0, min_lr=1e-6, factor=0.5, patience=3),
    collocation_distribution="uniform", adaptive_weights=AdaptiveWeightsConfig(enabled=True, strategy="lrw"),
    loss_weights={"residual": 1.0, "boundary": 1.0, "initial": 1.0, "smoothness": 0.0}, optimizer="adam", lbfgs=LBFGSConfig())
cfg = Config.__new__(Config); cfg.device = DEV
cfg.model = ModelConfig(input_dim=2, hidden_dim=16, output_dim=1, num_layers=3, activation="tanh", architecture="feedforward")
cfg.training = tc
pde = BurgersEquation(config=PDEConfig(name="burgers", domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0], parameters={"nu": 0.01},
    boundary_conditions={"dirichlet": {"value": 0.0}}, initial_condition={"type": "sine", "amplitude": -1.0, "frequency": 1.0},
    exact_solution={}, dimension=1, device=DEV, training=tc))
torch.manual_seed(0)
model = PINNModel(config=cfg, device=DEV)
m0 = copy.deepcopy(model)
g = torch.Generator().manual_seed(1)
X = torch.rand(64, 1, generator=g) * 2 - 1; T = torch.rand(64, 1, generator=g)
pde.generate_collocation_points = lambda n, strategy="uniform", **kw: (X.clone(), T.clone())
tr = PDETrainer(model=model, pde=pde, optimizer_config={"learning_rate": 1e-2, "weight_decay": 0.0}, config=cfg, device=DEV, validation_frequency=1000)
import io, contextlib
with contextlib.redirect_stdout(io.StringIO()):
    tr.train(num_epochs=1, batch_size=64, num_points=64, experiment_dir=tempfile.mkdtemp())
got = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
# emulate on the untouched copy with the reference's own compute_loss
params = list(m0.parameters())
losses = pde.compute_loss(m0, X.clone(), T.clone())
gs = [torch.cat([(torch.zeros_like(p) if gr is None else gr).reshape(-1) for p, gr in zip(params, torch.autograd.grad(losses[k], params, retain_graph=True, allow_unused=True))]) for k in ("residual", "boundary", "initial")]
w = torch.tensor([0.5, 0.3, 0.2])
def clipped(flat):
    n = flat.norm(); return flat * min(1.0, 1.0 / (float(n) + 1e-6))
a = clipped(sum(wk * gk for wk, gk in zip(w, gs)))
b = clipped(sum(wk * gk for wk, gk in zip(w, gs)) + gs[2])
print("rel diff to  w@G          :", float((got - a).norm() / got.norm()))
print("rel diff to  w@G + g_init :", float((got - b).norm() / got.norm()))

"""Run the UNMODIFIED reference trainer (pinnrl from baseline/_ref) -- stock or with patch_reference() -- and return
its loss history.  Test infrastructure (SURVEY section 7 step 9 / VERDICT r01 item 1): the same seeds, the same torch
RNG stream (the patched path consumes none), the same ``PDETrainer.train`` loop; only ``compute_residual`` /
``compute_loss`` differ.

``dtype=torch.float64`` runs the stock reference in double precision as the yardstick: the model is built in fp32
from the same seed and then cast, collocation points are still drawn in fp32 (same generator stream) and cast, and
everything the PDE builds itself (boundary / initial rows) follows the default dtype.
"""
from __future__ import annotations

import contextlib
import io
import logging
import math

import torch

CONFIGS = {
    # BASELINE configs[0] (README quick-start, README.md:99-141): Heat 1-D, fourier 4x128
    "c1_heat_fourier": dict(pde="heat", arch="fourier", hidden=128, layers=4, domain=[[0.0, 1.0]], time=[0.0, 1.0],
                            params={"alpha": 0.01}, bcs={"type": "dirichlet"},
                            ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0},
                            exact={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0},
                            lr=1e-3, wd=1e-4, nb=500, ni=500),
    # BASELINE configs[1] network and PDE at the README's batch sizes: Burgers, feedforward 8x128
    "c2_burgers_ff": dict(pde="burgers", arch="feedforward", hidden=128, layers=8, domain=[[-1.0, 1.0]], time=[0.0, 1.0],
                          params={"nu": 0.01 / math.pi}, bcs={"dirichlet": {"value": 0.0}},
                          ic={"type": "sine", "amplitude": -1.0, "frequency": 1.0}, exact={},
                          lr=1e-3, wd=1e-4, nb=500, ni=500),
}


def _classes():
    from oracle import ref_env
    ref_env.activate()
    from pinnrl.pdes.heat_equation import HeatEquation
    from pinnrl.pdes.burgers_equation import BurgersEquation
    return {"heat": HeatEquation, "burgers": BurgersEquation}


def build(name: str, device, epochs: int, batch_size: int = 2048, num_points: int = 5000, seed: int = 0,
          dtype=torch.float32, adaptive=None, distribution: str = "uniform"):
    """(trainer, model, pde) of the reference, built the way README.md:99-141 does."""
    classes = _classes()
    from pinnrl.config import (AdaptiveWeightsConfig, Config, EarlyStoppingConfig, LBFGSConfig,
                               LearningRateSchedulerConfig, ModelConfig, TrainingConfig)
    from pinnrl.neural_networks import PINNModel
    from pinnrl.pdes.pde_base import PDEConfig
    from pinnrl.training.trainer import PDETrainer
    c = CONFIGS[name]
    device = torch.device(device)
    training = TrainingConfig(
        num_epochs=epochs, batch_size=batch_size, num_collocation_points=num_points, num_boundary_points=c["nb"],
        num_initial_points=c["ni"], learning_rate=c["lr"], weight_decay=c["wd"], gradient_clipping=1.0,
        early_stopping=EarlyStoppingConfig(enabled=False, patience=10 ** 6, min_delta=1e-7),
        learning_rate_scheduler=LearningRateSchedulerConfig(type="cosine", warmup_epochs=0, min_lr=1e-6, factor=0.5, patience=3),
        collocation_distribution=distribution,
        adaptive_weights=AdaptiveWeightsConfig(enabled=adaptive is not None, strategy=adaptive or "rbw"),
        loss_weights={"residual": 1.0, "boundary": 10.0, "initial": 10.0, "smoothness": 0.0}, optimizer="adam",
        lbfgs=LBFGSConfig())
    pde_config = PDEConfig(name=c["pde"], domain=c["domain"], time_domain=c["time"], parameters=dict(c["params"]),
                           boundary_conditions=dict(c["bcs"]), initial_condition=dict(c["ic"]),
                           exact_solution=dict(c["exact"]), dimension=1, device=device, training=training)
    config = Config.__new__(Config)
    config.device = device
    config.model = ModelConfig(input_dim=2, hidden_dim=c["hidden"], output_dim=1, num_layers=c["layers"],
                               activation="tanh", architecture=c["arch"])
    config.training = training
    config.pde_config = pde_config
    torch.manual_seed(seed)
    if device.type == "cuda":
        torch.cuda.manual_seed_all(seed)
    with contextlib.redirect_stdout(io.StringIO()):
        pde = classes[c["pde"]](config=pde_config)
        model = PINNModel(config=config, device=device)
    if dtype == torch.float64:
        model = model.double()
        sample = pde.generate_collocation_points

        def sample64(*a, **kw):          # same fp32 draws from the same generator stream, cast afterwards
            torch.set_default_dtype(torch.float32)
            try:
                x, t = sample(*a, **kw)
            finally:
                torch.set_default_dtype(torch.float64)
            return x.double(), t.double()
        pde.generate_collocation_points = sample64
    trainer = PDETrainer(model=model, pde=pde, optimizer_config={"learning_rate": c["lr"], "weight_decay": c["wd"]},
                         config=config, device=device, validation_frequency=10,
                         early_stopping_config={"enabled": False, "patience": 10 ** 6})
    return trainer, model, pde


def run(name: str, device, epochs: int, batch_size: int = 2048, num_points: int = 5000, seed: int = 0,
        dtype=torch.float32, adaptive=None):
    """Per-epoch train loss and validation loss of ``PDETrainer.train`` (experiment_dir=None: no files, no plots)."""
    from tqdm import tqdm
    tqdm.__init__.__defaults__  # noqa: B018  (tqdm is a real dependency of the reference)
    logging.disable(logging.CRITICAL)
    old = torch.get_default_dtype()
    try:
        trainer, model, pde = build(name, device, epochs, batch_size, num_points, seed, dtype, adaptive)
        if dtype == torch.float64:
            torch.set_default_dtype(torch.float64)
        torch.manual_seed(seed + 1)
        if torch.device(device).type == "cuda":
            torch.cuda.manual_seed_all(seed + 1)
        import os
        os.environ.setdefault("TQDM_DISABLE", "1")
        with contextlib.redirect_stdout(io.StringIO()), contextlib.redirect_stderr(io.StringIO()):
            trainer.train(num_epochs=epochs, batch_size=batch_size, num_points=num_points, experiment_dir=None)
    finally:
        torch.set_default_dtype(old)
        logging.disable(logging.NOTSET)
    hist = trainer.history
    flat = torch.cat([p.detach().double().reshape(-1).cpu() for p in model.parameters()])
    return {"train_loss": [float(v) for v in hist["train_loss"]], "val_loss": [float(v) for v in hist["val_loss"]],
            "params": flat}


def pointwise(a, b):
    """|a - b| / |b| per epoch."""
    return [abs(x - y) / max(abs(y), 1e-300) for x, y in zip(a, b)]


def window_medians(dev, width=100):
    """{(lo, hi): median of the pointwise deviations in epochs [lo, hi)} -- robust to the loss spikes of the chaotic phase."""
    out = {}
    for lo in range(0, len(dev), width):
        w = sorted(dev[lo:lo + width])
        if w:
            out[(lo, min(lo + width, len(dev)))] = w[len(w) // 2]
    return out


def deviation(a, b):
    """max over epochs of |a - b| / |b| for the windows [0, E): {E: dev}."""
    out = {}
    n = min(len(a), len(b))
    worst = 0.0
    marks = {e for e in (10, 50, 100, 200, 300, 500, n) if e <= n}
    for i in range(n):
        worst = max(worst, abs(a[i] - b[i]) / max(abs(b[i]), 1e-300))
        if i + 1 in marks:
            out[i + 1] = worst
    return out

"""GPU: the finite-difference smoothness regulariser of HeatEquation.compute_loss (heat_equation.py:625-650; the shipped YAML
sets its weight to 0.1, config.yaml:338-343) through compute_loss, through the fused trainer step (paired mean-|e| segments on
value-only rows, functional._smoothness_calls) and against the unmodified reference's values (x_heat_smoothness.npz).

The term divides value differences by eps = 1e-4, so fp32 round-off is amplified 10^4 times: the reference's own fp32 run is
6e-5 (value) / 1.6e-4 (gradient) away from its fp64 run.  Gate: err(cuda, ref64) <= max(1e-5, 2 err(ref32, ref64))."""
import os

import numpy as np
import pytest
import torch

import parity_log
from helpers import GOLDEN, flat_grad, rel

pytestmark = pytest.mark.gpu


def _setup(dev, smooth=0.1):
    import pinns_rl_pde_b200 as pk
    z = np.load(os.path.join(GOLDEN, "x_heat_smoothness.npz"))
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    model = pk.make_model("fourier", 2, 128, 3, dev, mapping_size=32, scale=10.0)
    model.load_state_dict(state)
    training = {"num_collocation_points": 300, "num_boundary_points": 40, "num_initial_points": 40,
                "loss_weights": {"residual": 1.0, "boundary": 10.0, "initial": 10.0, "smoothness": smooth}}
    cfg = pk.PDEConfig(name="heat", domain=[[0.0, 1.0]], time_domain=[0.0, 1.0], parameters={"alpha": 0.01},
                       boundary_conditions={"dirichlet": {"type": "dirichlet"}},
                       initial_condition={"type": "sine", "amplitude": 1.0, "frequency": 2.0},
                       exact_solution={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0}, dimension=1,
                       device=dev, training=training)
    pde = pk.HeatEquation(cfg)
    return z, model, pde, torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["t"]).to(dev)


def test_compute_loss_with_smoothness_matches_the_reference():
    dev = torch.device("cuda:0")
    z, model, pde, x, t = _setup(dev)
    L = pde.compute_loss(model, x, t)
    L["total"].backward()
    g = flat_grad(model)
    names = ("residual", "boundary", "initial", "smoothness", "total")
    for i, k in enumerate(names):
        err = abs(float(L[k]) - z["loss64"][i]) / abs(z["loss64"][i])
        floor = abs(z["loss32"][i] - z["loss64"][i]) / abs(z["loss64"][i])
        gate = max(1e-5, 2 * floor)
        parity_log.log(f"[heat smoothness] {k}: cuda-vs-ref64 {err:.3e} | ref32-vs-ref64 {floor:.3e} | gate {gate:.3e}")
        assert err <= gate, (k, err, gate)
    eg, gate = rel(g, z["gtotal64"]), max(1e-5, 2 * float(z["gtotal32_err"]))
    parity_log.log(f"[heat smoothness] grad(total, smoothness weight 0.1): cuda-vs-ref64 {eg:.3e} | ref32-vs-ref64 "
                   f"{float(z['gtotal32_err']):.3e} | gate {gate:.3e}")
    assert eg <= gate


def test_fused_step_carries_the_smoothness_term():
    """loss_step_flat (the fused trainer step's loss + gradient pass) with the term on: component 3 and the flat gradient
    equal compute_loss + backward of the same library (the model values of both routes are bit-identical, so the signs of
    the finite differences agree), and the smoothness gradient alone matches the reference's fp64 gradient."""
    from pinns_rl_pde_b200 import functional as F
    dev = torch.device("cuda:0")
    z, model, pde, x, t = _setup(dev)
    L = pde.compute_loss(model, x, t)
    L["total"].backward()
    g_auto = flat_grad(model).clone()
    comp, w, flat = F.loss_step_flat(pde, model, x, t)
    assert w == (1.0, 10.0, 10.0, 0.1)
    for i, k in enumerate(("residual", "boundary", "initial", "smoothness")):
        assert abs(float(comp[i]) - float(L[k])) <= 2e-6 * abs(float(L[k])), (k, float(comp[i]), float(L[k]))
    e = rel(flat, g_auto)
    parity_log.log(f"[heat smoothness] fused step gradient vs compute_loss + backward: {e:.3e}")
    assert e <= 2e-6
    # the term alone: weights (0, 0, 0) for the physics components leave w_smooth * d smoothness / d theta
    _, model2, pde2, _, _ = _setup(dev)
    pde2.config.training["loss_weights"] = {"residual": 0.0, "boundary": 0.0, "initial": 0.0, "smoothness": 1.0}
    comp2, _, flat2 = F.loss_step_flat(pde2, model2, x, t)
    es = abs(float(comp2[3]) - float(z["smooth64"])) / float(z["smooth64"])
    eg, gate = rel(flat2, z["gsmooth64"]), max(1e-5, 2 * float(z["gsmooth32_err"]))
    parity_log.log(f"[heat smoothness] term alone: value cuda-vs-ref64 {es:.3e}, gradient {eg:.3e} | ref32-vs-ref64 "
                   f"{float(z['gsmooth32_err']):.3e} | gate {gate:.3e}")
    assert eg <= gate and es <= max(1e-5, 2 * abs(float(z["smooth32"]) - float(z["smooth64"])) / float(z["smooth64"]))


def test_fused_trainer_with_smoothness_follows_the_autograd_route():
    import copy
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    _, m1, pde, _, _ = _setup(dev)
    m2 = copy.deepcopy(m1)
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=1e-4, gradient_clipping=1.0, scheduler="none",
                            loss_weights={"residual": 1.0, "boundary": 10.0, "initial": 10.0, "smoothness": 0.1})
    pde.config.training = cfg
    t1 = pk.PDETrainer(m1, pde, config=cfg, device=dev, fused=True)
    t2 = pk.PDETrainer(m2, pde, config=cfg, device=dev, fused=False)
    g = torch.Generator().manual_seed(3)
    for it in range(10):
        x, t = torch.rand(2000, 1, generator=g).to(dev), torch.rand(2000, 1, generator=g).to(dev)
        l1, l2 = t1.train_step(x, t), t2.train_step(x, t)
        assert float(l1["smoothness"]) > 0
        assert abs(float(l1["total"]) - float(l2["total"])) <= 5e-5 * abs(float(l2["total"])), (it, float(l1["total"]), float(l2["total"]))
    p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    parity_log.log(f"[heat smoothness] 10 fused trainer steps vs autograd route: parameters {rel(p1, p2):.3e}")
    assert rel(p1, p2) <= 1e-4

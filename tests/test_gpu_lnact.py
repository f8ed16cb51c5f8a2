"""GPU: LayerNorm + activation as one kernel each way (resnet.py:45-65) against the separate LayerNorm and activation kernels of the
same library on identical seeded inputs.  Two fused routes: the feature-per-thread kernels of lnact_feat.cu (the default where they
apply: tanh, width 128 / 256 / 512, two-direction jets) and the warp-per-point pair lnact_fwd_kernel / lnact_bwd_kernel
(PINNK_ENABLE_LNACT=1, measured slower, kept for comparison); PINNK_ENABLE_LNACT=0 selects the separate kernels.  The switch is
read per call, so all routes run in one process.  Both routes are also covered against the oracle by test_gpu_parity.py's ResNet
cases (default route) ."""
import numpy as np
import pytest
import torch

import parity_log
from helpers import flat_grad, product_pde

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


def _step(pde, model, x, t):
    for p in model.parameters():
        p.grad = None
    losses = pde.compute_loss(model, x, t)
    losses["total"].backward()
    torch.cuda.synchronize()
    return ({k: float(losses[k]) for k in ("residual", "boundary", "initial", "total")}, flat_grad(model).clone(),
            pde.compute_residual(model, x, t).detach().clone(), pde.score_residual(model, x, t)[0].clone())


@pytest.mark.parametrize("route", ["feat", "warp"])
@pytest.mark.parametrize("pde_name,width,blocks,n", [("kdv", 256, 2, 3000), ("burgers", 128, 3, 2500), ("cahn_hilliard", 256, 1, 1100),
                                                     ("heat", 512, 1, 700), ("wave", 256, 1, 900), ("convection", 128, 2, 1300),
                                                     ("pendulum", 128, 1, 600)])
def test_fused_layernorm_activation_matches_separate_kernels(monkeypatch, pde_name, width, blocks, n, route):
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    torch.manual_seed(3)
    model = pk.make_model("resnet", 2, width, blocks, dev)
    with torch.no_grad():          # LayerNorm affine parameters away from their (1, 0) initial values
        for m in model.modules():
            if isinstance(m, torch.nn.LayerNorm):
                m.weight.add_(0.3 * torch.randn_like(m.weight))
                m.bias.add_(0.2 * torch.randn_like(m.bias))
    pde = product_pde(pde_name, dev)
    g = torch.Generator().manual_seed(5)
    lo, hi = pde.domain[0]
    x = (lo + (hi - lo) * torch.rand(n, 1, generator=g)).to(dev)
    t = (pde.time_domain[0] + (pde.time_domain[1] - pde.time_domain[0]) * torch.rand(n, 1, generator=g)).to(dev)
    from pinns_rl_pde_b200 import _lib
    if route == "warp":
        if width > 256:
            pytest.skip("the warp-per-point pair covers widths up to 256")
        monkeypatch.setenv("PINNK_ENABLE_LNACT", "1")
    else:
        monkeypatch.delenv("PINNK_ENABLE_LNACT", raising=False)
    before = _lib.launch_count()
    fused = _step(pde, model, x, t)
    n_fused = _lib.launch_count() - before
    monkeypatch.setenv("PINNK_ENABLE_LNACT", "0")
    before = _lib.launch_count()
    plain = _step(pde, model, x, t)
    n_plain = _lib.launch_count() - before
    assert n_fused < n_plain, (n_fused, n_plain)              # the fused route really ran (fewer launches)
    for k in fused[0]:
        assert abs(fused[0][k] - plain[0][k]) <= 2e-6 * abs(plain[0][k]) + 1e-12, (k, fused[0][k], plain[0][k])
    eg, er, es = _rel(fused[1], plain[1]), _rel(fused[2], plain[2]), _rel(fused[3], plain[3])
    parity_log.log(f"[lnact {route} {pde_name} resnet {blocks}x{width}] fused vs separate kernels: grad {eg:.2e}, residual {er:.2e}, "
                   f"|r| scores {es:.2e}; launches {n_fused} vs {n_plain}")
    assert eg <= 3e-6 and er <= 3e-6 and es <= 3e-6, (eg, er, es)
    assert np.isfinite(fused[1].cpu().numpy()).all()


@pytest.mark.parametrize("graph", [False, True])
def test_fused_trainer_on_a_residual_network_follows_the_autograd_route(graph):
    """The fused trainer step (pinnk_loss_step_flags with its KEEP / REUSE stash flows, optionally replayed as a CUDA graph) on a
    residual network -- LayerNorm + tanh through lnact_feat.cu -- against compute_loss + backward + clip + Adam on a copy."""
    import copy
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    m1 = pk.make_model("resnet", 2, 256, 2, dev)
    m2 = copy.deepcopy(m1)
    pde = product_pde("kdv", dev)
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=1e-4, gradient_clipping=1.0, scheduler="none")
    pde.config.training = cfg
    t1 = pk.PDETrainer(m1, pde, config=cfg, device=dev, fused=True, graph=graph)
    t2 = pk.PDETrainer(m2, pde, config=cfg, device=dev, fused=False)
    g = torch.Generator().manual_seed(3)
    lo, hi = pde.domain[0]
    t_lo, t_hi = pde.time_domain
    for it in range(10):
        x = (lo + (hi - lo) * torch.rand(1536, 1, generator=g)).to(dev)
        t = (t_lo + (t_hi - t_lo) * torch.rand(1536, 1, generator=g)).to(dev)
        l1, l2 = t1.train_step(x, t), t2.train_step(x, t)
        assert abs(float(l1["total"]) - float(l2["total"])) <= 5e-5 * abs(float(l2["total"])), (it, float(l1["total"]), float(l2["total"]))
    p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    e = _rel(p1, p2)
    parity_log.log(f"[lnact trainer kdv resnet 2x256, graph={graph}] 10 fused steps vs autograd route: parameters {e:.3e}")
    assert e <= 1e-4

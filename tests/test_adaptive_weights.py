"""Adaptive loss re-weighting (trainer.py:580-634, components/adaptive_weights.py).
CPU: oracle port and the product's AdaptiveLossWeights against the fixture generated from the unmodified reference.
GPU: PDETrainer's re-weighted step (per-component gradient matrix G, weights from its row norms / the losses, w @ G) against
the same algorithm assembled from compute_loss + per-component autograd.grad + torch clip / Adam."""
import copy
import os

import numpy as np
import pytest
import torch

from oracle import ref_port

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "x_adaptive_weights.npz")


@pytest.mark.parametrize("strategy", ["rbw", "lrw"])
def test_weights_follow_the_reference_sequence(strategy):
    from pinns_rl_pde_b200.training import AdaptiveLossWeights
    z = np.load(GOLD)
    seq, want = torch.from_numpy(z[strategy + "_in"]), torch.from_numpy(z[strategy + "_w"])
    port = ref_port.AdaptiveLossWeightsPort(strategy, 0.9, 1e-5, [0.5, 0.3, 0.2])
    mine = AdaptiveLossWeights(strategy, 0.9, 1e-5, [0.5, 0.3, 0.2])
    for v, w in zip(seq, want):
        kw = {"losses": v} if strategy == "rbw" else {"gradients": v}
        assert torch.equal(port.update(**kw), w)
        assert torch.equal(mine.update(**kw), w)
    with pytest.raises(ValueError):
        mine.update(**({"gradients": seq[0]} if strategy == "rbw" else {"losses": seq[0]}))


def test_config_guards():
    import pinns_rl_pde_b200 as pk
    with pytest.raises(ValueError):
        pk.AdaptiveLossWeights("softadapt")
    cfg = pk.TrainingConfig(adaptive_weights=pk.AdaptiveWeightsConfig(enabled=True))
    assert cfg.adaptive_weights.initial_weights == [0.5, 0.3, 0.2]


@pytest.mark.gpu
@pytest.mark.parametrize("strategy,fused", [("lrw", False), ("lrw", True), ("rbw", True)])
def test_gpu_reweighted_step(strategy, fused):
    import pinns_rl_pde_b200 as pk
    from helpers import product_pde
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    m1 = pk.make_model("feedforward", 2, 128, 3, dev)
    m2 = copy.deepcopy(m1)
    pde1, pde2 = product_pde("burgers", dev), product_pde("burgers", dev)
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=0.0, gradient_clipping=1.0, scheduler="none",
                            adaptive_weights=pk.AdaptiveWeightsConfig(enabled=True, strategy=strategy))
    tr = pk.PDETrainer(m1, pde1, config=cfg, device=dev, fused=fused)
    port = ref_port.AdaptiveLossWeightsPort(strategy, 0.9, 1e-5, [0.5, 0.3, 0.2])
    opt = torch.optim.Adam(m2.parameters(), lr=1e-3)
    params = list(m2.parameters())
    g = torch.Generator().manual_seed(4)
    for it in range(4):
        x = (torch.rand(1200, 1, generator=g) * 2 - 1).to(dev)
        t = torch.rand(1200, 1, generator=g).to(dev)
        got = tr.train_step(x, t)
        # the reference's step (trainer.py:578-694) spelled out on the autograd route
        opt.zero_grad(set_to_none=True)
        losses = pde2.compute_loss(m2, x, t)
        names = ("residual", "boundary", "initial")
        comps = torch.stack([losses[k] for k in names]).detach()
        grads = [torch.cat([gr.reshape(-1) for gr in torch.autograd.grad(losses[k], params, retain_graph=True)]) for k in names]
        if strategy == "lrw":
            w = port.update(gradients=torch.stack([gk.norm() for gk in grads]))
        else:
            w = port.update(losses=comps)
        flat = sum(wk * gk for wk, gk in zip(w, grads))
        if strategy == "lrw":
            flat = flat + grads[2]            # trainer.py:611-622,689: the last component's gradient is still in .grad
        off = 0
        for p in params:
            p.grad = flat[off:off + p.numel()].view_as(p).clone()
            off += p.numel()
        torch.nn.utils.clip_grad_norm_(params, 1.0)
        opt.step()
        want_total = float((w * comps).sum())
        assert abs(float(got["total"]) - want_total) <= 2e-5 * abs(want_total), (it, float(got["total"]), want_total)
        assert torch.allclose(tr.history["loss_weights"][-1].cpu(), w.detach().cpu().to(torch.float32), rtol=1e-4, atol=1e-7)
    p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    assert float((p1 - p2).norm()) <= 1e-3 * float(p2.norm())

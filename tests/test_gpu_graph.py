"""CUDA-graph replay of the fused trainer step (PDETrainer(fused=True, graph=True)) against the same step launched
eagerly: same batches, same initial weights -> same losses and parameters (the only arithmetic difference is powf of
the Adam bias correction evaluated on the device instead of the host)."""
import copy

import pytest
import torch

from helpers import product_pde

pytestmark = pytest.mark.gpu


def _trainers(pde_name, arch, layers, extra, dev):
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(0)
    m_a = pk.make_model(arch, 2, 128, layers, dev, **extra)
    m_b = copy.deepcopy(m_a)
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=1e-4, gradient_clipping=1.0, scheduler="none")
    t_a = pk.PDETrainer(m_a, product_pde(pde_name, dev), config=cfg, device=dev, fused=True)
    t_b = pk.PDETrainer(m_b, product_pde(pde_name, dev), config=cfg, device=dev, fused=True, graph=True)
    return m_a, m_b, t_a, t_b


@pytest.mark.parametrize("pde_name,arch,layers,extra", [
    ("burgers", "feedforward", 4, {}),
    ("heat", "fourier", 4, {"mapping_size": 32, "scale": 10.0}),
])
def test_graph_replay_matches_eager_steps(pde_name, arch, layers, extra):
    dev = torch.device("cuda:0")
    m_a, m_b, t_a, t_b = _trainers(pde_name, arch, layers, extra, dev)
    g = torch.Generator().manual_seed(3)
    shapes = [2025] * 7 + [777] * 4 + [2025] * 2 + [5000] * 4 + [2025] * 2   # other shapes capture their own graphs (5000 regrows
    # the engines: the first graph must keep its workspaces alive)
    for i, n in enumerate(shapes):
        x = torch.rand(n, 1, generator=g).to(dev)
        t = torch.rand(n, 1, generator=g).to(dev)
        if i == 9:                                         # learning-rate change between replays
            t_a.optimizer.param_groups[0]["lr"] = 5e-4
            t_b.optimizer.param_groups[0]["lr"] = 5e-4
        l_a = t_a.train_step(x, t)
        l_b = t_b.train_step(x, t)
        for k in ("residual", "boundary", "initial", "total"):
            a, b = float(l_a[k]), float(l_b[k])
            assert abs(a - b) <= 2e-5 * max(abs(a), 1e-12), (i, k, a, b)
    assert sum(1 for v in t_b._graphs.values() if v["graph"] is not None) == 3
    assert t_a.optimizer.step_count == t_b.optimizer.step_count == len(shapes)
    assert float(t_b.optimizer._dyn[0]) == len(shapes)
    # norm-wise: Adam turns the round-off of an (analytically) zero gradient element into +-lr steps, and the wgrad
    # reductions are atomic, so single elements legitimately differ between ANY two runs
    for p_a, p_b in zip(m_a.parameters(), m_b.parameters()):
        assert float((p_a - p_b).norm()) <= 5e-3 * float(p_a.norm()) + 1e-4, (float((p_a - p_b).norm()), float(p_a.norm()))
    assert float((t_a._flat - t_b._flat).norm()) <= 2e-2 * float(t_a._flat.norm())


def test_graph_needs_fused():
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    model = pk.make_model("feedforward", 2, 128, 2, dev)
    with pytest.raises(ValueError):
        pk.PDETrainer(model, product_pde("burgers", dev), config=pk.TrainingConfig(), device=dev, graph=True)


def test_validation_loss_and_live_snapshot(tmp_path):
    """trainer.py:140-162 / :171-279 through forward-only passes: same numbers as the differentiable calls."""
    import numpy as np
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import _lib
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 128, 4, dev)
    pde = product_pde("burgers", dev)
    tr = pk.PDETrainer(model, pde, config=pk.TrainingConfig(), device=dev)
    torch.manual_seed(5)
    val = tr._compute_validation_loss(1000)
    torch.manual_seed(5)
    x, t = pde.generate_collocation_points(1000)
    ref = pde.compute_loss(model, x.to(dev), t.to(dev))
    for k_val, k_ref in (("total_loss", "total"), ("residual_loss", "residual"), ("boundary_loss", "boundary"),
                         ("initial_loss", "initial")):
        assert abs(val[k_val] - float(ref[k_ref])) <= 1e-6 * abs(float(ref[k_ref])), (k_val, val[k_val], float(ref[k_ref]))
    assert model.training
    before = _lib.launch_count()
    tr._save_live_snapshot(str(tmp_path), epoch=7, grid_size=20)
    assert _lib.launch_count() > before
    snap = np.load(tmp_path / "live_snapshot.npz")
    assert set(snap.files) == {"axis_x", "axis_y", "u_pred", "residual", "epoch", "dimension", "x_label", "y_label", "fixed_t"}
    assert snap["u_pred"].shape == (20, 20) and snap["residual"].shape == (20, 20) and int(snap["epoch"]) == 7
    xx, tt = np.meshgrid(snap["axis_x"], snap["axis_y"], indexing="xy")
    xg = torch.tensor(xx.reshape(-1, 1), device=dev)
    tg = torch.tensor(tt.reshape(-1, 1), device=dev)
    r = pde.compute_residual(model, xg, tg).detach().cpu().numpy().reshape(20, 20)
    u = model(torch.cat([xg, tg], dim=1)).detach().cpu().numpy().reshape(20, 20)
    assert np.allclose(snap["residual"], r, rtol=1e-5, atol=1e-6)
    assert np.allclose(snap["u_pred"], u, rtol=1e-5, atol=1e-6)

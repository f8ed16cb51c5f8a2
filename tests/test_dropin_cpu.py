"""CPU: patch_reference() swaps the hot path into the reference's own classes (the unmodified reference installed to
baseline/_ref by oracle/install_reference.py); with no GPU the patched methods must raise -- never fall back."""
import os
import sys

import pytest
import torch

from oracle import ref_env

needs_ref = pytest.mark.skipif(not ref_env.available(), reason="baseline/_ref not installed")


@needs_ref
def test_patch_and_unpatch_reference():
    ref_env.activate()
    try:
        import pinnrl.pdes.burgers_equation as be
        from pinnrl.config import Config, ModelConfig
        from pinnrl.neural_networks import PINNModel
        from pinnrl.pdes.pde_base import PDEConfig
        import pinns_rl_pde_b200 as pk
        from pinns_rl_pde_b200 import _lib, dropin
        from pinns_rl_pde_b200.program import compile_network
        orig = be.BurgersEquation.compute_residual
        names = pk.patch_reference()
        assert "BurgersEquation" in names and be.BurgersEquation.compute_residual is not orig
        c = Config.__new__(Config)
        c.device = torch.device("cpu")
        c.model = ModelConfig(2, 32, 1, 3, "tanh", architecture="feedforward")
        model = PINNModel(config=c, device=torch.device("cpu"))
        prog = compile_network(model)                      # the reference's own module lowers to the op program
        assert len(prog.ops) == 7 and prog.grad_floats == sum(p.numel() for p in model.parameters())
        pde = be.BurgersEquation(config=PDEConfig(name="burgers", domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0],
                                                  parameters={"nu": 0.01}, boundary_conditions={"dirichlet": {"value": 0.0}},
                                                  initial_condition={"type": "sine"}, exact_solution={}, dimension=1,
                                                  device=torch.device("cpu")))
        with pytest.raises(_lib.PinnkError):               # CPU tensors: loud failure, no fallback
            pde.compute_residual(model, torch.zeros(4, 1), torch.zeros(4, 1))
        dropin.unpatch_reference()
        assert be.BurgersEquation.compute_residual is orig
        r = pde.compute_residual(model, torch.zeros(4, 1), torch.zeros(4, 1))      # the reference path again
        assert r.shape == (4, 1)
    finally:
        pass


@needs_ref
def test_patched_select_action_keeps_the_agents_forward_for_unsupported_q_networks():
    """dropin: RLAgent.select_action after patch_reference() must not raise for policy networks pinnk_dqn_forward does
    not cover -- it falls back to the reference's own method, here on a CPU agent and, via the
    UnsupportedQNetwork probe, for any non-DQNNetwork-shaped policy net."""
    ref_env.activate()
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import dropin, rl
    from pinnrl.rl.rl_agent import RLAgent
    agent = RLAgent(state_dim=2, action_dim=1, hidden_dim=512, epsilon_start=0.0, epsilon_end=0.0, device=torch.device("cpu"))
    pk.patch_reference()
    try:
        out = agent.select_action(torch.rand(50, 2))
        assert out.shape == (1, 50)
        with pytest.raises(rl.UnsupportedQNetwork):
            rl._lower(torch.nn.Linear(2, 1))
    finally:
        dropin.unpatch_reference()


@needs_ref
def test_reference_trainer_harness_runs_the_unmodified_trainer_on_cpu():
    """tests/ref_trainer_harness.py (the stock arm of the GPU trajectory test): 2 epochs of pinnrl's own PDETrainer.train
    in fp32 and fp64 from the same seed agree to fp32 round-off, i.e. both arms see the same points and weights."""
    import ref_trainer_harness as H
    a = H.run("c1_heat_fourier", "cpu", 2, batch_size=128, num_points=300)
    b = H.run("c1_heat_fourier", "cpu", 2, batch_size=128, num_points=300, dtype=torch.float64)
    assert len(a["train_loss"]) == 2 and len(a["val_loss"]) == 1
    assert max(H.deviation(a["train_loss"], b["train_loss"]).values()) < 1e-5


def test_cached_rows_are_keyed_on_what_they_depend_on():
    """functional._cached_rows / _rows_key (host logic of the small-batch path): boundary rows are rebuilt when the boundary
    functions, the initial condition, the domain or the requested sizes change, and reused otherwise."""
    import types
    from pinns_rl_pde_b200 import functional as F
    calls = []

    def build():
        calls.append(1)
        return torch.zeros(3)

    bc = {"dirichlet": lambda x, t: x * 0}
    pde = types.SimpleNamespace(domain=[[-1.0, 1.0]], time_domain=[0.0, 1.0], boundary_conditions=bc, compat="reference",
                                config=types.SimpleNamespace(initial_condition={"type": "sine", "amplitude": 1.0, "frequency": 1.0}))
    dev = torch.device("cpu")
    a = F._cached_rows(pde, F._rows_key(pde, dev, "1d"), build)
    b = F._cached_rows(pde, F._rows_key(pde, dev, "1d"), build)
    assert a is b and len(calls) == 1
    pde.config.initial_condition["amplitude"] = 2.0                       # mutated in place: new key
    F._cached_rows(pde, F._rows_key(pde, dev, "1d"), build)
    assert len(calls) == 2
    pde.boundary_conditions["dirichlet"] = lambda x, t: x * 0 + 1         # replaced function: new key
    F._cached_rows(pde, F._rows_key(pde, dev, "1d"), build)
    assert len(calls) == 3
    F._cached_rows(pde, F._rows_key(pde, dev, "heat1d", 50, 100), build)  # sizes are part of the key
    F._cached_rows(pde, F._rows_key(pde, dev, "heat1d", 60, 100), build)
    assert len(calls) == 5
    for i in range(12):                                                   # bounded
        F._cached_rows(pde, F._rows_key(pde, dev, "n", i), build)
    assert len(pde._pinnk_rows_cache) <= 8


def test_graph_trainer_needs_the_fused_step():
    import pinns_rl_pde_b200 as pk
    model = torch.nn.Linear(2, 1)
    with pytest.raises(ValueError):
        pk.PDETrainer(model, object.__new__(pk.BurgersEquation), config=pk.TrainingConfig(), device=torch.device("cpu"), graph=True)


def test_live_snapshot_layout_matches_the_reference_file(monkeypatch):
    """trainer.py:171-279: the mirror's PDETrainer.live_snapshot (grid construction, ordering, keys) executed on the CPU with
    the network / residual evaluations routed to the oracle, against the npz the UNMODIFIED reference wrote
    (tests/golden/x_live_snapshot.npz, made by make_golden.py snapshot).  The libpinnk evaluations themselves are covered by
    tests/test_gpu_graph.py::test_validation_loss_and_live_snapshot."""
    import types
    import numpy as np
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import training
    from oracle import ref_port
    from helpers import GOLDEN, PDES
    z = np.load(os.path.join(GOLDEN, "x_live_snapshot.npz"))
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    want = {k[6:]: z[k] for k in z.files if k.startswith("snap::")}
    model = ref_port.PINNModel("feedforward", 2, 32, 3)
    model.load_state_dict(state)
    s = PDES["burgers"]
    pde = types.SimpleNamespace(dimension=1, domain=s["domain"], time_domain=s["time"],
                                compute_residual=lambda m, x, t: ref_port.burgers_residual(m, x, t, nu=s["params"]["nu"]).detach())
    monkeypatch.setattr(training.F, "model_forward", lambda m, xt: m(xt))
    tr = object.__new__(pk.PDETrainer)
    tr.model, tr.pde, tr.device = model, pde, torch.device("cpu")
    snap = tr.live_snapshot(epoch=3, grid_size=12)
    assert set(snap) == set(want)
    for k in ("axis_x", "axis_y"):
        assert np.array_equal(snap[k], want[k])
    assert int(snap["epoch"]) == int(want["epoch"]) and int(snap["dimension"]) == int(want["dimension"])
    assert str(snap["x_label"]) == str(want["x_label"]) and str(snap["y_label"]) == str(want["y_label"])
    assert np.isnan(snap["fixed_t"]) and np.isnan(want["fixed_t"])
    assert np.allclose(snap["u_pred"], want["u_pred"], rtol=0, atol=1e-6)
    # the oracle residual needs grad-enabled inputs: live_snapshot runs it under no_grad in the product (forward-mode jets), so
    # here the comparison is made with the oracle called directly on the snapshot's grid, in the snapshot's ordering
    xx, tt = np.meshgrid(snap["axis_x"], snap["axis_y"], indexing="xy")
    r = ref_port.burgers_residual(model, torch.tensor(xx.reshape(-1, 1)), torch.tensor(tt.reshape(-1, 1)), nu=s["params"]["nu"])
    assert np.allclose(r.detach().numpy().reshape(12, 12), want["residual"], rtol=1e-5, atol=1e-6)

"""Small-batch compute_loss + backward (functional._LazyLossFn: one forward-only sweep over all row sets, one reverse sweep
seeded with the component weights) against the eager per-component route (PINNK_EAGER_LOSS_GRAD=1) of the same library:
total gradient, per-component backward with retain_graph (the reference's LRW re-weighting, trainer.py:611-622), a backward
after the engine was used for something else, and the in-place-modification guard."""
import pytest
import torch

from helpers import flat_grad, product_pde, rel

pytestmark = pytest.mark.gpu


def _setup(pde_name, arch, extra):
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = pk.make_model(arch, 2, 128, 3, dev, **extra)
    pde = product_pde(pde_name, dev)
    g = torch.Generator().manual_seed(5)
    x = (torch.rand(1500, 1, generator=g) * 2 - 1).to(dev) if pde_name != "heat" else torch.rand(1500, 1, generator=g).to(dev)
    t = torch.rand(1500, 1, generator=g).to(dev)
    return model, pde, x, t


@pytest.mark.parametrize("pde_name,arch,extra", [("burgers", "feedforward", {}), ("heat", "fourier", {"mapping_size": 32, "scale": 10.0}),
                                                ("kdv", "resnet", {"num_blocks": 2})])
def test_lazy_route_matches_eager_route(pde_name, arch, extra, monkeypatch):
    from pinns_rl_pde_b200 import _lib
    model, pde, x, t = _setup(pde_name, arch, extra)
    monkeypatch.setenv("PINNK_EAGER_LOSS_GRAD", "1")
    model.zero_grad(set_to_none=True)
    before = _lib.launch_count()
    l_e = pde.compute_loss(model, x, t)
    l_e["total"].backward()
    eager_launches = _lib.launch_count() - before
    g_e = flat_grad(model).clone()
    per_e = []
    for k in ("residual", "boundary", "initial"):
        model.zero_grad(set_to_none=True)
        pde.compute_loss(model, x, t)[k].backward()
        per_e.append(flat_grad(model).clone())
    monkeypatch.setenv("PINNK_EAGER_LOSS_GRAD", "0")
    model.zero_grad(set_to_none=True)
    before = _lib.launch_count()
    l_l = pde.compute_loss(model, x, t)
    l_l["total"].backward()
    lazy_launches = _lib.launch_count() - before
    g_l = flat_grad(model).clone()
    for k in ("residual", "boundary", "initial", "total"):
        assert abs(float(l_l[k]) - float(l_e[k])) <= 2e-6 * abs(float(l_e[k])) + 1e-12, k
    assert rel(g_l, g_e) <= 5e-6
    # one component at a time out of ONE forward (retain_graph): first backward reuses the stash, the others recompute
    model.zero_grad(set_to_none=True)
    losses = pde.compute_loss(model, x, t)
    for i, k in enumerate(("residual", "boundary", "initial")):
        model.zero_grad(set_to_none=True)
        losses[k].backward(retain_graph=True)
        assert rel(flat_grad(model), per_e[i]) <= 5e-6, k
    assert lazy_launches < 0.7 * eager_launches, (lazy_launches, eager_launches)


def test_backward_after_other_use_of_the_engine_and_inplace_guard():
    model, pde, x, t = _setup("burgers", "feedforward", {})
    model.zero_grad(set_to_none=True)
    pde.compute_loss(model, x, t)["total"].backward()
    want = flat_grad(model).clone()
    model.zero_grad(set_to_none=True)
    losses = pde.compute_loss(model, x, t)
    with torch.no_grad():                                   # validation-style calls in between overwrite the workspace
        pde.compute_loss(model, x[:700], t[:700])
        pde.compute_residual(model, x, t)
    losses["total"].backward()
    assert rel(flat_grad(model), want) <= 5e-6
    losses = pde.compute_loss(model, x, t)
    with torch.no_grad():
        next(model.parameters()).mul_(1.0)
    with pytest.raises(RuntimeError):
        losses["total"].backward()


def test_lbfgs_closure_step_follows_the_reference_algorithm():
    """trainer.py:373-389 (_lbfgs_step): torch.optim.LBFGS re-enters compute_loss + backward through its closure.  Same
    start, same rows, same settings as the oracle (reference algorithm, fp64, CPU): the loss goes down and the two
    trajectories stay together over two steps of five strong-Wolfe iterations."""
    import math
    import pinns_rl_pde_b200 as pk
    from oracle import ref_port
    from helpers import PDES
    dev = torch.device("cuda:0")
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 128, 3, dev)
    pde = product_pde("burgers", dev)
    g = torch.Generator().manual_seed(9)
    x = torch.rand(600, 1, generator=g) * 2 - 1
    t = torch.rand(600, 1, generator=g)
    cfg = pk.TrainingConfig(optimizer="lbfgs", learning_rate=1.0, lbfgs=pk.LBFGSConfig(history_size=10, max_iter=5))
    trainer = pk.PDETrainer(model, pde, config=cfg, device=dev)
    # oracle
    om = ref_port.PINNModel("feedforward", 2, 128, 3)
    om.load_state_dict({k: v.detach().cpu() for k, v in model.state_dict().items()})
    om = om.double()
    s = PDES["burgers"]
    fns = ref_port.boundary_condition_fns("burgers", s["bcs"], s["ic"], [tuple(d) for d in s["domain"]], s["params"])
    opt = torch.optim.LBFGS(om.parameters(), lr=1.0, history_size=10, max_iter=5, line_search_fn="strong_wolfe",
                            tolerance_grad=1e-7, tolerance_change=1e-9)
    xd, td = x.double(), t.double()

    def oracle_total():
        r = ref_port.burgers_residual(om, xd, td, nu=s["params"]["nu"])
        return ref_port.base_compute_loss(om, r, [tuple(d) for d in s["domain"]], tuple(s["time"]), fns)["total"]

    def closure():
        opt.zero_grad()
        total = oracle_total()
        total.backward()
        return total

    start = float(pde.compute_loss(model, x.to(dev), t.to(dev))["total"])
    assert abs(start - float(oracle_total())) <= 1e-5 * abs(start)
    got, want = [], []
    for _ in range(2):
        trainer.train_step(x.to(dev), t.to(dev))
        with torch.no_grad():
            got.append(float(pde.compute_loss(model, x.to(dev), t.to(dev))["total"]))
        opt.step(closure)
        want.append(float(oracle_total()))
    assert got[0] < start and got[1] < got[0], (start, got)
    for a, b in zip(got, want):
        assert abs(a - b) <= 5e-3 * abs(b), (start, got, want)

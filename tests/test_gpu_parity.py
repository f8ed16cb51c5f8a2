"""GPU parity: the CUDA path (through the C ABI) against
  (1) the committed golden fixtures generated from the UNMODIFIED reference, and
  (2) the oracle run on the same seeded inputs at sizes it finishes in seconds.

Tolerance (north_star: 1e-5 relative in fp32): errors are norm-wise and judged against the
reference evaluated in fp64:  err(cuda, ref64) <= max(1e-5, 2 * err(ref32, ref64)).
For LayerNorm nets the fp64 target is the corrected oracle.
"""
import math

import numpy as np
import pytest
import torch

import parity_log

from helpers import PDES, fixtures, flat_grad, load_fixture, port_model, product_model, product_pde, rel

pytestmark = pytest.mark.gpu
TOL = 1e-5


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from pinns_rl_pde_b200 import build
    build.build()
    return torch.device("cuda:0")


def _tol(z, what):
    floor = {"residual": rel(z["residual32"], z["residual64"]), "grad": rel(z["grad32"], z["grad64"])}[what]
    return max(TOL, 2 * floor)


def gate(label, got, want64, floor, floor_name="ref32"):
    """err(cuda, ref64) <= max(1e-5, 2 * err(ref32, ref64)), norm-wise.  ``floor`` is the fp32 reference's own
    error against the fp64 target (a number, or the fp32 tensor).  Every measured error is printed so the GPU test log
    shows how far inside the gate the kernels are."""
    if not isinstance(floor, float):
        floor = rel(floor, want64)
    err = rel(got, want64)
    tol = max(TOL, 2.0 * floor)
    parity_log.log(f"[parity] {label}: cuda-vs-ref64 {err:.3e} | {floor_name}-vs-ref64 {floor:.3e} | gate {tol:.3e}")
    assert err <= tol, (label, err, floor, tol)
    return err


@pytest.mark.parametrize("tag", fixtures())
def test_residual_and_gradient_vs_golden(dev, tag):
    z, meta, state = load_fixture(tag)
    model = product_model(meta, state, dev)
    pde = product_pde(meta["pde"], dev, meta["dimension"])
    x, t = torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["t"]).to(dev)
    r = pde.compute_residual(model, x, t)
    assert r.shape == (meta["n"], 1) and r.requires_grad
    has_ln = meta["arch"] == "resnet"
    want_r = z["residual64_corrected"] if has_ln else z["residual64"]
    gate(f"golden {tag} residual", r.detach().cpu(), want_r, rel(z["residual32"], z["residual64"]))
    (r ** 2).mean().backward()
    want_g = z["grad64_mse_corrected"] if has_ln else z["grad64_mse"]
    gate(f"golden {tag} grad(mean r^2)", flat_grad(model).cpu(), want_g, rel(z["grad32"], z["grad64"]))


@pytest.mark.parametrize("tag", [f for f in fixtures() if f.startswith(("c1_", "c2_", "x_"))])
def test_compute_loss_vs_golden(dev, tag):
    z, meta, state = load_fixture(tag)
    if meta["mode"] != "loss":
        pytest.skip("fixture stores residual only")
    model = product_model(meta, state, dev)
    pde = product_pde(meta["pde"], dev)
    x, t = torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["t"]).to(dev)
    losses = pde.compute_loss(model, x, t)
    assert set(losses) == {"residual", "boundary", "initial", "smoothness", "data", "total"}
    for k in ("residual", "boundary", "initial", "total"):
        want = float(z[f"loss64_{k}"])
        assert abs(losses[k].item() - want) <= max(TOL, 2 * abs(float(z[f"loss32_{k}"]) - want) / abs(want)) * abs(want), k
    losses["total"].backward()
    gate(f"golden {tag} grad(total loss)", flat_grad(model).cpu(), z["grad64"], rel(z["grad32"], z["grad64"]))


def test_2d_cahn_hilliard_math_operator(dev):
    """compat='math': the intended 2-D operator (18 jet columns) vs the corrected multi-dim autograd oracle."""
    z, meta, state = load_fixture("c4_ch2d_siren_small")
    model = product_model(meta, state, dev)
    pde = product_pde("cahn_hilliard", dev, 2, compat="math")
    x, t = torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["t"]).to(dev)
    r = pde.compute_residual(model, x, t)
    # fp32 noise floor of this operator: the independent Taylor-mode oracle evaluated in fp32 on the CPU (omega_0 = 30
    # amplifies each derivative order; the corrected multi-dim autograd oracle is what the fixture's fp64 target is)
    from oracle import jets_oracle
    m32 = port_model(meta, state, torch.float32)
    r32 = jets_oracle.residual(m32, "cahn_hilliard", torch.from_numpy(z["x"]), torch.from_numpy(z["t"]), meta["params"], 2, "math")
    g32 = jets_oracle.flat_grad(m32, (r32 ** 2).mean())
    gate("c4 2-D Cahn-Hilliard math residual", r.detach().cpu(), z["residual64_math"], r32.detach(), "jets32")
    (r ** 2).mean().backward()
    gate("c4 2-D Cahn-Hilliard math grad", flat_grad(model).cpu(), z["grad64_mse_math"], g32.detach(), "jets32")


def test_heat_math_operator(dev):
    from oracle import jets_oracle
    z, meta, state = load_fixture("c1_heat_fourier")
    model = product_model(meta, state, dev)
    pde = product_pde("heat", dev, compat="math")
    x, t = torch.from_numpy(z["x"]).to(dev), torch.from_numpy(z["t"]).to(dev)
    r = pde.compute_residual(model, x, t)
    m64 = port_model(meta, state, torch.float64)
    want = jets_oracle.residual(m64, "heat", torch.from_numpy(z["x"]).double(), torch.from_numpy(z["t"]).double(),
                                meta["params"], 1, "math")
    m32 = port_model(meta, state, torch.float32)
    r32 = jets_oracle.residual(m32, "heat", torch.from_numpy(z["x"]), torch.from_numpy(z["t"]), meta["params"], 1, "math")
    gate("c1 heat math residual", r.detach().cpu(), want.detach(), r32.detach(), "jets32")


FULL = [  # (pde, arch, hidden, layers, dimension, n, extra)  -- BASELINE configs at full network size
    ("heat", "fourier", 128, 4, 1, 1500, {"mapping_size": 32, "scale": 10.0}),
    ("burgers", "feedforward", 128, 8, 1, 3000, {}),
    ("kdv", "resnet", 256, 6, 1, 700, {"num_blocks": 6}),
    ("cahn_hilliard", "siren", 256, 5, 2, 1500, {"omega_0": 30.0}),
    ("cahn_hilliard", "siren", 256, 5, 1, 700, {"omega_0": 30.0}),
    ("allen_cahn", "feedforward", 128, 8, 1, 3000, {}),
    # the residual network the reference's YAML ships (config.yaml:15-19: hidden_dim 512; fewer blocks to keep the CPU oracle short)
    ("burgers", "resnet", 512, 2, 1, 600, {"num_blocks": 2}),
    # the Fourier network of the YAML (config.yaml:25-31: mapping_size 512 -> 1024 features, hidden 512, scale 4; fewer layers)
    ("heat", "fourier", 512, 3, 1, 500, {"mapping_size": 512, "scale": 4.0}),
]


@pytest.mark.parametrize("pde_name,arch,hidden,layers,dim,n,extra", FULL)
def test_full_size_configs_vs_oracle(dev, monkeypatch, pde_name, arch, hidden, layers, dim, n, extra):
    """Seeded weights/points; oracle = the reference's autograd algorithm (ref_port) in fp64 on the CPU
    (with the primitive LayerNorm where the reference's nn.LayerNorm is inexact).  The weight gradients use the
    fixed-order reduction (PINNK_DETERMINISTIC=1), so the measured error is the kernels' and not the arrival order's."""
    import pinns_rl_pde_b200 as pk
    from oracle import ref_port
    monkeypatch.setenv("PINNK_DETERMINISTIC", "1")
    torch.manual_seed(3)
    model = pk.make_model(arch, dim + 1, hidden, layers, dev, **extra)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    meta = dict(arch=arch, hidden=hidden, layers=layers, dimension=dim, extra=extra)
    m64 = port_model(meta, state, torch.float64, corrected=(arch == "resnet"))
    m32 = port_model(meta, state, torch.float32, corrected=(arch == "resnet"))
    s = PDES[pde_name]
    g = torch.Generator().manual_seed(4)
    lo, hi = s["domain"][0]
    x = torch.rand(n, dim, generator=g) * (hi - lo) + lo
    t = torch.rand(n, 1, generator=g) * (s["time"][1] - s["time"][0]) + s["time"][0]
    pde = product_pde(pde_name, dev, dim)
    r = pde.compute_residual(model, x.to(dev), t.to(dev))
    kw = {k: v for k, v in s["params"].items() if k != "speed"}
    want = ref_port.RESIDUALS[pde_name](m64, x.double(), t.double(), dimension=dim, **kw)
    ref32 = ref_port.RESIDUALS[pde_name](m32, x.clone(), t.clone(), dimension=dim, **kw)     # the reference algorithm in fp32
    label = f"full {pde_name}/{arch} {layers}x{hidden} dim {dim} n {n}"
    gate(label + " residual", r.detach().cpu(), want.detach(), ref32.detach())
    (r ** 2).mean().backward()
    (want ** 2).mean().backward()
    (ref32 ** 2).mean().backward()
    cat = lambda m: torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in m.parameters()])
    gate(label + " grad", flat_grad(model).cpu(), cat(m64), cat(m32))


def test_chunking_is_invisible(dev):
    """Rows beyond the plan's chunk are processed in several passes with identical results."""
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import engine
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 64, 3, dev)
    pde = product_pde("burgers", dev)
    g = torch.Generator().manual_seed(1)
    x, t = (torch.rand(5000, 1, generator=g) * 2 - 1).to(dev), torch.rand(5000, 1, generator=g).to(dev)
    l1 = pde.compute_loss(model, x, t)
    l1["total"].backward()
    g1 = flat_grad(model).clone()
    model.zero_grad()
    old = engine.MAX_CHUNK_POINTS
    engine.MAX_CHUNK_POINTS = 1024
    engine._CACHE.clear()
    try:
        l2 = pde.compute_loss(model, x, t)
        l2["total"].backward()
    finally:
        engine.MAX_CHUNK_POINTS = old
        engine._CACHE.clear()
    assert abs(l1["total"].item() - l2["total"].item()) <= 1e-6 * abs(l1["total"].item())
    assert rel(flat_grad(model), g1) <= 2e-6


@pytest.mark.parametrize("arch,extra,pde_name", [("feedforward", {}, "burgers"), ("resnet", {"num_blocks": 2}, "kdv"),
                                                 ("siren", {"omega_0": 30.0}, "cahn_hilliard")])
def test_residual_backward_reuses_the_forward_stash(dev, arch, extra, pde_name):
    """r = compute_residual(...); loss(r).backward(): when the rows fit one chunk the reverse pass starts from the stash the
    forward left in the workspace (pinnk_loss_step_flags KEEP_STASH / REUSE_STASH) and must give the gradient of the route
    that recomputes the forward -- which is what runs when anything touched the engine, the rows or the parameters."""
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import engine, functional as F
    torch.manual_seed(3)
    model = pk.make_model(arch, 2, 128, 3, dev, **extra)
    pde = product_pde(pde_name, dev)
    g = torch.Generator().manual_seed(5)
    x, t = (torch.rand(3000, 1, generator=g) * 2 - 1).to(dev), torch.rand(3000, 1, generator=g).to(dev)
    eng = engine.get_engine(model, F.residual_spec(pde)[0], 3000)

    def grad(disturb):
        model.zero_grad()
        r = pde.compute_residual(model, x, t)
        token = eng._stash_token
        assert token is not None                       # the forward kept its stash
        if disturb == "engine":                        # another call on the engine overwrites the workspace
            pde.score_residual(model, x[:100], t[:100])
            assert eng._stash_token is None
        elif disturb == "params":                      # an in-place parameter update bumps the version counter
            with torch.no_grad():
                next(model.parameters()).add_(0.0)
        (r ** 2).mean().backward()
        return flat_grad(model).clone(), float((r.detach() ** 2).mean())

    g_reuse, l_reuse = grad(None)
    g_engine, l_engine = grad("engine")
    g_params, _ = grad("params")
    assert l_reuse == l_engine
    assert torch.isfinite(g_reuse).all() and float(g_reuse.norm()) > 0
    assert rel(g_reuse, g_engine) <= 2e-6 and rel(g_params, g_engine) <= 2e-6
    with torch.no_grad():                              # forward-only callers keep nothing
        pde.compute_residual(model, x, t)
    assert eng._stash_token is None


def test_model_forward_and_custom_loss_autograd(dev):
    """model(x) and jets() are differentiable w.r.t. parameters for callers that write their own loss."""
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(0)
    model = pk.make_model("siren", 2, 64, 3, dev, omega_0=30.0)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    m64 = port_model(dict(arch="siren", hidden=64, layers=3, dimension=1, extra={"omega_0": 30.0}), state, torch.float64)
    xt = torch.rand(333, 2, generator=torch.Generator().manual_seed(2))
    u = model(xt.to(dev))
    want = m64(xt.double())
    assert u.shape == (333, 1) and rel(u.detach().cpu(), want.detach()) <= TOL
    (u.sin().sum()).backward()
    want.sin().sum().backward()
    og = torch.cat([p.grad.reshape(-1) for p in m64.parameters()])
    assert rel(flat_grad(model).cpu(), og) <= TOL
    U = pk.jets(model, xt.to(dev), [((1.0, 0.0), 2), ((0.0, 1.0), 1)])
    X = xt.double().requires_grad_(True)
    uu = m64(X)
    du = torch.autograd.grad(uu.sum(), X, create_graph=True)[0]
    uxx = torch.autograd.grad(du[:, 0].sum(), X)[0][:, 0]
    assert rel(U[:, 1].cpu(), du[:, 0].detach()) <= TOL and rel(2 * U[:, 2].cpu(), uxx) <= TOL
    assert rel(U[:, 3].cpu(), du[:, 1].detach()) <= TOL


def test_scoring_matches_residual(dev):
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 64, 4, dev)
    pde = product_pde("allen_cahn", dev)
    g = torch.Generator().manual_seed(5)
    x, t = (torch.rand(4097, 1, generator=g) * 2 - 1).to(dev), torch.rand(4097, 1, generator=g).to(dev)
    r = pde.compute_residual(model, x, t).detach().abs().reshape(-1)
    mag, stats = pde.score_residual(model, x, t)
    assert torch.equal(mag, r)
    s = stats.cpu()
    assert abs(s[0].item() - r.double().sum().item()) <= 1e-9 * r.double().sum().item() + 1e-12
    assert abs(s[1].item() - (r.double() ** 2).sum().item()) <= 1e-9 * (r.double() ** 2).sum().item() + 1e-12
    assert s[2].item() == float(r.max().item()) and s[3].item() == 4097
    xs, ts = pde.generate_collocation_points(1000, strategy="residual_based", model=model)
    assert xs.shape == (1000, 1) and ts.shape == (1000, 1)


def test_edge_cases(dev):
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import _lib
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 32, 2, dev)
    pde = product_pde("burgers", dev)
    # empty batch
    r = pde.compute_residual(model, torch.zeros(0, 1, device=dev), torch.zeros(0, 1, device=dev))
    assert r.shape == (0, 1)
    # single point, ragged size
    for n in (1, 7, 257):
        r = pde.compute_residual(model, torch.rand(n, 1, device=dev), torch.rand(n, 1, device=dev))
        assert r.shape == (n, 1) and torch.isfinite(r).all()
    # wrong dtype / shape fail loudly
    with pytest.raises(_lib.PinnkError):
        model(torch.zeros(4, 3, device=dev))
    # compute_loss in 2-D fails like the reference does
    m3 = pk.make_model("feedforward", 3, 32, 2, dev)
    with pytest.raises(RuntimeError):
        product_pde("burgers", dev, 2).compute_loss(m3, torch.rand(8, 2, device=dev), torch.rand(8, 1, device=dev))
    # frozen parameters receive no gradient and do not break the reverse pass
    for p in list(model.parameters())[:2]:
        p.requires_grad_(False)
    pde.compute_loss(model, torch.rand(64, 1, device=dev), torch.rand(64, 1, device=dev))["total"].backward()
    assert all((p.grad is None) == (not p.requires_grad) for p in model.parameters())


def test_loss_functions_and_training_configs(dev):
    """mae / huber reductions and the three shapes config.training can take (dataclass, dict, None)."""
    import pinns_rl_pde_b200 as pk
    from oracle import ref_port
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 32, 3, dev)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    m64 = port_model(dict(arch="feedforward", hidden=32, layers=3, dimension=1, extra={}), state, torch.float64)
    m32 = port_model(dict(arch="feedforward", hidden=32, layers=3, dimension=1, extra={}), state, torch.float32)
    g = torch.Generator().manual_seed(6)
    x, t = torch.rand(500, 1, generator=g) * 2 - 1, torch.rand(500, 1, generator=g)
    s = PDES["burgers"]
    fns = ref_port.boundary_condition_fns("burgers", s["bcs"], s["ic"], s["domain"], s["params"])
    for fn_name, training, weights in [("mae", {"loss_function": "mae"}, (1.0, 10.0, 10.0)),
                                       ("huber", {"loss_function": "huber", "huber_delta": 0.05}, (1.0, 10.0, 10.0)),
                                       ("mse", pk.TrainingConfig(loss_weights={"residual": 15.0, "boundary": 20.0, "initial": 10.0}),
                                        (15.0, 20.0, 10.0))]:
        pde = product_pde("burgers", dev, training=training)
        model.zero_grad()
        L = pde.compute_loss(model, x.to(dev), t.to(dev))
        L["total"].backward()
        r = ref_port.burgers_residual(m64, x.double(), t.double(), nu=s["params"]["nu"])
        delta = 0.05 if fn_name == "huber" else 1.0
        want = ref_port.base_compute_loss(m64, r, s["domain"], s["time"], fns, weights, fn_name, delta)
        for p in m64.parameters():
            p.grad = None
        want["total"].backward()
        og = torch.cat([p.grad.reshape(-1) for p in m64.parameters()])
        r32 = ref_port.burgers_residual(m32, x.clone(), t.clone(), nu=s["params"]["nu"])
        w32 = ref_port.base_compute_loss(m32, r32, s["domain"], s["time"], fns, weights, fn_name, delta)
        for p in m32.parameters():
            p.grad = None
        w32["total"].backward()
        og32 = torch.cat([p.grad.reshape(-1) for p in m32.parameters()])
        gate(f"{fn_name} total loss", L["total"].detach().cpu().reshape(1), want["total"].detach().reshape(1), w32["total"].detach().reshape(1))
        gate(f"{fn_name} grad(total loss)", flat_grad(model).cpu(), og, og32)


def test_loss_trajectory_500_epochs(dev):
    """north_star: loss trajectories over 500 epochs within 1e-4 relative of the reference algorithm.

    Adam on a fixed seeded batch, identical hyper-parameters.  Three runs: CUDA (fp32), the oracle port in
    fp32 (== the reference, bit for bit) and in fp64, both on the CPU.  The reference ITSELF is chaotic at this
    horizon: its own fp32 and fp64 runs agree to ~3e-7 for ~200 epochs and then drift apart to ~1e-2
    (measured: DESIGN.md "trajectory parity").  So, following SURVEY F9, the criterion is
        dev(cuda, ref64) <= 1e-4                                 while ref32 itself stays within 1e-5 of ref64
        worst dev(cuda, ref64) <= max(1e-4, 2 * worst dev(ref32, ref64))   over all 500 epochs."""
    import pinns_rl_pde_b200 as pk
    from oracle import ref_port
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 32, 3, dev)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    meta = dict(arch="feedforward", hidden=32, layers=3, dimension=1, extra={})
    m64, m32 = port_model(meta, state, torch.float64), port_model(meta, state, torch.float32)
    pde = product_pde("burgers", dev)
    s = PDES["burgers"]
    fns = ref_port.boundary_condition_fns("burgers", s["bcs"], s["ic"], s["domain"], s["params"])
    g = torch.Generator().manual_seed(7)
    x, t = torch.rand(256, 1, generator=g) * 2 - 1, torch.rand(256, 1, generator=g)
    xd, td = x.to(dev), t.to(dev)
    opt_a = torch.optim.Adam(model.parameters(), lr=1e-3)
    opts = {id(m): torch.optim.Adam(m.parameters(), lr=1e-3) for m in (m32, m64)}

    def oracle_step(m, xx, tt):
        opts[id(m)].zero_grad()
        r = ref_port.burgers_residual(m, xx, tt, nu=s["params"]["nu"])
        loss = ref_port.base_compute_loss(m, r, s["domain"], s["time"], fns)["total"]
        loss.backward()
        opts[id(m)].step()
        return loss.item()

    worst_cuda = worst_ref32 = 0.0
    for epoch in range(500):
        opt_a.zero_grad()
        la = pde.compute_loss(model, xd, td)["total"]
        la.backward()
        opt_a.step()
        l64 = oracle_step(m64, x.double(), t.double())
        l32 = oracle_step(m32, x, t)
        d_cuda, d_ref = abs(la.item() - l64) / abs(l64), abs(l32 - l64) / abs(l64)
        worst_cuda, worst_ref32 = max(worst_cuda, d_cuda), max(worst_ref32, d_ref)
        if worst_ref32 <= 1e-5:
            assert d_cuda <= 1e-4, (epoch, d_cuda)
    parity_log.log(f"[trajectory toy 3x32, 256 fixed points] worst dev cuda-vs-ref64 {worst_cuda:.3e}, ref32-vs-ref64 {worst_ref32:.3e}")
    assert worst_cuda <= max(1e-4, 2 * worst_ref32), (worst_cuda, worst_ref32)


def test_full_size_1m_points_properties(dev):
    """BASELINE config 2 at its full size (1 M collocation points, 8x128 tanh): size-independent properties.
      * the residual of a random 2 048-row subsample taken from the 1 M-row call equals the oracle's (fp64, CPU)
      * shard additivity: loss and gradient of the whole batch = row-weighted sum over two halves (what the
        data-parallel all-reduce relies on), and the forward-only scoring statistics agree with the residual."""
    import pinns_rl_pde_b200 as pk
    from oracle import ref_port
    from pinns_rl_pde_b200 import functional as F
    torch.manual_seed(11)
    model = pk.make_model("feedforward", 2, 128, 8, dev)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    m64 = port_model(dict(arch="feedforward", hidden=128, layers=8, dimension=1, extra={}), state, torch.float64)
    pde = product_pde("burgers", dev)
    n = 1 << 20
    g = torch.Generator(device="cuda").manual_seed(12)
    x = torch.rand(n, 1, generator=g, device=dev) * 2 - 1
    t = torch.rand(n, 1, generator=g, device=dev)
    r = pde.compute_residual(model, x, t).detach()
    idx = torch.randint(0, n, (2048,), generator=g, device=dev)
    want = ref_port.burgers_residual(m64, x[idx].cpu().double(), t[idx].cpu().double(), nu=PDES["burgers"]["params"]["nu"])
    assert rel(r[idx], want) <= TOL
    mag, stats = pde.score_residual(model, x, t)
    assert torch.equal(mag, r.abs().reshape(-1))
    assert abs(stats[1].item() - float((r.double() ** 2).sum())) <= 1e-9 * stats[1].item()
    # whole batch
    comp, _ = F.loss_components(pde, model, x, t)
    (gw,) = torch.autograd.grad(comp[0], [p for p in model.parameters()][4:5])          # one 128x128 weight is enough
    h = n // 2
    c1, _ = F.loss_components(pde, model, x[:h], t[:h], n_global=n)
    c2, _ = F.loss_components(pde, model, x[h:], t[h:], n_global=n)
    (g1,) = torch.autograd.grad(c1[0], [p for p in model.parameters()][4:5])
    (g2,) = torch.autograd.grad(c2[0], [p for p in model.parameters()][4:5])
    assert abs(comp[0].item() - 0.5 * (c1[0].item() + c2[0].item())) <= 2e-6 * abs(comp[0].item())
    assert abs(comp[0].item() - float((r.double() ** 2).mean())) <= 2e-6 * abs(comp[0].item())
    # two independent fp32 evaluations of the same gradient: the wgrad partial sums are combined with atomic reductions, so
    # even identical calls differ by 2e-6 .. 4e-6 from run to run (profiles/flaky_probe.py); the gate is the path's 1e-5
    assert rel(0.5 * (g1 + g2), gw) <= 1e-5
    assert abs(comp[1].item() - c1[1].item()) <= 1e-7 * abs(comp[1].item()) + 1e-12       # BC/IC terms do not depend on the shard


def test_fused_trainer_step_matches_autograd_route(dev):
    """PDETrainer(fused=True): loss + weighted gradient in one pass and clip + Adam in libpinnk must follow the same
    trajectory as compute_loss().backward() + torch clip_grad_norm_ + torch.optim.Adam (trainer.py:577-578,689-694)."""
    import copy
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(0)
    m1 = pk.make_model("feedforward", 2, 128, 3, dev)
    m2 = copy.deepcopy(m1)
    pde = product_pde("burgers", dev)
    cfg = pk.TrainingConfig(learning_rate=2e-3, weight_decay=1e-4, gradient_clipping=0.5, scheduler="none")
    t1 = pk.PDETrainer(m1, pde, config=cfg, device=dev, fused=True)
    t2 = pk.PDETrainer(m2, pde, config=cfg, device=dev, fused=False)
    g = torch.Generator().manual_seed(3)
    for it in range(20):
        x, t = (torch.rand(3000, 1, generator=g) * 2 - 1).to(dev), torch.rand(3000, 1, generator=g).to(dev)
        l1 = t1.train_step(x, t)
        l2 = t2.train_step(x, t)
        assert abs(l1["total"].item() - l2["total"].item()) <= 2e-5 * abs(l2["total"].item()), it
    p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    assert rel(p1, p2) <= 1e-5


def test_heat_nd_compute_loss_assembly(dev):
    """HeatEquation.compute_loss, N-D branch (heat_equation.py:446-473,516-535): periodic value matching on random
    face points per axis and a random-point IC term.  The kernels are covered elsewhere; here the assembly (same RNG
    calls in the same order, pair segments, targets) is checked against a plain evaluation through model()."""
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 3, 64, 3, dev)
    pde = product_pde("heat", dev, 2)
    n = 400
    x, t = torch.rand(n, 2, device=dev), torch.rand(n, 1, device=dev)
    torch.manual_seed(123)
    L = pde.compute_loss(model, x, t)
    assert set(L) == {"residual", "boundary", "initial", "smoothness", "data", "total"}
    # replay the reference's point generation
    torch.manual_seed(123)
    nb, ni, dim = max(n // 10, 10), max(n // 5, 10), 2
    per_axis = max(nb // (2 * dim), 1)
    bl = torch.zeros((), device=dev)
    with torch.no_grad():
        for axis in range(dim):
            free = torch.empty(per_axis, dim, device=dev)
            for d in range(dim):
                free[:, d] = torch.rand(per_axis, device=dev)
            ta = torch.rand(per_axis, 1, device=dev)
            lo, hi = free.clone(), free.clone()
            lo[:, axis], hi[:, axis] = 0.0, 1.0
            bl = bl + ((model(torch.cat([lo, ta], 1)) - model(torch.cat([hi, ta], 1))) ** 2).mean()
        xi = torch.empty(ni, dim, device=dev)
        for d in range(dim):
            xi[:, d] = torch.rand(ni, device=dev)
        ti = torch.zeros(ni, 1, device=dev)
        il = ((model(torch.cat([xi, ti], 1)) - pde.boundary_conditions["initial"](xi, ti)) ** 2).mean()
        r = pde.compute_residual(model, x, t)
    assert abs(L["boundary"].item() - bl.item()) <= 1e-5 * abs(bl.item())
    assert abs(L["initial"].item() - il.item()) <= 1e-5 * abs(il.item())
    assert abs(L["residual"].item() - (r ** 2).mean().item()) <= 1e-5 * (r ** 2).mean().item()
    assert abs(L["total"].item() - (L["residual"] + 10 * L["boundary"] + 10 * L["initial"]).item()) <= 1e-5 * abs(L["total"].item())
    L["total"].backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in model.parameters())


def test_adaptive_sampling_with_agent(dev):
    """generate_collocation_points('adaptive') with an attached agent (pde_base.py:961-1072; the reference's trainer
    never attaches it) and the RAR route through the CUDA scoring kernel."""
    import pinns_rl_pde_b200 as pk

    class Agent:
        def __init__(self):
            self.calls, self.eps = 0, []

        def select_action(self, pts):
            self.calls += 1
            return torch.exp(-((pts[:, 0:1]) ** 2) * 8.0)       # prefers x near 0

        def update_epsilon(self, k):
            self.eps.append(k)

        def update(self, state, reward):
            self.last = (state.shape, float(reward))

    pde = product_pde("burgers", dev)
    assert pde.generate_collocation_points(100, "adaptive")[0].shape == (100, 1)      # no agent: uniform fallback
    pde.rl_agent = Agent()
    x, t = pde.generate_collocation_points(2000, "adaptive")
    assert x.shape == (2000, 1) and t.shape == (2000, 1) and pde.rl_agent.calls == 1
    assert x.abs().mean().item() < 0.35 and x.min() >= -1 and x.max() <= 1 and t.min() >= 0 and t.max() <= 1
    pde.generate_collocation_points(2000, "adaptive")
    assert pde.rl_agent.eps == [2]
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 64, 3, dev)
    r = pde.compute_residual(model, x, t)
    pde.update_sampling_strategy(x, t, r.detach())
    assert pde.rl_agent.last[0] == (2000, 2) and pde.rl_agent.last[1] > 0
    xs, ts = pde.generate_collocation_points(500, "stratified")
    assert xs.shape == (500, 1)
    with pytest.raises(ValueError):
        pde.generate_collocation_points(10, "nonsense")

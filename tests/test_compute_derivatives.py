"""PDEBase.compute_derivatives shim (pde_base.py:590-794; the plugin-PDE surface of CONTRIBUTING.md:152-244).

CPU: the key / order bookkeeping (incl. the F1 "dx2 holds u_x" quirk and the F2 zeros) is checked against the oracle port
of the reference routine with ``functional.jets`` replaced by the fp64 Taylor-mode oracle (no compute on a GPU needed).
GPU: the same comparison through libpinnk, plus parameter gradients of a plugin-style residual written on top of the dict."""
import pytest
import torch

import parity_log
from helpers import flat_grad, port_model, rel

CASES = [  # (dimension, temporal, spatial)
    (1, [1], [2]),            # HeatEquation's request: "dx2" / "laplacian" hold u_x
    (1, [1], [1, 2]),         # Burgers
    (1, [1], [1, 2, 3]),      # KdV
    (1, [1, 2], [2, 4]),      # listed orders 2, 4 -> first and second derivative
    (1, [2], [1, 3]),         # "dt2" alone holds u_t; "dx3" holds u_xx
    (1, [0, 1], [0, 1, 2, 3, 4]),
    (2, [1], [1, 2]),         # every spatial entry is zero
    (2, [1, 2], [2]),
]


def _oracle_jets(model64):
    from oracle import jets_oracle

    def jets(model, xt, directions):
        u, dirs = jets_oracle.network_jet(model64, jets_oracle.seed_jet(xt.double(), directions))
        return torch.cat([u] + [a for d in dirs for a in d], dim=1)
    return jets


def _setup(dim, dev):
    import pinns_rl_pde_b200 as pk
    torch.manual_seed(5)
    model = pk.make_model("feedforward", dim + 1, 32, 3, dev)
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    m64 = port_model(dict(arch="feedforward", hidden=32, layers=3, dimension=dim, extra={}), state, torch.float64)
    g = torch.Generator().manual_seed(6)
    return model, m64, torch.rand(97, dim, generator=g) * 2 - 1, torch.rand(97, 1, generator=g)


@pytest.mark.parametrize("dim,temporal,spatial", CASES)
def test_keys_and_orders_match_the_reference_routine_cpu(monkeypatch, dim, temporal, spatial):
    from oracle import ref_port
    from pinns_rl_pde_b200 import functional as F
    from helpers import product_pde
    model, m64, x, t = _setup(dim, torch.device("cpu"))
    monkeypatch.setattr(F, "jets", _oracle_jets(m64))
    monkeypatch.setattr(F, "_prep", lambda model, x, t: (x, t))
    pde = product_pde("burgers", torch.device("cpu"), dim)
    got = pde.compute_derivatives(model, x, t, temporal_derivatives=temporal, spatial_derivatives=set(spatial))
    want = ref_port.compute_derivatives(m64, x.double(), t.double(), spatial, temporal, dimension=dim)
    want = {k: v for k, v in want.items() if not k.startswith("_")}
    assert set(got) == set(want), (sorted(got), sorted(want))
    for k in want:
        w = want[k] if want[k] is not None else torch.zeros(97, 1, dtype=torch.float64)
        assert got[k].shape == (97, 1)
        assert float((got[k].double() - w.detach()).abs().max()) <= 1e-10 * max(1.0, float(w.abs().max())), k
    if 2 in spatial and dim == 1:
        assert got["laplacian"] is got["dx2"]
    pde.compat = "math"      # what the keys say
    got = pde.compute_derivatives(model, x, t, temporal_derivatives=temporal, spatial_derivatives=set(spatial))
    X = torch.cat([x, t], dim=1).double().requires_grad_(True)
    d1 = torch.autograd.grad(m64(X).sum(), X, create_graph=True)[0]
    if 1 in temporal:
        assert rel(got["dt"], d1[:, dim:dim + 1].detach()) < 1e-10
    key2 = "dx2" if dim == 1 else "dx1x1"
    if 2 in spatial:
        d2 = torch.autograd.grad(d1[:, 0].sum(), X)[0][:, 0:1]
        assert rel(got[key2], d2) < 1e-10
    with pytest.raises(ValueError):
        pde.compute_derivatives(model, x, t, temporal_derivatives=[3])
    with pytest.raises(ValueError):
        pde.compute_derivatives(model, x, t, spatial_derivatives=[5])


@pytest.mark.gpu
@pytest.mark.parametrize("dim,temporal,spatial", CASES)
def test_compute_derivatives_on_libpinnk(dim, temporal, spatial):
    from oracle import ref_port
    from helpers import product_pde
    dev = torch.device("cuda:0")
    model, m64, x, t = _setup(dim, dev)
    pde = product_pde("burgers", dev, dim)
    got = pde.compute_derivatives(model, x.to(dev), t.to(dev), temporal_derivatives=temporal, spatial_derivatives=set(spatial))
    want = ref_port.compute_derivatives(m64, x.double(), t.double(), spatial, temporal, dimension=dim)
    want = {k: v for k, v in want.items() if not k.startswith("_")}
    assert set(got) == set(want)
    worst = 0.0
    for k, w in want.items():
        w = w.detach() if w is not None else torch.zeros(97, 1, dtype=torch.float64)
        if float(w.abs().max()) == 0.0:
            assert float(got[k].abs().max()) == 0.0, k
        else:
            worst = max(worst, rel(got[k].cpu(), w))
    parity_log.log(f"[compute_derivatives] dim {dim} temporal {temporal} spatial {spatial}: worst rel err vs ref64 {worst:.3e}")
    assert worst <= 1e-5


@pytest.mark.gpu
def test_plugin_style_residual_on_the_dict_is_differentiable():
    """A residual written per CONTRIBUTING.md:152-244 on top of compute_derivatives: r = dt + u*dx - nu*dx2 with the
    reference's bookkeeping; its parameter gradient through pinnk_jets_vjp against autograd through the oracle."""
    from oracle import ref_port
    from helpers import product_pde
    dev = torch.device("cuda:0")
    model, m64, x, t = _setup(1, dev)
    pde = product_pde("burgers", dev, 1)
    d = pde.compute_derivatives(model, x.to(dev), t.to(dev), temporal_derivatives=[1], spatial_derivatives={1, 2})
    u = model(torch.cat([x, t], dim=1).to(dev))
    r = d["dt"] + u * d["dx"] - 0.01 * d["dx2"]
    (r ** 2).mean().backward()
    w = ref_port.compute_derivatives(m64, x.double(), t.double(), [1, 2], [1])
    u64 = m64(torch.cat([w["_x"], w["_t"]], dim=1))
    r64 = w["dt"] + u64 * w["dx"] - 0.01 * w["dx2"]
    (r64 ** 2).mean().backward()
    og = torch.cat([p.grad.reshape(-1) for p in m64.parameters()])
    e_r, e_g = rel(r.detach().cpu(), r64.detach()), rel(flat_grad(model).cpu(), og)
    parity_log.log(f"[compute_derivatives] plugin-style Burgers residual rel {e_r:.3e}, grad rel {e_g:.3e}")
    assert e_r <= 1e-5 and e_g <= 1e-5

"""CPU: the oracle (oracle/ref_port.py, oracle/jets_oracle.py) against the committed golden fixtures that
tests/golden/make_golden.py generated from the UNMODIFIED reference."""
import pytest
import torch

from oracle import jets_oracle, ref_port
from helpers import PDES, fixtures, load_fixture, port_model, rel


def _kw(params):
    return {k: v for k, v in params.items() if k != "speed"}


@pytest.mark.parametrize("tag", fixtures())
def test_port_reproduces_reference_fp32(tag):
    z, meta, state = load_fixture(tag)
    m = port_model(meta, state)
    x, t = torch.from_numpy(z["x"]), torch.from_numpy(z["t"])
    r = ref_port.RESIDUALS[meta["pde"]](m, x.clone(), t.clone(), dimension=meta["dimension"], **_kw(meta["params"]))
    # same torch ops in the same order: bit-equal on the generating machine, ulp-level elsewhere
    assert rel(r.detach(), z["residual32"]) < 2e-6
    s = PDES[meta["pde"]]
    if meta["mode"] == "loss":
        if meta["pde"] == "heat":
            n = meta["n"]
            L = ref_port.heat_compute_loss(m, r, s["domain"], s["time"],
                                           ref_port.initial_condition_fn("heat", s["ic"], s["domain"], meta["params"]),
                                           max(n // 10, 10), max(n // 5, 10))
        else:
            L = ref_port.base_compute_loss(m, r, s["domain"], s["time"],
                                           ref_port.boundary_condition_fns(meta["pde"], s["bcs"], s["ic"], s["domain"],
                                                                           meta["params"], s["exact"]))
        for k in ("residual", "boundary", "initial", "total"):
            assert abs(L[k].item() - float(z[f"loss32_{k}"])) <= 2e-6 * abs(float(z[f"loss32_{k}"])) + 1e-12
        total = L["total"]
    else:
        total = (r ** 2).mean()
    g = jets_oracle.flat_grad(m, total)
    assert rel(g, z["grad32"]) < 1e-5


@pytest.mark.parametrize("tag", fixtures())
def test_jets_oracle_matches_reference_fp64(tag):
    z, meta, state = load_fixture(tag)
    has_ln = meta["arch"] == "resnet"
    m = port_model(meta, state, torch.float64, corrected=has_ln)
    x, t = torch.from_numpy(z["x"]).double(), torch.from_numpy(z["t"]).double()
    r = jets_oracle.residual(m, meta["pde"], x, t, meta["params"], meta["dimension"], "reference")
    want = z["residual64_corrected"] if has_ln else z["residual64"]
    assert rel(r.detach(), want) < 1e-10
    g = jets_oracle.flat_grad(m, (r ** 2).mean())
    assert rel(g, z["grad64_mse_corrected"] if has_ln else z["grad64_mse"]) < 1e-9
    if "residual64_math" in z.files:
        rm = jets_oracle.residual(m, meta["pde"], x, t, meta["params"], meta["dimension"], "math")
        assert rel(rm.detach(), z["residual64_math"]) < 1e-9
        assert rel(jets_oracle.flat_grad(m, (rm ** 2).mean()), z["grad64_mse_math"]) < 1e-8


def test_layernorm_reference_inexactness_is_recorded():
    """the unmodified reference differs from exact math through nn.LayerNorm at 3rd order."""
    z, meta, _ = load_fixture("c3_kdv_resnet_small")
    assert rel(z["residual64"], z["residual64_corrected"]) > 1e-3


def test_heat_smoothness_port_matches_reference_fixture():
    """oracle/ref_port.heat_smoothness_loss (heat_equation.py:625-650) against the values the UNMODIFIED reference produced
    (tests/golden/x_heat_smoothness.npz, written by make_golden.py after asserting bit-equality in fp32)."""
    import os
    import numpy as np
    from helpers import GOLDEN
    z = np.load(os.path.join(GOLDEN, "x_heat_smoothness.npz"))
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    x, t = torch.from_numpy(z["x"]), torch.from_numpy(z["t"])
    dom = PDES["heat"]["domain"]
    for dtype, key, tol in ((torch.float32, "smooth32", 2e-6), (torch.float64, "smooth64", 1e-12)):
        m = ref_port.PINNModel("fourier", 2, 128, 3, 1, "tanh", mapping_size=32, scale=10.0)
        m.load_state_dict(state)
        m = m.to(dtype)
        s = ref_port.heat_smoothness_loss(m, x.to(dtype), t.to(dtype), dom)
        assert abs(float(s) - float(z[key])) <= tol * abs(float(z[key])), (key, float(s), float(z[key]))
        if dtype == torch.float64:
            g = torch.autograd.grad(s, list(m.parameters()), allow_unused=True)
            g = torch.cat([(torch.zeros_like(p) if gi is None else gi).reshape(-1) for p, gi in zip(m.parameters(), g)])
            assert rel(g, z["gsmooth64"]) <= 1e-6          # (the fixture stores the fp64 gradient rounded to fp32)

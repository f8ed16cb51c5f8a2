"""GPU: the K-split launch of 256-wide contractions (linear_rows_ts_ksplit_kernel: CTA pairs, partial products through an
L2-resident ring; opt-in with PINNK_ENABLE_KSPLIT=1 because it measured slower) against the two K-half passes of the same
library on identical seeded inputs.  The
arithmetic is the same (own accumulator + partner's partial product), so residuals and loss components must agree to the
bit; gradients go through the weight-gradient kernel's arrival-order reduction and agree to its run-to-run noise."""
import pytest
import torch

import parity_log
from helpers import flat_grad, product_pde

pytestmark = pytest.mark.gpu


def _rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / b.norm().clamp_min(1e-300))


CASES = [  # name, pde, arch, layers, dimension, n, compat, mode, extra
    ("resnet 2x256 / KdV (plain forward + LayerNorm, dgrad with and without the tanh adjoint, 5 jet columns)",
     "kdv", "resnet", 2, 1, 9000, "reference", "loss", {"num_blocks": 2}),
    ("siren 3x256 / Burgers (fused Linear + sin jets, dgrad + sin adjoint, 4 jet columns)",
     "burgers", "siren", 3, 1, 8192, "reference", "loss", {"omega_0": 30.0}),
    ("siren 3x256 / 2-D Cahn-Hilliard, intended operator (18 jet columns: plain GEMMs, ragged last tile)",
     "cahn_hilliard", "siren", 3, 2, 2500, "math", "mse", {"omega_0": 30.0}),
    ("feedforward 3x256 / Allen-Cahn (fused Linear + tanh jets incl. the output layer folded in)",
     "allen_cahn", "feedforward", 3, 1, 8192, "reference", "loss", {}),
]


@pytest.mark.parametrize("case", CASES, ids=[c[0].split(" (")[0] for c in CASES])
def test_ksplit_matches_two_pass(monkeypatch, case):
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import _lib
    name, pde_name, arch, layers, dim, n, compat, mode, extra = case
    dev = torch.device("cuda:0")
    torch.manual_seed(11)
    model = pk.make_model(arch, dim + 1, 256, layers, dev, **extra)
    pde = product_pde(pde_name, dev, dim, compat=compat)
    g = torch.Generator().manual_seed(7)
    lo, hi = pde.domain[0]
    x = (lo + (hi - lo) * torch.rand(n, dim, generator=g)).to(dev)
    t = (pde.time_domain[0] + (pde.time_domain[1] - pde.time_domain[0]) * torch.rand(n, 1, generator=g)).to(dev)

    def step():
        for p in model.parameters():
            p.grad = None
        before = _lib.launch_count()
        if mode == "loss":
            L = pde.compute_loss(model, x, t)
            loss = L["total"]
        else:
            loss = (pde.compute_residual(model, x, t) ** 2).mean()
        loss.backward()
        torch.cuda.synchronize()
        launches = _lib.launch_count() - before
        return float(loss), flat_grad(model).clone(), pde.compute_residual(model, x, t).detach().clone(), \
            pde.score_residual(model, x, t)[0].clone(), launches

    monkeypatch.setenv("PINNK_DETERMINISTIC", "1")
    monkeypatch.setenv("PINNK_ENABLE_KSPLIT", "1")
    ks = step()
    monkeypatch.delenv("PINNK_ENABLE_KSPLIT")
    two = step()
    er, es, eg = _rel(ks[2], two[2]), _rel(ks[3], two[3]), _rel(ks[1], two[1])
    parity_log.log(f"[ksplit {name.split(' (')[0]}] vs two K-half passes: residual {er:.2e}, scores {es:.2e}, loss "
                   f"{abs(ks[0] - two[0]) / abs(two[0]):.2e}, grad {eg:.2e}; launches {ks[4]} vs {two[4]}")
    assert torch.isfinite(ks[1]).all() and torch.isfinite(ks[2]).all()
    assert er <= 1e-7 and es <= 1e-7 and abs(ks[0] - two[0]) <= 1e-6 * abs(two[0]) and eg <= 2e-6, (er, es, eg)

"""GPU: hidden widths that are not multiples of 128 (the reference's YAML ships SIREN 124 and an autoencoder 124 / 248,
config.yaml:22,43) run on the tcgen05 tiles through zero-padded shadow parameters (program._pad_program): values, losses and the
gradients of the real entries must equal the unpadded route of the same library (PINNK_DISABLE_PAD=1: exact-fp32 CUDA-core GEMMs)
and the fp64 oracle."""
import pytest
import torch

import parity_log
from helpers import PDES, flat_grad, port_model, product_pde, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("arch,hidden,layers,pde_name,extra", [
    ("siren", 124, 4, "burgers", {"omega_0": 30.0}),
    ("feedforward", 248, 3, "kdv", {}),
    ("feedforward", 100, 3, "allen_cahn", {}),
])
def test_padded_widths_match_the_unpadded_route_and_the_oracle(monkeypatch, arch, hidden, layers, pde_name, extra):
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import engine, program
    from oracle import ref_port
    dev = torch.device("cuda:0")
    torch.manual_seed(5)
    model = pk.make_model(arch, 2, hidden, layers, dev, **extra)
    assert program.compile_network(model).padded
    pde = product_pde(pde_name, dev)
    s = PDES[pde_name]
    g = torch.Generator().manual_seed(6)
    n = 2500
    lo, hi = s["domain"][0]
    x = torch.rand(n, 1, generator=g) * (hi - lo) + lo
    t = torch.rand(n, 1, generator=g) * (s["time"][1] - s["time"][0]) + s["time"][0]
    xd, td = x.to(dev), t.to(dev)

    def run():
        model.zero_grad()
        L = pde.compute_loss(model, xd, td)
        L["total"].backward()
        r = pde.compute_residual(model, xd, td).detach()
        return float(L["total"]), flat_grad(model).clone(), r, pde.score_residual(model, xd, td)[0].clone()

    padded = run()
    monkeypatch.setenv("PINNK_DISABLE_PAD", "1")
    engine._CACHE.clear()
    assert not program.compile_network(model).padded
    plain = run()
    monkeypatch.delenv("PINNK_DISABLE_PAD")
    engine._CACHE.clear()
    el, eg, er = abs(padded[0] - plain[0]) / abs(plain[0]), rel(padded[1], plain[1]), rel(padded[2], plain[2])
    # oracle (fp64 and fp32 of the reference algorithm) for the residual
    state = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    meta = dict(arch=arch, hidden=hidden, layers=layers, dimension=1, extra=extra)
    kw = {k: v for k, v in s["params"].items() if k != "speed"}
    want = ref_port.RESIDUALS[pde_name](port_model(meta, state, torch.float64), x.double(), t.double(), dimension=1, **kw).detach()
    r32 = ref_port.RESIDUALS[pde_name](port_model(meta, state, torch.float32), x.clone(), t.clone(), dimension=1, **kw).detach()
    e64, floor = rel(padded[2].cpu(), want), rel(r32, want)
    gate = max(1e-5, 2 * floor)
    parity_log.log(f"[padding {arch} {layers}x{hidden} / {pde_name}] padded (tcgen05) vs unpadded (CUDA-core fp32): loss {el:.2e}, grad "
                   f"{eg:.2e}, residual {er:.2e}; residual cuda-vs-ref64 {e64:.3e} | ref32-vs-ref64 {floor:.3e} | gate {gate:.3e}")
    assert el <= 1e-5 and eg <= 1e-5 and er <= 1e-5 and rel(padded[3], plain[3]) <= 1e-5
    assert e64 <= gate


def test_padded_network_trains_with_the_fused_step():
    """The fused trainer step (flat gradient -> clip + Adam through raw pointers) on a padded network follows the autograd route."""
    import copy
    import pinns_rl_pde_b200 as pk
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    m1 = pk.make_model("siren", 2, 124, 3, dev, omega_0=30.0)
    m2 = copy.deepcopy(m1)
    pde = product_pde("burgers", dev)
    cfg = pk.TrainingConfig(learning_rate=1e-4, weight_decay=0.0, gradient_clipping=1.0, scheduler="none")
    t1 = pk.PDETrainer(m1, pde, config=cfg, device=dev, fused=True)
    t2 = pk.PDETrainer(m2, pde, config=cfg, device=dev, fused=False)
    g = torch.Generator().manual_seed(3)
    for it in range(8):
        x, t = (torch.rand(3000, 1, generator=g) * 2 - 1).to(dev), torch.rand(3000, 1, generator=g).to(dev)
        l1, l2 = t1.train_step(x, t), t2.train_step(x, t)
        assert abs(float(l1["total"]) - float(l2["total"])) <= 5e-5 * abs(float(l2["total"])), (it, float(l1["total"]), float(l2["total"]))
    p1 = torch.cat([p.detach().reshape(-1) for p in m1.parameters()])
    p2 = torch.cat([p.detach().reshape(-1) for p in m2.parameters()])
    parity_log.log(f"[padding] 8 fused trainer steps on siren 3x124 vs autograd route: parameters {rel(p1, p2):.3e}")
    assert rel(p1, p2) <= 1e-5

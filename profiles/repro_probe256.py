"""Bit-reproducibility of the 256-wide routes (two K-half launches, tiled TMA, separate activation / LayerNorm kernels)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib
from helpers import product_pde
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
M, JC = 240000, 4

def report(name, outs):
    ref = outs[0]
    for i, o in enumerate(outs[1:], 1):
        ne = (o != ref)
        n = int(ne.sum())
        if n == 0:
            print(f"{name}: run {i} identical"); continue
        d = (o - ref).abs().max().item()
        if ne.dim() == 2:
            rows = ne.any(dim=1).nonzero().reshape(-1)
            print(f"{name}: run {i} differs in {n} of {ref.numel()} elements, {rows.numel()} rows (first {rows[:8].tolist()}, row%64 {[r % 64 for r in rows[:10].tolist()]}), max abs diff {d:.3e} (max |ref| {ref.abs().max().item():.3e}); cols differing in first row {int(ne[rows[0]].sum())}")
        else:
            print(f"{name}: run {i} differs in {n} of {ref.numel()} elements, max abs diff {d:.3e} (max |ref| {ref.abs().max().item():.3e})")

X = torch.randn(M, 256, generator=g, device=dev); W = torch.randn(256, 256, generator=g, device=dev) / 16; b = torch.randn(256, generator=g, device=dev)
dZ = torch.randn(M, 256, generator=g, device=dev)
report("plain fwd K=256 N=256", [_lib.debug_linear_fwd(X, W, b, JC, 1) for _ in range(3)])
report("plain dgrad 256", [_lib.debug_linear_dgrad(dZ, W, 1) for _ in range(3)])
report("wgrad 256 (atomics)", [_lib.debug_linear_wgrad(dZ, X, JC, 1)[0] for _ in range(3)])
X1 = X[:, :128].contiguous(); W1 = torch.randn(128, 128, generator=g, device=dev) / 11
report("wgrad 128 (atomics)", [_lib.debug_linear_wgrad(dZ[:, :128].contiguous(), X1, JC, 1)[0] for _ in range(3)])
for layers in (1, 2, 3):
    torch.manual_seed(0)
    model = pk.make_model("siren", 2, 256, layers, dev, omega_0=30.0)
    xt = torch.rand(60000, 2, generator=g, device=dev)
    with torch.no_grad():
        report(f"siren 256 x {layers} hidden: jets forward (order 2,1)", [pk.jets(model, xt, [((1.0, 0.0), 2), ((0.0, 1.0), 1)]) for _ in range(3)])
        report(f"siren 256 x {layers} hidden: value forward", [model(xt) for _ in range(3)])
os.environ["PINNK_DETERMINISTIC"] = "1"
for nb in (1,):
    torch.manual_seed(0)
    model = pk.make_model("resnet", 2, 256, nb, dev, num_blocks=nb)
    pde = product_pde("kdv", dev)
    x, t = torch.rand(60000, 1, generator=g, device=dev) * 30 - 15, torch.rand(60000, 1, generator=g, device=dev) * 5
    outs = []
    for _ in range(3):
        model.zero_grad()
        pde.compute_loss(model, x, t)["total"].backward()
        torch.cuda.synchronize()
        outs.append({n: p.grad.clone() for n, p in model.named_parameters()})
    for n in outs[0]:
        report(f"resnet {nb} block {n}", [o[n] for o in outs])

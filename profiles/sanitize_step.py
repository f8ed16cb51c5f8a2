"""Smallest shapes for compute-sanitizer (SURVEY section 5 "race detection / sanitizers"):
    compute-sanitizer --tool memcheck  python profiles/sanitize_step.py
    compute-sanitizer --tool racecheck python profiles/sanitize_step.py
One loss + backward step of (a) the smoke network (Burgers, feed-forward 4x128: the tcgen05 rows kernels incl. the
loss-fused and input-layer-fused ones, TS-mode wgrad) and (b) a 256-wide ResNet with LayerNorm on KdV (K = 256 GEMMs, LayerNorm
/ activation kernels), then a forward-only scoring call and the deterministic wgrad reduction."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import pinns_rl_pde_b200 as pk  # noqa: E402
from helpers import product_pde  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
g = torch.Generator().manual_seed(1)
for arch, hidden, layers, extra, pde_name, n in [("feedforward", 128, 4, {}, "burgers", 2048),
                                                 ("resnet", 256, 2, {"num_blocks": 2}, "kdv", 512)]:
    model = pk.make_model(arch, 2, hidden, layers, dev, **extra)
    pde = product_pde(pde_name, dev)
    x, t = (torch.rand(n, 1, generator=g) * 2 - 1).to(dev), torch.rand(n, 1, generator=g).to(dev)
    for det in ("0", "1"):
        os.environ["PINNK_DETERMINISTIC"] = det
        model.zero_grad()
        losses = pde.compute_loss(model, x, t)
        losses["total"].backward()
        torch.cuda.synchronize()
        gn = float(torch.cat([p.grad.reshape(-1) for p in model.parameters()]).norm())
        print(f"{arch} {layers}x{hidden} {pde_name} n={n} det={det}: loss {losses['total'].item():.6e} |grad| {gn:.6e}", flush=True)
    mag, stats = pde.score_residual(model, x, t)
    torch.cuda.synchronize()
    print(f"  score: sum|r| {stats[0].item():.6e} max {stats[2].item():.6e}", flush=True)
    big = 70000          # > 32768 rows: the eager per-component route with the loss-fused last layer
    xb, tb = (torch.rand(big, 1, generator=g) * 2 - 1).to(dev), torch.rand(big, 1, generator=g).to(dev)
    if arch == "feedforward":
        model.zero_grad()
        pde.compute_loss(model, xb, tb)["total"].backward()
        torch.cuda.synchronize()
        print("  70000-row step ok", flush=True)
print("sanitize_step: done")

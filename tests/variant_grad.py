"""Helper of test_gpu_variants.py: gradient + residual of one seeded Burgers / Allen-Cahn problem, written to argv[1].
The library reads its PINNK_* switches once per process, so every variant runs in its own interpreter."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import pinns_rl_pde_b200 as pk  # noqa: E402
from helpers import flat_grad, product_pde  # noqa: E402

dev = torch.device("cuda:0")
out = {}
for name, layers, n in (("burgers", 8, 6000), ("allen_cahn", 4, 4100)):
    torch.manual_seed(7)
    model = pk.make_model("feedforward", 2, 128, layers, dev)
    with torch.no_grad():                                   # push some units into saturation (w0 ~ 1e-4 .. 0)
        model.model.layers[2].weight.mul_(6.0)
    pde = product_pde(name, dev)
    g = torch.Generator().manual_seed(11)
    x = (torch.rand(n, 1, generator=g) * 2 - 1).to(dev)
    t = torch.rand(n, 1, generator=g).to(dev)
    losses = pde.compute_loss(model, x, t)
    losses["total"].backward()
    out[name + "_grad"] = flat_grad(model).cpu().numpy()
    out[name + "_loss"] = np.array([float(losses[k]) for k in ("residual", "boundary", "initial", "total")])
    out[name + "_res"] = pde.compute_residual(model, x, t).detach().cpu().numpy()
np.savez(sys.argv[1], **out)

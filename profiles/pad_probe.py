"""Shipped-YAML SIREN (7 hidden layers of 124 units, config.yaml:21-23) on Burgers, 262 144 points: compute_loss + backward with
the widths padded onto the tcgen05 tiles (default) and on the exact-fp32 CUDA-core GEMMs (PINNK_DISABLE_PAD=1)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import engine
from helpers import product_pde
dev = torch.device('cuda:0')
n = 1 << 18
for label, env in (("padded to 128 (tcgen05)", None), ("unpadded (CUDA-core fp32)", "1")):
    if env: os.environ["PINNK_DISABLE_PAD"] = env
    else: os.environ.pop("PINNK_DISABLE_PAD", None)
    engine._CACHE.clear()
    torch.manual_seed(0)
    model = pk.make_model("siren", 2, 124, 8, dev, omega_0=30.0)
    pde = product_pde("burgers", dev)
    x = torch.rand(n, 1, device=dev) * 2 - 1; t = torch.rand(n, 1, device=dev)
    def step():
        model.zero_grad(set_to_none=True)
        pde.compute_loss(model, x, t)["total"].backward()
    for _ in range(2): step()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): step()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print(f"siren 8x124 / burgers, {n} points, {label}: {ms:.2f} ms/step -> {n / ms / 1e3:.2f} Mpts/s", flush=True)

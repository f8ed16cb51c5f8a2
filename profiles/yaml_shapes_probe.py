"""Step time (compute_loss + backward) of the network shapes the reference's YAML ships (config.yaml:8-31), 65 536 points."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib
from helpers import product_pde
dev = torch.device('cuda:0')
n = 1 << 16
for name, arch, hidden, layers, pde_name, extra in (
        ("feedforward 7x128 / burgers", "feedforward", 128, 8, "burgers", {}),
        ("resnet 7 blocks x 512 / burgers", "resnet", 512, 7, "burgers", {"num_blocks": 7}),
        ("siren 7x124 / kdv", "siren", 124, 8, "kdv", {"omega_0": 30.0}),
        ("fourier 4x512, 512 features / heat", "fourier", 512, 5, "heat", {"mapping_size": 512, "scale": 4.0})):
    torch.manual_seed(0)
    model = pk.make_model(arch, 2, hidden, layers, dev, **extra)
    pde = product_pde(pde_name, dev)
    lo, hi = pde.domain[0]
    x = torch.rand(n, 1, device=dev) * (hi - lo) + lo; t = torch.rand(n, 1, device=dev) * (pde.time_domain[1] - pde.time_domain[0])
    def step():
        model.zero_grad(set_to_none=True)
        pde.compute_loss(model, x, t)["total"].backward()
    for _ in range(2): step()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): step()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 3
    _lib.prof_enable(True); torch.cuda.synchronize(); step(); torch.cuda.synchronize()
    p = _lib.prof_collect(); _lib.prof_enable(False)
    print(f"{name}: {ms:.2f} ms/step -> {n / ms / 1e3:.2f} Mpts/s   ", {k: round(v[0], 1) for k, v in p.items() if v[0] >= 0.05 * ms}, flush=True)

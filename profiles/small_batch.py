"""Latency of the hot path at the batch sizes the reference's own trainer uses (README: 5000 points -> 2025 per step)."""
import os, sys, time
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib
from helpers import product_pde
dev = torch.device('cuda:0')
def run(name, pde_name, arch, layers, n, extra, fused, graph=False):
    torch.manual_seed(0)
    model = pk.make_model(arch, 2, 128, layers, dev, **extra)
    pde = product_pde(pde_name, dev)
    x = torch.rand(n, 1, device=dev); t = torch.rand(n, 1, device=dev)
    cfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=0.0, gradient_clipping=1.0, scheduler="none")
    tr = pk.PDETrainer(model, pde, config=cfg, device=dev, fused=fused, graph=graph)
    for _ in range(5): tr.train_step(x, t)
    torch.cuda.synchronize()
    l0 = _lib.launch_count(); t0 = time.perf_counter()
    for _ in range(50): tr.train_step(x, t)
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"{name}: n={n} {'fused+graph' if graph else 'fused' if fused else 'autograd-route'} step {1e3 * (t2 - t0) / 50:.3f} ms (host enqueue {1e3 * (t1 - t0) / 50:.3f} ms), "
          f"{(_lib.launch_count() - l0) / 50:.0f} libpinnk launches/step")
for graph in (True,):
    run("C1 heat/fourier 4x128", "heat", "fourier", 4, 4900, {"mapping_size": 32, "scale": 10.0}, True, graph)
    run("C1 heat/fourier 4x128", "heat", "fourier", 4, 2025, {"mapping_size": 32, "scale": 10.0}, True, graph)
    run("C2 burgers/ff 8x128", "burgers", "feedforward", 8, 2025, {}, True, graph)
    run("C2 burgers/ff 8x128", "burgers", "feedforward", 8, 65536, {}, True, graph)
for fused in (True, False):
    run("C1 heat/fourier 4x128", "heat", "fourier", 4, 4900, {"mapping_size": 32, "scale": 10.0}, fused)
    run("C1 heat/fourier 4x128", "heat", "fourier", 4, 2025, {"mapping_size": 32, "scale": 10.0}, fused)
    run("C2 burgers/ff 8x128", "burgers", "feedforward", 8, 2025, {}, fused)

"""Throughput of the other BASELINE configs (step = compute_residual-MSE or compute_loss + backward), one B200."""
import sys, math, time, torch
import os; _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import pinns_rl_pde_b200 as pk
from helpers import product_pde
dev = torch.device('cuda:0')
def run(name, pde_name, arch, hidden, layers, dim, n, mode, extra, compat="reference"):
    torch.manual_seed(0)
    model = pk.make_model(arch, dim + 1, hidden, layers, dev, **extra)
    pde = product_pde(pde_name, dev, dim, compat=compat)
    x = torch.rand(n, dim, device=dev); t = torch.rand(n, 1, device=dev)
    def step():
        model.zero_grad(set_to_none=True)
        if mode == "loss": pde.compute_loss(model, x, t)["total"].backward()
        elif mode == "mse": (pde.compute_residual(model, x, t) ** 2).mean().backward()
        else: pde.score_residual(model, x, t, want_abs=True)
    for _ in range(2): step()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(3): step()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 3
    print(f"{name}: n={n} {ms:.1f} ms/step -> {n / ms / 1e3:.2f} Mpts/s")
run("C1 heat/fourier 4x128 (loss)", "heat", "fourier", 128, 4, 1, 4900, "loss", {"mapping_size": 32, "scale": 10.0})
run("C1 heat/fourier 4x128 (loss, 1M)", "heat", "fourier", 128, 4, 1, 1 << 20, "loss", {"mapping_size": 32, "scale": 10.0})
run("C2 burgers/ff 8x128 (loss)", "burgers", "feedforward", 128, 8, 1, 1 << 20, "loss", {})
run("C3 kdv/resnet 6x256 (loss)", "kdv", "resnet", 256, 6, 1, 1 << 19, "loss", {"num_blocks": 6})
run("C4 ch2d/siren 5x256 as-written (mse)", "cahn_hilliard", "siren", 256, 5, 2, 1 << 20, "mse", {"omega_0": 30.0})
run("C4 ch2d/siren 5x256 math 18 cols (mse)", "cahn_hilliard", "siren", 256, 5, 2, 1 << 17, "mse", {"omega_0": 30.0}, "math")
run("C5 allen-cahn/ff 8x128 scoring", "allen_cahn", "feedforward", 128, 8, 1, 1 << 22, "score", {})
run("next: wave/ff 8x128 (loss, 5 jet columns)", "wave", "feedforward", 128, 8, 1, 1 << 19, "loss", {})
run("next: convection/ff 8x128 (loss, 3 jet columns)", "convection", "feedforward", 128, 8, 1, 1 << 20, "loss", {})

from pinns_rl_pde_b200 import _lib
def prof(name, pde_name, arch, hidden, layers, dim, n, extra, compat="reference"):
    torch.manual_seed(0)
    model = pk.make_model(arch, dim + 1, hidden, layers, dev, **extra)
    pde = product_pde(pde_name, dev, dim, compat=compat)
    x = torch.rand(n, dim, device=dev); t = torch.rand(n, 1, device=dev)
    def step():
        model.zero_grad(set_to_none=True)
        (pde.compute_residual(model, x, t) ** 2).mean().backward()
    step()
    _lib.prof_enable(True); torch.cuda.synchronize()
    step(); torch.cuda.synchronize()
    p = _lib.prof_collect(); _lib.prof_enable(False)
    tot = sum(v[0] for v in p.values())
    print(name, {k: (round(v[0], 1), v[1]) for k, v in p.items() if v[1]}, "total ms", round(tot, 1))
if len(sys.argv) > 1:
    prof("C3 kdv/resnet", "kdv", "resnet", 256, 6, 1, 1 << 18, {"num_blocks": 6})
    prof("C4 ch2d/siren math", "cahn_hilliard", "siren", 256, 5, 2, 1 << 17, {"omega_0": 30.0}, "math")

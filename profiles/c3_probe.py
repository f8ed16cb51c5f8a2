"""C3 (KdV / ResNet 6x256) and C4-math (2-D Cahn-Hilliard / SIREN 5x256, 18 jet columns): step time and per-kernel-class
CUDA-event times, with the switches given as NAME=VALUE arguments applied for an A/B column (e.g. PINNK_ENABLE_LNACT=1)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib
from helpers import product_pde
dev = torch.device('cuda:0')
ab = dict(a.split("=", 1) for a in sys.argv[1:] if "=" in a)
which = [a for a in sys.argv[1:] if "=" not in a] or ["c3", "c4"]


def run(name, pde_name, arch, hidden, layers, dim, n, mode, extra, compat="reference"):
    torch.manual_seed(0)
    model = pk.make_model(arch, dim + 1, hidden, layers, dev, **extra)
    pde = product_pde(pde_name, dev, dim, compat=compat)
    x = torch.rand(n, dim, device=dev); t = torch.rand(n, 1, device=dev)

    def step():
        model.zero_grad(set_to_none=True)
        if mode == "loss": pde.compute_loss(model, x, t)["total"].backward()
        else: (pde.compute_residual(model, x, t) ** 2).mean().backward()

    for label, env in (("default", {}), ("A/B " + " ".join(f"{k}={v}" for k, v in ab.items()), ab)):
        if label != "default" and not ab: continue
        for k, v in env.items(): os.environ[k] = v
        for _ in range(2): step()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(3): step()
        e.record(); torch.cuda.synchronize()
        ms = a.elapsed_time(e) / 3
        _lib.prof_enable(True); torch.cuda.synchronize()
        step(); torch.cuda.synchronize()
        p = _lib.prof_collect(); _lib.prof_enable(False)
        for k in env: os.environ.pop(k)
        print(f"{name} [{label}]: n={n} {ms:.1f} ms/step -> {n / ms / 1e3:.2f} Mpts/s")
        print("   ", {k: (round(v[0], 1), v[1]) for k, v in p.items() if v[1]}, "sum", round(sum(v[0] for v in p.values()), 1))


if "c3" in which: run("C3 kdv/resnet 6x256 (loss)", "kdv", "resnet", 256, 6, 1, 1 << 18, "loss", {"num_blocks": 6})
if "c4" in which: run("C4 ch2d/siren 5x256 math (mse)", "cahn_hilliard", "siren", 256, 5, 2, 1 << 17, "mse", {"omega_0": 30.0}, "math")
if "c4w" in which: run("C4 ch2d/siren 5x256 as written (mse)", "cahn_hilliard", "siren", 256, 5, 2, 1 << 20, "mse", {"omega_0": 30.0})
if "c1" in which: run("C1 heat/fourier 4x128 (loss, 1M)", "heat", "fourier", 128, 4, 1, 1 << 20, "loss", {"mapping_size": 32, "scale": 10.0})
if "c2" in which: run("C2 burgers/ff 8x128 (loss)", "burgers", "feedforward", 128, 8, 1, 1 << 20, "loss", {})

"""Per-kernel-class CUDA-event time of C5 scoring (Allen-Cahn, feedforward 8x128, forward-only residual scoring)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
import torch
import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib
from helpers import product_pde
dev = torch.device('cuda:0')
torch.manual_seed(0)
model = pk.make_model("feedforward", 2, 128, 8, dev)
pde = product_pde("allen_cahn", dev, 1)
n = 1 << 22
x = torch.rand(n, 1, device=dev) * 2 - 1; t = torch.rand(n, 1, device=dev)
def step():
    return pde.score_residual(model, x, t, want_abs=True)
for _ in range(3): step()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(5): step()
e.record(); torch.cuda.synchronize()
ms = a.elapsed_time(e) / 5
print(f"score {n} pts: {ms:.2f} ms -> {n / ms / 1e3:.1f} Mpts/s")
_lib.prof_enable(True); torch.cuda.synchronize()
step(); torch.cuda.synchronize()
p = _lib.prof_collect(); _lib.prof_enable(False)
print({k: (round(v[0], 2), v[1]) for k, v in p.items() if v[1]}, "sum", round(sum(v[0] for v in p.values()), 2))

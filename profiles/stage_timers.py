"""Where does a rows-kernel CTA wait?  Needs a library built with PINNK_NVCC_EXTRA=-DPINNK_STAGE_TIMERS."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
import torch
from pinns_rl_pde_b200 import _lib
dev = torch.device('cuda:0')
M, K, N, C = 4 * (1 << 17), 128, 128, 4
X = torch.randn(M, K, device=dev); dZ = torch.randn(M, N, device=dev)
W = torch.randn(N, K, device=dev) / K ** 0.5; b = torch.randn(N, device=dev)
def show(tag, which, fn):
    for _ in range(2): fn()
    torch.cuda.synchronize(); _lib.stage_timers(which)
    fn(); torch.cuda.synchronize()
    rc, t = _lib.stage_timers(which)
    n = max(1, t["tiles"])
    print(tag, "rc", rc, {k: round(v / n) for k, v in t.items() if k != "tiles"}, "tiles", t["tiles"])
    L = max(1, t["launches"])
    print("   per launch: loop %.1f us, prologue %.1f us, kernel %.1f us; slowest CTA loop of any launch %.1f us" % (t["total_ns"] / L / 1e3, t["prologue_ns"] / L / 1e3, t["kernel_ns"] / L / 1e3, t["max_loop_ns"] / 1e3))
show("fwd plain", 0, lambda: _lib.debug_linear_fwd(X, W, b, C, 1))
show("dgrad plain", 1, lambda: _lib.debug_linear_dgrad(dZ, W, 1))
show("wgrad (32-row tiles)", 2, lambda: _lib.debug_linear_wgrad(dZ, X, C, 1))

# the fused kernels inside a real C2 step (Linear+tanh jets forward, dgrad+tanh adjoint)
sys.path.insert(0, os.path.join(_R, 'tests'))
import pinns_rl_pde_b200 as pk
from helpers import product_pde
torch.manual_seed(0)
model = pk.make_model("feedforward", 2, 128, 8, dev)
pde = product_pde("burgers", dev, 1)
n = 1 << 19
x = torch.rand(n, 1, device=dev); t = torch.rand(n, 1, device=dev)
def step():
    model.zero_grad(set_to_none=True)
    pde.compute_loss(model, x, t)["total"].backward()
for _ in range(2): step()
torch.cuda.synchronize(); _lib.stage_timers(0); _lib.stage_timers(1)
step(); torch.cuda.synchronize()
for which, tag in ((0, "step fwd (Linear+tanh jets)"), (1, "step bwd (dgrad+tanh adjoint)")):
    rc, tm = _lib.stage_timers(which)
    nt = max(1, tm["tiles"])
    print(tag, {k: round(v / nt) for k, v in tm.items() if k != "tiles"}, "tiles", tm["tiles"])
    L = max(1, tm["launches"])
    print("   per launch: loop %.1f us, prologue %.1f us, kernel %.1f us, launches %d; slowest CTA loop of any launch %.1f us" % (tm["total_ns"] / L / 1e3, tm["prologue_ns"] / L / 1e3, tm["kernel_ns"] / L / 1e3, L, tm["max_loop_ns"] / 1e3))

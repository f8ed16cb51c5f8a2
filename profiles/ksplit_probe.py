"""One 256 x 256 Linear over 1.31 M rows (C3: 262 144 points x 5 jet columns): forward and dgrad GEMM as two K-half passes
vs the K-split launch, L2 flushed between timed launches.  `python profiles/ksplit_probe.py once` runs each variant once
(for ncu)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
import torch
from pinns_rl_pde_b200 import _lib
os.environ.setdefault("PINNK_ENABLE_KSPLIT", "1")
dev = torch.device('cuda:0')
M = 262144 * 5
g = torch.Generator(device='cuda').manual_seed(0)
X = torch.randn(M, 256, generator=g, device=dev)
W = torch.randn(256, 256, generator=g, device=dev) / 16
b = torch.randn(256, generator=g, device=dev)
ring = torch.empty(74 * 8 * 8192, device=dev)
Z = torch.empty(M, 256, device=dev)
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
once = len(sys.argv) > 1


def timed(fn, reps=5):
    if once:
        fn(); torch.cuda.synchronize(); return 0.0
    for _ in range(2): fn()
    torch.cuda.synchronize(); tot = 0.0
    for _ in range(reps):
        flush.zero_(); torch.cuda.synchronize()
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); e.record(); torch.cuda.synchronize(); tot += a.elapsed_time(e)
    return tot / reps


for trans in (False, True):
    name = "dgrad" if trans else "fwd"
    t2 = timed(lambda: _lib.debug_linear_ks(X, W, None if trans else b, 5, trans, None, Z))
    z2 = Z.clone()
    tk = timed(lambda: _lib.debug_linear_ks(X, W, None if trans else b, 5, trans, ring, Z))
    same = bool(torch.equal(z2, Z))
    flops = 2.0 * M * 256 * 256
    t2 = max(t2, 1e-9)
    print(f"{name}: two passes {t2:.3f} ms ({flops / t2 / 1e9:.0f} TFLOP/s) | K-split {tk:.3f} ms ({flops / max(tk, 1e-9) / 1e9:.0f} TFLOP/s) | "
          f"bit-identical {same}", flush=True)

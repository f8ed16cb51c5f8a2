"""Error of the 3xTF32 tcgen05 GEMM against fp64, with its SIGN: norm-wise relative error, and the mean signed relative error
on all-positive operands (where a truncating accumulator shows up as a systematic shrink).  CUDA-core fp32 GEMM alongside."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
import torch
from pinns_rl_pde_b200 import _lib
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
for K in (128, 256):
    for kind in ("signed", "positive"):
        M = 1 << 16
        X = torch.randn(M, K, generator=g, device=dev)
        W = torch.randn(K if K == 128 else 256, K, generator=g, device=dev) / K ** 0.5
        if kind == "positive":
            X, W = X.abs(), W.abs()
        ref = X.double() @ W.double().t()
        for mode, name in ((1, "tcgen05 3xTF32"), (0, "CUDA-core fp32")):
            Y = _lib.debug_linear_fwd(X, W, None, 1, mode).double()
            err = float((Y - ref).norm() / ref.norm())
            signed = float(((Y - ref) / ref.abs().clamp_min(1e-30)).mean()) if kind == "positive" else float("nan")
            print(f"K={K} {kind:8s} {name:15s}: rel err {err:.3e}   mean signed rel err {signed:+.3e}", flush=True)

"""Is each tcgen05 rows kernel bit-reproducible from launch to launch?  Runs a kernel several times on identical inputs and
reports differing elements (count, where, how much)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
import torch
from pinns_rl_pde_b200 import _lib
dev = torch.device('cuda:0')
P, JC = int(sys.argv[1]) if len(sys.argv) > 1 else (1 << 18), 4
M = P * JC
g = torch.Generator(device='cuda').manual_seed(0)
Y = torch.empty(P, JC, 128, device=dev)
BENIGN = os.environ.get("PROBE_BENIGN", "0") == "1"
if BENIGN:      # moderate units: |y0| < 0.76, 1/w0 < 2.4 -- no amplification in the tanh adjoint
    Y[:, 0] = torch.tanh(torch.rand(P, 128, generator=g, device=dev) * 2 - 1); Y[:, 1:] = torch.randn(P, JC - 1, 128, generator=g, device=dev) * 0.3
else:
    Y[:, 0] = torch.tanh(torch.randn(P, 128, generator=g, device=dev) * 1.5); Y[:, 1:] = torch.randn(P, JC - 1, 128, generator=g, device=dev) * 0.3
Y = Y.reshape(M, 128).contiguous()
dZ = torch.randn(M, 128, generator=g, device=dev); W = torch.randn(128, 128, generator=g, device=dev) / 11.3
bias = torch.randn(128, generator=g, device=dev)

def report(name, outs):
    ref = outs[0]
    for i, o in enumerate(outs[1:], 1):
        ne = (o != ref)
        n = int(ne.sum())
        if n == 0:
            print(f"{name}: run {i} identical to run 0")
            continue
        idx = ne.nonzero()[:8].tolist()
        d = (o - ref).abs()
        rows = ne.any(dim=1).nonzero().reshape(-1)
        print(f"{name}: run {i} differs in {n} elements of {ref.numel()} ({rows.numel()} rows; first rows {rows[:6].tolist()}, row%64 {[(r % 64) for r in rows[:12].tolist()]}); "
              f"max abs diff {d.max().item():.3e} (max |ref| {ref.abs().max().item():.3e}); sample idx {idx[:4]}")
        r, c = idx[0]
        print(f"    e.g. [{r},{c}]: {ref[r, c].item():.9e} vs {o[r, c].item():.9e}; cols differing in that row: {int(ne[r].sum())}")

report("dgrad+tanh adjoint (EPI_ACTBWD_Y)", [_lib.debug_bwd_layer(dZ, W, Y, 2, 1, False)[0] for _ in range(4)])
print("env:", {k: v for k, v in os.environ.items() if k.startswith(("PINNK_", "PROBE_"))})
if os.environ.get("PROBE_ONLY_ADJ"):
    sys.exit(0)
report("plain dgrad (EPI_PLAIN, TRANS_W)", [_lib.debug_linear_dgrad(dZ, W, 1) for _ in range(4)])
report("plain forward (EPI_PLAIN)", [_lib.debug_linear_fwd(dZ, W, bias, JC, 1) for _ in range(4)])
report("paired kernel dZprev", [_lib.debug_bwd_layer(dZ, W, Y, 2, 1, True)[0] for _ in range(3)])
a = _lib.debug_bwd_layer(dZ, W, Y, 2, 1, False)[0]
# fp64 reference of the plain dgrad on a slice
ref = (dZ[:4096].double() @ W.double())
got = _lib.debug_linear_dgrad(dZ, W, 1)[:4096].double()
print("plain dgrad vs fp64 (first 4096 rows): rel", float((got - ref).norm() / ref.norm()))

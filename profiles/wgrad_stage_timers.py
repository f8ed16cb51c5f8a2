import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pinns_rl_pde_b200 import _lib
dev = torch.device('cuda:0')
M, K, N, C = 4 * (1 << 20), 128, 128, 4
X = torch.randn(M, K, device=dev); dZ = torch.randn(M, N, device=dev)
def fn(): _lib.debug_linear_wgrad(dZ, X, C, 1)
for _ in range(2): fn()
torch.cuda.synchronize(); _lib.stage_timers(2)
fn(); torch.cuda.synchronize()
rc, t = _lib.stage_timers(2)
n = max(1, t["tiles"])
print("wgrad TS", "rc", rc, {k: round(v / n) for k, v in t.items() if k != "tiles"}, "tiles", t["tiles"])

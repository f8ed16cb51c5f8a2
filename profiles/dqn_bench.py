"""RL sampler Q-network over the 100 x 100 candidate grid (rl_agent.py:15-88,214-229): pinnk_dqn_forward (one launch) against the
same torch modules, eval mode and with live dropout (the reference's default)."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R)
import torch
from pinns_rl_pde_b200 import rl
dev = torch.device('cuda:0')
torch.manual_seed(0)
def timed(fn, reps=200):
    for _ in range(20): fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3
for hidden, n in ((128, 10000), (128, 1000000), (256, 10000), (512, 10000), (512, 1000000)):
    net = rl.DQNNetwork(2, 1, hidden).to(dev)
    twin = torch.nn.Sequential(*[torch.nn.Sequential(*list(g)) if isinstance(g, torch.nn.Sequential) else g for g in net.layers])
    x = torch.rand(n, 2, device=dev)
    for mode in ("eval", "train"):
        net.train(mode == "train"); twin.train(mode == "train")
        with torch.no_grad():
            us_lib = timed(lambda: rl.dqn_forward(net, x))
            us_ref = timed(lambda: twin(x))
        print(f"hidden={hidden} n={n} {mode}: libpinnk {us_lib:.1f} us ({n / us_lib:.1f} M states/s), torch modules {us_ref:.1f} us "
              f"-> x{us_ref / us_lib:.2f}")

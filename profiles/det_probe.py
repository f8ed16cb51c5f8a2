"""Which parameter gradients differ between two identical PINNK_DETERMINISTIC=1 steps?"""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import torch
import pinns_rl_pde_b200 as pk
from helpers import product_pde
os.environ["PINNK_DETERMINISTIC"] = "1"
dev = torch.device("cuda:0")
torch.manual_seed(0)
ARCH = sys.argv[1] if len(sys.argv) > 1 else "feedforward"
if ARCH == "siren":
    model = pk.make_model("siren", 2, 256, 4, dev, omega_0=30.0)
    pde = product_pde("allen_cahn", dev)
elif ARCH == "resnet":
    model = pk.make_model("resnet", 2, 256, 2, dev, num_blocks=2)
    pde = product_pde("kdv", dev)
else:
    model = pk.make_model("feedforward", 2, 128, 8, dev)
    pde = product_pde("burgers", dev)
g = torch.Generator().manual_seed(1)
n = 60000 if ARCH != 'feedforward' else 200000
x, t = (torch.rand(n, 1, generator=g) * 2 - 1).to(dev), torch.rand(n, 1, generator=g).to(dev)
def run():
    model.zero_grad()
    pde.compute_loss(model, x, t)["total"].backward()
    torch.cuda.synchronize()
    return [p.grad.clone() for p in model.parameters()]
a, b = run(), run()
for (name, _), u, v in zip(model.named_parameters(), a, b):
    d = (u - v).abs().max().item()
    print(f"{name:28s} {tuple(u.shape)} max|diff| {d:.3e} rel {d / max(u.abs().max().item(), 1e-30):.2e}")
r1 = pde.compute_residual(model, x, t).detach(); r2 = pde.compute_residual(model, x, t).detach()
print("residual identical:", torch.equal(r1, r2))

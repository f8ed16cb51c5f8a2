import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import torch
import pinns_rl_pde_b200 as pk
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
print("env:", {k: v for k, v in os.environ.items() if k.startswith("PINNK_")})
for arch, hidden, layers, extra in (("siren", 256, 3, {"omega_0": 30.0}), ("siren", 128, 3, {"omega_0": 30.0}), ("feedforward", 256, 3, {}), ("feedforward", 128, 3, {})):
    torch.manual_seed(0)
    model = pk.make_model(arch, 2, hidden, layers, dev, **extra)
    xt = torch.rand(60000, 2, generator=g, device=dev)
    with torch.no_grad():
        for name, fn in (("value", lambda: model(xt)), ("jets(2,1)", lambda: pk.jets(model, xt, [((1.0, 0.0), 2), ((0.0, 1.0), 1)]))):
            outs = [fn() for _ in range(4)]
            diffs = [int((o != outs[0]).sum()) for o in outs[1:]]
            mx = max(float((o - outs[0]).abs().max()) for o in outs[1:])
            print(f"{arch} {layers}x{hidden} {name}: differing elements vs run 0: {diffs}, max abs diff {mx:.3e}", flush=True)

import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import torch
import pinns_rl_pde_b200 as pk
dev = torch.device('cuda:0')
g = torch.Generator(device='cuda').manual_seed(0)
print("env:", {k: v for k, v in os.environ.items() if k.startswith("PINNK_")})
torch.manual_seed(0)
model = pk.make_model("siren", 2, 256, 2, dev, omega_0=30.0)
xt = torch.rand(60000, 2, generator=g, device=dev)
out_lin = [m for m in model.modules() if isinstance(m, torch.nn.Linear)][-1]
rows_by_f = {}
with torch.no_grad():
    for f in (0, 5, 31, 32, 64, 100, 127, 128, 129, 200, 255):
        out_lin.weight.zero_(); out_lin.weight[0, f] = 1.0; out_lin.bias.zero_()
        outs = [model(xt) for _ in range(6)]
        bad = set()
        for o in outs[1:]:
            bad |= set((o != outs[0]).reshape(-1).nonzero().reshape(-1).tolist())
        rows_by_f[f] = sorted(bad)
        ex = ""
        if bad:
            r = sorted(bad)[0]
            ex = " e.g. row %d values %s" % (r, [round(float(o[r]), 5) for o in outs])
        print(f"feature {f:3d}: rows differing in any of 5 reruns: {len(bad)} {sorted(bad)[:12]} tile-row {[r % 64 for r in sorted(bad)[:12]]}{ex}", flush=True)

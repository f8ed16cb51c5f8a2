"""C2 step time (compute_loss + backward, 1M points) as a function of the chunk size (PINNK_MAX_CHUNK_POINTS)."""
import os, sys, subprocess
if len(sys.argv) > 1:
    _R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, 'tests'))
    import torch
    import pinns_rl_pde_b200 as pk
    from helpers import product_pde
    dev = torch.device('cuda:0')
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 128, 8, dev)
    pde = product_pde("burgers", dev, 1)
    n = 1 << 20
    x = torch.rand(n, 1, device=dev); t = torch.rand(n, 1, device=dev)
    def step():
        model.zero_grad(set_to_none=True)
        pde.compute_loss(model, x, t)["total"].backward()
    for _ in range(3): step()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(5): step()
    e.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(e) / 5
    print(f"chunk={sys.argv[1]}: {ms:.2f} ms/step -> {n / ms / 1e3:.2f} Mpts/s", flush=True)
else:
    for c in (4096, 8192, 16384, 32768, 65536, 131072):
        env = dict(os.environ, PINNK_MAX_CHUNK_POINTS=str(c))
        subprocess.run([sys.executable, __file__, str(c)], env=env)

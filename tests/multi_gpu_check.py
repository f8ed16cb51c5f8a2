"""Run under torchrun on N GPUs: the sharded (NCCL) loss/gradient equals the single-GPU result.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py
"""
import math
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def main():
    import pinns_rl_pde_b200 as pk
    from pinns_rl_pde_b200 import parallel
    from helpers import flat_grad, product_pde
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 128, 4, dev)
    pde = product_pde("burgers", dev)
    g = torch.Generator().manual_seed(1)
    n = 50001
    x, t = (torch.rand(n, 1, generator=g) * 2 - 1).to(dev), torch.rand(n, 1, generator=g).to(dev)
    losses = parallel.sharded_loss_backward(pde, model, x, t)
    sharded = flat_grad(model).clone()
    model.zero_grad()
    ref = pde.compute_loss(model, x, t)
    ref["total"].backward()
    single = flat_grad(model)
    err = float((sharded - single).norm() / single.norm())
    dl = abs(losses["total"].item() - ref["total"].item()) / abs(ref["total"].item())
    ok = err < 5e-6 and dl < 5e-6
    print(f"rank {dist.get_rank()}/{dist.get_world_size()}: sharded-vs-single grad rel {err:.2e}, loss rel {dl:.2e} -> {'OK' if ok else 'FAIL'}")
    # sharded candidate scoring: global statistics equal the single-GPU ones; the two-level RAR draw returns the request
    ac = product_pde("allen_cahn", dev)
    lo, hi = parallel.shard_bounds(n)
    _, st_sh = parallel.sharded_score(ac, model, x[lo:hi], t[lo:hi])
    _, st_1 = ac.score_residual(model, x, t)
    ds = float(((st_sh - st_1).abs() / st_1.abs().clamp_min(1e-30)).max())
    xs, ts = parallel.sharded_residual_sample(ac, model, x[lo:hi], t[lo:hi], 4096, gather=True)
    ok2 = ds < 1e-6 and xs.shape == (4096, 1) and ts.shape == (4096, 1)
    print(f"rank {dist.get_rank()}: sharded score stats rel {ds:.2e}, RAR selection {tuple(xs.shape)} -> {'OK' if ok2 else 'FAIL'}")
    ok = ok and ok2
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

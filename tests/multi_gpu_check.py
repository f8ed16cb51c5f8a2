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
    # Heat with the smoothness regulariser (heat_equation.py:625-650, weight 0.1 as the reference's YAML ships it): the fused
    # data-parallel trainer step (rows sharded, one in-place all-reduce of [gradient || 4 loss sums]) follows the single-process one
    import copy
    training = {"num_collocation_points": 5000, "num_boundary_points": 200, "num_initial_points": 200,
                "loss_weights": {"residual": 1.0, "boundary": 10.0, "initial": 10.0, "smoothness": 0.1}}
    cfgh = pk.PDEConfig(name="heat", domain=[[0.0, 1.0]], time_domain=[0.0, 1.0], parameters={"alpha": 0.01},
                        boundary_conditions={"dirichlet": {"type": "dirichlet"}},
                        initial_condition={"type": "sine", "amplitude": 1.0, "frequency": 2.0},
                        exact_solution={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0}, dimension=1, device=dev,
                        training=training)
    heat = pk.HeatEquation(cfgh)
    torch.manual_seed(2)
    mh = pk.make_model("fourier", 2, 128, 3, dev, mapping_size=32, scale=10.0)
    m1 = copy.deepcopy(mh)
    tcfg = pk.TrainingConfig(learning_rate=1e-3, weight_decay=0.0, gradient_clipping=1.0, scheduler="none",
                             loss_weights=dict(training["loss_weights"]))
    tr_sh = pk.PDETrainer(mh, heat, config=tcfg, device=dev, fused=True)
    nh = 20000
    xh, th = torch.rand(nh, 1, generator=g).to(dev), torch.rand(nh, 1, generator=g).to(dev)
    lo, hi = parallel.shard_bounds(nh)
    l_sh = tr_sh.train_step(xh[lo:hi], th[lo:hi], n_global=nh)
    # single-process reference of the same step: loss_step_flat on all rows (no collective), then the same optimizer step
    from pinns_rl_pde_b200 import functional as F
    comp, wts, flat = F.loss_step_flat(heat, m1, xh, th)
    tot1 = float(wts[0] * comp[0] + wts[1] * comp[1] + wts[2] * comp[2] + wts[3] * comp[3])
    dl2 = abs(float(l_sh["total"]) - tot1) / abs(tot1)
    ds2 = abs(float(l_sh["smoothness"]) - float(comp[3])) / abs(float(comp[3]))
    ok3 = dl2 < 2e-5 and ds2 < 2e-5 and float(comp[3]) > 0
    print(f"rank {dist.get_rank()}: sharded Heat + smoothness step: total rel {dl2:.2e}, smoothness rel {ds2:.2e} -> {'OK' if ok3 else 'FAIL'}")
    ok = ok and ok3
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()

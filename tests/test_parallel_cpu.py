"""CPU, world_size 2 over gloo: the data-parallel host logic (row sharding, the single fused all-reduce of
[flat gradient || loss sums], re-weighting of shard means) reproduces the single-process loss and gradient.
The per-shard loss components come from the oracle here (there is no GPU); on the B200 box the same code
path runs with functional.loss_components on libpinnk."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from helpers import PDES


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _oracle_components(pde, model, x, t, n_global=None):
    """Stand-in for functional.loss_components built on the CPU oracle (same return contract)."""
    from oracle import ref_port
    s = PDES["burgers"]
    r = ref_port.burgers_residual(model, x, t, nu=s["params"]["nu"]) if x.shape[0] else torch.zeros(0, 1)
    fns = ref_port.boundary_condition_fns("burgers", s["bcs"], s["ic"], s["domain"], s["params"])
    if x.shape[0]:
        L = ref_port.base_compute_loss(model, r, s["domain"], s["time"], fns)
    else:
        dummy = model(torch.zeros(1, 2)) * 0.0
        L = ref_port.base_compute_loss(model, dummy, s["domain"], s["time"], fns)
        L["residual"] = dummy.sum()
    comp = torch.stack([L["residual"], L["boundary"], L["initial"]])
    return comp, (1.0, 10.0, 10.0, 0.0, False)


def _worker(rank, world, port, n, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import ref_port
        from pinns_rl_pde_b200 import parallel
        torch.manual_seed(0)
        model = ref_port.PINNModel("feedforward", 2, 16, 2)
        g = torch.Generator().manual_seed(1)
        x, t = torch.rand(n, 1, generator=g) * 2 - 1, torch.rand(n, 1, generator=g)
        lo, hi = parallel.shard_bounds(n)
        assert parallel.world_size() == world and parallel.rank() == rank
        losses = parallel.sharded_loss_backward(None, model, x, t, components=_oracle_components)
        flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
        if rank == 0:
            torch.save({"flat": flat, "losses": {k: float(v) for k, v in losses.items()}, "bounds": (lo, hi)}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [101, 1])
def test_sharded_loss_backward_matches_single_process(tmp_path, n):
    from oracle import ref_port
    out = str(tmp_path / "r0.pt")
    mp.spawn(_worker, args=(2, _free_port(), n, out), nprocs=2, join=True)
    got = torch.load(out)
    torch.manual_seed(0)
    model = ref_port.PINNModel("feedforward", 2, 16, 2)
    g = torch.Generator().manual_seed(1)
    x, t = torch.rand(n, 1, generator=g) * 2 - 1, torch.rand(n, 1, generator=g)
    comp, (w_r, w_b, w_i, _, _) = _oracle_components(None, model, x, t)
    total = w_r * comp[0] + w_b * comp[1] + w_i * comp[2]
    total.backward()
    flat = torch.cat([p.grad.reshape(-1) for p in model.parameters()])
    assert torch.allclose(got["flat"], flat, rtol=2e-5, atol=1e-7)
    assert abs(got["losses"]["total"] - float(total)) <= 2e-6 * abs(float(total))
    assert abs(got["losses"]["residual"] - float(comp[0])) <= 2e-6 * abs(float(comp[0])) + 1e-12


def test_shard_bounds_cover_rows_exactly():
    from pinns_rl_pde_b200 import parallel
    for n in (0, 1, 7, 8, 1000003):
        for w in (1, 2, 3, 8):
            b = [parallel.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [hi - lo for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1


def test_multinomial_large_matches_distribution():
    """Two-level draw used for candidate pools beyond torch.multinomial's 2^24 limit."""
    from pinns_rl_pde_b200.pdes import multinomial_large
    torch.manual_seed(0)
    w = torch.rand(5000) ** 3
    sel = multinomial_large(w, 200000, block=512)            # small blocks force the two-level path ... only above 2^24
    assert sel.min() >= 0 and sel.max() < 5000
    n = (1 << 24) + 4096
    w2 = torch.zeros(n)
    hot = torch.tensor([5, 1 << 23, n - 3])
    w2[hot] = torch.tensor([1.0, 2.0, 3.0])
    sel2 = multinomial_large(w2, 60000)
    counts = torch.stack([(sel2 == h).sum() for h in hot]).double()
    assert counts.sum() == 60000
    assert torch.allclose(counts / 60000, torch.tensor([1 / 6, 2 / 6, 3 / 6], dtype=torch.float64), atol=0.01)

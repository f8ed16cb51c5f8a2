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


# ---------------------------------------------------------------------------------- sharded residual scoring / RAR
def _stub_score(pde, model, x, t, want_abs=True):
    """Stand-in for functional.score_residual (same return contract): r = x^3 * (1 + t)."""
    r = (x.double() ** 3 * (1 + t.double())).reshape(-1)
    a = r.abs()
    stats = torch.stack([a.sum(), (r * r).sum(), a.max() if a.numel() else torch.zeros((), dtype=torch.float64),
                         torch.tensor(float(a.numel()), dtype=torch.float64)])
    return (a.to(torch.float32) if want_abs else None), stats


def _pool(n):
    g = torch.Generator().manual_seed(5)
    return torch.rand(n, 1, generator=g) * 2 - 1, torch.rand(n, 1, generator=g)


def _score_worker(rank, world, port, n, k, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from pinns_rl_pde_b200 import parallel
        x, t = _pool(n)
        lo, hi = parallel.shard_bounds(n)
        mag, stats = parallel.sharded_score(None, None, x[lo:hi], t[lo:hi], score_fn=_stub_score)
        assert mag.shape[0] == hi - lo
        torch.manual_seed(100 + rank)
        xs, ts = parallel.sharded_residual_sample(None, None, x[lo:hi], t[lo:hi], k, gather=True,
                                                  generator=torch.Generator().manual_seed(9), score_fn=_stub_score)
        xl, tl = parallel.sharded_residual_sample(None, None, x[lo:hi], t[lo:hi], k, gather=False,
                                                  generator=torch.Generator().manual_seed(9), score_fn=_stub_score)
        cnt = torch.tensor([xl.shape[0]])
        dist.all_reduce(cnt)
        assert int(cnt) == k                                  # the per-rank draws add up to the request
        if rank == 0:
            torch.save({"stats": stats, "xs": xs, "ts": ts}, out)
    finally:
        dist.destroy_process_group()


def test_sharded_score_and_two_level_rar_sampling(tmp_path):
    n, k = 20001, 40000
    out = str(tmp_path / "score.pt")
    mp.spawn(_score_worker, args=(2, _free_port(), n, k, out), nprocs=2, join=True)
    got = torch.load(out)
    x, t = _pool(n)
    mag, stats = _stub_score(None, None, x, t)
    assert torch.allclose(got["stats"], stats, rtol=1e-12)                     # global statistics = single-process ones
    xs, ts = got["xs"], got["ts"]
    assert xs.shape == (k, 1) and ts.shape == (k, 1)
    # every selected point is a pool point, and the selection follows |r| + 1e-8: compare the mass drawn from the
    # region |x| > 0.8 (which carries most of the |x|^3 weight) with its exact probability
    pool = {(round(float(a), 6), round(float(b), 6)) for a, b in zip(x[:, 0], t[:, 0])}
    assert all((round(float(a), 6), round(float(b), 6)) in pool for a, b in zip(xs[:200, 0], ts[:200, 0]))
    w = mag.double() + 1e-8
    p_hot = float(w[(x[:, 0].abs() > 0.8)].sum() / w.sum())
    f_hot = float((xs[:, 0].abs() > 0.8).double().mean())
    assert abs(f_hot - p_hot) < 0.01, (f_hot, p_hot)


def test_sharded_score_single_process_is_plain_score():
    from pinns_rl_pde_b200 import parallel
    x, t = _pool(1000)
    mag, stats = parallel.sharded_score(None, None, x, t, score_fn=_stub_score)
    m2, s2 = _stub_score(None, None, x, t)
    assert torch.equal(mag, m2) and torch.equal(stats, s2)
    xs, ts = parallel.sharded_residual_sample(None, None, x, t, 50, score_fn=_stub_score)
    assert xs.shape == (50, 1) and ts.shape == (50, 1)

"""Data parallelism over collocation rows: one process per GPU, rows sharded by
rank, weights replicated, ONE all-reduce per step of the flat buffer [grad || loss sums].

The collective is torch.distributed (NCCL over NVLink on the B200 box, gloo in the CPU tests); the
payload is 0.17-3.2 MB so the step is latency-bound and a single fused buffer is the right shape.
"""
from __future__ import annotations

from typing import Callable, Dict, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def shard_bounds(n: int, r: Optional[int] = None, w: Optional[int] = None) -> Tuple[int, int]:
    """Rows [lo, hi) of an n-row set owned by rank r of w: contiguous, sizes differ by at most one."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    lo = r * base + min(r, rem)
    return lo, lo + base + (1 if r < rem else 0)


# Optional timing of the step's collective (bench.py: `collective_ms`): when TIMING is a list, every all-reduce of the
# gradient buffer appends a (start, end) pair of CUDA events recorded on the current stream around it.
TIMING: Optional[list] = None


def _timed_all_reduce(buf: torch.Tensor, group=None):
    if TIMING is not None and buf.is_cuda:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)
        b.record()
        TIMING.append((a, b))
    else:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)


def reduce_inplace(buf: torch.Tensor, group=None) -> torch.Tensor:
    """All-reduce(sum) IN PLACE of a step buffer laid out [flat_grad || loss sums] (the fused step writes its three loss
    sums into the buffer's tail, so there is no concatenation and no copy around the collective)."""
    if world_size() > 1:
        _timed_all_reduce(buf, group)
    return buf


def reduce_flat(local_flat: torch.Tensor, local_sums: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-reduce(sum) of [flat_grad || loss sums] in one call; returns the two views."""
    if world_size() == 1:
        return local_flat, local_sums
    buf = torch.cat([local_flat.reshape(-1), local_sums.reshape(-1).to(local_flat.dtype)])
    _timed_all_reduce(buf, group)
    n = local_flat.numel()
    return buf[:n].view_as(local_flat), buf[n:].view_as(local_sums)


def sharded_loss_backward(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor,
                          components: Optional[Callable] = None, group=None,
                          n_global: Optional[int] = None) -> Dict[str, torch.Tensor]:
    """Data-parallel compute_loss + backward (trainer.py:578,689 on N GPUs).

    With ``n_global=None``, ``x, t`` are the GLOBAL collocation rows (identical on every rank, e.g. drawn
    from a shared seed) and rank r takes rows [lo_r, hi_r); with ``n_global`` given, ``x, t`` already are
    this rank's shard.  The few hundred boundary/initial rows are evaluated on every rank (replicated
    work, no exchange).  With frac_r = n_r / N the global objective is
        total = sum_r [ frac_r * w_res * res_r + (w_bc * bc + w_ic * ic) / W ]
    so each rank differentiates its bracket locally and ONE all-reduce(sum) of
    [flat_grad || frac_r*res_r || bc/W || ic/W] yields the global gradient and the global components.
    ``param.grad`` is set to views of the reduced buffer.  ``components`` (tests only) replaces
    functional.loss_components."""
    from . import functional as F
    comp_fn = components or F.loss_components
    if components is None:
        F._require_physics_only(pde, "sharded_loss_backward (data-parallel step)")
    w, r = world_size(), rank()
    if n_global is None:
        n = x.shape[0]
        lo, hi = shard_bounds(n, r, w)
        x, t = x[lo:hi], t[lo:hi]
    else:
        n = int(n_global)
        lo, hi = 0, x.shape[0]
    comp, (w_res, w_bc, w_ic, w_smooth, adaptive) = comp_fn(pde, model, x, t, n_global=n)
    if adaptive:
        w_res = w_bc = w_ic = 1.0
    frac = (hi - lo) / max(n, 1)
    local = frac * w_res * comp[0] + (w_bc * comp[1] + w_ic * comp[2]) / w
    # Heat's finite-difference smoothness term (heat_equation.py:625-650) is a mean over the collocation rows, so it shards
    # like the residual: each rank evaluates it on its rows and contributes frac_r * smooth_r
    smooth = None
    if w_smooth and F.pde_name(pde) == "heat" and components is None:
        xs, ts = F._prep(model, x, t)
        smooth = F._heat_smoothness(pde, model, xs, ts)
        local = local + frac * w_smooth * smooth
    params = [p for p in model.parameters() if p.requires_grad]
    grads = torch.autograd.grad(local, params, allow_unused=True)
    flat = torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1) for p, g in zip(params, grads)])
    sm = frac * smooth.detach() if smooth is not None else torch.zeros((), device=flat.device, dtype=comp.dtype)
    sums = torch.stack([frac * comp[0].detach(), comp[1].detach() / w, comp[2].detach() / w, sm.to(comp.dtype)]).to(flat.dtype)
    flat, sums = reduce_flat(flat, sums, group)
    off = 0
    for p in params:
        p.grad = flat[off:off + p.numel()].view_as(p)
        off += p.numel()
    zero = torch.zeros((), device=flat.device)
    total = w_res * sums[0] + w_bc * sums[1] + w_ic * sums[2]
    if smooth is not None:
        total = total + w_smooth * sums[3]
    return {"residual": sums[0], "boundary": sums[1], "initial": sums[2], "smoothness": sums[3], "data": zero,
            "total": total}


# ---------------------------------------------------------------------------------- sharded residual scoring
def sharded_score(pde, model: nn.Module, x: torch.Tensor, t: torch.Tensor, want_abs: bool = True, group=None,
                  score_fn: Optional[Callable] = None):
    """Forward-only residual scoring of a candidate set sharded over the ranks (the callers are
    pde_base.py:895-935 RAR, :1364-1377 the RL reward, trainer.py:210-262 live snapshots).

    ``x, t`` are THIS rank's candidates.  Returns ``(abs_r_local, stats)``: ``|r|`` of the local candidates (or None)
    and the GLOBAL statistics ``[sum|r|, sum r^2, max|r|, count]`` (fp64), reduced with two tiny collectives
    (sum of 3 doubles, max of 1).  The scores themselves never move.  ``score_fn`` (tests only) replaces
    functional.score_residual."""
    from . import functional as F
    fn = score_fn or F.score_residual
    mag, stats = fn(pde, model, x, t, want_abs)
    stats = stats.to(torch.float64).clone()
    if world_size() > 1:
        sums = torch.stack([stats[0], stats[1], stats[3]])
        mx = stats[2:3].clone()
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
        stats = torch.stack([sums[0], sums[1], mx[0], sums[2]])
    return mag, stats


def sharded_residual_sample(pde, model: nn.Module, x_pool: torch.Tensor, t_pool: torch.Tensor, num_points: int,
                            gather: bool = False, generator: Optional[torch.Generator] = None, group=None,
                            score_fn: Optional[Callable] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """Residual-adaptive refinement over a candidate pool sharded over the ranks (pde_base.py:895-935 on N GPUs).

    The reference draws ``num_points`` candidates with probability proportional to ``|r| + 1e-8`` from ONE pool.  Here the
    draw is two-level and exact in distribution: (1) every rank computes the mass of its shard, the W masses are
    all-gathered, and the per-rank sample counts are drawn from the multinomial over shard masses -- by rank 0,
    broadcast, so that all ranks agree; (2) each rank draws its count locally, proportionally to its own scores
    (``multinomial_large``, no 2^24 limit).  Only W doubles and W integers cross the fabric.

    Returns this rank's selected ``(x, t)`` (the data-parallel trainer consumes per-rank shards directly), or with
    ``gather=True`` the whole selection on every rank (``num_points`` rows, rank order)."""
    from .pdes import multinomial_large
    w, r = world_size(), rank()
    mag, _ = sharded_score(pde, model, x_pool, t_pool, True, group, score_fn)
    wts = mag.reshape(-1).to(torch.float32) + 1e-8
    mass = wts.sum(dtype=torch.float64).reshape(1)
    if w > 1:
        masses = [torch.zeros_like(mass) for _ in range(w)]
        dist.all_gather(masses, mass, group=group)
        masses = torch.cat(masses)
        counts = torch.zeros(w, dtype=torch.int64, device=mass.device)
        if r == 0:
            # per-rank counts ~ Multinomial(num_points, masses / sum) as W - 1 conditional binomials (exact in distribution;
            # drawing num_points single categories on the host cost 0.3 s at 16 M points -- more than scoring a 32 M shard)
            p = (masses / masses.sum()).to(torch.float64).cpu()
            rem_n, rem_p, cs = float(num_points), 1.0, []
            for i in range(w - 1):
                q = min(max(float(p[i]) / rem_p, 0.0), 1.0) if rem_p > 0 else 0.0
                c = float(torch.binomial(torch.tensor([rem_n], dtype=torch.float64), torch.tensor([q], dtype=torch.float64),
                                         generator=generator)) if rem_n > 0 else 0.0
                cs.append(int(c))
                rem_n -= c
                rem_p -= float(p[i])
            cs.append(int(rem_n))
            counts = torch.tensor(cs, dtype=torch.int64).to(mass.device)
        dist.broadcast(counts, 0, group=group)
        k = int(counts[r].item())
    else:
        k = int(num_points)
    if k > 0:
        sel = multinomial_large(wts, k)
        xs, ts = x_pool[sel].detach(), t_pool[sel].detach()
    else:
        xs, ts = x_pool[:0].detach(), t_pool[:0].detach()
    if not gather or w == 1:
        return xs, ts
    # variable-size all-gather of the (small) selection: pad to the largest count
    kmax = int(counts.max().item())
    pad = torch.zeros(kmax, x_pool.shape[1] + 1, dtype=x_pool.dtype, device=x_pool.device)
    pad[:k, :-1], pad[:k, -1:] = xs, ts
    bufs = [torch.zeros_like(pad) for _ in range(w)]
    dist.all_gather(bufs, pad, group=group)
    allp = torch.cat([b[:int(counts[i].item())] for i, b in enumerate(bufs)], dim=0)
    return allp[:, :-1].contiguous(), allp[:, -1:].contiguous()

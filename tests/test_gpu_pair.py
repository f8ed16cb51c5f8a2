"""GPU: the paired reverse kernel (bwd_pair_kernel: dgrad + tanh adjoint on one CTA of a cluster, wgrad on the other, one
shared tile stream) against the two-launch route of the same library and against a torch fp64 evaluation of the same
per-layer reverse step (loss["total"].backward() through one Linear(128, 128) + tanh, reference trainer.py:689)."""
import pytest
import torch

import parity_log
from helpers import rel

pytestmark = pytest.mark.gpu


def _inputs(points, jc, dev, seed=0):
    g = torch.Generator(device="cuda").manual_seed(seed)
    M = points * jc
    Y = torch.empty(points, jc, 128, device=dev)
    Y[:, 0] = torch.tanh(torch.randn(points, 128, generator=g, device=dev) * 1.5)      # value column of a tanh output
    if jc > 1:
        Y[:, 1:] = torch.randn(points, jc - 1, 128, generator=g, device=dev) * 0.3
    dZ = torch.randn(M, 128, generator=g, device=dev)
    W = torch.randn(128, 128, generator=g, device=dev) / 128 ** 0.5
    return dZ, W, Y.reshape(M, 128).contiguous()


@pytest.mark.parametrize("k0,k1,points", [(2, 1, 16), (2, 1, 1000), (2, 1, 262144), (1, 0, 4099), (0, 0, 777), (3, 0, 5000),
                                          (2, 1, 1 << 20)])
def test_pair_kernel_matches_two_launch_route(k0, k1, points):
    from pinns_rl_pde_b200 import _lib
    dev = torch.device("cuda:0")
    jc = 1 + k0 + k1
    dZ, W, Y = _inputs(points, jc, dev)
    a = _lib.debug_bwd_layer(dZ, W, Y, k0, k1, pair=True)
    b = _lib.debug_bwd_layer(dZ, W, Y, k0, k1, pair=False)
    torch.cuda.synchronize()
    assert torch.equal(a[0], b[0]), "dZprev differs between the paired and the two-launch route"      # same tile arithmetic
    # the weight gradient is the same sum over a different CTA partition: fp32 round-off apart; both against fp64
    dW64 = dZ.double().t() @ Y.double()
    db64 = dZ.double().reshape(points, jc, 128)[:, 0].sum(0)
    e_pair, e_split = rel(a[1], dW64), rel(b[1], dW64)
    parity_log.log(f"[pair kernel] jets ({k0},{k1}) {points} points: dZprev bit-identical; dW vs fp64 pair {e_pair:.2e} / split "
                   f"{e_split:.2e}; db pair {rel(a[2], db64):.2e}")
    assert e_pair <= 2e-6 and e_split <= 2e-6
    assert rel(a[2], db64) <= 2e-6


def test_pair_kernel_step_speed_and_traffic_note():
    """Timing of one 4 M-row layer step, paired vs two launches (reported, not gated: the clock of the box varies)."""
    from pinns_rl_pde_b200 import _lib
    dev = torch.device("cuda:0")
    dZ, W, Y = _inputs(1 << 20, 4, dev)
    flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)

    def timed(pair):
        for _ in range(2):
            _lib.debug_bwd_layer(dZ, W, Y, 2, 1, pair)
        tot = 0.0
        for _ in range(5):
            flush.zero_()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _lib.debug_bwd_layer(dZ, W, Y, 2, 1, pair)
            b.record()
            torch.cuda.synchronize()
            tot += a.elapsed_time(b)
        return tot / 5
    t_pair, t_split = timed(True), timed(False)
    rows = 4 << 20
    parity_log.log(f"[pair kernel] 4 Mi-row layer reverse step: paired {t_pair:.3f} ms ({rows * 1536 / t_pair / 1e9:.2f} TB/s of "
                   f"1536 B/row), two launches {t_split:.3f} ms ({rows * 2560 / t_split / 1e9:.2f} TB/s of 2560 B/row)")
    assert t_pair > 0 and t_split > 0

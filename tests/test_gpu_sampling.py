"""GPU: on-device samplers (reference pde_base.py:806-935).
  * pinnk_sample_weighted: inverse-CDF draw with replacement -- exact distribution, beyond torch.multinomial's 2^24 limit
  * pinnk_jittered_grid:   bit-identical to the reference's torch ops on the same generator state
  * the RAR sampler end to end (no host round trip in the draw)."""
import math

import pytest
import torch

import parity_log
from helpers import product_pde

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_weighted_draw_small_matches_the_distribution():
    from pinns_rl_pde_b200.pdes import weighted_sample_device
    torch.manual_seed(0)
    w = torch.rand(5000, device=DEV) ** 3
    w[100:200] = 0.0                                   # categories with no mass are never drawn
    m = 2_000_000
    idx = weighted_sample_device(w, 0.0, m)
    assert idx.dtype == torch.int64 and idx.shape == (m,) and int(idx.min()) >= 0 and int(idx.max()) < 5000
    counts = torch.bincount(idx, minlength=5000).double()
    assert float(counts[100:200].sum()) == 0.0
    p = (w.double() / w.double().sum())
    z = (counts - m * p) / torch.sqrt(m * p * (1 - p) + 1e-30)
    worst = float(z[p > 1e-6].abs().max())
    chi2 = float(((counts - m * p) ** 2 / (m * p + 1e-30))[p > 1e-6].sum()) / int((p > 1e-6).sum())
    parity_log.log(f"[device sampler] 5000 categories, 2e6 draws: worst |z| {worst:.2f}, chi^2/dof {chi2:.3f}")
    assert worst < 6.0 and 0.9 < chi2 < 1.1


def test_weighted_draw_beyond_2_24_candidates():
    """2^25 candidates (torch.multinomial raises here): block-level masses and the eps floor are respected."""
    from pinns_rl_pde_b200.pdes import weighted_sample_device
    n = (1 << 25) + 12345                             # ragged tail
    g = torch.Generator(device=DEV).manual_seed(1)
    w = torch.rand(n, device=DEV, generator=g)
    w[: n // 2] *= 3.0                                 # first half carries 3/4 of the mass
    w[n - 5000:] = 0.0                                 # only eps there
    with pytest.raises(RuntimeError):
        torch.multinomial(w, 10, replacement=True)
    m = 4_000_000
    idx = weighted_sample_device(w, 1e-8, m)
    assert int(idx.min()) >= 0 and int(idx.max()) < n
    frac_first = float((idx < n // 2).double().mean())
    want = float(w[: n // 2].double().sum() / (w.double().sum() + 1e-8 * n))
    # 1024 coarse buckets
    edges = torch.linspace(0, n, 1025, device=DEV).long()
    mass = torch.stack([w[edges[i]:edges[i + 1]].double().sum() + 1e-8 * float(edges[i + 1] - edges[i]) for i in range(1024)])
    p = mass / mass.sum()
    counts = torch.bincount(torch.bucketize(idx, edges[1:-1], right=True), minlength=1024).double()
    z = (counts - m * p) / torch.sqrt(m * p * (1 - p))
    parity_log.log(f"[device sampler] 2^25+12345 candidates, 4e6 draws: P(first half) {frac_first:.5f} (exact {want:.5f}), "
                   f"worst bucket |z| {float(z.abs().max()):.2f}, tail draws {int((idx >= n - 5000).sum())}")
    assert abs(frac_first - want) < 5 * math.sqrt(want * (1 - want) / m)
    assert float(z.abs().max()) < 6.0


def test_jittered_grid_is_bit_identical_to_the_torch_ops():
    pde = product_pde("burgers", torch.device(DEV))
    for n in (100, 4097, 250000):
        torch.manual_seed(7)
        x, t = pde._sample_uniform(n)
        torch.manual_seed(7)
        n_side = int(math.sqrt(n))
        xs = torch.linspace(-1.0, 1.0, n_side, device=DEV).reshape(-1, 1)
        ts = torch.linspace(0.0, 1.0, n_side, device=DEV).reshape(-1, 1)
        X, T = torch.meshgrid(xs.squeeze(), ts.squeeze(), indexing="ij")
        xr, tr = X.reshape(-1, 1), T.reshape(-1, 1)
        xr = xr + torch.randn_like(xr) * (1.0 - -1.0) * 0.01
        tr = tr + torch.randn_like(tr) * (1.0 - 0.0) * 0.01
        xr, tr = torch.clamp(xr, -1.0, 1.0), torch.clamp(tr, 0.0, 1.0)
        assert x.shape == xr.shape and torch.equal(x, xr) and torch.equal(t, tr), n
    torch.manual_seed(9)                                 # and the generator is left where the reference's two randn calls leave it
    pde._sample_uniform(900)
    s1 = torch.cuda.get_rng_state(0)
    torch.manual_seed(9)
    torch.randn(900, 1, device=DEV), torch.randn(900, 1, device=DEV)
    assert torch.equal(s1, torch.cuda.get_rng_state(0))


def test_rar_sampler_prefers_high_residual_regions_and_stays_on_device(monkeypatch):
    import pinns_rl_pde_b200 as pk
    dev = torch.device(DEV)
    torch.manual_seed(0)
    model = pk.make_model("feedforward", 2, 64, 3, dev)
    pde = product_pde("burgers", dev)
    # same generator state -> the draw is the reference's torch.multinomial on (|r| + 1e-8) / sum  (pde_base.py:924-931)
    torch.manual_seed(5)
    x, t = pde.generate_collocation_points(4000, strategy="residual_based", model=model)
    assert x.shape == (4000, 1) and t.shape == (4000, 1)
    torch.manual_seed(5)
    xq, tq = pde._sample_uniform(16000)
    mag = pde.score_residual(model, xq, tq)[0].reshape(-1) + 1e-8
    sel = torch.multinomial(mag / mag.sum(), 4000, replacement=True)
    assert torch.equal(x, xq[sel]) and torch.equal(t, tq[sel])
    # the device sampler (the route beyond 2^24 candidates) on the same pool: another draw of the same distribution
    monkeypatch.setenv("PINNK_DEVICE_SAMPLER", "1")
    x, t = pde.generate_collocation_points(4000, strategy="residual_based", model=model)
    monkeypatch.delenv("PINNK_DEVICE_SAMPLER")
    assert x.shape == (4000, 1) and t.shape == (4000, 1)
    # the selected points carry larger residuals on average than the pool they were drawn from
    xp, tp = pde._sample_uniform(16000)
    r_pool = pde.score_residual(model, xp, tp)[0].mean().item()
    r_sel = pde.score_residual(model, x, t)[0].mean().item()
    parity_log.log(f"[device sampler] RAR: mean |r| of the selected points {r_sel:.4e} vs pool {r_pool:.4e}")
    assert r_sel > r_pool

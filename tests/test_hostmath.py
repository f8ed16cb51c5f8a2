"""CPU check of the product's device math (csrc/jet_math.cuh compiled for the host).

The jet recurrences and, above all, their HAND-WRITTEN adjoints are compared in fp64 with
torch.autograd through the oracle's formulas (oracle/jets_oracle.py).  No GPU needed.
"""
import ctypes
import os
import subprocess

import numpy as np
import pytest
import torch

from oracle import jets_oracle as jo

HERE = os.path.dirname(os.path.abspath(__file__))
D = ctypes.c_double
PD = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def hm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("hostmath") / "hostmath.so")
    subprocess.check_call(["g++", "-x", "c++", "-O2", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "hostmath", "hostmath.cpp")])
    lib = ctypes.CDLL(out)
    lib.hm_pde.restype = D
    lib.hm_pde_x.restype = D
    lib.hm_rho.restype = D
    return lib


def _p(a):
    return a.ctypes.data_as(PD)


def _rand_jet(rng, K, scale=1.0):
    return rng.standard_normal(K + 1) * scale


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_tanh_jet_and_adjoint(hm, K):
    rng = np.random.default_rng(K)
    for _ in range(20):
        z = _rand_jet(rng, K, 0.8)
        yb = rng.standard_normal(K + 1)
        y = np.zeros(K + 1)
        zb = np.zeros(K + 1)
        for maxk in range(K, 5):
            hm.hm_tanh_k(maxk, K, _p(z), _p(yb), _p(y), _p(zb))
            zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
            y0, dirs = jo.tanh_jet((zt[0:1], [[zt[k:k + 1] for k in range(1, K + 1)]]))
            yt = torch.cat([y0] + dirs[0])
            (g,) = torch.autograd.grad((yt * torch.tensor(yb)).sum(), zt)
            np.testing.assert_allclose(y, yt.detach().numpy(), rtol=1e-12, atol=1e-13)
            np.testing.assert_allclose(zb, g.numpy(), rtol=1e-11, atol=1e-12)


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_tanh_preactivation_recovered_from_output_jets(hm, K):
    """The reverse kernels of fused Linear+tanh layers keep only the OUTPUT jets: tanh_dir_recover must give back the
    pre-activation coefficients z_1..z_K (and the w series) the forward recurrence consumed."""
    rng = np.random.default_rng(40 + K)
    for trial in range(20):
        z = _rand_jet(rng, K, 0.8 if trial < 15 else 3.0)
        y, zb = np.zeros(K + 1), np.zeros(K + 1)
        hm.hm_tanh(K, _p(z), _p(np.zeros(K + 1)), _p(y), _p(zb))
        zr, w = np.zeros(K + 1), np.zeros(K + 1)
        hm.hm_tanh_recover(K, _p(y), _p(zr), _p(w))
        np.testing.assert_allclose(zr[1:], z[1:], rtol=1e-9, atol=1e-11)
        # w is the Taylor series of 1 - tanh^2
        wt = [-(sum(y[i] * y[k - i] for i in range(k + 1))) + (1.0 if k == 0 else 0.0) for k in range(K + 1)]
        np.testing.assert_allclose(w, wt, rtol=1e-12, atol=1e-13)
    # a saturated unit (w0 == 0): every recovered coefficient is 0, nothing is inf / nan
    y = np.zeros(K + 1); y[0] = 1.0
    zr, w = np.ones(K + 1), np.ones(K + 1)
    hm.hm_tanh_recover(K, _p(y), _p(zr), _p(w))
    assert np.all(np.isfinite(zr)) and np.all(zr[1:] == 0.0)


@pytest.mark.parametrize("K", [1, 2, 3, 4])
@pytest.mark.parametrize("omega", [1.0, 30.0])
def test_sincos_jet_and_adjoint(hm, K, omega):
    rng = np.random.default_rng(10 + K)
    for _ in range(20):
        z = _rand_jet(rng, K, 0.5)
        sb, cb = rng.standard_normal(K + 1), rng.standard_normal(K + 1)
        s, c, zb = np.zeros(K + 1), np.zeros(K + 1), np.zeros(K + 1)
        hm.hm_sin(K, D(omega), _p(z), _p(sb), _p(cb), _p(s), _p(c), _p(zb))
        zt = torch.tensor(z, dtype=torch.float64, requires_grad=True)
        (s0, sd), (c0, cd) = jo.sincos_jet((zt[0:1], [[zt[k:k + 1] for k in range(1, K + 1)]]), omega)
        st, ct = torch.cat([s0] + sd[0]), torch.cat([c0] + cd[0])
        (g,) = torch.autograd.grad((st * torch.tensor(sb)).sum() + (ct * torch.tensor(cb)).sum(), zt)
        np.testing.assert_allclose(s, st.detach().numpy(), rtol=1e-11, atol=1e-12 * omega ** K)
        np.testing.assert_allclose(c, ct.detach().numpy(), rtol=1e-11, atol=1e-12 * omega ** K)
        np.testing.assert_allclose(zb, g.numpy(), rtol=1e-10, atol=1e-11 * omega ** (K + 1))


@pytest.mark.parametrize("K", [1, 2, 3, 4])
def test_rsqrt_series_and_adjoint(hm, K):
    rng = np.random.default_rng(20 + K)
    for _ in range(20):
        v = _rand_jet(rng, K, 0.5)
        v[0] = abs(v[0]) + 0.5
        sb = rng.standard_normal(K + 1)
        s, vb = np.zeros(K + 1), np.zeros(K + 1)
        hm.hm_rsqrt(K, _p(v), _p(sb), _p(s), _p(vb))
        # reference: Taylor coefficients of (v(e))^-1/2 by direct differentiation
        vt = torch.tensor(v, dtype=torch.float64, requires_grad=True)
        ser = [vt[0] ** -0.5]
        for k in range(1, K + 1):
            ser.append(sum((-0.5 * j - (k - j)) * vt[j] * ser[k - j] for j in range(1, k + 1)) / (k * vt[0]))
        st = torch.stack(ser)
        (g,) = torch.autograd.grad((st * torch.tensor(sb)).sum(), vt)
        np.testing.assert_allclose(s, st.detach().numpy(), rtol=1e-12)
        np.testing.assert_allclose(vb, g.numpy(), rtol=1e-10, atol=1e-12)
        # and the series really is v^-1/2: compare with a polynomial evaluation at small e
        e = 1e-3
        val = sum(v[k] * e ** k for k in range(K + 1)) ** -0.5
        approx = sum(s[k] * e ** k for k in range(K + 1))
        assert abs(val - approx) < 50 * e ** (K + 1) * max(1.0, np.abs(s).max())


PDE_CASES = [  # (kind, compat, pde name, dims, compat string, orders)
    (0, 0, "heat", 1, "reference", [1, 1]), (0, 1, "heat", 1, "math", [2, 1]),
    (1, 0, "burgers", 1, "reference", [2, 1]), (2, 0, "kdv", 1, "reference", [3, 1]),
    (3, 0, "allen_cahn", 1, "reference", [2, 1]), (4, 0, "cahn_hilliard", 1, "reference", [4, 1]),
    (5, 0, "cahn_hilliard", 2, "reference", [1]), (6, 0, "allen_cahn", 2, "reference", [1]),
    (7, 0, "cahn_hilliard", 2, "math", [4, 4, 4, 4, 1]),
    (10, 0, "wave", 1, "reference", [2, 2]), (11, 0, "convection", 1, "reference", [1, 1]),
]


@pytest.mark.parametrize("kind,compat,name,dim,cstr,orders", PDE_CASES)
def test_pde_epilogue_and_partials(hm, kind, compat, name, dim, cstr, orders):
    rng = np.random.default_rng(kind)
    params = {"alpha": 0.37, "nu": 0.37, "epsilon": 0.37, "c": 0.37, "velocity": 0.37}
    ncols = 1 + sum(orders)
    for trial in range(10):
        U = rng.standard_normal(ncols)
        if trial == 0:
            U[0] = 12.0  # outside the Cahn-Hilliard clamp
        dU = np.zeros(ncols)
        iords = (ctypes.c_int * len(orders))(*orders)
        r = hm.hm_pde(kind, compat, D(0.37), len(orders), iords, dim + 1, _p(U), _p(dU))
        Ut = torch.tensor(U, dtype=torch.float64, requires_grad=True)
        dirs, col = [], 1
        for o in orders:
            dirs.append([Ut[col + k:col + k + 1] for k in range(o)])
            col += o
        rt = jo.residual_from_jet(name, (Ut[0:1], dirs), params, dim, cstr)
        (g,) = torch.autograd.grad(rt.sum(), Ut, allow_unused=True)
        assert abs(r - rt.item()) <= 1e-6 * max(1.0, abs(rt.item()))   # PdeDesc carries fp32 parameters
        np.testing.assert_allclose(dU, g.numpy(), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("kind,name,orders", [(12, "black_scholes", [2, 1]), (13, "pendulum", [2])])
def test_pde_epilogues_with_coordinate_or_single_direction(hm, kind, name, orders):
    """Black-Scholes (coefficients depend on the point coordinate S) and the pendulum ODE (one jet direction, t, order 2)."""
    rng = np.random.default_rng(kind)
    params = {"sigma": 0.37, "r": 0.21, "g": 9.81, "L": 9.81 / 0.37}
    ncols = 1 + sum(orders)
    for trial in range(10):
        U = rng.standard_normal(ncols)
        xs = 0.5 + rng.random() * 2
        dU = np.zeros(ncols)
        iords = (ctypes.c_int * len(orders))(*orders)
        r = hm.hm_pde_x(kind, D(0.37), D(0.21), len(orders), iords, 2, _p(U), _p(dU), D(xs))
        Ut = torch.tensor(U, dtype=torch.float64, requires_grad=True)
        dirs, col = [], 1
        for o in orders:
            dirs.append([Ut[col + k:col + k + 1] for k in range(o)])
            col += o
        rt = jo.residual_from_jet(name, (Ut[0:1], dirs), params, 1, "reference", x=torch.tensor([xs], dtype=torch.float64))
        (g,) = torch.autograd.grad(rt.sum(), Ut, allow_unused=True)
        g = torch.zeros_like(Ut) if g is None else g
        assert abs(r - rt.item()) <= 1e-6 * max(1.0, abs(rt.item()))
        np.testing.assert_allclose(dU, g.numpy(), rtol=1e-6, atol=1e-7)


def test_loss_rho(hm):
    for kind, fn in [(0, lambda e: e ** 2), (1, lambda e: e.abs()),
                     (2, lambda e: torch.nn.functional.huber_loss(e, torch.zeros_like(e), delta=0.7))]:
        for ev in [-2.0, -0.3, 0.4, 1.5]:
            d = D(0.0)
            rho = hm.hm_rho(kind, D(0.7), D(ev), ctypes.byref(d))
            et = torch.tensor(ev, dtype=torch.float64, requires_grad=True)
            lt = fn(et)
            (g,) = torch.autograd.grad(lt, et)
            assert abs(rho - lt.item()) < 1e-12 and abs(d.value - g.item()) < 1e-12

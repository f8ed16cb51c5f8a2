"""TEST INFRASTRUCTURE ONLY -- second, independent oracle: forward Taylor-mode jets.

Plain torch (fp64 by default) restatement of the jet recurrences the CUDA
kernels implement.  It exists to
  * pin the *math* (Faa-di-Bruno recurrences, LayerNorm jets, PDE epilogues)
    against the reference's nested-autograd path (``ref_port`` / pinnrl itself),
  * give parameter gradients of that math through ordinary ``torch.autograd``
    (the ground truth for the hand-written reverse pass), and
  * provide the *corrected* oracles where the reference is inexact or
    degenerate: exact LayerNorm jets, true multi-dim operators.

A jet is ``(a0, [[a_{d,1..K_d}] for d in directions])`` with normalised Taylor
coefficients: ``d^k u / ds^k = k! * a_k`` along direction ``d``.

Parity status: pinned by ``tests/test_oracle.py`` against the committed golden
fixtures generated from the unmodified reference.
"""

from __future__ import annotations

import math
from typing import Dict, List, Sequence, Tuple

import torch
import torch.nn as nn

Jet = Tuple[torch.Tensor, List[List[torch.Tensor]]]


# ---------------------------------------------------------------- primitives
def seed_jet(xt: torch.Tensor, directions: Sequence[Tuple[Sequence[float], int]]) -> Jet:
    """Input jet: value = xt, first coefficient = direction vector, rest zero."""
    dirs = []
    for vec, order in directions:
        v = torch.as_tensor(vec, dtype=xt.dtype, device=xt.device).expand_as(xt)
        dirs.append([v if k == 1 else torch.zeros_like(xt) for k in range(1, order + 1)])
    return xt, dirs


def linear_jet(j: Jet, W: torch.Tensor, b) -> Jet:
    a0, dirs = j
    y0 = a0 @ W.t()
    if b is not None:
        y0 = y0 + b
    return y0, [[a @ W.t() for a in d] for d in dirs]


def tanh_jet(j: Jet) -> Jet:
    z0, dirs = j
    y0 = torch.tanh(z0)
    w0 = 1 - y0 * y0
    out = []
    for zd in dirs:
        z = [z0] + list(zd)
        y, w = [y0], [w0]
        for k in range(1, len(z)):
            yk = sum(jj * z[jj] * w[k - jj] for jj in range(1, k + 1)) / k
            y.append(yk)
            w.append(-sum(y[i] * y[k - i] for i in range(k + 1)))
        out.append(y[1:])
    return y0, out


def sincos_jet(j: Jet, omega: float = 1.0) -> Tuple[Jet, Jet]:
    z0, dirs = j
    s0, c0 = torch.sin(omega * z0), torch.cos(omega * z0)
    so, co = [], []
    for zd in dirs:
        z = [omega * z0] + [omega * a for a in zd]
        s, c = [s0], [c0]
        for k in range(1, len(z)):
            s.append(sum(jj * z[jj] * c[k - jj] for jj in range(1, k + 1)) / k)
            c.append(-sum(jj * z[jj] * s[k - jj] for jj in range(1, k + 1)) / k)
        so.append(s[1:])
        co.append(c[1:])
    return (s0, so), (c0, co)


def layernorm_jet(j: Jet, gamma, beta, eps: float) -> Jet:
    z0, dirs = j
    c0 = z0 - z0.mean(-1, keepdim=True)
    v0 = (c0 * c0).mean(-1, keepdim=True) + eps
    s0 = v0 ** -0.5
    y0 = gamma * c0 * s0 + beta
    out = []
    for zd in dirs:
        c = [c0] + [a - a.mean(-1, keepdim=True) for a in zd]
        K = len(c) - 1
        v = [v0] + [sum((c[i] * c[k - i]).mean(-1, keepdim=True) for i in range(k + 1))
                    for k in range(1, K + 1)]
        s = [s0]
        for k in range(1, K + 1):
            s.append(sum((-0.5 * jj - (k - jj)) * v[jj] * s[k - jj] for jj in range(1, k + 1)) / (k * v0))
        out.append([gamma * sum(c[i] * s[k - i] for i in range(k + 1)) for k in range(1, K + 1)])
    return y0, out


def add_jet(a: Jet, b: Jet) -> Jet:
    return a[0] + b[0], [[x + y for x, y in zip(da, db)] for da, db in zip(a[1], b[1])]


def cat_jet(a: Jet, b: Jet) -> Jet:
    return (torch.cat([a[0], b[0]], -1),
            [[torch.cat([x, y], -1) for x, y in zip(da, db)] for da, db in zip(a[1], b[1])])


# ---------------------------------------------------------------- networks
def _inner(model: nn.Module) -> nn.Module:
    return model.model if hasattr(model, "model") and isinstance(model.model, nn.Module) else model


def network_jet(model: nn.Module, j: Jet) -> Jet:
    """Push a jet through one of the four in-scope architectures, dispatching on structure
    (works for the reference's modules, for ``ref_port`` and for the product's mirrors)."""
    m = _inner(model)
    name = type(m).__name__

    def act(jj, fn):
        n = type(fn).__name__
        if n == "Tanh":
            return tanh_jet(jj)
        raise NotImplementedError(f"activation {n} is outside the oracle")

    def ln(jj, mod):
        return layernorm_jet(jj, mod.weight, mod.bias, mod.eps)

    if name == "FeedForwardNetwork":
        for mod in m.layers:
            n = type(mod).__name__
            if n == "Linear":
                j = linear_jet(j, mod.weight, mod.bias)
            elif n in ("LayerNorm", "PrimitiveLayerNorm"):
                j = ln(j, mod)
            elif n == "Dropout":
                assert mod.p == 0.0
            else:
                j = act(j, mod)
        return j
    if name == "ResNet":
        j = act(linear_jet(j, m.input_layer.weight, m.input_layer.bias), m.activation_fn)
        for blk in m.blocks:
            L = blk.layers
            h = linear_jet(j, L[0].weight, L[0].bias)
            h = act(ln(h, L[1]), L[2])
            h = linear_jet(h, L[4].weight, L[4].bias)
            h = ln(h, L[5])
            j = act(add_jet(j, h), blk.activation_fn)
        return linear_jet(j, m.output_layer.weight, m.output_layer.bias)
    if name == "SIREN":
        for lyr in list(m.layers)[:-1]:
            j, _ = sincos_jet(linear_jet(j, lyr.linear.weight, lyr.linear.bias), lyr.omega_0)
        last = m.layers[-1]
        return linear_jet(j, last.weight, last.bias)
    if name == "FourierNetwork":
        s, c = sincos_jet(linear_jet(j, m.fourier.B.t(), None), 1.0)
        j = cat_jet(s, c)
        for lyr in list(m.layers)[:-1]:
            j = act(linear_jet(j, lyr.weight, lyr.bias), m.activation_fn)
        last = m.layers[-1]
        return linear_jet(j, last.weight, last.bias)
    raise NotImplementedError(name)


# ---------------------------------------------------------------- PDE epilogues
def jet_spec(pde: str, dimension: int = 1, compat: str = "reference"):
    """Directions (vector in (x..., t) input space, order) each PDE needs."""
    d = dimension
    e = lambda i: [1.0 if k == i else 0.0 for k in range(d + 1)]
    t_dir = (e(d), 1)
    if pde == "pendulum":
        return [(e(d), 2)]
    if d == 1:
        order = {"heat": 1 if compat == "reference" else 2, "burgers": 2, "kdv": 3,
                 "allen_cahn": 2, "cahn_hilliard": 4, "wave": 2, "convection": 1, "black_scholes": 2}[pde]
        return [(e(0), order), (e(d), 2) if pde == "wave" else t_dir]
    if compat == "reference":
        # every multi-dim residual degenerates; only u and u_t enter.
        return [t_dir]
    if pde == "cahn_hilliard" and d == 2:
        return [([1, 0, 0], 4), ([0, 1, 0], 4), ([1, 1, 0], 4), ([1, -1, 0], 4), t_dir]
    order = {"heat": 2, "burgers": 2, "kdv": 3, "allen_cahn": 2}[pde]
    return [(e(i), order) for i in range(d)] + [t_dir]


def residual_from_jet(pde: str, j: Jet, params: Dict[str, float], dimension: int = 1,
                      compat: str = "reference", x=None) -> torch.Tensor:
    u, dirs = j
    fact = [1.0, 1.0, 2.0, 6.0, 24.0]
    D = lambda d, k: fact[k] * dirs[d][k - 1]
    if pde == "pendulum":
        return D(0, 2) + (params.get("g", 9.81) / params.get("L", 1.0)) * torch.sin(u)
    if pde == "black_scholes" and dimension == 1:
        sg, rf = params.get("sigma", 0.2), params.get("r", 0.05)
        return D(1, 1) + 0.5 * sg ** 2 * x ** 2 * D(0, 2) + rf * x * D(0, 1) - rf * u
    if dimension == 1:
        u_t = D(1, 1)
        if pde == "heat":
            a = params.get("alpha", 0.01)
            return u_t - a * (D(0, 1) if compat == "reference" else D(0, 2))
        if pde == "burgers":
            return u_t + u * D(0, 1) - params.get("nu", 0.01) * D(0, 2)
        if pde == "kdv":
            return u_t + 6 * u * D(0, 1) + D(0, 3)
        if pde == "allen_cahn":
            return u_t - params.get("epsilon", 0.1) ** 2 * D(0, 2) - u + u ** 3
        if pde == "cahn_hilliard":
            e2 = params.get("epsilon", 0.1) ** 2
            inside = (u.abs() <= 10.0).to(u.dtype)
            return u_t + e2 * D(0, 4) - inside * ((3 * u * u - 1) * D(0, 2) + 6 * u * D(0, 1) ** 2)
        if pde == "wave":
            return D(1, 2) - params.get("c", 1.0) ** 2 * D(0, 2)
        if pde == "convection":
            v = params.get("velocity", 1.0)
            return u_t + (v[0] if isinstance(v, (list, tuple)) else v) * D(0, 1)
        raise NotImplementedError(pde)
    if compat == "reference":
        u_t = D(0, 1)
        if pde == "allen_cahn":
            return u_t - u + u ** 3
        return u_t                                   # heat / burgers / kdv / cahn-hilliard (F2)
    if pde == "cahn_hilliard" and dimension == 2:
        e2 = params.get("epsilon", 0.1) ** 2
        u_t = D(4, 1)
        bih = (2.0 / 3.0) * (D(0, 4) + D(1, 4)) + (D(2, 4) + D(3, 4)) / 6.0
        lap = D(0, 2) + D(1, 2)
        inside = (u.abs() <= 10.0).to(u.dtype)
        lap_mu = -e2 * bih + inside * ((3 * u * u - 1) * lap + 6 * u * (D(0, 1) ** 2 + D(1, 1) ** 2))
        return u_t - lap_mu
    raise NotImplementedError((pde, dimension, compat))


def residual(model, pde: str, x, t, params=None, dimension=1, compat="reference"):
    xt = torch.cat([x, t], dim=1)
    j = network_jet(model, seed_jet(xt, jet_spec(pde, dimension, compat)))
    return residual_from_jet(pde, j, params or {}, dimension, compat, x=x)


def flat_grad(model: nn.Module, loss: torch.Tensor) -> torch.Tensor:
    ps = [p for p in model.parameters() if p.requires_grad]
    gs = torch.autograd.grad(loss, ps, allow_unused=True)
    return torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1) for p, g in zip(ps, gs)])

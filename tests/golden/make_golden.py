"""Generate the golden fixtures under tests/golden/ from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case it
  1. builds the reference model (pinnrl ``PINNModel`` via ``ModelConfig``, SURVEY
     Appendix C.5) and the reference PDE object, runs the reference hot path in
     fp32 and fp64 and records residuals, loss components and parameter gradients;
  2. loads the same weights into ``oracle/ref_port.py`` and asserts the port is
     BIT-IDENTICAL to the reference in fp32 (this is what pins the oracle);
  3. runs ``oracle/jets_oracle.py`` in fp64 and asserts it matches the reference
     (or, where the reference is inexact/degenerate -- SURVEY F2/F4 -- the corrected
     autograd oracle) to ~1e-12;
  4. writes ``<case>.npz`` (weights, points, expected outputs) and a summary line in
     ``golden_report.json``.

Fixture networks are kept small (<= 0.5 MB each); the full-size BASELINE configs are
re-checked here at generation time and their agreement is recorded in the report.
"""

from __future__ import annotations

import copy
import json
import math
import os
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
for _m in ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "plotly", "plotly.graph_objects",
           "plotly.subplots", "plotly.express"):
    if _m not in sys.modules:
        mm = MagicMock()
        mm.__path__ = []
        sys.modules[_m] = mm

from pinnrl.config import Config, ModelConfig  # noqa: E402
from pinnrl.neural_networks import PINNModel as RefPINNModel  # noqa: E402
from pinnrl.pdes.pde_base import PDEConfig  # noqa: E402
from pinnrl.pdes.heat_equation import HeatEquation  # noqa: E402
from pinnrl.pdes.burgers_equation import BurgersEquation  # noqa: E402
from pinnrl.pdes.kdv_equation import KdVEquation  # noqa: E402
from pinnrl.pdes.cahn_hilliard import CahnHilliardEquation  # noqa: E402
from pinnrl.pdes.allen_cahn import AllenCahnEquation  # noqa: E402
from pinnrl.pdes.wave_equation import WaveEquation  # noqa: E402
from pinnrl.pdes.convection_equation import ConvectionEquation  # noqa: E402
from pinnrl.pdes.black_scholes import BlackScholesEquation  # noqa: E402
from pinnrl.pdes.pendulum_equation import PendulumEquation  # noqa: E402

from oracle import ref_port, jets_oracle  # noqa: E402

torch.set_num_threads(8)
CPU = torch.device("cpu")


def ref_model(arch, input_dim, hidden, layers, **extra):
    c = Config.__new__(Config)
    c.device = CPU
    c.model = ModelConfig(input_dim, hidden, 1, layers, "tanh", architecture=arch)
    for k, v in extra.items():
        setattr(c.model, k, v)
    return RefPINNModel(config=c, device=CPU)


PDES = {
    "heat": dict(cls=HeatEquation, domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"alpha": 0.01},
                 bcs={"dirichlet": {"type": "dirichlet"}},
                 ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0},
                 exact={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0}),
    "burgers": dict(cls=BurgersEquation, domain=[[-1.0, 1.0]], time=[0.0, 1.0],
                    params={"nu": 0.01 / math.pi}, bcs={"dirichlet": {"value": 0.0}},
                    ic={"type": "sine", "amplitude": -1.0, "frequency": 1.0}, exact={}),
    "kdv": dict(cls=KdVEquation, domain=[[-15.0, 15.0]], time=[0.0, 5.0], params={"speed": 1.0},
                bcs={"dirichlet": {"value": 0.0}}, ic={"type": "soliton", "speed": 1.0},
                exact={}),
    "cahn_hilliard": dict(cls=CahnHilliardEquation, domain=[[0.0, 1.0]], time=[0.0, 1.0],
                          params={"epsilon": 0.1}, bcs={"dirichlet": {"value": 0.0}},
                          ic={"type": "tanh", "epsilon": 0.1}, exact={}),
    "allen_cahn": dict(cls=AllenCahnEquation, domain=[[-1.0, 1.0]], time=[0.0, 1.0],
                       params={"epsilon": 0.1}, bcs={"dirichlet": {"value": 0.0}},
                       ic={"type": "tanh", "epsilon": 0.1}, exact={}),
    # the next PDE epilogues (second-order time jets; first-order transport)
    "wave": dict(cls=WaveEquation, domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"c": 1.5},
                 bcs={"dirichlet": {"value": 0.0}}, ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0}, exact={}),
    "convection": dict(cls=ConvectionEquation, domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"velocity": 0.7},
                       bcs={"dirichlet": {"value": 0.0}}, ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0}, exact={}),
    "black_scholes": dict(cls=BlackScholesEquation, domain=[[0.5, 2.0]], time=[0.0, 1.0], params={"sigma": 0.3, "r": 0.04},
                          bcs={"dirichlet": {"value": 0.0}}, ic={"type": "call_option", "strike_price": 1.0}, exact={}),
    "pendulum": dict(cls=PendulumEquation, domain=[[0.0, 1.0]], time=[0.0, 2.0], params={"g": 9.81, "L": 2.0},
                     bcs={"dirichlet": {"value": 0.0}}, ic={"type": "small_angle", "initial_angle": 0.2}, exact={}),
}


def ref_pde(name, dimension=1):
    s = PDES[name]
    dom = s["domain"] * dimension
    cfg = PDEConfig(name=name, domain=copy.deepcopy(dom), time_domain=list(s["time"]),
                    parameters=dict(s["params"]), boundary_conditions=copy.deepcopy(s["bcs"]),
                    initial_condition=dict(s["ic"]), exact_solution=dict(s["exact"]),
                    dimension=dimension, device=CPU, training=None)
    return s["cls"](config=cfg)


def port_model(arch, input_dim, hidden, layers, state, dtype=torch.float32, corrected=False, **extra):
    m = ref_port.PINNModel(arch, input_dim, hidden, layers, 1, "tanh",
                           omega_0=extra.get("omega_0", 30.0),
                           mapping_size=extra.get("mapping_size", 32), scale=extra.get("scale", 10.0),
                           corrected_layernorm=corrected)
    m.load_state_dict(state)
    return m.to(dtype)


def points(name, n, dimension, seed):
    s = PDES[name]
    g = torch.Generator().manual_seed(seed)
    lo, hi = s["domain"][0]
    x = torch.rand(n, dimension, generator=g) * (hi - lo) + lo
    t = torch.rand(n, 1, generator=g) * (s["time"][1] - s["time"][0]) + s["time"][0]
    return x, t


def grads_of(model, loss):
    ps = list(model.parameters())
    gs = torch.autograd.grad(loss, ps, allow_unused=True)
    return torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1) for p, g in zip(ps, gs)])


def rel(a, b):
    a, b = a.double().reshape(-1), b.double().reshape(-1)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def corrected_multidim_residual(name, model, x, t, params):
    """differentiate w.r.t. the full input tensor."""
    X = torch.cat([x, t], dim=1).detach().requires_grad_(True)
    u = model(X)
    g = lambda f: torch.autograd.grad(f, X, torch.ones_like(f), create_graph=True)[0]
    d = x.shape[1]
    du = g(u)
    u_t = du[:, d:d + 1]
    lap = sum(g(du[:, k:k + 1])[:, k:k + 1] for k in range(d))
    if name == "cahn_hilliard":
        uc = torch.clamp(u, -10.0, 10.0)
        mu = -params["epsilon"] ** 2 * lap + uc ** 3 - uc
        dmu = g(mu)
        lap_mu = sum(g(dmu[:, k:k + 1])[:, k:k + 1] for k in range(d))
        return u_t - lap_mu
    raise NotImplementedError


def run_case(tag, pde_name, arch, hidden, layers, n, dimension=1, mode="loss", save=True, seed=0, **extra):
    torch.manual_seed(seed)
    in_dim = dimension + 1
    rm = ref_model(arch, in_dim, hidden, layers, **extra)
    state = copy.deepcopy(rm.state_dict())
    rp = ref_pde(pde_name, dimension)
    x, t = points(pde_name, n, dimension, seed + 1)
    params = dict(PDES[pde_name]["params"])
    s = PDES[pde_name]

    def loss_of(pde, model, xx, tt):
        if mode == "loss":
            L = pde.compute_loss(model, xx, tt)
            return L
        r = pde.compute_residual(model, xx.clone(), tt.clone())
        return {"residual": (r ** 2).mean(), "total": (r ** 2).mean()}

    # --- reference fp32
    r32 = rp.compute_residual(rm, x.clone(), t.clone()).detach()
    L32 = loss_of(rp, rm, x, t)
    g32 = grads_of(rm, L32["total"])
    # --- reference fp64
    rm64 = copy.deepcopy(rm).double()
    rp64 = ref_pde(pde_name, dimension)
    x64, t64 = x.double(), t.double()
    torch.set_default_dtype(torch.float64)   # reference builds BC/IC points with the default dtype
    try:
        r64 = rp64.compute_residual(rm64, x64.clone(), t64.clone()).detach()
        if mode == "loss" and pde_name != "heat":
            # base compute_loss hard-codes float32 boundary tensors (pde_base.py:1103-1107);
            # evaluate the loss through the port in fp64 instead (same formulas).
            pm64 = port_model(arch, in_dim, hidden, layers, state, torch.float64, **extra)
            res = ref_port.RESIDUALS[pde_name](pm64, x64.clone(), t64.clone(),
                                               **{k: v for k, v in params.items() if k != "speed"})
            fns = ref_port.boundary_condition_fns(pde_name, s["bcs"], s["ic"], s["domain"], params, s["exact"])
            L64 = _base_loss64(pm64, res, s, fns)
            g64 = grads_of(pm64, L64["total"])
        else:
            L64 = loss_of(rp64, rm64, x64, t64)
            g64 = grads_of(rm64, L64["total"])
    finally:
        torch.set_default_dtype(torch.float32)

    # --- port fp32 must be bit-identical to the reference
    pm = port_model(arch, in_dim, hidden, layers, state, **extra)
    kw = {k: v for k, v in params.items() if k != "speed"}
    pr32 = ref_port.RESIDUALS[pde_name](pm, x.clone(), t.clone(), dimension=dimension, **kw)
    assert torch.equal(pr32.detach(), r32), f"{tag}: port residual differs from reference"
    if mode == "loss":
        if pde_name == "heat":
            nb, ni = max(n // 10, 10), max(n // 5, 10)
            PL = ref_port.heat_compute_loss(pm, pr32, s["domain"], s["time"],
                                            ref_port.initial_condition_fn(pde_name, s["ic"], s["domain"], params), nb, ni)
        else:
            PL = ref_port.base_compute_loss(pm, pr32, s["domain"], s["time"],
                                            ref_port.boundary_condition_fns(pde_name, s["bcs"], s["ic"], s["domain"], params, s["exact"]))
    else:
        PL = {"residual": (pr32 ** 2).mean(), "total": (pr32 ** 2).mean()}
    pg32 = grads_of(pm, PL["total"])
    for k in ("residual", "boundary", "initial", "total"):
        if k in PL and k in L32:
            assert torch.equal(PL[k].detach(), L32[k].detach()), f"{tag}: port loss[{k}] differs"
    assert torch.equal(pg32, g32), f"{tag}: port gradient differs from reference"

    # --- jets oracle fp64 vs (corrected) reference fp64
    has_ln = arch == "resnet"
    report = {"case": tag, "n": n, "ref32_vs_ref64_residual": rel(r32, r64),
              "ref32_vs_ref64_grad": rel(g32, g64)}
    jm = port_model(arch, in_dim, hidden, layers, state, torch.float64, corrected=has_ln, **extra)
    jr = jets_oracle.residual(jm, pde_name, x64, t64, params, dimension, "reference")
    if has_ln:
        # corrected autograd oracle = port with primitive LayerNorm
        cr = ref_port.RESIDUALS[pde_name](jm, x64.clone(), t64.clone(), dimension=dimension, **kw).detach()
        report["jets_vs_corrected64_residual"] = rel(jr, cr)
        report["unmodified_ref64_vs_corrected64_residual"] = rel(r64, cr)
        gc = grads_of(jm, (ref_port.RESIDUALS[pde_name](jm, x64.clone(), t64.clone(), dimension=dimension, **kw) ** 2).mean())
        gj = grads_of(jm, (jets_oracle.residual(jm, pde_name, x64, t64, params, dimension) ** 2).mean())
        report["jets_vs_corrected64_grad_mse"] = rel(gj, gc)
        assert report["jets_vs_corrected64_residual"] < 1e-10, report
        assert report["jets_vs_corrected64_grad_mse"] < 1e-9, report
        extra_out = {"residual64_corrected": cr.numpy(), "grad64_mse_corrected": gc.numpy()}
    else:
        report["jets_vs_ref64_residual"] = rel(jr, r64)
        assert report["jets_vs_ref64_residual"] < 1e-10, report
        gj = grads_of(jm, (jets_oracle.residual(jm, pde_name, x64, t64, params, dimension) ** 2).mean())
        gr = grads_of(rm64, (rp64.compute_residual(rm64, x64.clone(), t64.clone()) ** 2).mean())
        report["jets_vs_ref64_grad_mse"] = rel(gj, gr)
        assert report["jets_vs_ref64_grad_mse"] < 1e-9, report
        extra_out = {"grad64_mse": gr.numpy()}
    if dimension == 2 and pde_name == "cahn_hilliard":
        cm = corrected_multidim_residual(pde_name, jm, x64, t64, params).detach()
        jmth = jets_oracle.residual(jm, pde_name, x64, t64, params, dimension, "math").detach()
        report["jets_math_vs_corrected_multidim64"] = rel(jmth, cm)
        assert report["jets_math_vs_corrected_multidim64"] < 1e-9, report
        extra_out["residual64_math"] = cm.numpy()
        gm = grads_of(jm, (corrected_multidim_residual(pde_name, jm, x64, t64, params) ** 2).mean())
        extra_out["grad64_mse_math"] = gm.numpy()

    if save:
        out = {f"w::{k}": v.numpy() for k, v in state.items()}
        out.update(x=x.detach().numpy(), t=t.detach().numpy(), residual32=r32.numpy(), residual64=r64.numpy(),
                   grad32=g32.numpy(), grad64=g64.numpy())
        for k in ("residual", "boundary", "initial", "total"):
            if k in L32:
                out[f"loss32_{k}"] = np.float64(L32[k].item())
                out[f"loss64_{k}"] = np.float64(L64[k].item())
        out.update(extra_out)
        meta = dict(tag=tag, pde=pde_name, arch=arch, hidden=hidden, layers=layers, n=n,
                    dimension=dimension, mode=mode, params=params, extra=extra)
        out["meta"] = np.array(json.dumps(meta))
        np.savez_compressed(os.path.join(HERE, f"{tag}.npz"), **out)
    print(json.dumps(report))
    return report


def _base_loss64(model, residual, s, fns):
    """ref_port.base_compute_loss with float64 boundary tensors."""
    dev = residual.device
    dom, td = s["domain"], s["time"]
    res_loss = (residual ** 2).mean()
    xb = torch.tensor([dom[0][0], dom[0][1]], dtype=torch.float64).reshape(-1, 1)
    tb = torch.linspace(td[0], td[1], 100, dtype=torch.float32).double().reshape(-1, 1)
    xb = xb.repeat_interleave(len(tb), dim=0)
    tb = tb.repeat(len(xb) // len(tb), 1)
    b = sum(((model(torch.cat([xb, tb], 1)) - fn(xb, tb)) ** 2).mean() for fn in fns.values())
    xi = torch.linspace(dom[0][0], dom[0][1], 100, dtype=torch.float32).double().reshape(-1, 1)
    ti = torch.zeros_like(xi)
    i = ((model(torch.cat([xi, ti], 1)) - fns["initial"](xi, ti)) ** 2).mean()
    return {"residual": res_loss, "boundary": b, "initial": i, "total": res_loss + 10 * b + 10 * i}


def run_dqn_case(tag="x_dqn", state_dim=2, hidden=128, grid=24, seed=0):
    """Q-network of the RL sampler (rl_agent.py:15-88) in eval mode over a grid x grid candidate mesh: the unmodified
    reference vs oracle/ref_port.dqn_forward_port (bit-identical in fp32) -> weights, states, Q-values (fp32 / fp64)."""
    from pinnrl.rl.rl_agent import DQNNetwork as RefDQN
    torch.manual_seed(seed)
    net = RefDQN(state_dim, 1, hidden).eval()
    with torch.no_grad():                                   # non-trivial LayerNorm affine and biases (init is 1 / 0)
        for name, p in net.named_parameters():
            if name.endswith("bias") or ".1.weight" in name:
                p.add_(0.1 * torch.randn_like(p))
    axes = [torch.linspace(-1.0, 1.0, grid), torch.linspace(0.0, 1.0, grid)] + [torch.linspace(0.0, 1.0, grid)] * (state_dim - 2)
    pts = torch.stack([g.flatten() for g in torch.meshgrid(*axes[:state_dim], indexing="ij")], dim=1)
    with torch.no_grad():
        q32 = net(pts)
        port32 = ref_port.dqn_forward_port(net.state_dict(), pts)
        net64 = copy.deepcopy(net).double()
        q64 = net64(pts.double())
        port64 = ref_port.dqn_forward_port(net64.state_dict(), pts.double())
    assert torch.equal(q32, port32), "oracle port of DQNNetwork.forward is not bit-identical to the reference (fp32)"
    assert torch.equal(q64, port64)
    out = {"states": pts.numpy(), "q32": q32.numpy(), "q64": q64.numpy()}
    for k, v in net.state_dict().items():
        out["w::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    return {"case": tag, "port_bit_identical_fp32": True, "ref32_vs_ref64": rel(q32.double(), q64), "states": int(pts.shape[0])}


def run_adaptive_case(tag="x_adaptive_weights", steps=12, seed=0):
    """components/adaptive_weights.py (the re-weighting of trainer.py:580-634): unmodified reference vs the oracle port on a
    random sequence of component losses / gradient norms, both strategies -> inputs and the weights after every update."""
    from pinnrl.components.adaptive_weights import AdaptiveLossWeights as RefALW
    g = torch.Generator().manual_seed(seed)
    out = {}
    for strategy in ("rbw", "lrw"):
        seq = torch.rand(steps, 3, generator=g) * torch.tensor([5.0, 0.5, 0.05])
        ref = RefALW(strategy=strategy, alpha=0.9, eps=1e-5, initial_weights=[0.5, 0.3, 0.2])
        port = ref_port.AdaptiveLossWeightsPort(strategy=strategy, alpha=0.9, eps=1e-5, initial_weights=[0.5, 0.3, 0.2])
        ws = []
        for v in seq:
            kw = {"losses": v} if strategy == "rbw" else {"gradients": v}
            a, b = ref.update(**kw), port.update(**kw)
            assert torch.equal(a, b), "oracle port of AdaptiveLossWeights differs from the reference"
            ws.append(a.clone())
        out[strategy + "_in"] = seq.numpy()
        out[strategy + "_w"] = torch.stack(ws).numpy()
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    return {"case": tag, "port_bit_identical_fp32": True, "steps": steps}


def run_snapshot_case(tag="x_live_snapshot", grid=12, seed=0):
    """trainer.py:171-279 ``_save_live_snapshot`` of the UNMODIFIED reference (Burgers, feed-forward 3x32, CPU): the npz it writes
    plus the weights, so that the mirror's grid construction / ordering / keys can be checked against the real file."""
    import tempfile
    from pinnrl.training.trainer import PDETrainer as RefTrainer
    torch.manual_seed(seed)
    model = ref_model("feedforward", 2, 32, 3)
    pde = ref_pde("burgers")
    tr = RefTrainer.__new__(RefTrainer)                      # only the attributes the method reads
    tr.model, tr.pde, tr.device = model, pde, CPU
    import logging
    tr.logger = logging.getLogger("golden")
    d = tempfile.mkdtemp()
    tr._save_live_snapshot(d, epoch=3, grid_size=grid)
    z = np.load(os.path.join(d, "live_snapshot.npz"))
    out = {"snap::" + k: z[k] for k in z.files}
    for k, v in model.state_dict().items():
        out["w::" + k] = v.numpy()
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    return {"case": tag, "keys": sorted(z.files), "grid": grid}


def run_smoothness_case(tag="x_heat_smoothness", n=300, seed=0):
    """HeatEquation._compute_smoothness_loss (heat_equation.py:625-650) and the full HeatEquation.compute_loss with the
    shipped smoothness weight 0.1 (config.yaml:338-343): unmodified reference (fp32 / fp64) vs oracle/ref_port (bit-identical
    in fp32) -> weights, points, smoothness value, its parameter gradient, and the loss dict with the term switched on."""
    torch.manual_seed(seed)
    model = ref_model("fourier", 2, 128, 3, mapping_size=32, scale=10.0)
    pde = ref_pde("heat")
    x, t = points("heat", n, 1, seed + 1)
    x[:3] = torch.tensor([[0.0], [1.0], [0.99995]])                       # rows the clamp acts on
    out = {"x": x.numpy(), "t": t.numpy()}
    for k, v in model.state_dict().items():
        out["w::" + k] = v.numpy()
    port = port_model("fourier", 2, 128, 3, model.state_dict(), mapping_size=32, scale=10.0)
    s_ref = pde._compute_smoothness_loss(model, x, t)
    s_port = ref_port.heat_smoothness_loss(port, x, t, PDES["heat"]["domain"])
    assert torch.equal(s_ref, s_port), "oracle port of the smoothness regulariser is not bit-identical to the reference (fp32)"
    g_ref, g_port = grads_of(model, s_ref), grads_of(port, s_port)
    assert torch.equal(g_ref, g_port)
    m64 = copy.deepcopy(model).double()
    s64 = pde._compute_smoothness_loss(m64, x.double(), t.double())
    g64 = grads_of(m64, s64)
    # (fp64 gradients are stored rounded to fp32 -- 6e-8, far below every gate -- and the reference's own fp32 results as
    # their distance to fp64: the fixture stays ~0.3 MB)
    out.update(smooth32=s_ref.detach().numpy(), smooth64=s64.detach().numpy(), gsmooth64=g64.float().numpy(),
               gsmooth32_err=np.array(rel(g_ref.double(), g64)))
    # the whole compute_loss with the term on (fixed weights incl. smoothness 0.1), as the trainer calls it
    pde.config.training = {"num_collocation_points": n, "num_boundary_points": 40, "num_initial_points": 40,
                           "loss_weights": {"residual": 1.0, "boundary": 10.0, "initial": 10.0, "smoothness": 0.1}}
    for dt, m, xx, tt in ((torch.float32, model, x, t), (torch.float64, m64, x.double(), t.double())):
        torch.set_default_dtype(dt)              # the reference builds its boundary / initial rows with the default dtype
        try:
            L = pde.compute_loss(m, xx, tt)
            sfx = "32" if dt == torch.float32 else "64"
            out["loss" + sfx] = np.array([float(L[k]) for k in ("residual", "boundary", "initial", "smoothness", "total")])
            gt = grads_of(m, L["total"])
            if dt == torch.float64:
                out["gtotal64"] = gt.float().numpy()
                out["gtotal32_err"] = np.array(rel(g32_total.double(), gt))
            else:
                g32_total = gt
        finally:
            torch.set_default_dtype(torch.float32)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **out)
    return {"case": tag, "port_bit_identical_fp32": True, "ref32_vs_ref64_value": abs(float(s_ref) - float(s64)) / abs(float(s64)),
            "ref32_vs_ref64_grad": rel(g_ref.double(), g64)}


def main_next():
    """`python tests/golden/make_golden.py next`: only the fixtures of the SURVEY 8(f).4 PDEs (existing files untouched)."""
    reports = [run_case("x_wave_ff_small", "wave", "feedforward", 32, 3, 96),
               run_case("x_wave_ff128", "wave", "feedforward", 128, 3, 64),
               run_case("x_convection_ff128", "convection", "feedforward", 128, 4, 96),
               run_case("x_convection_siren_small", "convection", "siren", 32, 3, 64, omega_0=30.0),
               run_case("x_black_scholes_ff128", "black_scholes", "feedforward", 128, 3, 96),
               run_case("x_pendulum_ff128", "pendulum", "feedforward", 128, 3, 64)]
    path = os.path.join(HERE, "golden_report.json")
    old = json.load(open(path)) if os.path.exists(path) else []
    old = [r for r in old if r["case"] not in {r2["case"] for r2 in reports}]
    with open(path, "w") as f:
        json.dump(old + reports, f, indent=1)


def main():
    reports = []
    # committed fixtures (small networks / few points)
    reports.append(run_case("c1_heat_fourier", "heat", "fourier", 128, 4, 256, mapping_size=32, scale=10.0))
    reports.append(run_case("c2_burgers_ff", "burgers", "feedforward", 128, 8, 192))
    reports.append(run_case("c3_kdv_resnet_small", "kdv", "resnet", 64, 2, 128, num_blocks=2))
    reports.append(run_case("c4_ch1d_siren_small", "cahn_hilliard", "siren", 64, 3, 128, mode="residual", omega_0=30.0))
    reports.append(run_case("c4_ch2d_siren_small", "cahn_hilliard", "siren", 64, 3, 128, dimension=2, mode="residual", omega_0=30.0))
    reports.append(run_case("c5_allen_cahn_ff_small", "allen_cahn", "feedforward", 64, 4, 128, mode="residual"))
    reports.append(run_case("x_burgers_siren_small", "burgers", "siren", 32, 3, 64, omega_0=30.0))
    reports.append(run_case("x_kdv_ff_small", "kdv", "feedforward", 32, 3, 64))
    # full-size configs: checked now, not stored
    reports.append(run_case("full_c3_kdv_resnet", "kdv", "resnet", 256, 6, 256, save=False, num_blocks=6))
    reports.append(run_case("full_c4_ch2d_siren", "cahn_hilliard", "siren", 256, 5, 256, dimension=2,
                            mode="residual", save=False, omega_0=30.0))
    reports.append(run_case("full_c4_ch1d_siren", "cahn_hilliard", "siren", 256, 5, 128,
                            mode="residual", save=False, omega_0=30.0))
    with open(os.path.join(HERE, "golden_report.json"), "w") as f:
        json.dump(reports, f, indent=1)


if __name__ == "__main__":
    if sys.argv[1:] in (["dqn"], ["adaptive"], ["snapshot"], ["smoothness"]):
        rep = {"dqn": run_dqn_case, "adaptive": run_adaptive_case, "snapshot": run_snapshot_case,
               "smoothness": run_smoothness_case}[sys.argv[1]]()
        path = os.path.join(HERE, "golden_report.json")
        old = [r for r in json.load(open(path)) if r["case"] != rep["case"]]
        json.dump(old + [rep], open(path, "w"), indent=1)
        print(rep)
    else:
        main_next() if sys.argv[1:] == ["next"] else main()

"""Shared helpers for the parity tests: load golden fixtures, rebuild models / PDEs from them."""
import json
import math
import os

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

PDES = {
    "heat": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"alpha": 0.01},
                 bcs={"dirichlet": {"type": "dirichlet"}}, ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0},
                 exact={"type": "sin_exp_decay", "amplitude": 1.0, "frequency": 2.0}),
    "burgers": dict(domain=[[-1.0, 1.0]], time=[0.0, 1.0], params={"nu": 0.01 / math.pi},
                    bcs={"dirichlet": {"value": 0.0}}, ic={"type": "sine", "amplitude": -1.0, "frequency": 1.0}, exact={}),
    "kdv": dict(domain=[[-15.0, 15.0]], time=[0.0, 5.0], params={"speed": 1.0},
                bcs={"dirichlet": {"value": 0.0}}, ic={"type": "soliton", "speed": 1.0}, exact={}),
    "cahn_hilliard": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"epsilon": 0.1},
                          bcs={"dirichlet": {"value": 0.0}}, ic={"type": "tanh", "epsilon": 0.1}, exact={}),
    "allen_cahn": dict(domain=[[-1.0, 1.0]], time=[0.0, 1.0], params={"epsilon": 0.1},
                       bcs={"dirichlet": {"value": 0.0}}, ic={"type": "tanh", "epsilon": 0.1}, exact={}),
    "wave": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"c": 1.5},
                 bcs={"dirichlet": {"value": 0.0}}, ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0}, exact={}),
    "convection": dict(domain=[[0.0, 1.0]], time=[0.0, 1.0], params={"velocity": 0.7},
                       bcs={"dirichlet": {"value": 0.0}}, ic={"type": "sine", "amplitude": 1.0, "frequency": 2.0}, exact={}),
    "black_scholes": dict(domain=[[0.5, 2.0]], time=[0.0, 1.0], params={"sigma": 0.3, "r": 0.04},
                          bcs={"dirichlet": {"value": 0.0}}, ic={"type": "call_option", "strike_price": 1.0}, exact={}),
    "pendulum": dict(domain=[[0.0, 1.0]], time=[0.0, 2.0], params={"g": 9.81, "L": 2.0},
                     bcs={"dirichlet": {"value": 0.0}}, ic={"type": "small_angle", "initial_angle": 0.2}, exact={}),
}


def fixtures():
    """PDE hot-path fixtures (x_dqn.npz, the RL sampler's Q-network, x_adaptive_weights.npz, x_live_snapshot.npz and
    x_heat_smoothness.npz have their own tests)."""
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f not in ("x_dqn.npz", "x_adaptive_weights.npz", "x_live_snapshot.npz", "x_heat_smoothness.npz"))


def load_fixture(tag):
    z = np.load(os.path.join(GOLDEN, tag + ".npz"))
    meta = json.loads(str(z["meta"]))
    state = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("w::")}
    return z, meta, state


def _t64(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().double().reshape(-1)
    return torch.as_tensor(np.asarray(a)).double().reshape(-1)


def rel(a, b):
    a, b = _t64(a), _t64(b)
    return float((a - b).norm() / max(float(b.norm()), 1e-300))


def port_model(meta, state, dtype=torch.float32, corrected=False):
    from oracle import ref_port
    ex = meta["extra"]
    m = ref_port.PINNModel(meta["arch"], meta["dimension"] + 1, meta["hidden"], meta["layers"], 1, "tanh",
                           omega_0=ex.get("omega_0", 30.0), mapping_size=ex.get("mapping_size", 32),
                           scale=ex.get("scale", 10.0), corrected_layernorm=corrected)
    m.load_state_dict(state)
    return m.to(dtype)


def product_model(meta, state, device):
    import pinns_rl_pde_b200 as pk
    ex = dict(meta["extra"])
    m = pk.make_model(meta["arch"], meta["dimension"] + 1, meta["hidden"], meta["layers"], device, **ex)
    m.load_state_dict(state)
    return m


def product_pde(name, device, dimension=1, training=None, compat="reference"):
    import pinns_rl_pde_b200 as pk
    s = PDES[name]
    cfg = pk.PDEConfig(name=name, domain=[list(d) for d in s["domain"] * dimension], time_domain=list(s["time"]),
                       parameters=dict(s["params"]), boundary_conditions={k: dict(v) for k, v in s["bcs"].items()},
                       initial_condition=dict(s["ic"]), exact_solution=dict(s["exact"]), dimension=dimension,
                       device=device, training=training)
    pde = pk.create_pde(name, cfg)
    pde.compat = compat
    return pde


def flat_grad(model):
    return torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in model.parameters()])

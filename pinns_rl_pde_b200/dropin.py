"""Swap the B200 hot path into an installed ``pinnrl`` (the reference) in place.

    import pinns_rl_pde_b200 as pk
    pk.patch_reference()          # pinnrl's PDE classes now evaluate residual / loss on libpinnk

After patching, the reference's own ``PDETrainer``, RAR sampler and live snapshot run unchanged on
top of the CUDA path (they only call ``compute_residual`` / ``compute_loss``).
Models must be one of the four hot-path architectures and live on a CUDA device.
"""
from __future__ import annotations

import importlib

from . import functional as F

_PATCHED = {}


def patch_reference(compat: str = "reference"):
    mods = {"heat_equation": "HeatEquation", "burgers_equation": "BurgersEquation", "kdv_equation": "KdVEquation",
            "allen_cahn": "AllenCahnEquation", "cahn_hilliard": "CahnHilliardEquation",
            "wave_equation": "WaveEquation", "convection_equation": "ConvectionEquation",
            "black_scholes": "BlackScholesEquation", "pendulum_equation": "PendulumEquation"}
    try:
        base = importlib.import_module("pinnrl.pdes.pde_base")
    except ImportError as e:
        raise ImportError("patch_reference() needs the reference package `pinnrl` importable") from e
    for mod, cls_name in mods.items():
        cls = getattr(importlib.import_module(f"pinnrl.pdes.{mod}"), cls_name)
        if cls in _PATCHED:
            continue
        _PATCHED[cls] = (cls.__dict__.get("compute_residual"), cls.__dict__.get("compute_loss"))
        cls.compat = compat
        cls.compute_residual = lambda self, model, x, t: F.compute_residual(self, model, x, t)
        cls.compute_loss = lambda self, model, x, t: F.compute_loss(self, model, x, t)
    # plugin PDEs (CONTRIBUTING.md:152-244) write compute_residual on top of PDEBase.compute_derivatives: one jet pass
    # (same keys, same F1 / F2 bookkeeping under compat="reference") instead of nested autograd.grad chains
    if base.PDEBase not in _PATCHED:
        _PATCHED[base.PDEBase] = ("compute_derivatives", base.PDEBase.__dict__.get("compute_derivatives"))
        if not hasattr(base.PDEBase, "compat"):
            base.PDEBase.compat = compat
        base.PDEBase.compute_derivatives = (lambda self, model, x, t, temporal_derivatives=None, spatial_derivatives=None:
                                            F.compute_derivatives(self, model, x, t, temporal_derivatives, spatial_derivatives))
    # RL sampler: RLAgent.select_action (rl/rl_agent.py:214-229) scores the candidate grid through pinnk_dqn_forward when
    # the agent lives on a GPU; a CPU agent keeps the reference's own torch forward (the agent is not on the hot path then).
    try:
        agent_cls = importlib.import_module("pinnrl.rl.rl_agent").RLAgent
    except Exception:          # the module imports matplotlib at the top; without it there is no agent to patch
        agent_cls = None
    if agent_cls is not None and agent_cls not in _PATCHED:
        original = agent_cls.select_action
        _PATCHED[agent_cls] = (original,)

        def select_action(self, state, _orig=original):
            import torch
            from . import rl
            if torch.device(self.device).type == "cuda":
                try:
                    rl._lower(getattr(self, "policy_net", None))
                except rl.UnsupportedQNetwork:
                    # not a DQNNetwork-shaped policy net (or a width pinnk_dqn_forward does not cover): the agent's own
                    # torch forward, exactly as before the patch
                    return _orig(self, state)
                return rl.select_action(self, state)
            return _orig(self, state)
        agent_cls.select_action = select_action
    return sorted(c.__name__ for c in _PATCHED)


def unpatch_reference():
    for cls, saved in list(_PATCHED.items()):
        if len(saved) == 2 and saved[0] == "compute_derivatives":
            if saved[1] is not None:
                cls.compute_derivatives = saved[1]
            _PATCHED.pop(cls)
            continue
        if len(saved) == 1:                      # RLAgent.select_action
            cls.select_action = saved[0]
            _PATCHED.pop(cls)
            continue
        res, loss = saved
        if res is not None:
            cls.compute_residual = res
        if loss is not None:
            cls.compute_loss = loss
        else:
            try:
                del cls.compute_loss
            except AttributeError:
                pass
        _PATCHED.pop(cls)

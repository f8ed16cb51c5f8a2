"""Q-network scoring of the RL sampler's candidate grid.

The reference's adaptive sampler asks its DQN for one score per grid point
(``PDEBase.generate_collocation_points("adaptive")``, pdes/pde_base.py:961-1018 ->
``RLAgent.select_action``, rl/rl_agent.py:214-229 -> ``DQNNetwork.forward``, rl_agent.py:15-88): a
``[Linear -> LayerNorm -> ReLU -> Dropout] x 2 -> Linear`` network over <= 100^(d+1) states, ~12 small launches
with ``[N, hidden]`` round trips.  Here the whole network is ONE libpinnk launch (``pinnk_dqn_forward``,
csrc/kernels_dqn.cuh).  Replay buffer, target network and the agent's optimiser are the agent's own business
(not on the path); this module only replaces the forward over the grid.

Dropout: the reference never puts ``policy_net`` into eval mode, so dropout (p = 0.1) is live while it scores the
grid.  To stay a drop-in, masks are drawn with the same torch RNG calls the reference's ``nn.Dropout`` modules make
(one ``[N, hidden]`` draw per hidden group, in order) and applied inside the kernel; in eval mode (or p = 0) no mask is
drawn and nothing is consumed from the generator.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L


class DQNNetwork(nn.Module):
    """Mirror of rl_agent.py:15-88 (same sub-module layout, state-dict keys and initialisation: Xavier-normal
    weights, zero biases), so checkpoints of the reference's policy / target networks load unchanged."""

    def __init__(self, state_dim: int, action_dim: int, hidden_dim: int, num_layers: int = 3, dropout: float = 0.1):
        super().__init__()
        groups: List[nn.Module] = []
        width_in = state_dim
        for _ in range(num_layers - 1):
            groups.append(nn.Sequential(nn.Linear(width_in, hidden_dim), nn.LayerNorm(hidden_dim), nn.ReLU(),
                                        nn.Dropout(dropout)))
            width_in = hidden_dim
        groups.append(nn.Linear(hidden_dim, action_dim))
        self.layers = nn.Sequential(*groups)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_normal_(m.weight, gain=1.0)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return dqn_forward(self, x)


class UnsupportedQNetwork(ValueError):
    pass


MAX_HIDDEN = 128       # widest Q-network the one-launch kernel is used for (benchmarks/sampling.py:124 builds hidden_dim = 64)
MAX_HIDDEN_WIDE = 1024  # multiples of 128 above that run their hidden Linear layers on the tcgen05 rows kernel
                        # (pinnk_dqn_forward_wide; config.yaml:363 / train.py:348-351 build hidden_dim = 512)


def _lower(net: nn.Module) -> Tuple[List[Tuple[nn.Linear, nn.LayerNorm, float]], nn.Linear]:
    """By structure, so the reference's own ``DQNNetwork`` instances are accepted."""
    seq = getattr(net, "layers", None)
    if not isinstance(seq, nn.Sequential) or len(seq) < 2:
        raise UnsupportedQNetwork("expected DQNNetwork.layers = Sequential(groups..., Linear)")
    groups = []
    for g in list(seq)[:-1]:
        mods = list(g) if isinstance(g, nn.Sequential) else []
        names = [type(m).__name__ for m in mods]
        if names != ["Linear", "LayerNorm", "ReLU", "Dropout"]:
            raise UnsupportedQNetwork(f"hidden group {names} is not Linear-LayerNorm-ReLU-Dropout")
        lin, ln, _, drop = mods
        if tuple(ln.normalized_shape) != (lin.out_features,):
            raise UnsupportedQNetwork("LayerNorm must normalise the feature axis")
        groups.append((lin, ln, float(drop.p)))
    out = seq[-1]
    if type(out).__name__ != "Linear":
        raise UnsupportedQNetwork("the Q-network must end in a Linear layer")
    if not 1 <= len(groups) <= 8:
        raise UnsupportedQNetwork("1..8 hidden groups")
    h = groups[0][0].out_features
    if h > MAX_HIDDEN and not (h % 128 == 0 and h <= MAX_HIDDEN_WIDE and groups[0][0].in_features <= 8):
        # the one-launch kernel is FFMA-bound (x1.8 over the torch modules at hidden 128 on a 100 x 100 grid, x0.6 at 256); wider
        # nets take the tcgen05 route when their width is a multiple of 128, anything else keeps the agent's own forward
        raise UnsupportedQNetwork(f"hidden width {h}: not covered by pinnk_dqn_forward (<= {MAX_HIDDEN}) nor by "
                                  f"pinnk_dqn_forward_wide (multiples of 128 up to {MAX_HIDDEN_WIDE})")
    if any(lin.out_features != h for lin, _, _ in groups):
        raise UnsupportedQNetwork("hidden groups of different widths")
    return groups, out


def _ptr(tn: Optional[torch.Tensor]):
    if tn is None:
        return None
    if not (tn.is_cuda and tn.dtype == torch.float32 and tn.is_contiguous()):
        raise L.PinnkError("Q-network tensors must be contiguous float32 CUDA tensors (there is no CPU path)")
    return tn.data_ptr()


def dqn_forward(net: nn.Module, states: torch.Tensor) -> torch.Tensor:
    """``net(states)`` -> [N, action_dim] in one libpinnk launch.  No autograd graph (the callers on the path run under
    ``no_grad``; the agent's own TD update keeps using torch)."""
    groups, out = _lower(net)
    if not states.is_cuda:
        raise L.PinnkError("dqn_forward needs CUDA states (there is no CPU or PyTorch fallback for this path)")
    x = states.detach().to(torch.float32).reshape(-1, states.shape[-1]).contiguous()
    n = x.shape[0]
    if x.shape[1] != groups[0][0].in_features:
        raise ValueError(f"states have {x.shape[1]} columns, the network takes {groups[0][0].in_features}")
    layers = (L.PinnkDqnLayer * len(groups))()
    keep = []
    for i, (lin, ln, p) in enumerate(groups):
        c = layers[i]
        c.weight, c.bias = _ptr(lin.weight.detach()), _ptr(None if lin.bias is None else lin.bias.detach())
        c.ln_weight = _ptr(None if ln.weight is None else ln.weight.detach())
        c.ln_bias = _ptr(None if ln.bias is None else ln.bias.detach())
        c.eps, c.in_dim, c.out_dim = float(ln.eps), lin.in_features, lin.out_features
        if net.training and p > 0.0 and n > 0:
            # the same generator call nn.Dropout makes on this group's [N, hidden] activation
            mask = torch.nn.functional.dropout(torch.ones(n, lin.out_features, device=x.device), p, True)
            keep.append(mask)
            c.dropout_mask = mask.data_ptr()
    q = torch.empty(n, out.out_features, dtype=torch.float32, device=x.device)
    if n == 0:
        return q.reshape(*states.shape[:-1], out.out_features)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    h = groups[0][0].out_features
    if h > MAX_HIDDEN:
        ws = torch.empty(2 * n * h, dtype=torch.float32, device=x.device)
        L.check(L.load().pinnk_dqn_forward_wide(layers, len(groups), _ptr(out.weight.detach()),
                                                _ptr(None if out.bias is None else out.bias.detach()), out.out_features,
                                                x.data_ptr(), n, q.data_ptr(), ws.data_ptr(), ws.numel(), stream),
                "pinnk_dqn_forward_wide")
    else:
        L.check(L.load().pinnk_dqn_forward(layers, len(groups), _ptr(out.weight.detach()),
                                           _ptr(None if out.bias is None else out.bias.detach()), out.out_features,
                                           x.data_ptr(), n, q.data_ptr(), stream), "pinnk_dqn_forward")
    return q.reshape(*states.shape[:-1], out.out_features)


def select_action(agent, state: torch.Tensor) -> torch.Tensor:
    """rl_agent.py:214-229 with the policy network evaluated by libpinnk: same epsilon draw (``torch.rand(1)`` on the
    CPU generator), same ``[1, N * action_dim]`` result, same uniform ``[1, 1]`` exploration action."""
    if torch.rand(1).to(agent.device).item() > agent.epsilon:
        with torch.no_grad():
            return dqn_forward(agent.policy_net, state.to(agent.device)).view(1, -1)
    return torch.rand(1, 1, device=agent.device)


def grid_scores(agent, points: torch.Tensor) -> torch.Tensor:
    """What the adaptive sampler feeds to its multinomial draw (pde_base.py:1005-1008): |scores| normalised to 1.
    Agents whose policy network is not a DQNNetwork-shaped module keep their own ``select_action``."""
    try:
        _lower(getattr(agent, "policy_net", None))
        scores = select_action(agent, points)
    except UnsupportedQNetwork:
        scores = agent.select_action(points)
    probs = torch.abs(scores)
    return probs / torch.sum(probs)

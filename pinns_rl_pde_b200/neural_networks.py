"""Host-side mirror of ``pinnrl.neural_networks`` for the four hot-path architectures.

Same constructor arguments, attribute names, parameter names/shapes (so ``state_dict`` files are
interchangeable) and default initialisation as the reference
(pinnrl/neural_networks/{__init__,base_network,feedforward,resnet,siren,fourier}.py).  The modules
own ordinary ``nn.Parameter``s; ``forward`` runs through libpinnk on the GPU and is differentiable
w.r.t. the parameters.  There is no CPU forward.
"""
from __future__ import annotations

import math
from typing import Any, Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import functional as F


class ModelConfig:
    """pinnrl/config/__init__.py:172-253 -- including its quirk that ``omega_0``,
    ``mapping_size``, ``scale``, ``num_heads``... are class attributes the constructor never sets."""
    omega_0: Optional[float] = None
    num_blocks: Optional[int] = None
    mapping_size: int = 32
    scale: float = 10.0

    def __init__(self, input_dim, hidden_dim, output_dim, num_layers, activation, fourier_features=0,
                 fourier_scale=1.0, dropout=0.0, layer_norm=False, architecture="feedforward"):
        self.input_dim, self.hidden_dim, self.output_dim = input_dim, hidden_dim, output_dim
        self.num_layers, self.activation = num_layers, activation
        self.fourier_features, self.fourier_scale = fourier_features, fourier_scale
        self.dropout, self.layer_norm, self.architecture = dropout, layer_norm, architecture
        self.hidden_dims = [hidden_dim] * num_layers
        if architecture in ("resnet", "fno"):
            self.num_blocks = num_layers

    def get(self, key, default=None):
        return getattr(self, key, default)

    def __getitem__(self, key):
        return getattr(self, key)


class Config:
    """Minimal stand-in for ``pinnrl.config.Config``: ``.device`` and ``.model``."""

    def __init__(self, model: ModelConfig, device=None, training=None):
        self.model, self.device, self.training = model, device, training


def _activation(name: str) -> nn.Module:
    if name == "tanh":
        return nn.Tanh()
    if name in ("relu", "leaky_relu", "sigmoid", "gelu"):
        raise ValueError(f"activation '{name}' exists in pinnrl but is not on the B200 hot path (tanh, SIREN sine)")
    raise ValueError(f"Unsupported activation: {name}")


class BaseNetwork(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.device = config.get("device", torch.device("cpu"))

    def _prepare_input(self, x):
        if not isinstance(x, torch.Tensor):
            x = torch.tensor(np.asarray(x), dtype=torch.float32, device=self.device)
        return x

    def forward(self, x):
        return F.model_forward(self, self._prepare_input(x))

    def count_parameters(self) -> int:
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def save_state(self, path: str) -> None:
        torch.save({"model_state_dict": self.state_dict(), "config": self.config}, path)

    def load_state(self, path: str) -> None:
        state = torch.load(path, map_location=self.device, weights_only=False)
        self.load_state_dict(state["model_state_dict"])
        self.config = state["config"]


class FeedForwardNetwork(BaseNetwork):
    def __init__(self, config):
        super().__init__(config)
        self.input_dim, self.hidden_dims, self.output_dim = config["input_dim"], config["hidden_dims"], config["output_dim"]
        self.dropout_rate = config.get("dropout", 0.1)
        self.use_layer_norm = config.get("layer_norm", True)
        act = config.get("activation", "relu")
        mods: List[nn.Module] = []
        prev = self.input_dim
        for h in self.hidden_dims:
            mods.append(nn.Linear(prev, h))
            if self.use_layer_norm:
                mods.append(nn.LayerNorm(h))
            mods.append(_activation(act))
            if self.dropout_rate > 0.0:
                mods.append(nn.Dropout(self.dropout_rate))
            prev = h
        mods.append(nn.Linear(prev, self.output_dim))
        self.layers = nn.Sequential(*mods)
        self.to(self.device)


class ResNetBlock(nn.Module):
    def __init__(self, in_dim, hidden_dim, activation="relu", dropout=0.1):
        super().__init__()
        self.activation_fn = _activation(activation)
        self.layers = nn.Sequential(nn.Linear(in_dim, hidden_dim), nn.LayerNorm(hidden_dim), self.activation_fn,
                                    nn.Dropout(dropout), nn.Linear(hidden_dim, in_dim), nn.LayerNorm(in_dim),
                                    nn.Dropout(dropout))


class ResNet(BaseNetwork):
    def __init__(self, config):
        super().__init__(config)
        self.input_dim = config["input_dim"]
        if "hidden_dim" in config:
            self.hidden_dim = config["hidden_dim"]
        elif isinstance(config.get("hidden_dims"), list) and config["hidden_dims"]:
            self.hidden_dim = config["hidden_dims"][0]
        else:
            self.hidden_dim = 124
        self.num_blocks = config["num_blocks"] if "num_blocks" in config else config.get("num_layers", 4)
        self.output_dim = config["output_dim"]
        act = config.get("activation", "relu")
        self.activation_fn = _activation(act)
        self.dropout = config.get("dropout", 0.1)
        self.input_layer = nn.Linear(self.input_dim, self.hidden_dim)
        self.blocks = nn.ModuleList([ResNetBlock(self.hidden_dim, self.hidden_dim, act, self.dropout)
                                     for _ in range(self.num_blocks)])
        self.output_layer = nn.Linear(self.hidden_dim, self.output_dim)


class SIRENLayer(nn.Module):
    def __init__(self, in_features, out_features, omega_0=30.0):
        super().__init__()
        self.omega_0 = omega_0
        self.linear = nn.Linear(in_features, out_features)
        with torch.no_grad():
            b = np.sqrt(6 / in_features) / self.omega_0     # TypeError when omega_0 is None, like the reference (F5)
            self.linear.weight.uniform_(-b, b)


class SIREN(BaseNetwork):
    def __init__(self, config):
        super().__init__(config)
        self.input_dim, self.hidden_dims, self.output_dim = config["input_dim"], config["hidden_dims"], config["output_dim"]
        self.omega_0 = config.get("omega_0", 30.0)
        self.layers = nn.ModuleList()
        prev = self.input_dim
        for h in self.hidden_dims:
            self.layers.append(SIRENLayer(prev, h, omega_0=self.omega_0))
            prev = h
        self.layers.append(nn.Linear(prev, self.output_dim))


class FourierFeatures(nn.Module):
    def __init__(self, input_dim, mapping_size, scale=10.0, device=None):
        super().__init__()
        self.input_dim, self.mapping_size, self.scale = input_dim, mapping_size, scale
        self.register_buffer("B", torch.randn(input_dim, mapping_size, device=device) * scale)
        self.output_dim = 2 * mapping_size


class FourierNetwork(BaseNetwork):
    def __init__(self, config):
        super().__init__(config)
        self.input_dim = config["input_dim"]
        self.mapping_size = config.get("mapping_size", 32)
        self.hidden_dim = config["hidden_dim"]
        self.num_layers = config.get("num_layers", 4)
        self.output_dim = config["output_dim"]
        self.activation_fn = _activation(config.get("activation", "relu"))
        self.scale = config.get("scale", 10.0)
        self.fourier = FourierFeatures(self.input_dim, self.mapping_size, self.scale, device=self.device)
        self.layers = nn.ModuleList()
        prev = 2 * self.mapping_size
        for _ in range(self.num_layers - 1):
            self.layers.append(nn.Linear(prev, self.hidden_dim))
            prev = self.hidden_dim
        self.layers.append(nn.Linear(prev, self.output_dim))
        self.to(self.device)


class PINNModel(BaseNetwork):
    """pinnrl/neural_networks/__init__.py:61-154, restricted to the hot-path architectures."""

    def __init__(self, config, device=None, **kwargs):
        dev = device if device is not None else config.device
        mc = config.model
        mc.device = dev
        super().__init__(mc)
        self.config = config
        self.device = dev
        self.architecture = self.architecture_name = mc.architecture
        if self.architecture == "fourier":
            self.model = FourierNetwork(mc)
        elif self.architecture == "resnet":
            rc = {"input_dim": mc.input_dim, "hidden_dim": mc.hidden_dim, "output_dim": mc.output_dim,
                  "activation": mc.activation, "dropout": mc.dropout, "device": dev}
            nb = getattr(mc, "num_blocks", None)
            rc["num_blocks"] = nb if nb is not None else mc.num_layers
            if getattr(mc, "hidden_dims", None) is not None:
                rc["hidden_dims"] = mc.hidden_dims
            self.model = ResNet(rc)
        elif self.architecture == "siren":
            self.model = SIREN(mc)
        elif self.architecture in ("attention", "autoencoder", "fno"):
            raise ValueError(f"architecture '{self.architecture}' is outside the B200 hot path "
                             "(feedforward, resnet, siren, fourier)")
        else:
            self.model = FeedForwardNetwork(mc)
        self.model = self.model.to(dev)
        self.to(dev)


def make_model(architecture: str, input_dim: int, hidden_dim: int, num_layers: int, device,
               activation: str = "tanh", **extra) -> PINNModel:
    """Build a model the way every reference caller does."""
    mc = ModelConfig(input_dim, hidden_dim, 1, num_layers, activation, architecture=architecture)
    for k, v in extra.items():
        setattr(mc, k, v)
    return PINNModel(Config(mc, device=device), device=device)

"""Turn a PINN ``nn.Module`` into the op program libpinnk executes.

Works by *structure*, not by type identity, so it accepts the reference's own modules
(``pinnrl.neural_networks.PINNModel`` and the four in-scope architectures it wraps,
pinnrl/neural_networks/{feedforward,resnet,siren,fourier}.py) as well as this package's
mirrors.  Parameters are read by pointer at call time (never copied), gradients are laid
out in ``model.parameters()`` order inside one flat buffer (the unit of the data-parallel
all-reduce).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib as L


@dataclass
class OpSpec:
    kind: int
    in_dim: int
    out_dim: int
    act: int = 0
    scale: float = 1.0
    eps: float = 0.0
    weight: Optional[torch.Tensor] = None
    bias: Optional[torch.Tensor] = None
    w_transposed: bool = False


@dataclass
class NetProgram:
    ops: List[OpSpec]
    in_dim: int
    tensors: List[torch.Tensor] = field(default_factory=list)       # params[] passed to the library
    grad_params: List[nn.Parameter] = field(default_factory=list)   # trainable, in model.parameters() order
    grad_offsets: List[int] = field(default_factory=list)
    grad_floats: int = 0
    c_ops: Optional[object] = None                                   # ctypes array
    # Width padding (hidden widths such as SIREN's 124 or the autoencoder's 248 in the reference's YAML, config.yaml:22,43):
    # the library sees zero-padded SHADOW copies of the parameters whose hidden widths are multiples of 128, so the layers run
    # on the tcgen05 tiles instead of the CUDA-core GEMM.  Padded units compute act(0) = 0 with zero jets and feed zero weights,
    # so values and the gradients of the real entries are unchanged.  pad_entries: (real tensor, shadow, real flat offset or
    # -1, padded flat offset or -1); pad_floats: length of the padded flat gradient the library accumulates into.
    pad_entries: List[Tuple[torch.Tensor, torch.Tensor, int, int]] = field(default_factory=list)
    pad_floats: int = 0

    def split_flat(self, flat: torch.Tensor) -> List[torch.Tensor]:
        return [flat[o:o + p.numel()].view_as(p) for p, o in zip(self.grad_params, self.grad_offsets)]

    @property
    def padded(self) -> bool:
        return bool(self.pad_entries)

    def refresh_shadows(self):
        """Copy the real parameters into the leading block of their zero-padded shadows (every call: the fused optimizer
        updates parameters through raw pointers, so version counters cannot be trusted)."""
        for real, shadow, _, _ in self.pad_entries:
            src = real.detach()
            if shadow.device != src.device or shadow.dtype != src.dtype:      # the model was moved after the program was compiled
                shadow.data = torch.zeros(shadow.shape, dtype=src.dtype, device=src.device)
            if src.dim() == 2:
                shadow[:src.shape[0], :src.shape[1]].copy_(src)
            else:
                shadow[:src.shape[0]].copy_(src)

    def unpad_add(self, flat_pad: torch.Tensor, flat_real: torch.Tensor):
        """flat_real += the real entries of the padded flat gradient."""
        for real, shadow, off, poff in self.pad_entries:
            if off < 0 or poff < 0:
                continue
            blk = flat_pad[poff:poff + shadow.numel()].view_as(shadow)
            dst = flat_real[off:off + real.numel()].view_as(real)
            if real.dim() == 2:
                dst.add_(blk[:real.shape[0], :real.shape[1]])
            else:
                dst.add_(blk[:real.shape[0]])


class UnsupportedNetwork(ValueError):
    pass


def _unwrap(model: nn.Module) -> nn.Module:
    inner = getattr(model, "model", None)
    return inner if isinstance(inner, nn.Module) else model


def _act_of(mod: nn.Module) -> Tuple[int, float]:
    name = type(mod).__name__
    if name == "Tanh":
        return L.ACT_TANH, 1.0
    raise UnsupportedNetwork(f"activation {name} is not on the B200 hot path (tanh and SIREN sine are)")


def _linear(mod, ops: List[OpSpec], transposed_weight=None):
    ops.append(OpSpec(L.OP_LINEAR, mod.in_features, mod.out_features, weight=mod.weight, bias=mod.bias))


def _layernorm(mod, ops, width):
    if tuple(mod.weight.shape) != (width,):
        raise UnsupportedNetwork("LayerNorm over more than the feature axis")
    ops.append(OpSpec(L.OP_LAYERNORM, width, width, eps=float(mod.eps), weight=mod.weight, bias=mod.bias))


def _lower(m: nn.Module) -> Tuple[List[OpSpec], int]:
    name = type(m).__name__
    ops: List[OpSpec] = []
    if name == "FeedForwardNetwork":
        width = None
        for mod in m.layers:
            n = type(mod).__name__
            if n == "Linear":
                _linear(mod, ops)
                width = mod.out_features
            elif n in ("LayerNorm", "PrimitiveLayerNorm"):
                _layernorm(mod, ops, width)
            elif n == "Dropout":
                if mod.p > 0.0:
                    raise UnsupportedNetwork("Dropout with p > 0 is not supported (all hot-path configs use p = 0)")
            else:
                a, s = _act_of(mod)
                ops.append(OpSpec(L.OP_ACT, width, width, act=a, scale=s))
        return ops, m.layers[0].in_features
    if name == "ResNet":
        a, s = _act_of(m.activation_fn)
        _linear(m.input_layer, ops)
        h = m.input_layer.out_features
        ops.append(OpSpec(L.OP_ACT, h, h, act=a, scale=s))
        for blk in m.blocks:
            seq = blk.layers
            ba, bs = _act_of(blk.activation_fn)
            for drop in (seq[3], seq[6]):
                if drop.p > 0.0:
                    raise UnsupportedNetwork("Dropout with p > 0 is not supported")
            ops.append(OpSpec(L.OP_SKIP_SAVE, h, h))
            _linear(seq[0], ops)
            hh = seq[0].out_features
            _layernorm(seq[1], ops, hh)
            ops.append(OpSpec(L.OP_ACT, hh, hh, act=ba, scale=bs))
            _linear(seq[4], ops)
            _layernorm(seq[5], ops, h)
            ops.append(OpSpec(L.OP_SKIP_ADD, h, h))
            ops.append(OpSpec(L.OP_ACT, h, h, act=ba, scale=bs))
        _linear(m.output_layer, ops)
        return ops, m.input_layer.in_features
    if name == "SIREN":
        layers = list(m.layers)
        for lyr in layers[:-1]:
            _linear(lyr.linear, ops)
            w = lyr.linear.out_features
            ops.append(OpSpec(L.OP_ACT, w, w, act=L.ACT_SIN, scale=float(lyr.omega_0)))
        _linear(layers[-1], ops)
        return ops, layers[0].linear.in_features
    if name == "FourierNetwork":
        B = m.fourier.B
        in_dim, msz = int(B.shape[0]), int(B.shape[1])
        ops.append(OpSpec(L.OP_LINEAR, in_dim, msz, weight=B, bias=None, w_transposed=True))
        ops.append(OpSpec(L.OP_SINCOS, msz, 2 * msz))
        a, s = _act_of(m.activation_fn)
        layers = list(m.layers)
        for lyr in layers[:-1]:
            _linear(lyr, ops)
            ops.append(OpSpec(L.OP_ACT, lyr.out_features, lyr.out_features, act=a, scale=s))
        _linear(layers[-1], ops)
        return ops, in_dim
    raise UnsupportedNetwork(
        f"{name} is not one of the hot-path architectures (feedforward, resnet, siren, fourier)")


def _pad_width(w: int) -> int:
    """Hidden width the tcgen05 tiles run at: the next multiple of 128 when that costs at most 32 extra units (124 -> 128,
    248 -> 256); anything else keeps its width (and the CUDA-core GEMM)."""
    import os
    if os.environ.get("PINNK_DISABLE_PAD", "0") == "1":
        return w
    wp = -(-w // 128) * 128
    return wp if (w % 128 != 0 and w >= 96 and wp - w <= 32) else w


def _pad_program(ops: List[OpSpec]) -> Tuple[List[OpSpec], Dict[int, torch.Tensor]]:
    """Zero-padded shadow parameters for plain Linear / activation chains (feed-forward, SIREN) whose hidden widths are not
    multiples of 128.  Returns the op list the library runs (padded widths, shadow tensors) and {id(real): shadow}."""
    if any(o.kind not in (L.OP_LINEAR, L.OP_ACT) for o in ops) or any(o.w_transposed for o in ops):
        return ops, {}
    lin = [o for o in ops if o.kind == L.OP_LINEAR]
    if len(lin) < 2 or all(_pad_width(o.out_dim) == o.out_dim for o in lin[:-1]):
        return ops, {}
    shadows: Dict[int, torch.Tensor] = {}
    out: List[OpSpec] = []
    width = None                              # padded width flowing between ops
    for i, o in enumerate(ops):
        if o.kind == L.OP_LINEAR:
            last = o is lin[-1]
            in_p = o.in_dim if width is None else width
            out_p = o.out_dim if last else _pad_width(o.out_dim)
            w = torch.zeros(out_p, in_p, dtype=o.weight.dtype, device=o.weight.device)
            shadows[id(o.weight)] = w
            b = None
            if o.bias is not None:
                b = torch.zeros(out_p, dtype=o.bias.dtype, device=o.bias.device)
                shadows[id(o.bias)] = b
            out.append(OpSpec(L.OP_LINEAR, in_p, out_p, weight=w, bias=b))
            width = out_p
        else:
            out.append(OpSpec(o.kind, width, width, act=o.act, scale=o.scale, eps=o.eps))
    return out, shadows


def compile_network(model: nn.Module) -> NetProgram:
    """Lower ``model`` to a NetProgram.  Raises UnsupportedNetwork for anything outside the path."""
    inner = _unwrap(model)
    ops, in_dim = _lower(inner)
    if ops[-1].kind != L.OP_LINEAR or ops[-1].out_dim != 1:
        raise UnsupportedNetwork("the network must end in Linear(width -> 1)")
    real_ops = ops
    ops, shadows = _pad_program(ops)
    prog = NetProgram(ops=ops, in_dim=in_dim)
    # gradient layout: model.parameters() order, trainable parameters only
    offsets: Dict[int, int] = {}
    off = 0
    for p in model.parameters():
        if p.requires_grad and id(p) not in offsets:
            offsets[id(p)] = off
            prog.grad_params.append(p)
            prog.grad_offsets.append(off)
            off += p.numel()
    prog.grad_floats = off
    index: Dict[int, int] = {}

    def idx(tn: Optional[torch.Tensor]) -> int:
        if tn is None:
            return -1
        if id(tn) not in index:
            index[id(tn)] = len(prog.tensors)
            prog.tensors.append(tn)
        return index[id(tn)]

    # padded programs: the library's flat gradient is laid out over the SHADOW tensors (same parameter order)
    pad_offsets: Dict[int, int] = {}
    if shadows:
        poff = 0
        for p in model.parameters():
            if id(p) in shadows and id(p) in offsets and id(shadows[id(p)]) not in pad_offsets:
                pad_offsets[id(shadows[id(p)])] = poff
                poff += shadows[id(p)].numel()
        prog.pad_floats = poff
        seen = set()
        for o in real_ops:
            for tn in (o.weight, o.bias):
                if tn is not None and id(tn) in shadows and id(tn) not in seen:
                    seen.add(id(tn))
                    sh = shadows[id(tn)]
                    prog.pad_entries.append((tn, sh, offsets.get(id(tn), -1), pad_offsets.get(id(sh), -1)))
    goff = pad_offsets if shadows else offsets
    arr = (L.PinnkOp * len(ops))()
    for i, o in enumerate(ops):
        c = arr[i]
        c.kind, c.in_dim, c.out_dim, c.act = o.kind, o.in_dim, o.out_dim, o.act
        c.scale, c.eps = o.scale, o.eps
        c.w_index, c.b_index = idx(o.weight), idx(o.bias)
        c.w_transposed = 1 if o.w_transposed else 0
        c.gw_offset = goff.get(id(o.weight), -1) if o.weight is not None else -1
        c.gb_offset = goff.get(id(o.bias), -1) if o.bias is not None else -1
    prog.c_ops = arr
    used = {id(o.weight) for o in real_ops if o.weight is not None} | {id(o.bias) for o in real_ops if o.bias is not None}
    for p in prog.grad_params:
        if id(p) not in used:
            raise UnsupportedNetwork("model has a trainable parameter the op program does not use")
    return prog

"""JetEngine: one (network program, jet spec) plan of libpinnk plus its workspace.

PyTorch is plumbing here: it owns device memory (parameters, points, workspace, outputs) and the
CUDA stream; all arithmetic of the path happens inside libpinnk.so.
"""
from __future__ import annotations

import ctypes as C
import os
import weakref
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn as nn

from . import _lib as L
from .program import NetProgram, compile_network

Direction = Tuple[Tuple[float, ...], int]

# upper bound for the per-engine workspace; the chunk size is derived from it.  A chunk is one launch per kernel, so
# bigger chunks mean fewer launches / prologues / tails: 1 M-point chunks are 3 % faster than 256 K-point chunks on C2
# (gpurun_out/chunk2.log) and the 40 GB workspace is small change on a 180 GB B200.
MAX_WORKSPACE_BYTES = int(os.environ.get("PINNK_MAX_WORKSPACE_GB", 48)) << 30
MAX_CHUNK_POINTS = int(os.environ.get("PINNK_MAX_CHUNK_POINTS", 1 << 20))


@dataclass
class Segment:
    """One error functional over a row range of a call (PinnkSegment)."""
    kind: int
    row_start: int
    row_count: int
    component: int = 0
    weight: float = 1.0
    p0: float = 0.0
    p1: float = 0.0
    compat_math: int = 0
    pair_offset: int = 0
    target: Optional[torch.Tensor] = None
    error_out: Optional[torch.Tensor] = None
    error_grad: Optional[torch.Tensor] = None
    loss_kind: int = L.LOSS_MSE
    huber_delta: float = 1.0


def _require_cuda_f32(t: torch.Tensor, what: str):
    if not t.is_cuda:
        raise L.PinnkError(f"{what} must live on a CUDA device (got {t.device}); this path has no CPU fallback")
    if t.dtype != torch.float32:
        raise L.PinnkError(f"{what} must be float32 (got {t.dtype})")


class JetEngine:
    def __init__(self, program: NetProgram, directions: Sequence[Direction], chunk_points: int,
                 device: torch.device):
        self.lib = L.load()
        self.program = program
        self.directions = tuple((tuple(float(v) for v in vec), int(order)) for vec, order in directions)
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise L.PinnkError("JetEngine needs a CUDA device; this path has no CPU fallback")
        js = L.PinnkJetSpec()
        js.ndirs = len(self.directions)
        for d, (vec, order) in enumerate(self.directions):
            js.order[d] = order
            for i, v in enumerate(vec):
                js.vec[d][i] = v
        self.ncols = 1 + sum(o for _, o in self.directions)
        self.chunk = int(chunk_points)
        handle = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        L.check(self.lib.pinnk_plan_create(program.c_ops, len(program.ops), program.in_dim, C.byref(js),
                                           self.chunk, dev_index, C.byref(handle)), "pinnk_plan_create")
        self.handle = handle
        self._finalizer = weakref.finalize(self, self.lib.pinnk_plan_destroy, handle)
        self.ws_bytes = int(self.lib.pinnk_plan_workspace_bytes(handle))
        assert int(self.lib.pinnk_plan_ncols(handle)) == self.ncols
        self.workspace = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.device)
        self._ptrs = (C.c_void_p * max(1, len(program.tensors)))()
        # token of the forward-only call whose stash the workspace still holds (loss_step(keep_stash=True)); any other
        # call on this engine clears it
        self._stash_token = None
        self._token_count = 0

    # ------------------------------------------------------------------ helpers
    def _params(self):
        if self.program.padded:
            self.program.refresh_shadows()
        for i, tn in enumerate(self.program.tensors):
            _require_cuda_f32(tn, "model parameter")
            if not tn.is_contiguous():
                raise L.PinnkError("model parameters must be contiguous")
            self._ptrs[i] = tn.data_ptr()
        return C.cast(self._ptrs, C.c_void_p)

    def _xt(self, x: torch.Tensor, t: Optional[torch.Tensor]):
        _require_cuda_f32(x, "x")
        x = x.detach().contiguous()
        n = x.shape[0]
        if t is None:
            if x.dim() != 2 or x.shape[1] != self.program.in_dim:
                raise L.PinnkError(f"input must be [n, {self.program.in_dim}]")
            return x, None, n
        _require_cuda_f32(t, "t")
        t = t.detach().contiguous()
        if x.dim() != 2 or x.shape[1] != self.program.in_dim - 1 or t.numel() != n:
            raise L.PinnkError(f"x must be [n, {self.program.in_dim - 1}] and t [n, 1]")
        return x, t, n

    @staticmethod
    def _stream() -> C.c_void_p:
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    @staticmethod
    def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
        return None if t is None else t.data_ptr()

    # ------------------------------------------------------------------ entry points
    def jets_forward(self, x, t=None) -> torch.Tensor:
        self._stash_token = None
        x, t, n = self._xt(x, t)
        out = torch.empty(n, self.ncols, dtype=torch.float32, device=self.device)
        if n:
            L.check(self.lib.pinnk_jets_forward(self.handle, self._params(), x.data_ptr(), self._ptr(t), n,
                                                out.data_ptr(), self.workspace.data_ptr(), self.ws_bytes,
                                                self._stream()), "pinnk_jets_forward")
        return out

    def jets_vjp(self, x, t, adj: torch.Tensor, flat_grad: Optional[torch.Tensor] = None) -> torch.Tensor:
        self._stash_token = None
        x, t, n = self._xt(x, t)
        _require_cuda_f32(adj, "adj_jets")
        adj = adj.contiguous()
        if tuple(adj.shape) != (n, self.ncols):
            raise L.PinnkError(f"adj_jets must be [{n}, {self.ncols}]")
        if flat_grad is None:
            flat_grad = torch.zeros(self.program.grad_floats, dtype=torch.float32, device=self.device)
        if n:
            target = self._pad_target(flat_grad)
            L.check(self.lib.pinnk_jets_vjp(self.handle, self._params(), x.data_ptr(), self._ptr(t), n,
                                            adj.data_ptr(), target.data_ptr(), self.workspace.data_ptr(),
                                            self.ws_bytes, self._stream()), "pinnk_jets_vjp")
            self._pad_finish(target, flat_grad)
        return flat_grad

    # width-padded programs (program.pad_entries): the library accumulates into a zeroed flat gradient laid out over the
    # shadow tensors; the real entries are then added into the caller's buffer (model.parameters() layout)
    def _pad_target(self, flat_grad: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        if flat_grad is None or not self.program.padded:
            return flat_grad
        if getattr(self, "_flat_pad", None) is None:
            self._flat_pad = torch.zeros(self.program.pad_floats, dtype=torch.float32, device=self.device)
        else:
            self._flat_pad.zero_()
        return self._flat_pad

    def _pad_finish(self, target: Optional[torch.Tensor], flat_grad: Optional[torch.Tensor]):
        if flat_grad is not None and target is not flat_grad:
            self.program.unpad_add(target, flat_grad)

    def stash_signature(self):
        """What has to be unchanged between a keep_stash forward and the reverse pass that reuses its stash."""
        # (width-padded programs: the REAL parameters -- the shadows are rewritten from them at every call)
        tens = [real for real, _, _, _ in self.program.pad_entries] if self.program.padded else self.program.tensors
        return tuple((tn.data_ptr(), tn._version) for tn in tens)

    def loss_step(self, x, t, segments: Sequence[Segment], n_components: int, want_grad: bool,
                  grad_scale: Optional[Sequence[float]] = None, flat_grad: Optional[torch.Tensor] = None,
                  loss_sums: Optional[torch.Tensor] = None, keep_stash: bool = False, reuse_token=None):
        """Returns (loss_sums fp64 [n_components], flat_grad or None).

        ``keep_stash`` (forward-only call, n <= chunk): the workspace keeps what a reverse pass needs and
        ``self._stash_token`` names this call; ``reuse_token``: run the reverse pass from that stash when the token is
        still current (nothing else used the engine), else recompute the forward as usual."""
        flags = 0
        if reuse_token is not None and want_grad and reuse_token == self._stash_token:
            flags = L.STEP_REUSE_STASH
        self._stash_token = None
        x, t, n = self._xt(x, t)
        if keep_stash and not want_grad and 0 < n <= self.chunk:
            flags = L.STEP_KEEP_STASH
        if loss_sums is None:
            loss_sums = torch.zeros(n_components, dtype=torch.float64, device=self.device)
        if want_grad and flat_grad is None:
            flat_grad = torch.zeros(self.program.grad_floats, dtype=torch.float32, device=self.device)
        segs = (L.PinnkSegment * len(segments))()
        keep = []
        for i, s in enumerate(segments):
            c = segs[i]
            c.pde.kind, c.pde.compat_math, c.pde.p0, c.pde.p1 = s.kind, s.compat_math, s.p0, s.p1
            c.component, c.loss_kind, c.huber_delta, c.weight = s.component, s.loss_kind, s.huber_delta, s.weight
            c.row_start, c.row_count, c.pair_offset = s.row_start, s.row_count, s.pair_offset
            for name in ("target", "error_out", "error_grad"):
                tn = getattr(s, name)
                if tn is not None:
                    _require_cuda_f32(tn, name)
                    if not tn.is_contiguous() or tn.numel() != s.row_count:
                        raise L.PinnkError(f"segment {name} must be contiguous with row_count elements")
                    keep.append(tn)
                    setattr(c, name, tn.data_ptr())
        gs = None
        if grad_scale is not None:
            gs = (C.c_float * n_components)(*[float(g) for g in grad_scale])
        if n:
            target = self._pad_target(flat_grad if want_grad else None)
            L.check(self.lib.pinnk_loss_step_flags(self.handle, self._params(), x.data_ptr(), self._ptr(t), n, segs,
                                                   len(segments), C.cast(gs, C.c_void_p) if gs is not None else None,
                                                   loss_sums.data_ptr(), self._ptr(target) if want_grad else None,
                                                   self.workspace.data_ptr(), self.ws_bytes, self._stream(), flags),
                    "pinnk_loss_step")
            self._pad_finish(target, flat_grad if want_grad else None)
            if flags == L.STEP_KEEP_STASH:
                self._token_count += 1
                self._stash_token = (self._token_count, x.data_ptr(), self._ptr(t), n, self.stash_signature())
        return loss_sums, (flat_grad if want_grad else None)

    def score(self, x, t, kind: int, p0: float = 0.0, compat_math: int = 0, want_abs: bool = True,
              stats: Optional[torch.Tensor] = None, p1: float = 0.0):
        """Forward-only |r| and stats = [sum|r|, sum r^2, max|r|, count] (fp64, device)."""
        self._stash_token = None
        x, t, n = self._xt(x, t)
        abs_r = torch.empty(n, dtype=torch.float32, device=self.device) if want_abs else None
        if stats is None:
            stats = torch.zeros(4, dtype=torch.float64, device=self.device)
        pde = L.PinnkPde(kind, compat_math, p0, p1)
        if n:
            L.check(self.lib.pinnk_score(self.handle, self._params(), x.data_ptr(), self._ptr(t), n, C.byref(pde),
                                         self._ptr(abs_r), stats.data_ptr(), self.workspace.data_ptr(),
                                         self.ws_bytes, self._stream()), "pinnk_score")
        return abs_r, stats


# ---------------------------------------------------------------------- per-model engine cache
def _bytes_per_point(program: NetProgram, ncols: int) -> int:
    stash, width, maxw = 0, program.in_dim, 1
    for o in program.ops[:-1]:
        if o.kind in (L.OP_LINEAR, L.OP_ACT, L.OP_LAYERNORM, L.OP_SINCOS):
            width = o.out_dim
            stash += ncols * width
            maxw = max(maxw, width)
    return 4 * (stash + 2 * ncols + 3 * ncols * maxw)


def _signature(model: nn.Module):
    return tuple((id(p), p.requires_grad) for p in model.parameters())


_CACHE: "weakref.WeakKeyDictionary[nn.Module, Dict]" = weakref.WeakKeyDictionary()


def get_program(model: nn.Module) -> NetProgram:
    cache = _CACHE.setdefault(model, {})
    sig = _signature(model)
    if cache.get("sig") != sig:
        cache.clear()
        cache["sig"] = sig
        cache["program"] = compile_network(model)
    return cache["program"]


def get_engine(model: nn.Module, directions: Sequence[Direction], n_points: int,
               max_chunk: Optional[int] = None, whole: bool = False, program: Optional[NetProgram] = None) -> JetEngine:
    """Engine for (model, jet spec) able to process ``n_points`` rows per call; cached per model.
    ``whole``: the call must fit ONE chunk (paired periodic-BC rows reference each other), so only the
    workspace budget caps the chunk size."""
    if program is None:                # callers that already validated the model this call pass its program (the
        program = get_program(model)   # parameter-signature walk is the dominant host cost of a small-batch step)
    cache = _CACHE[model]
    dirs = tuple((tuple(float(v) for v in vec), int(order)) for vec, order in directions)
    ncols = 1 + sum(o for _, o in dirs)
    cap = max(1024, MAX_WORKSPACE_BYTES // _bytes_per_point(program, ncols))
    if not whole:
        cap = min(MAX_CHUNK_POINTS if max_chunk is None else max_chunk, cap)
    want = min(cap, max(256, -(-n_points // 256) * 256))
    eng: Optional[JetEngine] = cache.get(dirs)
    if eng is None or eng.chunk < want:
        device = next(model.parameters()).device
        eng = JetEngine(program, dirs, want, device)
        cache[dirs] = eng
    return eng

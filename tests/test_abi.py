"""CPU: the C-ABI library loads, exports every symbol include/pinnk.h declares, and validates programs
(no compute calls: there is no GPU here)."""
import ctypes as C
import os
import re

import pytest
import torch

import pinns_rl_pde_b200 as pk
from pinns_rl_pde_b200 import _lib as L
from pinns_rl_pde_b200.program import UnsupportedNetwork, compile_network

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from pinns_rl_pde_b200 import build
    build.build()
    return L.load()


def test_header_symbols_are_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "pinnk.h")).read()
    declared = set(re.findall(r"\b(pinnk_[a-z_]+)\s*\(", hdr))
    assert declared == set(L.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.pinnk_abi_version() == L.ABI_VERSION


def _plan(lib, model, dirs, chunk=1024):
    prog = compile_network(model)
    js = L.PinnkJetSpec()
    js.ndirs = len(dirs)
    for d, (vec, order) in enumerate(dirs):
        js.order[d] = order
        for i, v in enumerate(vec):
            js.vec[d][i] = v
    h = C.c_void_p()
    rc = lib.pinnk_plan_create(prog.c_ops, len(prog.ops), prog.in_dim, C.byref(js), chunk, 0, C.byref(h))
    return rc, h, prog


@pytest.mark.parametrize("arch", ["feedforward", "resnet", "siren", "fourier"])
def test_plan_create_for_every_architecture(lib, arch):
    m = pk.make_model(arch, 2, 64, 3, torch.device("cpu"), omega_0=30.0)
    rc, h, prog = _plan(lib, m, [((1.0, 0.0), 3), ((0.0, 1.0), 1)])
    assert rc == 0, lib.pinnk_last_error()
    assert lib.pinnk_plan_ncols(h) == 5
    assert lib.pinnk_plan_grad_floats(h) == sum(p.numel() for p in m.parameters() if p.requires_grad) == prog.grad_floats
    assert lib.pinnk_plan_workspace_bytes(h) > 0
    lib.pinnk_plan_destroy(h)


def test_plan_rejects_bad_programs(lib):
    m = pk.make_model("feedforward", 2, 64, 3, torch.device("cpu"))
    rc, _, _ = _plan(lib, m, [((1.0, 0.0), 5)])
    assert rc == L.ABI_VERSION * 0 - 1 and b"order" in lib.pinnk_last_error()
    m2 = pk.make_model("feedforward", 2, 30, 2, torch.device("cpu"))     # width not a multiple of 4
    rc, _, _ = _plan(lib, m2, [((1.0, 0.0), 1)])
    assert rc == -1 and b"multiples of 4" in lib.pinnk_last_error()


def test_unsupported_networks_fail_loudly():
    with pytest.raises(ValueError):
        pk.make_model("feedforward", 2, 32, 2, torch.device("cpu"), activation="relu")
    with pytest.raises(ValueError):
        pk.make_model("attention", 2, 32, 2, torch.device("cpu"))
    with pytest.raises(UnsupportedNetwork):
        compile_network(torch.nn.Sequential(torch.nn.Linear(2, 1)))


def test_no_cpu_fallback():
    """A CPU model must raise, never silently compute somewhere else."""
    m = pk.make_model("feedforward", 2, 32, 2, torch.device("cpu"))
    with pytest.raises(L.PinnkError):
        m(torch.zeros(4, 2))


def test_state_dict_keys_match_reference_layout():
    keys = {a: list(pk.make_model(a, 2, 32, 2, torch.device("cpu"), omega_0=30.0).state_dict()) for a in
            ["feedforward", "resnet", "siren", "fourier"]}
    assert keys["feedforward"][:2] == ["model.layers.0.weight", "model.layers.0.bias"]
    assert "model.blocks.0.layers.1.weight" in keys["resnet"] and "model.input_layer.weight" in keys["resnet"]
    assert keys["siren"][0] == "model.layers.0.linear.weight"
    assert keys["fourier"][0] == "model.fourier.B"

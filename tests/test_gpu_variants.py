"""GPU: the optimised routes against the plain ones of the same library, on identical seeded inputs.

  default                      tcgen05 kernels, pre-activation stash elided (reverse pass works from the OUTPUT jets),
                               specialised first / last layer kernels
  PINNK_KEEP_Z=1               pre-activations stashed, adjoint from the stashed z jets
  PINNK_DISABLE_EDGE_FAST=1    generic (run-time jet layout) first / last layer kernels (also keeps the stash)
  PINNK_DISABLE_OUT_FUSE=1     output layer nn.Linear(width, 1) as its own pass instead of folded into the last hidden
                               layer's epilogue
  PINNK_DISABLE_FIRST_FUSE=1   reverse of the input layer as its own kernel instead of inside the first hidden layer's dgrad
  PINNK_DISABLE_LOSS_FUSE=1    PDE residual, loss sums, seeds and the output layer's reverse as separate kernels instead of
                               inside the last hidden layer's forward epilogue
  PINNK_ENABLE_PAIR=1          dgrad + tanh adjoint and wgrad of a hidden layer in one launch of CTA pairs sharing the tile
                               stream (bwd_pair_kernel; measured slower than the two launches, so not the default)
  PINNK_DISABLE_TC=1           every GEMM on the exact-fp32 CUDA-core kernel, unfused activations

All eight must give the same residuals, loss components and parameter gradients to fp32 round-off; the model has one
layer scaled into saturation so that units with w0 = 1 - tanh^2 down to exactly 0 are exercised."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _run(tmp_path, tag, env):
    out = str(tmp_path / f"{tag}.npz")
    e = dict(os.environ)
    e.update(env)
    subprocess.check_call([sys.executable, os.path.join(HERE, "variant_grad.py"), out], env=e)
    return np.load(out)


def _rel(a, b):
    a, b = a.astype(np.float64).ravel(), b.astype(np.float64).ravel()
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def test_optimised_routes_agree_with_plain_routes(tmp_path):
    assert torch.cuda.is_available()
    from pinns_rl_pde_b200 import build
    build.build()
    base = _run(tmp_path, "default", {})
    exact = _run(tmp_path, "no_tc", {"PINNK_DISABLE_TC": "1"})
    for tag, env in (("keep_z", {"PINNK_KEEP_Z": "1"}), ("generic_edges", {"PINNK_DISABLE_EDGE_FAST": "1"}),
                     ("separate_output_layer", {"PINNK_DISABLE_OUT_FUSE": "1"}),
                     ("separate_input_layer_reverse", {"PINNK_DISABLE_FIRST_FUSE": "1"}),
                     ("separate_loss_kernels", {"PINNK_DISABLE_LOSS_FUSE": "1"}),
                     ("paired_dgrad_wgrad", {"PINNK_ENABLE_PAIR": "1"})):
        v = _run(tmp_path, tag, env)
        for k in base.files:
            assert np.all(np.isfinite(v[k])) and np.all(np.isfinite(base[k])), (tag, k)
            assert _rel(base[k], v[k]) <= 2e-6, (tag, k, _rel(base[k], v[k]))
    for k in base.files:                              # 3xTF32 tensor-core path vs exact fp32 FFMA path
        assert _rel(base[k], exact[k]) <= 5e-6, (k, _rel(base[k], exact[k]))

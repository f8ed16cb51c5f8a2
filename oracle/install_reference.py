"""Install the UNMODIFIED reference package (pinnrl) into ``baseline/_ref`` -- test infrastructure, not product.

    python oracle/install_reference.py            # build container only (needs /root/reference)

``baseline/_ref`` is git-ignored (never in history) but travels with the gpurun snapshot, so the GPU box can run the
reference's own ``PDETrainer`` / samplers side by side with the patched path (tests/test_gpu_reference_trainer.py) and
``bench.py --impl reference`` can time the reference's own CPU implementation (``cpu_baseline.kind = "reference"``).

The reference declares ``hatchling`` as its build backend, which is not in this image's offline wheelhouse, so the
install runs from a scratch copy under /tmp whose ``[build-system]`` table (build metadata only -- no source file is
touched) is pointed at setuptools:

    pip install --no-index --no-build-isolation --no-deps --target baseline/_ref /tmp/<copy>

``--no-deps``: matplotlib / plotly / gymnasium are absent from the image; ``oracle/ref_env.py`` stubs the plotting
modules the trainer imports at module scope (they are never called on the paths the tests and the bench exercise).
Nothing under the product package imports this file or ``baseline/_ref``.
"""
from __future__ import annotations

import os
import re
import shutil
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = "/root/reference"
TARGET = os.path.join(ROOT, "baseline", "_ref")


def installed() -> bool:
    return os.path.isfile(os.path.join(TARGET, "pinnrl", "__init__.py"))


def install(force: bool = False) -> str:
    """Returns one line describing the outcome (recorded in DESIGN.md section 8)."""
    if installed() and not force:
        return "already installed"
    if not os.path.isdir(REFERENCE):
        return "reference tree absent (GPU box): using the prebuilt baseline/_ref if it travelled"
    tmp = tempfile.mkdtemp(prefix="pinnrl_ref_")
    try:
        src = os.path.join(tmp, "reference")
        shutil.copytree(REFERENCE, src, ignore=shutil.ignore_patterns(".git", "notebooks", "docs", "__pycache__"))
        pp = os.path.join(src, "pyproject.toml")
        text = open(pp).read()
        text = re.sub(r"\[build-system\].*?(?=\n\[)", '[build-system]\nrequires = ["setuptools"]\n'
                      'build-backend = "setuptools.build_meta"\n', text, count=1, flags=re.S)
        text += '\n[tool.setuptools.packages.find]\ninclude = ["pinnrl*"]\n\n[tool.setuptools.package-data]\n"*" = ["*.yaml", "*.yml", "*.json"]\n'
        open(pp, "w").write(text)
        if os.path.isdir(TARGET):
            shutil.rmtree(TARGET)
        os.makedirs(os.path.dirname(TARGET), exist_ok=True)
        cmd = [sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "--find-links",
               "/opt/wheelhouse", "--target", TARGET, src]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0 or not installed():
            raise RuntimeError("pip install of the reference failed:\n" + res.stdout[-2000:] + res.stderr[-2000:])
        return "pip install --no-deps --target baseline/_ref (setuptools backend on a /tmp copy): ok"
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    print(install(force="--force" in sys.argv))

"""Import the UNMODIFIED reference (pinnrl 0.3.1) from ``baseline/_ref`` -- test infrastructure, not product.

``baseline/_ref`` is produced by ``oracle/install_reference.py`` in the build container, is git-ignored, and travels
to the GPU box with the gpurun snapshot.  Only ``tests/`` and ``bench.py --impl reference`` / ``cpu_baseline`` import
this module; nothing under ``pinns_rl_pde_b200/`` does.

The trainer / RL agent modules import matplotlib and plotly at module scope; neither is in the image.  They are
stubbed with ``MagicMock`` packages -- no code path the tests or the bench exercise calls them
(``PDETrainer.train(experiment_dir=None)`` skips every plot / file write).
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_DIR = os.path.join(ROOT, "baseline", "_ref")
_STUBS = ("matplotlib", "matplotlib.pyplot", "matplotlib.colors", "matplotlib.animation", "matplotlib.gridspec",
          "matplotlib.cm", "matplotlib.figure", "mpl_toolkits", "mpl_toolkits.mplot3d", "plotly", "plotly.graph_objects",
          "plotly.subplots", "plotly.express", "plotly.io", "gymnasium", "gymnasium.spaces", "seaborn")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_DIR, "pinnrl", "__init__.py"))


def activate():
    """Put the installed reference on sys.path (ahead of anything else named pinnrl) and stub the plotting deps.
    Returns the imported ``pinnrl`` package; raises ImportError when ``baseline/_ref`` is absent."""
    if not available():
        raise ImportError("baseline/_ref/pinnrl not found: run `python oracle/install_reference.py` in the build container")
    for name in _STUBS:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mm = MagicMock()
                mm.__path__ = []
                sys.modules[name] = mm
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import pinnrl
    got = os.path.realpath(os.path.dirname(pinnrl.__file__))
    if not got.startswith(os.path.realpath(REF_DIR)):
        raise ImportError(f"pinnrl resolved to {got}, not to baseline/_ref")
    return pinnrl

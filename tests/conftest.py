import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HERE = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_terminal_summary(terminalreporter, exitstatus, config):
    """Every measured parity error of the run (tests/parity_log.py), also for passing tests."""
    import parity_log
    if not parity_log.LINES:
        return
    terminalreporter.write_sep("=", "measured parity errors (SURVEY F9 gate: max(1e-5, 2 x fp32 reference floor))")
    for line in parity_log.LINES:
        terminalreporter.write_line(line)
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "parity_summary.log"), "a") as f:
            f.write("\n".join(parity_log.LINES) + "\n")

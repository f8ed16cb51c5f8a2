"""Measured parity errors collected during a test run; conftest.pytest_terminal_summary prints them after the result line
so that the driver's `pytest -q` log shows how far inside each gate the kernels are (passing tests' stdout is captured)."""
LINES = []


def log(line: str):
    LINES.append(line)
    print(line)

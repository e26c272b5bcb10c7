"""pytest configuration: registers the `gpu` marker and makes the repo importable.

`-m "not gpu"` : oracle vs golden vectors, host logic, C-ABI symbol export (CPU only).
`-m gpu`       : parity tests proper, calling the CUDA path through the C-ABI on a B200.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run by the driver at round end)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN

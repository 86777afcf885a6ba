import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import __graft_entry__ as entry  # noqa: E402


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (test infrastructure)."""
    return entry.load_oracle()


@pytest.fixture(scope="session")
def g():
    """The product package (host mirror over libgenb200.so); built on demand."""
    import subprocess
    subprocess.check_call([sys.executable, os.path.join(ROOT, "gen.jl_b200", "build.py")])
    return entry.load_package()

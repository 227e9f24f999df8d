import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def built():
    """Make sure the shared library and the C oracle exist (cross-compiles without a GPU)."""
    import subprocess

    from cgmres_cpp_b200._lib import LIB_PATH

    if not os.path.exists(LIB_PATH):
        subprocess.run(["make", "-C", os.path.join(ROOT, "cgmres_cpp_b200", "csrc")], check=True)
    from oracle import pyoracle as po

    po.load("port")
    return True


@pytest.fixture(scope="session")
def oracle_port(built):
    from oracle import pyoracle as po

    return po.load("port")


@pytest.fixture(scope="session")
def oracle_best(built):
    from oracle import pyoracle as po

    return po.best()

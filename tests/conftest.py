import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

FIX = os.path.join(HERE, "golden", "ref_fixtures")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def fixture_bytes(name):
    with open(os.path.join(FIX, name), "rb") as f:
        return f.read()


@pytest.fixture(scope="session")
def oracle():
    import oracle_binding
    oracle_binding.build()
    return oracle_binding


@pytest.fixture(scope="session")
def sim_engine():
    """Engine over the CPU *simulation* of the kernels (tests/sim): kernel-logic tests only."""
    so = os.path.join(HERE, "sim", "libbz2b200_sim.so")
    srcdir = os.path.join(ROOT, "compressjs_flattened_b200", "csrc")
    newest = max(os.path.getmtime(os.path.join(srcdir, f)) for f in os.listdir(srcdir))
    if not os.path.exists(so) or os.path.getmtime(so) < newest:
        subprocess.check_call([os.path.join(HERE, "sim", "build_sim.sh")])
    from compressjs_flattened_b200 import _native
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    return Bzip2Engine(0, _native.Library(so))


@pytest.fixture(scope="session")
def gpu_engine():
    """Engine over the real CUDA library; fails loudly if it is missing or no GPU is present."""
    from compressjs_flattened_b200.bzip2 import Bzip2Engine
    return Bzip2Engine(0)

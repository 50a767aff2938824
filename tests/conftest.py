import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def oracle_port():
    """Build (if needed) and return the oracle restatement driver class factory."""
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "port"], check=True)
    from oracle.hdrref import CpuStretch

    return lambda: CpuStretch("orc")


@pytest.fixture(scope="session")
def emu_libs():
    """CPU emulator builds of the CUDA sources (tests/cuda_emu) -- test infrastructure only."""
    subprocess.run(["sh", os.path.join(ROOT, "tests", "cuda_emu", "build.sh")], check=True)
    d = os.path.join(ROOT, "tests", "cuda_emu", "_build")
    return {"float": os.path.join(d, "libb200stretch_emu.so"), "exact": os.path.join(d, "libb200stretch_emu_exactfft.so")}


@pytest.fixture(scope="session")
def cuda_lib():
    """The product library, cross-compiled for sm_100a (no GPU needed to build or dlopen)."""
    import signalsmith_stretch_b200 as pkg

    return pkg.build_library()

import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


# exercise the composite (3 x 2^a) Bluestein lengths from L = 1536 up in the tests (the product default starts at 3072)
os.environ.setdefault("SB_FFT3_MINL", "2048")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def emu_lib():
    """TEST-ONLY CPU emulation build of the kernel sources (never loaded by the product package)."""
    from scythe_jl_b200 import _lib, build
    return _lib.load(build.build_emu())


@pytest.fixture(scope="session")
def gpu_lib():
    """The product library on a real device; fails (not skips) if it is missing."""
    from scythe_jl_b200 import _lib
    lib = _lib.load()
    assert lib.sb_device_count() > 0, "no CUDA device visible: -m gpu tests must run on the GPU box"
    return lib

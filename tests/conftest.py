import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "blokus-engine_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def orc():
    """The CPU oracle (test infrastructure)."""
    from oracle import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def emu_lib():
    """The kernel sources compiled for the CPU warp emulator (tests/warp_emu) — logic checks only."""
    d = os.path.join(ROOT, "tests", "warp_emu")
    subprocess.run(["make", "-C", d, "libblokus_emu.so"], check=True, capture_output=True)
    from blokus_self_play import Lib
    return Lib(os.path.join(d, "libblokus_emu.so"))


@pytest.fixture(scope="session")
def cuda_lib():
    """The sm_100a library on a real device.  Fails (not skips) if it is missing on a GPU box."""
    from blokus_self_play import Lib, DEFAULT_LIB
    lib = Lib(DEFAULT_LIB)
    assert not lib.missing, lib.missing
    lib.require_device()
    return lib

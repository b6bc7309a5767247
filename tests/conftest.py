import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run under gpurun)")


@pytest.fixture(scope="session")
def emu():
    """The SIMT-emulator build of the kernel sources (tests/simt_emu): CPU check of kernel LOGIC.
    Test infrastructure only -- the product library is zarc_b200/libzarcgpu.so."""
    from zarc_b200 import build, _lib

    return _lib.Lib(build.build_emu(), strict=False)


@pytest.fixture(scope="session")
def gpu():
    """The product library on a real GPU; fails loudly if the extension or the device is missing."""
    from zarc_b200 import _lib

    lib = _lib.lib()
    assert lib.zg_device_count() > 0, "no CUDA device visible"
    assert lib.zg_build_info() == b"sm_100a"
    return lib

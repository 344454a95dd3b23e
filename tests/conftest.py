"""Test configuration: `gpu` marker, import paths, shared helpers.

CPU tests (-m "not gpu") cover the oracle against the golden anchors, host logic, and that the C-ABI
library loads and exports every declared symbol.  GPU tests (-m gpu) are the parity tests proper:
CUDA path (through the C ABI) vs the oracle on the same seeded inputs.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

REFERENCE = "/root/reference"
HAS_REFERENCE = os.path.isdir(REFERENCE)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _has_gpu():
    try:
        import ctypes as C
        from prt_b200 import capi
        n = C.c_int()
        return capi.load().prt_device_count(C.byref(n)) == 0 and n.value > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    # A gpu-marked test on a box without a GPU is an error in how the suite was invoked, not a skip:
    # the driver selects with -m gpu / -m "not gpu".  Only skip when the user did not select by marker.
    if config.getoption("-m"):
        return
    if not _has_gpu():
        skip = pytest.mark.skip(reason="no CUDA device")
        for it in items:
            if "gpu" in it.keywords:
                it.add_marker(skip)


@pytest.fixture(scope="session")
def orc():
    import orc_py
    orc_py.build()
    return orc_py

"""tests/conftest.py -- markers, paths and the shared builders.

Tiers:  `-m "not gpu"`  oracle vs the reference's golden vectors, host logic through the host EMULATION of
                        the kernel source (tests/emu, scaffolding), ABI exports, 2-rank gloo sharding;
        `-m gpu`        the parity tests proper: CUDA library through the C ABI vs the oracle.
"""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_lib
    oracle_lib.build()
    return oracle_lib


@pytest.fixture(scope="session")
def emu_lib():
    """The product's host + kernel sources compiled for the host (tests/emu): logic checks without a GPU."""
    import ctypes
    from audiosdr_b200 import api
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d], check=True)
    return api._bind(ctypes.CDLL(os.path.join(d, "libsdr_emu.so")))


@pytest.fixture(scope="session")
def emu_contract_lib():
    """The same emulation built with -DSDR_CONTRACT: the arithmetic of the opt-in contracting build (tolerance tests)."""
    import ctypes
    from audiosdr_b200 import api
    d = os.path.join(ROOT, "tests", "emu")
    subprocess.run(["make", "-s", "-C", d, "libsdr_emu_contract.so"], check=True)
    return api._bind(ctypes.CDLL(os.path.join(d, "libsdr_emu_contract.so")))


@pytest.fixture(scope="session")
def cuda_lib():
    """The real thing: audiosdr_b200/libsdr_batch.so on a CUDA device.  No fallback: absence is a failure."""
    import torch
    assert torch.cuda.is_available(), "gpu-marked test running without a CUDA device"
    from audiosdr_b200 import api
    assert os.path.exists(api.lib_path()), "libsdr_batch.so missing: __graft_entry__.build() must run before the GPU tier"
    return api.load_library()

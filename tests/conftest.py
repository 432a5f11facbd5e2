import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def _ensure_built():
    """Build the parity checkers and the C-ABI library if they are missing (CPU box: nvcc
    cross-compiles). On the GPU box the prebuilt .so files travel with the snapshot."""
    from oracle import refpy
    if not os.path.exists(refpy.ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "oracle"])
    if not os.path.exists(refpy.REF_SO) and os.path.exists("/root/reference/src/LibHLA.cpp"):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "ref"])
    from hibag_b200 import api
    if not os.path.exists(api.LIB_PATH):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "hibag_b200", "csrc"), "-j", "8"])


@pytest.fixture(scope="session")
def built():
    _ensure_built()
    return True


@pytest.fixture(scope="session")
def orc(built):
    from oracle import refpy
    return refpy.OracleLib()


@pytest.fixture(scope="session")
def ref(built):
    """The compiled unmodified reference; pinned to the bit-exact target 'base'."""
    from oracle import refpy
    if not os.path.exists(refpy.REF_SO):
        pytest.skip("oracle/_ref/libhibag_ref.so not built (needs /root/reference at build time)")
    r = refpy.RefLib()
    r.set_target("base")
    r.set_gpu_procs(None)
    return r


@pytest.fixture(scope="session")
def gpu(built):
    from hibag_b200 import api
    if api.device_count() < 1:
        pytest.fail("no CUDA device: the -m gpu tests need the B200 box (no CPU fallback exists)")
    api.set_device(0)
    return api

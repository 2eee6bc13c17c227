import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests skip (instead of erroring) on a machine without a CUDA device."""
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    """The CPU checkers (oracle/), built on demand."""
    from oracle import joint_oracle
    if not os.path.exists(os.path.join(ROOT, "oracle", "liboracle.so")):
        joint_oracle.build()
    joint_oracle.lib()
    return joint_oracle


def _build_host_sim(name, defines=()):
    import ctypes
    src = os.path.join(ROOT, "tests", "host_sim", "core_sim.cpp")
    so = os.path.join(ROOT, "tests", "host_sim", name)
    deps = [src] + [os.path.join(ROOT, "bayhunter_b200", "csrc", f)
                    for f in ("bh_common.cuh", "bh_math.cuh", "swd_core.cuh", "swd_general_core.cuh", "sampler_core.cuh", "rf_core.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off"] +
                       ["-D" + d for d in defines] + ["-o", so, src, "-lm"], check=True)
    return ctypes.CDLL(so)


@pytest.fixture(scope="session")
def host_sim():
    """Host lock-step simulation of the device cores, device formulation of the secular functions."""
    return _build_host_sim("libcore_sim.so")


@pytest.fixture(scope="session")
def host_sim_reforder():
    """Same, with the secular functions in the Fortran's operation order (bit-equal to the oracle)."""
    return _build_host_sim("libcore_sim_ref.so", ("BH_SECULAR_REFERENCE_ORDER",))


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def has_cuda():
    try:
        from bayhunter_b200 import _lib
        return _lib.load().bh_device_count() > 0
    except Exception:
        return False

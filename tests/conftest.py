import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
GOLDEN_CASES = ["pairA", "pairB", "synth_clean", "synth_outage", "synth_outage_start",
                "synth_sharp_turn", "synth_sparse_gnss", "synth_long"]


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


@pytest.fixture(scope="session")
def hostmath():
    """tests/hostmath/libhostmath.so: the kernels' scalar math compiled with g++."""
    import ctypes
    import subprocess
    d = os.path.join(ROOT, "tests", "hostmath")
    so = os.path.join(d, "libhostmath.so")
    src = os.path.join(d, "hostmath.cpp")
    deps = [src] + [os.path.join(ROOT, "gps_optimize_slam_b200", "csrc", f) for f in ("gsf_common.cuh", "gsf_ekf_strict.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(p) > os.path.getmtime(so) for p in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-x", "c++", src, "-o", so])
    lib = ctypes.CDLL(so)
    lib.hm_yaw_zyx.restype = ctypes.c_double
    return lib

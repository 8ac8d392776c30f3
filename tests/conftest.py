import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # native pieces are built in-tree; (re)build what is missing so that a fresh checkout can run the suite
    if not os.path.exists(os.path.join(ROOT, "oracle", "librd_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    lib = os.path.join(ROOT, "rectdetect_b200", "librectdetect_b200.so")
    syn = os.path.join(ROOT, "rectdetect_b200", "librd_synth.so")
    if not (os.path.exists(lib) and os.path.exists(syn)):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "rectdetect_b200", "csrc"), "-j8"])


@pytest.fixture(scope="session")
def rd():
    import rectdetect_b200
    return rectdetect_b200


@pytest.fixture(scope="session")
def gpu_dev(rd):
    if rd.device_count() < 1:
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    d = rd.Device(0)
    yield d
    d.close()

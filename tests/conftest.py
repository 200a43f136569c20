import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_libraries():
    """Build the oracle (always) and the product library (if missing) before any test runs."""
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    if not os.path.exists(os.path.join(ROOT, "yoxel-voxel_b200", "libyv_b200.so")):
        subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "yoxel-voxel_b200", "csrc")],
                              stdout=subprocess.DEVNULL)
    yield


def has_gpu():
    import yoxel_voxel_b200 as yv
    try:
        return yv.device_count() > 0
    except Exception:
        return False

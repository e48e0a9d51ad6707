import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    # build the native libraries if they are not in-tree yet (nvcc cross-compiles without a GPU)
    import aa_admm_b200 as A
    if not (os.path.exists(A.LIB_CUDA) and os.path.exists(A.LIB_HOST)):
        A.build()
    port = os.path.join(ROOT, "oracle", "liboracle_port.so")
    if not os.path.exists(port) or (os.path.isdir("/root/reference") and not os.path.isdir(os.path.join(ROOT, "oracle", "_ref"))):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "all"], check=False,
                       stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def A():
    import aa_admm_b200
    return aa_admm_b200


@pytest.fixture(scope="session")
def gpu(A):
    if A.device_count() <= 0:
        pytest.skip("no CUDA device")
    return A


@pytest.fixture(scope="session")
def ref():
    from oracle import refbind
    if not refbind.have_ref():
        pytest.skip("oracle/_ref (compiled reference) not present")
    return refbind

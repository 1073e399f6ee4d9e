import os
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(REPO, "advanced-cpu-raytracing_b200"))
sys.path.insert(0, os.path.join(REPO, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "ref: needs the compiled reference oracle/_ref (skipped when absent)")


@pytest.fixture(scope="session", autouse=True)
def _build_native():
    """CPU-side pieces (host mirror + C oracle) are built on demand; the CUDA library is built by
    __graft_entry__.build() / make and must already exist for -m gpu."""
    import subprocess
    need = [os.path.join(REPO, "advanced-cpu-raytracing_b200", "libdthost.so"), os.path.join(REPO, "oracle", "libdtoracle.so")]
    if not all(os.path.exists(p) for p in need):
        subprocess.run(["make", "-C", REPO, "host", "oracle"], check=True, stdout=subprocess.DEVNULL)
    yield

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built_oracle():
    """The oracle's C shim must exist (built by `make -C oracle` / __graft_entry__.build())."""
    so = os.path.join(ROOT, "oracle", "_build", "libsicmath.so")
    if not os.path.isfile(so):
        import subprocess
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
    yield


# GPU tests that have not yet passed on a B200 run LAST under `-m gpu -x`, so that a surprise in them cannot hide the
# verified parity tests behind an early stop.  Empty: every GPU test passed on a B200 (gpurun_out/r2_gputest*.log).
_FIRST_TIME_ON_GPU = ()


def pytest_collection_modifyitems(config, items):
    def late(item):
        return any(item.name == n or item.name.startswith(n + "[") for n in _FIRST_TIME_ON_GPU)
    items[:] = [i for i in items if not late(i)] + [i for i in items if late(i)]

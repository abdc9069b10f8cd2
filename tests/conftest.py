import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: long-running (exhaustive sweeps); opt in with -m slow")


@pytest.fixture(scope="session")
def workdir(tmp_path_factory):
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else None
    if base:
        import tempfile
        d = tempfile.mkdtemp(prefix="qtipb200_", dir=base)
        yield d
        import shutil
        shutil.rmtree(d, ignore_errors=True)
    else:
        yield str(tmp_path_factory.mktemp("qtipb200"))

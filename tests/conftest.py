import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def _ensure_built():
    import __graft_entry__ as g

    g.build_if_needed()


@pytest.fixture(scope="session")
def oracle():
    _ensure_built()
    from oracle import oracle as o

    return o


@pytest.fixture(scope="session")
def lsio():
    _ensure_built()
    import lsio_b200

    return lsio_b200


@pytest.fixture(scope="session")
def cuda_dev():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


def data_dirs():
    """Directories to exercise: tmpfs always (accepts any O_DIRECT alignment), plus a disk-backed one."""
    dirs = []
    if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK):
        dirs.append("/dev/shm")
    dirs.append(os.environ.get("LSIO_TEST_DISK_DIR", "/tmp"))
    return dirs


@pytest.fixture(params=data_dirs())
def data_dir(request, tmp_path_factory):
    import tempfile

    d = tempfile.mkdtemp(prefix="lsio_test_", dir=request.param)
    yield d
    import shutil

    shutil.rmtree(d, ignore_errors=True)

import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Make sure the native libraries exist (built in-tree; they travel to the GPU box)."""
    pkg = os.path.join(ROOT, "cuda-vp9_b200")
    need = [os.path.join(pkg, "libvp9b200_host.so"), os.path.join(ROOT, "oracle", "_build", "libvp9ref.so")]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()


def clip_names():
    d = os.path.join(GOLDEN, "clips")
    return sorted(f[:-4] for f in os.listdir(d) if f.endswith(".ivf"))


def clip_path(name):
    return os.path.join(GOLDEN, "clips", name + ".ivf")


def golden_md5(name):
    with open(os.path.join(GOLDEN, "md5", name + ".md5")) as f:
        return [l.split()[0] for l in f if l.strip()]

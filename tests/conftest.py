import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def built_libraries():
    """Make sure the oracle, the CUDA library and the host library exist (no GPU needed to build)."""
    need = [
        os.path.join(ROOT, "oracle", "liboracle.so"),
        os.path.join(ROOT, "smpl_b200", "lib", "libsmplgpu.so"),
        os.path.join(ROOT, "smpl_b200", "lib", "libsmplhost.so"),
    ]
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__
        __graft_entry__.build()
    yield


def has_gpu():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True, timeout=20)
        return out.returncode == 0 and "GPU" in out.stdout
    except Exception:
        return False

"""CPU tests of the drop-in boundary: the C-ABI libraries load and export every symbol
include/*.h declares; without a GPU the product path fails loudly (no fallback)."""
import ctypes as C
import os
import re

import pytest

from conftest import ROOT, has_gpu
from smpl_b200 import api


def declared(header, prefix):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(%s_[a-z0-9_]+)\s*\(" % prefix, src)))


def test_smplgpu_exports_every_declared_symbol():
    L = api.gpu_lib()
    names = declared("smplgpu.h", "smplgpu")
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libsmplgpu.so does not export " + n


def test_smplhost_exports_every_declared_symbol():
    H = api.host_lib()
    names = declared("smplhost.h", "smplhost")
    assert len(names) >= 15
    for n in names:
        assert hasattr(H, n), "libsmplhost.so does not export " + n


def test_cuda_library_is_sm100a_only():
    out = os.popen("cuobjdump -lelf %s 2>/dev/null" % os.path.join(ROOT, "smpl_b200", "lib", "libsmplgpu.so")).read()
    if not out.strip():
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out


@pytest.mark.skipif(has_gpu(), reason="checks the no-GPU failure mode")
def test_no_gpu_means_loud_failure_not_fallback():
    with pytest.raises(api.SmplGpuError) as e:
        api.GpuContext(0)
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_null_context_is_rejected():
    L = api.gpu_lib()
    assert L.smplgpu_synchronize(None) < 0
    assert L.smplgpu_bfs_last_levels(None) < 0
    assert L.smplgpu_is_states_valid(None, None, 0, None) < 0

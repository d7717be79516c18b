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


def test_round2_entry_points_reject_a_null_context():
    """the entry points added in round 2 follow the same convention: negative code, no crash, no fallback"""
    L = api.gpu_lib()
    assert L.smplgpu_set_motion_primitives(None, None, 0) < 0
    assert L.smplgpu_expand_state(None, None, 0, None) < 0
    assert L.smplgpu_scene_epoch(None) < 0
    assert L.smplgpu_set_lattice(None, None, None) < 0
    assert L.smplgpu_is_lattice_states_valid(None, None, 0, None) < 0
    assert L.smplgpu_is_lattice_edges_valid(None, None, None, 0, None, 0, None, None) < 0
    assert L.smplgpu_collision_distance(None, None, 0, None) < 0
    assert L.smplgpu_reserve_distance_field(None, 1, 1, 1, None, None) < 0
    assert L.smplgpu_set_distance_field_l2_persistence(None, 1, None) < 0
    for name in ("smplgpu_lattice_max_slots", "smplgpu_lattice_create", "smplgpu_lattice_begin", "smplgpu_lattice_expand_submit",
                 "smplgpu_lattice_expand_wait", "smplgpu_lattice_states"):
        assert hasattr(L, name)
    L.smplgpu_lattice_max_slots.argtypes = [C.c_void_p, C.c_int]
    assert L.smplgpu_lattice_max_slots(None, 100) < 0
    L.smplgpu_lattice_expand_submit.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    assert L.smplgpu_lattice_expand_submit(None, None, None, 0, 0) < 0


def test_succ_info_layout_matches_the_header():
    """api.SuccInfoC mirrors smplgpu_succ_info (include/smplgpu.h): 16 + 6 + 3 doubles, 3 ints, 4 bytes"""
    assert C.sizeof(api.SuccInfoC) == 25 * 8 + 3 * 4 + 4
    src = open(os.path.join(ROOT, "include", "smplgpu.h")).read()
    assert "#define SMPLGPU_MAX_DOF 16" in src

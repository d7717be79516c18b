"""Pins oracle/distance_map.cpp (the field every validity verdict reads) to the REFERENCE's own
EuclidDistanceMap (smpl/src/distance_map/*.cpp + detail/distance_map.hpp, compiled where they lie by
`make -C oracle ref`): the whole squared-distance field after adds, incremental adds and removals."""
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle
from oracle_api import RefDistanceMap, ref_distmap_lib
from smpl_b200 import scenes

GOLD = os.path.join(ROOT, "tests", "golden", "distmap_reference.npz")
needs_ref = pytest.mark.skipif(ref_distmap_lib() is None, reason="oracle/_ref/libref_distmap.so not built (needs /root/reference)")


def small_scene():
    """PR2 clutter scene geometry at a size the reference propagates in about a second."""
    s = scenes.Scene("pr2", "right_arm", scenes.PR2_RIGHT_ARM_JOINTS, (-0.3, -0.8, 0.2), (1.0, 1.2, 0.9), 0.02, 0.3)
    scenes._pr2_common(s)
    rng = np.random.Generator(np.random.PCG64(5))
    for _ in range(14):
        size = rng.uniform(0.04, 0.3, 3)
        center = np.array(s.origin) + rng.uniform(0.0, 1.0, 3) * np.array(s.size)
        s.add_box(center, size)
    return s


def obstacle_points(o):
    """World points (cell centres as the reference's gridToWorld defines them) of the oracle's obstacle cells."""
    _, origin, res, _ = o.grid_info()
    cells = np.argwhere(o.df_d2() == 0)
    return origin + cells * res


@needs_ref
def test_field_equals_reference_field_after_adds_and_removals():
    scene = small_scene()
    o = make_oracle(scene)                 # world boxes + the voxels of the robot's fixed links
    ref = RefDistanceMap(scene.origin, scene.size, scene.res, scene.max_dist)
    assert ref.dims == tuple(int(d) for d in o.grid_info()[0])
    empty = ref.d2()
    assert empty.max() > 0 and empty[0, 0, 0] == 1      # border cells count as obstacles: distance 1 next to the border
    pts = obstacle_points(o)
    assert all(np.array_equal(ref.world_to_grid(p), c) for p, c in zip(pts[:50], o.world_to_grid(pts[:50])))
    ref.add_points(pts)
    assert np.array_equal(ref.d2(), o.df_d2())
    # incremental adds (points inside, outside, duplicates)
    rng = np.random.default_rng(8)
    more = np.array(scene.origin) - 0.05 + rng.random((300, 3)) * (np.array(scene.size) + 0.1)
    more = np.concatenate([more, more[:20]])
    o.add_points(more)
    ref.add_points(more)
    assert np.array_equal(ref.d2(), o.df_d2())
    # removals: the propagating field is updated, not rebuilt
    o.remove_points(more[:150])
    ref.remove_points(more[:150])
    assert np.array_equal(ref.d2(), o.df_d2())
    # metric distance as CollisionSpace reads it
    q = np.array(scene.origin) - 0.1 + rng.random((300, 3)) * (np.array(scene.size) + 0.2)
    d2 = o.df_d2()
    g = o.world_to_grid(q)
    inb = np.all((g >= 0) & (g < np.array(d2.shape)), axis=1)
    gc = np.clip(g, 0, np.array(d2.shape) - 1)
    want = np.where(inb, scene.res * np.sqrt(d2[gc[:, 0], gc[:, 1], gc[:, 2]].astype(np.float64)), 0.0)   # outside: 0
    got = np.array([ref.distance(*p) for p in q])
    assert np.array_equal(got, want) and inb.any() and not inb.all()
    ref.close()


@needs_ref
def test_empty_field_and_capping():
    o_scene = scenes.Scene("pr2", "right_arm", scenes.PR2_RIGHT_ARM_JOINTS, (0.0, 0.0, 0.0), (0.5, 0.4, 0.6), 0.02, 0.1)
    ref = RefDistanceMap(o_scene.origin, o_scene.size, o_scene.res, o_scene.max_dist)
    d = ref.d2()
    dmax = int(np.ceil(0.1 / 0.02))
    assert d.max() == dmax * dmax                      # capped at ceil(max_dist / res)^2
    ref.add_points([[0.25, 0.2, 0.3]])
    c = ref.world_to_grid([0.25, 0.2, 0.3])
    assert ref.d2()[tuple(c)] == 0
    ref.close()


def test_golden_fixture_from_reference_build():
    """Field of the reference build for the small scene's obstacle set (tools/gen_golden.py)."""
    g = np.load(GOLD)
    scene = small_scene()
    o = make_oracle(scene)
    assert np.array_equal(o.df_d2().astype(np.uint16), g["d2_after_add"])
    rng = np.random.default_rng(8)
    more = np.array(scene.origin) - 0.05 + rng.random((300, 3)) * (np.array(scene.size) + 0.1)
    more = np.concatenate([more, more[:20]])
    o.add_points(more)
    o.remove_points(more[:150])
    assert np.array_equal(o.df_d2().astype(np.uint16), g["d2_after_remove"])


"""Pins oracle/voxelize.{h,cpp} (scene ingest) to the REFERENCE's own voxeliser
(smpl/src/geometry/voxelize.cpp + mesh_utils.cpp compiled where they lie against the arithmetic Eigen stand-in of
oracle/ref_stubs/eigen_arith): the same voxel centres in the same order for boxes at arbitrary poses, random triangle
soups, degenerate triangles, both voxel grids (pivot / half-res) and with ScanFill."""
import os

import numpy as np
import pytest

from conftest import ROOT
from oracle_api import box_mesh, ref_voxelize_lib, voxelize_box, voxelize_mesh

GOLD = os.path.join(ROOT, "tests", "golden", "voxelize_reference.npz")
needs_ref = pytest.mark.skipif(ref_voxelize_lib() is None, reason="oracle/_ref/libref_voxelize.so not built (needs /root/reference)")


def random_pose(rng, spread=1.0):
    """A rigid pose 3x4 from a random unit quaternion (or an axis-aligned one every third draw)."""
    q = rng.normal(size=4)
    q /= np.linalg.norm(q)
    w, x, y, z = q
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    if rng.integers(0, 3) == 0:
        R = np.eye(3)
    return np.concatenate([R, rng.uniform(-spread, spread, (3, 1))], axis=1)


def box_case(seed):
    rng = np.random.default_rng(seed)
    size = rng.uniform(0.03, 0.6, 3)
    if seed % 5 == 0:
        size = np.round(size / 0.02) * 0.02          # faces on cell boundaries
    pose = random_pose(rng)
    if seed % 7 == 0:
        pose[:, 3] = np.round(pose[:, 3] / 0.02) * 0.02
    res = float(rng.choice([0.02, 0.01, 0.025 / np.sqrt(2), 0.05]))
    origin = None if seed % 4 == 3 else rng.uniform(-1.0, 0.0, 3)
    return size, pose, res, origin, bool(seed % 6 == 1)


def soup_case(seed):
    rng = np.random.default_rng(1000 + seed)
    nv = int(rng.integers(3, 30))
    v = rng.uniform(-0.4, 0.4, (nv, 3))
    t = rng.integers(0, nv, (int(rng.integers(1, 24)), 3))
    t[0] = (0, 0, 1)                                  # degenerate: two equal vertices
    if len(t) > 2:
        v[t[1][2]] = 0.5 * (v[t[1][0]] + v[t[1][1]])  # colinear
    res = float(rng.choice([0.02, 0.04]))
    origin = None if seed % 2 else rng.uniform(-1.0, 0.0, 3)
    return v, t, res, origin, bool(seed % 3 == 0)


N_BOX, N_SOUP = 36, 16


@needs_ref
def test_box_mesh_and_box_voxels_equal_reference():
    v0, t0 = box_mesh("reference", (0.3, 0.2, 0.1))
    v1, t1 = box_mesh("oracle", (0.3, 0.2, 0.1))
    assert np.array_equal(v0, v1) and np.array_equal(t0, t1)
    total = 0
    for seed in range(N_BOX):
        size, pose, res, origin, fill = box_case(seed)
        ref = voxelize_box("reference", size, pose, res, origin, fill)
        got = voxelize_box("oracle", size, pose, res, origin, fill)
        assert ref.shape == got.shape and np.array_equal(ref, got), seed
        total += len(ref)
    assert total > 20000


@needs_ref
def test_triangle_soups_equal_reference():
    total = 0
    for seed in range(N_SOUP):
        v, t, res, origin, fill = soup_case(seed)
        ref = voxelize_mesh("reference", v, t, res, origin, fill)
        got = voxelize_mesh("oracle", v, t, res, origin, fill)
        assert ref.shape == got.shape and np.array_equal(ref, got), seed
        total += len(ref)
    assert total > 5000


def test_surface_voxels_of_an_aligned_box_form_a_shell():
    # 0.2 m cube centred on a cell centre at 2 cm: a closed shell one to two cells thick, empty inside
    pose = np.concatenate([np.eye(3), [[0.5], [0.5], [0.5]]], axis=1)
    vox = voxelize_box("oracle", (0.2, 0.2, 0.2), pose, 0.02, (0.0, 0.0, 0.0), False)
    cells = np.rint(vox / 0.02).astype(int)
    assert np.abs(vox - cells * 0.02).max() < 1e-12       # pivot grid: centres on origin + i res
    d = np.abs(cells - 25).max(axis=1)
    assert d.max() <= 6 and d.min() >= 4
    filled = voxelize_box("oracle", (0.2, 0.2, 0.2), pose, 0.02, (0.0, 0.0, 0.0), True)
    assert len(filled) > len(vox) and np.abs(np.rint(filled / 0.02).astype(int) - 25).max(axis=1).min() == 0


def test_golden_fixture_from_reference_build():
    g = np.load(GOLD)
    for seed in range(N_BOX):
        size, pose, res, origin, fill = box_case(seed)
        assert np.array_equal(voxelize_box("oracle", size, pose, res, origin, fill), g["box_%d" % seed]), seed
    for seed in range(N_SOUP):
        v, t, res, origin, fill = soup_case(seed)
        assert np.array_equal(voxelize_mesh("oracle", v, t, res, origin, fill), g["soup_%d" % seed]), seed


SHAPES = [(0, (0.3, 0.2, 0.1)), (1, (0.17,)), (2, (0.09, 0.42)), (3, (0.12, 0.3))]


@needs_ref
@pytest.mark.parametrize("kind,dims", SHAPES)
def test_primitive_meshes_equal_reference(kind, dims):
    """Oracle restatement AND the product's host-side builder (smplhost_shape_meshes) against the reference's
    CreateIndexed*Mesh: same vertices (bitwise) and the same triangles in the same order."""
    from oracle_api import shape_mesh
    from smpl_b200 import api
    v_ref, t_ref = shape_mesh("reference", kind, dims)
    v_or, t_or = shape_mesh("oracle", kind, dims)
    assert np.array_equal(v_ref, v_or) and np.array_equal(t_ref, t_or)
    d = np.zeros(3)
    d[:len(dims)] = dims
    row = np.concatenate([[kind], d, np.eye(4)[:3].ravel()])
    v, t = api.shape_meshes(row[None, :])
    assert np.array_equal(v, v_ref) and np.array_equal(t, t_ref)
    # two shapes in one call: the second one's triangles are offset by the first one's vertices, the pose moves it
    pose = np.concatenate([np.eye(3), [[0.5], [-0.25], [1.0]]], axis=1)
    v2, t2 = api.shape_meshes(np.stack([row, np.concatenate([[kind], d, pose.ravel()])]))
    assert np.array_equal(v2[:len(v)], v_ref) and np.array_equal(v2[len(v):], v_ref + pose[:, 3])
    assert np.array_equal(t2[len(t):], t_ref + len(v))
    # VoxelizeSphere / VoxelizeCylinder / VoxelizeCone = this mesh through VoxelizeMesh
    for res, origin in ((0.02, (-0.5, -1.0, 0.0)), (0.025 / np.sqrt(2), (0.0, 0.0, 0.0)), (0.01, None)):
        vox_ref = voxelize_mesh("reference", v2[len(v):], t_ref, res, origin, False)
        vox = voxelize_mesh("oracle", v2[len(v):], t_ref, res, origin, False)
        assert len(vox_ref) > 100 and np.array_equal(vox, vox_ref)

"""Scene ingest on the device: the voxels the CUDA voxeliser finds for a mesh, and the
distance field built from a scene's box objects, must equal the reference-shaped CPU code of the oracle
(oracle/voxelize.cpp, pinned against the reference's own voxelize.cpp) voxel for voxel and cell for cell."""
import numpy as np
import pytest

from helpers import make_oracle
from oracle_api import voxelize_box, voxelize_mesh
from smpl_b200 import api, scenes
from test_oracle_voxelize import N_BOX, N_SOUP, box_case, soup_case

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = api.GpuContext(0)
    yield c
    c.close()


def test_box_voxels_equal_oracle(ctx):
    total = 0
    for seed in range(N_BOX):
        size, pose, res, origin, fill = box_case(seed)
        v, t = api.box_meshes(np.concatenate([size, pose.ravel()])[None, :])
        got = ctx.voxelize_mesh(v, t, res, origin)
        ref = voxelize_box("oracle", size, pose, res, origin, False)
        assert got.shape == ref.shape and np.array_equal(got, ref), seed    # same centres, same (ExtractVoxels) order
        total += len(got)
    assert total > 20000


def test_triangle_soups_equal_oracle(ctx):
    total = 0
    for seed in range(N_SOUP):
        v, t, res, origin, _ = soup_case(seed)
        got = ctx.voxelize_mesh(v, t, res, origin)
        ref = voxelize_mesh("oracle", v, t, res, origin, False)
        assert got.shape == ref.shape and np.array_equal(got, ref), seed
        total += len(got)
    assert total > 5000


@pytest.mark.parametrize("kind,dims", [(1, (0.17, 0, 0)), (2, (0.09, 0.42, 0)), (3, (0.12, 0.3, 0))])
def test_sphere_cylinder_cone_voxels_equal_oracle(ctx, kind, dims):
    rng = np.random.default_rng(kind)
    from test_oracle_voxelize import random_pose
    for res, origin in ((0.02, (-0.5, -1.0, 0.0)), (0.025 / np.sqrt(2), (0.0, 0.0, 0.0)), (0.01, None)):
        pose = random_pose(rng, 0.5)
        v, t = api.shape_meshes(np.concatenate([[kind], dims, pose.ravel()])[None, :])
        got = ctx.voxelize_mesh(v, t, res, origin)
        ref = voxelize_mesh("oracle", v, t, res, origin, False)
        assert len(ref) > 100 and np.array_equal(got, ref)


def test_degenerate_and_empty_meshes(ctx):
    v = np.array([[0.0, 0, 0], [1, 0, 0], [2, 0, 0], [0.5, 0, 0]])
    assert len(ctx.voxelize_mesh(v, [[0, 1, 2], [0, 0, 3]], 0.05, (0, 0, 0))) == 0      # colinear / repeated vertex
    assert len(ctx.voxelize_mesh(v, np.zeros((0, 3), np.int32), 0.05, (0, 0, 0))) == 0
    with pytest.raises(api.SmplGpuError):
        ctx.voxelize_mesh(v, [[0, 1, 4]], 0.05, (0, 0, 0))                              # vertex index out of range
    # a large thin triangle: many candidate cells per triangle
    big = np.array([[0.0, 0.0, 0.013], [3.0, 0.1, 0.013], [0.2, 2.5, 0.4]])
    got = ctx.voxelize_mesh(big, [[0, 1, 2]], 0.01, (0.0, 0.0, 0.0))
    ref = voxelize_mesh("oracle", big, [[0, 1, 2]], 0.01, (0.0, 0.0, 0.0), False)
    assert len(ref) > 40000 and np.array_equal(got, ref)


@pytest.mark.parametrize("make_scene", [scenes.pr2_tabletop_env_scene, scenes.pr2_shelf_objects_scene])
def test_distance_field_from_box_objects(make_scene):
    """Ingest parity is exact: the occupied cells are the reference's.  The field built on the device is the EXACT
    Euclidean transform of that set; the reference's propagating distance map (pinned in
    tests/test_oracle_distance_map.py) over-estimates a few cells per million by one unit of d^2 when obstacles are
    not axis-aligned (directional vector propagation), so away from obstacles exact <= reference, and verdicts are
    compared both ways: device-built field, and the reference's own field uploaded (the drop-in path)."""
    scene = make_scene()
    o = make_oracle(scene, with_kdl=False)
    c, tables = api.setup_context(scene)
    try:
        ref = o.df_d2()
        got = c.download_distance_field().astype(np.int32)
        assert got.shape == ref.shape
        assert np.array_equal(got == 0, ref == 0), "occupied cells differ"
        assert (ref == 0).sum() > 1000
        differ = got != ref
        assert np.all(got <= ref) and np.all(ref[differ] - got[differ] <= 2)
        assert differ.sum() <= 1e-5 * ref.size, "%d cells differ" % int(differ.sum())
        lo, hi, cont = tables.limits()
        q = scenes.random_states(20000, lo, hi, cont, seed=4)
        v_ref = o.is_states_valid(q)
        v = c.is_states_valid(q)
        assert (v != v_ref).sum() <= differ.sum()           # a flip needs a sphere centre in one of those cells
        assert 0.05 < v.mean() < 0.95
        _, origin, res, dmax_sq = o.grid_info()
        c.set_distance_field(ref.astype(np.uint16), origin, res, dmax_sq)
        assert np.array_equal(c.is_states_valid(q), v_ref)   # reference field in, reference verdicts out
        print("%s: %d occupied cells, %d / %d cells where the reference's propagation is inexact" % (
            make_scene.__name__, int((ref == 0).sum()), int(differ.sum()), ref.size))
    finally:
        c.close()


def test_surface_ingest_differs_from_filled_cells():
    """The reference ingests SURFACE voxels (fill = false): the inside of a thick box is free space in its field."""
    s = scenes.Scene("pr2", "right_arm", scenes.PR2_RIGHT_ARM_JOINTS, (-0.5, -1.0, 0.0), (2.0, 2.0, 2.0), 0.02, 0.4)
    scenes._pr2_common(s)
    s.add_box_object((0.9, 0.4, 1.0), (0.4, 0.4, 0.4))
    c, tables = api.setup_context(s)
    try:
        d2 = c.download_distance_field()
        centre = api.world_to_grid([[0.9, 0.4, 1.0]], s.origin, s.res)[0]
        face = api.world_to_grid([[0.7, 0.4, 1.0]], s.origin, s.res)[0]
        assert d2[tuple(face)] == 0 and d2[tuple(centre)] >= 64      # 10 cells from every face, minus the shell
    finally:
        c.close()


def test_attached_box_sphere_model_generated_from_voxels():
    """attachBody (attached_bodies_collision_model.cpp:264-309): the body's spheres sit on its surface voxels at
    0.025 / sqrt(2); generated on the device for the product, by the oracle's voxeliser for the check."""
    scene = scenes.ubr1_tabletop_scene(attach=False)
    pose = np.concatenate([np.eye(3), [[0.26], [0.0], [0.0]]], axis=1)      # the box in the wrist_roll_link frame
    size = (0.05, 0.05, 0.20)
    o = make_oracle(scene, with_kdl=False)
    n_ref = o.attach_box("object", "wrist_roll_link", size, pose)
    grasp = ("wrist_roll_link", "gripper_link", "left_gripper_finger_link", "right_gripper_finger_link")
    for link in grasp:
        o.acm_set("object", link, True)
    c = api.GpuContext(0)
    try:
        tables = api.build_tables(scene)
        n = tables.attach_box(c, "object", "wrist_roll_link", size, pose)
        assert n == n_ref and n > 50
        for link in grasp:
            tables.set_acm_entry("object", link, True)
        c.set_robot(tables)
        c.build_distance_field(api.scene_cells(scene, tables), scene.dims, scene.origin, scene.res, scene.max_dist,
                               scene.padding)
        lo, hi, cont = tables.limits()
        q = scenes.random_states(20000, lo, hi, cont, seed=9)
        v = c.is_states_valid(q)
        assert np.array_equal(v, o.is_states_valid(q))
        q0, q1 = scenes.mprim_edges(q[:4000])
        e, cnt = c.is_edges_valid(q0, q1)
        e_ref, cnt_ref = o.is_edges_valid(q0, q1)
        assert np.array_equal(e, e_ref) and np.array_equal(cnt, cnt_ref)
        assert 0.02 < v.mean() < 0.9
    finally:
        c.close()


def test_incremental_add_and_remove_on_the_resident_field():
    """OccupancyGrid::addPointsToField / removePointsFromField (occupancy_grid.cpp:357-410 -> distance_map.hpp:305-367)
    on the device: cells leave and enter the obstacle set of the resident field (objects removed from / inserted into
    the scene).  Checked against the oracle's restatement of the reference's incremental propagation (pinned to the
    reference build in tests/test_oracle_distance_map.py) with the same exactness caveat as the bulk build, against a
    fresh device build of the final set (bit for bit), on a field uploaded from outside, and through the verdicts."""
    scene = scenes.pr2_tabletop_env_scene()
    o = make_oracle(scene, with_kdl=False)
    c, tables = api.setup_context(scene)
    try:
        dims, origin, res, dmax_sq = o.grid_info()
        rng = np.random.default_rng(21)
        occ = np.argwhere(o.df_d2() == 0)
        gone = occ[rng.choice(len(occ), len(occ) // 3, replace=False)]          # an object taken away ...
        new = np.stack([rng.integers(-2, d + 2, 400) for d in dims], axis=1)    # ... clutter added (some cells outside)
        new[:, 0] = rng.integers(2 * dims[0] // 3, dims[0] + 2, len(new))       # in the far third of the workspace
        new = np.concatenate([new, new[:100], occ[:50]])                        # duplicates, cells that are obstacles already
        world = lambda cells: np.asarray(origin) + np.asarray(cells) * res      # cell centres, as gridToWorld defines them
        lo, hi, cont = tables.limits()
        q = scenes.random_states(20000, lo, hi, cont, seed=5)
        v_before = c.is_states_valid(q)
        o.remove_points(world(gone))
        c.distance_field_remove_cells(gone)
        o.add_points(world(new))
        c.distance_field_add_cells(new)
        ref = o.df_d2()
        got = c.download_distance_field().astype(np.int32)
        assert np.array_equal(got == 0, ref == 0), "obstacle sets differ"
        differ = got != ref
        assert np.all(got <= ref) and np.all(ref[differ] - got[differ] <= 2)
        assert differ.sum() <= 2e-5 * ref.size, "%d cells differ" % int(differ.sum())
        # the same set built from scratch on the device: bit for bit
        c2, _ = api.setup_context(scene)
        try:
            c2.build_distance_field(np.argwhere(ref == 0), scene.dims, scene.origin, scene.res, scene.max_dist, scene.padding)
            assert np.array_equal(c2.download_distance_field(), got)
        finally:
            c2.close()
        v_ref = o.is_states_valid(q)
        v = c.is_states_valid(q)
        assert (v != v_ref).sum() <= differ.sum()
        assert 0.0 < v.mean() < 0.95 and (v != v_before).sum() > 20      # the update is what the verdicts see
        # on a field that came from outside (the reference's own), removing and re-adding the same cells
        c.set_distance_field(ref.astype(np.uint16), origin, res, dmax_sq)
        some = np.argwhere(ref == 0)[::7]
        c.distance_field_remove_cells(some)
        assert (c.download_distance_field()[tuple(some.T)] > 0).all()
        c.distance_field_add_cells(some)
        again = c.download_distance_field().astype(np.int32)
        assert np.array_equal(again == 0, ref == 0) and np.all(again <= ref) and (again != ref).sum() <= 2e-5 * ref.size
        print("incremental update: %d cells removed, %d added, %d / %d cells where the reference's propagation is inexact" % (
            len(gone), len(new), int(differ.sum()), ref.size))
    finally:
        c.close()

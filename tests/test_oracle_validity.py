"""CPU tests: the oracle against its committed golden vectors, the host-side model builder
(product) against the oracle's model tables, and oracle-internal consistency properties."""
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle
from smpl_b200 import api, scenes

GOLD = os.path.join(ROOT, "tests", "golden")


def load_csv(name):
    return np.loadtxt(os.path.join(GOLD, name))


@pytest.fixture(scope="module")
def pr2():
    scene = scenes.pr2_clutter_scene()
    return scene, make_oracle(scene)


def test_golden_pr2_states(pr2):
    _, o = pr2
    g = load_csv("pr2_right_arm_validity.csv")
    v = o.is_states_valid(g[:, :7])
    assert np.array_equal(v, g[:, 7].astype(np.uint8))
    assert 0.2 < v.mean() < 0.8


def test_golden_pr2_edges(pr2):
    _, o = pr2
    g = load_csv("pr2_right_arm_edges.csv")
    v, c = o.is_edges_valid(g[:, :7], g[:, 7:14])
    assert np.array_equal(v, g[:, 14].astype(np.uint8))
    assert np.array_equal(c, g[:, 15].astype(np.int32))


def test_golden_ubr1_attached():
    scene = scenes.ubr1_tabletop_scene()
    o = make_oracle(scene)
    g = load_csv("ubr1_attached_validity.csv")
    assert np.array_equal(o.is_states_valid(g[:, :7]), g[:, 7].astype(np.uint8))


def test_early_out_and_exhaustive_evaluation_agree(pr2):
    """The verdict must not depend on the reference's DFS order (SURVEY.md section 7 'order-independence')."""
    scene, o = pr2
    lo, hi, cont = o.joint_limits()
    q = scenes.random_states(3000, lo, hi, cont, seed=5)
    v = o.is_states_valid(q)
    v2, L, cm, pm = o.report_states(q)
    assert np.array_equal(v, v2)
    assert L.min() >= 8 and L.max() <= 38  # PR2 right arm: 8 trees, 38 nodes
    q0, q1 = scenes.mprim_edges(q)
    e, c = o.is_edges_valid(q0, q1)
    e2, c2, _ = o.report_edges(q0, q1)
    assert np.array_equal(e, e2) and np.array_equal(c, c2)


def test_edge_semantics(pr2):
    scene, o = pr2
    lo, hi, cont = o.joint_limits()
    q = scenes.random_states(200, lo, hi, cont, seed=6)
    # zero motion => 0 waypoints => valid without any check (collision_space.cpp:538-581)
    e, c = o.is_edges_valid(q, q)
    assert (c == 0).all() and (e == 1).all()
    # an edge is valid iff every waypoint is valid, endpoints included
    q0, q1 = scenes.mprim_edges(q)
    e, c = o.is_edges_valid(q0, q1)
    for i in range(50):
        w = o.edge_waypoints(q0[i], q1[i])
        assert len(w) == c[i] >= 2
        assert np.array_equal(w[0], q0[i])
        assert bool(e[i]) == bool(o.is_states_valid(w).all())
    # continuous joints interpolate along the shortest arc
    a = q[0].copy()
    b = q[0].copy()
    a[4], b[4] = 3.0, -3.0
    w = o.edge_waypoints(a, b)
    assert np.all(np.abs(np.diff(w[:, 4])) < 0.2) and w[-1, 4] > 3.0


@pytest.mark.parametrize("maker", [scenes.pr2_clutter_scene, scenes.pr2_tabletop_scene, scenes.ubr1_tabletop_scene,
                                   scenes.pr2_dual_arm_scene])
def test_product_tables_equal_oracle_tables(maker):
    """Host-side model builder (smpl_b200/host) vs the oracle: sphere trees, motion weights, tree pairs,
    limits, and the set of occupied cells contributed by out-of-group links -- all bit-identical."""
    scene = maker()
    if scene.chain_root is None:
        o = make_oracle(scene, with_kdl=False)
    else:
        o = make_oracle(scene)
    t = api.build_tables(scene)
    no, npd = o.node_table(), t.node_table()
    assert no.shape == npd.shape
    assert np.array_equal(no[:, :7], npd[:, :7])
    wo, to = o.motion_weights()
    wp, tp = t.motion_weights()
    assert np.array_equal(wo, wp)
    # oracle JointType (REVOLUTE=1, PRISMATIC=2, CONTINUOUS=3) vs SMPLGPU_VAR_* (0,2,1)
    assert np.array_equal(np.array([{1: 0, 2: 2, 3: 1}[int(x)] for x in to]), tp)
    assert np.array_equal(o.checked_pairs(), t.pairs())
    if scene.chain_root is not None:
        lo, hi, c = o.joint_limits()
        lo2, hi2, c2 = t.limits()
        assert np.array_equal(lo, lo2) and np.array_equal(hi, hi2) and np.array_equal(c, c2)
    occ_o = set(map(tuple, np.argwhere(o.df_d2() == 0)))
    occ_p = set(map(tuple, np.unique(api.scene_cells(scene, t), axis=0)))
    assert occ_o == occ_p


def test_distance_field_removal_restores_field():
    """add/remove symmetry -- the behaviour the reference's distance_map_test.cpp prints (:83-160)."""
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    before = o.df_d2().copy()
    rng = np.random.default_rng(3)
    pts = np.array(scene.origin) + 0.05 + rng.random((400, 3)) * (np.array(scene.size) - 0.1)
    o.add_points(pts)
    assert not np.array_equal(o.df_d2(), before)
    # only remove points whose cell was free before (the grid is not reference counted)
    g = o.world_to_grid(pts)
    free = before[g[:, 0], g[:, 1], g[:, 2]] > 0
    o.remove_points(pts[free])
    assert np.array_equal(o.df_d2(), before)


def test_world_to_grid_plumbing_matches_oracle():
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    rng = np.random.default_rng(4)
    pts = np.array(scene.origin) - 0.2 + rng.random((20000, 3)) * (np.array(scene.size) + 0.4)
    assert np.array_equal(o.world_to_grid(pts), api.world_to_grid(pts, scene.origin, scene.res))

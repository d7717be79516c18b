"""GPU parity tests of the validity path (kernels 1-3) through the C ABI, against the oracle.

Bar (BASELINE.json north_star): verdicts bit-exact except states that have a sphere within 1e-5 m
of a decision threshold (a worldToGrid cell boundary or a sphere-pair contact); those are counted.
"""
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import flips_within_tolerance, make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def pr2():
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    yield scene, o, ctx, tables
    ctx.close()


def test_device_distance_field_equals_oracle_field(pr2):
    scene, o, ctx, _ = pr2
    assert np.array_equal(ctx.download_distance_field().astype(np.int32), o.df_d2())


def test_distance_field_upload_path(pr2):
    scene, o, ctx, tables = pr2
    ctx2 = api.GpuContext(0)
    ctx2.set_robot(tables)
    _, origin, res, dmax_sq = o.grid_info()
    ctx2.set_distance_field(o.df_d2().astype(np.uint16), origin, res, dmax_sq)
    lo, hi, cont = tables.limits()
    q = scenes.random_states(2000, lo, hi, cont, seed=21)
    assert np.array_equal(ctx2.is_states_valid(q), ctx.is_states_valid(q))
    ctx2.close()


def test_fk_sphere_centres_match_oracle(pr2):
    """Kernel (1).  CUDA sin/cos may differ from glibc in the last place, so centres are compared to 1e-13 m
    and the fraction of bit-identical coordinates is reported."""
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    q = scenes.random_states(512, lo, hi, cont, seed=22)
    g = ctx.fk_sphere_centers(q)
    c = o.sphere_centers(q)
    assert g.shape == c.shape == (512, 38, 3)
    assert np.abs(g - c).max() < 1e-13
    print("bit-identical centre coordinates: %.4f" % float((g == c).mean()))
    assert (g == c).mean() > 0.5


def test_golden_vectors_through_the_abi(pr2):
    scene, o, ctx, _ = pr2
    g = np.loadtxt(os.path.join(GOLD, "pr2_right_arm_validity.csv"))
    assert np.array_equal(ctx.is_states_valid(g[:, :7]), g[:, 7].astype(np.uint8))
    e = np.loadtxt(os.path.join(GOLD, "pr2_right_arm_edges.csv"))
    v, c = ctx.is_edges_valid(e[:, :7], e[:, 7:14])
    assert np.array_equal(v, e[:, 14].astype(np.uint8))
    assert np.array_equal(c, e[:, 15].astype(np.int32))


def test_states_parity_with_flip_accounting(pr2):
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    q = scenes.random_states(60000, lo, hi, cont, seed=23)
    gpu = ctx.is_states_valid(q)
    cpu, L, cm, pm = o.report_states(q)
    flips, unexplained = flips_within_tolerance(gpu, cpu, cm, pm)
    print("states: %d flips of %d (all within 1e-5 m of a threshold: %s), mean L=%.2f, valid=%.3f" % (
        flips, len(q), unexplained == 0, L.mean(), cpu.mean()))
    assert unexplained == 0
    assert flips <= 2
    st = ctx.last_validity_stats()
    assert st["waypoints"] == len(q)
    assert len(q) <= st["df_lookups"] <= int(L.sum())  # early-out never does more than the exhaustive count


def test_edges_parity(pr2):
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    q = scenes.random_states(20000, lo, hi, cont, seed=24)
    q0, q1 = scenes.mprim_edges(q)
    gv, gc = ctx.is_edges_valid(q0, q1)
    cv, cc = o.is_edges_valid(q0, q1)
    assert np.array_equal(gc, cc)
    assert np.array_equal(gv, cv)
    # long edges (> 5 waypoints: the strided order of collision_space.cpp:561-570) and wrap-around edges
    big = q.copy()
    big[:, 0] = np.clip(big[:, 0] + 0.6, lo[0], hi[0])
    big[:, 4] = -q[:, 4]
    gv, gc = ctx.is_edges_valid(q, big)
    cv, cc = o.is_edges_valid(q, big)
    assert gc.max() > 10
    assert np.array_equal(gc, cc) and np.array_equal(gv, cv)


def test_edges_given_as_parent_plus_motion_primitive(pr2):
    """smplgpu_is_mprim_edges_valid forms q1 = q0 + delta[prim] on the device: same verdicts and counts as
    shipping q1, and as the oracle."""
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    q = scenes.random_states(300000, lo, hi, cont, seed=29)     # several chunks of the host pipeline
    deltas = scenes.pr2_mprim_deltas()
    pid = (np.arange(len(q)) % len(deltas)).astype(np.int32)
    pid[::1000] = -1                                            # out-of-table id: zero-length edge
    q1 = q + np.where(pid[:, None] >= 0, deltas[np.maximum(pid, 0)], 0.0)
    v_a, c_a = ctx.is_mprim_edges_valid(q, pid, deltas)
    v_b, c_b = ctx.is_edges_valid(q, q1)
    assert np.array_equal(v_a, v_b) and np.array_equal(c_a, c_b)
    assert (c_a[::1000] == 0).all() and (v_a[::1000] == 1).all()
    v_o, c_o = o.is_edges_valid(q[:4000], q1[:4000])
    assert np.array_equal(v_a[:4000], v_o) and np.array_equal(c_a[:4000], c_o)


def test_lattice_states_and_edges_as_16_bit_coordinates(pr2):
    """smplgpu_is_lattice_states_valid / _edges_valid take RobotCoord (16-bit) + a primitive byte and form the joint
    values on the device with ManipLattice::coordToState (manip_lattice.cpp:1245-1261): same verdicts and waypoint
    counts as the double entry points on coordToState(coords), and as the oracle."""
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    res = scenes.PlanParams(scene.dof).resolutions
    vals = ctx.set_lattice(res)
    want_vals, _, _ = scenes.lattice_discretisation(lo, hi, cont, res)
    assert np.array_equal(vals, want_vals)
    coords, q = scenes.random_lattice_coords(300000, lo, hi, cont, res, seed=33)   # several pipeline chunks
    deltas = scenes.pr2_mprim_deltas()
    pid = (np.arange(len(q)) % len(deltas)).astype(np.uint8)
    pid[::1000] = 255                                           # out-of-table id: zero-length edge
    q1 = q + np.where(pid[:, None] < len(deltas), deltas[np.minimum(pid, len(deltas) - 1)], 0.0)
    v_l = ctx.is_lattice_states_valid(coords)
    assert np.array_equal(v_l, ctx.is_states_valid(q))
    e_l, c_l = ctx.is_lattice_edges_valid(coords, pid, deltas)
    e_d, c_d = ctx.is_edges_valid(q, q1)
    assert np.array_equal(e_l, e_d) and np.array_equal(c_l, c_d)
    assert (c_l[::1000] == 0).all() and (e_l[::1000] == 1).all()
    assert np.array_equal(v_l[:4000], o.is_states_valid(q[:4000]))
    e_o, c_o = o.is_edges_valid(q[:4000], q1[:4000])
    assert np.array_equal(e_l[:4000], e_o) and np.array_equal(c_l[:4000], c_o)
    assert 0.05 < v_l.mean() < 0.95
    # empty batch, and a context without a lattice
    assert len(ctx.is_lattice_states_valid(np.zeros((0, scene.dof), np.int16))) == 0


def test_edge_cases(pr2):
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    q = scenes.random_states(257, lo, hi, cont, seed=25)
    # empty batch
    assert len(ctx.is_states_valid(np.zeros((0, 7)))) == 0
    # ragged sizes around the block size
    for n in (1, 31, 127, 128, 129, 257):
        assert np.array_equal(ctx.is_states_valid(q[:n]), o.is_states_valid(q[:n]))
    # zero-motion edges: 0 waypoints, valid without a check
    v, c = ctx.is_edges_valid(q, q)
    assert (c == 0).all() and (v == 1).all()
    # states far outside the grid: every lookup is out of bounds => d2 = 0 => invalid
    far = q.copy()
    far[:, 0] = 0.0
    ctx_v = ctx.is_states_valid(far)
    assert np.array_equal(ctx_v, o.is_states_valid(far))
    # values beyond +-2*pi on continuous joints (fmod path of normalize_angle)
    wrap = q.copy()
    wrap[:, 6] += 40.0
    v, c = ctx.is_edges_valid(q, wrap)
    cv, cc = o.is_edges_valid(q, wrap)
    assert np.array_equal(c, cc) and np.array_equal(v, cv)


def test_large_batch_chunked_host_path(pr2):
    """> 2^18 states exercises the double-buffered pinned staging of the host-pointer entry points."""
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    n = (1 << 18) * 2 + 12345
    q = scenes.random_states(n, lo, hi, cont, seed=26)
    v = ctx.is_states_valid(q)
    idx = np.random.default_rng(0).choice(n, 6000, replace=False)
    idx = np.concatenate([idx, [0, (1 << 18) - 1, 1 << 18, (1 << 19) - 1, 1 << 19, n - 1]])
    assert np.array_equal(v[idx], o.is_states_valid(q[idx]))
    # idempotence: same input, same output
    assert np.array_equal(v, ctx.is_states_valid(q))


def test_joint_limits_and_heuristic(pr2):
    scene, o, ctx, tables = pr2
    lo, hi, cont = tables.limits()
    rng = np.random.default_rng(27)
    q = scenes.random_states(5000, lo, hi, cont, seed=27)
    q += rng.normal(0.0, 0.3, q.shape) * (rng.random(q.shape) < 0.2)
    q[::7, 4] += 12.0
    assert np.array_equal(ctx.check_joint_limits(q), o.check_joint_limits(q))
    # planning-link FK
    p_gpu, p_cpu = ctx.planning_frame_fk(q), o.planning_frame_fk(q)
    assert np.abs(p_gpu[:, :3] - p_cpu[:, :3]).max() < 1e-13
    # BFS heuristic through the whole chain: walls from the field, goal from the demo, gather
    walls_gpu = ctx.bfs_set_walls_from_df(scene.inflation_radius)
    assert walls_gpu == o.heur_init(scene.inflation_radius, scene.cost_per_cell)
    goal = (0.4, -0.2, 0.8)
    gcell = o.world_to_grid([goal])[0]
    o.heur_set_goal(*goal)
    ctx.bfs_run([gcell])
    assert np.array_equal(ctx.bfs_download(), o.heur_grid())
    h_gpu, h_cpu = ctx.goal_heuristics(q, scene.cost_per_cell), o.goal_heuristics(q)
    flips = np.flatnonzero(h_gpu != h_cpu)
    print("heuristic mismatches: %d of %d" % (len(flips), len(q)))
    assert len(flips) == 0
    assert (h_cpu == 32767).any() and (h_cpu < 32767).any()


def test_padding_and_acm_variants():
    scene = scenes.pr2_clutter_scene()
    scene.padding = 0.015
    scene.use_desc_acm = False        # default ACM: adjacent links only => many more tree pairs
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    assert len(tables.pairs()) > 10
    lo, hi, cont = tables.limits()
    q = scenes.random_states(20000, lo, hi, cont, seed=28)
    gpu = ctx.is_states_valid(q)
    cpu, L, cm, pm = o.report_states(q)
    flips, unexplained = flips_within_tolerance(gpu, cpu, cm, pm)
    assert unexplained == 0 and flips <= 1
    assert ctx.last_validity_stats()["pair_tests"] > 0
    ctx.close()


def test_ubr1_attached_body():
    scene = scenes.ubr1_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    assert np.array_equal(ctx.download_distance_field().astype(np.int32), o.df_d2())
    g = np.loadtxt(os.path.join(GOLD, "ubr1_attached_validity.csv"))
    assert np.array_equal(ctx.is_states_valid(g[:, :7]), g[:, 7].astype(np.uint8))
    lo, hi, cont = tables.limits()
    q = scenes.random_states(20000, lo, hi, cont, seed=29)
    gpu = ctx.is_states_valid(q)
    cpu, L, cm, pm = o.report_states(q)
    flips, unexplained = flips_within_tolerance(gpu, cpu, cm, pm)
    assert unexplained == 0 and flips <= 1
    q0, q1 = scenes.mprim_edges(q[:5000])
    gv, gc = ctx.is_edges_valid(q0, q1)
    cv, cc = o.is_edges_valid(q0, q1)
    assert np.array_equal(gc, cc) and np.array_equal(gv, cv)
    c_gpu, c_cpu = ctx.fk_sphere_centers(q[:64]), o.sphere_centers(q[:64])
    assert c_gpu.shape == c_cpu.shape and np.abs(c_gpu - c_cpu).max() < 1e-13
    ctx.close()


def test_pr2_dual_arm_15dof():
    scene = scenes.pr2_dual_arm_scene()
    o = make_oracle(scene, with_kdl=False)
    ctx, tables = api.setup_context(scene)
    assert np.array_equal(ctx.download_distance_field().astype(np.int32), o.df_d2())
    lo, hi, cont = tables.limits()
    q = scenes.random_states(8000, lo, hi, cont, seed=30)
    gpu = ctx.is_states_valid(q)
    cpu, L, cm, pm = o.report_states(q)
    flips, unexplained = flips_within_tolerance(gpu, cpu, cm, pm)
    print("15-DOF: valid=%.3f mean L=%.1f flips=%d" % (cpu.mean(), L.mean(), flips))
    assert unexplained == 0 and flips <= 1
    d = np.zeros((22, 15))
    for k in range(15):
        d[k % 22, k] = 0.07
    q0 = q[:4000]
    q1 = q0 + d[np.arange(4000) % 22]
    gv, gc = ctx.is_edges_valid(q0, q1)
    cv, cc = o.is_edges_valid(q0, q1)
    assert np.array_equal(gc, cc) and np.array_equal(gv, cv)
    ctx.close()


def test_missing_scene_is_an_error_not_a_fallback():
    ctx = api.GpuContext(0)
    with pytest.raises(api.SmplGpuError):
        ctx.dof = 7
        ctx.is_states_valid(np.zeros((4, 7)))
    ctx.close()


def test_df_lookup_rate_probe():
    """bench.py's L2 roofline denominator: needs a scene, reports a plausible rate."""
    scene = scenes.pr2_clutter_scene()
    ctx, _ = api.setup_context(scene)
    try:
        assert 1e10 < ctx.probe_df_lookup_rate() < 1e13
    finally:
        ctx.close()
    bare = api.GpuContext(0)
    with pytest.raises(api.SmplGpuError):
        bare.probe_df_lookup_rate()
    bare.close()

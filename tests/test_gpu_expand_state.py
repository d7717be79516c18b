"""smplgpu_expand_state (one launch per expansion for UNCHANGED callers) and the ExpansionCache behind the adapters:
every field of a record must equal what the per-call entry points -- and hence the oracle -- answer for the same
state, and the adapters must give the same answers with the cache as without it while launching once per parent."""
import numpy as np
import pytest

from helpers import make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu


def table_with_converses(params):
    """ManipLatticeActionSpace::addMotionPrim with add_converse: the converse follows each primitive."""
    rows = []
    for d in np.asarray(params.mprims):
        rows.append(d)
        rows.append(-d)
    return np.array(rows)


@pytest.fixture(scope="module")
def rig():
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    goal = (0.5, -0.3, 0.8)
    o.heur_init(scene.inflation_radius, scene.cost_per_cell)
    o.heur_set_goal(*goal)
    ctx.bfs_set_walls_from_df(scene.inflation_radius)
    ctx.bfs_run([api.world_to_grid([goal], scene.origin, scene.res)[0]])
    yield scene, o, ctx, tables
    ctx.close()


def test_records_equal_the_per_call_entry_points_and_the_oracle(rig):
    scene, o, ctx, tables = rig
    deltas = table_with_converses(scenes.PlanParams(scene.dof))
    ctx.set_motion_primitives(deltas)
    lo, hi, cont = tables.limits()
    parents = scenes.random_states(60, lo, hi, cont, seed=71)
    parents[::7, 1] = hi[1] - 0.03          # successors beyond a joint limit
    parents[3] = np.zeros(scene.dof)
    n_invalid_edges = 0
    for q in parents:
        l0 = ctx.launch_count()
        rec = ctx.expand_state(q, scene.cost_per_cell)
        assert ctx.launch_count() - l0 == 1            # ONE launch answers the whole expansion
        succ = q[None, :] + deltas                       # the host's addition == the device's
        states = np.vstack([q[None, :], succ])
        assert np.array_equal(rec["state"], states)
        assert list(rec["is_parent"]) == [1] + [0] * len(deltas)
        # edges parent -> successor: verdict and waypoint count
        e_cpu, c_cpu = o.is_edges_valid(np.repeat(q[None, :], len(deltas), 0), succ)
        assert np.array_equal(rec["edge_valid"][1:], e_cpu)
        assert np.array_equal(rec["waypoints"][1:], c_cpu)
        n_invalid_edges += int((e_cpu == 0).sum())
        # the parent's own validity
        assert rec["state_valid"][0] == o.is_states_valid(q[None, :])[0]
        # joint limits, planning-frame FK, heuristic of every record
        assert np.array_equal(rec["limits_ok"], o.check_joint_limits(states))
        assert np.array_equal(rec["pose"], ctx.planning_frame_fk(states))     # same device code: bit for bit
        assert np.abs(rec["pose"] - o.planning_frame_fk(states)).max() < 1e-12
        assert np.array_equal(rec["h"], o.goal_heuristics(states))
        # metric goal distance in cells at the planning link (getMetricGoalDistance / res)
        cells = api.world_to_grid(rec["link_xyz"], scene.origin, scene.res)
        assert np.array_equal(rec["goal_dist_cells"], ctx.bfs_distances(cells))
    assert n_invalid_edges > 0


def test_adapters_answer_the_same_with_and_without_the_cache(rig):
    """The reference's call pattern for one expansion (manip_lattice_action_space.cpp:385-396, manip_lattice.cpp:1520-1565,
    1582-1640, arastar.cpp:613-618), one virtual at a time, through the adapters."""
    scene, o, ctx, tables = rig
    deltas = table_with_converses(scenes.PlanParams(scene.dof))
    lo, hi, cont = tables.limits()
    parents = scenes.random_states(25, lo, hi, cont, seed=72)
    goal = (0.5, -0.3, 0.8)

    def walk(ad):
        out = []
        for q in parents:
            pose = ad.compute_planning_link_fk(q)
            out.append(tuple(pose))
            out.append(ad.metric_goal_distance(pose[0], pose[1], pose[2]))
            for d in deltas:
                s = d + q
                ok = ad.check_joint_limits(s)
                out.append(ok)
                if not ok:
                    continue
                v = ad.is_state_to_state_valid(q, s)
                out.append(v)
                if v:
                    out.append(tuple(ad.compute_planning_link_fk(s)))
                    out.append(ad.goal_heuristic(s))
            out.append(ad.is_state_valid(q))
        return out

    plain = api.Adapters(ctx, scene, tables)
    plain.update_goal(goal)
    l0 = ctx.launch_count()
    want = walk(plain)
    launches_plain = ctx.launch_count() - l0
    plain.close()

    cached = api.Adapters(ctx, scene, tables)
    cached.update_goal(goal)
    cached.enable_expansion_cache(deltas)
    l0 = ctx.launch_count()
    got = walk(cached)
    launches_cached = ctx.launch_count() - l0
    n_launch, n_hit = cached.expansion_cache_counters()
    # arbitrary edges (not parent + primitive) still take the per-call path, with the same answer
    a, b = parents[0], parents[1]
    assert cached.is_state_to_state_valid(a, b) == bool(o.is_edges_valid(a[None, :], b[None, :])[0][0])
    # a new goal invalidates the record: the heuristic follows the new BFS
    cached.update_goal((0.6, 0.1, 0.9))
    o.heur_set_goal(0.6, 0.1, 0.9)
    assert cached.goal_heuristic(parents[0]) == o.goal_heuristics(parents[:1])[0]
    o.heur_set_goal(*goal)
    cached.close()
    assert got == want
    assert launches_cached == n_launch == len(parents)          # one launch per expanded state
    assert n_hit > 20 * len(parents)
    assert launches_plain > 30 * launches_cached

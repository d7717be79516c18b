"""The drop-in boundary as the reference's planner sees it: the C++ adapter objects (GpuCollisionSpace :
CollisionChecker, GpuRobotModel : ForwardKinematicsInterface, GpuBfsHeuristic : RobotHeuristic) called one
virtual at a time must answer exactly like the oracle's CollisionSpace / KDLRobotModel / BfsHeuristic."""
import numpy as np
import pytest

from helpers import make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rig():
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    ad = api.Adapters(ctx, scene, tables)
    yield scene, o, ctx, tables, ad
    ad.close()
    ctx.close()


def test_collision_checker_virtuals(rig):
    scene, o, ctx, tables, ad = rig
    lo, hi, cont = tables.limits()
    q = scenes.random_states(300, lo, hi, cont, seed=51)
    q0, q1 = scenes.mprim_edges(q)
    v = o.is_states_valid(q)
    e, counts = o.is_edges_valid(q0, q1)
    assert [ad.is_state_valid(x) for x in q] == list(v)
    assert [ad.is_state_to_state_valid(a, b) for a, b in zip(q0, q1)] == list(e)
    assert v.any() and not v.all()
    # the batched entry points behind the same object
    assert np.array_equal(ad.is_states_valid(q), v)
    assert np.array_equal(ad.is_edges_valid(q0, q1), e)
    # interpolatePath returns the waypoints isStateToStateValid checks
    for k in (0, 7, 19):
        wp = ad.interpolate_path(q0[k], q1[k])
        ref = o.edge_waypoints(q0[k], q1[k])
        assert len(wp) == counts[k] == len(ref)
        assert np.array_equal(wp, ref)
    # wrong-sized state is `false`, not an exception
    assert ad.interpolate_path(q0[0], q0[0]).shape == (0, 7)


def test_robot_model_virtuals(rig):
    scene, o, ctx, tables, ad = rig
    lo, hi, cont = tables.limits()
    q = scenes.random_states(200, lo, hi, cont, seed=52)
    q[::5, 1] += 2.0
    assert [ad.check_joint_limits(x) for x in q] == list(o.check_joint_limits(q))
    pose = np.array([ad.compute_planning_link_fk(x) for x in q])
    assert np.abs(pose - o.planning_frame_fk(q)).max() < 1e-12


def test_heuristic_virtuals(rig):
    scene, o, ctx, tables, ad = rig
    lo, hi, cont = tables.limits()
    q = scenes.random_states(200, lo, hi, cont, seed=53)
    goal = (0.5, -0.3, 0.8)
    o.heur_init(scene.inflation_radius, scene.cost_per_cell)
    o.heur_set_goal(*goal)
    ad.update_goal(goal)
    assert [ad.goal_heuristic(x) for x in q] == list(o.goal_heuristics(q))
    grid = o.heur_grid()
    cell = o.world_to_grid([goal])[0]
    assert ad.metric_goal_distance(*goal) == 0.0
    far = (0.9, 0.4, 1.2)
    c = o.world_to_grid([far])[0]
    assert ad.metric_goal_distance(*far) == float(grid[c[2] + 1, c[1] + 1, c[0] + 1]) * scene.res
    assert ad.metric_goal_distance(50.0, 0.0, 0.0) == float(0x7FFFFFFF) * scene.res
    assert cell.min() >= 0

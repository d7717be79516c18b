"""Pins the query level -- oracle/lattice.cpp (ManipLattice bookkeeping), the BfsHeuristic of oracle/kdl_model.cpp and
oracle/arastar.h working together -- against the REFERENCE's own planning stack: ManipLattice, RobotPlanningSpace,
BfsHeuristic + BFS_3D, ARAStar and CollisionSpace compiled where they lie (oracle/ref_planner_shim.cpp, part of
oracle/_ref/libref_collision.so).  The RobotModel plug-in (KDL is absent) and the action space (fork defect 2) are the
shim's, as its header explains.  Same success flag, expansion count, cost, lattice size, state-id path and extracted
joint path.

tests/golden/plans_reference.json holds the reference build's results (tools/gen_golden_plans_reference.py); the
restatement is checked against it where oracle/_ref is absent, and tests/golden/pr2_tabletop_plans.json -- the plans the
GPU batch planner has to reproduce (tests/test_gpu_planner.py) -- is checked to be what the reference build returns.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle
from oracle_api import ref_collision_lib
from smpl_b200 import scenes
from test_oracle_collision import UBR1_BOX, make_reference, make_restatement

GOLD = os.path.join(ROOT, "tests", "golden")
needs_ref = pytest.mark.skipif(ref_collision_lib() is None, reason="oracle/_ref/libref_collision.so not built")


def plan_cases():
    """name -> (scene, attach, params, starts, goals)"""
    pr2 = scenes.pr2_tabletop_scene()
    p1 = scenes.PlanParams(pr2.dof)
    p1.max_expansions = 2000
    s1, g1 = scenes.tabletop_queries(8, seed=3)
    ubr1 = scenes.ubr1_tabletop_scene()          # config 4 shape: extra ACM entries, grasped object
    ubr1.attached = None                          # ... attached through attachObject / attach_box on both sides
    attach = ("object", "wrist_roll_link") + UBR1_BOX
    p2 = scenes.PlanParams(ubr1.dof)
    p2.max_expansions = 800
    s2, g2 = scenes.ubr1_tabletop_queries(6, seed=31)
    return {"pr2_tabletop": (pr2, None, p1, s1, g1), "ubr1_tabletop": (ubr1, attach, p2, s2, g2)}


def summary(p):
    return [int(p["success"]), int(p["expansions"]), int(p["cost"]), int(p["num_states"]), [int(i) for i in p["path_ids"]]]


@needs_ref
@pytest.mark.parametrize("name", ["pr2_tabletop", "ubr1_tabletop"])
def test_restatement_plans_equal_reference_build(name):
    scene, attach, params, starts, goals = plan_cases()[name]
    o = make_restatement(scene, attach, with_kdl=True)
    r = make_reference(scene, attach)
    solved = 0
    for s, g in zip(starts, goals):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        a = o.plan(s, g, params)
        b = r.plan(scene, s, g, params)
        assert summary(a) == summary(b)
        assert np.array_equal(a["path_states"], b["path_states"])
        solved += a["success"]
    assert solved >= 2


@needs_ref
def test_committed_golden_plans_are_the_reference_builds():
    gold = json.load(open(os.path.join(GOLD, "pr2_tabletop_plans.json")))
    scene = scenes.pr2_tabletop_scene()
    r = make_reference(scene, None)
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = gold["max_expansions"]
    for s, g, res in zip(gold["starts"], gold["goals"], gold["results"]):
        assert summary(r.plan(scene, np.array(s), np.array(g), params)) == res[:5]


def weighted_params(params):
    """the primitive file's weight column (manip_lattice_action_space.cpp:182-190): edge cost = int(1000 * weight)"""
    import copy
    p = copy.copy(params)
    p.weights = [(1.0, 2.5, 0.4, 1.7)[i % 4] for i in range(len(params.mprims))]
    return p


@needs_ref
def test_restatement_plans_with_action_weights_equal_reference_build():
    """cost(parent, succ, actionWeight, goal) = DefaultCostMultiplier * actionWeight, truncated to int
    (manip_lattice.cpp:296, 1414-1437): non-unit weights change which path ARA* returns; restatement and reference
    build must agree, and differ from the unit-weight plans."""
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    wp = weighted_params(params)
    o = make_restatement(scene, attach, with_kdl=True)
    r = make_reference(scene, attach)
    changed = 0
    for s, g in list(zip(starts, goals))[:5]:
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        a = o.plan(s, g, wp)
        b = r.plan(scene, s, g, wp)
        assert summary(a) == summary(b)
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        changed += summary(o.plan(s, g, params)) != summary(a)
    assert changed >= 2


@needs_ref
@pytest.mark.parametrize("name", ["pr2_tabletop", "ubr1_tabletop"])
def test_lazy_restatement_plans_equal_reference_build(name):
    """oracle/lazy_arastar.h + ManipLatticePlanner::getLazySuccs / getTrueCost against the reference's own
    lazy_arastar.cpp + ManipLattice::GetLazySuccs / GetTrueCost (compiled from its sources), query by query: success,
    expansions, evaluations, cost, lattice size, id path, joint path."""
    scene, attach, params, starts, goals = plan_cases()[name]
    o = make_restatement(scene, attach, with_kdl=True)
    r = make_reference(scene, attach)
    solved = 0
    for s, g in zip(starts, goals):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        a = o.plan(s, g, params, lazy=True)
        b = r.plan(scene, s, g, params, lazy=True)
        assert summary(a) + [a["evaluations"]] == summary(b) + [b["evaluations"]]
        assert np.array_equal(a["path_states"], b["path_states"])
        solved += a["success"]
    assert solved >= 2


def test_lazy_restatement_plans_equal_golden_reference_outputs():
    gold = json.load(open(os.path.join(GOLD, "plans_reference_lazy.json")))
    for name, (scene, attach, params, starts, goals) in plan_cases().items():
        o = make_restatement(scene, attach, with_kdl=True)
        for s, g, want in list(zip(starts, goals, gold[name]))[:5]:
            o.heur_init(scene.inflation_radius, scene.cost_per_cell)
            p = o.plan(s, g, params, lazy=True)
            assert summary(p) + [p["evaluations"]] == want


@needs_ref
def test_committed_lazy_golden_plans_are_the_reference_builds():
    """The reference's lazy successors -- ManipLattice::GetLazySuccs / GetTrueCost (manip_lattice.cpp:1012-1167) under its
    in-tree LazyARAStar (search/lazy_arastar.cpp), compiled from the reference's sources -- reproduce the committed
    fixture (tools/gen_golden_plans_reference.py).  The GPU side of this row is tests/test_gpu_dropin.py: the same
    reference code with the product's adapters answering."""
    gold = json.load(open(os.path.join(GOLD, "plans_reference_lazy.json")))
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    r = make_reference(scene, attach)
    lazier = 0
    for s, g, want in list(zip(starts, goals, gold["pr2_tabletop"]))[:4]:
        p = r.plan(scene, s, g, params, lazy=True)
        assert summary(p) + [p["evaluations"]] == want
        # laziness: far fewer edges evaluated than an eager expansion generates (~65 per expansion)
        lazier += p["evaluations"] < 4 * max(1, p["expansions"])
    assert lazier == 4


def test_restatement_plans_equal_golden_reference_outputs():
    gold = json.load(open(os.path.join(GOLD, "plans_reference.json")))
    for name, (scene, attach, params, starts, goals) in plan_cases().items():
        o = make_restatement(scene, attach, with_kdl=True)
        for s, g, res in zip(starts, goals, gold[name]):
            o.heur_init(scene.inflation_radius, scene.cost_per_cell)
            assert summary(o.plan(s, g, params)) == res


def _wandering_paths(anchors, n_paths, dof, seed):
    """random walks, jittered lines and detours between valid states (20 to 60 points)"""
    rng = np.random.default_rng(seed)
    paths = []
    for p in range(n_paths):
        m = int(rng.integers(20, 61))
        a, b = anchors[rng.integers(0, len(anchors), 2)]
        if p % 3 == 0:
            pts = a + np.cumsum(rng.normal(0.0, 0.06, (m, dof)), axis=0)
        elif p % 3 == 1:
            pts = a + np.linspace(0.0, 1.0, m)[:, None] * (b - a) * 0.5 + rng.normal(0.0, 0.02, (m, dof))
        else:
            t = np.concatenate([np.linspace(0, 1, m // 2 + 1), np.linspace(1, 0.2, m - m // 2 - 1)])[:, None]
            pts = a + t * (b - a) * 0.4
        paths.append(np.ascontiguousarray(pts[:m]))
    return paths


@needs_ref
@pytest.mark.parametrize("kind", [0, 1])
def test_shortcut_paths_equal_reference_post_processing(kind):
    """ShortcutPath(rm, cc, pin, pout, JOINT_SPACE / JOINT_POSITION_VELOCITY_SPACE) of the reference's
    own post_processing.cpp (generators, cost functions and shortcut templates, over its own CollisionSpace) against
    oracle/shortcut.h: the same points."""
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    r = make_reference(scene, None)
    lo, hi, cont = o.joint_limits()
    q = scenes.random_states(4000, lo, hi, cont, seed=8)
    anchors = q[o.is_states_valid(q) == 1]
    shortened = 0
    for p in _wandering_paths(anchors, 30, scene.dof, seed=4):
        idx, _ = o.shortcut_path(p, cont, kind=kind)
        ref = r.post_process(scene, p, kind)
        assert ref.shape == p[idx].shape and np.array_equal(ref, p[idx])
        shortened += len(idx) < len(p)
    assert shortened >= 20


@needs_ref
def test_reference_interpolate_path_rejects_legal_paths():
    """Fork defect 7 (collision_space.cpp:592-597), observed on the reference build itself: InterpolatePath fails on a
    path whose points are all within the joint limits, which is why oracle and product leave that test out."""
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    r = make_reference(scene, None)
    lo, hi, cont = o.joint_limits()
    q = scenes.random_states(2000, lo, hi, cont, seed=8)
    a = q[o.is_states_valid(q) == 1][0]
    path = np.stack([a, a + 0.05, a + 0.1])
    assert o.check_joint_limits(path).all()
    assert r.post_process(scene, path, 2) is None
    assert len(o.interpolate_path(path)) >= 3


@needs_ref
@pytest.mark.parametrize("robot", ["pr2", "ubr1"])
def test_kdl_robot_model_equals_reference_build(robot):
    """the reference's own kdl_robot_model.cpp (compiled against the KDL / kdl_parser
    stand-ins of oracle/ref_stubs/collision/kdl) against oracle/kdl_model.cpp: the limits the planner sees, the
    planning-link pose (x y z roll pitch yaw, bit for bit -- including the segment-index quirk of
    computePlanningLinkFK) and checkJointLimits for states inside, outside and several turns away from the limits."""
    scene = scenes.pr2_tabletop_scene() if robot == "pr2" else scenes.ubr1_tabletop_scene(attach=False)
    o = make_oracle(scene)
    r = make_reference(scene, None)
    lo, hi, cont = o.joint_limits()
    rlo, rhi, rcont = r.kdl_limits(scene)
    assert np.array_equal(lo, rlo) and np.array_equal(hi, rhi) and np.array_equal(cont, rcont)
    rng = np.random.default_rng(5)
    q = scenes.random_states(3000, lo, hi, cont, seed=12)
    q[::3] += rng.normal(0.0, 2.5, q[::3].shape)
    pose, ok = r.kdl_fk_and_limits(scene, q)
    assert np.array_equal(o.planning_frame_fk(q), pose)
    assert np.array_equal(o.check_joint_limits(q), ok)
    assert 0.2 < ok.mean() < 0.9


@needs_ref
def test_goal_heuristic_values_equal_reference_bfs_heuristic():
    """BfsHeuristic::GetGoalHeuristic of the reference, asked by lattice state id on the reference's
    own ManipLattice after ManipLattice::setGoal -> updateGoal -> BFS_3D::run, against the oracle's heuristic on the
    same joint states: cost_per_cell * distance, Infinity (32767) for planning-frame positions in walls or outside
    the grid, and cost_per_cell * (-1) -- negative -- for free cells the wavefront never reached (goal buried in the
    table).  (An out-of-bounds goal makes the reference's BFS_3D throw std::system_error, fork defect 6: not pinned.)"""
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    r = make_reference(scene, None)
    lo, hi, cont = o.joint_limits()
    q = scenes.random_states(4000, lo, hi, cont, seed=21)
    res = scenes.PlanParams(scene.dof).resolutions
    seen_negative = seen_infinity = False
    for goal in ((0.4, -0.2, 0.9), (0.6, 0.3, 0.75), (0.55, 0.0, 0.6)):     # two free cells, one inside the table top
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        o.heur_set_goal(*goal)
        h_ref, h_goal_state, _ = r.goal_heuristics(scene, goal, res, q)
        assert np.array_equal(o.goal_heuristics(q), h_ref)
        assert h_goal_state == 0
        seen_negative |= bool((h_ref < 0).any())
        seen_infinity |= bool((h_ref == 32767).any())
    assert seen_negative and seen_infinity

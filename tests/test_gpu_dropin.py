"""The drop-in claim, run literally (BASELINE north_star: "existing planners link unchanged"): the REFERENCE's own
ManipLattice + RobotPlanningSpace + ARAStar, compiled from /root/reference, plan queries with the product's plug-ins
behind the reference's own interfaces -- GpuCollisionSpace as CollisionChecker, GpuRobotModel as RobotModel +
ForwardKinematicsInterface, GpuBfsHeuristic as RobotHeuristic (smpl_b200/host/gpu_adapters.cpp compiled against the
reference's real headers) -- so every isStateToStateValid / GetGoalHeuristic / computePlanningLinkFK / checkJointLimits
call of the reference's search is answered by a CUDA kernel through the C ABI (oracle/ref_dropin_shim.cpp ->
oracle/_ref/libref_dropin.so).  The plans must equal the all-reference run's (tests/golden/plans_reference.json, written
by the reference's own CollisionSpace + BfsHeuristic + BFS_3D): success, expansions, cost, lattice size, id path."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from smpl_b200 import api
from test_oracle_planner_reference import plan_cases

pytestmark = pytest.mark.gpu

LIB = os.path.join(ROOT, "oracle", "_ref", "libref_dropin.so")


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def dropin_plan(lib, ctx, scene, start, goal, params, max_path=4096, batched=False, lazy=False):
    dof = scene.dof
    start = np.ascontiguousarray(start, np.float64)
    goal = np.ascontiguousarray(goal, np.float64)
    origin = np.ascontiguousarray(scene.origin, np.float64)
    dims = np.ascontiguousarray(scene.dims, np.int32)
    off = np.ascontiguousarray(scene.xyz_offset, np.float64)
    res = np.ascontiguousarray(params.resolutions, np.float64)
    prims = np.ascontiguousarray(params.mprims, np.float64)
    flags = np.ascontiguousarray(params.short_flags, np.uint8)
    tol = np.ascontiguousarray(params.xyz_tolerance, np.float64)
    summary = np.zeros(8, np.int32)
    path = np.zeros(max_path, np.int32)
    pstates = np.zeros((max_path, dof), np.float64)
    fn = lib.refdrop_plan_lazy if lazy else lib.refdrop_plan
    rc = fn(ctx.h, scene.robot_path.encode(), scene.group.encode(), ",".join(scene.planning_joints).encode(),
                          scene.planning_link.encode(), _p(origin, C.c_double), C.c_double(scene.res), _p(dims, C.c_int32),
                          C.c_double(scene.inflation_radius), int(scene.cost_per_cell),
                          _p(start, C.c_double), _p(goal, C.c_double), _p(off, C.c_double),
                          _p(res, C.c_double), _p(prims, C.c_double), _p(flags, C.c_uint8), len(prims),
                          int(params.use_short_dist), C.c_double(params.short_dist_thresh), C.c_double(params.epsilon),
                          int(params.max_expansions), _p(tol, C.c_double), _p(summary, C.c_int32), _p(path, C.c_int32),
                          max_path, _p(pstates, C.c_double), int(batched))
    assert rc == 0, "refdrop_plan refused step %d: %s" % (-rc, ctx.L.smplgpu_last_error(ctx.h))
    n = int(summary[3])
    dropin_plan.last_batched = (int(summary[6]), int(summary[7]))
    dropin_plan.last_evaluations = int(lib.refdrop_last_lazy_evaluations()) if lazy else 0
    return [int(summary[0]), int(summary[1]), int(summary[2]), int(summary[4]), [int(i) for i in path[:n]]]


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_dropin.so not built")
def test_reference_planner_over_gpu_plugins_returns_the_reference_plans():
    lib = C.CDLL(LIB)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plans_reference.json")))
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    assert attach is None
    ctx, tables = api.setup_context(scene)
    try:
        launches0 = ctx.launch_count()
        solved = 0
        for s, g, want in zip(starts, goals, gold["pr2_tabletop"]):
            got = dropin_plan(lib, ctx, scene, s, g, params)
            assert got == want
            solved += got[0]
        assert solved >= 4
        assert ctx.launch_count() - launches0 > 10000     # the reference's search really ran on the device
    finally:
        ctx.close()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_dropin.so not built")
def test_reference_lattice_with_batched_get_succs_returns_the_same_plans():
    """INTEGRATION.md section 4 on the reference's own lattice: GetSuccs collects the successors of an expansion and
    checks them in one GpuCollisionSpace::isEdgesValid call (BatchedManipLattice in oracle/ref_dropin_shim.cpp, every
    other line of the expansion is the reference's).  Same plans with fewer device calls: the edge checks shrink to
    one launch group per expansion; what remains are the reference's per-state calls (checkJointLimits and
    computePlanningLinkFK per successor, GetGoalHeuristic per new state), which only a caller that batches those too --
    smplgpu_expand_batch behind smplhost_plan_batch -- gets rid of."""
    import time
    lib = C.CDLL(LIB)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plans_reference.json")))
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    ctx, tables = api.setup_context(scene)
    try:
        secs = {}
        for batched in (False, True):
            t0 = time.perf_counter()
            l0 = ctx.launch_count()
            for s, g, want in zip(starts, goals, gold["pr2_tabletop"]):
                assert dropin_plan(lib, ctx, scene, s, g, params, batched=batched) == want
                if batched:
                    calls, edges = dropin_plan.last_batched
                    assert calls == want[1] or not want[0] or calls <= want[1]   # at most one call per expansion
                    assert edges >= calls
            secs[batched] = (time.perf_counter() - t0, ctx.launch_count() - l0)
        print("reference lattice over the GPU plug-ins, 8 queries: per-successor calls %.2f s (%d launches), "
              "batched GetSuccs %.2f s (%d launches)" % (secs[False] + secs[True]))
        assert secs[True][1] < 0.75 * secs[False][1]
    finally:
        ctx.close()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_dropin.so not built")
def test_unchanged_reference_planner_with_expansion_cache_is_one_launch_per_expansion():
    """The UNCHANGED reference lattice (ManipLattice::GetSuccs as shipped, one virtual call per question) with the
    adapters sharing a smplhost::ExpansionCache: the first question about a state triggers one smplgpu_expand_state
    launch, the other ~65 calls of the expansion are served from its record.  Same plans as the all-reference run; at
    most ~one launch per expansion (plus the goal / start set-up), and the wall time is printed beside the per-call
    path's (11.7 s for these eight queries in round 1; the reference's own CPU stack needs ~0.6 s)."""
    import time
    lib = C.CDLL(LIB)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plans_reference.json")))
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    ctx, tables = api.setup_context(scene)
    try:
        dropin_plan(lib, ctx, scene, starts[0], goals[0], params, batched=2)     # warm-up (allocations, module load)
        l0 = ctx.launch_count()
        t0 = time.perf_counter()
        expansions = cache_launches = 0
        for s, g, want in zip(starts, goals, gold["pr2_tabletop"]):
            got = dropin_plan(lib, ctx, scene, s, g, params, batched=2)
            assert got == want
            expansions += got[1]
            cache_launches += dropin_plan.last_batched[0]
        secs = time.perf_counter() - t0
        launches = ctx.launch_count() - l0
        print("unchanged reference planner over cached GPU plug-ins, 8 queries: %d expansions in %.3f s "
              "(%.0f expansions/s), %d launches (%d by the cache)" % (expansions, secs, expansions / secs, launches,
                                                                      cache_launches))
        # per query: walls + BFS (a handful of launches) + one record per expanded state (+ start, + goal-state h, + the
        # few states ARA* asks about out of expansion order: 6616 records for 6491 expansions on the B200)
        assert cache_launches <= 1.05 * expansions + 4 * len(starts)
        assert launches <= cache_launches + 40 * len(starts)
    finally:
        ctx.close()


@pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libref_dropin.so not built")
def test_reference_lazy_planner_over_gpu_plugins_returns_the_reference_plans():
    """SURVEY 8f row 2, second half: the reference's LAZY successors -- ManipLattice::GetLazySuccs / GetTrueCost
    (manip_lattice.cpp:1012-1167) under its in-tree LazyARAStar (search/lazy_arastar.cpp), all compiled from the
    reference -- with the product's adapters answering, first call by call, then with the shared ExpansionCache (a
    GetTrueCost on a parent expanded earlier re-runs that parent's record: one launch per evaluated edge group).  Plans,
    expansion and evaluation counts must equal the all-reference run's (tests/golden/plans_reference_lazy.json)."""
    import time
    lib = C.CDLL(LIB)
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "plans_reference_lazy.json")))
    scene, attach, params, starts, goals = plan_cases()["pr2_tabletop"]
    ctx, tables = api.setup_context(scene)
    try:
        for mode in (0, 2):
            dropin_plan(lib, ctx, scene, starts[1], goals[1], params, batched=mode, lazy=True)   # warm-up
            l0 = ctx.launch_count()
            t0 = time.perf_counter()
            expansions = evaluations = solved = 0
            for s, g, want in zip(starts, goals, gold["pr2_tabletop"]):
                got = dropin_plan(lib, ctx, scene, s, g, params, batched=mode, lazy=True)
                assert got + [dropin_plan.last_evaluations] == want
                expansions += got[1]
                evaluations += dropin_plan.last_evaluations
                solved += got[0]
            secs = time.perf_counter() - t0
            print("reference LazyARAStar over the GPU plug-ins (%s), 8 queries: %d expansions + %d edge evaluations in "
                  "%.3f s, %d launches" % ("expansion cache" if mode == 2 else "one call per question", expansions,
                                           evaluations, secs, ctx.launch_count() - l0))
            assert solved >= 4
    finally:
        ctx.close()

"""Query-level parity: the lock-step batch planner over the CUDA path must
return, for every query, the same path (lattice state ids), cost and expansion count as the sequential
reference-shaped planner of the oracle (ManipLattice + ARA*, first solution at the initial epsilon)."""
import json
import os

import numpy as np
import pytest

from conftest import ROOT
from helpers import make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def tabletop():
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    yield scene, o, ctx, tables
    ctx.close()


def test_bfs_bank_equals_single_runs(tabletop):
    scene, o, ctx, tables = tabletop
    walls = ctx.bfs_set_walls_from_df(scene.inflation_radius)
    assert ctx.bfs_bank_create(5, scene.inflation_radius) == walls
    _, goals = scenes.tabletop_queries(4, seed=31)
    seeds = api.world_to_grid(goals, scene.origin, scene.res)
    seeds = np.concatenate([seeds, [[-1, 0, 0]]]).astype(np.int32)   # slot 4: out-of-bounds seed => no search
    assert ctx.bfs_bank_run(seeds) == 4
    rng = np.random.default_rng(5)
    cells = rng.integers(0, np.asarray(scene.dims), (4000, 3)).astype(np.int32)
    for s in range(4):
        ctx.bfs_set_walls_from_df(scene.inflation_radius)   # fresh walls: a seed on a wall un-walls it (bfs3d.cpp:181-187)
        ctx.bfs_run([seeds[s]])
        single = ctx.bfs_distances(cells)
        bank = ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells)
        assert np.array_equal(single, bank)
        assert (single > 0).any()
    d = ctx.bfs_bank_distances(np.full(len(cells), 4, np.int32), cells)
    assert np.all((d == -1) | (d == 0x7FFFFFFF))


def test_asynchronous_bank_runs_equal_synchronous_ones(tabletop):
    """smplgpu_bfs_bank_run_slots_async: same distances as the blocking call, other slots untouched, one run in
    flight per context, and two contexts on one GPU take turns."""
    scene, o, ctx, tables = tabletop
    ctx.bfs_bank_create(4, scene.inflation_radius)
    _, goals = scenes.tabletop_queries(6, seed=41)
    seeds = api.world_to_grid(goals, scene.origin, scene.res).astype(np.int32)
    rng = np.random.default_rng(6)
    cells = rng.integers(0, np.asarray(scene.dims), (3000, 3)).astype(np.int32)
    ctx.bfs_bank_run(seeds[:4])
    want = [ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells) for s in range(4)]
    # re-run slots 1 and 3 with other goals, asynchronously
    ctx.bfs_bank_run_slots_async([1, 3], seeds[4:6])
    with pytest.raises(api.SmplGpuError):
        ctx.bfs_bank_run_slots_async([0], seeds[:1])
    ctx.bfs_bank_run_wait()
    assert ctx.bfs_bank_run_done()
    got = [ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells) for s in range(4)]
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[2], want[2])
    ctx.bfs_bank_run_slots([1, 3], seeds[4:6])
    for s in (1, 3):
        assert np.array_equal(got[s], ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells))
        assert not np.array_equal(got[s], want[s])
    # a second context on the same GPU: its run waits for its turn, both finish
    other = api.clone_context(ctx, scene, tables)
    try:
        other.bfs_bank_create(2, scene.inflation_radius)
        ctx.bfs_bank_run_slots_async([0, 1, 2, 3], seeds[:4])
        other.bfs_bank_run_slots_async([0, 1], seeds[:2])
        while not (ctx.bfs_bank_run_done() and other.bfs_bank_run_done()):
            pass
        for s in range(2):
            assert np.array_equal(other.bfs_bank_distances(np.full(len(cells), s, np.int32), cells), want[s])
        assert np.array_equal(ctx.bfs_bank_distances(np.full(len(cells), 3, np.int32), cells), want[3])
    finally:
        other.close()


def test_bank_max_slots_bounds_the_stacked_grid(tabletop):
    """The stacked bank keeps BFS_3D's int node indices (bfs3d.h:213-220): the slot limit the planner clamps to."""
    scene, o, ctx, tables = tabletop
    nx, ny, nz = scene.dims
    m = ctx.bfs_bank_max_slots()
    assert 1 <= m <= 0x7FFFFFFF // ((nx + 2) * (ny + 2) * (nz + 2))
    with pytest.raises(api.SmplGpuError):
        ctx.bfs_bank_create(0x7FFFFFFF // ((nx + 2) * (ny + 2) * (nz + 2)) + 1, scene.inflation_radius)


def test_bank_slot_refill_leaves_other_slots_untouched(tabletop):
    scene, o, ctx, tables = tabletop
    ctx.bfs_bank_create(3, scene.inflation_radius)
    _, goals = scenes.tabletop_queries(4, seed=33)
    seeds = api.world_to_grid(goals, scene.origin, scene.res).astype(np.int32)
    ctx.bfs_bank_run(seeds[:3])
    rng = np.random.default_rng(6)
    cells = rng.integers(0, np.asarray(scene.dims), (3000, 3)).astype(np.int32)
    before = [ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells) for s in range(3)]
    # slot 1 goes to a new query; slots 0 and 2 keep searching with their own distances
    sl = np.array([1], np.int32)
    assert ctx._ck(ctx.L.smplgpu_bfs_bank_run_slots(ctx.h, api._ip(sl), api._ip(np.ascontiguousarray(seeds[3:4])), 1), "run_slots") == 1
    after = [ctx.bfs_bank_distances(np.full(len(cells), s, np.int32), cells) for s in range(3)]
    assert np.array_equal(before[0], after[0]) and np.array_equal(before[2], after[2])
    ctx.bfs_set_walls_from_df(scene.inflation_radius)
    ctx.bfs_run([seeds[3]])
    assert np.array_equal(after[1], ctx.bfs_distances(cells))
    assert not np.array_equal(before[1], after[1])


def test_expand_batch_matches_oracle(tabletop):
    scene, o, ctx, tables = tabletop
    lo, hi, cont = tables.limits()
    n = 3000
    q = scenes.random_states(n, lo, hi, cont, seed=41)
    q0, q1 = scenes.mprim_edges(q)
    ctx.bfs_bank_create(2, scene.inflation_radius)
    goals = np.array([[0.5, -0.3, 0.8], [0.6, 0.2, 1.0]])
    seeds = api.world_to_grid(goals, scene.origin, scene.res)
    ctx.bfs_bank_run(seeds)
    slot = (np.arange(n) % 2).astype(np.int32)
    v, h, g, off = ctx.expand_batch(q0, q1, slot, scene.cost_per_cell)
    ev, _ = o.is_edges_valid(q0, q1)
    assert np.array_equal(v, ev)
    pose = o.planning_frame_fk(q1)
    assert np.abs(off - pose[:, :3]).max() < 1e-13
    for s in range(2):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        o.heur_set_goal(*goals[s])
        sel = slot == s
        assert np.array_equal(h[sel], o.goal_heuristics(q1[sel]))


def _run_oracle(o, scene, params, starts, goals):
    out = []
    for s, g in zip(starts, goals):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)   # a fresh BfsHeuristic per query
        out.append(o.plan(s, g, params))
    return out


def test_batch_planner_matches_sequential_oracle(tabletop):
    scene, o, ctx, tables = tabletop
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 4000
    starts, goals = scenes.tabletop_queries(24, seed=13)
    goals[5] = (0.4, -0.2, 0.36)        # the demo's goal (pr2_goal.yaml:38-44): under the table top, hard
    goals[7] = (5.0, 0.0, 1.0)          # outside the grid: heuristic is Infinity everywhere
    starts[9, 1] = 3.0                  # start violates joint limits: setStart fails
    ref = _run_oracle(o, scene, params, starts, goals)
    got, stats = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=10, n_threads=3)
    n_ok = 0
    for i, (a, b) in enumerate(zip(ref, got)):
        assert a["success"] == b["success"], i
        assert a["expansions"] == b["expansions"], (i, a["expansions"], b["expansions"])
        assert a["cost"] == b["cost"], i
        assert a["num_states"] == b["num_states"], i
        assert np.array_equal(a["path_ids"], b["path_ids"]), i
        n_ok += a["success"]
    assert n_ok >= 12 and not ref[9]["success"]
    assert stats["rounds"] > 0 and stats["edges_submitted"] > 0
    print("planner parity: %d queries, %d solved, %d rounds, %d edges, device %.3fs host %.3fs" % (
        len(ref), n_ok, stats["rounds"], stats["edges_submitted"], stats["device_seconds"], stats["host_seconds"]))


def test_batch_planner_with_action_weights_matches_oracle(tabletop):
    """Edge cost = int(1000 * weight) with the primitive file's weight column (manip_lattice.cpp:296, 1414-1437;
    manip_lattice_action_space.cpp:182-190): the device tells the host which primitive set an expansion's successor
    words belong to (SMPLGPU_LATTICE_SHORT_FLAG in the count word), the host looks the weight up.  The oracle with the
    same weights is pinned against the reference build in tests/test_oracle_planner_reference.py."""
    scene, o, ctx, tables = tabletop
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 3000
    params.weights = [(1.0, 2.5, 0.4, 1.7)[i % 4] for i in range(len(params.mprims))]
    starts, goals = scenes.tabletop_queries(16, seed=13)
    ref = _run_oracle(o, scene, params, starts, goals)
    unit = scenes.PlanParams(scene.dof)
    unit.max_expansions = 3000
    ref_unit = _run_oracle(o, scene, unit, starts, goals)
    got, _ = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=6, n_threads=2)
    n_ok = 0
    for i, (a, b) in enumerate(zip(ref, got)):
        assert (a["success"], a["expansions"], a["cost"], a["num_states"]) == \
               (b["success"], b["expansions"], b["cost"], b["num_states"]), i
        assert np.array_equal(a["path_ids"], b["path_ids"]), i
        n_ok += a["success"]
    assert n_ok >= 8
    # the weights matter: costs are no longer multiples of 1000, and some plans change
    assert any(a["success"] and a["cost"] % 1000 != 0 for a in ref)
    assert sum(a["expansions"] != u["expansions"] for a, u in zip(ref, ref_unit)) >= 3


def test_one_planner_thread_per_context_gives_the_same_plans(tabletop):
    scene, o, ctx, tables = tabletop
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 800
    starts, goals = scenes.tabletop_queries(30, seed=17)
    single, _ = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=8)
    ctxs = [ctx] + [api.clone_context(ctx, scene, tables) for _ in range(2)]
    multi, stats = api.plan_batch(ctxs, scene, tables, params, starts, goals, max_concurrent=4)
    for a, b in zip(single, multi):
        assert (a["success"], a["expansions"], a["cost"], a["num_states"]) == \
               (b["success"], b["expansions"], b["cost"], b["num_states"])
        assert np.array_equal(a["path_ids"], b["path_ids"])
    assert stats["edges_submitted"] > 0
    for c in ctxs[1:]:
        c.close()


def _device_count():
    import ctypes as C
    n = C.c_int(0)
    try:
        rt = C.CDLL("libcudart.so")
        return n.value if rt.cudaGetDeviceCount(C.byref(n)) == 0 else 0
    except OSError:
        import torch
        return torch.cuda.device_count()


@pytest.mark.skipif(_device_count() < 2, reason="needs two GPUs in one process")
def test_contexts_on_different_gpus_in_one_process_give_the_same_plans(tabletop):
    """smplhost_plan_batch_multi with contexts on DIFFERENT devices: a C++ caller uses an 8-GPU box from one process
    (one planner thread per context, each bound to its context's device) without torchrun."""
    scene, o, ctx, tables = tabletop
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 800
    starts, goals = scenes.tabletop_queries(30, seed=17)
    single, _ = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=8)
    other, _ = api.setup_context(scene, device=1)
    ctxs = [ctx, other, api.clone_context(other, scene, tables, device=1)]
    multi, stats = api.plan_batch(ctxs, scene, tables, params, starts, goals, max_concurrent=4)
    for a, b in zip(single, multi):
        assert (a["success"], a["expansions"], a["cost"], a["num_states"]) == \
               (b["success"], b["expansions"], b["cost"], b["num_states"])
        assert np.array_equal(a["path_ids"], b["path_ids"])
    for c in ctxs[1:]:
        c.close()
    ctx.L.smplgpu_bind_thread(ctx.h)   # creating a context on device 1 made it this thread's current device


def test_ubr1_with_attached_object_queries_match_oracle():
    """Config 4 shape: UBR1 arm, attached box, extra ACM entries; queries dealt round-robin as on 2 GPUs."""
    from smpl_b200 import sharding
    scene = scenes.ubr1_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 1500
    starts, goals = scenes.ubr1_tabletop_queries(20, seed=13)
    ref = _run_oracle(o, scene, params, starts, goals)
    solved = 0
    for rank in range(2):
        mine = sharding.round_robin_shard(len(starts), rank, 2)
        got, _ = api.plan_batch(ctx, scene, tables, params, starts[mine], goals[mine], max_concurrent=6, n_threads=2)
        for i, b in zip(mine, got):
            a = ref[i]
            assert (a["success"], a["expansions"], a["cost"], a["num_states"]) == \
                   (b["success"], b["expansions"], b["cost"], b["num_states"]), i
            assert np.array_equal(a["path_ids"], b["path_ids"]), i
            solved += a["success"]
    assert solved >= 10
    ctx.close()


def test_batch_planner_reproduces_golden_plans(tabletop):
    scene, o, ctx, tables = tabletop
    gold = json.load(open(os.path.join(GOLD, "pr2_tabletop_plans.json")))
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = gold["max_expansions"]
    starts, goals = np.array(gold["starts"]), np.array(gold["goals"])
    got, _ = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=64)
    for g, r in zip(got, gold["results"]):
        assert [int(g["success"]), g["expansions"], g["cost"], g["num_states"]] == r[:4]
        assert list(map(int, g["path_ids"])) == r[4]

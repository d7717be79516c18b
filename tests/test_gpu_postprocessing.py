"""Path post-processing over the CUDA path: ManipLattice::extractPath of the batch
planner, ShortcutPath (joint-space variants) and InterpolatePath for many paths at once must give what the
reference-shaped sequential code of the oracle gives, point for point."""
import numpy as np
import pytest

from helpers import make_oracle
from smpl_b200 import api, scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def clutter():
    scene = scenes.pr2_clutter_scene()
    ctx, tables = api.setup_context(scene)
    o = make_oracle(scene, with_kdl=False)
    yield scene, o, ctx, tables
    ctx.close()


@pytest.fixture(scope="module")
def tabletop():
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    ctx, tables = api.setup_context(scene)
    yield scene, o, ctx, tables
    ctx.close()


def wandering_paths(tables, ctx, n_paths, seed, max_len=40):
    """Joint-space paths of mixed quality: random walks over the 22 motion primitives (short steps, many valid
    shortcuts), straight lines between random states cut into pieces with jitter, and degenerate ones."""
    lo, hi, cont = tables.limits()
    rng = np.random.default_rng(seed)
    paths = []
    q = scenes.random_states(4 * n_paths, lo, hi, cont, seed=seed)
    valid = ctx.is_states_valid(q).astype(bool)
    anchors = q[valid]
    for p in range(n_paths):
        n = int(rng.integers(3, max_len))
        a, b = anchors[rng.integers(0, len(anchors), 2)]
        kind = p % 3
        if kind == 0:      # random walk of small steps around a valid state
            steps = rng.normal(0.0, 0.06, (n, tables.dof))
            steps[rng.random(n) < 0.15] = 0.0            # repeated points (zero-length segments)
            pts = a + np.cumsum(steps, axis=0)
        elif kind == 1:    # a straight line with jitter
            t = np.linspace(0.0, 1.0, n)[:, None]
            pts = a + t * (b - a) * 0.5 + rng.normal(0.0, 0.02, (n, tables.dof))
        else:              # a detour: out and back
            t = np.concatenate([np.linspace(0, 1, n // 2 + 1), np.linspace(1, 0.2, n - n // 2 - 1)])[:, None]
            pts = a + t * (b - a) * 0.4
        paths.append(np.ascontiguousarray(pts[:n]))
    paths.append(anchors[:1].copy())                    # one point
    paths.append(anchors[:2].copy())                    # two points
    paths.append(np.zeros((0, tables.dof)))             # empty
    return paths


def test_indexed_edges_equal_plain_edges_and_oracle(clutter):
    scene, o, ctx, tables = clutter
    lo, hi, cont = tables.limits()
    pts = scenes.random_states(300, lo, hi, cont, seed=5)
    near = pts[:150] + np.random.default_rng(1).normal(0, 0.1, (150, tables.dof))
    pts = np.concatenate([pts, near])
    rng = np.random.default_rng(2)
    a = rng.integers(0, len(pts), 5000).astype(np.int32)
    b = np.where(rng.random(5000) < 0.6, (a + 300) % len(pts), rng.integers(0, len(pts), 5000)).astype(np.int32)
    a[:10] = b[:10]                                     # zero-length motions
    v, c = ctx.is_indexed_edges_valid(pts, a, b)
    v2, c2 = ctx.is_edges_valid(pts[a], pts[b])
    assert np.array_equal(v, v2) and np.array_equal(c, c2)
    ev, ec = o.is_edges_valid(pts[a], pts[b])
    assert np.array_equal(v, ev) and np.array_equal(c, ec)
    assert 0 < v.sum() < len(v)
    with pytest.raises(api.SmplGpuError):
        ctx.is_indexed_edges_valid(pts, [0, len(pts)], [1, 2])


@pytest.mark.parametrize("kind", [0, 1])
def test_shortcut_paths_match_sequential_oracle(clutter, kind):
    scene, o, ctx, tables = clutter
    _, _, cont = tables.limits()
    paths = wandering_paths(tables, ctx, 36, seed=21 + kind)
    got, stats = api.shortcut_paths(ctx, tables, paths, kind=kind)
    n_short, checks_seq = 0, 0
    for p, g in zip(paths, got):
        ref, checks = o.shortcut_path(p, cont, kind=kind) if len(p) else (np.zeros(0, np.int32), 0)
        assert np.array_equal(g, ref), (len(p), g, ref)
        n_short += len(g) < len(p)
        checks_seq += checks
    assert n_short >= 10                     # shortcuts happen ...
    assert any(2 < len(g) for g in got)      # ... and some motions are refused
    assert stats["device_calls"] == 1 and stats["edges_checked"] >= checks_seq // 2
    print("shortcut kind %d: %d paths, %d shortened, %d candidate motions in one call (sequential oracle: %d calls)" % (
        kind, len(paths), n_short, stats["edges_checked"], checks_seq))


def test_interpolate_paths_match_sequential_oracle(clutter):
    scene, o, ctx, tables = clutter
    paths = wandering_paths(tables, ctx, 18, seed=33, max_len=12)
    got, stats = api.interpolate_paths(ctx, tables, paths)
    grew = 0
    for p, g in zip(paths, got):
        ref = o.interpolate_path(p) if len(p) else np.zeros((0, tables.dof))
        assert g.shape == ref.shape and np.array_equal(g, ref)   # waypoints are formed with the same IEEE operations
        grew += len(g) > len(p)
    assert grew >= 6 and stats["device_calls"] == 1


def test_extract_path_and_shortcut_of_planned_paths(tabletop):
    scene, o, ctx, tables = tabletop
    _, _, cont = tables.limits()
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 3000
    starts, goals = scenes.tabletop_queries(16, seed=13)
    got, _ = api.plan_batch(ctx, scene, tables, params, starts, goals, max_concurrent=8, want_states=True)
    paths = []
    for i, (s, g, r) in enumerate(zip(starts, goals, got)):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        ref = o.plan(s, g, params)
        assert ref["success"] == r["success"] and np.array_equal(ref["path_ids"], r["path_ids"])
        if r["success"]:
            # ManipLattice::extractPath: identical joint values, the goal id resolved to a real lattice state
            assert r["path_states"].shape == ref["path_states"].shape
            assert np.array_equal(r["path_states"], ref["path_states"]), i
            assert np.array_equal(r["path_states"][0], s)
            paths.append(r["path_states"])
    assert len(paths) >= 6
    for kind in (0, 1):
        short, _ = api.shortcut_paths(ctx, tables, paths, kind=kind)
        for p, g in zip(paths, short):
            ref, _ = o.shortcut_path(p, cont, kind=kind)
            assert np.array_equal(g, ref)
            assert g[0] == 0 and g[-1] == len(p) - 1
        assert sum(len(g) for g in short) < sum(len(p) for p in paths)

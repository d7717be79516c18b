#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/.

Runs in the authoring container only:
  * bfs_reference_24.npz      -- inputs + outputs of the REFERENCE's own BFS_3D (compiled from
                                 /root/reference by `make -C oracle ref`), three seeds on a 24^3 grid;
  * pr2_right_arm_validity.csv -- `benchmark_cc export` format (benchmark_cc.cpp:1390-1409): q0..q6 verdict
                                 at 12 significant digits; verdicts from the oracle (parity unpinned:
                                 the reference ships no such file);
  * pr2_right_arm_edges.csv    -- q0[7] q1[7] verdict waypoint_count, same convention;
  * ubr1_attached_validity.csv -- UBR1 arm + attached box (config 4 shape).
The CSVs freeze the oracle's behaviour so that a later change to oracle/ cannot silently move the
target the CUDA path is compared with.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import make_oracle  # noqa: E402
from oracle_api import RefBfs  # noqa: E402
from smpl_b200 import scenes  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def round12(a):
    return np.array([[float("%.12g" % v) for v in row] for row in a])


def gen_plans():
    """pr2_tabletop_plans.json -- ARA* results of the oracle's ManipLattice driver on the config-1 scene."""
    import json
    scene = scenes.pr2_tabletop_scene()
    o = make_oracle(scene)
    params = scenes.PlanParams(scene.dof)
    params.max_expansions = 3000
    starts, goals = scenes.tabletop_queries(12, seed=113)
    starts, goals = round12(starts), round12(goals)
    results = []
    for s, g in zip(starts, goals):
        o.heur_init(scene.inflation_radius, scene.cost_per_cell)
        p = o.plan(s, g, params)
        results.append([int(p["success"]), p["expansions"], p["cost"], p["num_states"], [int(v) for v in p["path_ids"]]])
    json.dump({"max_expansions": params.max_expansions, "starts": starts.tolist(), "goals": goals.tolist(),
               "results": results}, open(os.path.join(OUT, "pr2_tabletop_plans.json"), "w"))


def gen_arastar():
    """arastar_reference.npz -- outputs of the REFERENCE's own ARA* (oracle/_ref/libref_arastar.so) on the graphs
    of tests/test_oracle_arastar.py."""
    from oracle_api import arastar_search
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_oracle_arastar as T
    out = {}
    for k, (seed, eps, max_exp) in enumerate(T.CASES):
        off, dst, cost, h, start, goal = T.lattice_graph(seed, weird_h=seed != 3)
        r = arastar_search("reference", off, dst, cost, h, start, goal, eps, max_exp)
        out["summary_%d" % k] = np.array([int(r["found"]), r["cost"], r["expansions"]], np.int32)
        out["path_%d" % k] = r["path"]
        print("arastar case %d: found=%s cost=%d expansions=%d" % (k, r["found"], r["cost"], r["expansions"]))
    np.savez_compressed(os.path.join(OUT, "arastar_reference.npz"), **out)


def gen_shortcut():
    """shortcut_reference.npz -- outputs of the REFERENCE's own shortcut templates (oracle/_ref/libref_shortcut.so)
    on the tables of tests/test_oracle_shortcut.py."""
    from oracle_api import shortcut_table
    import test_oracle_shortcut as T
    out = {}
    for k, (seed, algo, gran) in enumerate(T.CASES):
        costs, valid, pair = T.table_case(seed)
        out["out_%d" % k] = shortcut_table("reference", costs, valid, pair, algo, gran)
    np.savez_compressed(os.path.join(OUT, "shortcut_reference.npz"), **out)
    print("shortcut: %d cases" % len(T.CASES))


def gen_voxelize():
    """voxelize_reference.npz -- voxel lists of the REFERENCE's own voxeliser (oracle/_ref/libref_voxelize.so)."""
    from oracle_api import voxelize_box, voxelize_mesh
    import test_oracle_voxelize as T
    out = {}
    for seed in range(T.N_BOX):
        size, pose, res, origin, fill = T.box_case(seed)
        out["box_%d" % seed] = voxelize_box("reference", size, pose, res, origin, fill)
    for seed in range(T.N_SOUP):
        v, t, res, origin, fill = T.soup_case(seed)
        out["soup_%d" % seed] = voxelize_mesh("reference", v, t, res, origin, fill)
    np.savez_compressed(os.path.join(OUT, "voxelize_reference.npz"), **out)
    print("voxelize: %d voxels" % sum(len(v) for v in out.values()))


def gen_distmap():
    """distmap_reference.npz -- fields of the REFERENCE's own EuclidDistanceMap (oracle/_ref/libref_distmap.so)."""
    from oracle_api import RefDistanceMap
    import test_oracle_distance_map as T
    scene = T.small_scene()
    o = make_oracle(scene)
    ref = RefDistanceMap(scene.origin, scene.size, scene.res, scene.max_dist)
    ref.add_points(T.obstacle_points(o))
    out = {"d2_after_add": ref.d2().astype(np.uint16)}
    rng = np.random.default_rng(8)
    more = np.array(scene.origin) - 0.05 + rng.random((300, 3)) * (np.array(scene.size) + 0.1)
    more = np.concatenate([more, more[:20]])
    ref.add_points(more)
    ref.remove_points(more[:150])
    out["d2_after_remove"] = ref.d2().astype(np.uint16)
    np.savez_compressed(os.path.join(OUT, "distmap_reference.npz"), **out)
    print("distmap: dims", ref.dims, "max d2", int(out["d2_after_add"].max()))


def gen_bfs():
    rng = np.random.default_rng(24)
    walls = (rng.random((24, 24, 24)) < 0.3).astype(np.uint8)
    walls[10:12, :, :] = 1
    walls[10:12, 5:8, 5:8] = 0
    seeds = np.array([scenes.first_free_cell(walls, (12, 12, 12)), (0, 0, 0), (23, 23, 5)], np.int32)
    out = {"walls": walls, "seeds": seeds}
    for k, s in enumerate(seeds):
        b = RefBfs(24, 24, 24)
        b.set_walls(walls)
        b.run(*s)
        out["dist_%d" % k] = b.grid()
        b.close()
    np.savez_compressed(os.path.join(OUT, "bfs_reference_24.npz"), **out)


def gen_pr2():
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    lo, hi, cont = o.joint_limits()
    q = round12(scenes.random_states(768, lo, hi, cont, seed=99))
    v = o.is_states_valid(q)
    with open(os.path.join(OUT, "pr2_right_arm_validity.csv"), "w") as f:
        for row, vv in zip(q, v):
            f.write(" ".join("%.12g" % x for x in row) + " %d\n" % vv)
    q0 = q[:512]
    q1 = round12(scenes.mprim_edges(q0)[1])
    ev, ec = o.is_edges_valid(q0, q1)
    with open(os.path.join(OUT, "pr2_right_arm_edges.csv"), "w") as f:
        for a, b, vv, cc in zip(q0, q1, ev, ec):
            f.write(" ".join("%.12g" % x for x in np.concatenate([a, b])) + " %d %d\n" % (vv, cc))
    print("pr2: %.1f%% valid states, %.1f%% valid edges" % (100 * v.mean(), 100 * ev.mean()))


def gen_ubr1():
    scene = scenes.ubr1_tabletop_scene()
    o = make_oracle(scene)
    lo, hi, cont = o.joint_limits()
    q = round12(scenes.random_states(512, lo, hi, cont, seed=98))
    v = o.is_states_valid(q)
    with open(os.path.join(OUT, "ubr1_attached_validity.csv"), "w") as f:
        for row, vv in zip(q, v):
            f.write(" ".join("%.12g" % x for x in row) + " %d\n" % vv)
    print("ubr1: %.1f%% valid states" % (100 * v.mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1:] or ["bfs", "pr2", "ubr1", "plans", "arastar", "distmap", "shortcut", "voxelize"]
    for name in which:
        {"bfs": gen_bfs, "pr2": gen_pr2, "ubr1": gen_ubr1, "plans": gen_plans, "arastar": gen_arastar, "distmap": gen_distmap,
         "shortcut": gen_shortcut, "voxelize": gen_voxelize}[name]()

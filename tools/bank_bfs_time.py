"""Time one BFS-bank run (one BFS per planning query, all slots at once) with both wavefront kernels."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
from smpl_b200 import api, scenes
slots = int(sys.argv[1]) if len(sys.argv) > 1 else 64
scene = scenes.pr2_tabletop_scene()
ctx, tables = api.setup_context(scene)
_, goals = scenes.tabletop_queries(slots, seed=13)
seeds = api.world_to_grid(goals, scene.origin, scene.res).astype(np.int32)
ref = None
rng = np.random.default_rng(1)
cells = rng.integers(0, np.asarray(scene.dims), (20000, 3)).astype(np.int32)
sl = rng.integers(0, slots, 20000).astype(np.int32)
for mode in (ctx.BFS_LEVELS, ctx.BFS_TILES):
    ctx.bfs_set_mode(mode)
    ctx.bfs_bank_create(slots, scene.inflation_radius)
    ctx.bfs_bank_run(seeds)
    t0 = time.perf_counter()
    for _ in range(3):
        ctx.bfs_bank_run(seeds)
    dt = (time.perf_counter() - t0) / 3
    d = ctx.bfs_bank_distances(sl, cells)
    same = True if ref is None else bool(np.array_equal(ref, d))
    ref = d if ref is None else ref
    print("bank of %d x 150^3, mode %d: %.3f ms per run (%.3f ms per query), same=%s" % (slots, mode, dt * 1e3, dt * 1e3 / slots, same))

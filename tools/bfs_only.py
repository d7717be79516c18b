import sys, time
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from smpl_b200 import api, scenes
n = int(sys.argv[1]) if len(sys.argv) > 1 else 400
ctx = api.GpuContext(0)
walls = scenes.bfs_clutter_walls(n, seed=11)
seed = scenes.first_free_cell(walls, (n // 2, n // 2, n // 2))
modes = [int(a) for a in sys.argv[2:]] or [ctx.BFS_TILES]
ref = None
for mode in modes:
    ctx.bfs_set_mode(mode)
    ctx.bfs_set_walls(walls)
    for _ in range(2):
        ctx.bfs_run([seed])
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        ctx.bfs_run([seed])
    dt = (time.perf_counter() - t0) / 5
    print("bfs %d^3 mode %d: %.3f ms, %d levels, %.1f Mvoxel/s" % (n, mode, dt * 1e3, ctx.bfs_last_levels(), n ** 3 / dt / 1e6))
    if len(modes) > 1:
        d = ctx.bfs_download()
        if ref is None:
            ref = d
        else:
            print("  same distances as first mode:", bool(np.array_equal(ref, d)))

"""e2e of the lattice entry points with K host threads, one context each (the reference's threading model: one
CollisionSpace per thread), each validating 1/K of the step's states and edges from pinned host buffers."""
import ctypes as C
import sys
import threading
import time

sys.path.insert(0, "/root/repo")
sys.path.insert(0, "/root/repo/tests")
import numpy as np
import torch

import bench
from smpl_b200 import api, scenes

n = 1 << 20
scene = scenes.pr2_clutter_scene()
ctx, tables = api.setup_context(scene)
lo, hi, cont = tables.limits()
res, coords, q, pid8, deltas, q1 = bench.sweep_inputs(lo, hi, cont, n, 0)
c_i16_p = C.POINTER(C.c_int16)
for K in (1, 2, 3, 4, 6, 8):
    ctxs = [ctx] + [api.clone_context(ctx, scene, tables) for _ in range(K - 1)]
    for c in ctxs:
        c.set_lattice(res)
    hc = torch.from_numpy(coords).pin_memory()
    hp = torch.from_numpy(pid8).pin_memory()
    hv = torch.empty(n, dtype=torch.uint8).pin_memory()
    hev = torch.empty(n, dtype=torch.uint8).pin_memory()
    dp = deltas.ctypes.data_as(api.c_double_p)
    bounds = np.linspace(0, n, K + 1).astype(int)
    L = ctx.L
    steps = 10
    start = threading.Barrier(K + 1)

    def work(k):
        c = ctxs[k]
        L.smplgpu_bind_thread(c.h)
        b, e = int(bounds[k]), int(bounds[k + 1])
        m = e - b
        pc = C.cast(hc.data_ptr() + b * coords.shape[1] * 2, c_i16_p)
        pp = C.cast(hp.data_ptr() + b, api.c_uint8_p)
        pv = C.cast(hv.data_ptr() + b, api.c_uint8_p)
        pe = C.cast(hev.data_ptr() + b, api.c_uint8_p)
        for it in range(steps + 2):
            if it == 2:
                start.wait()
            r = L.smplgpu_is_lattice_states_valid(c.h, pc, m, pv)
            r |= L.smplgpu_is_lattice_edges_valid(c.h, pc, pp, m, dp, len(deltas), pe, None)
            assert r == 0

    ts = [threading.Thread(target=work, args=(k,)) for k in range(K)]
    for t in ts:
        t.start()
    start.wait()
    t0 = time.perf_counter()
    for t in ts:
        t.join()
    dt = (time.perf_counter() - t0) / steps
    units = n + int(ctx.is_edges_valid(q[:65536], q1[:65536])[1].sum()) * 16
    print("K=%d: %.3f ms per step, ~%.2f G validated states/s" % (K, dt * 1e3, units / dt / 1e9))
    for c in ctxs[1:]:
        c.close()

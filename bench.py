#!/usr/bin/env python3
"""Benchmark of the validity + BFS hot path (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W          our CUDA path
    python bench.py --impl reference ...                    the reference's CPU algorithm on the host cores

A step = one pass of the hot path over one batch of synthetic input of the shape
BASELINE.json configs[1] names: `--states` random PR2 right-arm states plus as many
motion-primitive edges, against the 2 m^3 / 2 cm clutter scene.  Every rank (GPU)
processes its own full batch (weak scaling: independent queries, no collective on
the data path; the distance field is built on rank 0 and broadcast once over NCCL).
The unit is a validated state: one per state plus `waypoint_count` per edge, the
states CollisionSpace::isStateToStateValid accounts for (collision_space.cpp:538-581).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "states validated/s"
UNIT = "states/s"


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for k, nm in enumerate(names):
                if f[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_oracle(scene, prime_q):
    from oracle_api import OracleScene
    o = OracleScene(scene.robot_path, scene.group, scene.planning_joints, scene.origin, scene.size, scene.res,
                    scene.max_dist)
    for k, v in scene.fixed_joints.items():
        o.set_joint(k, v)
    if scene.use_desc_acm:
        o.use_desc_acm()
    for a, b, allowed in scene.acm_extra:
        o.acm_set(a, b, allowed)
    if len(scene.cells):
        o.add_cells(scene.cells)
    if len(getattr(scene, "boxes", [])):
        o.insert_boxes(scene.boxes)
    o.prime(prime_q)
    return o


def make_reference_checker(scene, prime_q):
    """The reference's OWN collision checker (sbpl_collision_checking compiled from its sources into
    oracle/_ref/libref_collision.so, see oracle/ref_collision_shim.cpp) set up for `scene`, or None when that library
    was not built (then the CPU legs time the oracle port)."""
    from oracle_api import RefCollisionScene, ref_collision_lib
    if ref_collision_lib() is None or os.environ.get("SMPL_BENCH_CPU_PORT"):
        return None
    r = RefCollisionScene(scene.robot_path, scene.group, scene.planning_joints, scene.origin, scene.size, scene.res,
                          scene.max_dist)
    for k, v in scene.fixed_joints.items():
        r.set_joint(k, v)
    if scene.use_desc_acm:
        r.use_desc_acm()
    for a, b, allowed in scene.acm_extra:
        r.acm_set(a, b, allowed)
    if len(scene.cells):
        r.add_cells(scene.cells)
    if len(getattr(scene, "boxes", [])):
        r.insert_boxes(scene.boxes)
    r.prime(prime_q)
    return r


def make_cpu_checker(scene, prime_q):
    """-> (checker, kind, description): the compiled reference when available ("reference"), else the port."""
    r = make_reference_checker(scene, prime_q)
    if r is not None:
        return r, "reference", "the reference's own sbpl_collision_checking CollisionSpace (oracle/_ref/libref_collision.so)"
    return make_oracle(scene, prime_q), "port", "oracle (CPU port of sbpl_collision_checking)"


def cpu_validity_rate(scene, q, q0, q1, threads, checkers=None):
    """The CPU checker (one instance per thread: CollisionSpace is not reentrant) on `threads` host threads;
    returns (units/s, units, seconds)."""
    n = len(q)
    parts = np.array_split(np.arange(n), threads)
    oracles = checkers if checkers is not None else [make_cpu_checker(scene, q[0])[0] for _ in range(threads)]
    units = [0] * threads

    def work(i):
        idx = parts[i]
        if len(idx) == 0:
            return
        oracles[i].time_states_valid(q[idx])
        _, _, c = oracles[i].time_edges_valid(q0[idx], q1[idx])
        units[i] = len(idx) + int(c.sum())

    t0 = time.perf_counter()
    if threads == 1:
        work(0)
    else:
        ts = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
    dt = time.perf_counter() - t0
    return sum(units) / dt, sum(units), dt


def cpu_bfs_rate(n):
    """The reference's own BFS_3D (oracle/_ref) when present, else the port; Mvoxel/s on one core (+1 search thread)."""
    from oracle_api import OracleBfs, RefBfs, ref_lib
    from smpl_b200 import scenes
    walls = scenes.bfs_clutter_walls(n, seed=11)
    seed = scenes.first_free_cell(walls, (n // 2, n // 2, n // 2))
    kind = "reference" if ref_lib() is not None else "port"
    b = RefBfs(n, n, n) if kind == "reference" else OracleBfs(n, n, n)
    b.set_walls(walls)
    dt = b.time_run(*seed)
    b.close()
    return n ** 3 / dt / 1e6, kind, dt


class StdoutToStderr:
    """stdout carries exactly one JSON line: anything a library prints while this is active (NCCL's version
    banner at the first collective, at file-descriptor level) goes to stderr instead."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def setup_shared_scene(scene, local_rank, rank, world, dev):
    """Context for `scene` on this rank's GPU: rank 0 builds the distance field on its GPU, every other rank
    receives it in ONE broadcast (NCCL over NVLink) -- the only collective of the data path."""
    import torch
    from smpl_b200 import api, sharding
    ctx = api.GpuContext(local_rank)
    tables = api.build_tables(scene)
    ctx.set_robot(tables)
    if rank == 0:
        ctx.build_distance_field(api.scene_cells(scene, tables), scene.dims, scene.origin, scene.res, scene.max_dist,
                                 scene.padding)
    if world > 1:
        d2 = ctx.download_distance_field() if rank == 0 else None
        df_t = sharding.broadcast_distance_field(d2, scene.dims, src=0, device=dev)
        torch.cuda.synchronize()
        if rank != 0:
            dmax = int(np.ceil(scene.max_dist * (1.0 / scene.res)))
            ctx.set_distance_field_dev(df_t.data_ptr(), scene.dims, scene.origin, scene.res, dmax * dmax, scene.padding)
    return ctx, tables


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU collision checker (oracle/_ref/libref_collision.so; the oracle port
    when that was not built) on all host threads, one CollisionSpace per thread."""
    if rank != 0:
        return
    from smpl_b200 import scenes
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene, np.zeros(scene.dof))
    o.init_kdl(scene.chain_root, scene.chain_tip, scene.planning_link, scene.T_kin_to_planning)
    lo, hi, cont = o.joint_limits()
    n = args.ref_sample
    threads = os.cpu_count() or 1
    q = scenes.random_states(n, lo, hi, cont, seed=20260101)
    q0, q1 = scenes.mprim_edges(q)
    made = [make_cpu_checker(scene, q[0]) for _ in range(threads)]
    checkers, kind, what = [m[0] for m in made], made[0][1], made[0][2]
    for _ in range(args.warmup):
        cpu_validity_rate(scene, q[: n // 8], q0[: n // 8], q1[: n // 8], threads, checkers)
    total_units, total_t = 0, 0.0
    for _ in range(args.steps):
        _, u, dt = cpu_validity_rate(scene, q, q0, q1, threads, checkers)
        total_units += u
        total_t += dt
    value = total_units / total_t
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total_t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "config[1] validity sweep: PR2 right arm states + mprim edges vs 2 m^3 clutter scene @ 2 cm",
                   "states_per_step": n, "edges_per_step": n},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind,
                         "sample": "%d states + %d edges per step, %s on %d threads" % (n, n, what, threads)},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def wandering_paths(anchors, n_paths, dof, seed):
    """Joint-space paths for the shortcut leg: random walks of small steps, jittered straight lines and detours
    between valid states (20 to 60 points), the shapes tests/test_gpu_postprocessing.py checks against the oracle."""
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


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="smpl_b200", choices=["smpl_b200", "reference"])
    ap.add_argument("--states", type=int, default=1 << 20, help="states (and edges) per step per GPU")
    ap.add_argument("--bfs-n", type=int, default=400)
    ap.add_argument("--cpu-sample", type=int, default=1 << 19, help="states (and edges) timed on the CPU oracle")
    ap.add_argument("--ref-sample", type=int, default=1 << 17)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--plan-queries", type=int, default=2048, help="planning queries per GPU (0 = skip)")
    ap.add_argument("--plan-concurrent", type=int, default=2048, help="queries in flight per GPU (one BFS grid each)")
    ap.add_argument("--plan-max-expansions", type=int, default=2000)
    ap.add_argument("--plan-cpu-queries", type=int, default=12)
    ap.add_argument("--post-paths", type=int, default=1024, help="joint-space paths shortcut in one call (0 = skip)")
    ap.add_argument("--post-cpu-paths", type=int, default=48)
    ap.add_argument("--no-ingest", dest="ingest", action="store_false", help="skip the scene-ingest leg")
    ap.add_argument("--plan-threads", type=int, default=0, help="planner threads (= contexts) per GPU; 0 = 75 %% of the rank's cores, at most 12")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "smpl_b200" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from smpl_b200 import api, scenes, sharding

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    # ---- scene: rank 0 builds the distance field on its GPU, then ONE broadcast over NCCL ----
    scene = scenes.pr2_clutter_scene()
    with StdoutToStderr():
        if world > 1:
            dist.init_process_group("nccl", device_id=dev)
        ctx, tables = setup_shared_scene(scene, local_rank, rank, world, dev)
        if world > 1:
            dist.barrier()
    # time on ONE explicit stream shared by torch (events) and the library (kernels, copies)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    lo, hi, cont = tables.limits()

    n = args.states
    q = scenes.random_states(n, lo, hi, cont, seed=20260101 + rank)
    q0, q1 = scenes.mprim_edges(q)
    dof = scene.dof

    # ---- resident inputs (value) ----
    d_q = torch.from_numpy(q).to(dev)
    d_q0 = torch.from_numpy(q0).to(dev)
    d_q1 = torch.from_numpy(q1).to(dev)
    d_v = torch.empty(n, dtype=torch.uint8, device=dev)
    d_ev = torch.empty(n, dtype=torch.uint8, device=dev)
    d_cnt = torch.empty(n, dtype=torch.int32, device=dev)

    def step_resident():
        ctx.is_states_valid_dev(d_q.data_ptr(), n, d_v.data_ptr())
        ctx.is_edges_valid_dev(d_q0.data_ptr(), d_q1.data_ptr(), n, d_ev.data_ptr(), d_cnt.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_resident()
    barrier()
    units_per_step = n + int(d_cnt.sum().item())
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    barrier()
    ev[0].record()
    for s in range(args.steps):
        ctx.is_states_valid_dev(d_q.data_ptr(), n, d_v.data_ptr())
        ev[2 * s + 1].record()
        ctx.is_edges_valid_dev(d_q0.data_ptr(), d_q1.data_ptr(), n, d_ev.data_ptr(), d_cnt.data_ptr())
        ev[2 * s + 2].record()
    barrier()
    gpu_launches = ctx.launch_count() - launches0
    total_ms = ev[0].elapsed_time(ev[-1])
    states_ms = np.mean([ev[2 * s].elapsed_time(ev[2 * s + 1]) for s in range(args.steps)])
    edges_ms = np.mean([ev[2 * s + 1].elapsed_time(ev[2 * s + 2]) for s in range(args.steps)])
    t_ms = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    total_ms_max = float(t_ms.item())
    gpu_stats = ctx.last_validity_stats()
    gpu_stats["f64_resolved_edges_last_launch"] = ctx.last_f64_resolved()
    cert_in_use, e_pos, eps_cells = ctx.certified_bounds()
    gpu_stats["certified_f32"] = {"in_use": cert_in_use, "e_pos_m": e_pos, "eps_cells": eps_cells}
    df_lookup_peak = ctx.probe_df_lookup_rate() if rank == 0 else None   # independent random lookups/s on this field

    # ---- end to end through the C ABI with HOST buffers (pinned), copies inside the timed region ----
    hq = torch.from_numpy(q).pin_memory()
    hq0 = torch.from_numpy(q0).pin_memory()
    deltas = np.ascontiguousarray(scenes.pr2_mprim_deltas())
    hpid = torch.from_numpy((np.arange(n) % len(deltas)).astype(np.int32)).pin_memory()   # edge i = state i + primitive i mod 22
    hv = torch.empty(n, dtype=torch.uint8).pin_memory()
    hev = torch.empty(n, dtype=torch.uint8).pin_memory()
    L = ctx.L
    import ctypes as C

    def step_e2e():
        r = L.smplgpu_is_states_valid(ctx.h, C.cast(hq.data_ptr(), api.c_double_p), n, C.cast(hv.data_ptr(), api.c_uint8_p))
        # edges as GetSuccs produces them: (parent state, motion primitive id) against the primitive table
        r |= L.smplgpu_is_mprim_edges_valid(ctx.h, C.cast(hq0.data_ptr(), api.c_double_p), C.cast(hpid.data_ptr(), api.c_int32_p),
                                            n, deltas.ctypes.data_as(api.c_double_p), len(deltas),
                                            C.cast(hev.data_ptr(), api.c_uint8_p), None)
        if r != 0:
            raise RuntimeError(L.smplgpu_last_error(ctx.h).decode())

    for _ in range(2):
        step_e2e()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(3, args.steps // 2)
    e0.record()
    for _ in range(e2e_steps):
        step_e2e()
    e1.record()
    barrier()
    e2e_ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_ms, op=dist.ReduceOp.MAX)
    e2e_value = world * units_per_step * e2e_steps / (float(e2e_ms.item()) * 1e-3)
    assert torch.equal(hv.to(dev), d_v) and torch.equal(hev.to(dev), d_ev), "host-buffer path disagrees with resident path"
    verdict_s, verdict_e, counts_e = hv.numpy().copy(), hev.numpy().copy(), d_cnt.cpu().numpy()

    # ---- BFS (config[2]): 400^3 cluttered occupancy, rank 0 only for the side metric ----
    bfs = None
    if rank == 0 and args.bfs_n > 0:
        nb = args.bfs_n
        walls = scenes.bfs_clutter_walls(nb, seed=11)
        seed = scenes.first_free_cell(walls, (nb // 2, nb // 2, nb // 2))
        ctx.bfs_set_walls(walls)
        for _ in range(2):
            ctx.bfs_run([seed])
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 5
        torch.cuda.synchronize()
        b0.record()
        for _ in range(reps):
            ctx.bfs_run([seed])
        b1.record()
        torch.cuda.synchronize()
        bfs_ms = b0.elapsed_time(b1) / reps
        alg_bytes = (nb + 2) ** 3 * (1.0 / 8 + 4)
        bfs = {"grid": "%d^3" % nb, "levels": ctx.bfs_last_levels(), "ms": bfs_ms,
               "mvoxel_s": nb ** 3 / (bfs_ms * 1e-3) / 1e6,
               "algorithmic_bytes": alg_bytes, "achieved_gbs": alg_bytes / (bfs_ms * 1e-3) / 1e9}

    # ---- path post-processing: shortcut many paths with ONE batch of candidate motions ----
    post = None
    post_paths = None
    if rank == 0 and args.post_paths > 0:
        post_paths = wandering_paths(q[d_v.cpu().numpy().astype(bool)], args.post_paths, dof, seed=5)
        api.shortcut_paths(ctx, tables, post_paths, kind=0)     # warm-up at full size (buffers, lazily loaded kernels)
        t0 = time.perf_counter()
        short, pst = api.shortcut_paths(ctx, tables, post_paths, kind=0)
        dt = time.perf_counter() - t0
        post = {"paths": len(post_paths), "points": int(sum(len(p) for p in post_paths)),
                "points_after": int(sum(len(g) for g in short)), "candidate_motions": pst["edges_checked"],
                "seconds": dt, "device_seconds": pst["device_seconds"], "paths_per_s": len(post_paths) / dt,
                "call": "smplhost_shortcut_paths (ShortcutPath JOINT_SPACE): all point pairs of every path in one "
                        "smplgpu_is_indexed_edges_valid call", "_short": short}

    # ---- scene ingest: box objects -> surface voxels -> distance field, all on the device ----
    ingest = None
    if rank == 0 and args.ingest:
        iscene = scenes.pr2_shelf_objects_scene()
        ictx = api.GpuContext(local_rank)
        itables = api.build_tables(iscene)
        ictx.set_robot(itables)
        icells = api.scene_cells(iscene, itables)
        iv, it = api.box_meshes(iscene.boxes)
        for _ in range(2):
            ictx.build_distance_field_from_meshes(iv, it, icells, iscene.dims, iscene.origin, iscene.res, iscene.max_dist)
        t0 = time.perf_counter()
        reps = 10
        for _ in range(reps):
            ictx.build_distance_field_from_meshes(iv, it, icells, iscene.dims, iscene.origin, iscene.res, iscene.max_dist)
        dt = (time.perf_counter() - t0) / reps
        d2 = ictx.download_distance_field()
        ingest = {"scene": "34 box objects (12 triangles each, arbitrary poses) in 2 m^3 @ 2 cm", "triangles": int(len(it)),
                  "grid_cells": int(d2.size), "occupied_cells": int((d2 == 0).sum()), "ms": dt * 1e3,
                  "call": "smplgpu_build_distance_field_from_meshes: voxelise + addPointsToField + distance field, host call to field ready",
                  "_d2": d2, "_scene": iscene}
        ictx.close()

    clocks = sampler.stop() if rank == 0 else None   # sampled across the validity, end-to-end, BFS, shortcut and ingest regions

    # ---- plan queries/s (config[0]/[3] shape): PR2 right arm on the tabletop scene, queries sharded over ranks ----
    plan = None
    if args.plan_queries > 0:
        pscene = scenes.pr2_tabletop_scene()
        pctx, ptables = setup_shared_scene(pscene, local_rank, rank, world, dev)
        pparams = scenes.PlanParams(pscene.dof)
        pparams.max_expansions = args.plan_max_expansions
        nq_total = args.plan_queries * world
        starts_all, goals_all = scenes.tabletop_queries(nq_total, seed=13)
        mine = sharding.round_robin_shard(nq_total, rank, world)   # no collective
        # one planner thread per context (the reference's threading model), all on this rank's GPU
        # the planner is bound by the host-side lattice / OPEN-list work once enough queries are in flight
        # (measured on 16 cores, 2048 queries: 8 threads 916 q/s, 12 threads 1229 q/s)
        n_thr = args.plan_threads if args.plan_threads > 0 else max(1, min(12, int(0.75 * (os.cpu_count() or 1) / max(1, world))))
        # a context's BFS bank keeps int node indices (<= 611 slots of 150^3): enough contexts for the queries in flight
        n_thr = max(n_thr, (args.plan_concurrent + 511) // 512)
        pctxs = [pctx] + [api.clone_context(pctx, pscene, ptables, device=local_rank) for _ in range(n_thr - 1)]
        per_ctx = max(1, (args.plan_concurrent + n_thr - 1) // n_thr)
        # warm-up with the same bank shape: the BFS bank (a scene-level allocation of per_ctx grids per context) is
        # created here and reused by the timed call
        wq = min(len(mine), per_ctx * n_thr)
        wparams = scenes.PlanParams(pscene.dof)
        wparams.max_expansions = 20
        api.plan_batch(pctxs, pscene, ptables, wparams, starts_all[mine][:wq], goals_all[mine][:wq], max_concurrent=per_ctx)
        barrier()
        psampler = ClockSampler(local_rank)
        if rank == 0:
            psampler.start()
        t0 = time.perf_counter()
        pres, pstats = api.plan_batch(pctxs, pscene, ptables, pparams, starts_all[mine], goals_all[mine],
                                      max_concurrent=per_ctx)
        dt = time.perf_counter() - t0
        pclocks = psampler.stop() if rank == 0 else None
        pstats["planner_threads"] = n_thr
        t_plan = torch.tensor([dt], device=dev, dtype=torch.float64)
        n_exp = torch.tensor([float(sum(r["expansions"] for r in pres)), float(sum(r["success"] for r in pres))],
                             device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t_plan, op=dist.ReduceOp.MAX)
            dist.all_reduce(n_exp, op=dist.ReduceOp.SUM)
        plan = {"scene": "PR2 right arm, tabletop env, 150^3 field @ 2 cm, ARA* eps 100, first solution, <= %d expansions" % args.plan_max_expansions,
                "queries": nq_total, "solved": int(n_exp[1].item()), "expansions": int(n_exp[0].item()),
                "seconds": float(t_plan.item()), "queries_per_s": nq_total / float(t_plan.item()),
                "expansions_per_s": float(n_exp[0].item()) / float(t_plan.item()), "concurrent_per_gpu": args.plan_concurrent,
                "rank0": pstats, "clocks": pclocks}
        for c in pctxs:
            c.close()

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel (rank 0) ----
    peak, peak_src = load_peaks()
    cpu = None
    Lbar_state, Lbar_edge = None, None
    if not args.no_cpu:
        m = min(args.cpu_sample, n)
        o = make_oracle(scene, q[0])
        checker, cpu_kind, cpu_what = make_cpu_checker(scene, q[0])
        t_s, v_s = checker.time_states_valid(q[:m])
        t_e, v_e, c = checker.time_edges_valid(q0[:m], q1[:m])
        cpu_units = m + int(c.sum())
        cpu_rate = cpu_units / (t_s + t_e)
        cpu_agree = None
        if verdict_s is not None:      # the timed CPU run doubles as a parity check of the device verdicts
            cpu_agree = {"states_differing": int((v_s != verdict_s[:m]).sum()), "edges_differing": int((v_e != verdict_e[:m]).sum()),
                         "waypoint_counts_differing": int((c != counts_e[:m]).sum()), "of": m}
        # L-bar: DF lookups the reference semantics requires (no early-out), from the oracle on a sub-sample
        ms = min(m, 1 << 15)
        _, Ls, _, _ = o.report_states(q[:ms])
        _, _, Le = o.report_edges(q0[:ms], q1[:ms])
        Lbar_state, Lbar_edge = float(Ls.mean()), float(Le.mean())
        cpu = {"value": cpu_rate, "unit": UNIT, "cores": 1, "kind": cpu_kind,
               "sample": "%d states + %d edges of the same workload, %s, 1 thread; "
                         "states %.0f/s, edge waypoints %.0f/s" % (m, m, cpu_what, m / t_s, int(c.sum()) / t_e),
               "device_verdicts_vs_this_run": cpu_agree}
        if plan is not None and args.plan_cpu_queries > 0:
            po = make_oracle(pscene, np.zeros(pscene.dof))
            po.init_kdl(pscene.chain_root, pscene.chain_tip, pscene.planning_link, pscene.T_kin_to_planning, pscene.xyz_offset)
            k = min(args.plan_cpu_queries, len(starts_all))
            secs, cexp = 0.0, 0
            same = 0
            # the reference's own ManipLattice + BfsHeuristic + ARAStar + CollisionSpace (oracle/ref_planner_shim.cpp)
            # when its build is there, else the oracle's restatement
            pref = make_reference_checker(pscene, np.zeros(pscene.dof))
            for qi, (s_, g_) in enumerate(zip(starts_all[:k], goals_all[:k])):
                po.heur_init(pscene.inflation_radius, pscene.cost_per_cell)
                t0 = time.perf_counter()
                pr = pref.plan(pscene, s_, g_, pparams) if pref is not None else po.plan(s_, g_, pparams)
                secs += time.perf_counter() - t0      # includes the per-query BFS, as the GPU figure does
                cexp += pr["expansions"]
                if world == 1 or qi % world == 0:     # rank 0 planned queries 0, world, 2 world, ...
                    gr = pres[qi // world]
                    same += int((gr["success"], gr["expansions"], gr["cost"]) == (pr["success"], pr["expansions"], pr["cost"])
                                and np.array_equal(gr["path_ids"], pr["path_ids"]))
                else:
                    same += 1
            plan["parity_checked_queries"] = k
            plan["parity_identical"] = same
            plan["cpu_queries_per_s"] = k / secs
            plan["cpu_expansions_per_s"] = cexp / secs
            plan["cpu_kind"] = "reference" if pref is not None else "port"
            plan["cpu_sample"] = "first %d queries, %s, 1 thread" % (
                k, "the reference's own ManipLattice + BfsHeuristic + ARAStar + CollisionSpace (oracle/_ref/libref_collision.so; "
                   "RobotModel and action-space plug-ins from the oracle)" if pref is not None else "oracle ManipLattice + ARA*")
        if post is not None:
            k = min(args.post_cpu_paths, len(post_paths))
            t0 = time.perf_counter()
            same = 0
            # the reference's own ShortcutPath (post_processing.cpp over its CollisionSpace) when its build is there
            if cpu_kind == "reference":
                for p_, g_ in zip(post_paths[:k], post["_short"][:k]):
                    same += int(np.array_equal(checker.post_process(scene, p_, 0), p_[g_]))
            else:
                for p_, g_ in zip(post_paths[:k], post["_short"][:k]):
                    ref_idx, _ = o.shortcut_path(p_, cont, kind=0)
                    same += int(np.array_equal(ref_idx, g_))
            dt = time.perf_counter() - t0
            post["cpu_paths_per_s"] = k / dt
            post["cpu_kind"] = cpu_kind
            post["cpu_sample"] = "first %d paths, %s ShortcutPath JOINT_SPACE (one isStateToStateValid per request), 1 thread" % (
                k, "the reference's own" if cpu_kind == "reference" else "oracle")
            post["parity_identical"] = "%d / %d" % (same, k)
        if ingest is not None:
            t0 = time.perf_counter()
            io = make_reference_checker(ingest["_scene"], np.zeros(dof))   # WorldCollisionModel::insertObject per box
            ingest_kind = "reference"
            if io is None:
                io = make_oracle(ingest["_scene"], np.zeros(dof))  # VoxelizeBox per object + addPointsToField + propagation
                ingest_kind = "port"
            dt = time.perf_counter() - t0
            ref_d2 = io.df_d2()
            ingest["cpu_ms"] = dt * 1e3
            ingest["cpu_kind"] = ingest_kind
            ingest["cpu_sample"] = "%s: scene construction incl. insertObject (VoxelizeBox + addPointsToField) + DistanceMap propagation, 1 thread" % (
                "the reference's own CollisionSpace + OccupancyGrid" if ingest_kind == "reference" else "oracle")
            ingest["occupied_cells_identical"] = bool(np.array_equal(ref_d2 == 0, ingest["_d2"] == 0))
            ingest["cells_where_reference_propagation_is_inexact"] = int((ref_d2 != ingest["_d2"]).sum())
        if bfs is not None:
            mv, kind, dt = cpu_bfs_rate(args.bfs_n)
            bfs["cpu_mvoxel_s"] = mv
            bfs["cpu_kind"] = kind
            bfs["cpu_seconds"] = dt
    if Lbar_state is None:
        Lbar_state = gpu_stats["df_lookups"] / max(1, gpu_stats["waypoints"])
        Lbar_edge = Lbar_state * (units_per_step - n) / n
    state_bytes = n * (8 * dof + 1 + 32 * Lbar_state)
    edge_bytes = n * (16 * dof + 1 + 32 * Lbar_edge)
    k_states = {"kernel": "states_valid32_kernel (+ f64 resolve pass)", "ms": float(states_ms), "algorithmic_bytes": state_bytes,
                "achieved": state_bytes / (states_ms * 1e-3) / 1e9}
    k_edges = {"kernel": "edges_valid32_kernel (+ f64 resolve pass)", "ms": float(edges_ms), "algorithmic_bytes": edge_bytes,
               "achieved": edge_bytes / (edges_ms * 1e-3) / 1e9}
    dom, other = (k_edges, k_states) if edges_ms >= states_ms else (k_states, k_edges)
    # DRAM traffic of the dominant kernel from the committed ncu --set full capture (profiles/), scaled to this
    # launch's item count (the traffic is the streamed inputs; the distance field stays in L2)
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        key = "edges_valid32_kernel" if dom is k_edges else "states_valid32_kernel"
        traffic = tr[key]["dram_bytes"] * n / tr[key]["items"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom["achieved"], "peak": peak, "unit": "GB/s",
                "frac": dom["achieved"] / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dom["algorithmic_bytes"], "launch_ms": dom["ms"],
                "Lbar_state": Lbar_state, "Lbar_edge": Lbar_edge,
                "note": "issue/latency bound (FK arithmetic + dependent L2 lookups), not HBM bound: the HBM fraction is reported "
                        "as the contract asks; see DESIGN.md and profiles/ for pipe utilisation and stall reasons",
                "other_kernel": other}
    # SURVEY 8d's two fractions for the dominant kernel: HBM = streamed bytes only; L2 = the lookups the reference
    # semantics requires (one 32-byte sector each) against the measured rate of INDEPENDENT random lookups on the
    # same field (smplgpu_probe_df_lookup_rate): the kernels' lookups are dependent, so this ceiling is not reachable
    if df_lookup_peak:
        Lb, per_item = (Lbar_edge, 16 * dof + 1) if dom is k_edges else (Lbar_state, 8 * dof + 1)
        roofline["hbm_frac_streamed_bytes"] = n * per_item / (dom["ms"] * 1e-3) / 1e9 / peak
        roofline["l2"] = {"required_lookups_per_launch": n * Lb, "achieved_glookups_s": n * Lb / (dom["ms"] * 1e-3) / 1e9,
                          "peak_glookups_s": df_lookup_peak / 1e9, "frac": n * Lb / (dom["ms"] * 1e-3) / df_lookup_peak,
                          "achieved_gbs": 32 * n * Lb / (dom["ms"] * 1e-3) / 1e9, "peak_gbs": 32 * df_lookup_peak / 1e9,
                          "peak_source": "measured live: independent random lookups on the loaded field, 8 in flight per thread"}
    if bfs is not None:
        roofline["bfs"] = {"bound": "hbm", "achieved": bfs["achieved_gbs"], "peak": peak, "unit": "GB/s",
                           "frac": bfs["achieved_gbs"] / peak}

    value = world * units_per_step * args.steps / (total_ms_max * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32 certified, f64 resolves the undecidable items (verdicts = all-f64)", "data": "synthetic",
        "config": {"workload": "config[1] validity sweep: PR2 right arm (7-DOF) states + mprim edges vs 2 m^3 clutter scene @ 2 cm",
                   "states_per_step_per_gpu": n, "edges_per_step_per_gpu": n, "validated_states_per_step_per_gpu": units_per_step,
                   "l2_policy": "inputs (%.0f MB/step) exceed the 126 MB L2; the 2 MB distance field is L2-resident by design" % ((3 * n * dof * 8) / 1e6),
                   "valid_fraction_states": float(d_v.float().mean().item()), "valid_fraction_edges": float(d_ev.float().mean().item())},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 2 * n * dof * 8 + 4 * n, "d2h_bytes_per_step": 2 * n,
                "calls": "smplgpu_is_states_valid(q) + smplgpu_is_mprim_edges_valid(q, primitive ids, table): host buffers in, verdicts out"},
        "gpu_launches": int(gpu_launches),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "bfs": bfs,
        "plan": plan,
        "post_processing": post,
        "scene_ingest": ingest,
        "host_cores": os.cpu_count(),
        "gpu_stats_last_launch": gpu_stats,
    }
    if post is not None:
        post.pop("_short", None)
    if ingest is not None:
        ingest.pop("_d2", None)
        ingest.pop("_scene", None)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

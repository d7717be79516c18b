"""N > 1 host-side logic on CPU (gloo, world_size 2): sharding of states / queries, the single scene broadcast,
and that per-rank results reassemble into the single-process answer (the oracle stands in for the device)."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT
from smpl_b200 import sharding


def test_shards_partition_the_work():
    for n in (0, 1, 7, 4096, 1000003):
        for world in (1, 2, 4, 8):
            cont = [sharding.contiguous_shard(n, r, world) for r in range(world)]
            assert cont[0][0] == 0 and cont[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(cont, cont[1:]))
            rr = np.concatenate([sharding.round_robin_shard(n, r, world) for r in range(world)])
            assert np.array_equal(np.sort(rr), np.arange(n))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist
    from helpers import make_oracle
    from smpl_b200 import scenes, sharding as sh

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene, with_kdl=False)
    # rank 0 owns the scene; everyone else receives the field in ONE broadcast
    d2 = o.df_d2().astype(np.uint16) if rank == 0 else None
    t = sh.broadcast_distance_field(d2, scene.dims, src=0)
    field = t.numpy().view(np.uint16).reshape(scene.dims)
    same = np.array_equal(field.astype(np.int32), o.df_d2())
    # the in-place form bench.py uses on the GPUs: every rank owns a buffer (the library's resident field on the source,
    # a reserved one elsewhere) and the broadcast lands in it directly
    own = o.df_d2().astype(np.uint16).copy() if rank == 0 else np.zeros(scene.dims, np.uint16)
    sh.broadcast_field_in_place(own.ctypes.data, own.nbytes, src=0, device=None)
    same = same and np.array_equal(own.astype(np.int32), o.df_d2())
    q = scenes.random_states(4001, *np.load(os.path.join(out_dir, "limits.npy"), allow_pickle=True), seed=71)
    b, e = sh.contiguous_shard(len(q), rank, world)
    v = o.is_states_valid(q[b:e])
    total = sh.gather_counts(len(v))
    np.save(os.path.join(out_dir, "v%d.npy" % rank), v)
    np.save(os.path.join(out_dir, "meta%d.npy" % rank), np.array([same, total, b, e]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_reassemble_the_single_process_answer(tmp_path):
    from helpers import make_oracle
    from smpl_b200 import scenes
    scene = scenes.pr2_clutter_scene()
    o = make_oracle(scene)
    lo, hi, cont = o.joint_limits()
    np.save(os.path.join(tmp_path, "limits.npy"), np.array([lo, hi, cont], dtype=object), allow_pickle=True)
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    q = scenes.random_states(4001, lo, hi, cont, seed=71)
    want = o.is_states_valid(q)
    parts, metas = [], []
    for r in range(2):
        parts.append(np.load(os.path.join(tmp_path, "v%d.npy" % r)))
        metas.append(np.load(os.path.join(tmp_path, "meta%d.npy" % r)))
    assert all(m[0] == 1 for m in metas), "broadcast field differs from the source field"
    assert all(m[1] == len(q) for m in metas), "all_reduce of the unit counts"
    assert np.array_equal(np.concatenate(parts), want)

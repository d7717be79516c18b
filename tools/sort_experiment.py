"""How much would coherent warps buy?  Times the resident validity kernels on the bench workload as is and with the
items reordered by bins of the most proximal joints (results per item are unchanged by the order)."""
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np, torch
from smpl_b200 import api, scenes

scene = scenes.pr2_clutter_scene()
ctx, tables = api.setup_context(scene)
lo, hi, cont = tables.limits()
n = 1 << 20
q = scenes.random_states(n, lo, hi, cont, seed=20260101)
q0, q1 = scenes.mprim_edges(q)
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)


def run(tag, order):
    d_q = torch.from_numpy(np.ascontiguousarray(q[order])).to(dev)
    d_q0 = torch.from_numpy(np.ascontiguousarray(q0[order])).to(dev)
    d_q1 = torch.from_numpy(np.ascontiguousarray(q1[order])).to(dev)
    d_v = torch.empty(n, dtype=torch.uint8, device=dev)
    d_e = torch.empty(n, dtype=torch.uint8, device=dev)
    for _ in range(3):
        ctx.is_states_valid_dev(d_q.data_ptr(), n, d_v.data_ptr())
        ctx.is_edges_valid_dev(d_q0.data_ptr(), d_q1.data_ptr(), n, d_e.data_ptr(), None)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    torch.cuda.synchronize()
    ev[0].record()
    for _ in range(10):
        ctx.is_states_valid_dev(d_q.data_ptr(), n, d_v.data_ptr())
    ev[1].record()
    for _ in range(10):
        ctx.is_edges_valid_dev(d_q0.data_ptr(), d_q1.data_ptr(), n, d_e.data_ptr(), None)
    ev[2].record()
    torch.cuda.synchronize()
    print("%-28s states %.3f ms  edges %.3f ms   valid %.3f / %.3f" % (
        tag, ev[0].elapsed_time(ev[1]) / 10, ev[1].elapsed_time(ev[2]) / 10, d_v.float().mean().item(), d_e.float().mean().item()))


def bins(k, nb):
    key = np.zeros(n, np.int64)
    lo_ = np.where(np.asarray(cont, bool), -np.pi, lo)
    hi_ = np.where(np.asarray(cont, bool), np.pi, hi)
    for j in range(k):
        b = np.clip(((q[:, j] - lo_[j]) / (hi_[j] - lo_[j]) * nb).astype(np.int64), 0, nb - 1)
        key = key * nb + b
    return np.argsort(key, kind="stable")


run("as generated", np.arange(n))
for k, nb in ((1, 64), (2, 16), (3, 16), (4, 8), (7, 4)):
    run("binned: %d joints x %d bins" % (k, nb), bins(k, nb))

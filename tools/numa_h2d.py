"""Does the NUMA node of the page-locked buffer matter for the e2e leg?  Pinned H2D bandwidth with the allocating
thread bound to each NUMA node in turn, next to the GPU's own CPU affinity as NVML reports it."""
import glob
import os

import torch


def cpulist(s):
    out = []
    for part in s.strip().split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def h2d_gbs(n=128 << 20, reps=20):
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h.fill_(1)   # first touch on the current node
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    return n * reps / (e0.elapsed_time(e1) * 1e-3) / 1e9


torch.cuda.init()
allowed = sorted(os.sched_getaffinity(0))
print("allowed cpus:", len(allowed), allowed[:4], "...", allowed[-4:])
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    gpu_cpus = [64 * w + b for w, x in enumerate(words) for b in range(64) if (x >> b) & 1]
    print("GPU0 cpu affinity (NVML):", len(gpu_cpus), gpu_cpus[:4], "...", gpu_cpus[-4:])
except Exception as e:  # noqa: BLE001
    print("NVML affinity unavailable:", e)
print("unbound: %.1f GB/s" % h2d_gbs())
for node in sorted(glob.glob("/sys/devices/system/node/node[0-9]*")):
    cpus = [c for c in cpulist(open(node + "/cpulist").read()) if c in allowed]
    if not cpus:
        print(os.path.basename(node), "no allowed cpus")
        continue
    os.sched_setaffinity(0, cpus)
    print("%s (%d allowed cpus): %.1f GB/s" % (os.path.basename(node), len(cpus), h2d_gbs()))
    os.sched_setaffinity(0, allowed)

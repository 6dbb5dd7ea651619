#!/usr/bin/env python
"""Host topology probe: does the placement of pinned buffers (NUMA node of the allocating thread) change the
PCIe copy bandwidth?  Prints the topology and, per CPU set, H2D / D2H GB/s of a 2 GiB pinned buffer."""
import os
import subprocess
import time

import torch


def sh(cmd):
    try:
        return subprocess.run(cmd, shell=True, capture_output=True, text=True, timeout=20).stdout.strip()
    except Exception as e:  # noqa: BLE001
        return "ERR %s" % e


print(sh("nvidia-smi topo -m | head -8"))
print(sh("lscpu | grep -i -E 'numa|socket|model name|^CPU\\(s\\)'"))
print("affinity:", sorted(os.sched_getaffinity(0)))
try:
    import pynvml

    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    n = (os.cpu_count() + 63) // 64
    mask = pynvml.nvmlDeviceGetCpuAffinity(h, n)
    cpus = [i * 64 + b for i, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
    print("nvml ideal cpus for GPU0:", cpus[:64], "... total", len(cpus))
except Exception as e:  # noqa: BLE001
    print("nvml affinity unavailable:", e)
    cpus = []
nodes = {}
for d in sorted(os.listdir("/sys/devices/system/node")) if os.path.isdir("/sys/devices/system/node") else []:
    if d.startswith("node"):
        nodes[d] = open("/sys/devices/system/node/%s/cpulist" % d).read().strip()
print("nodes:", nodes)

torch.cuda.set_device(0)
dev = torch.device("cuda", 0)
allowed = sorted(os.sched_getaffinity(0))


def parse(cl):
    out = []
    for part in cl.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


sets = {"all": allowed}
for name, cl in nodes.items():
    s = [c for c in parse(cl) if c in allowed]
    if s:
        sets[name] = s
if cpus:
    s = [c for c in cpus if c in allowed]
    if s:
        sets["nvml_ideal"] = s
d = torch.empty(1 << 31, dtype=torch.uint8, device=dev)
for name, s in sets.items():
    os.sched_setaffinity(0, s)
    h = torch.empty(1 << 31, dtype=torch.uint8, pin_memory=True)
    h.fill_(1)
    res = []
    for direction in ("h2d", "d2h"):
        best = 0
        for _ in range(3):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if direction == "h2d":
                d.copy_(h, non_blocking=True)
            else:
                h.copy_(d, non_blocking=True)
            torch.cuda.synchronize()
            best = max(best, h.numel() / (time.perf_counter() - t0) / 1e9)
        res.append("%s %.1f GB/s" % (direction, best))
    # both directions at once
    h2 = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True)
    d2 = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    with torch.cuda.stream(s1):
        d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2):
        h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res.append("duplex: h2d 2GiB + d2h 1GiB in %.1f ms" % (dt * 1e3))
    print(name, "cpus", s[:8], "...", len(s), "|", " | ".join(res), flush=True)
    del h, h2, d2
os.sched_setaffinity(0, allowed)

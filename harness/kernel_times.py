#!/usr/bin/env python
"""Per-kernel device time of one call of a bench.py secondary workload, from CUPTI (torch.profiler sees every kernel of the
process, including the C-ABI library's): python harness/kernel_times.py <contacts|broadphase|mixed|mesh_contacts> [K=V,K=V ...]"""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from torch.profiler import ProfilerActivity, profile
import bench
import parry_b200

FN = {"contacts": lambda *a: bench.also_contacts(*a, e2e=False), "broadphase": bench.also_broadphase, "mixed": bench.also_mixed,
      "mesh_contacts": bench.also_mesh_contacts}
ctx = parry_b200.Context(0)
stream = ctx.torch_stream()


def timed_once(fn, steps=1, warmup=2, flush=None):
    for _ in range(warmup):
        fn()
    ctx.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        ctx.synchronize()
        torch.cuda.synchronize()
    rows = sorted(prof.key_averages(), key=lambda e: -e.device_time_total)
    for e in rows[:14]:
        print("   %10.3f ms  x%-4d %s" % (e.device_time_total / 1e3, e.count, e.key[:110]))
    return sum(e.device_time_total for e in rows) / 1e3


for setting in (sys.argv[2:] or ["-"]):
    keys = []
    if setting != "-":
        for kv in setting.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
            keys.append(k)
    print("==", setting, flush=True)
    FN[sys.argv[1]](ctx, stream, timed_once, None, 6553.6)
    for k in keys:
        del os.environ[k]

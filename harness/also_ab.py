#!/usr/bin/env python
"""A/B of one of bench.py's secondary workloads under environment switches:
python harness/also_ab.py <name> [K=V[,K=V] ...]   (name: contacts | broadphase | mixed | mesh_contacts; '-' = defaults)"""
import json
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import parry_b200

FN = {"contacts": lambda *a: bench.also_contacts(*a, e2e=False), "broadphase": bench.also_broadphase, "mixed": bench.also_mixed,
      "mesh_contacts": bench.also_mesh_contacts, "siblings": bench.also_siblings, "manifolds": bench.also_manifolds}
ctx = parry_b200.Context(0)
stream = ctx.torch_stream()
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for setting in ["-"] + sys.argv[2:] + ["-"]:
    keys = []
    if setting != "-":
        for kv in setting.split(","):
            k, v = kv.split("=")
            os.environ[k] = v
            keys.append(k)
    r = FN[sys.argv[1]](ctx, stream, bench.make_timed(ctx, stream), flush, 6553.6)
    print(setting, json.dumps({k: v for k, v in r.items() if k in ("ms", "value", "contacts_fraction", "pairs_per_frame", "contacts_per_frame") or isinstance(v, dict)}), flush=True)
    for k in keys:
        del os.environ[k]

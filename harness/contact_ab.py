#!/usr/bin/env python
"""A/B timing of the 4M-pair contact config under environment switches: python harness/contact_ab.py KEY=V [KEY=V ...]
(each argument is one setting compared against the default; '-' is the default)."""
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import parry_b200
from harness import scenes

ctx = parry_b200.Context(0)
stream = ctx.torch_stream()
pts, radii = scenes.hull_pool(4096)
G = parry_b200.Shapes(ctx, [parry_b200.ConvexPolyhedron(p) for p in pts])
n = 1 << 22
a, b, p1, p2 = scenes.hull_pairs(n, radii, seed=4)
da, db = torch.from_numpy(a.astype(np.int32)).cuda(), torch.from_numpy(b.astype(np.int32)).cuda()
dp1, dp2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
for setting in ["-"] + sys.argv[1:] + ["-"]:
    if setting != "-":
        k, v = setting.split("=")
        os.environ[k] = v
    ts = []
    for it in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        out, st = parry_b200.contact(G, da, dp1, db, dp2, 0.01)
        e1.record(stream)
        ctx.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%s: min %.3f ms, checksum %d %.6f" % (setting, min(ts[1:]), int(st.to(torch.int64).sum().item()), float(out.double().nan_to_num().sum().item())), flush=True)
    if setting != "-":
        del os.environ[setting.split("=")[0]]

#!/usr/bin/env python
"""Device-resident vs host-buffer ray casts on the bench terrain for a list of ray seeds (they must agree bit for bit):
python harness/ray_paths_check.py 12 13"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import parry_b200
from harness import scenes
FMAX = float(np.finfo(np.float32).max)
ctx = parry_b200.Context(0)
v, i = scenes.terrain(2001, 2001)
mesh = parry_b200.TriMesh(ctx, v, i)
m = 1 << 23
for seed in [int(x) for x in sys.argv[1:]]:
    rays = scenes.terrain_rays(m, seed=seed)
    with torch.cuda.stream(ctx.torch_stream()):
        rd = torch.from_numpy(rays).cuda()
    ctx.synchronize()
    res = []
    for rep in range(3):
        t, k = mesh.cast_local_ray(rd, FMAX)
        ctx.synchronize()
        res.append((t.cpu().numpy().copy(), k.cpu().numpy().view(np.uint32).copy()))
    ht, hk = mesh.cast_local_ray(rays, FMAX)
    ht, hk = np.asarray(ht), np.asarray(hk).view(np.uint32)
    for rep in range(3):
        dt, dk = res[rep]
        bad = np.nonzero((dt.view(np.uint32) != ht.view(np.uint32)) | (dk != hk))[0]
        print("seed %d rep %d: %d mismatches device vs host" % (seed, rep, len(bad)), bad[:8], dt[bad[:4]], ht[bad[:4]], dk[bad[:4]], hk[bad[:4]])
    d01 = np.nonzero((res[0][0].view(np.uint32) != res[1][0].view(np.uint32)) | (res[0][1] != res[1][1]))[0]
    print("   device run 0 vs run 1: %d mismatches" % len(d01))

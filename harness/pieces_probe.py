#!/usr/bin/env python
"""Single-GPU probe of the piece-wise ray cast used by the multi-GPU gather (no peers): what do the pieces themselves cost?"""
import ctypes as C
import os
import sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import parry_b200
from harness import scenes

FMAX = float(np.finfo(np.float32).max)
ctx = parry_b200.Context(0)
stream = ctx.torch_stream()
v, i = scenes.terrain(2001, 2001)
m = 1 << 23
rays = torch.from_numpy(scenes.terrain_rays(m, seed=6)).cuda()
mesh = parry_b200.TriMesh(ctx, v, i)
toi = torch.empty(m, dtype=torch.float32, device="cuda")
tri = torch.empty(m, dtype=torch.int32, device="cuda")
# a second pair of buffers on the same device plays the peer: the pushes and their flag waits run for real
toi2 = torch.zeros(m, dtype=torch.float32, device="cuda")
tri2 = torch.zeros(m, dtype=torch.int32, device="cuda")
p_toi = (C.c_void_p * 2)(toi.data_ptr(), toi2.data_ptr())
p_tri = (C.c_void_p * 2)(tri.data_ptr(), tri2.data_ptr())


def timeit(fn, name):
    for _ in range(3):
        fn()
    ctx.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        ctx.synchronize()
        ts.append(e0.elapsed_time(e1))
    print("%s: min %.3f ms med %.3f ms" % (name, min(ts), float(np.median(ts))), flush=True)


timeit(lambda: mesh.cast_local_ray(rays, FMAX, out=(toi, tri)), "plain")
ref_toi, ref_tri = toi.clone(), tri.clone()
for chunks in [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8, 16]:
    toi.zero_(); tri.zero_(); toi2.zero_(); tri2.zero_()
    torch.cuda.synchronize()
    timeit(lambda: mesh.cast_local_ray_allgather(rays, FMAX, p_toi, p_tri, 0, 0, chunks), "pieces=%d" % chunks)
    ok = bool((toi.view(torch.int32) == ref_toi.view(torch.int32)).all() and (tri == ref_tri).all())
    pushed = bool((toi2.view(torch.int32) == ref_toi.view(torch.int32)).all() and (tri2 == ref_tri).all())
    print("   results identical to the plain cast: %s; pushed copy identical: %s" % (ok, pushed), flush=True)

#!/usr/bin/env python
"""Small workloads through every kernel family, for compute-sanitizer (memcheck / racecheck / synccheck):
compute-sanitizer --tool racecheck python harness/sanitize_probe.py"""
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
what = set(sys.argv[1:]) or {"rays", "pieces", "contacts", "siblings", "persistence"}
if "rays" in what or "pieces" in what:
    v, i = scenes.terrain(65, 65)
    mesh = parry_b200.TriMesh(ctx, v, i)
    m = 40000
    rays = scenes.terrain_rays(m, seed=21)
    if "rays" in what:
        toi, tri, n, f = mesh.cast_local_ray_and_get_normal(rays, FMAX)
        print("rays:", int((tri != 0xFFFFFFFF).sum()), "hits", flush=True)
    if "pieces" in what:
        rd = torch.from_numpy(rays).cuda()
        t1, i1 = torch.zeros(m, device="cuda"), torch.zeros(m, dtype=torch.int32, device="cuda")
        t2, i2 = torch.zeros(m, device="cuda"), torch.zeros(m, dtype=torch.int32, device="cuda")
        p_toi = (C.c_void_p * 2)(t1.data_ptr(), t2.data_ptr())
        p_tri = (C.c_void_p * 2)(i1.data_ptr(), i2.data_ptr())
        mesh.cast_local_ray_allgather(rd, FMAX, p_toi, p_tri, 0, 0, 4)
        ctx.synchronize()
        print("pieces: pushed copy identical:", bool((t1 == t2).all() and (i1 == i2).all()), flush=True)
if "contacts" in what:
    pts, radii = scenes.hull_pool(64)
    G = parry_b200.Shapes(ctx, [parry_b200.ConvexPolyhedron(p) for p in pts])
    a, b, p1, p2 = scenes.hull_pairs(6000, radii, seed=4)
    out, st = parry_b200.contact(G, a % 64, p1, b % 64, p2, 0.01)
    print("contacts:", int((st == 1).sum()), "of 6000", flush=True)
if "siblings" in what:
    z = np.load(os.path.join(ROOT, "tests", "golden", "siblings_3000.npz"))
    pu = z["params"].view(np.uint32)
    spec = [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p[:3]) if k == 1 else parry_b200.ConvexPolyhedron(z["points"][u[0]:u[0] + u[1]])
            for k, p, u in zip(z["kinds"], z["params"], pu)]
    G = parry_b200.Shapes(ctx, spec)
    o, st = parry_b200.cast_shapes(G, z["shape1"], z["pos1"], z["vel1"], z["shape2"], z["pos2"], z["vel2"])
    compounds = [[(z["part_pose"][f + i], int(z["part_shape"][f + i])) for i in range(c)] for f, c in zip(z["comp_first"], z["comp_count"])]
    Cc = parry_b200.Compounds(ctx, G, compounds)
    o2, st2, part = Cc.contact_shapes(z["compound_id"], z["pos1"], z["shape2"], z["pos2_compound"], 0.05)
    nr, cnt, mp, st3 = parry_b200.contact_manifolds(G, z["man_shape1"], z["pos1"], z["man_shape2"], z["man_pos2"], 0.05)
    cp, kind, st4 = parry_b200.closest_points(G, z["shape1"], z["pos1"], z["shape2"], z["pos2"], 0.5)
    print("siblings:", int((st != 0).sum()), int((st2 == 1).sum()), int(cnt.sum()), np.bincount(kind).tolist(), flush=True)
if "persistence" in what:
    # the paths that first run on hardware in round 2: second-frame manifolds (host and device arrays) and Compound vs Compound
    z = np.load(os.path.join(ROOT, "tests", "golden", "siblings_3000.npz"))
    pu = z["params"].view(np.uint32)
    spec = [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p[:3]) if k == 1 else parry_b200.ConvexPolyhedron(z["points"][u[0]:u[0] + u[1]])
            for k, p, u in zip(z["kinds"], z["params"], pu)]
    G = parry_b200.Shapes(ctx, spec)
    nr, cnt, mp, st3 = parry_b200.contact_manifolds(G, z["man_shape1"], z["pos1"], z["man_shape2"], z["man_pos2"], 0.05)
    moved = z["man_pos2"].copy()
    moved[::2, 4:] += np.float32(2.0e-4)
    moved[1::2, 4:] += np.float32(0.03)
    u = parry_b200.contact_manifolds_update(G, z["man_shape1"], z["pos1"], z["man_shape2"], moved, 0.05, nr, cnt, mp)
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    ud = parry_b200.contact_manifolds_update(G, dev(z["man_shape1"]), dev(z["pos1"]), dev(z["man_shape2"]), dev(moved), 0.05, dev(nr), dev(cnt), dev(mp))
    ctx.synchronize()
    kept, q = parry_b200.manifolds_try_update(ctx, z["pos1"], moved, nr, cnt, mp)
    compounds = [[(z["part_pose"][f + i], int(z["part_shape"][f + i])) for i in range(c)] for f, c in zip(z["comp_first"], z["comp_count"])]
    Cc = parry_b200.Compounds(ctx, G, compounds)
    ids2 = np.roll(z["compound_id"], 1)
    o5, st5, parts = Cc.contact_compounds(z["compound_id"], z["pos1"], ids2, z["pos2_compound"], 0.05)
    print("persistence:", int(u[4].sum()), "kept,", int((ud[4].cpu().numpy() == u[4]).all()), "device == host,", int(kept.sum()), "try_update kept,",
          int((st5 == 1).sum()), "compound-compound contacts", flush=True)
print("done", flush=True)

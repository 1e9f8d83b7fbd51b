#!/usr/bin/env python
"""Small single-workload driver for ncu captures / quick timings: python harness/prof.py <workload> [iters]
workloads: rays_terrain, rays_sphere1m, contacts, broadphase, mesh_queries, colliders_queries"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
FMAX = float(np.finfo(np.float32).max)


def main():
    import torch
    import parry_b200
    from harness import scenes
    wl = sys.argv[1]
    iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    ctx = parry_b200.Context(0)
    stream = ctx.torch_stream()

    def timeit(fn, name):
        fn()
        ctx.synchronize()
        ts = []
        for _ in range(iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            fn()
            e1.record(stream)
            ctx.synchronize()
            ts.append(e0.elapsed_time(e1))
        print("%s: min %.3f ms  med %.3f ms" % (name, min(ts), float(np.median(ts))))
        return min(ts)

    if wl.startswith("rays"):
        if wl == "rays_terrain":
            v, i = scenes.terrain(2001, 2001)
            m = 1 << 23
            rays = scenes.terrain_rays(m, seed=6)
        else:
            v, i = scenes.uv_sphere(708, 707)
            m = 1 << int(os.environ.get("PB2_PROF_RAYS_LOG2", "20"))
            rays = scenes.sphere_rays(m, seed=1)
        for rep in range(2):
            t0 = time.perf_counter()
            mesh = parry_b200.TriMesh(ctx, v, i)
            ctx.synchronize()
            print("TriMesh::new (host arrays, incl. upload) %.2f ms" % ((time.perf_counter() - t0) * 1e3))
        rd = torch.from_numpy(rays).cuda()
        toi = torch.empty(m, dtype=torch.float32, device="cuda")
        tri = torch.empty(m, dtype=torch.int32, device="cuda")
        # PB2_SWEEP="K=V,K=V;K=V;..." runs one timing per ';'-separated group of environment settings
        sweep = os.environ.get("PB2_SWEEP")
        configs = [None] + (sweep.split(";") if sweep else [])
        touched = set()
        for cfg in configs:
            for k in touched:
                os.environ.pop(k, None)
            if cfg:
                for kv in cfg.split(","):
                    k, v = kv.split("=")
                    os.environ[k] = v
                    touched.add(k)
            ms = timeit(lambda: mesh.cast_local_ray(rd, FMAX, out=(toi, tri)), "%s %s" % (wl, cfg))
            print("   %.1f Mrays/s  checksum %d %.6f" % (m / ms / 1e3, int(tri.to(torch.int64).sum().item()), float(toi.double().sum().item())))
    elif wl == "contacts":
        pts, radii = scenes.hull_pool(4096)
        G = parry_b200.Shapes(ctx, [parry_b200.ConvexPolyhedron(p) for p in pts])
        n = 1 << 22
        a, b, p1, p2 = scenes.hull_pairs(n, radii, seed=4)
        da, db = torch.from_numpy(a.astype(np.int32)).cuda(), torch.from_numpy(b.astype(np.int32)).cuda()
        dp1, dp2 = torch.from_numpy(p1).cuda(), torch.from_numpy(p2).cuda()
        res = {}

        def run():
            res["o"] = parry_b200.contact(G, da, dp1, db, dp2, 0.01)
        ms = timeit(run, wl)
        print("%.1f Mpairs/s" % (n / ms / 1e3))
        out, st = res["o"]
        print("checksum", int(st.to(torch.int64).sum().item()), float(out.double().nan_to_num().sum().item()))
    elif wl == "broadphase":
        n = 1 << 20
        kinds, params, poses, _ = scenes.colliders(n, seed=2)
        shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) for k, p in zip(kinds, params)])
        ids = torch.arange(n, dtype=torch.int32, device="cuda")
        dposes = torch.from_numpy(poses).cuda()
        aabbs = shapes.compute_aabbs(ids, dposes)
        bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
        st = {}

        def frame():
            a = shapes.compute_aabbs(ids, dposes)
            bvh.insert_or_update_partially(a, ids, 0.0)
            bvh.rebuild()
            st["p"] = bvh.traverse_bvtt_single_tree(capacity=16 * n, like=a)
        ms = timeit(frame, wl)
        print("%.1f MAABB/s, %d pairs" % (n / ms / 1e3, st["p"].shape[0]))
    elif wl == "colliders_queries":
        # Bvh::cast_ray over typed leaves: 2^20 rays against 2^20 colliders, a third of them 32-vertex hulls
        n = 1 << 20
        g = scenes.rng(12)
        pts, _ = scenes.hull_pool(256, 32, seed=13)
        kinds = g.integers(0, 3, n)
        base = [parry_b200.Ball(0.4), parry_b200.Cuboid([0.3, 0.5, 0.4])] + [parry_b200.ConvexPolyhedron(np.asarray(p, np.float32) * 0.6) for p in pts]
        shapes = parry_b200.Shapes(ctx, base)
        sid = np.where(kinds == 0, 0, np.where(kinds == 1, 1, 2 + g.integers(0, len(pts), n))).astype(np.int32)
        side = (n ** (1 / 3)) * 1.6
        poses = np.concatenate([scenes.random_unit_quaternions(g, n), g.random((n, 3)) * side], axis=1).astype(np.float32)
        dsid, dposes = torch.from_numpy(sid).cuda(), torch.from_numpy(poses).cuda()
        aabbs = shapes.compute_aabbs(dsid, dposes)
        bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
        o = g.random((n, 3)) * side
        d = g.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
        rays = torch.from_numpy(np.concatenate([o, d], axis=1).astype(np.float32)).cuda()
        res = {}
        for with_normal in (False, True):
            ms = timeit(lambda: res.__setitem__("r", bvh.cast_ray(shapes, dsid, dposes, rays, FMAX, solid=True, with_normal=with_normal)),
                        "Bvh::cast_ray typed leaves, normal=%s" % with_normal)
            print("   %.1f M rays/s, %d hits, checksum %.6f" % (n / ms / 1e3, int((res["r"][1] != -1).sum().item()), float(res["r"][0].double().sum().item())))
        q = torch.from_numpy((g.random((n, 3)) * side).astype(np.float32)).cuda()
        ms = timeit(lambda: res.__setitem__("p", bvh.project_point(shapes, dsid, dposes, q, FMAX, solid=True)), "Bvh::project_point typed leaves")
        print("   %.1f M points/s" % (n / ms / 1e3))
    elif wl == "mesh_queries":
        # the composite-shape queries against a 2 M-triangle terrain: shapes scattered within a few shape sizes of the surface
        v, i = scenes.terrain(1001, 1001)
        mesh = parry_b200.TriMesh(ctx, v, i)
        g = scenes.rng(9)
        pts, _ = scenes.hull_pool(64, 32, seed=10)
        spec = [parry_b200.Ball(0.4), parry_b200.Cuboid([0.3, 0.5, 0.4])] + [parry_b200.ConvexPolyhedron(np.asarray(p, np.float32) * 0.6) for p in pts]
        G = parry_b200.Shapes(ctx, spec)
        n = 1 << 20
        sid = g.integers(0, len(spec), n).astype(np.int32)
        anchor = np.asarray(v)[g.integers(0, len(v), n)]
        t = anchor + np.stack([g.standard_normal(n) * 0.3, g.random(n) * 4.0 - 0.3, g.standard_normal(n) * 0.3], axis=1)
        poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
        vel = np.stack([g.standard_normal(n) * 0.5, -(g.random(n) * 2.0 + 0.2), g.standard_normal(n) * 0.5], axis=1).astype(np.float32)
        dsid, dposes, dvel = torch.from_numpy(sid).cuda(), torch.from_numpy(poses).cuda(), torch.from_numpy(vel).cuda()
        ident = torch.tensor([0, 0, 0, 1, 0, 0, 0], dtype=torch.float32, device="cuda")
        zero = torch.zeros(3, dtype=torch.float32, device="cuda")
        res = {}
        ms = timeit(lambda: res.__setitem__("c", mesh.contact_shapes(ident, G, dsid, dposes, 0.05)), "trimesh contact_shapes")
        print("   %.1f M queries/s, %d contacts" % (n / ms / 1e3, int((res["c"][1] == 1).sum().item())))
        ms = timeit(lambda: res.__setitem__("s", mesh.cast_shapes(ident, zero, G, dsid, dposes, dvel)), "trimesh cast_shapes")
        print("   %.1f M queries/s, %d hits" % (n / ms / 1e3, int((res["s"][1] != 0).sum().item())))
        ms = timeit(lambda: res.__setitem__("d", mesh.distance_shapes(ident, G, dsid, dposes)), "trimesh distance_shapes")
        print("   %.1f M queries/s, %d apart" % (n / ms / 1e3, int((res["d"][0] > 0).sum().item())))
        q = torch.from_numpy((t + g.standard_normal((n, 3)) * 0.5).astype(np.float32)).cuda()
        ms = timeit(lambda: res.__setitem__("p", mesh.project_local_point(q)), "trimesh project_point")
        print("   %.1f M points/s" % (n / ms / 1e3))


if __name__ == "__main__":
    main()

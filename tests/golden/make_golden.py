#!/usr/bin/env python
"""Generates tests/golden/*.npz from the CPU oracle AFTER it has been pinned against the reference's own known
answers (tests/test_oracle_kats.py). parry3d cannot be executed here (no Rust toolchain, nalgebra not vendored), so
these vectors freeze the pinned oracle's outputs on small seeded inputs; they guard both the oracle and the GPU path
against silent drift. Re-run: python tests/golden/make_golden.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from harness import oracle, scenes  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))
FMAX = float(np.finfo(np.float32).max)


def main():
    oracle.build()
    # rays vs a small sphere mesh
    v, i = scenes.uv_sphere(24, 16)
    rays = scenes.sphere_rays(2000, seed=101)
    m = oracle.TriMesh(v, i)
    toi, tri, n, f = m.cast_rays(None, rays, FMAX, with_normal=True)
    np.savez_compressed(os.path.join(HERE, "rays_sphere24x16.npz"), vertices=v, indices=i, rays=rays, toi=toi, tri=tri, normal=n, feature=f)
    # broad-phase pairs
    kinds, params, poses, _ = scenes.colliders(1500, seed=102)
    aabbs = oracle.shape_aabbs(kinds, params, poses)
    b = oracle.Bvh(aabbs)
    pairs = b.self_pairs()
    p = np.sort(pairs.astype(np.int64), axis=1)
    order = np.lexsort((p[:, 1], p[:, 0]))
    np.savez_compressed(os.path.join(HERE, "pairs_1500.npz"), kinds=kinds, params=params, poses=poses, aabbs=aabbs, pairs=p[order].astype(np.uint32))
    # contacts over all dispatch arms
    g = scenes.rng(103)
    pts, _ = scenes.hull_pool(6, 32, seed=104)
    spec = [("ball", float(r)) for r in g.random(4) * 0.5 + 0.3] + [("cuboid", list(h)) for h in g.random((4, 3)) * 0.6 + 0.2] + [("convex", q) for q in pts]
    T = oracle.ShapeTable(spec)
    n = 3000
    a = g.integers(0, len(spec), n).astype(np.uint32)
    bb = g.integers(0, len(spec), n).astype(np.uint32)
    q1, q2 = scenes.random_unit_quaternions(g, n), scenes.random_unit_quaternions(g, n)
    t1 = (g.random((n, 3)) - 0.5) * 10
    t2 = t1 + g.standard_normal((n, 3)) * 0.9
    p1 = np.concatenate([q1, t1], 1).astype(np.float32)
    p2 = np.concatenate([q2, t2], 1).astype(np.float32)
    out, st = T.contact(a, p1, bb, p2, 0.02)
    np.savez_compressed(os.path.join(HERE, "contacts_mixed_3000.npz"), kinds=T.kinds, params=T.params, points=T.points, shape1=a, shape2=bb,
                        pos1=p1, pos2=p2, prediction=np.float32(0.02), contacts=out, status=st)
    more()
    print("golden vectors written")


def more():
    """Later additions (culling ray casts, distance / intersection_test, TriMesh-vs-shape contacts); separate files so that the
    first set stays byte-identical."""
    z = np.load(os.path.join(HERE, "rays_sphere24x16.npz"))
    m = oracle.TriMesh(z["vertices"], z["indices"])
    g = scenes.rng(105)
    o = (g.random((2000, 3)) - 0.5) * 2.4
    rays = np.concatenate([o, g.standard_normal((2000, 3))], axis=1).astype(np.float32)
    res = {}
    for mode, name in ((1, "ignore_backfaces"), (2, "ignore_frontfaces")):
        toi, tri, n, f = m.cast_rays(None, rays, FMAX, with_normal=True, mode=1 + mode)
        res.update({name + "_toi": toi, name + "_tri": tri, name + "_normal": n, name + "_feature": f})
    np.savez_compressed(os.path.join(HERE, "rays_culling_sphere24x16.npz"), rays=rays, **res)
    z = np.load(os.path.join(HERE, "contacts_mixed_3000.npz"))
    T = oracle.ShapeTable([])
    T.kinds, T.params, T.points = z["kinds"].copy(), z["params"].copy(), np.ascontiguousarray(z["points"])
    d, ds = T.distance(z["shape1"], z["pos1"], z["shape2"], z["pos2"])
    h, hs = T.intersection_test(z["shape1"], z["pos1"], z["shape2"], z["pos2"])
    np.savez_compressed(os.path.join(HERE, "distance_mixed_3000.npz"), dist=d, dist_status=ds, hit=h, hit_status=hs)
    # TriMesh-vs-shape contacts on a small terrain
    v, i = scenes.terrain(17, 17, extent=12.0)
    v = v.copy(); v[:, 1] *= 0.1
    pts, _ = scenes.hull_pool(4, 32, seed=106)
    g = scenes.rng(107)
    spec = [("ball", float(r)) for r in g.random(3) * 0.4 + 0.2] + [("cuboid", list(h_)) for h_ in g.random((3, 3)) * 0.4 + 0.15]
    spec += [("convex", np.asarray(q, np.float32) * 0.5) for q in pts]
    T2 = oracle.ShapeTable(spec)
    n = 1500
    sid = g.integers(0, len(spec), n).astype(np.uint32)
    t = v[g.integers(0, len(v), n)] + np.stack([g.standard_normal(n) * 0.2, (g.random(n) - 0.35) * 1.2, g.standard_normal(n) * 0.2], axis=1)
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    mesh_pose = np.array([0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0], np.float32)
    om = oracle.TriMesh(v, i)
    out, st, part = om.contact_shapes(mesh_pose, T2, sid, poses, 0.05)
    out2, st2, part2 = om.contact_shapes(mesh_pose, T2, sid, poses, 0.05, min_index_ties=True)
    np.savez_compressed(os.path.join(HERE, "mesh_contacts_1500.npz"), vertices=v, indices=i, kinds=T2.kinds, params=T2.params, points=T2.points,
                        shape=sid, poses=poses, mesh_pose=mesh_pose, prediction=np.float32(0.05), contacts=out, status=st, part=part,
                        contacts_min_index=out2, part_min_index=part2)


def shape_rays():
    """Bvh::cast_ray over ball / cuboid / ConvexPolyhedron leaves (solid and non-solid), 400 colliders, 3000 rays."""
    n, H = 400, 16
    g = scenes.rng(201)
    hulls, _ = scenes.hull_pool(H, 16, seed=202)
    hulls = (hulls * 0.5).astype(np.float32)
    kinds = g.integers(0, 3, n).astype(np.uint8)
    params = (g.random((n, 3)) * 0.3 + 0.15).astype(np.float32)
    hid = g.integers(0, H, n)
    side = (n ** (1 / 3)) * 1.2
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), g.random((n, 3)) * side], axis=1).astype(np.float32)
    points = np.concatenate([hulls[h] for h in hid]).astype(np.float32)
    first = (np.arange(n) * 16).astype(np.uint32)
    count = np.full(n, 16, np.uint32)
    aabbs = oracle.shape_aabbs(kinds, params, poses, points, first, count)
    ob = oracle.Bvh(aabbs)
    d = g.standard_normal((3000, 3))
    d[::2] /= np.linalg.norm(d[::2], axis=1, keepdims=True)
    rays = np.concatenate([g.random((3000, 3)) * side, d], axis=1).astype(np.float32)
    res = {}
    for solid in (True, False):
        toi, leaf, nrm, feat = ob.cast_rays_shapes(kinds, params, poses, rays, FMAX, solid=solid, with_normal=True, points=points, first=first,
                                                   count=count)
        tag = "solid" if solid else "hollow"
        res.update({tag + "_toi": toi, tag + "_leaf": leaf, tag + "_normal": nrm, tag + "_feature": feat})
    np.savez_compressed(os.path.join(HERE, "rays_shapes_400.npz"), kinds=kinds, params=params, poses=poses, points=points, first=first, count=count,
                        aabbs=aabbs, rays=rays, **res)


if __name__ == "__main__":
    import sys
    if "shape_rays" in sys.argv:
        shape_rays()
    else:
        main()

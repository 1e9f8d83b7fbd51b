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


def distance_golden():
    """distance / intersection_test on the contact golden's pairs. Regenerated in round 2 (`python make_golden.py distance`) when the
    oracle gained the SAT-based cuboid-cuboid arms (distance_cuboid_cuboid.rs, intersection_test_cuboid_cuboid.rs): those pairs used
    to be recorded as status 3 (host)."""
    z = np.load(os.path.join(HERE, "contacts_mixed_3000.npz"))
    T = oracle.ShapeTable([])
    T.kinds, T.params, T.points = z["kinds"].copy(), z["params"].copy(), np.ascontiguousarray(z["points"])
    d, ds = T.distance(z["shape1"], z["pos1"], z["shape2"], z["pos2"])
    h, hs = T.intersection_test(z["shape1"], z["pos1"], z["shape2"], z["pos2"])
    np.savez_compressed(os.path.join(HERE, "distance_mixed_3000.npz"), dist=d, dist_status=ds, hit=h, hit_status=hs)


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
    distance_golden()
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


def siblings():
    """cast_shapes (two option sets), Compound contacts (both orders) and ball / cuboid contact manifolds on one mixed scene."""
    g = scenes.rng(301)
    pts, _ = scenes.hull_pool(6, 16, seed=302)
    spec = [("ball", 0.4), ("ball", 0.25), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2])] + [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    T = oracle.ShapeTable(spec)
    ns, n = len(spec), 3000
    s1, s2 = g.integers(0, ns, n).astype(np.uint32), g.integers(0, ns, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - 0.5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sep = np.where(g.random(n) < 0.25, g.random(n) * 0.9, 1.0 + g.random(n) * 3.0)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * sep[:, None]], axis=1).astype(np.float32)
    v1 = (d * (0.5 + g.random((n, 1)) * 3.0) + g.standard_normal((n, 3)) * 0.5).astype(np.float32)
    v2 = (g.standard_normal((n, 3)) * 0.3).astype(np.float32)
    out = dict(kinds=T.kinds, params=T.params, points=T.points, shape1=s1, shape2=s2, pos1=p1, pos2=p2, vel1=v1, vel2=v2)
    o, st = T.cast_shapes(s1, p1, v1, s2, p2, v2)
    out.update(cast_default=o, cast_default_status=st)
    o, st = T.cast_shapes(s1, p1, v1, s2, p2, v2, target_distance=0.05, stop_at_penetration=False)
    out.update(cast_target=o, cast_target_status=st)
    # compounds: 24 of 1-5 parts
    first, count, psid, ppose = [], [], [], []
    for c in range(24):
        k = int(g.integers(1, 6))
        first.append(len(psid)); count.append(k)
        psid += [int(x) for x in g.integers(0, ns, k)]
        ppose.append(np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.6], axis=1))
    first, count, psid = np.asarray(first, np.uint32), np.asarray(count, np.uint32), np.asarray(psid, np.uint32)
    ppose = np.concatenate(ppose).astype(np.float32)
    cid = g.integers(0, 24, n).astype(np.uint32)
    p2c = p2.copy()
    p2c[:, 4:] = p1[:, 4:] + d * (g.random((n, 1)) * 2.2 + 0.1)
    out.update(comp_first=first, comp_count=count, part_shape=psid, part_pose=ppose, compound_id=cid, pos2_compound=p2c)
    for second in (False, True):
        o, st, part = T.contact_compound(first, count, psid, ppose, cid, p1, s2, p2c, 0.05, compound_second=second)
        tag = "compound_second" if second else "compound_first"
        out.update({tag: o, tag + "_status": st, tag + "_part": part})
    # manifolds on the ball / cuboid subset of the pair list, closer together
    m1, m2 = (s1 % 4).astype(np.uint32), (s2 % 4).astype(np.uint32)
    p2m = p2.copy()
    p2m[:, 4:] = p1[:, 4:] + d * (g.random((n, 1)) * 1.2 + 0.2)
    p2m[::4, :4] = p1[::4, :4]
    nr, cnt, mp, st = T.contact_manifolds(m1, p1, m2, p2m, 0.05, max_points=8)
    out.update(man_shape1=m1, man_shape2=m2, man_pos2=p2m, man_normals=nr, man_counts=cnt, man_points=mp, man_status=st)
    np.savez_compressed(os.path.join(HERE, "siblings_3000.npz"), **out)


def pfm():
    """pfm_pfm contact manifolds: cuboids, hulls of cuboid corners and 16-point hulls, with their face topology."""
    g = scenes.rng(401)
    pts, _ = scenes.hull_pool(8, 16, seed=402)
    hes = [np.array([0.3, 0.5, 0.4], np.float32), np.array([0.6, 0.2, 0.2], np.float32)]
    corners = lambda he: np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    spec = [("cuboid", hes[0]), ("cuboid", hes[1]), ("convex", corners(hes[0])), ("convex", corners(hes[1]))]
    spec += [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    T = oracle.ShapeTable(spec)
    topo = T.hull_topology()
    n, ns = 2000, len(spec)
    s1, s2 = g.integers(0, ns, n).astype(np.uint32), g.integers(2, ns, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.2 + 0.2)], axis=1).astype(np.float32)
    p2[::4, :4] = p1[::4, :4]
    nr, cnt, mp, st = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=12, topology=topo)
    np.savez_compressed(os.path.join(HERE, "manifolds_pfm_2000.npz"), kinds=T.kinds, params=T.params, points=T.points, shape1=s1, shape2=s2, pos1=p1,
                        pos2=p2, normals=nr, counts=cnt, man_points=mp, status=st, **{"topo_" + k: v for k, v in topo.items()})


def second_frame():
    """Second-frame manifolds (persistent dispatch, match_contacts) on the pfm golden's scene and first-frame manifolds, and Compound
    vs Compound contacts on the sibling golden's compounds; inputs are read from those two files, only the new poses / ids and
    the oracle's outputs are stored."""
    z = np.load(os.path.join(HERE, "manifolds_pfm_2000.npz"))
    T = oracle.ShapeTable([])
    T.kinds, T.params, T.points = z["kinds"].copy(), z["params"].copy(), np.ascontiguousarray(z["points"])
    topo = {k[5:]: z[k] for k in z.files if k.startswith("topo_")}
    g = scenes.rng(501)
    n = len(z["shape1"])
    moved = z["pos2"].copy()
    moved[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.5, 2.0e-4, 0.05)).astype(np.float32)
    rn, rc, rp, rs, rk, rm = T.contact_manifolds_update(z["shape1"], z["pos1"], z["shape2"], moved, 0.05, z["normals"], z["counts"], z["man_points"],
                                                        topology=topo, seed_gjk=True)   # round 2: the reference's GJK seed (pfm_pfm.rs:63-66)
    y = np.load(os.path.join(HERE, "siblings_3000.npz"))
    T2 = oracle.ShapeTable([])
    T2.kinds, T2.params, T2.points = y["kinds"].copy(), y["params"].copy(), np.ascontiguousarray(y["points"])
    ids2 = g.integers(0, len(y["comp_first"]), len(y["compound_id"])).astype(np.uint32)
    co, cs, cp = T2.contact_compound_compound(y["comp_first"], y["comp_count"], y["part_shape"], y["part_pose"], y["compound_id"], y["pos1"], ids2,
                                              y["pos2_compound"], 0.05)
    np.savez_compressed(os.path.join(HERE, "second_frame_2000.npz"), moved_pos2=moved, normals=rn, counts=rc, man_points=rp, status=rs, kept=rk,
                        match=rm, compound_id2=ids2, cc_contacts=co, cc_status=cs, cc_parts=cp)


if __name__ == "__main__":
    import sys
    if "distance" in sys.argv:
        distance_golden()
    elif "second_frame" in sys.argv:
        second_frame()
    elif "pfm" in sys.argv:
        pfm()
    elif "shape_rays" in sys.argv:
        shape_rays()
    elif "siblings" in sys.argv:
        siblings()
    else:
        main()

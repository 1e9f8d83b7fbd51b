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
    print("golden vectors written")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Diagnostics for pb2_trimesh_cast_shapes vs the oracle: how often the winning triangle differs and whether those are exact ties.
python harness/mesh_cast_check.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import parry_b200
    from harness import oracle
    import test_zzzz_mesh_cast_gpu as t
    oracle.build()
    ctx = parry_b200.Context(0)
    for seed, n, lift, second, opts in [(401, 6000, (0.5, 6.0), False, {}), (403, 3000, (0.5, 6.0), True, {}),
                                        (404, 3000, (0.5, 6.0), False, {"max_toi": 1.5, "target_distance": 0.1})]:
        v, idx, spec, sid, poses, vel = t.make_scene(n, seed, lift)
        T, G = t.tables(ctx, oracle, spec)
        gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
        mpose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
        mvel = np.zeros(3, np.float32)
        o = t.oracle_casts(om, T, mpose, mvel, sid, poses, vel, second, **opts)
        go = parry_b200.ShapeCastOptions(max_time_of_impact=opts.get("max_toi", t.FMAX), target_distance=opts.get("target_distance", 0.0))
        g = [np.asarray(x) for x in gm.cast_shapes(mpose, mvel, G, sid, poses, vel, go, mesh_second=second)]
        some = (g[1] == o[1]) & (o[1] != 0)
        same = g[2].astype(np.uint32)[some] == o[2][some]
        gt, ot = g[0][some][:, 12], o[0][some][:, 12]
        kinds = np.array([0 if k == "ball" else 1 if k == "cuboid" else 2 for k, _ in spec])[sid][some]
        print("seed", seed, "hits", some.sum(), "same triangle %.3f" % same.mean(), "status mismatches", (g[1] != o[1]).sum())
        print("   differing: toi bit-equal %.3f, gpu smaller %.3f, oracle smaller %.3f; max rel diff %.2e" % (
            (gt[~same] == ot[~same]).mean(), (gt[~same] < ot[~same]).mean(), (gt[~same] > ot[~same]).mean(),
            np.max(np.abs(gt - ot) / np.maximum(1e-6, np.abs(ot)))))
        for kk, name in enumerate(("ball", "cuboid", "hull")):
            m = kinds == kk
            print("   %s: same triangle %.3f of %d" % (name, same[m].mean(), m.sum()))
        w = slice(0, 3) if not second else slice(3, 6)
        d = np.abs(g[0][some][~same][:, w] - o[0][some][~same][:, w]).max(axis=1)
        print("   differing: witness on the mesh within 1e-3: %.3f" % (d < 1e-3).mean())


if __name__ == "__main__":
    main()

"""GPU parity of query::cast_shapes with a TriMesh on one side (the composite arms of
DefaultQueryDispatcher::cast_shapes, default_query_dispatcher.rs:498-515 -> shape_cast_composite_shape_shape.rs:14-105) through
pb2_trimesh_cast_shapes against the CPU oracle, whose restatement is pinned by the reference's own trimesh_trimesh_toi.rs
(tests/test_oracle_kats.py::test_trimesh_trimesh_toi_issue_194): outcome and status exact, time of impact within 1e-5, the hit
triangle exact but for equal times, witnesses / normals within 1e-4 where the triangle agrees."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)


def make_scene(n, seed, lift):
    g = scenes.rng(seed)
    v, idx = scenes.terrain(65, 65, extent=40.0)
    v = v.copy()
    v[:, 1] *= 0.2
    pts, _ = scenes.hull_pool(8, 24, seed=seed + 1)
    spec = [("ball", 0.35), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])]
    spec += [("convex", np.asarray(p, np.float32) * 0.5) for p in pts]
    sid = g.integers(0, len(spec), n).astype(np.uint32)
    anchor = v[g.integers(0, len(v), n)]
    t = anchor + np.stack([g.standard_normal(n) * 2.0, lift[0] + g.random(n) * (lift[1] - lift[0]), g.standard_normal(n) * 2.0], axis=1)
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    vel = np.stack([g.standard_normal(n) * 0.6, -(g.random(n) * 2.0 + 0.2), g.standard_normal(n) * 0.6], axis=1).astype(np.float32)
    vel[g.random(n) < 0.1] *= -1.0                      # some move away
    return v, idx, spec, sid, poses, vel


def tables(ctx, oracle, spec):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(p) if k == "ball" else parry_b200.Cuboid(p) if k == "cuboid" else parry_b200.ConvexPolyhedron(p)
                                for k, p in spec])
    return T, G


def oracle_casts(om, T, mpose, mvel, sid, poses, vel, mesh_second, **opts):
    n = len(sid)
    out = np.zeros((n, 13), np.float32)
    st = np.zeros(n, np.uint8)
    part = np.full(n, 0xFFFFFFFF, np.uint32)
    for k in range(n):
        r = om.cast_shapes(mpose, mvel, poses[k], vel[k], table=T, shape=int(sid[k]), mesh_second=mesh_second, **opts)
        if r is not None:
            out[k], st[k], part[k] = r[0], r[1], r[2]
    return out, st, part


def check(g, o, min_hits):
    go, gs, gp = (np.asarray(x) for x in g)
    oo, os_, op = o
    gp = gp.astype(np.uint32)
    assert (os_ != 0).mean() > min_hits, (os_ != 0).mean()
    # a time of impact within rounding of max_toi or of the 1e-5 penetration threshold may fall on either side
    differ = np.nonzero(gs != os_)[0]
    assert len(differ) <= max(1, len(gs) // 500), (len(differ), differ[:10], gs[differ][:10], os_[differ][:10])
    ok = gs == os_
    some = ok & (os_ != 0)
    assert (gp[ok & (os_ == 0)] == 0xFFFFFFFF).all() and (go[ok & (os_ == 0)] == 0).all()
    np.testing.assert_allclose(go[some][:, 12], oo[some][:, 12], rtol=1e-5, atol=2e-6)
    # Equal times of impact are common on a mesh — a shape that starts in touch with several triangles hits all of them at time 0, an
    # impact on a shared edge is one event for both triangles (measured with harness/mesh_cast_check.py: 18-35 % of the hits here, all
    # of them bit-equal times) — and resolve to the reference's first-in-tree-order in the oracle, to the smallest index here.
    same = gp[some] == op[some]
    assert same.mean() > 0.5, same.mean()
    assert (go[some][~same][:, 12] == oo[some][~same][:, 12]).all()
    a, b = go[some], oo[some]
    rows_ok = (np.abs(a[:, :12] - b[:, :12]) < 1e-4).all(axis=1)
    assert rows_ok[same].mean() > 0.99, rows_ok[same].mean()
    return some, same, rows_ok


@pytest.mark.parametrize("seed,n,lift,mesh_second,opts", [
    (401, 6000, (0.5, 6.0), False, {}),                                   # falling onto the terrain
    (402, 3000, (-0.3, 1.0), False, {}),                                   # many start touching / penetrating: impact geometry from EPA
    (403, 3000, (0.5, 6.0), True, {}),                                     # the shape is shape 1: swapped hit
    (404, 3000, (0.5, 6.0), False, {"max_toi": 1.5, "target_distance": 0.1}),
    (405, 2000, (-0.3, 1.0), True, {"compute_impact_geometry_on_penetration": False}),
])
def test_trimesh_cast_shapes_vs_oracle(ctx, oracle, seed, n, lift, mesh_second, opts):
    import parry_b200
    v, idx, spec, sid, poses, vel = make_scene(n, seed, lift)
    T, G = tables(ctx, oracle, spec)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    mq = np.array([0.01, -0.02, 0.015, 1.0]); mq /= np.linalg.norm(mq)
    mpose = np.concatenate([mq, [0.05, -0.1, 0.08]]).astype(np.float32)
    mvel = np.array([0.1, 0.05, -0.08], np.float32)
    o = oracle_casts(om, T, mpose, mvel, sid, poses, vel, mesh_second, **opts)
    go = parry_b200.ShapeCastOptions(max_time_of_impact=opts.get("max_toi", FMAX), target_distance=opts.get("target_distance", 0.0),
                                     compute_impact_geometry_on_penetration=opts.get("compute_impact_geometry_on_penetration", True))
    g = gm.cast_shapes(mpose, mvel, G, sid, poses, vel, go, mesh_second=mesh_second)
    some, same, rows_ok = check(g, o, 0.2)
    if lift[0] < 0:
        assert (o[1][some] == 2).mean() > 0.1               # PenetratingOrWithinTargetDist is exercised


def test_trimesh_cast_shapes_edge_cases(ctx, oracle):
    import torch
    import parry_b200
    v, idx, spec, sid, poses, vel = make_scene(500, 406, (0.5, 6.0))
    T, G = tables(ctx, oracle, spec)
    gm = parry_b200.TriMesh(ctx, v, idx)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    zero = np.zeros(3, np.float32)
    h = gm.cast_shapes(ident, zero, G, sid, poses, vel)
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    d = gm.cast_shapes(dev(ident), dev(zero), G, dev(sid), dev(poses), dev(vel))
    ctx.synchronize()
    assert (d[1].cpu().numpy() == h[1]).all() and (d[2].cpu().numpy().view(np.uint32) == np.asarray(h[2]).view(np.uint32)).all()
    assert (d[0].cpu().numpy().view(np.uint32) == np.asarray(h[0]).view(np.uint32)).all()
    # zero relative velocity: only shapes already touching report a hit (toi 0); unknown shape id: status 3
    still = gm.cast_shapes(ident, zero, G, sid, poses, np.zeros_like(vel))
    assert (np.asarray(still[0])[:, 12] == 0).all()
    bad = sid.copy(); bad[3] = 9999
    b = gm.cast_shapes(ident, zero, G, bad, poses, vel)
    assert np.asarray(b[1])[3] == 3 and (np.delete(np.asarray(b[1]), 3) == np.delete(np.asarray(h[1]), 3)).all()
    with pytest.raises(parry_b200.Pb2Error):
        gm.cast_shapes(ident, zero, G, sid, poses, vel, parry_b200.ShapeCastOptions(stop_at_penetration=False))
    # the reference's own composite cast fixture, one side as plain triangles: a pyramid mesh hit by a ball moving along -x
    pts = np.array([[0, 1, 0], [-1, -0.5, 0], [0, -0.5, -1], [1, -0.5, 0]], np.float32)
    pidx = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1]], np.uint32)
    pm, po = parry_b200.TriMesh(ctx, pts, pidx), oracle.TriMesh(pts, pidx)
    pose = np.array([[0, 0, 0, 1, 10.0, 0.0, 0.0]], np.float32)
    r = po.cast_shapes(ident, zero, pose[0], np.array([-2.0, 0, 0], np.float32), table=T, shape=0)
    gq = pm.cast_shapes(ident, zero, G, np.array([0], np.uint32), pose, np.array([[-2.0, 0, 0]], np.float32))
    assert r is not None and np.asarray(gq[1])[0] == r[1]
    np.testing.assert_allclose(np.asarray(gq[0])[0], r[0], rtol=1e-5, atol=1e-6)


def pyramid():
    pts = np.array([[0, 1, 0], [-1, -0.5, 0], [0, -0.5, -1], [1, -0.5, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1]], np.uint32)
    return pts, idx


def test_trimesh_trimesh_toi_issue_194_on_the_gpu(ctx, oracle):
    """crates/parry3d/tests/geometry/trimesh_trimesh_toi.rs through pb2_trimesh_cast_trimesh: two pyramids 1000 apart, one moving at
    100000 along x: `assert_eq!(time_of_impact, Some(0.00998))`, exact; and the opposite direction misses."""
    import parry_b200
    pts, idx = pyramid()
    a, b = parry_b200.TriMesh(ctx, pts, idx), parry_b200.TriMesh(ctx, pts, idx)
    p1 = np.array([[0, 0, 0, 1, 0, 0, 0]] * 2, np.float32)
    p2 = np.array([[0, 0, 0, 1, 1000.0, 0, 0]] * 2, np.float32)
    v1 = np.array([[100000.0, 0, 0], [-100000.0, 0, 0]], np.float32)
    v2 = np.zeros((2, 3), np.float32)
    out, st, parts = (np.asarray(x) for x in a.cast_trimesh(p1, v1, b, p2, v2))
    assert st[0] == 1 and out[0, 12] == np.float32(0.00998)
    assert (out[0, :6] == np.array([1, -0.5, 0, -1, -0.5, 0], np.float32)).all()          # the two base corners that meet
    assert st[1] == 0 and (parts[1] == 0xFFFFFFFF).all()
    oa, ob = oracle.TriMesh(pts, idx), oracle.TriMesh(pts, idx)
    r = oa.cast_shapes(p1[0], v1[0], p2[0], v2[0], other_mesh=ob)
    assert r is not None and (r[0].view(np.uint32) == out[0].view(np.uint32)).all()


@pytest.mark.parametrize("lift,seed", [((2.6, 6.5), 411), ((0.9, 3.0), 412)])
def test_trimesh_cast_trimesh_vs_oracle(ctx, oracle, lift, seed):
    """Random poses / velocities of two small meshes (a pyramid against a patch of terrain) against the oracle's nested descent; with
    the lower lift a good part of the pairs start in touch and take their geometry from the triangle-triangle contact."""
    import parry_b200
    g = scenes.rng(seed)
    pts, idx = pyramid()
    v, tidx = scenes.terrain(17, 17, extent=12.0)
    v = v.copy()
    v[:, 1] *= 0.2
    ga, gb = parry_b200.TriMesh(ctx, pts, idx), parry_b200.TriMesh(ctx, v, tidx)
    oa, ob = oracle.TriMesh(pts, idx), oracle.TriMesh(v, tidx)
    n = 600
    lo, hi = v.min(axis=0), v.max(axis=0)
    anchor = v[g.integers(0, len(v), n)]
    t = anchor + np.stack([g.standard_normal(n) * 0.3, lift[0] + g.random(n) * (lift[1] - lift[0]), g.standard_normal(n) * 0.3], axis=1)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    p2 = np.tile(np.array([0.02, -0.01, 0.03, 1.0, 0.1, 0.0, -0.1], np.float32), (n, 1))
    p2[:, :4] /= np.linalg.norm(p2[:, :4], axis=1, keepdims=True)
    v1 = np.stack([g.standard_normal(n) * 0.5, -(g.random(n) * 2.0 + 0.3), g.standard_normal(n) * 0.5], axis=1).astype(np.float32)
    v1[g.random(n) < 0.15] *= -1.0
    v2 = (g.standard_normal((n, 3)) * 0.1).astype(np.float32)
    for first, second, pa, va, pb, vb, fa, fb in ((ga, gb, p1, v1, p2, v2, oa, ob), (gb, ga, p2, v2, p1, v1, ob, oa)):
        out, st, parts = (np.asarray(x) for x in first.cast_trimesh(pa, va, second, pb, vb))
        oo = np.zeros((n, 13), np.float32); os_ = np.zeros(n, np.uint8); op = np.full(n, 0xFFFFFFFF, np.uint32)
        for k in range(n):
            r = fa.cast_shapes(pa[k], va[k], pb[k], vb[k], other_mesh=fb)
            if r is not None:
                oo[k], os_[k], op[k] = r[0], r[1], r[2]
        assert 0.3 < (os_ != 0).mean() < 0.98
        if lift[0] < 2:
            assert (os_ == 2).mean() > 0.1
        assert (st == os_).all(), np.nonzero(st != os_)[0][:10]
        hit = os_ != 0
        assert (out[hit][:, 12] == oo[hit][:, 12]).mean() > 0.99
        np.testing.assert_allclose(out[hit][:, 12], oo[hit][:, 12], rtol=1e-5, atol=2e-6)
        same = parts[hit][:, 0] == op[hit]                     # (a pyramid's edges and apex belong to 2-3 triangles: ties, see check())
        assert same.mean() > (0.5 if lift[0] > 2 else 0.25)
        assert (out[hit][~same][:, 12] == oo[hit][~same][:, 12]).mean() > 0.98
        rows_ok = (np.abs(out[hit][same][:, :12] - oo[hit][same][:, :12]) < 1e-4).all(axis=1)
        # (the oracle reports the triangle of the first mesh only: an equal-time pair with another triangle of the second mesh passes
        # `same` with different geometry)
        assert rows_ok.mean() > (0.9 if lift[0] > 2 else 0.6), rows_ok.mean()


@pytest.mark.parametrize("mesh_second", [False, True])
def test_trimesh_distance_vs_oracle(ctx, oracle, mesh_second):
    """query::distance with a TriMesh on one side through pb2_trimesh_distance_shapes: the distance is a pure minimum (no tie
    can change it), so it must be bit-identical to the oracle's descent of the reference's own tree."""
    import parry_b200
    v, idx, spec, sid, poses, _ = make_scene(8000, 421, (-0.5, 6.0))
    T, G = tables(ctx, oracle, spec)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    mq = np.array([0.01, -0.02, 0.015, 1.0]); mq /= np.linalg.norm(mq)
    mpose = np.concatenate([mq, [0.05, -0.1, 0.08]]).astype(np.float32)
    od, op = om.distance_shapes(mpose, T, sid, poses, mesh_second=mesh_second, threads=8)
    gd, gs, gp = (np.asarray(x) for x in gm.distance_shapes(mpose, G, sid, poses, mesh_second=mesh_second))
    assert (gs == 0).all()
    assert 0.05 < (od == 0).mean() < 0.6
    assert (gd.view(np.uint32) == od.view(np.uint32)).all(), np.nonzero(gd != od)[0][:10]
    pos = od > 0
    # (query::distance returns no part; the closest point of a convex shape over a terrain is on a shared edge or vertex about half of
    # the time, where 2-6 triangles are equally close: first in the reference's tree order there, smallest index here)
    assert (gp.astype(np.uint32)[pos] == op[pos]).mean() > 0.4
    for lo, hi in ((0, 2), (2, 4), (4, 12)):                          # every arm is exercised
        m = (sid >= lo) & (sid < hi)
        assert (od[m] > 0).mean() > 0.3
    bad = sid[:4].copy(); bad[1] = 9999
    b = gm.distance_shapes(mpose, G, bad, poses[:4], mesh_second=mesh_second)
    assert np.asarray(b[1])[1] == 2 and (np.asarray(b[0])[[0, 2, 3]] == gd[[0, 2, 3]]).all()

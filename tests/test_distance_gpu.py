"""query::distance / query::intersection_test: oracle pins from the reference's examples and doc-tests (CPU) and
GPU parity through the C ABI."""
import numpy as np
import pytest

from harness import scenes

I4 = [0.0, 0.0, 0.0, 1.0]


def P(t):
    return np.array([I4 + list(t)], np.float32)


def test_oracle_pins_from_reference_examples():
    from harness import oracle
    oracle.build()
    T = oracle.ShapeTable([("ball", 1.0), ("cuboid", [1, 1, 1]), ("ball", 2.0), ("cuboid", [0.5, 0.5, 0.5])])
    # examples/distance_query3d.rs
    assert T.distance([0], P([0, 1, 0]), [1], P([0, 0, 0]))[0][0] == 0.0
    assert abs(T.distance([0], P([0, 3, 0]), [1], P([0, 0, 0]))[0][0] - 1.0) <= 1e-7
    # examples/proximity_query3d.rs
    assert T.intersection_test([0], P([1, 1, 1]), [1], P([0, 0, 0]))[0][0] == 1
    assert T.intersection_test([0], P([3, 3, 3]), [1], P([0, 0, 0]))[0][0] == 0
    # distance.rs doc-test: balls r = 1 and r = 2, centres 10 apart => exactly 7
    assert T.distance([0], P([0, 0, 0]), [2], P([10, 0, 0]))[0][0] == np.float32(7.0)
    # intersection_test.rs doc-test: two unit balls 1.5 apart intersect, 5 apart do not
    assert T.intersection_test([0], P([0, 0, 0]), [0], P([1.5, 0, 0]))[0][0] == 1
    assert T.intersection_test([0], P([0, 0, 0]), [0], P([5.0, 0, 0]))[0][0] == 0
    # cuboid-cuboid is the SAT arm (distance_cuboid_cuboid.rs): half extents 1 and 0.5, centres 5 apart on x => 3.5
    d, st = T.distance([1], P([0, 0, 0]), [3], P([5, 0, 0]))
    assert st[0] == 0 and d[0] == np.float32(3.5)
    assert T.intersection_test([1], P([0, 0, 0]), [3], P([1.4, 0, 0]))[0][0] == 1
    assert T.intersection_test([1], P([0, 0, 0]), [3], P([1.6, 0, 0]))[0][0] == 0


def make_pairs(seed, n):
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(12, 32, seed=seed + 1)
    spec = [("ball", float(r)) for r in g.random(6) * 0.5 + 0.3]
    spec += [("cuboid", list(h)) for h in g.random((6, 3)) * 0.6 + 0.2]
    spec += [("convex", p) for p in pts]
    a = g.integers(0, len(spec), n).astype(np.uint32)
    b = g.integers(0, len(spec), n).astype(np.uint32)
    q1, q2 = scenes.random_unit_quaternions(g, n), scenes.random_unit_quaternions(g, n)
    t1 = (g.random((n, 3)) - 0.5) * 10
    t2 = t1 + g.standard_normal((n, 3)) * 1.2
    return spec, a, b, np.concatenate([q1, t1], 1).astype(np.float32), np.concatenate([q2, t2], 1).astype(np.float32)


@pytest.mark.gpu
def test_distance_and_intersection_match_oracle(ctx, oracle):
    import parry_b200
    spec, a, b, p1, p2 = make_pairs(91, 80000)
    gshapes = [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v) for k, v in spec]
    G, O = parry_b200.Shapes(ctx, gshapes), oracle.ShapeTable(spec)
    gd, gds = parry_b200.distance(G, a, p1, b, p2)
    od, ods = O.distance(a, p1, b, p2, threads=8)
    assert (np.asarray(gds) == ods).all()
    ok = ods == 0
    assert ok.all()                                  # no pair of these shapes goes back to the host (round 1: cuboid-cuboid did)
    kinds = np.array([0 if k == "ball" else 1 if k == "cuboid" else 2 for k, _ in spec])
    cc = (kinds[a] == 1) & (kinds[b] == 1)           # the SAT-based arm (distance_cuboid_cuboid.rs, intersection_test_cuboid_cuboid.rs)
    assert 0.05 < cc.mean() < 0.2 and 0.2 < (od[cc] > 0).mean() < 0.9
    assert (np.asarray(gd)[cc].view(np.uint32) == od[cc].view(np.uint32)).mean() > 0.999
    assert 0.2 < (od[ok] > 0).mean() < 0.9
    np.testing.assert_allclose(np.asarray(gd)[ok], od[ok], rtol=1e-5, atol=1e-6)
    assert (np.asarray(gd)[ok].view(np.uint32) == od[ok].view(np.uint32)).mean() > 0.999
    gi, gis = parry_b200.intersection_test(G, a, p1, b, p2)
    oi, ois = O.intersection_test(a, p1, b, p2, threads=8)
    assert (np.asarray(gis) == ois).all()
    assert (np.asarray(gi) == oi).all()              # boolean answers: exact
    ok = ois == 0
    # consistency of the two queries (same GJK, different exit rules): intersecting <=> distance 0 up to grazing cases
    agree = (oi[ok] == 1) == (od[ok] == 0.0)
    assert agree.mean() > 0.999
    # bad shape id
    d, st = parry_b200.distance(G, np.array([999], np.uint32), p1[:1], b[:1], p2[:1])
    assert np.asarray(st)[0] == 2


@pytest.mark.gpu
def test_distance_device_resident(ctx, oracle):
    import torch
    import parry_b200
    spec, a, b, p1, p2 = make_pairs(92, 20000)
    gshapes = [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v) for k, v in spec]
    G = parry_b200.Shapes(ctx, gshapes)
    h = parry_b200.distance(G, a, p1, b, p2)
    T = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    d = parry_b200.distance(G, T(a), T(p1), T(b), T(p2))
    ctx.synchronize()
    assert (d[0].cpu().numpy().view(np.uint32) == h[0].view(np.uint32)).all() and (d[1].cpu().numpy() == h[1]).all()


@pytest.mark.gpu
def test_closest_points_vs_oracle(ctx, oracle):
    """query::closest_points (closest_points_shape_shape.rs:220-231): kinds exact, points 1e-5; the doc examples of that file as
    known answers (balls 0.5 at x = 0 / 12 within 15 -> (0.5, 0, 0), (11.5, 0, 0); ball 2 at x = 5 vs unit cuboid -> (3, 0, 0),
    (1, 0, 0))."""
    import parry_b200
    I = [0, 0, 0, 1]
    pose = lambda t: np.array(I + list(t), np.float32)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(0.5), parry_b200.Ball(2.0), parry_b200.Cuboid([1, 1, 1])])
    o, k, s = parry_b200.closest_points(G, np.array([0, 1], np.uint32), np.stack([pose([0, 0, 0]), pose([5, 0, 0])]), np.array([0, 2], np.uint32),
                                        np.stack([pose([12, 0, 0]), pose([0, 0, 0])]), 15.0)
    assert (k == 1).all() and (s == 1).all()
    np.testing.assert_allclose(o, [[0.5, 0, 0, 11.5, 0, 0], [3, 0, 0, 1, 0, 0]], rtol=0, atol=1e-6)
    g = scenes.rng(5)
    pts, _ = scenes.hull_pool(6, 16, seed=6)
    spec = [("ball", 0.4), ("ball", 0.25), ("cuboid", [0.3, 0.5, 0.4])] + [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(v) if kk == "ball" else parry_b200.Cuboid(v) if kk == "cuboid" else parry_b200.ConvexPolyhedron(v)
                                for kk, v in spec])
    n = 30000
    s1, s2 = g.integers(0, len(spec), n).astype(np.uint32), g.integers(0, len(spec), n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 2.5 + 0.2)], axis=1).astype(np.float32)
    for margin in (0.8, 0.0):
        ro, rk, rs = T.closest_points(s1, p1, s2, p2, margin, threads=8)
        go, gk, gs = parry_b200.closest_points(G, s1, p1, s2, p2, margin)
        assert (rs == 1).all() and (gs == 1).all() and (gk == rk).all()
        assert (np.bincount(rk, minlength=3) > (0 if margin == 0.0 else 2000))[[0, 2]].all()
        np.testing.assert_allclose(go, ro, rtol=1e-5, atol=2e-6)
    bad = s1.copy()
    bad[11] = 10 ** 6
    _, bk, bs = parry_b200.closest_points(G, bad, p1, s2, p2, 0.8)
    assert bs[11] == 2 and bk[11] == 0

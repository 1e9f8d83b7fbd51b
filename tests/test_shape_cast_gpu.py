"""GPU parity of query::cast_shapes through pb2_cast_shapes_batch against the CPU oracle: statuses exact,
time of impact / witnesses / normals within 1e-5; the reference's own shape-cast tests as known answers."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)


def _pose(t):
    return np.array([0, 0, 0, 1] + list(t), np.float32)


def make_table(oracle, ctx):
    import parry_b200
    pts, _ = scenes.hull_pool(16, 16, seed=82)
    spec_o = [("ball", 0.4), ("ball", 0.25), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2])] + [("convex", p) for p in pts]
    spec_g = [parry_b200.Ball(0.4), parry_b200.Ball(0.25), parry_b200.Cuboid([0.3, 0.5, 0.4]), parry_b200.Cuboid([0.6, 0.2, 0.2])]
    spec_g += [parry_b200.ConvexPolyhedron(p) for p in pts]
    return oracle.ShapeTable(spec_o), parry_b200.Shapes(ctx, spec_g), len(spec_o)


def make_pairs(n, n_shapes, seed):
    g = scenes.rng(seed)
    s1, s2 = g.integers(0, n_shapes, n).astype(np.uint32), g.integers(0, n_shapes, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - 0.5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    sep = np.where(g.random(n) < 0.25, g.random(n) * 0.9, 1.0 + g.random(n) * 3.0)   # a quarter start overlapping
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * sep[:, None]], axis=1).astype(np.float32)
    v1 = (d * (0.5 + g.random((n, 1)) * 3.0) + g.standard_normal((n, 3)) * 0.5).astype(np.float32)
    v2 = (g.standard_normal((n, 3)) * 0.3).astype(np.float32)
    v1[::17] = 0.0
    v2[::17] = 0.0     # no relative motion: None
    return s1, p1, v1, s2, p2, v2


def test_reference_shape_cast_tests_on_gpu(ctx):
    import parry_b200
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(0.5), parry_b200.Ball(1.0), parry_b200.Cuboid([1, 1, 1]), parry_b200.Cuboid([.5, .5, .5])])
    s1 = np.array([0, 1, 1, 1, 3, 3, 3], np.uint32)
    s2 = np.array([0, 2, 2, 2, 3, 3, 3], np.uint32)
    p1 = np.stack([_pose([0, 0, 0]), _pose([1, 1, 1]), _pose([2, 2, 2]), _pose([3, 3, 3])] + [_pose([0, 1.1, 0])] * 3)
    p2 = np.stack([_pose([0, 10, 0])] + [_pose([0, 0, 0])] * 6)
    v1 = np.array([[0, 10, 0], [2, 2, 2], [-.5, -.5, -.5], [2, 2, 2], [0, 0, 0], [0, 1, 0], [0, -1, 0]], np.float32)
    v2 = np.array([[0, 0, 0], [-1, 1, 1], [1, 1, 1], [-1, 1, 1], [0, 0, 0], [0, 0, 0], [0, 0, 0]], np.float32)
    out, st = parry_b200.cast_shapes(G, s1, p1, v1, s2, p2, v2)
    assert st[0] == 1 and out[0, 12] == np.float32(0.9)                      # ball_ball_toi.rs
    assert st[1] != 0 and out[1, 12] == 0.0                                   # time_of_impact3.rs
    expect = (np.sqrt(np.float32(3.0)) - np.float32(1.0)) / np.linalg.norm(np.array([-1.5, -1.5, -1.5], np.float32))
    assert st[2] == 1 and abs(out[2, 12] - expect) <= np.finfo(np.float32).eps * abs(expect)
    assert st[3] == 0
    assert st[4] == 0 and st[5] == 0 and st[6] == 1                           # still_objects_toi.rs


@pytest.mark.parametrize("opts", [dict(), dict(stop_at_penetration=False), dict(compute_impact_geometry_on_penetration=False),
                                  dict(target_distance=0.05), dict(max_time_of_impact=0.4),
                                  dict(stop_at_penetration=False, compute_impact_geometry_on_penetration=False, target_distance=0.02)])
def test_cast_shapes_mixed_pairs(ctx, oracle, opts):
    import parry_b200
    T, G, ns = make_table(oracle, ctx)
    s1, p1, v1, s2, p2, v2 = make_pairs(20000, ns, seed=83)
    ro, rs = T.cast_shapes(s1, p1, v1, s2, p2, v2, threads=8, **opts)
    go, gs = parry_b200.cast_shapes(G, s1, p1, v1, s2, p2, v2, parry_b200.ShapeCastOptions(**opts))
    assert (rs == 1).mean() > 0.15 and (rs == 2).mean() > 0.05
    assert (gs == rs).all(), np.nonzero(gs != rs)[0][:10]
    hit = rs != 0
    np.testing.assert_allclose(go[hit, 12], ro[hit, 12], rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(go[hit, :12], ro[hit, :12], rtol=1e-5, atol=2e-6)
    assert (go[~hit] == 0).all()


def test_cast_shapes_device_resident_and_bad_ids(ctx, oracle):
    import torch
    import parry_b200
    T, G, ns = make_table(oracle, ctx)
    s1, p1, v1, s2, p2, v2 = make_pairs(5000, ns, seed=84)
    ho, hs = parry_b200.cast_shapes(G, s1, p1, v1, s2, p2, v2)
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    do, ds = parry_b200.cast_shapes(G, dev(s1), dev(p1), dev(v1), dev(s2), dev(p2), dev(v2))
    ctx.synchronize()
    assert (ds.cpu().numpy() == hs).all() and (do.cpu().numpy().view(np.uint32) == ho.view(np.uint32)).all()
    bad = s1.copy()
    bad[7] = 10 ** 6
    _, bs = parry_b200.cast_shapes(G, bad, p1, v1, s2, p2, v2)
    assert bs[7] == 3 and (np.delete(bs, 7) == np.delete(hs, 7)).all()

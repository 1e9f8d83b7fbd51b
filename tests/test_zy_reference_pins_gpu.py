"""Reference pins added after this round's GPU budget was spent (they call entry points that have run on hardware, on inputs that
have not): the doc examples of contact_manifold.rs and tests/geometry/ball_triangle_toi.rs, the same checks tests/test_oracle_kats.py
runs on the oracle. Kept in a late file so that a surprise here cannot hide the suites before it under `pytest -x`."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _pose(t):
    return np.array([0, 0, 0, 1] + list(t), np.float32)


def test_manifold_doc_examples(ctx):
    """contact_manifold.rs:270-297 and :305-329, the reference's own pins (same checks as tests/test_oracle_kats.py on the oracle):
    unit balls 1.5 apart give one contact of dist -0.5 along +x; 2.1 apart with prediction 0.2 a predicted contact of dist 0.1."""
    import parry_b200
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(1.0)])
    z = np.zeros(3, np.uint32)
    ident = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), (3, 1))
    p2 = ident.copy()
    p2[:, 4] = [1.5, 2.1, 2.1]
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, z[:1], ident[:1], z[:1], p2[:1], 0.0, max_points=4)
    assert st[0] == 0 and cnt[0] == 1 and pts[0, 0, 6] == np.float32(-0.5)
    assert (nr[0] == np.array([1, 0, 0, -1, 0, 0], np.float32)).all() and (pts[0, 0, :6] == np.array([1, 0, 0, -1, 0, 0], np.float32)).all()
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, z[1:2], ident[1:2], z[1:2], p2[1:2], 0.2, max_points=4)
    assert cnt[0] == 1 and pts[0, 0, 6] > 0 and abs(pts[0, 0, 6] - 0.1) < 1e-6
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, z[2:], ident[2:], z[2:], p2[2:], 0.0, max_points=4)
    assert cnt[0] == 0


def test_ball_triangle_toi_issue_123(ctx):
    """crates/parry3d/tests/geometry/ball_triangle_toi.rs: a denormal velocity must answer None (and terminate); the Triangle is a
    3-point ConvexPolyhedron here, as in tests/test_oracle_kats.py."""
    import parry_b200
    tri = np.array([[0.5, -0.5, 0], [-0.5, -0.5, 0], [-0.5, 0.5, 0]], np.float32)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(0.375), parry_b200.ConvexPolyhedron(tri)])
    vel = np.array([[0.0, 6.925e-42, 0.0]], np.float32)
    out, st = parry_b200.cast_shapes(G, np.array([0], np.uint32), _pose([0, 0, 0])[None], vel, np.array([1], np.uint32),
                                     _pose([11.5, 5.5, 0])[None], np.zeros((1, 3), np.float32))
    assert st[0] == 0


def test_shape_cast_doc_examples(ctx):
    """query/shape_cast/shape_cast.rs:196-252: unit balls 10 apart at speed 2 meet at exactly 4.0; overlapping balls answer 0.0 with
    PenetratingOrWithinTargetDist (status 2)."""
    import parry_b200
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(1.0), parry_b200.Ball(2.0)])
    s = np.array([0, 1], np.uint32)
    p1 = np.stack([_pose([0, 0, 0])] * 2)
    p2 = np.stack([_pose([10, 0, 0]), _pose([3, 0, 0])])
    v1 = np.array([[2, 0, 0], [1, 0, 0]], np.float32)
    out, st = parry_b200.cast_shapes(G, s, p1, v1, s, p2, np.zeros((2, 3), np.float32))
    assert st[0] == 1 and out[0, 12] == 4.0
    assert st[1] == 2 and out[1, 12] == 0.0


def test_solid_point_query_example_through_the_ball_arm(ctx):
    """examples/solid_point_query3d.rs through the ball-convex contact arm (see tests/test_oracle_kats.py): Cuboid(1, 2, 2) and a
    Ball of radius 0 at the origin / at (2, 2, 2): dist exactly -1.0 / 1.0, in both argument orders."""
    import parry_b200
    G = parry_b200.Shapes(ctx, [parry_b200.Cuboid([1.0, 2.0, 2.0]), parry_b200.Ball(0.0)])
    ident = _pose([0, 0, 0])
    out, st = parry_b200.contact(G, np.array([0, 0, 1], np.uint32), np.stack([ident, ident, _pose([2, 2, 2])]), np.array([1, 1, 0], np.uint32),
                                 np.stack([ident, _pose([2, 2, 2]), ident]), 2.0)
    assert (st == 1).all()
    assert out[0, 12] == -1.0 and out[1, 12] == 1.0 and out[2, 12] == 1.0
    assert tuple(out[1, 0:3]) == (1.0, 2.0, 2.0)

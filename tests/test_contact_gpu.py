"""GPU parity: query::contact through the C ABI vs the CPU oracle (GJK/EPA + closed forms) on the same inputs.
Pair membership (status) is bit-exact; f32 outputs are required within 1e-5 relative (they are bit-identical in
practice because both sides evaluate the same un-fused arithmetic in the same order)."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu
I4 = [0.0, 0.0, 0.0, 1.0]


def build_tables(ctx, oracle, spec):
    import parry_b200
    gshapes = []
    for k, v in spec:
        gshapes.append(parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v))
    return parry_b200.Shapes(ctx, gshapes), oracle.ShapeTable(spec)


def compare(g, o, rtol=1e-5):
    gout, gst = np.asarray(g[0]), np.asarray(g[1])
    oout, ost = o[0], o[1]
    assert (gst == ost).all(), "status mismatch at %s" % np.nonzero(gst != ost)[0][:10]
    some = ost == 1
    np.testing.assert_allclose(gout[some], oout[some], rtol=rtol, atol=1e-6)
    return (gout[some].view(np.uint32) == oout[some].view(np.uint32)).all(axis=1).mean() if some.any() else 1.0


def test_reference_known_answers(ctx, oracle):
    """crates/parry3d/tests/geometry/epa3.rs:8-23 (exact -0.5 / -1.8), examples/contact_query3d.rs, contact_shape_shape.rs doc."""
    import parry_b200
    G, O = build_tables(ctx, oracle, [("cuboid", [2, 1, 1]), ("ball", 1.0), ("cuboid", [1, 1, 1]), ("ball", 0.5)])
    ident = np.array([I4 + [0, 0, 0]], dtype=np.float32)
    # cuboid_cuboid_EPA: m1 = translation(3.5, 0, 0), m2 = identity
    out, st = parry_b200.contact(G, [0], np.array([I4 + [3.5, 0, 0]], np.float32), [0], ident, 10.0)
    assert st[0] == 1 and out[0, 12] == np.float32(-0.5)
    assert (out[0, 6:9] == [-1, 0, 0]).all()
    out, st = parry_b200.contact(G, [0], np.array([I4 + [0, 0.2, 0]], np.float32), [0], ident, 10.0)
    assert st[0] == 1 and out[0, 12] == np.float32(-1.8)
    assert (out[0, 6:9] == [0, -1, 0]).all()
    # contact_query3d.rs: ball vs unit cube
    for t, check in (([1, 1, 1], lambda s, d: s == 1 and d <= 0), ([2, 2, 2], lambda s, d: s == 1 and d >= 0), ([3, 3, 3], lambda s, d: s == 0)):
        out, st = parry_b200.contact(G, [1], np.array([I4 + t], np.float32), [2], ident, 1.0)
        assert check(st[0], out[0, 12])
    # contact_shape_shape.rs doc: two balls (r = 0.5) with centres 3.2 apart => gap 2.2
    p2 = np.array([I4 + [3.2, 0, 0]], np.float32)
    for pred, expect in ((0.0, 0), (0.5, 0), (3.0, 1)):
        out, st = parry_b200.contact(G, [3], ident, [3], p2, pred)
        assert st[0] == expect
        if expect:
            assert 0 < out[0, 12] <= 3.0 and abs(out[0, 12] - 2.2) < 1e-6
    # bad shape index => Unsupported
    out, st = parry_b200.contact(G, [99], ident, [0], ident, 0.1)
    assert st[0] == 2


def random_pairs(n, n_shapes, seed, spread):
    g = scenes.rng(seed)
    a = g.integers(0, n_shapes, n).astype(np.uint32)
    b = g.integers(0, n_shapes, n).astype(np.uint32)
    q1, q2 = scenes.random_unit_quaternions(g, n), scenes.random_unit_quaternions(g, n)
    t1 = (g.random((n, 3)) - 0.5) * 10
    t2 = t1 + g.standard_normal((n, 3)) * spread
    return a, b, np.concatenate([q1, t1], 1).astype(np.float32), np.concatenate([q2, t2], 1).astype(np.float32)


def mixed_spec(seed, n_each=8):
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(n_each, 32, seed=seed + 1)
    spec = [("ball", float(r)) for r in g.random(n_each) * 0.5 + 0.3]
    spec += [("cuboid", list(h)) for h in g.random((n_each, 3)) * 0.6 + 0.2]
    spec += [("convex", p) for p in pts]
    return spec


def test_all_dispatch_arms_random(ctx, oracle):
    import parry_b200
    spec = mixed_spec(40)
    G, O = build_tables(ctx, oracle, spec)
    for seed, spread, pred in ((41, 1.0, 0.01), (42, 1.8, 0.3), (43, 0.3, 0.0)):
        a, b, p1, p2 = random_pairs(60000, len(spec), seed, spread)
        g = parry_b200.contact(G, a, p1, b, p2, pred)
        o = O.contact(a, p1, b, p2, pred, threads=8)
        assert 0.05 < (o[1] == 1).mean() < 0.999
        exact = compare(g, o)
        assert exact > 0.999, exact
        assert (o[1] < 2).all()


def test_dispatcher_level_contact_in_local_frames(ctx, oracle):
    """pb2_contact_batch_local = QueryDispatcher::contact(pos12, g1, g2, prediction) (query_dispatcher.rs:430-436): what a dispatcher
    chained in front of DefaultQueryDispatcher must return — relative pose in, contact in the shapes' local frames out."""
    import parry_b200
    spec = mixed_spec(40)
    G, O = build_tables(ctx, oracle, spec)
    a, b, p1, p2 = random_pairs(40000, len(spec), 47, 1.2)
    ident = np.tile(np.array(I4 + [0, 0, 0], np.float32), (len(a), 1))
    pos12 = np.concatenate([p2[:, :4], p2[:, 4:] - p1[:, 4:]], axis=1).astype(np.float32)   # a relative pose per pair
    g = parry_b200.contact_local(G, a, b, pos12, 0.05)
    o = O.contact_local(a, ident, b, pos12, 0.05, threads=8)
    assert 0.05 < (o[1] == 1).mean() < 0.999
    assert compare(g, o) > 0.999
    # and query::contact is that result moved to the world frames: consistent with pb2_contact_batch on (identity, pos12) up to the
    # transform of point2 / normal2
    w = parry_b200.contact(G, a, ident, b, pos12, 0.05)
    assert (np.asarray(w[1]) == np.asarray(g[1])).all()
    some = np.asarray(g[1]) == 1
    assert (np.asarray(w[0])[some][:, [0, 1, 2, 6, 7, 8, 12]] == np.asarray(g[0])[some][:, [0, 1, 2, 6, 7, 8, 12]]).all()   # (== : -0.0 vs +0.0 allowed)


def test_cuboid_cuboid_goes_through_gjk_epa(ctx, oracle):
    """The SAT arm is commented out in the reference (default_query_dispatcher.rs:314-317): axis-aligned cuboid pairs are
    full of exact ties (EPA heap order, support copy-sign) and must still match."""
    import parry_b200
    spec = [("cuboid", [1, 1, 1]), ("cuboid", [2, 1, 0.5]), ("cuboid", [0.25, 0.5, 3.0])]
    G, O = build_tables(ctx, oracle, spec)
    g0 = scenes.rng(44)
    n = 20000
    a = g0.integers(0, 3, n).astype(np.uint32)
    b = g0.integers(0, 3, n).astype(np.uint32)
    t = np.round((g0.random((n, 3)) - 0.5) * 8, 1)  # lattice offsets => many exact ties
    p1 = np.tile(np.array(I4 + [0, 0, 0], np.float32), (n, 1))
    p2 = np.concatenate([np.tile(I4, (n, 1)), t], axis=1).astype(np.float32)
    g = parry_b200.contact(G, a, p1, b, p2, 0.05)
    o = O.contact(a, p1, b, p2, 0.05, threads=8)
    assert compare(g, o) > 0.999


def test_ball_arms_edge_cases(ctx, oracle):
    """Coincident centres, ball centre inside / on the surface of a cuboid, zero prediction boundaries."""
    import parry_b200
    spec = [("ball", 0.5), ("ball", 1.0), ("cuboid", [1, 1, 1])]
    G, O = build_tables(ctx, oracle, spec)
    cases = [
        (0, 1, [0, 0, 0], [0, 0, 0]),          # coincident balls: normal = +x
        (0, 1, [0, 0, 0], [1.5, 0, 0]),        # exactly touching: d^2 < (r1+r2+pred)^2 is strict
        (0, 2, [0.2, 0.1, 0.3], [0, 0, 0]),    # ball centre inside the cuboid
        (2, 0, [0, 0, 0], [1.0, 0.2, 0.1]),    # centre on a face: degenerate branch -> feature normal
        (2, 0, [0, 0, 0], [1.0, 1.0, 0.3]),    # centre on an edge
        (2, 0, [0, 0, 0], [1.0, 1.0, 1.0]),    # centre on a vertex
        (0, 2, [1.0, -1.0, 0.0], [0, 0, 0]),   # same, ball first (flipped)
        (2, 0, [0, 0, 0], [1.5, 0, 0]),        # dist == prediction (non-strict <=)
    ]
    a = np.array([c[0] for c in cases], np.uint32)
    b = np.array([c[1] for c in cases], np.uint32)
    p1 = np.array([I4 + c[2] for c in cases], np.float32)
    p2 = np.array([I4 + c[3] for c in cases], np.float32)
    for pred in (0.0, 0.01):
        g = parry_b200.contact(G, a, p1, b, p2, pred)
        o = O.contact(a, p1, b, p2, pred)
        compare(g, o)
    assert o[1][0] == 1 and (o[0][0, 6:9] == [1, 0, 0]).all()


def test_ball_centre_on_a_hull_needs_no_host(ctx, oracle):
    """Ball centre exactly on a ConvexPolyhedron's vertex / edge / face (and one ulp around them): the degenerate branch of
    contact_convex_polyhedron_ball (contact_ball_convex_polyhedron.rs:47-53). The hull's feature there is always
    FeatureId::Unknown (point_support_map.rs:62-77 fails the same normalisation), so the normal is the normalised projection, else
    +y — no status 3, identical to the oracle, also through closest_points / distance."""
    import parry_b200
    cube = np.array([[x, y, z] for x in (-1, 1) for y in (-1, 1) for z in (-1, 1)], np.float32)
    shifted = cube + np.array([1, 1, 1], np.float32)                     # one vertex at the origin: the +y fall-back
    spec = [("ball", 0.5), ("convex", cube), ("convex", shifted)]
    G, O = build_tables(ctx, oracle, spec)
    spots = [[1, 1, 1], [1, 1, 0.25], [1, 0.5, -0.25], [-1, -1, 1], [1, 0, 0], [0.5, 0.25, 0.125]]
    cases = [(1, 0, s) for s in spots] + [(0, 1, [-x for x in s]) for s in spots] + [(2, 0, [0, 0, 0]), (2, 0, [2, 2, 2]), (0, 2, [0, 0, 0])]
    up = np.nextafter(np.float32(1), np.float32(2)); dn = np.nextafter(np.float32(1), np.float32(0))
    cases += [(1, 0, [up, 1, 1]), (1, 0, [dn, dn, dn]), (1, 0, [up, 0.5, 0.5]), (1, 0, [dn, 0.5, 0.5])]
    a = np.array([c[0] for c in cases], np.uint32)
    b = np.array([c[1] for c in cases], np.uint32)
    p1 = np.tile(np.array(I4 + [0, 0, 0], np.float32), (len(cases), 1))
    p2 = np.array([I4 + list(map(float, c[2])) for c in cases], np.float32)
    for pred in (0.0, 0.01):
        g = parry_b200.contact(G, a, p1, b, p2, pred)
        o = O.contact(a, p1, b, p2, pred)
        assert (np.asarray(g[1]) != 3).all() and (o[1] != 3).all()
        compare(g, o)
    # centre on a face / an edge: dist = -radius, normal1 = the normalised projection; projection at the origin: +y. (Exactly on a
    # vertex GJK stops on a 0-dimensional simplex and EPA answers with its placeholder, epa3.rs:451-455: proj = origin, as the oracle.)
    assert o[1][4] == 1 and o[0][4, 12] == np.float32(-0.5) and (o[0][4, 6:9] == [1, 0, 0]).all()
    np.testing.assert_allclose(o[0][1, 6:9], np.array([1, 1, 0.25]) / np.linalg.norm([1, 1, 0.25]), atol=1e-6)
    k = 2 * len(spots)
    assert o[1][k] == 1 and (o[0][k, 6:9] == [0, 1, 0]).all()
    # rotated poses too
    g0 = scenes.rng(77)
    q = scenes.random_unit_quaternions(g0, len(cases)).astype(np.float32)
    r1 = np.concatenate([q, (g0.random((len(cases), 3)) * 2).astype(np.float32)], axis=1)
    hull_first = a == 1
    sel = np.nonzero(hull_first)[0]
    # ball pose = hull pose * local spot (exactly representable offsets keep some of the cases degenerate, the others nearly so)
    from harness.scenes import compose_pose
    r2 = np.stack([compose_pose(r1[i], p2[i]) for i in sel])
    g = parry_b200.contact(G, a[sel], r1[sel], b[sel], r2, 0.01)
    o = O.contact(a[sel], r1[sel], b[sel], r2, 0.01)
    assert (np.asarray(g[1]) != 3).all() and (o[1] != 3).all()
    compare(g, o)
    d_g = parry_b200.distance(G, a, p1, b, p2)
    d_o = O.distance(a, p1, b, p2)
    assert (np.asarray(d_g[1]) != 3).all() and (np.asarray(d_g[0]).view(np.uint32) == d_o[0].view(np.uint32)).all()


def test_hull_pairs_config3_slice_and_compact(ctx, oracle):
    """BASELINE config[2] inputs (32-vertex hulls, prediction 0.01) on a slice the oracle finishes in seconds."""
    import parry_b200
    pts, radii = scenes.hull_pool(512)
    spec = [("convex", p) for p in pts]
    G, O = build_tables(ctx, oracle, spec)
    a, b, p1, p2 = scenes.hull_pairs(200000, radii, seed=4)
    g = parry_b200.contact(G, a, p1, b, p2, 0.01)
    o = O.contact(a, p1, b, p2, 0.01, threads=8)
    exact = compare(g, o)
    assert exact > 0.9999, exact
    assert (o[1] == 3).sum() == 0
    cg, idx = parry_b200.contact_compact(G, a, p1, b, p2, 0.01)
    order = np.argsort(idx)
    assert (idx[order] == np.nonzero(o[1] == 1)[0]).all()
    assert (cg[order].view(np.uint32) == g[0][o[1] == 1].view(np.uint32)).all()


def test_device_resident_contacts(ctx, oracle):
    import torch
    import parry_b200
    spec = mixed_spec(50)
    G, O = build_tables(ctx, oracle, spec)
    a, b, p1, p2 = random_pairs(30000, len(spec), 51, 0.8)
    h = parry_b200.contact(G, a, p1, b, p2, 0.02)
    d = parry_b200.contact(G, torch.from_numpy(a.astype(np.int32)).cuda(), torch.from_numpy(p1).cuda(),
                           torch.from_numpy(b.astype(np.int32)).cuda(), torch.from_numpy(p2).cuda(), 0.02)
    ctx.synchronize()
    assert (d[1].cpu().numpy() == h[1]).all()
    assert (d[0].cpu().numpy().view(np.uint32) == h[0].view(np.uint32)).all()


def test_full_size_config3_properties(ctx, oracle):
    """BASELINE config[2] at full size (2^22 pairs, 4096 hulls): oracle parity on a random 2^17 subset; on all pairs the
    size-independent properties of a contact: unit normals, normal1 == -normal2 in world space, dist == (p2-p1).n1."""
    import parry_b200
    pts, radii = scenes.hull_pool(4096)
    spec = [("convex", p) for p in pts]
    G, O = build_tables(ctx, oracle, spec)
    n = 1 << 22
    a, b, p1, p2 = scenes.hull_pairs(n, radii, seed=4)
    out, st = parry_b200.contact(G, a, p1, b, p2, 0.01)
    assert (st <= 1).all()
    some = st == 1
    assert 0.5 < some.mean() < 0.8
    c = out[some].astype(np.float64)
    n1, n2 = c[:, 6:9], c[:, 9:12]
    assert np.abs(np.linalg.norm(n1, axis=1) - 1).max() < 1e-5
    assert np.abs(n1 + n2).max() < 1e-5
    d = np.einsum("ij,ij->i", c[:, 3:6] - c[:, 0:3], n1)
    assert np.abs(d - c[:, 12]).max() < 2e-5
    assert c[:, 12].max() <= 0.0101
    sub = scenes.rng(77).choice(n, 1 << 17, replace=False)
    o = O.contact(a[sub], p1[sub], b[sub], p2[sub], 0.01, threads=8)
    assert compare((out[sub], st[sub]), o) > 0.9999

"""CPU: pins the oracle (C++ restatement of parry3d) against every known-answer the reference's own tests, examples and
doc-tests hold for the hot path. These are the only reference-provided vectors: parry3d itself cannot be
compiled here (no cargo, nalgebra not vendored)."""
import numpy as np

from harness import scenes

FMAX = float(np.finfo(np.float32).max)
I4 = [0.0, 0.0, 0.0, 1.0]
INVALID = 0xFFFFFFFF


def test_cuboid_cuboid_epa_exact(oracle):
    """crates/parry3d/tests/geometry/epa3.rs:8-23 — exact equality, as in the reference's assert_eq!."""
    T = oracle.ShapeTable([("cuboid", [2, 1, 1])])
    st, c = T.dispatch_contact(0, 0, I4 + [-3.5, 0, 0], 10.0)  # m1.inv_mul(m2), m1 = translation(3.5,0,0)
    assert st == 1 and c[12] == np.float32(-0.5) and (c[6:9] == [-1, 0, 0]).all()
    st, c = T.dispatch_contact(0, 0, I4 + [0, -0.2, 0], 10.0)
    assert st == 1 and c[12] == np.float32(-1.8) and (c[6:9] == [0, -1, 0]).all()


def test_triangle_vertex_touches_triangle_edge_epa(oracle):
    """epa3.rs:25-59 (issues #253/#246): must produce ClosestPoints with a.y ~ 1.0 and a.x ~ -2.349647 (1e-3)."""
    t1 = [[-13.174434, 1.0, 8.736801], [3.5251038, 1.0, 12.1], [3.2048466, 1.0, 12.218325]]
    t2 = [[-1.63, 0.0, 11.19], [-2.349647, 0.0, 11.037681], [-2.349647, 1.0, 11.037681]]
    T = oracle.ShapeTable([("triangle", t1), ("triangle", t2)])
    st, c = T.dispatch_contact(0, 1, I4 + [0, 0, 0], 0.00999999977)
    assert st == 1
    assert abs(c[1] - 1.0) < 1e-3 and abs(c[0] - (-2.349647)) < 1e-3


def test_contact_query3d_example(oracle):
    """crates/parry3d/examples/contact_query3d.rs: ball (r=1) vs unit cube, prediction 1."""
    T = oracle.ShapeTable([("ball", 1.0), ("cuboid", [1, 1, 1])])
    ident = np.array([I4 + [0, 0, 0]], np.float32)
    out, st = T.contact([0], [I4 + [1, 1, 1]], [1], ident, 1.0)
    assert st[0] == 1 and out[0, 12] <= 0.0
    out, st = T.contact([0], [I4 + [2, 2, 2]], [1], ident, 1.0)
    assert st[0] == 1 and out[0, 12] >= 0.0
    out, st = T.contact([0], [I4 + [3, 3, 3]], [1], ident, 1.0)
    assert st[0] == 0


def test_contact_doc_examples(oracle):
    """contact_shape_shape.rs doc (two r=0.5 balls, gap 2.2) and contact.rs:37-63 (overlapping balls => dist < 0)."""
    T = oracle.ShapeTable([("ball", 0.5), ("ball", 1.0)])
    p1, p2 = [I4 + [0, 0, 0]], [I4 + [3.2, 0, 0]]
    assert T.contact([0], p1, [0], p2, 0.0)[1][0] == 0
    assert T.contact([0], p1, [0], p2, 0.5)[1][0] == 0
    out, st = T.contact([0], p1, [0], p2, 3.0)
    assert st[0] == 1 and 0 < out[0, 12] <= 3.0
    out, st = T.contact([1], p1, [1], [I4 + [1.5, 0, 0]], 0.0)
    assert st[0] == 1 and out[0, 12] < 0
    assert np.allclose(out[0, 6:9], [1, 0, 0]) and np.allclose(out[0, 9:12], [-1, 0, 0])


def test_ball_cuboid_both_argument_orders(oracle):
    """crates/parry2d/tests/geometry/ball_cuboid_contact.rs semantics (3D flavour): a contact must exist in both orders."""
    T = oracle.ShapeTable([("ball", 0.5), ("cuboid", [1, 1, 1])])
    for t in ([1.2, 0, 0], [0, 1.4, 0.1], [0.9, 0.9, 0.9]):
        a = T.contact([0], [I4 + t], [1], [I4 + [0, 0, 0]], 0.0)
        b = T.contact([1], [I4 + [0, 0, 0]], [0], [I4 + t], 0.0)
        assert a[1][0] == 1 and b[1][0] == 1
        assert np.allclose(a[0][0, 12], b[0][0, 12])
        assert np.allclose(a[0][0, 0:3], b[0][0, 3:6], atol=1e-6)


def test_solid_ray_cast3d_example(oracle):
    """crates/parry3d/examples/solid_ray_cast3d.rs: cuboid (1,2,1)."""
    he = [1, 2, 1]
    inside = [0, 0, 0, 0, 1, 0]
    miss = [2, 2, 2, 1, 1, 1]
    assert oracle.shape_cast_ray_toi(1, he, I4 + [0, 0, 0], inside, FMAX, True) == 0.0
    assert oracle.shape_cast_ray_toi(1, he, I4 + [0, 0, 0], inside, FMAX, False) == 2.0
    assert oracle.shape_cast_ray_toi(1, he, I4 + [0, 0, 0], miss, FMAX, False) is None
    assert oracle.shape_cast_ray_toi(1, he, I4 + [0, 0, 0], miss, FMAX, True) is None


def test_bvh_node_cast_ray_doc(oracle):
    """bvh_tree.rs:1155-1171 doc-test: node box [5,6]x[-1,1]^2, ray from the origin along +x => toi == 5.0."""
    b = oracle.Bvh(np.array([[5, -1, -1, 6, 1, 1]], np.float32))
    toi, leaf = b.cast_rays_shapes([1], [[0.5, 1, 1]], [I4 + [5.5, 0, 0]], [[0, 0, 0, 1, 0, 0]], FMAX)
    assert leaf[0] == 0 and toi[0] == 5.0


def test_intersect_aabb_doc_examples(oracle):
    """bvh_queries.rs:130-155 doc-test: the far object is culled."""
    boxes = np.array([[0, 0, 0, 1, 1, 1], [2, 0, 0, 3, 1, 1], [100, 0, 0, 101, 1, 1]], np.float32)
    b = oracle.Bvh(boxes)
    offs, ids = b.intersect_aabbs(np.array([[-10, -10, -10, 10, 10, 10]], np.float32))
    assert sorted(ids.tolist()) == [0, 1]
    b2 = oracle.Bvh(np.array([[0, 0, 0, 1, 1, 1], [1.5, 0, 0, 2.5, 1, 1], [100, 0, 0, 101, 1, 1]], np.float32))
    offs, ids = b2.intersect_aabbs(np.array([[-2, -2, -2, 3, 3, 3]], np.float32))
    assert sorted(ids.tolist()) == [0, 1]


def test_shape_ray_cast_points_to_surface(oracle):
    """crates/parry3d/tests/geometry/cuboid_ray_cast.rs:7-90 property (issue #242) with our seeded RNG: the hit point,
    nudged outward along the normal and re-cast away from the shape, must not hit again; nudged inward it is inside."""
    g = scenes.rng(42)
    shapes = [(0, [1.0, 0, 0]), (1, [1, 1, 1]), (1, [1, 1, 0.5]), (1, [0.5, 1, 0.5])]
    for kind, p in shapes:
        for _ in range(1000):
            o = g.random(3)
            o = o / np.linalg.norm(o) * 5.0
            ray = np.concatenate([o, -o]).astype(np.float32)
            q = g.random(4)
            q = np.array([0, 0, 0, 1.0]) if g.random() < 0.01 else q / np.linalg.norm(q)
            pose = np.concatenate([q, [0, 0, 0]]).astype(np.float32)
            hit = oracle.shape_cast_ray(kind, p, pose, ray, FMAX, True)
            assert hit is not None
            toi, n, _ = hit
            pt = ray[:3] + ray[3:] * np.float32(toi)
            out = pt + n * np.float32(0.001)
            new_ray = np.concatenate([out, ray[:3] - out]).astype(np.float32)
            assert oracle.shape_cast_ray(kind, p, pose, new_ray, FMAX, True) is None
            inn = pt - n * np.float32(0.001)
            # contains_point: a ray starting inside a solid shape reports toi == 0
            assert oracle.shape_cast_ray_toi(kind, p, pose, np.concatenate([inn, [1, 0, 0]]).astype(np.float32), FMAX, True) == 0.0


def test_single_triangle_mesh_faces(oracle):
    """ray_trimesh.rs:187-212 scene (one triangle, rays from both sides): front face => Face(0), back face => Face(0 + 1)."""
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    i = np.array([[0, 1, 2]], np.uint32)
    m = oracle.TriMesh(v, i)
    up = np.array([[0.1, 0.1, -1, 0, 0, 1]], np.float32)     # along +z: hits the back face (normal is +z)
    down = np.array([[0.1, 0.1, 1, 0, 0, -1]], np.float32)
    toi, tri, n, f = m.cast_rays(None, up, 1000.0, with_normal=True)
    assert tri[0] == 0 and toi[0] == 1.0 and f[0] == 1 and (n[0] == [0, 0, -1]).all()
    toi, tri, n, f = m.cast_rays(None, down, 1000.0, with_normal=True)
    assert tri[0] == 0 and toi[0] == 1.0 and f[0] == 0 and (n[0] == [0, 0, 1]).all()
    # toi-only variant post-filters toi < max_toi (ray_composite_shape.rs:38)
    assert m.cast_rays(None, down, 1.0)[1][0] == INVALID
    assert m.cast_rays(None, down, 1.0001)[1][0] == 0


def test_bvh_build_well_formed_all_sizes(oracle):
    """bvh_tests.rs:34-123 (build part): Binned and Ploc, len 1..=100, assert_well_formed (bvh_validation.rs:61-134)."""
    from helpers import assert_well_formed
    g = scenes.rng(7)
    for strategy in (0, 1):
        for n in list(range(1, 101)) + [1000]:
            c = g.random((n, 3)) * 10
            aabbs = np.concatenate([c - 0.5, c + 0.5], axis=1).astype(np.float32)
            b = oracle.Bvh(aabbs, strategy)
            nodes = b.nodes()
            assert len(nodes) == (1 if n <= 2 else n - 1)
            if n > 2:
                assert_well_formed(nodes, b.parents(), b.leaf_node_indices())
            # every leaf reachable through intersect_aabb with an all-enclosing box (test_leaves_iteration analogue)
            offs, ids = b.intersect_aabbs(np.array([[-100, -100, -100, 100, 100, 100]], np.float32))
            assert sorted(ids.tolist()) == list(range(n))


def test_pairs_match_brute_force(oracle):
    from helpers import sorted_pairs
    g = scenes.rng(8)
    for n in (2, 3, 10, 500):
        c = g.random((n, 3)) * (n ** (1 / 3)) * 1.2
        h = g.random((n, 3)) * 0.6 + 0.1
        a = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
        for strategy in (0, 1):
            b = oracle.Bvh(a, strategy)
            got = sorted_pairs(b.self_pairs())
            ref = []
            for i in range(n):
                for j in range(i + 1, n):
                    if (a[i, :3] <= a[j, 3:]).all() and (a[i, 3:] >= a[j, :3]).all():
                        ref.append((i, j))
            assert (got == sorted_pairs(np.array(ref).reshape(-1, 2))).all()


def test_ray_ball_analytic(oracle):
    """Unit ball, axis rays: closed-form distances (ray_ball.rs)."""
    assert oracle.shape_cast_ray_toi(0, [1.0], None, [-3, 0, 0, 1, 0, 0], FMAX, True) == 2.0
    assert oracle.shape_cast_ray_toi(0, [1.0], None, [0, 0, 0, 1, 0, 0], FMAX, True) == 0.0
    assert oracle.shape_cast_ray_toi(0, [1.0], None, [0, 0, 0, 1, 0, 0], FMAX, False) == 1.0
    assert oracle.shape_cast_ray_toi(0, [1.0], None, [-3, 0, 0, -1, 0, 0], FMAX, True) is None
    assert oracle.shape_cast_ray_toi(0, [1.0], None, [-3, 0, 0, 1, 0, 0], 1.5, True) is None
    t, n, f = oracle.shape_cast_ray(0, [1.0], None, [-3, 0, 0, 1, 0, 0], FMAX, True)
    assert (n == [-1, 0, 0]).all() and f == 0


def test_convex_ray_cast_matches_cuboid_closed_form(oracle):
    """No reference test pins ray casts on a ConvexPolyhedron (ray_support_map.rs); the restatement of the GJK ray cast
    (gjk.rs:660-795) is cross-checked against the closed-form cuboid cast on the hull of a cuboid's corners: same hit/miss,
    toi within GJK's tolerance, for solid and non-solid casts (unit directions: the non-solid branch of
    ray_support_map.rs:31-56 mixes distance and ray-parameter units otherwise, which the restatement keeps)."""
    g = scenes.rng(5)
    he = np.array([0.7, 0.4, 1.1], np.float32)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    hits = 0
    for solid in (True, False):
        for k in range(1500):
            o = (g.random(3) - 0.5) * (1.0 if k % 4 == 0 else 6.0)
            d = g.standard_normal(3)
            d /= np.linalg.norm(d)
            ray = np.concatenate([o, d]).astype(np.float32)
            pose = np.concatenate([scenes.random_unit_quaternions(g, 1)[0], g.random(3) - 0.5]).astype(np.float32)
            a = oracle.convex_cast_ray(corners, pose, ray, FMAX, solid)
            b = oracle.shape_cast_ray(1, he, pose, ray, FMAX, solid)
            assert (a is None) == (b is None)
            if a is not None:
                hits += 1
                assert abs(float(a[0]) - float(b[0])) < 2e-3
    assert hits > 500


def test_convex_ray_cast_points_to_surface(oracle):
    """cuboid_ray_cast.rs:7-90 property on random 16-point hulls: the hit point nudged outward and re-cast away from the
    shape does not hit again; nudged inward, a solid cast from there reports toi == 0."""
    g = scenes.rng(43)
    pts, _ = scenes.hull_pool(8, 16, seed=44)
    for h in range(8):
        for _ in range(150):
            o = g.standard_normal(3)
            o = o / np.linalg.norm(o) * 5.0
            ray = np.concatenate([o, -o]).astype(np.float32)
            pose = np.concatenate([scenes.random_unit_quaternions(g, 1)[0], [0, 0, 0]]).astype(np.float32)
            hit = oracle.convex_cast_ray(pts[h], pose, ray, FMAX, True)
            assert hit is not None
            toi, n = hit
            pt = ray[:3] + ray[3:] * np.float32(toi)
            out = pt + n * np.float32(0.002)
            assert oracle.convex_cast_ray(pts[h], pose, np.concatenate([out, ray[:3] - out]).astype(np.float32), FMAX, True) is None
            inn = pt - n * np.float32(0.002)
            back = oracle.convex_cast_ray(pts[h], pose, np.concatenate([inn, [1, 0, 0]]).astype(np.float32), FMAX, True)
            assert back is not None and back[0] == 0.0


def _pose(t):
    return np.array([0, 0, 0, 1] + list(t), np.float32)


def test_cast_shapes_reference_tests(oracle):
    """crates/parry3d/tests/geometry/ball_ball_toi.rs (exact 0.9), time_of_impact3.rs (Some(0.0), relative_eq value, None) and
    still_objects_toi.rs (issue #141: None, None, Some)."""
    T = oracle.ShapeTable([("ball", 0.5)])
    out, st = T.cast_shapes([0], [_pose([0, 0, 0])], [[0, 10, 0]], [0], [_pose([0, 10, 0])], [[0, 0, 0]])
    assert st[0] == 1 and out[0, 12] == np.float32(0.9)
    T = oracle.ShapeTable([("ball", 1.0), ("cuboid", [1, 1, 1])])
    out, st = T.cast_shapes([0, 0, 0], [_pose([1, 1, 1]), _pose([2, 2, 2]), _pose([3, 3, 3])], [[2, 2, 2], [-.5, -.5, -.5], [2, 2, 2]],
                            [1, 1, 1], [_pose([0, 0, 0])] * 3, [[-1, 1, 1], [1, 1, 1], [-1, 1, 1]])
    assert st[0] != 0 and out[0, 12] == 0.0
    expect = (np.sqrt(np.float32(3.0)) - np.float32(1.0)) / np.linalg.norm(np.array([-1.5, -1.5, -1.5], np.float32))
    assert st[1] == 1 and abs(out[1, 12] - expect) <= np.finfo(np.float32).eps * max(abs(expect), abs(out[1, 12]))
    assert st[2] == 0
    T = oracle.ShapeTable([("cuboid", [.5, .5, .5])])
    got = [T.cast_shapes([0], [_pose([0, 1.1, 0])], [[0, vy, 0]], [0], [_pose([0, 0, 0])], [[0, 0, 0]])[1][0] for vy in (0.0, 1.0, -1.0)]
    assert got[0] == 0 and got[1] == 0 and got[2] != 0
    # ball_triangle_toi.rs (issue #123, once an infinite loop): a denormal velocity is "too small", the cast answers None. The
    # Triangle is given as a 3-point ConvexPolyhedron (same support function up to ties, which play no role here).
    tri = np.array([[0.5, -0.5, 0], [-0.5, -0.5, 0], [-0.5, 0.5, 0]], np.float32)
    T = oracle.ShapeTable([("ball", 0.375), ("convex", tri)])
    vel = np.array([[0.0, 6.925e-42, 0.0]], np.float32)
    assert vel[0, 1] > 0                                            # still a (denormal) non-zero f32
    out, st = T.cast_shapes([0], [_pose([0, 0, 0])], vel, [1], [_pose([11.5, 5.5, 0])], [[0, 0, 0]])
    assert st[0] == 0


def test_cast_shapes_hit_configuration_touches(oracle):
    """Property at the oracle level: advancing both shapes to the reported time of impact leaves them at distance ~
    target_distance, and the witnesses (local frames) coincide in world space up to that distance."""
    g = scenes.rng(71)
    pts, radii = scenes.hull_pool(8, 16, seed=72)
    T = oracle.ShapeTable([("ball", 0.4), ("cuboid", [0.3, 0.5, 0.4])] + [("convex", p) for p in pts])
    n = 400
    s1, s2 = g.integers(0, 10, n), g.integers(0, 10, n)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - 0.5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * 4.0], axis=1).astype(np.float32)
    v1 = (d * 3.0 + g.standard_normal((n, 3)) * 0.4).astype(np.float32)
    v2 = np.zeros((n, 3), np.float32)
    for target in (0.0, 0.05):
        out, st = T.cast_shapes(s1, p1, v1, s2, p2, v2, target_distance=target)
        hit = st == 1
        assert hit.mean() > 0.4
        q1, q2 = p1[hit].copy(), p2[hit].copy()
        q1[:, 4:] += v1[hit] * out[hit, 12:13]
        dist, ds = T.distance(s1[hit], q1, s2[hit], q2)
        ok = ds == 0
        assert np.abs(dist[ok] - target).max() < 5e-3


def test_compound_single_part_equals_plain_contact(oracle):
    """contact_composite_shape_shape.rs:14-45 with one part at the identity pose degenerates to the part's own contact (bit for
    bit); with the part moved by a pose P it equals the plain contact of the part posed at pos1 * P up to rounding."""
    g = scenes.rng(3)
    pts, _ = scenes.hull_pool(4, 16, seed=5)
    T = oracle.ShapeTable([("ball", 0.4), ("cuboid", [0.3, 0.5, 0.4])] + [("convex", p * 0.6) for p in pts])
    n = 2000
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    cf, cc = np.arange(6, dtype=np.uint32), np.ones(6, np.uint32)
    cid, sid = g.integers(0, 6, n).astype(np.uint32), g.integers(0, 6, n).astype(np.uint32)
    pc = np.concatenate([scenes.random_unit_quaternions(g, n), g.random((n, 3)) - .5], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ps = np.concatenate([scenes.random_unit_quaternions(g, n), pc[:, 4:] + d * (g.random((n, 1)) * 1.2 + 0.2)], axis=1).astype(np.float32)
    o1, s1, p1 = T.contact_compound(cf, cc, np.arange(6, dtype=np.uint32), np.tile(ident, (6, 1)), cid, pc, sid, ps, 0.05)
    o2, s2 = T.contact(cid, pc, sid, ps, 0.05)
    assert (s1 == s2).all() and (s2 == 1).mean() > 0.3 and (o1.view(np.uint32) == o2.view(np.uint32)).all()
    assert (p1[s1 == 1] == 0).all() and (p1[s1 != 1] == 0xFFFFFFFF).all()
    o3, s3, _ = T.contact_compound(cf, cc, np.arange(6, dtype=np.uint32), np.tile(ident, (6, 1)), cid, pc, sid, ps, 0.05, compound_second=True)
    o4, s4 = T.contact(sid, ps, cid, pc, 0.05)
    assert (s3 == s4).all()
    np.testing.assert_allclose(o3[s4 == 1], o4[s4 == 1], rtol=1e-3, atol=2e-4)


def test_manifold_restatement_properties(oracle):
    """No reference test pins contact manifolds; the restatement (ball-ball, ball-cuboid, cuboid-cuboid SAT + face clipping) is
    checked through what the manifolds must satisfy: a manifold exists exactly when query::contact finds a contact (up to the
    strict / non-strict prediction comparison), its deepest point has query::contact's distance (exactly for the closed-form
    arms, within the SAT-vs-GJK feature difference for cuboid pairs), every local_p1 / local_p2 lies on its shape, and
    local_p2 - local_p1 (in frame 1) is dist * local_n1."""
    g = scenes.rng(7)
    T = oracle.ShapeTable([("ball", 0.4), ("ball", 0.25), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("cuboid", [0.5, 0.5, 0.5])])
    he = {2: [0.3, 0.5, 0.4], 3: [0.6, 0.2, 0.2], 4: [0.5, 0.5, 0.5]}
    rad = {0: 0.4, 1: 0.25}
    n = 6000
    s1, s2 = g.integers(0, 5, n).astype(np.uint32), g.integers(0, 5, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.3 + 0.2)], axis=1).astype(np.float32)
    p2[::5, :4] = p1[::5, :4]
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05)
    co, cs = T.contact(s1, p1, s2, p2, 0.05)
    assert (st == 0).all() and cnt.max() <= 8
    has = cnt > 0
    # (the 15 SAT axes under-estimate the distance of vertex-vertex / vertex-edge configurations, hence a few manifolds without contact)
    assert (has & (cs != 1)).sum() <= 0.002 * n and ((~has) & (cs == 1)).sum() <= 0.002 * n
    valid = np.arange(16)[None, :] < cnt[:, None]
    mind = np.min(np.where(valid, pts[:, :, 6], np.inf), axis=1)
    both = has & (cs == 1)
    cc = (s1 >= 2) & (s2 >= 2)
    assert np.abs(mind[both & ~cc] - co[both & ~cc, 12]).max() < 1e-6
    assert np.quantile(np.abs(mind[both & cc] - co[both & cc, 12]), 0.99) < 1e-5

    def on_surface(shape, p):
        if shape in rad:
            return abs(np.linalg.norm(p) - rad[shape]) < 1e-5
        h = np.array(he[shape])
        return (np.abs(p) <= h + 1e-5).all() and (np.abs(np.abs(p) - h) < 1e-5).any()
    checked = 0
    for k in np.nonzero(has)[0][:1500]:
        for i in range(cnt[k]):
            lp1, lp2, dist = pts[k, i, :3], pts[k, i, 3:6], pts[k, i, 6]
            # closed-form arms put both points on the surfaces; face clipping puts one of them on its face and the other on the
            # line through it along the normal
            ok1, ok2 = on_surface(int(s1[k]), lp1), on_surface(int(s2[k]), lp2)
            assert ok1 or ok2
            if not cc[k]:
                assert ok1 and ok2
            checked += 1
    assert checked > 2000


def test_closest_points_doc_examples_and_consistency(oracle):
    """Doc examples of closest_points_shape_shape.rs:136-200 and consistency with query::distance: WithinMargin points are
    `distance` apart, Intersecting pairs have distance 0, Disjoint pairs are farther apart than the margin."""
    pose = lambda t: np.array([0, 0, 0, 1] + list(t), np.float32)
    T = oracle.ShapeTable([("ball", 0.5), ("ball", 2.0), ("cuboid", [1, 1, 1])])
    o, k, s = T.closest_points([0], [pose([0, 0, 0])], [0], [pose([12, 0, 0])], 15.0)
    assert k[0] == 1 and np.allclose(o[0], [0.5, 0, 0, 11.5, 0, 0], atol=1e-6)
    o, k, s = T.closest_points([0], [pose([0, 0, 0])], [0], [pose([12, 0, 0])], 5.0)
    assert k[0] == 0
    o, k, s = T.closest_points([1], [pose([5, 0, 0])], [2], [pose([0, 0, 0])], 10.0)
    assert k[0] == 1 and np.allclose(o[0], [3, 0, 0, 1, 0, 0], atol=1e-6)
    g = scenes.rng(5)
    pts, _ = scenes.hull_pool(4, 16, seed=6)
    T = oracle.ShapeTable([("ball", 0.4), ("cuboid", [0.3, 0.5, 0.4])] + [("convex", p * 0.6) for p in pts])
    n = 5000
    s1, s2 = g.integers(0, 6, n).astype(np.uint32), g.integers(0, 6, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 2.5 + 0.2)], axis=1).astype(np.float32)
    o, k, s = T.closest_points(s1, p1, s2, p2, 0.8)
    dist, ds = T.distance(s1, p1, s2, p2)
    assert (s == 1).all() and (np.bincount(k) > 500).all()
    w = (k == 1) & (ds == 0)
    assert np.abs(np.linalg.norm(o[w, :3] - o[w, 3:], axis=1) - dist[w]).max() < 5e-6
    assert (dist[(k == 2) & (ds == 0)] == 0).all() and (dist[(k == 0) & (ds == 0)] > 0.8 - 1e-5).all()


def test_pfm_manifold_matches_cuboid_sat_manifold(oracle):
    """contact_manifolds_pfm_pfm.rs has no reference test; the restatement (GJK/EPA contact -> support faces -> face clipping ->
    + witness pair) is cross-checked against the independent cuboid-cuboid SAT arm: the hull of a cuboid's corners (topology from
    harness/hull_topology.py: 6 quads, 18 edges of which 6 merged diagonals) must give the cuboid manifold's normal, deepest
    distance and points, plus exactly one more point (the witness pair)."""
    g = scenes.rng(11)
    hes = [np.array([0.3, 0.5, 0.4], np.float32), np.array([0.6, 0.2, 0.2], np.float32)]
    corners = lambda he: np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    T = oracle.ShapeTable([("cuboid", hes[0]), ("cuboid", hes[1]), ("convex", corners(hes[0])), ("convex", corners(hes[1]))])
    topo = T.hull_topology()
    assert (topo["hull_face_count"][2:] == 6).all() and (topo["face_count"] == 4).all()
    n = 4000
    a, b = g.integers(0, 2, n), g.integers(0, 2, n)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    nc, cc, pc, sc = T.contact_manifolds(a.astype(np.uint32), p1, b.astype(np.uint32), p2, 0.05)
    nh, ch, ph, sh = T.contact_manifolds((a + 2).astype(np.uint32), p1, (b + 2).astype(np.uint32), p2, 0.05, topology=topo)
    nu, cu, pu, su = T.contact_manifolds((a + 2).astype(np.uint32), p1, (b + 2).astype(np.uint32), p2, 0.05)
    assert (sc == 0).all() and (sh == 0).all() and (su == 2).all() and (cu == 0).all()     # without topology: unsupported
    both = (cc > 0) & (ch > 0)
    assert both.sum() > 2000 and ((cc > 0) != (ch > 0)).sum() <= 0.005 * n
    agree = np.abs(nc[both, :3] - nh[both, :3]).max(axis=1) < 1e-3
    assert agree.mean() > 0.995
    idx = np.nonzero(both)[0][agree]
    assert (ch[idx] == cc[idx] + 1).all()
    valid = lambda c: np.arange(16)[None, :] < c[:, None]
    mc = np.min(np.where(valid(cc), pc[:, :, 6], np.inf), axis=1)
    mh = np.min(np.where(valid(ch), ph[:, :, 6], np.inf), axis=1)
    assert np.quantile(np.abs(mc[idx] - mh[idx]), 0.99) < 1e-5
    # the clipped points are the cuboid manifold's points (the hull's face starts at another corner, so in another order); the
    # last one is the witness pair with UNKNOWN features
    def rows(a):
        a = np.round(a.astype(np.float64), 3)
        return a[np.lexsort(a.T[::-1])]
    for k in idx[:500]:
        np.testing.assert_allclose(rows(ph[k, :cc[k], :7]), rows(pc[k, :cc[k], :7]), rtol=0, atol=2.1e-3)
        assert (ph[k, cc[k], 7:].view(np.uint32) == 0).all()


def test_ball_hull_manifold_matches_cuboid_closed_form(oracle):
    """contact_manifolds_convex_ball.rs with a ConvexPolyhedron (GJK/EPA projection + support_feature_id_toward) against the same
    arm with a Cuboid (closed-form projection, point_aabb.rs): on the hull of the cuboid's corners both give the same contact, and
    the same kind of feature (vertex / edge / face) away from the 1-degree thresholds; and it agrees with query::contact."""
    g = scenes.rng(17)
    pts, _ = scenes.hull_pool(6, 16, seed=18)
    he = np.array([0.3, 0.5, 0.4], np.float32)
    corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    T = oracle.ShapeTable([("ball", 0.3), ("cuboid", he), ("convex", corners)] + [("convex", p * 0.6) for p in pts])
    topo = T.hull_topology()
    n = 6000
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.05)], axis=1).astype(np.float32)
    z = np.zeros(n, np.uint32)
    nb, cb, pb, sb = T.contact_manifolds(z, p1, z + 1, p2, 0.05)
    nh, ch, ph, sh = T.contact_manifolds(z, p1, z + 2, p2, 0.05, topology=topo)
    assert (sb == 0).all() and (sh == 0).all() and (cb == ch).all() and (cb > 0).sum() > 2000
    both = cb > 0
    assert np.abs(nb[both] - nh[both]).max() < 3e-3 and np.abs(pb[both, 0, :7] - ph[both, 0, :7]).max() < 1e-3
    kb, kh = pb[both, 0, 8].view(np.uint32) >> 30, ph[both, 0, 8].view(np.uint32) >> 30      # flipped call: the solid's feature is fid2
    assert (kb == kh).mean() > 0.98 and (np.bincount(kh, minlength=4)[1:] > 50).all()
    s1 = g.integers(3, 9, n).astype(np.uint32)
    nr, cr, pr, sr = T.contact_manifolds(s1, p1, z, p2, 0.05, topology=topo)
    co, cs = T.contact(s1, p1, z, p2, 0.05)
    assert (sr == 0).all() and ((cr > 0) == (cs == 1)).all()
    b2 = cr > 0
    assert (pr[b2, 0, 6] == co[b2, 12]).all()


def test_hull_topology_builder_invariants():
    """harness/hull_topology.py (restatement of ConvexPolyhedron::from_convex_mesh over Qhull triangles): Euler's formula on the
    merged faces, outward unit normals, consistent vertex <-> face / edge adjacency; a cube gives 6 quads and 12 live edges."""
    from harness import hull_topology as ht
    he = np.array([0.7, 0.4, 1.1], np.float32)
    cube = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    t = ht.from_convex_mesh(cube, ht.hull_triangles(cube))
    assert len(t["face_first"]) == 6 and (t["face_count"] == 4).all() and t["num_edges"] == 18
    assert len(set(t["edges_adj_to_face"].tolist())) == 12          # the 6 face diagonals are deleted edges
    assert sorted(np.abs(t["face_normal"]).argmax(axis=1).tolist()) == [0, 0, 1, 1, 2, 2]
    pts, _ = scenes.hull_pool(6, 24, seed=77)
    for p in pts:
        t = ht.from_convex_mesh(p, ht.hull_triangles(p))
        nf, live = len(t["face_first"]), len(set(t["edges_adj_to_face"].tolist()))
        assert len(p) - live + nf == 2                                # V - E + F = 2 on the merged faces
        assert np.abs(np.linalg.norm(t["face_normal"], axis=1) - 1).max() < 1e-5
        for f in range(nf):
            vs = t["vertices_adj_to_face"][t["face_first"][f]:t["face_first"][f] + t["face_count"][f]]
            c = p[vs].mean(axis=0)
            assert np.dot(t["face_normal"][f], c) > 0                 # outward (the hull contains the origin)
            # coplanar within what the reference's merge rule lets through (adjacent triangle normals within ~1.5 degrees, chained)
            assert np.abs((p[vs] - c) @ t["face_normal"][f]).max() < 5e-2
        # vertex side: every (vertex, face) incidence appears exactly once, with the face's edge leaving that vertex
        assert int(t["vert_count"].sum()) == len(t["vertices_adj_to_face"])
        for v in range(len(p)):
            fs = t["faces_adj_to_vertex"][t["vert_first"][v]:t["vert_first"][v] + t["vert_count"][v]]
            assert len(fs) >= 3
            for f in fs:
                vs = t["vertices_adj_to_face"][t["face_first"][f]:t["face_first"][f] + t["face_count"][f]]
                assert v in vs
        assert np.abs(np.linalg.norm(t["edge_dir"], axis=1) - 1).max() < 1e-5


def test_manifold_try_update_contacts(oracle):
    """ContactManifold::try_update_contacts (contact_manifold.rs:652-699), oracle groundwork for manifold persistence: an unchanged
    pose keeps every manifold (values re-derived within a few ulps); a 1e-4 drift keeps most of them and the refreshed dists track a fresh computation; a
    5-degree turn drops them all; empty manifolds are never kept."""
    g = scenes.rng(23)
    T = oracle.ShapeTable([("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("cuboid", [0.5, 0.5, 0.5])])
    n = 4000
    s1, s2 = g.integers(0, 3, n).astype(np.uint32), g.integers(0, 3, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05)
    has = cnt > 0
    kept, q = oracle.ShapeTable.manifolds_try_update(p1, p2, nr, cnt, pts)
    assert (kept[has] == 1).all() and (kept[~has] == 0).all()
    # (dist and local_p1 are re-derived from local_p2 and the normal: same values up to a few ulps of the coordinates)
    assert np.abs(q[has] - pts[has])[:, :, :7].max() < 2e-5 and (q[:, :, 7:].view(np.uint32) == pts[:, :, 7:].view(np.uint32)).all()
    moved = p2.copy()
    moved[:, 4:] += (d * 1.0e-4).astype(np.float32)
    kept2, q2 = oracle.ShapeTable.manifolds_try_update(p1, moved, nr, cnt, pts)
    assert kept2[has].mean() > 0.9
    nr3, cnt3, pts3, _ = T.contact_manifolds(s1, p1, s2, moved, 0.05)
    k = (kept2 == 1) & (cnt3 == cnt)
    valid = np.arange(pts.shape[1])[None, :] < cnt[:, None]
    # (a fresh computation may pick another separating axis near a tie, hence a quantile and not the maximum)
    assert np.quantile(np.abs((q2[:, :, 6] - pts3[:, :, 6])[k][valid[k]]), 0.99) < 2e-4
    c5, s5 = np.cos(np.radians(2.5)), np.sin(np.radians(2.5))
    turn = np.array([s5, 0.0, 0.0, c5])                      # 5 degrees about x, composed onto every pose of shape 2
    def qmul(a, b):
        ai, aj, ak, aw = a[:, 0], a[:, 1], a[:, 2], a[:, 3]
        bi, bj, bk, bw = b
        return np.stack([aw * bi + ai * bw + aj * bk - ak * bj, aw * bj - ai * bk + aj * bw + ak * bi,
                         aw * bk + ai * bj - aj * bi + ak * bw, aw * bw - ai * bi - aj * bj - ak * bk], axis=1)
    turned = p2.copy()
    turned[:, :4] = qmul(p2[:, :4].astype(np.float64), turn).astype(np.float32)
    kept5, _ = oracle.ShapeTable.manifolds_try_update(p1, turned, nr, cnt, pts)
    assert kept5[has].mean() < 0.35      # only manifolds whose normal is (nearly) the turn axis survive


def test_compound_compound_contact_against_part_pairs(oracle):
    """Compound vs Compound (oracle groundwork, default_query_dispatcher.rs:338-351 nested through contact_shape_composite_shape): the
    result must be the closest of the contacts between all part pairs — computed here with plain query::contact on composed poses —
    and single identity-part compounds on both sides must reproduce the plain contact."""
    g = scenes.rng(31)
    pts, _ = scenes.hull_pool(4, 16, seed=32)
    spec = [("ball", 0.3), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])] + [("convex", p * 0.5) for p in pts]
    T = oracle.ShapeTable(spec)
    ns = len(spec)
    first, count, psid, ppose = [], [], [], []
    for c in range(16):
        k = int(g.integers(1, 4))
        first.append(len(psid)); count.append(k)
        psid += [int(x) for x in g.integers(0, ns, k)]
        ppose.append(np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.2], axis=1))
    first, count, psid = np.asarray(first, np.uint32), np.asarray(count, np.uint32), np.asarray(psid, np.uint32)
    ppose = np.concatenate(ppose).astype(np.float32)
    n = 1500
    a, b = g.integers(0, 16, n).astype(np.uint32), g.integers(0, 16, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 2.0 + 0.2)], axis=1).astype(np.float32)
    out, st, parts = T.contact_compound_compound(first, count, psid, ppose, a, p1, b, p2, 0.05, threads=4)
    assert 0.15 < (st == 1).mean() < 0.9

    def compose(p, q):       # world pose of a part: p * q, float64
        def rot(qt, v):
            u, w = qt[:3], qt[3]
            t = 2.0 * np.cross(u, v)
            return v + w * t + np.cross(u, t)
        pq, pt, qq, qt = p[:4].astype(np.float64), p[4:].astype(np.float64), q[:4].astype(np.float64), q[4:].astype(np.float64)
        w = pq[3] * qq[3] - np.dot(pq[:3], qq[:3])
        v = pq[3] * qq[:3] + qq[3] * pq[:3] + np.cross(pq[:3], qq[:3])
        return np.concatenate([v, [w], pt + rot(pq, qt)]).astype(np.float32)
    for k in range(0, n, 5):
        best = None
        for i in range(count[a[k]]):
            for j in range(count[b[k]]):
                gi, gj = first[a[k]] + i, first[b[k]] + j
                o, s = T.contact([psid[gi]], [compose(p1[k], ppose[gi])], [psid[gj]], [compose(p2[k], ppose[gj])], 0.05)
                if s[0] == 1 and (best is None or o[0, 12] < best):
                    best = o[0, 12]
        if best is None or st[k] != 1:
            # borderline pairs (dist within rounding of the prediction) may flip between the two ways of composing the poses
            assert (best is None) == (st[k] != 1) or abs((best if best is not None else out[k, 12]) - 0.05) < 1e-4
        else:
            assert abs(out[k, 12] - best) < 2e-5
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    cf, cc = np.arange(ns, dtype=np.uint32), np.ones(ns, np.uint32)
    o1, s1, _ = T.contact_compound_compound(cf, cc, np.arange(ns, dtype=np.uint32), np.tile(ident, (ns, 1)), a % ns, p1, b % ns, p2, 0.05)
    o2, s2 = T.contact(a % ns, p1, b % ns, p2, 0.05)
    assert (s1 == s2).all()
    np.testing.assert_allclose(o1[s2 == 1], o2[s2 == 1], rtol=1e-4, atol=2e-5)


def _two_frame_scene(oracle, n, seed, drift=2.0e-4):
    """Mixed Ball / Cuboid / ConvexPolyhedron pairs in two consecutive frames: half of the pairs drift by `drift`, the rest jump."""
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(4, 16, seed=seed + 1)
    spec = [("ball", 0.3), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2])] + [("convex", p * 0.5) for p in pts]
    T = oracle.ShapeTable(spec)
    s1, s2 = g.integers(0, len(spec), n).astype(np.uint32), g.integers(0, len(spec), n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 0.9 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    step = np.where(g.random((n, 1)) < 0.5, drift, 0.05)
    p2b = p2.copy()
    p2b[:, 4:] += (g.standard_normal((n, 3)) * step).astype(np.float32)
    return T, s1, s2, p1, p2, p2b


def test_persistent_manifold_dispatch(oracle):
    """QueryDispatcher::contact_manifolds on a second frame (contact_manifolds_cuboid_cuboid.rs:28, contact_manifolds_pfm_pfm.rs:63-66,
    match_contacts contact_manifold.rs:761-770): kept manifolds equal try_update_contacts of last frame's, everything else equals a
    first-frame computation of the new poses (GJK not seeded) — seeded from last frame's normal the pfm_pfm arm lands within GJK's
    own convergence tolerance of it —, ball pairs are never kept, and match pairs new points with old ones by feature ids."""
    T, s1, s2, p1, p2, p2b = _two_frame_scene(oracle, 6000, 41)
    topo = T.hull_topology()
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05, topology=topo)
    assert (st == 0).all() and 0.2 < (cnt > 0).mean() < 0.95
    nr2, cnt2, pts2, st2, kept, match = T.contact_manifolds_update(s1, p1, s2, p2b, 0.05, nr, cnt, pts, topology=topo)
    ball = (T.kinds[s1] == 0) | (T.kinds[s2] == 0)
    assert (kept[ball] == 0).all() and (kept[cnt == 0] == 0).all() and (st2 == 0).all()
    assert 0.15 < kept[~ball & (cnt > 0)].mean() < 0.85
    # kept: exactly ContactManifold::try_update_contacts of the old manifold
    k0, q = oracle.ShapeTable.manifolds_try_update(p1, p2b, nr, cnt, pts)
    assert (k0[~ball] == kept[~ball]).all()
    kp = kept == 1
    assert (pts2[kp].view(np.uint32) == q[kp].view(np.uint32)).all() and (cnt2[kp] == cnt[kp]).all()
    assert (nr2[kp].view(np.uint32) == nr[kp].view(np.uint32)).all()
    # recomputed: exactly the first-frame result at the new poses
    nr3, cnt3, pts3, _ = T.contact_manifolds(s1, p1, s2, p2b, 0.05, topology=topo)
    assert (cnt2[~kp] == cnt3[~kp]).all() and (pts2[~kp].view(np.uint32) == pts3[~kp].view(np.uint32)).all()
    assert (nr2[~kp].view(np.uint32) == nr3[~kp].view(np.uint32)).all()
    # match: identity on kept manifolds; on recomputed ones the last old point with both feature ids equal, else -1
    mp = pts.shape[1]
    valid2 = np.arange(mp)[None, :] < cnt2[:, None]
    assert (match[kp] == np.where(valid2[kp], np.arange(mp)[None, :], -1)).all() and (match[~valid2] == -1).all()
    f_old, f_new = pts[:, :, 7:].view(np.uint32), pts2[:, :, 7:].view(np.uint32)
    matched = 0
    for k in np.nonzero(~kp & (cnt2 > 0))[0][:1500]:
        for i in range(cnt2[k]):
            js = [j for j in range(cnt[k]) if (f_old[k, j] == f_new[k, i]).all()]
            assert match[k, i] == (js[-1] if js else -1)
            matched += bool(js)
    assert matched > 100
    # the reference seeds the pfm_pfm recomputation with last frame's normal: same manifolds up to GJK's convergence tolerance
    nr4, cnt4, pts4, st4, kept4, _ = T.contact_manifolds_update(s1, p1, s2, p2b, 0.05, nr, cnt, pts, topology=topo, seed_gjk=True)
    assert (kept4 == kept).all()
    pfm = ~ball & ~kp & ~((T.kinds[s1] == 1) & (T.kinds[s2] == 1))
    assert (pts4[~pfm].view(np.uint32) == pts2[~pfm].view(np.uint32)).all()
    both = pfm & (cnt4 > 0) & (cnt2 > 0)
    assert ((cnt4 > 0) == (cnt2 > 0))[pfm].mean() > 0.99 and both.sum() > 100
    deepest = lambda p, c: np.where(np.arange(mp)[None, :] < c[:, None], p[:, :, 6], np.inf).min(axis=1)
    assert np.quantile(np.abs(deepest(pts4[both], cnt4[both]) - deepest(pts2[both], cnt2[both])), 0.99) < 2e-3


def test_manifold_doc_examples(oracle):
    """The doc examples of query/contact_manifolds/contact_manifold.rs, the only places the reference pins manifold values:
    :270-297 (two unit balls 1.5 apart, prediction 0: one contact), :305-329 (2.1 apart, prediction 0.2: a predicted contact with
    dist > 0) and :719-758 (frames at 1.9 and 1.85: match_contacts hands the old point's ContactData to the new point, i.e. the
    new point matches old point 0 by its feature ids; ball manifolds are always recomputed, never kept)."""
    T = oracle.ShapeTable([("ball", 1.0)])
    z = np.zeros(3, np.uint32)
    ident = np.tile(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), (3, 1))
    p2 = ident.copy()
    p2[:, 4] = [1.5, 2.1, 1.9]
    pred = 0.0
    nr, cnt, pts, st = T.contact_manifolds(z[:1], ident[:1], z[:1], p2[:1], pred)
    assert st[0] == 0 and cnt[0] == 1 and pts[0, 0, 6] == np.float32(-0.5)
    assert (nr[0] == np.array([1, 0, 0, -1, 0, 0], np.float32)).all()
    assert (pts[0, 0, :6] == np.array([1, 0, 0, -1, 0, 0], np.float32)).all()
    nr, cnt, pts, st = T.contact_manifolds(z[1:2], ident[1:2], z[1:2], p2[1:2], 0.2)
    assert cnt[0] == 1 and pts[0, 0, 6] > 0 and abs(pts[0, 0, 6] - 0.1) < 1e-6
    nr, cnt, pts, st = T.contact_manifolds(z[1:2], ident[1:2], z[1:2], p2[1:2], 0.0)
    assert cnt[0] == 0                                              # without prediction the separated balls give no contact
    nr, cnt, pts, st = T.contact_manifolds(z[2:], ident[2:], z[2:], p2[2:], 0.0)
    assert cnt[0] == 1
    p2b = p2[2:].copy()
    p2b[:, 4] = 1.85
    rn, rc, rp, rs, rk, rm = T.contact_manifolds_update(z[2:], ident[2:], z[2:], p2b, 0.0, nr, cnt, pts)
    assert rk[0] == 0 and rc[0] == 1 and rm[0, 0] == 0 and (rm[0, 1:] == -1).all()
    assert abs(rp[0, 0, 6] - (1.85 - 2.0)) < 1e-6


def test_clip_empty_aabb_line(oracle):
    """query/clip/clip_aabb_line.rs:192-206, the reference's unit test of the function under Cuboid ray casts with normals: a
    zero direction clips to Some exactly when the origin lies in the (here degenerate) box."""
    assert oracle.clip_aabb_line([0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]) is not None
    assert oracle.clip_aabb_line([1, 1, 1], [2, 2, 2], [0, 0, 0], [0, 0, 0]) is None
    # and a hand-checkable clip: the x axis through the unit box enters at 1 and leaves at 2 from the origin
    nf = oracle.clip_aabb_line([1, -1, -1], [2, 1, 1], [0, 0, 0], [1, 0, 0])
    assert nf is not None and nf[0] == 1.0 and nf[1] == 2.0


def test_compound_trimesh_contact_against_parts(oracle):
    """Compound vs TriMesh in both orders (oracle groundwork for SURVEY §8 f2's last open item; default_query_dispatcher.rs:338-351
    nested through contact_shape_composite_shape): the result must be the closest of the contacts of the mesh with every part
    placed at its composed world pose (TriMesh-vs-shape, already checked against brute force), with that part and its triangle as
    the winners; the two orders are each other's flipped()."""
    g = scenes.rng(37)
    v, idx = scenes.terrain(17, 17)
    mesh = oracle.TriMesh(v, idx)
    lo, hi = np.asarray(v).min(axis=0), np.asarray(v).max(axis=0)
    pts, _ = scenes.hull_pool(4, 16, seed=38)
    sc = float((hi - lo)[:2].max()) / 16.0            # shapes about the size of a terrain cell
    spec = [("ball", 0.5 * sc), ("cuboid", [0.4 * sc, 0.6 * sc, 0.5 * sc])] + [("convex", np.asarray(p, np.float32) * 0.8 * sc) for p in pts]
    T = oracle.ShapeTable(spec)
    ns = len(spec)
    first, count, psid, ppose = [], [], [], []
    for c in range(12):
        k = int(g.integers(1, 4))
        first.append(len(psid)); count.append(k)
        psid += [int(x) for x in g.integers(0, ns, k)]
        ppose.append(np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 2.0 * sc], axis=1))
    first, count, psid = np.asarray(first, np.uint32), np.asarray(count, np.uint32), np.asarray(psid, np.uint32)
    ppose = np.concatenate(ppose).astype(np.float32)
    n = 600
    ids = g.integers(0, 12, n).astype(np.uint32)
    # compound centres scattered over the terrain, within a shape size of its surface on average
    up = int(np.argmin((hi - lo)))                    # the height axis of the generated terrain
    plane = [a for a in range(3) if a != up]
    ctr = np.zeros((n, 3))
    ctr[:, plane] = lo[plane] + g.random((n, 2)) * (hi - lo)[plane]
    ctr[:, up] = lo[up] + g.random(n) * ((hi - lo)[up] + 2.0 * sc) - 0.5 * sc
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), ctr], axis=1).astype(np.float32)
    mpose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    pred = 0.1 * sc
    out, st, parts = mesh.contact_compounds(mpose, T, first, count, psid, ppose, ids, poses, pred, threads=4, min_index_ties=True)
    assert 0.15 < (st == 1).mean() < 0.95

    def compose(p, q):
        def rot(qt, x):
            u, w = qt[:3], qt[3]
            t = 2.0 * np.cross(u, x)
            return x + w * t + np.cross(u, t)
        pq, pt, qq, qt = p[:4].astype(np.float64), p[4:].astype(np.float64), q[:4].astype(np.float64), q[4:].astype(np.float64)
        w = pq[3] * qq[3] - np.dot(pq[:3], qq[:3])
        vv = pq[3] * qq[:3] + qq[3] * pq[:3] + np.cross(pq[:3], qq[:3])
        return np.concatenate([vv, [w], pt + rot(pq, qt)]).astype(np.float32)
    checked = same_tri = 0
    for k in range(0, n, 3):
        best, bi, bt = None, None, None
        for i in range(count[ids[k]]):
            gi = first[ids[k]] + i
            o, s, t = mesh.contact_shapes(mpose, T, [psid[gi]], [compose(poses[k], ppose[gi])], pred, min_index_ties=True)
            if s[0] == 1 and (best is None or o[0, 12] < best):
                best, bi, bt = o[0, 12], i, t[0]
        if best is None or st[k] != 1:
            assert (best is None) == (st[k] != 1) or abs((best if best is not None else out[k, 12]) - pred) < 1e-4 * sc + 1e-5
        else:
            assert abs(out[k, 12] - best) < 2e-5 * max(1.0, sc)
            same_tri += int(parts[k, 0] == bi and parts[k, 1] == bt)
            checked += 1
    # (a shape resting on two triangles has two equal dists; composing the poses in float64 here can tip such ties the other way)
    assert checked > 30 and same_tri > 0.85 * checked
    # the other order is the flipped contact (same dist; points and normals swapped)
    out2, st2, parts2 = mesh.contact_compounds(mpose, T, first, count, psid, ppose, ids, poses, pred, trimesh_first=True, threads=4,
                                               min_index_ties=True)
    both = (st == 1) & (st2 == 1)
    assert (st == st2).mean() > 0.99 and both.sum() > 50
    np.testing.assert_allclose(out2[both][:, 12], out[both][:, 12], rtol=0, atol=3e-5 * max(1.0, sc))
    agree = both & (parts == parts2).all(axis=1)
    assert agree.sum() > 0.9 * both.sum()
    np.testing.assert_allclose(out2[agree][:, 0:3], out[agree][:, 3:6], rtol=0, atol=2e-4 * max(1.0, sc))
    np.testing.assert_allclose(out2[agree][:, 6:9], out[agree][:, 9:12], rtol=0, atol=2e-3)


def test_trimesh_trimesh_toi_issue_194(oracle):
    """crates/parry3d/tests/geometry/trimesh_trimesh_toi.rs: two pyramids 1000 apart, one moving at 100000 along x:
    `assert_eq!(time_of_impact, Some(0.00998))`, exact. Pins the restatement of the composite shape casts
    (shape_cast_composite_shape_shape.rs: find_best over Minkowski-summed node boxes, nested through cast_shapes_shape_composite_shape),
    oracle groundwork for SURVEY §8 f3's open item. Also: a mesh against a plain shape equals the earliest of the per-triangle casts,
    and the mesh as second shape gives the swapped hit."""
    pts = np.array([[0, 1, 0], [-1, -0.5, 0], [0, -0.5, -1], [1, -0.5, 0]], np.float32)
    idx = np.array([[0, 1, 2], [0, 2, 3], [0, 3, 1]], np.uint32)
    a, b = oracle.TriMesh(pts, idx), oracle.TriMesh(pts, idx)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    r = a.cast_shapes(ident, [100000.0, 0, 0], _pose([1000.0, 0, 0]), [0, 0, 0], other_mesh=b)
    assert r is not None and r[1] == 1 and r[0][12] == np.float32(0.00998)
    assert (r[0][:6] == np.array([1, -0.5, 0, -1, -0.5, 0], np.float32)).all()          # the two base corners that meet
    # moving away: None
    assert a.cast_shapes(ident, [-100000.0, 0, 0], _pose([1000.0, 0, 0]), [0, 0, 0], other_mesh=b) is None
    # mesh vs plain shapes: the earliest per-triangle cast (triangles as 3-point hulls differ from shape::Triangle only on ties)
    g = scenes.rng(43)
    hp, _ = scenes.hull_pool(2, 16, seed=44)
    spec = [("ball", 0.3), ("cuboid", [0.3, 0.2, 0.4])] + [("convex", np.asarray(p, np.float32) * 0.4) for p in hp]
    tris = [pts[t] for t in idx]
    T = oracle.ShapeTable(spec + [("convex", t) for t in tris])
    n_some = 0
    for k in range(120):
        sid = int(g.integers(0, len(spec)))
        d = g.standard_normal(3); d /= np.linalg.norm(d)
        pose = np.concatenate([scenes.random_unit_quaternions(g, 1)[0], d * 4.0]).astype(np.float32)
        vel = (-d * 3.0 + g.standard_normal(3) * 0.6).astype(np.float32)
        r = a.cast_shapes(ident, [0, 0, 0], pose, vel, table=T, shape=sid)
        per = [T.cast_shapes([len(spec) + t], [ident], [[0, 0, 0]], [sid], [pose], [vel]) for t in range(3)]
        tois = [o[0, 12] for o, s in per if s[0] != 0]
        assert (r is None) == (len(tois) == 0)
        if r is not None:
            n_some += 1
            assert abs(r[0][12] - min(tois)) <= 1e-5 * max(1.0, min(tois))
            r2 = a.cast_shapes(ident, [0, 0, 0], pose, vel, table=T, shape=sid, mesh_second=True)
            # (the two orders run different GJK ray casts, each converged to gjk.rs's relative tolerance sqrt(10 eps) ~ 1e-3)
            assert r2 is not None and abs(r2[0][12] - r[0][12]) <= 1e-3 * max(1.0, r[0][12])
            np.testing.assert_allclose(r2[0][0:3], r[0][3:6], atol=5e-3)
            np.testing.assert_allclose(r2[0][3:6], r[0][0:3], atol=5e-3)
    assert n_some > 30


def test_getting_started_example(oracle):
    """crates/parry3d/examples/getting_started.rs: a ray from (0, 0, -1) along +z intersects the unit cube (origin on its face:
    solid cast, toi 0)."""
    FMAX = float(np.finfo(np.float32).max)
    toi = oracle.shape_cast_ray_toi(1, [1.0, 1.0, 1.0], None, [0, 0, -1, 0, 0, 1], FMAX, solid=True)
    assert toi is not None and toi == 0.0
    hit = oracle.shape_cast_ray(1, [1.0, 1.0, 1.0], None, [0, 0, -3, 0, 0, 1], FMAX, solid=True)
    assert hit is not None and hit[0] == 2.0 and tuple(hit[1]) == (0.0, 0.0, -1.0)


def test_ray_and_shape_cast_doc_examples(oracle):
    """Doc examples with exact values: query/ray/ray.rs:40-66 (ball at x = 5, toi == 4.0), :255-285 (cuboid from x = -5: toi == 4.0,
    normal == -x), query/shape_cast/shape_cast.rs:196-252 (unit balls 10 apart at speed 2: time_of_impact == 4.0; overlapping
    balls: 0.0 with PenetratingOrWithinTargetDist)."""
    toi = oracle.shape_cast_ray_toi(0, [1.0], _pose([5, 0, 0]), [0, 0, 0, 1, 0, 0], 100.0, solid=True)
    assert toi == 4.0
    hit = oracle.shape_cast_ray(1, [1.0, 1.0, 1.0], None, [-5, 0, 0, 1, 0, 0], 100.0, solid=True)
    assert hit is not None and hit[0] == 4.0 and tuple(hit[1]) == (-1.0, 0.0, 0.0)
    T = oracle.ShapeTable([("ball", 1.0), ("ball", 2.0)])
    out, st = T.cast_shapes([0, 1], [_pose([0, 0, 0])] * 2, [[2, 0, 0], [1, 0, 0]], [0, 1], [_pose([10, 0, 0]), _pose([3, 0, 0])], [[0, 0, 0]] * 2)
    assert st[0] == 1 and out[0, 12] == 4.0
    assert st[1] == 2 and out[1, 12] == 0.0


def test_bvh_traverse_doc_examples(oracle):
    """partitioning/bvh/bvh_traverse.rs:170-243: three unit boxes at x = 0, 5, 10; the region [-1, 7] x [-1, 2]^2 reaches the first
    two leaves (count == 2), and the point (5.5, 0.5, 0.5) lies in leaf 1 (a degenerate box query is the same prune rule)."""
    boxes = np.array([[0, 0, 0, 1, 1, 1], [5, 0, 0, 6, 1, 1], [10, 0, 0, 11, 1, 1]], np.float32)
    for strategy in (0, 1):
        bvh = oracle.Bvh(boxes, strategy)
        offs, ids = bvh.intersect_aabbs(np.array([[-1, -1, -1, 7, 2, 2], [5.5, 0.5, 0.5, 5.5, 0.5, 0.5]], np.float32))
        assert offs[1] - offs[0] == 2 and sorted(ids[offs[0]:offs[1]]) == [0, 1]
        assert offs[2] - offs[1] == 1 and ids[offs[1]] == 1


def test_solid_point_query_example_through_the_ball_arm(oracle):
    """crates/parry3d/examples/solid_point_query3d.rs pins Cuboid(1, 2, 2) point projection: the origin is at (non-solid) distance
    -1.0, the point (2, 2, 2) at 1.0. The contact path reaches the same function (point_aabb.rs:9-132, via
    project_local_point_and_get_feature) in its ball-convex arm, so a Ball of radius 0 at those points must report exactly these
    distances (contact_ball_convex_polyhedron.rs:37-51: dist = -len - r inside, len - r outside)."""
    T = oracle.ShapeTable([("cuboid", [1.0, 2.0, 2.0]), ("ball", 0.0)])
    ident = _pose([0, 0, 0])
    out, st = T.contact([0, 0, 1], [ident, ident, _pose([2, 2, 2])], [1, 1, 0], [ident, _pose([2, 2, 2]), ident], 2.0)
    assert (st == 1).all()
    assert out[0, 12] == -1.0 and out[1, 12] == 1.0 and out[2, 12] == 1.0          # third: the ball-first (flipped) arm
    assert tuple(out[1, 0:3]) == (1.0, 2.0, 2.0)                                    # the projection on the cuboid


def test_trimesh_distance_closed_forms_and_brute_force(oracle):
    """query::distance with a TriMesh (distance_composite_shape_shape.rs:13-77): closed forms over a flat mesh, and on a terrain the
    minimum over every triangle of the per-triangle distance (triangles as 3-point hulls in the plain pair path: same GJK for cuboids
    and hulls; balls go through the triangle's own projection instead of GJK, equal to rounding)."""
    v = np.array([[-4, 0, -4], [4, 0, -4], [4, 0, 4], [-4, 0, 4]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)
    om = oracle.TriMesh(v, idx)
    T = oracle.ShapeTable([("ball", 0.5), ("cuboid", [0.5, 0.25, 0.5])])
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    poses = np.array([[0, 0, 0, 1, 0.3, 2.0, -0.2], [0, 0, 0, 1, 1.0, 1.0, 1.0], [0, 0, 0, 1, 0.0, 0.1, 0.0], [0, 0, 0, 1, 6.0, 0.0, 0.0]], np.float32)
    d, part = om.distance_shapes(ident, T, [0, 1, 0, 0], poses)
    np.testing.assert_allclose(d, [1.5, 0.75, 0.0, 1.5], atol=1e-6)
    d2, _ = om.distance_shapes(ident, T, [0, 1, 0, 0], poses, mesh_second=True)
    np.testing.assert_allclose(d2, d, atol=1e-6)
    g = scenes.rng(51)
    tv, tidx = scenes.terrain(17, 17, extent=12.0)
    tv = tv.copy(); tv[:, 1] *= 0.2
    om = oracle.TriMesh(tv, tidx)
    hp, _ = scenes.hull_pool(2, 16, seed=52)
    spec = [("ball", 0.3), ("cuboid", [0.3, 0.2, 0.4])] + [("convex", np.asarray(p, np.float32) * 0.4) for p in hp]
    T = oracle.ShapeTable(spec + [("convex", tv[t]) for t in tidx])
    n = 60
    sid = g.integers(0, len(spec), n).astype(np.uint32)
    t = np.stack([(g.random(n) - 0.5) * 12, g.random(n) * 6 - 1.0, (g.random(n) - 0.5) * 12], axis=1)
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    d, part = om.distance_shapes(ident, T, sid, poses, threads=4)
    nt = len(tidx)
    for k in range(n):
        tri_ids = np.arange(len(spec), len(spec) + nt, dtype=np.uint32)
        bd, _ = T.distance(tri_ids, np.tile(ident, (nt, 1)), np.full(nt, sid[k], np.uint32), np.tile(poses[k], (nt, 1)))
        assert abs(bd.min() - d[k]) < 2e-5, (k, sid[k], bd.min(), d[k])
        if d[k] > 0 and (bd == bd.min()).sum() == 1 and sid[k] != 0:
            assert part[k] == int(np.argmin(bd))
    assert (d > 0).mean() > 0.4 and (d == 0).mean() > 0.05


def test_bvh_project_point_typed_leaves_closed_forms(oracle):
    """Bvh::project_point (bvh_queries.rs:213-227) over typed leaves: closed forms for a ball and a cuboid leaf, solid and not, and the
    max_distance cut-off (strict: find_best only accepts costs below it)."""
    I = [0.0, 0.0, 0.0, 1.0]
    poses = np.array([I + [0, 0, 0], I + [5, 0, 0]], np.float32)
    kinds = np.array([0, 1], np.uint8)
    params = np.array([[1.0, 0, 0], [0.5, 1.0, 1.0]], np.float32)
    aabbs = np.array([[-1, -1, -1, 1, 1, 1], [4.5, -1, -1, 5.5, 1, 1]], np.float32)
    b = oracle.Bvh(aabbs)
    pts = np.array([[3, 0, 0], [0.5, 0, 0], [5.1, 0.2, 0.3], [2.0, 0, 0], [-4, 0, 0]], np.float32)
    proj, inside, leaf = b.project_points_shapes(kinds, params, poses, pts, float(np.finfo(np.float32).max), solid=True)
    assert leaf.tolist() == [1, 0, 1, 0, 0] and inside.tolist() == [0, 1, 1, 0, 0]
    np.testing.assert_allclose(proj, [[4.5, 0, 0], [0.5, 0, 0], [5.1, 0.2, 0.3], [1, 0, 0], [-1, 0, 0]], atol=1e-6)
    proj, inside, leaf = b.project_points_shapes(kinds, params, poses, pts, float(np.finfo(np.float32).max), solid=False)
    assert inside.tolist() == [0, 1, 1, 0, 0]
    np.testing.assert_allclose(proj[1], [1, 0, 0], atol=1e-6)            # pushed to the sphere
    np.testing.assert_allclose(proj[2], [5.5, 0.2, 0.3], atol=1e-6)      # pushed to the nearest face
    proj, inside, leaf = b.project_points_shapes(kinds, params, poses, pts, 1.5, solid=True)
    assert leaf.tolist() == [0xFFFFFFFF, 0, 1, 0, 0xFFFFFFFF]            # 1.5 away is not < 1.5; 3 away is cut

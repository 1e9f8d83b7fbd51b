"""GPU parity of the closed-form contact-manifold arms (ball-ball, ball-cuboid, cuboid-cuboid SAT + face
clipping) through pb2_contact_manifolds_batch against the CPU oracle: statuses, point counts and feature ids exact, normals /
points / distances within 1e-5 (bit-identical in practice)."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu


def make_scene(n, seed):
    g = scenes.rng(seed)
    spec = [("ball", 0.4), ("ball", 0.25), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("cuboid", [0.5, 0.5, 0.5])]
    s1, s2 = g.integers(0, 5, n).astype(np.uint32), g.integers(0, 5, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.3 + 0.2)], axis=1).astype(np.float32)
    p2[::5, :4] = p1[::5, :4]          # same orientation: face-face stacks with 4 - 8 points
    p2[::25, 4:] = p1[::25, 4:] + np.array([0.0, 0.7, 0.0], np.float32)   # exactly aligned offsets
    p2[::50, 4:] = p1[::50, 4:]        # coincident centres (degenerate normals)
    return spec, s1, p1, s2, p2


def tables(ctx, oracle, spec):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v)
                                for k, v in spec])
    return T, G


@pytest.mark.parametrize("prediction", [0.05, 0.3])
def test_manifolds_vs_oracle(ctx, oracle, prediction):
    import parry_b200
    spec, s1, p1, s2, p2 = make_scene(30000, seed=7)
    T, G = tables(ctx, oracle, spec)
    rn, rc, rp, rs = T.contact_manifolds(s1, p1, s2, p2, prediction, max_points=8, threads=8)
    gn, gc, gp, gs = parry_b200.contact_manifolds(G, s1, p1, s2, p2, prediction, max_points=8)
    assert (rs == 0).all() and (gs == rs).all()
    assert (gc == rc).all(), np.nonzero(gc != rc)[0][:10]
    assert (rc > 0).mean() > 0.4 and (rc >= 4).mean() > 0.1 and rc.max() >= 6
    assert (gp[:, :, 7:].view(np.uint32) == rp[:, :, 7:].view(np.uint32)).all()      # feature ids, in the reference's point order
    np.testing.assert_allclose(gn, rn, rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp[:, :, :7], rp[:, :, :7], rtol=1e-5, atol=2e-6)


def test_manifold_statuses_and_capacity(ctx, oracle):
    import torch
    import parry_b200
    spec, s1, p1, s2, p2 = make_scene(5000, seed=9)
    pts, _ = scenes.hull_pool(1, 16, seed=3)
    spec = spec + [("convex", pts[0])]
    T, G = tables(ctx, oracle, spec)
    s1 = s1.copy()
    s1[::7] = 5                                 # a ConvexPolyhedron: pfm arm, not built
    gn, gc, gp, gs = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=8)
    rn, rc, rp, rs = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=8)
    assert (gs == rs).all() and (gs[::7] == 2).all() and (gc[::7] == 0).all() and (gc == rc).all()
    # capacity: with room for 4 points, richer manifolds report status 4 and keep the first four points
    g4 = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=4)
    over = rc > 4
    assert over.sum() > 50 and (g4[3][over] == 4).all() and (g4[1][over] == 4).all() and (g4[3][~over] == gs[~over]).all()
    assert (g4[2][over].view(np.uint32) == gp[over][:, :4].view(np.uint32)).all()
    # device-resident inputs give the same bits
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    dn, dc, dp, ds = parry_b200.contact_manifolds(G, dev(s1), dev(p1), dev(s2), dev(p2), 0.05, max_points=8)
    ctx.synchronize()
    assert (dp.cpu().numpy().view(np.uint32) == gp.view(np.uint32)).all() and (dc.cpu().numpy().view(np.uint32) == gc).all()


def test_pfm_manifolds_vs_oracle(ctx, oracle):
    """contact_manifolds_pfm_pfm.rs:42-162 on the GPU: pairs of 16-point hulls, hulls of cuboid corners (quad faces) and cuboids,
    with the hull topology supplied through pb2_shapes_set_hull_topology. Counts, statuses and feature ids exact, values 1e-5."""
    import parry_b200
    g = scenes.rng(13)
    pts, _ = scenes.hull_pool(12, 16, seed=14)
    hes = [np.array([0.3, 0.5, 0.4], np.float32), np.array([0.6, 0.2, 0.2], np.float32)]
    corners = lambda he: np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * he
    spec = [("ball", 0.3), ("cuboid", hes[0]), ("cuboid", hes[1]), ("convex", corners(hes[0])), ("convex", corners(hes[1]))]
    spec += [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    T, G = tables(ctx, oracle, spec)
    topo = T.hull_topology()
    G.set_hull_topology(topo)
    n = 20000
    ns = len(spec)
    s1, s2 = g.integers(0, ns, n).astype(np.uint32), g.integers(0, ns, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.2 + 0.2)], axis=1).astype(np.float32)
    p2[::4, :4] = p1[::4, :4]
    rn, rc, rp, rs = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=12, threads=8, topology=topo)
    gn, gc, gp, gs = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=12)
    kinds = np.array([0 if k == "ball" else 1 if k == "cuboid" else 2 for k, _ in spec])
    pfm = ((kinds[s1] == 2) | (kinds[s2] == 2)) & (kinds[s1] != 0) & (kinds[s2] != 0)
    ballhull = ((kinds[s1] == 2) | (kinds[s2] == 2)) & ~pfm      # contact_manifolds_convex_ball.rs with a ConvexPolyhedron
    assert pfm.mean() > 0.5 and (rs == 0).all() and (rc[pfm] > 0).mean() > 0.3
    assert ballhull.sum() > 1000 and (rc[ballhull] > 0).mean() > 0.2 and (rc[ballhull] <= 1).all()
    hull_feat = np.where(kinds[s1[ballhull]] == 2, rp[ballhull, 0, 7].view(np.uint32), rp[ballhull, 0, 8].view(np.uint32))[rc[ballhull] > 0] >> 30
    assert (np.bincount(hull_feat, minlength=4)[1:] > 10).all()      # vertex, edge and face features all occur
    host = gs == 3                                   # none: EPA runs beyond the hot arena go to the overflow kernel
    assert host.sum() == 0
    ok = ~host
    assert (gs[ok] == rs[ok]).all() and (gc[ok] == rc[ok]).all(), np.nonzero((gc != rc) & ok)[0][:10]
    assert (gp[ok][:, :, 7:].view(np.uint32) == rp[ok][:, :, 7:].view(np.uint32)).all()
    np.testing.assert_allclose(gn[ok], rn[ok], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp[ok][:, :, :7], rp[ok][:, :, :7], rtol=1e-5, atol=2e-6)
    # without topology the same pairs are status 2
    G2 = tables(ctx, oracle, spec)[1]
    _, c2, _, st2 = parry_b200.contact_manifolds(G2, s1, p1, s2, p2, 0.05, max_points=12)
    assert (st2[pfm | ballhull] == 2).all() and (c2[pfm | ballhull] == 0).all() and (st2[~pfm & ~ballhull] == 0).all()

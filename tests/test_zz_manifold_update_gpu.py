"""GPU parity of manifold persistence: ContactManifold::try_update_contacts through pb2_manifolds_try_update and the
second-frame dispatch (try to keep, else recompute, match_contacts) through pb2_contact_manifolds_update_batch, against the CPU
oracle on the same last-frame manifolds: kept flags, counts, feature ids and match indices exact; kept manifolds bit for bit;
recomputed ones within 1e-5 like the first-frame arms. (Written after this round's GPU budget was spent: the function the
try-update kernel calls is checked on the CPU by tests/test_hostcheck.py; the kernels themselves first run on hardware here.)"""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu


def two_frames(oracle, n, seed, hulls=True):
    g = scenes.rng(seed)
    spec = [("ball", 0.3), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("cuboid", [0.5, 0.5, 0.5])]
    if hulls:
        pts, _ = scenes.hull_pool(6, 16, seed=seed + 1)
        spec += [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    ns = len(spec)
    s1, s2 = g.integers(0, ns, n).astype(np.uint32), g.integers(0, ns, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    p2b = p2.copy()
    p2b[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.5, 2.0e-4, 0.05)).astype(np.float32)
    return spec, s1, p1, s2, p2, p2b


def tables(ctx, oracle, spec):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v)
                                for k, v in spec])
    return T, G


def test_try_update_vs_oracle(ctx, oracle):
    import torch
    import parry_b200
    spec, s1, p1, s2, p2, p2b = two_frames(oracle, 40000, 61, hulls=False)
    T, G = tables(ctx, oracle, spec)
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=8)
    kept_o, q_o = oracle.ShapeTable.manifolds_try_update(p1, p2b, nr, cnt, pts)
    kept_g, q_g = parry_b200.manifolds_try_update(ctx, p1, p2b, nr, cnt, pts)
    assert 0.2 < kept_o[cnt > 0].mean() < 0.9
    assert (kept_g == kept_o).all()
    assert (q_g.view(np.uint32) == q_o.view(np.uint32)).all()
    # unchanged poses keep every non-empty manifold; a custom threshold of cos(0) = 1 + eps keeps none
    kept_same, _ = parry_b200.manifolds_try_update(ctx, p1, p2, nr, cnt, pts)
    assert (kept_same == (cnt > 0)).all()
    kept_none, _ = parry_b200.manifolds_try_update(ctx, p1, p2, nr, cnt, pts, angle_dot_threshold=1.5)
    assert (kept_none == 0).all()
    # device-resident arrays: same bits, caller's arrays untouched
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    dpts = dev(pts)
    kept_d, q_d = parry_b200.manifolds_try_update(ctx, dev(p1), dev(p2b), dev(nr), dev(cnt), dpts)
    ctx.synchronize()
    assert (kept_d.cpu().numpy() == kept_o).all() and (q_d.cpu().numpy().view(np.uint32) == q_o.view(np.uint32)).all()
    assert (dpts.cpu().numpy().view(np.uint32) == pts.view(np.uint32)).all()


@pytest.mark.parametrize("hulls", [False, True])
def test_manifolds_update_vs_oracle(ctx, oracle, hulls):
    import parry_b200
    spec, s1, p1, s2, p2, p2b = two_frames(oracle, 30000, 71, hulls=hulls)
    T, G = tables(ctx, oracle, spec)
    topo = None
    if hulls:
        topo = T.hull_topology()
        G.set_hull_topology(topo)
    mp = 12
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=mp)
    first_ok = st == 0
    assert first_ok.mean() > 0.999
    # the reference's persistent dispatch, GJK seed included: a recomputed pfm_pfm pair starts GJK from last frame's manifold normal —
    # or from the direction an empty manifold cached when GJK answered NoIntersection (contact_manifolds_pfm_pfm.rs:63-66,151-154)
    rn, rc, rp, rs, rk, rm = T.contact_manifolds_update(s1, p1, s2, p2b, 0.05, nr, cnt, pts, threads=8, topology=topo, seed_gjk=True)
    gn, gc, gp, gs, gk, gm = parry_b200.contact_manifolds_update(G, s1, p1, s2, p2b, 0.05, nr, cnt, pts)
    ok = first_ok & (gs != 3)
    assert (~ok).sum() == 0
    kinds = T.kinds
    ball = (kinds[s1] == 0) | (kinds[s2] == 0)
    assert (gk[ok] == rk[ok]).all(), np.nonzero((gk != rk) & ok)[0][:10]
    assert (gk[ball] == 0).all() and 0.15 < rk[~ball & (cnt > 0)].mean() < 0.85
    assert (gs[ok] == rs[ok]).all() and (gc[ok] == rc[ok]).all(), np.nonzero((gc != rc) & ok)[0][:10]
    assert (gp[ok][:, :, 7:].view(np.uint32) == rp[ok][:, :, 7:].view(np.uint32)).all()
    assert (gm[ok] == rm[ok]).all()
    kp = ok & (rk == 1)
    assert (gp[kp].view(np.uint32) == rp[kp].view(np.uint32)).all() and (gn[kp].view(np.uint32) == rn[kp].view(np.uint32)).all()
    np.testing.assert_allclose(gn[ok], rn[ok], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp[ok][:, :, :7], rp[ok][:, :, :7], rtol=1e-5, atol=2e-6)
    # the recomputed part equals a first-frame computation at the new poses, bit for bit — for the arms without a GJK seed (a seeded
    # GJK stops at another iterate within its sqrt(10 eps) tolerance)
    fn, fc, fp, fs = parry_b200.contact_manifolds(G, s1, p1, s2, p2b, 0.05, max_points=mp)
    pfm = (kinds[s1] != 0) & (kinds[s2] != 0) & ((kinds[s1] == 2) | (kinds[s2] == 2))
    rec = (gk == 0) & (gs != 3) & (fs != 3) & ~pfm
    if hulls:
        # and the seed does matter: some recomputed hull pairs differ from the unseeded first-frame answer, all within GJK's tolerance
        both = (gk == 0) & pfm & (gc > 0) & (fc == gc)
        assert both.sum() > 100
        d_seeded, d_plain = gp[both][:, :, 6].min(axis=1), fp[both][:, :, 6].min(axis=1)
        assert np.abs(d_seeded - d_plain).max() < 5e-3
        # empty pfm manifolds carry the cached NoIntersection direction (unit length) instead of a zero normal
        empty = ok & pfm & (gc == 0) & (gk == 0)
        assert empty.sum() > 100 and np.allclose(np.linalg.norm(gn[empty][:, :3], axis=1), 1.0, atol=1e-5) and (gn[empty][:, 3:] == 0).all()
    assert (gc[rec] == fc[rec]).all() and (gs[rec] == fs[rec]).all() and (gp[rec].view(np.uint32) == fp[rec].view(np.uint32)).all()
    # without the match output
    out = parry_b200.contact_manifolds_update(G, s1, p1, s2, p2b, 0.05, nr, cnt, pts, with_match=False)
    assert out[5] is None and (out[4] == gk).all() and (out[2].view(np.uint32) == gp.view(np.uint32)).all()


def test_manifolds_update_device_resident(ctx, oracle):
    import torch
    import parry_b200
    spec, s1, p1, s2, p2, p2b = two_frames(oracle, 20000, 81, hulls=True)
    T, G = tables(ctx, oracle, spec)
    G.set_hull_topology(T.hull_topology())
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, s1, p1, s2, p2, 0.05, max_points=12)
    host = parry_b200.contact_manifolds_update(G, s1, p1, s2, p2b, 0.05, nr, cnt, pts)
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    dnr, dcnt, dpts = dev(nr), dev(cnt), dev(pts)
    d = parry_b200.contact_manifolds_update(G, dev(s1), dev(p1), dev(s2), dev(p2b), 0.05, dnr, dcnt, dpts)
    ctx.synchronize()
    for a, b in zip(host, d):
        b = b.cpu().numpy()
        assert (a.view(np.uint8) == b.view(np.uint8)).all()
    assert (dpts.cpu().numpy().view(np.uint32) == pts.view(np.uint32)).all()     # the wrapper works on private copies
    # empty batch and argument checks
    z = np.zeros((0, 7), np.float32)
    e = parry_b200.contact_manifolds_update(G, np.zeros(0, np.uint32), z, np.zeros(0, np.uint32), z, 0.05, np.zeros((0, 6), np.float32),
                                            np.zeros(0, np.uint32), np.zeros((0, 12, 9), np.float32))
    assert len(e[1]) == 0


def test_match_contacts_doc_example(ctx):
    """contact_manifold.rs:719-758, the reference's own pin for match_contacts: unit balls at 1.9 then 1.85 apart; the recomputed
    point carries the old point's feature ids, so it matches old point 0 (whose ContactData the reference hands over)."""
    import parry_b200
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(1.0)])
    z = np.zeros(1, np.uint32)
    ident = np.array([[0, 0, 0, 1, 0, 0, 0]], np.float32)
    f1, f2 = ident.copy(), ident.copy()
    f1[0, 4], f2[0, 4] = 1.9, 1.85
    nr, cnt, pts, st = parry_b200.contact_manifolds(G, z, ident, z, f1, 0.0, max_points=4)
    assert cnt[0] == 1
    gn, gc, gp, gs, gk, gm = parry_b200.contact_manifolds_update(G, z, ident, z, f2, 0.0, nr, cnt, pts)
    assert gk[0] == 0 and gc[0] == 1 and gm[0, 0] == 0 and (gm[0, 1:] == -1).all()
    assert abs(gp[0, 0, 6] - (1.85 - 2.0)) < 1e-6


def test_gpu_reproduces_second_frame_golden(ctx):
    """tests/golden/second_frame_2000.npz (frozen oracle output, tests/golden/make_golden.py second_frame) on the pfm golden's scene."""
    import os
    import parry_b200
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    z, w = np.load(os.path.join(gd, "manifolds_pfm_2000.npz")), np.load(os.path.join(gd, "second_frame_2000.npz"))
    pu = z["params"].view(np.uint32)
    spec = [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p[:3]) if k == 1 else parry_b200.ConvexPolyhedron(z["points"][u[0]:u[0] + u[1]])
            for k, p, u in zip(z["kinds"], z["params"], pu)]
    G = parry_b200.Shapes(ctx, spec)
    G.set_hull_topology({k[5:]: z[k] for k in z.files if k.startswith("topo_")})
    gn, gc, gp, gs, gk, gm = parry_b200.contact_manifolds_update(G, z["shape1"], z["pos1"], z["shape2"], w["moved_pos2"], 0.05, z["normals"],
                                                                 z["counts"], z["man_points"])
    ok = gs != 3
    assert (~ok).sum() <= 2
    assert (gk[ok] == w["kept"][ok]).all() and (gc[ok] == w["counts"][ok]).all() and (gs[ok] == w["status"][ok]).all()
    assert (gm[ok] == w["match"][ok]).all()
    assert (gp[ok][:, :, 7:].view(np.uint32) == w["man_points"][ok][:, :, 7:].view(np.uint32)).all()
    kp = ok & (gk == 1)
    assert (gp[kp].view(np.uint32) == w["man_points"][kp].view(np.uint32)).all()
    np.testing.assert_allclose(gn[ok], w["normals"][ok], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(gp[ok][:, :, :7], w["man_points"][ok][:, :, :7], rtol=1e-5, atol=2e-6)

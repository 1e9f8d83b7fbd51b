"""GPU parity of query::contact between two Compounds (the composite arm nested through contact_shape_composite_shape)
through pb2_compound_contact_compounds against the CPU oracle: statuses and winning parts exact, contacts within 1e-5. (Written
after this round's GPU budget was spent: the per-pair candidate / reduction functions the kernels call are checked on the CPU by
tests/test_hostcheck.py; the kernels themselves first run on hardware here.)"""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu


def make_scene(n, seed, n_compounds=48):
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(8, 16, seed=seed + 1)
    spec = [("ball", 0.3), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])] + [("convex", np.asarray(p, np.float32) * 0.5) for p in pts]
    ns = len(spec)
    compounds = []
    for c in range(n_compounds):
        k = int(g.integers(1, 6))
        poses = np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.4], axis=1).astype(np.float32)
        compounds.append([(poses[i], int(g.integers(0, ns))) for i in range(k)])
    a, b = g.integers(0, n_compounds, n).astype(np.uint32), g.integers(0, n_compounds, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - 0.5) * 4], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 2.6 + 0.1)], axis=1).astype(np.float32)
    return spec, compounds, a, p1, b, p2


def tables(ctx, oracle, spec, compounds):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v)
                                for k, v in spec])
    C = parry_b200.Compounds(ctx, G, compounds)
    return T, G, C


def test_compound_compound_vs_oracle(ctx, oracle):
    spec, compounds, a, p1, b, p2 = make_scene(20000, seed=101)
    T, G, C = tables(ctx, oracle, spec, compounds)
    ro, rs, rp = T.contact_compound_compound(C.first, C.count, C.part_shape, C.part_pose, a, p1, b, p2, 0.05, threads=8)
    go, gs, gp = C.contact_compounds(a, p1, b, p2, 0.05)
    assert 0.2 < (rs == 1).mean() < 0.9
    host = gs == 3                                   # none: EPA runs beyond the hot arena go to the overflow kernel
    assert host.sum() == 0
    ok = ~host
    assert (gs[ok] == rs[ok]).all(), np.nonzero((gs != rs) & ok)[0][:10]
    some = ok & (rs == 1)
    assert (gp[ok & (rs != 1)] == 0xFFFFFFFF).all() and (go[ok & (rs != 1)] == 0).all()
    same = (gp[some] == rp[some]).all(axis=1)
    assert same.mean() > 0.999
    np.testing.assert_allclose(go[some][same], ro[some][same], rtol=1e-5, atol=2e-6)
    assert (rp[some] > 0).any(axis=1).mean() > 0.3


def test_single_part_compounds_equal_plain_contact(ctx, oracle):
    """Compounds whose only part sits at the identity pose. The nested composite dispatch solves the leaf problem with swapped
    roles — contact_shape_composite_shape = contact_composite_shape_shape(pos12.inverse(), g2, g1).flipped()
    (contact_composite_shape_shape.rs:63-76) — so against plain contact(part_i, part_j) only the outcome and the distance agree
    (witness points of face / edge contacts are not unique and GJK/EPA pick others when the roles swap: round 1's failure). The
    bit-level checker for this layout is the oracle's contact_compound_compound on the same single-part compounds."""
    import torch
    import parry_b200
    spec, compounds, a, p1, b, p2 = make_scene(6000, seed=103)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    ns = len(spec)
    T, G, C = tables(ctx, oracle, spec, [[(ident, s)] for s in range(ns)])
    a, b = (a % ns).astype(np.uint32), (b % ns).astype(np.uint32)
    go, gs, gp = C.contact_compounds(a, p1, b, p2, 0.05)
    ro, rs, rp = T.contact_compound_compound(C.first, C.count, C.part_shape, C.part_pose, a, p1, b, p2, 0.05, threads=8)
    ok = gs != 3
    assert (~ok).sum() <= 2
    assert (gs[ok] == rs[ok]).all() and (gs == 1).mean() > 0.1
    some = ok & (rs == 1)
    assert (gp[some] == rp[some]).all()
    np.testing.assert_allclose(go[some], ro[some], rtol=1e-5, atol=2e-6)
    # against the plain dispatcher: same membership away from the prediction threshold, same distance to GJK's tolerance
    po, pst = parry_b200.contact(G, a, p1, b, p2, 0.05)
    both = ok & (gs == 1) & (pst == 1)
    assert ((gs == 1) != (pst == 1))[ok].mean() < 2e-3
    dist_c, dist_p = go[both][:, 12], po[both][:, 12]
    assert np.abs(dist_c - dist_p).max() < 5e-3
    # invalid compound ids are status 2; device-resident arrays give the same bits
    a2 = a.copy()
    a2[::9] = 1000
    g2 = C.contact_compounds(a2, p1, b, p2, 0.05)
    assert (g2[1][::9] == 2).all() and (g2[2][::9] == 0xFFFFFFFF).all() and (g2[1][1::9] == gs[1::9]).all()
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    d = C.contact_compounds(dev(a), dev(p1), dev(b), dev(p2), 0.05)
    ctx.synchronize()
    assert (d[0].cpu().numpy().view(np.uint32) == go.view(np.uint32)).all() and (d[1].cpu().numpy() == gs).all()
    assert (d[2].cpu().numpy().view(np.uint32) == gp).all()


def test_gpu_reproduces_compound_compound_golden(ctx):
    """tests/golden/second_frame_2000.npz (frozen oracle output) on the sibling golden's compounds."""
    import os
    import parry_b200
    gd = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    y, w = np.load(os.path.join(gd, "siblings_3000.npz")), np.load(os.path.join(gd, "second_frame_2000.npz"))
    pu = y["params"].view(np.uint32)
    spec = [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p[:3]) if k == 1 else parry_b200.ConvexPolyhedron(y["points"][u[0]:u[0] + u[1]])
            for k, p, u in zip(y["kinds"], y["params"], pu)]
    G = parry_b200.Shapes(ctx, spec)
    compounds = [[(y["part_pose"][f + i], int(y["part_shape"][f + i])) for i in range(c)] for f, c in zip(y["comp_first"], y["comp_count"])]
    Cc = parry_b200.Compounds(ctx, G, compounds)
    go, gs, gp = Cc.contact_compounds(y["compound_id"], y["pos1"], w["compound_id2"], y["pos2_compound"], 0.05)
    ok = gs != 3
    assert (~ok).sum() <= 2
    assert (gs[ok] == w["cc_status"][ok]).all()
    some = ok & (gs == 1)
    same = (gp[some] == w["cc_parts"][some]).all(axis=1)
    assert same.mean() > 0.999
    np.testing.assert_allclose(go[some][same], w["cc_contacts"][some][same], rtol=1e-5, atol=2e-6)

"""GPU parity of query::contact with a Compound on one side (composite arms of DefaultQueryDispatcher::contact)
through pb2_compound_contact_shapes against the CPU oracle: statuses and winning parts exact, contacts within 1e-5."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu


def make_scene(n, seed):
    """64 compounds of 1-6 parts (balls, cuboids, 16-point hulls at random part poses) against single shapes."""
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(8, 16, seed=seed + 1)
    spec = [("ball", 0.3), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])] + [("convex", np.asarray(p, np.float32) * 0.5) for p in pts]
    ns = len(spec)
    compounds = []
    for c in range(64):
        k = int(g.integers(1, 7))
        poses = np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.6], axis=1).astype(np.float32)
        if c == 0:
            poses[0] = [0, 0, 0, 1, 0, 0, 0]
        compounds.append([(poses[i], int(g.integers(0, ns))) for i in range(k)])
    cid = g.integers(0, 64, n).astype(np.uint32)
    sid = g.integers(0, ns, n).astype(np.uint32)
    pc = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - 0.5) * 4], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    ps = np.concatenate([scenes.random_unit_quaternions(g, n), pc[:, 4:] + d * (g.random((n, 1)) * 2.2 + 0.1)], axis=1).astype(np.float32)
    return spec, compounds, cid, pc, sid, ps


def tables(ctx, oracle, spec, compounds):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(v) if k == "ball" else parry_b200.Cuboid(v) if k == "cuboid" else parry_b200.ConvexPolyhedron(v)
                                for k, v in spec])
    C = parry_b200.Compounds(ctx, G, compounds)
    return T, G, C


@pytest.mark.parametrize("second", [False, True])
def test_compound_contacts_vs_oracle(ctx, oracle, second):
    spec, compounds, cid, pc, sid, ps = make_scene(20000, seed=91)
    T, G, C = tables(ctx, oracle, spec, compounds)
    ro, rs, rp = T.contact_compound(C.first, C.count, C.part_shape, C.part_pose, cid, pc, sid, ps, 0.05, compound_second=second, threads=8)
    go, gs, gp = C.contact_shapes(cid, pc, sid, ps, 0.05, compound_second=second)
    assert 0.2 < (rs == 1).mean() < 0.9
    assert (gs == rs).all(), np.nonzero(gs != rs)[0][:10]
    some = rs == 1
    assert (gp[~some] == 0xFFFFFFFF).all() and (go[~some] == 0).all()
    same = gp[some] == rp[some]
    assert same.mean() > 0.999     # equal dists between two parts are the only legitimate difference (both pick the smallest index)
    np.testing.assert_allclose(go[some][same], ro[some][same], rtol=1e-5, atol=2e-6)
    assert (C.count[cid[some]] > 1).mean() > 0.5 and (rp[some] > 0).mean() > 0.3


def test_single_part_compound_equals_plain_contact(ctx, oracle):
    """A compound whose only part sits at the identity pose gives the contact of the part itself, bit for bit."""
    import parry_b200
    spec, compounds, cid, pc, sid, ps = make_scene(4000, seed=93)
    ident = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    ns = len(spec)
    T, G, C = tables(ctx, oracle, spec, [[(ident, s)] for s in range(ns)])
    cid = (cid % ns).astype(np.uint32)
    go, gs, gp = C.contact_shapes(cid, pc, sid, ps, 0.05)
    po, pst = parry_b200.contact(G, cid, pc, sid, ps, 0.05)
    assert (gs == pst).all() and (gs == 1).mean() > 0.1
    assert (go.view(np.uint32) == po.view(np.uint32)).all()


def test_invalid_ids_and_empty_compounds(ctx, oracle):
    import parry_b200
    spec, compounds, cid, pc, sid, ps = make_scene(1000, seed=95)
    T, G, C = tables(ctx, oracle, spec, compounds)
    bad_c, bad_s = cid.copy(), sid.copy()
    bad_c[3] = 64
    bad_s[5] = 10 ** 6
    _, st, part = C.contact_shapes(bad_c, pc, bad_s, ps, 0.05)
    assert st[3] == 2 and st[5] == 2 and part[3] == 0xFFFFFFFF
    with pytest.raises(parry_b200.Pb2Error):
        parry_b200.Compounds(ctx, G, [[]])
    with pytest.raises(parry_b200.Pb2Error):
        parry_b200.Compounds(ctx, G, [[(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), 10 ** 6)]])

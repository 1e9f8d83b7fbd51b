"""GPU parity of query::contact(Compound, TriMesh) (the composite arm with the Compound first, nested through
contact_shape_composite_shape over the mesh; default_query_dispatcher.rs:338-351, contact_composite_shape_shape.rs:12-76) through
pb2_compound_contact_trimesh against the CPU oracle's contact_compound_trimesh: statuses exact, winning {part, triangle} exact but
for exact dist ties, contacts within 1e-5. The oracle side is pinned by tests/test_oracle_kats.py::
test_compound_trimesh_contact_against_parts."""
import numpy as np
import pytest

from harness import scenes

pytestmark = pytest.mark.gpu


def make_scene(n, seed, n_compounds=32):
    g = scenes.rng(seed)
    v, idx = scenes.terrain(65, 65, extent=40.0)
    v = v.copy()
    v[:, 1] *= 0.2
    pts, _ = scenes.hull_pool(8, 24, seed=seed + 1)
    spec = [("ball", 0.35), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])]
    spec += [("convex", np.asarray(p, np.float32) * 0.5) for p in pts]
    ns = len(spec)
    compounds = []
    for c in range(n_compounds):
        k = int(g.integers(1, 6))
        poses = np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.6], axis=1).astype(np.float32)
        compounds.append([(poses[i], int(g.integers(0, ns))) for i in range(k)])
    ids = g.integers(0, n_compounds, n).astype(np.uint32)
    anchor = v[g.integers(0, len(v), n)]
    t = anchor + np.stack([g.standard_normal(n) * 0.3, (g.random(n) - 0.3) * 2.4, g.standard_normal(n) * 0.3], axis=1)
    poses = np.concatenate([scenes.random_unit_quaternions(g, n), t], axis=1).astype(np.float32)
    return v, idx, spec, compounds, ids, poses


def tables(ctx, oracle, spec, compounds):
    import parry_b200
    T = oracle.ShapeTable(spec)
    G = parry_b200.Shapes(ctx, [parry_b200.Ball(p) if k == "ball" else parry_b200.Cuboid(p) if k == "cuboid" else parry_b200.ConvexPolyhedron(p)
                                for k, p in spec])
    C = parry_b200.Compounds(ctx, G, compounds)
    return T, G, C


@pytest.mark.parametrize("seed,n,prediction,mesh_moved", [(301, 12000, 0.05, False), (302, 4000, 0.3, True), (303, 3000, 0.0, True)])
def test_compound_trimesh_vs_oracle(ctx, oracle, seed, n, prediction, mesh_moved):
    import parry_b200
    v, idx, spec, compounds, ids, poses = make_scene(n, seed)
    T, G, C = tables(ctx, oracle, spec, compounds)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    mpose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    if mesh_moved:
        # a small mesh motion keeps most compounds near the surface
        mq = np.array([0.01, -0.02, 0.015, 1.0]); mq /= np.linalg.norm(mq)
        mpose = np.concatenate([mq, [0.05, -0.1, 0.08]]).astype(np.float32)
    ro, rs, rp = om.contact_compounds(mpose, T, C.first, C.count, C.part_shape, C.part_pose, ids, poses, prediction, threads=8, min_index_ties=True)
    go, gs, gp = C.contact_trimesh(ids, poses, gm, mpose, prediction)
    go, gs, gp = np.asarray(go), np.asarray(gs), np.asarray(gp).astype(np.uint32)
    assert 0.2 < (rs == 1).mean() < 0.98
    assert (gs != 3).all()
    assert (gs == rs).all(), np.nonzero(gs != rs)[0][:10]
    some = rs == 1
    assert (gp[~some] == 0xFFFFFFFF).all() and (go[~some] == 0).all()
    np.testing.assert_allclose(go[some][:, 12], ro[some][:, 12], rtol=1e-5, atol=2e-6)
    same = (gp[some] == rp[some]).all(axis=1)
    assert same.mean() > 0.995, same.mean()
    np.testing.assert_allclose(go[some][same], ro[some][same], rtol=1e-5, atol=3e-6)
    exact = (go[some].view(np.uint32) == ro[some].view(np.uint32)).all(axis=1).mean()
    assert exact > 0.98, exact
    assert (rp[some][:, 0] > 0).mean() > 0.2            # later parts win too


@pytest.mark.parametrize("seed,n,prediction,mesh_moved", [(311, 8000, 0.05, False), (312, 3000, 0.3, True)])
def test_trimesh_compound_vs_oracle(ctx, oracle, seed, n, prediction, mesh_moved):
    """The other argument order, query::contact(mesh_pose, &TriMesh, pose, Compound): the reference walks the triangles first and
    solves every leaf problem with the part as shape 1 and the triangle as shape 2 (contact_composite_shape_shape.rs:12-76 twice);
    against the oracle's contact_trimesh_compound."""
    import parry_b200
    v, idx, spec, compounds, ids, poses = make_scene(n, seed)
    T, G, C = tables(ctx, oracle, spec, compounds)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    mpose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    if mesh_moved:
        mq = np.array([0.01, -0.02, 0.015, 1.0]); mq /= np.linalg.norm(mq)
        mpose = np.concatenate([mq, [0.05, -0.1, 0.08]]).astype(np.float32)
    ro, rs, rp = om.contact_compounds(mpose, T, C.first, C.count, C.part_shape, C.part_pose, ids, poses, prediction, trimesh_first=True, threads=8,
                                      min_index_ties=True)
    go, gs, gp = C.contact_trimesh(ids, poses, gm, mpose, prediction, mesh_first=True)
    go, gs, gp = np.asarray(go), np.asarray(gs), np.asarray(gp).astype(np.uint32)
    assert 0.2 < (rs == 1).mean() < 0.98
    assert (gs != 3).all()
    assert (gs == rs).all(), np.nonzero(gs != rs)[0][:10]
    some = rs == 1
    assert (gp[~some] == 0xFFFFFFFF).all() and (go[~some] == 0).all()
    np.testing.assert_allclose(go[some][:, 12], ro[some][:, 12], rtol=1e-5, atol=2e-6)
    same = (gp[some] == rp[some]).all(axis=1)
    assert same.mean() > 0.99, same.mean()
    np.testing.assert_allclose(go[some][same], ro[some][same], rtol=1e-5, atol=3e-6)
    exact = (go[some].view(np.uint32) == ro[some].view(np.uint32)).all(axis=1).mean()
    assert exact > 0.98, exact
    assert (rp[some][:, 0] > 0).mean() > 0.2
    # and it is the other order's flipped contact up to the leaf problems' swapped roles (same dist to rounding)
    fo, fs, fp = C.contact_trimesh(ids, poses, gm, mpose, prediction)
    fo, fs = np.asarray(fo), np.asarray(fs)
    both = (fs == 1) & (gs == 1)
    assert (fs == gs).mean() > 0.99
    # (the two orders prefilter differently — part boxes against the mesh's root box there, triangle boxes against the compound's
    # here — and EPA witnesses depend on the roles: a handful of deep contacts may settle on another local minimum)
    assert (np.abs(fo[both][:, 12] - go[both][:, 12]) < 5e-5).mean() > 0.995


def test_compound_trimesh_edge_cases_and_device_memory(ctx, oracle):
    import torch
    import parry_b200
    v, idx, spec, compounds, ids, poses = make_scene(2000, 304)
    T, G, C = tables(ctx, oracle, spec, compounds)
    gm = parry_b200.TriMesh(ctx, v, idx)
    mpose = np.array([0, 0, 0, 1, 0.1, 0.0, -0.2], np.float32)
    h = C.contact_trimesh(ids, poses, gm, mpose, 0.05)
    dev = lambda x: torch.from_numpy(x.view(np.int32) if x.dtype == np.uint32 else x).cuda()
    d = C.contact_trimesh(dev(ids), dev(poses), gm, dev(mpose), 0.05)
    ctx.synchronize()
    assert (d[1].cpu().numpy() == h[1]).all()
    assert (d[2].cpu().numpy().view(np.uint32) == np.asarray(h[2]).view(np.uint32)).all()
    assert (d[0].cpu().numpy().view(np.uint32) == np.asarray(h[0]).view(np.uint32)).all()
    # far away: no part meets the mesh's root box; unknown compound id: Unsupported; n = 0 is a no-op
    far = poses[:4].copy(); far[:, 4:] += 1000.0
    out, st, parts = C.contact_trimesh(ids[:4], far, gm, mpose, 0.05)
    assert (np.asarray(st) == 0).all() and (np.asarray(parts).astype(np.uint32) == 0xFFFFFFFF).all()
    bad = ids[:4].copy(); bad[2] = 9999
    out, st, parts = C.contact_trimesh(bad, poses[:4], gm, mpose, 0.05)
    assert np.asarray(st)[2] == 2 and (np.asarray(st)[[0, 1, 3]] == np.asarray(h[1])[[0, 1, 3]]).all()
    out, st, parts = C.contact_trimesh(ids[:0], poses[:0], gm, mpose, 0.05)
    assert len(np.asarray(st)) == 0


def test_single_part_compound_equals_flipped_mesh_contact(ctx, oracle):
    """A compound whose only part sits at the identity pose: contact(compound, mesh) is contact(mesh, shape).flipped()."""
    import parry_b200
    v, idx, spec, _, _, poses = make_scene(3000, 305)
    compounds = [[(np.array([0, 0, 0, 1, 0, 0, 0], np.float32), s)] for s in range(len(spec))]
    T, G, C = tables(ctx, oracle, spec, compounds)
    gm = parry_b200.TriMesh(ctx, v, idx)
    g = scenes.rng(306)
    ids = g.integers(0, len(spec), len(poses)).astype(np.uint32)
    mpose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    co, cs, cp = C.contact_trimesh(ids, poses, gm, mpose, 0.05)
    mo, ms, mp = gm.contact_shapes(mpose, G, ids, poses, 0.05)
    co, cs, cp, mo, ms, mp = (np.asarray(x) for x in (co, cs, cp, mo, ms, mp))
    assert (cs == ms).all() and (cs == 1).mean() > 0.2
    some = cs == 1
    assert (cp[some][:, 0] == 0).all()
    # the nested dispatch composes the poses (inverse of an inv_mul), which can move the last bit of the shape's pose in the mesh
    # frame: equal-dist triangles (shared edges) may swap, everything else agrees to rounding
    same = cp[some][:, 1].astype(np.uint32) == mp[some].astype(np.uint32)
    # (measured: 95 % — a ball or a hull vertex above a shared terrain edge sees the same closest point from both triangles)
    assert same.mean() > 0.9, same.mean()
    # (coordinates reach 20: one ulp of a recomposed translation is 2e-6)
    np.testing.assert_allclose(co[some][:, 12], mo[some][:, 12], rtol=0, atol=3e-5)
    a, b = co[some][same], mo[some][same]
    flipped = np.concatenate([b[:, 3:6], b[:, 0:3], b[:, 9:12], b[:, 6:9]], axis=1)
    err = np.abs(a[:, :12] - flipped)
    rows_ok = (err[:, :6] < 1e-4).all(axis=1) & (err[:, 6:] < 1e-4).all(axis=1)
    # witnesses are not unique for face-face contacts: a last-bit pose change may pick another point of the same face pair
    assert rows_ok.mean() > 0.97, rows_ok.mean()

"""TriMesh-vs-shape contacts: the composite-shape arm of DefaultQueryDispatcher::contact
(contact_composite_shape_shape.rs:14-61) — mesh Bvh query with the loosened shape AABB, every reported triangle dispatched
as a shape::Triangle, smallest dist kept. CPU tests pin the oracle on closed-form cases; GPU tests compare
pb2_trimesh_contact_shapes with the oracle on the same inputs."""
import numpy as np
import pytest

from harness import scenes

I4 = [0.0, 0.0, 0.0, 1.0]


def flat_mesh(size=4.0):
    v = np.array([[-size, 0, -size], [size, 0, -size], [size, 0, size], [-size, 0, size]], np.float32)
    idx = np.array([[0, 2, 1], [0, 3, 2]], np.uint32)   # normals +y
    return v, idx


def test_oracle_ball_and_cuboid_on_a_flat_mesh():
    from harness import oracle
    oracle.build()
    v, idx = flat_mesh()
    om = oracle.TriMesh(v, idx)
    T = oracle.ShapeTable([("ball", 0.5), ("cuboid", [0.5, 0.5, 0.5])])
    ident = np.array(I4 + [0, 0, 0], np.float32)
    pos = np.array([I4 + [0.3, 0.4, -0.2], I4 + [1.0, 0.8, 1.0], I4 + [0.5, 0.45, 0.7], I4 + [0.0, 2.0, 0.0]], np.float32)
    out, st, part = om.contact_shapes(ident, T, [0, 0, 1, 1], pos, 0.05)
    assert st.tolist() == [1, 0, 1, 0]
    # ball: centre 0.4 above the plane, r = 0.5 => dist -0.1, normal1 = +y, point1 = projection of the centre
    assert abs(out[0, 12] + 0.1) < 1e-6 and np.allclose(out[0, 6:9], [0, 1, 0], atol=1e-6)
    assert np.allclose(out[0, 0:3], [0.3, 0.0, -0.2], atol=1e-6) and part[0] in (0, 1)
    # cuboid: bottom face 0.05 below the plane
    assert abs(out[2, 12] + 0.05) < 1e-5 and np.allclose(out[2, 6:9], [0, 1, 0], atol=1e-5)
    assert part[1] == 0xFFFFFFFF and part[3] == 0xFFFFFFFF
    # a mesh pose moves everything: lift the mesh by 1.0 => the far ball now touches, the near ones are deep inside
    lifted = np.array(I4 + [0, 1.55, 0], np.float32)
    out2, st2, _ = om.contact_shapes(lifted, T, [1], pos[3:4], 0.1)
    assert st2[0] == 1 and abs(out2[0, 12] - (-0.05)) < 1e-5


def make_case(seed, n, spread):
    v, idx = scenes.terrain(65, 65, extent=40.0)
    v = v.copy()
    v[:, 1] *= 0.2                                        # gentle hills: heights within +-8
    g = scenes.rng(seed)
    pts, _ = scenes.hull_pool(16, 32, seed=seed + 1)
    spec = [("ball", float(r)) for r in g.random(8) * 0.5 + 0.2]
    spec += [("cuboid", list(h)) for h in g.random((8, 3)) * 0.5 + 0.15]
    spec += [("convex", np.asarray(p, np.float32) * 0.6) for p in pts]
    sid = g.integers(0, len(spec), n).astype(np.uint32)
    anchor = v[g.integers(0, len(v), n)]
    t = anchor + np.stack([g.standard_normal(n) * 0.3, (g.random(n) - 0.35) * spread, g.standard_normal(n) * 0.3], axis=1)
    q = scenes.random_unit_quaternions(g, n)
    poses = np.concatenate([q, t], axis=1).astype(np.float32)
    mq = np.array([0.05, -0.1, 0.08, 0.99]); mq /= np.linalg.norm(mq)
    mesh_pose = np.concatenate([mq, [0.7, -0.3, 1.1]]).astype(np.float32)
    # shapes were placed in mesh-local space: move them with the mesh
    from harness import oracle
    return v, idx, spec, sid, poses, mesh_pose


def to_world(mesh_pose, poses):
    """pose_world = mesh_pose * pose_local (quaternion product + rotated translation), in float64 then rounded."""
    def qmul(a, b):
        ai, aj, ak, aw = a[..., 0], a[..., 1], a[..., 2], a[..., 3]
        bi, bj, bk, bw = b[..., 0], b[..., 1], b[..., 2], b[..., 3]
        return np.stack([aw * bi + ai * bw + aj * bk - ak * bj, aw * bj - ai * bk + aj * bw + ak * bi,
                         aw * bk + ai * bj - aj * bi + ak * bw, aw * bw - ai * bi - aj * bj - ak * bk], axis=-1)
    mq = mesh_pose[:4].astype(np.float64)
    q = qmul(np.broadcast_to(mq, poses[:, :4].shape), poses[:, :4].astype(np.float64))
    u = mq[:3]
    t = poses[:, 4:].astype(np.float64)
    tt = 2.0 * np.cross(u, t)
    rt = t + mq[3] * tt + np.cross(u, tt)
    return np.concatenate([q, rt + mesh_pose[4:]], axis=1).astype(np.float32)


@pytest.mark.gpu
@pytest.mark.parametrize("seed,n,prediction", [(81, 20000, 0.05), (82, 6000, 0.0), (83, 3000, 0.6)])
def test_trimesh_contact_shapes_matches_oracle(ctx, oracle, seed, n, prediction):
    import parry_b200
    v, idx, spec, sid, poses_local, mesh_pose = make_case(seed, n, 1.6)
    poses = to_world(mesh_pose, poses_local)
    gshapes = [parry_b200.Ball(p) if k == "ball" else parry_b200.Cuboid(p) if k == "cuboid" else parry_b200.ConvexPolyhedron(p) for k, p in spec]
    G, O = parry_b200.Shapes(ctx, gshapes), oracle.ShapeTable(spec)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    gout, gst, gpart = gm.contact_shapes(mesh_pose, G, sid, poses, prediction)
    oout, ost, opart = om.contact_shapes(mesh_pose, O, sid, poses, prediction, threads=8, min_index_ties=True)
    gst, gpart, gout = np.asarray(gst), np.asarray(gpart).astype(np.uint32), np.asarray(gout)
    assert 0.2 < (ost == 1).mean() < 0.999
    for kind_lo, kind_hi in ((0, 8), (8, 16), (16, 32)):      # every arm sees contacts
        m = (sid >= kind_lo) & (sid < kind_hi)
        assert (ost[m] == 1).mean() > 0.1
    assert (gst == ost).all(), np.nonzero(gst != ost)[0][:10]
    some = ost == 1
    np.testing.assert_allclose(gout[some], oout[some], rtol=1e-5, atol=1e-6)
    # the winning triangle: exact, except where two triangles give bit-equal dist up to the last ulps (shared edges)
    diff = some & (gpart != opart)
    assert diff.mean() < 0.002
    assert (np.abs(gout[diff, 12] - oout[diff, 12]) <= 1e-6).all()
    exact = (gout[some].view(np.uint32) == oout[some].view(np.uint32)).all(axis=1).mean()
    assert exact > 0.995, exact
    # reference order (first strictly smaller in its own BVH order) differs from min-index only on exact dist ties
    rout, rst, rpart = om.contact_shapes(mesh_pose, O, sid, poses, prediction, threads=8, min_index_ties=False)
    assert (rst == ost).all()
    assert (rout[some, 12] == oout[some, 12]).all()


@pytest.mark.gpu
def test_trimesh_contact_shapes_device_resident_and_edge_cases(ctx, oracle):
    import torch
    import parry_b200
    v, idx, spec, sid, poses_local, mesh_pose = make_case(84, 5000, 1.6)
    poses = to_world(mesh_pose, poses_local)
    gshapes = [parry_b200.Ball(p) if k == "ball" else parry_b200.Cuboid(p) if k == "cuboid" else parry_b200.ConvexPolyhedron(p) for k, p in spec]
    G = parry_b200.Shapes(ctx, gshapes)
    gm = parry_b200.TriMesh(ctx, v, idx)
    h = gm.contact_shapes(mesh_pose, G, sid, poses, 0.05)
    d = gm.contact_shapes(torch.from_numpy(mesh_pose).cuda(), G, torch.from_numpy(sid.view(np.int32)).cuda(), torch.from_numpy(poses).cuda(), 0.05)
    ctx.synchronize()
    assert (d[1].cpu().numpy() == h[1]).all() and (d[2].cpu().numpy().view(np.uint32) == h[2]).all()
    assert (d[0].cpu().numpy().view(np.uint32) == h[0].view(np.uint32)).all()
    # far away shapes: no candidates at all; unknown shape id: Unsupported
    far = poses[:4].copy(); far[:, 4:] += 1000.0
    out, st, part = gm.contact_shapes(mesh_pose, G, sid[:4], far, 0.05)
    assert (np.asarray(st) == 0).all() and (np.asarray(part).astype(np.uint32) == 0xFFFFFFFF).all()
    bad = sid[:4].copy(); bad[1] = 9999
    out, st, part = gm.contact_shapes(mesh_pose, G, bad, poses[:4], 0.05)
    assert np.asarray(st)[1] == 2

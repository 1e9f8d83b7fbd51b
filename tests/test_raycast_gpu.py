"""GPU parity: TriMesh::cast_ray through the C ABI vs the CPU oracle (reference-order traversal) on the same inputs."""
import numpy as np
import pytest

from harness import scenes
from helpers import INVALID, assert_well_formed, check_ray_parity

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)


@pytest.fixture(scope="module")
def sphere(ctx, oracle):
    import parry_b200
    v, i = scenes.uv_sphere(64, 48)
    return v, i, parry_b200.TriMesh(ctx, v, i), oracle.TriMesh(v, i)


def _brute(omesh, pose, rays, max_toi):
    def fn(idx):
        t, i, _, _ = omesh.cast_rays(pose, rays[idx], max_toi, with_normal=True, mode=1, threads=4)
        return t, i
    return fn


def test_trimesh_bvh_well_formed(sphere):
    v, i, gmesh, omesh = sphere
    nodes, parents, leaf_idx = gmesh.bvh().download()
    assert len(nodes) == len(i) - 1
    assert_well_formed(nodes, parents, leaf_idx)


def test_cast_local_ray_toi_and_ids(sphere):
    v, i, gmesh, omesh = sphere
    rays = scenes.sphere_rays(50000, seed=11)
    g = gmesh.cast_local_ray(rays, FMAX)
    r = omesh.cast_rays(None, rays, FMAX, threads=8)
    assert (np.asarray(r[1]) != INVALID).mean() > 0.5
    check_ray_parity(g, r, _brute(omesh, None, rays, FMAX))


def test_cast_ray_and_get_normal_with_pose(sphere):
    v, i, gmesh, omesh = sphere
    rays = scenes.sphere_rays(30000, seed=12)
    q = np.array([0.3, -0.2, 0.5, 0.7], dtype=np.float64)
    q /= np.linalg.norm(q)
    pose = np.concatenate([q, [0.1, -0.25, 0.3]]).astype(np.float32)
    g = gmesh.cast_ray_and_get_normal(pose, rays, FMAX)
    r = omesh.cast_rays(pose, rays, FMAX, with_normal=True, threads=8)
    check_ray_parity(g, r, _brute(omesh, pose, rays, FMAX))
    hit = np.asarray(g[1]) != INVALID
    # features: front faces => Face(tri), back faces => Face(tri + nt)
    f = np.asarray(g[3]).astype(np.uint32)[hit]
    t = np.asarray(g[1]).astype(np.uint32)[hit]
    assert ((f == t) | (f == t + len(i))).all()


def test_max_toi_limits_hits(sphere):
    v, i, gmesh, omesh = sphere
    rays = scenes.sphere_rays(20000, seed=13)
    for max_toi in (0.25, 0.5, 0.8):
        g = gmesh.cast_local_ray(rays, max_toi)
        r = omesh.cast_rays(None, rays, max_toi, threads=8)
        check_ray_parity(g, r, _brute(omesh, None, rays, max_toi))
        hit = np.asarray(g[1]) != INVALID
        assert (np.asarray(g[0])[hit] < max_toi).all()


def test_rays_from_inside_hit_backfaces(sphere):
    v, i, gmesh, omesh = sphere
    g0 = scenes.rng(14)
    d = g0.standard_normal((5000, 3))
    o = (g0.random((5000, 3)) - 0.5) * 0.5
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    g = gmesh.cast_local_ray_and_get_normal(rays, FMAX)
    r = omesh.cast_rays(None, rays, FMAX, with_normal=True, threads=8)
    check_ray_parity(g, r, _brute(omesh, None, rays, FMAX))
    assert (np.asarray(g[3]).astype(np.uint32) >= len(i)).all()  # all back faces


def test_device_resident_matches_host(sphere, ctx):
    import torch
    v, i, gmesh, omesh = sphere
    rays = scenes.sphere_rays(10000, seed=15)
    h = gmesh.cast_local_ray_and_get_normal(rays, FMAX)
    d = gmesh.cast_local_ray_and_get_normal(torch.from_numpy(rays).cuda(), FMAX)
    ctx.synchronize()
    assert (d[0].cpu().numpy().view(np.uint32) == h[0].view(np.uint32)).all()
    assert (d[1].cpu().numpy().view(np.uint32) == h[1]).all()
    assert (d[2].cpu().numpy().view(np.uint32) == h[2].view(np.uint32)).all()
    assert (d[3].cpu().numpy().view(np.uint32) == h[3]).all()


def test_tie_stress_axis_aligned_grid(ctx, oracle):
    """Rays through shared edges / vertices of an axis-aligned grid mesh: exact toi ties."""
    import parry_b200
    n = 17
    xs = np.arange(n, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    v = np.stack([X, np.zeros_like(X), Z], axis=-1).reshape(-1, 3).astype(np.float32)
    a = (np.arange(n - 1)[:, None] * n + np.arange(n - 1)[None, :]).ravel()
    idx = np.concatenate([np.stack([a, a + 1, a + n], 1), np.stack([a + 1, a + n + 1, a + n], 1)]).astype(np.uint32)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    # straight-down rays at every half-integer lattice point: vertices, edge midpoints, diagonals, interiors
    px, pz = np.meshgrid(np.arange(0, n - 1 + 0.01, 0.5), np.arange(0, n - 1 + 0.01, 0.5), indexing="ij")
    o = np.stack([px.ravel(), np.full(px.size, 3.0), pz.ravel()], axis=1)
    d = np.tile([0.0, -1.0, 0.0], (len(o), 1))
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    r = om.cast_rays(None, rays, FMAX, with_normal=True)
    b = om.cast_rays(None, rays, FMAX, with_normal=True, mode=1)
    gt, gi = np.asarray(g[0]), np.asarray(g[1]).astype(np.uint32)
    assert (gi != INVALID).all() and (gt == 3.0).all()
    assert (gt.view(np.uint32) == r[0].view(np.uint32)).all()
    # documented rule: smallest index among bit-equal minimal toi == brute force with min-index ties
    assert (gi == b[1]).all()
    assert (gi <= r[1]).all()


def test_cast_ray_with_culling(sphere, ctx, oracle):
    """TriMesh::cast_ray_with_culling: the reference's own test (ray_trimesh.rs:187-212) and oracle parity on the sphere
    for both culling modes, small (thread-per-ray kernel) and large (wide-tree kernel) batches, with a pose."""
    import parry_b200
    v = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0]], np.float32)
    tm = parry_b200.TriMesh(ctx, v, np.array([[0, 1, 2]], np.uint32))
    up = np.array([[0, 0, -1, 0, 0, 1]], np.float32)
    down = np.array([[0, 0, 1, 0, 0, -1]], np.float32)
    hit = lambda rays, mode: tm.cast_local_ray_with_culling(rays, 1000.0, mode)[1][0] != INVALID
    assert hit(up, tm.IGNORE_FRONTFACES) and not hit(down, tm.IGNORE_FRONTFACES)
    assert not hit(up, tm.IGNORE_BACKFACES) and hit(down, tm.IGNORE_BACKFACES)
    vs, i, gmesh, omesh = sphere
    g0 = scenes.rng(16)
    q = np.array([0.1, 0.7, -0.2, 0.6]); q /= np.linalg.norm(q)
    pose = np.concatenate([q, [0.05, 0.1, -0.2]]).astype(np.float32)
    for m in (3000, 40000):
        o = (g0.random((m, 3)) - 0.5) * 2.4          # origins inside and outside the sphere
        rays = np.concatenate([o, g0.standard_normal((m, 3))], axis=1).astype(np.float32)
        for mode in (1, 2):
            g = gmesh.cast_ray_with_culling(pose, rays, FMAX, mode)
            r = omesh.cast_rays(pose, rays, FMAX, with_normal=True, mode=1 + mode, threads=8)
            hitm = np.asarray(r[1]) != INVALID
            assert 0.05 < hitm.mean() < 0.95
            same = (np.asarray(g[1]).astype(np.uint32) == r[1]) & (np.asarray(g[0]).view(np.uint32) == r[0].view(np.uint32))
            assert same.mean() > 0.9995, same.mean()
            back = np.asarray(g[3]).astype(np.uint32)[hitm & same] >= len(i)
            assert back.all() if mode == 2 else not back.any()
            np.testing.assert_allclose(np.asarray(g[2])[same], r[2][same], rtol=1e-5, atol=1e-7)


@pytest.mark.parametrize("nt", [1, 2, 3, 5])
def test_tiny_meshes(ctx, oracle, nt):
    import parry_b200
    g0 = scenes.rng(100 + nt)
    v = g0.random((3 * nt, 3)).astype(np.float32)
    idx = np.arange(3 * nt, dtype=np.uint32).reshape(nt, 3)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    o = g0.random((4000, 3)) * 3 - 1
    t = g0.random((4000, 3))
    rays = np.concatenate([o, t - o], axis=1).astype(np.float32)
    g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    r = om.cast_rays(None, rays, FMAX, with_normal=True)
    check_ray_parity(g, r, _brute(om, None, rays, FMAX))


def test_empty_mesh_is_an_error(ctx):
    import parry_b200
    with pytest.raises(parry_b200.Pb2Error):
        parry_b200.TriMesh(ctx, np.zeros((3, 3), np.float32), np.zeros((0, 3), np.uint32))


def test_full_size_config1_properties(ctx, oracle):
    """BASELINE config[0]: 1M rays vs the 99,904-triangle sphere. Oracle parity on a 100k-ray slice, geometric
    property on all rays: every hit point lies on the unit sphere to within the facet sagitta."""
    import parry_b200
    v, i = scenes.uv_sphere(224, 224)
    gm, om = parry_b200.TriMesh(ctx, v, i), oracle.TriMesh(v, i)
    rays = scenes.sphere_rays(1 << 20, seed=1)
    g = gm.cast_local_ray(rays, FMAX)
    sl = slice(0, 100000)
    r = om.cast_rays(None, rays[sl], FMAX, threads=8)
    check_ray_parity((g[0][sl], g[1][sl]), r, _brute(om, None, rays[sl], FMAX))
    hit = g[1] != INVALID
    assert 0.7 < hit.mean() < 0.95
    p = rays[hit, :3].astype(np.float64) + rays[hit, 3:].astype(np.float64) * g[0][hit, None].astype(np.float64)
    rad = np.linalg.norm(p, axis=1)
    assert rad.max() <= 1.0 + 1e-5 and rad.min() >= 1.0 - 3e-4


def test_intersects_ray(sphere):
    """RayCast::intersects_ray == cast_ray(..).is_some()"""
    v, i, gm, om = sphere
    rays = scenes.sphere_rays(3000, seed=31)
    pose = np.array([0, 0, 0, 1, 0, 0, 0], np.float32)
    hit = gm.intersects_ray(pose, rays, FMAX)
    toi, tri = om.cast_rays(pose, rays, FMAX)[:2]
    assert hit.dtype == bool and (hit == (tri != INVALID)).all() and 0.1 < hit.mean() < 1.0

"""Deterministic synthetic inputs for the configs in BASELINE.json / SURVEY.md §8(d). All generators are seeded
numpy PCG64 streams (float64 draws cast to f32) so the oracle and the GPU see identical bits."""
import numpy as np


def rng(seed):
    return np.random.Generator(np.random.PCG64(seed))


def uv_sphere(ntheta, nphi, radius=1.0):
    """UV sphere: 2 poles + (nphi-1) rings of ntheta vertices; T = 2*ntheta*(nphi-1) triangles, outward CCW."""
    th = (np.arange(ntheta, dtype=np.float64) / ntheta) * 2.0 * np.pi
    ph = (np.arange(1, nphi, dtype=np.float64) / nphi) * np.pi
    st, ct = np.sin(th), np.cos(th)
    sp, cp = np.sin(ph), np.cos(ph)
    ring = np.stack([np.outer(sp, ct), np.repeat(cp[:, None], ntheta, 1), np.outer(sp, st)], axis=-1)  # (nphi-1, ntheta, 3)
    verts = np.concatenate([[[0.0, 1.0, 0.0]], [[0.0, -1.0, 0.0]], ring.reshape(-1, 3)], axis=0) * radius
    def vid(r, t):
        return 2 + r * ntheta + (t % ntheta)
    t = np.arange(ntheta)
    tris = []
    # top cap
    tris.append(np.stack([np.zeros(ntheta, dtype=np.int64), vid(0, t + 1), vid(0, t)], axis=1))
    for r in range(nphi - 2):
        a, b, c, d = vid(r, t), vid(r, t + 1), vid(r + 1, t), vid(r + 1, t + 1)
        tris.append(np.stack([a, b, c], axis=1))
        tris.append(np.stack([b, d, c], axis=1))
    tris.append(np.stack([np.ones(ntheta, dtype=np.int64), vid(nphi - 2, t), vid(nphi - 2, t + 1)], axis=1))
    idx = np.concatenate(tris, axis=0)
    return verts.astype(np.float32), idx.astype(np.uint32)


def sphere_rays(m, seed=1):
    """origin uniform in [-3,3]^3 rejected if |o| < 1.5, dir = target - origin (un-normalised),
    target uniform in the ball of radius 1.2."""
    g = rng(seed)
    out_o = np.empty((0, 3))
    while out_o.shape[0] < m:
        o = g.random((2 * m, 3)) * 6.0 - 3.0
        o = o[np.linalg.norm(o, axis=1) >= 1.5]
        out_o = np.concatenate([out_o, o], axis=0)
    o = out_o[:m]
    out_t = np.empty((0, 3))
    while out_t.shape[0] < m:
        t = g.random((3 * m, 3)) * 2.4 - 1.2
        t = t[np.linalg.norm(t, axis=1) <= 1.2]
        out_t = np.concatenate([out_t, t], axis=0)
    t = out_t[:m]
    rays = np.concatenate([o, t - o], axis=1).astype(np.float32)
    return np.ascontiguousarray(rays)


def terrain(nx, nz, extent=1000.0, seed=5):
    """Heightfield grid (nx x nz vertices) => 2*(nx-1)*(nz-1) triangles; height = 4 seeded sine octaves."""
    g = rng(seed)
    xs = np.linspace(-extent / 2, extent / 2, nx)
    zs = np.linspace(-extent / 2, extent / 2, nz)
    X, Z = np.meshgrid(xs, zs, indexing="ij")
    H = np.zeros_like(X)
    amp, freq = 20.0, 2.0 * np.pi / extent * 3.0
    for _ in range(4):
        px, pz = g.random(2) * 2.0 * np.pi
        ax, az = g.random(2) * 0.5 + 0.75
        H += amp * np.sin(X * freq * ax + px) * np.cos(Z * freq * az + pz)
        amp *= 0.5
        freq *= 2.1
    verts = np.stack([X, H, Z], axis=-1).reshape(-1, 3).astype(np.float32)
    i, j = np.meshgrid(np.arange(nx - 1), np.arange(nz - 1), indexing="ij")
    a = (i * nz + j).ravel()
    b = a + 1
    c = a + nz
    d = c + 1
    idx = np.concatenate([np.stack([a, b, c], 1), np.stack([b, d, c], 1)], axis=0).astype(np.uint32)
    return verts, np.ascontiguousarray(idx)


def terrain_rays(m, extent=1000.0, seed=6):
    """Incoherent rays: origins uniform above the terrain (y in [50,150]), directions uniform on the lower hemisphere."""
    g = rng(seed)
    o = np.stack([(g.random(m) - 0.5) * extent * 0.9, g.random(m) * 100.0 + 50.0, (g.random(m) - 0.5) * extent * 0.9], axis=1)
    z = -g.random(m)
    phi = g.random(m) * 2.0 * np.pi
    rxy = np.sqrt(np.maximum(0.0, 1.0 - z * z))
    d = np.stack([rxy * np.cos(phi), z, rxy * np.sin(phi)], axis=1)
    return np.ascontiguousarray(np.concatenate([o, d], axis=1).astype(np.float32))


def random_unit_quaternions(g, n):
    q = g.standard_normal((n, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    return q  # (i, j, k, w)


def compose_pose(p, q):
    """p * q for two (quaternion ijkw, translation) poses, in float64, rounded to float32."""
    pq, pt, qq, qt = (np.asarray(x, np.float64) for x in (p[:4], p[4:], q[:4], q[4:]))

    def rot(u4, x):
        u, w = u4[:3], u4[3]
        t = 2.0 * np.cross(u, x)
        return x + w * t + np.cross(u, t)
    w = pq[3] * qq[3] - np.dot(pq[:3], qq[:3])
    v = pq[3] * qq[:3] + qq[3] * pq[:3] + np.cross(pq[:3], qq[:3])
    return np.concatenate([v, [w], pt + rot(pq, qt)]).astype(np.float32)


def colliders(n, side=None, seed=2, hull_fraction=0.0, n_hulls=0):
    """n colliders with centres uniform in a cube (density ~1/unit^3), balls r in [0.2,0.6],
    cuboids h in [0.2,0.6]^3 with random rotations, optional hulls from a pool. Returns (kinds, params(n,3), poses(n,7),
    hull_ids)."""
    g = rng(seed)
    if side is None:
        side = float(n) ** (1.0 / 3.0)
    centres = g.random((n, 3)) * side
    u = g.random(n)
    if hull_fraction > 0:
        kinds = np.where(u < hull_fraction, 2, np.where(u < hull_fraction + (1 - hull_fraction) / 2, 0, 1)).astype(np.uint8)
    else:
        kinds = (u >= 0.5).astype(np.uint8)
    params = (g.random((n, 3)) * 0.4 + 0.2).astype(np.float32)
    q = random_unit_quaternions(g, n)
    q[kinds == 0] = [0.0, 0.0, 0.0, 1.0]
    poses = np.concatenate([q, centres], axis=1).astype(np.float32)
    hull_ids = (g.integers(0, max(1, n_hulls), n)).astype(np.uint32)
    return kinds, params, np.ascontiguousarray(poses), hull_ids


def hull_pool(n_hulls, n_points=32, seed=3):
    """each hull = n_points normalised random directions x radius in [0.5, 1] (all on the hull)."""
    g = rng(seed)
    d = g.standard_normal((n_hulls, n_points, 3))
    d /= np.linalg.norm(d, axis=2, keepdims=True)
    r = g.random(n_hulls) * 0.5 + 0.5
    pts = (d * r[:, None, None]).astype(np.float32)
    return pts, r.astype(np.float32)


def hull_pairs(n_pairs, radii, seed=4, s_lo=0.6, s_hi=2.4):
    """(hullA, hullB) uniform; poseA = identity rotation, random translation; poseB = random
    rotation, translation = tA + random dir * s, s uniform in [0.6, 2.4]*(rA+rB)/2 (67 % of the pairs penetrate and go
    through EPA; [1.7, 2.4] gives the shallow mix of a settled scene: 18 % penetrating, 20 % in contact)."""
    g = rng(seed)
    nh = len(radii)
    a = g.integers(0, nh, n_pairs).astype(np.uint32)
    b = g.integers(0, nh, n_pairs).astype(np.uint32)
    ta = (g.random((n_pairs, 3)) - 0.5) * 20.0
    qa = np.tile(np.array([0.0, 0.0, 0.0, 1.0]), (n_pairs, 1))
    qb = random_unit_quaternions(g, n_pairs)
    d = g.standard_normal((n_pairs, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    s = (g.random(n_pairs) * (s_hi - s_lo) + s_lo) * (radii[a] + radii[b]) / 2.0
    tb = ta + d * s[:, None]
    pos1 = np.ascontiguousarray(np.concatenate([qa, ta], axis=1).astype(np.float32))
    pos2 = np.ascontiguousarray(np.concatenate([qb, tb], axis=1).astype(np.float32))
    # re-normalise quaternions in f32 so both sides see the same (already rounded) unit quaternion bits
    return a, b, pos1, pos2

"""GPU parity of every ray-kernel variant (PB2_RAY_VARIANT: 0 thread-per-ray, 1 persistent binary tree, 2 = 1 + ray
reordering, 3 compressed 8-wide tree, 4 = 3 + ray reordering, 5 = 3 with the warp-shared triangle phase, 6 = 5 + ray reordering) against the CPU oracle and against each other."""
import os

import numpy as np
import pytest

from harness import scenes
from helpers import INVALID, check_ray_parity

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)
VARIANTS = [0, 1, 2, 3, 4, 5, 6]


class _Variant:
    def __init__(self, v):
        self.v = v

    def __enter__(self):
        self.old = os.environ.get("PB2_RAY_VARIANT")
        os.environ["PB2_RAY_VARIANT"] = str(self.v)

    def __exit__(self, *a):
        if self.old is None:
            os.environ.pop("PB2_RAY_VARIANT", None)
        else:
            os.environ["PB2_RAY_VARIANT"] = self.old


def _brute(omesh, rays, max_toi):
    def fn(idx):
        t, i, _, _ = omesh.cast_rays(None, rays[idx], max_toi, with_normal=True, mode=1, threads=4)
        return t, i
    return fn


@pytest.fixture(scope="module")
def terrain(ctx, oracle):
    import parry_b200
    v, i = scenes.terrain(129, 129)
    return v, i, parry_b200.TriMesh(ctx, v, i), oracle.TriMesh(v, i)


@pytest.mark.parametrize("variant", VARIANTS)
def test_terrain_variant_vs_oracle(terrain, variant):
    v, i, gm, om = terrain
    rays = scenes.terrain_rays(70000, seed=21)
    with _Variant(variant):
        g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    r = om.cast_rays(None, rays, FMAX, with_normal=True, threads=8)
    assert (np.asarray(r[1]) != INVALID).mean() > 0.3
    check_ray_parity(g, r, _brute(om, rays, FMAX))


def test_variants_agree_bit_for_bit(terrain):
    v, i, gm, om = terrain
    rays = scenes.terrain_rays(200000, seed=22)
    outs = []
    for variant in VARIANTS:
        with _Variant(variant):
            outs.append(gm.cast_local_ray(rays, FMAX))
    t0, i0 = outs[0]
    for t, i_ in outs[1:]:
        diff = (t.view(np.uint32) != t0.view(np.uint32)) | (i_ != i0)
        # only order-dependent ulp-level AABB culls may differ; none expected on this scene
        assert diff.sum() <= 2, int(diff.sum())


@pytest.mark.parametrize("variant", [1, 3, 5])
def test_tie_stress_large_batch(ctx, oracle, variant):
    """Exact toi ties (rays through shared edges / vertices) in batches large enough to take the persistent kernels."""
    import parry_b200
    n = 33
    xs = np.arange(n, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="ij")
    v = np.stack([X, np.zeros_like(X), Z], axis=-1).reshape(-1, 3).astype(np.float32)
    a = (np.arange(n - 1)[:, None] * n + np.arange(n - 1)[None, :]).ravel()
    idx = np.concatenate([np.stack([a, a + 1, a + n], 1), np.stack([a + 1, a + n + 1, a + n], 1)]).astype(np.uint32)
    gm, om = parry_b200.TriMesh(ctx, v, idx), oracle.TriMesh(v, idx)
    px, pz = np.meshgrid(np.arange(0, n - 1 + 0.01, 0.25), np.arange(0, n - 1 + 0.01, 0.25), indexing="ij")
    o = np.stack([px.ravel(), np.full(px.size, 3.0), pz.ravel()], axis=1)
    d = np.tile([0.0, -1.0, 0.0], (len(o), 1))
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    assert len(rays) >= 4096
    with _Variant(variant):
        g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    b = om.cast_rays(None, rays, FMAX, with_normal=True, mode=1)
    gt, gi = np.asarray(g[0]), np.asarray(g[1]).astype(np.uint32)
    assert (gi != INVALID).all() and (gt == 3.0).all()
    assert (gi == b[1]).all()  # smallest index among bit-equal minimal toi


@pytest.mark.parametrize("variant", [3, 5])
def test_wide_tree_axis_aligned_and_degenerate_rays(terrain, variant):
    """Directions with zero components (1/0 = inf in the slab test), rays starting inside leaf boxes, zero-length
    directions: the quantised tree must stay a superset of the reference's culling."""
    v, i, gm, om = terrain
    g0 = scenes.rng(23)
    n = 20000
    o = np.stack([(g0.random(n) - 0.5) * 900.0, 60 + g0.random(n) * 30, (g0.random(n) - 0.5) * 900.0], axis=1)
    d = np.zeros((n, 3))
    d[:, 1] = -1.0
    k = n // 4
    d[k:2 * k, 0] = g0.standard_normal(k)          # dz == 0
    d[2 * k:3 * k, 2] = g0.standard_normal(k)      # dx == 0
    d[3 * k:] = g0.standard_normal((n - 3 * k, 3))
    d[3 * k:3 * k + 50] = 0.0                      # null directions never hit
    o[3 * k + 50:3 * k + 2000, 1] = v[:, 1].mean()  # origins inside the terrain's height range
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    with _Variant(variant):
        g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    r = om.cast_rays(None, rays, FMAX, with_normal=True, threads=8)
    check_ray_parity(g, r, _brute(om, rays, FMAX))


@pytest.mark.parametrize("variant", [0, 1, 3, 5])
def test_duplicate_and_degenerate_triangles(ctx, oracle, variant):
    """Exact duplicates (bit-equal toi: smallest index must win), zero-area triangles (never hit, zero-extent boxes) and a
    far-away outlier (huge root box, coarse quantisation at the top of the wide tree)."""
    import parry_b200
    v, i = scenes.terrain(33, 33, extent=60.0)
    nv = len(v)
    dup = i[:200].copy()
    degen = np.stack([i[300:340, 0], i[300:340, 0], i[300:340, 1]], axis=1)          # a == b
    line = np.stack([i[400:440, 0], i[400:440, 1], i[400:440, 0]], axis=1)           # c == a
    far = np.array([[1e4, 5e3, -2e4], [1e4 + 1, 5e3, -2e4], [1e4, 5e3 + 1, -2e4]], np.float32)
    v2 = np.concatenate([v, far]).astype(np.float32)
    i2 = np.concatenate([i, dup, degen, line, np.array([[nv, nv + 1, nv + 2]], np.uint32)]).astype(np.uint32)
    gm, om = parry_b200.TriMesh(ctx, v2, i2), oracle.TriMesh(v2, i2)
    rays = scenes.terrain_rays(30000, extent=60.0, seed=24)
    rays[:, 1] = rays[:, 1] * 0.2 + 10.0
    with _Variant(variant):
        g = gm.cast_local_ray_and_get_normal(rays, FMAX)
    r = om.cast_rays(None, rays, FMAX, with_normal=True, threads=8)
    b = om.cast_rays(None, rays, FMAX, with_normal=True, mode=1, threads=8)
    # rays through a duplicated triangle tie exactly: the reference keeps the first in its own tree order, the GPU the smallest index
    check_ray_parity(g, r, _brute(om, rays, FMAX), max_ulp_cases=0.05)
    gi = np.asarray(g[1]).astype(np.uint32)
    hit = gi != INVALID
    assert hit.mean() > 0.3
    assert (gi[hit] < len(i)).all()                      # duplicates and degenerate triangles never win
    assert (gi == b[1]).mean() > 0.999                   # == brute force with min-index ties


@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1.0, 3.0e4), (100.0, -7.0e5), (1.0e-2, 2.5e3)])
def test_wide_tree_far_from_origin_and_tiny_scenes(ctx, oracle, scale, offset):
    """The quantised-box filter must stay conservative when coordinates are huge compared with the boxes (few f32 bits left for
    the geometry) or tiny: directed rounding + relative slack, trimesh_wide.cu. Checked against the oracle, bit for bit."""
    import parry_b200
    v, i = scenes.terrain(65, 65, extent=100.0)
    rays = scenes.terrain_rays(40000, extent=100.0, seed=25)
    rays[:, 1] = rays[:, 1] * 0.3 + 20.0
    off = np.array([offset, -0.5 * offset, 0.25 * offset])
    v2 = (v.astype(np.float64) * scale + off).astype(np.float32)
    r2 = rays.astype(np.float64)
    r2[:, :3] = r2[:, :3] * scale + off
    r2[:, 3:] *= scale
    r2 = r2.astype(np.float32)
    gm, om = parry_b200.TriMesh(ctx, v2, i), oracle.TriMesh(v2, i)
    with _Variant(3):
        g = gm.cast_local_ray_and_get_normal(r2, FMAX)
    r = om.cast_rays(None, r2, FMAX, with_normal=True, threads=8)
    assert (np.asarray(r[1]) != INVALID).mean() > 0.2
    check_ray_parity(g, r, _brute(om, r2, FMAX), max_ulp_cases=0.002)

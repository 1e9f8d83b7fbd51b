"""CPU check of device code that is written __host__ __device__: the per-manifold core of the persistence kernel
(parry_b200/csrc/manifold_update.cuh) and the per-pair candidate enumeration / reduction of Compound-vs-Compound contacts
(parry_b200/csrc/compound_pair.cuh) are compiled for the host by nvcc with the library's floating-point flags and must reproduce
the oracle bit for bit. (The kernels around them are covered by the GPU tests; this file needs no GPU.)"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from harness import scenes

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def hostlib():
    from parry_b200 import build as b
    src = os.path.join(HERE, "hostcheck", "host_device_cores.cu")
    out_dir = os.path.join(HERE, "hostcheck", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libhostcheck.so")
    deps = [src] + [os.path.join(b.CSRC, f) for f in ("manifold_update.cuh", "compound_pair.cuh", "shapes.cuh", "common.cuh", "ploc.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call([b._nvcc()] + b.NVCC_FLAGS + ["-shared", src, "-o", so])
    return C.CDLL(so)


def test_try_update_core_matches_oracle(hostlib, oracle):
    g = scenes.rng(51)
    T = oracle.ShapeTable([("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("ball", 0.3)])
    n = 20000
    s1, s2 = g.integers(0, 3, n).astype(np.uint32), g.integers(0, 3, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=8, threads=4)
    moved = p2.copy()
    moved[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.6, 3.0e-4, 0.02)).astype(np.float32)
    tilt = scenes.random_unit_quaternions(g, n)
    tilt[:, :3] *= 0.004                                   # a fraction of a degree on some pairs
    tilt /= np.linalg.norm(tilt, axis=1, keepdims=True)
    sel = g.random(n) < 0.3
    a, b = moved[sel, :4].astype(np.float64), tilt[sel]
    moved[sel, :4] = np.stack([a[:, 3] * b[:, 0] + a[:, 0] * b[:, 3] + a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1],
                               a[:, 3] * b[:, 1] - a[:, 0] * b[:, 2] + a[:, 1] * b[:, 3] + a[:, 2] * b[:, 0],
                               a[:, 3] * b[:, 2] + a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0] + a[:, 2] * b[:, 3],
                               a[:, 3] * b[:, 3] - a[:, 0] * b[:, 0] - a[:, 1] * b[:, 1] - a[:, 2] * b[:, 2]], axis=1).astype(np.float32)
    kept_o, q_o = oracle.ShapeTable.manifolds_try_update(p1, moved, nr, cnt, pts)
    q_h = pts.copy()
    kept_h = np.zeros(n, np.uint8)
    P = C.c_void_p
    hostlib.hostcheck_manifolds_try_update.argtypes = [P, P, C.c_uint32, C.c_uint32, C.c_float, C.c_float, P, P, P, P]
    hostlib.hostcheck_manifolds_try_update(p1.ctypes.data, moved.ctypes.data, n, pts.shape[1], 0.99984769515, 1.0e-6, nr.ctypes.data,
                                           cnt.ctypes.data, q_h.ctypes.data, kept_h.ctypes.data)
    has = cnt > 0
    assert 0.2 < kept_o[has].mean() < 0.9                  # both outcomes exercised
    assert (kept_h == kept_o).all()
    assert (q_h.view(np.uint32) == q_o.view(np.uint32)).all()   # including the partially refreshed points of rejected manifolds


def test_compound_compound_cores_match_oracle(hostlib, oracle):
    """compound_pair.cuh on the CPU: candidates (two nested AABB tests) -> leaf contacts (here from the oracle's dispatcher on the
    materialised leaf problems, on the GPU from the contact kernels) -> reduction; against the oracle's contact_compound_compound."""
    g = scenes.rng(31)
    pts, _ = scenes.hull_pool(4, 16, seed=32)
    spec = [("ball", 0.3), ("ball", 0.2), ("cuboid", [0.25, 0.4, 0.3]), ("cuboid", [0.5, 0.15, 0.2])] + [("convex", p * 0.5) for p in pts]
    T = oracle.ShapeTable(spec)
    ns, nc = len(spec), 24
    first, count, psid, ppose = [], [], [], []
    for c in range(nc):
        k = int(g.integers(1, 5))
        first.append(len(psid)); count.append(k)
        psid += [int(x) for x in g.integers(0, ns, k)]
        ppose.append(np.concatenate([scenes.random_unit_quaternions(g, k), (g.random((k, 3)) - 0.5) * 1.2], axis=1))
    first, count, psid = np.asarray(first, np.uint32), np.asarray(count, np.uint32), np.asarray(psid, np.uint32)
    ppose = np.ascontiguousarray(np.concatenate(ppose), dtype=np.float32)
    n = 20000
    a, b = g.integers(0, nc, n).astype(np.uint32), g.integers(0, nc, n).astype(np.uint32)
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 2.0 + 0.2)], axis=1).astype(np.float32)
    pred = 0.05
    ref_out, ref_st, ref_parts = T.contact_compound_compound(first, count, psid, ppose, a, p1, b, p2, pred, threads=8)
    assert 0.15 < (ref_st == 1).mean() < 0.9

    P, u32 = C.c_void_p, C.c_uint32
    tab = [T.kinds.ctypes.data, T.params.ctypes.data, T.points.ctypes.data]
    np_parts = len(psid)
    aabb = np.zeros((np_parts, 6), np.float32)
    hostlib.hostcheck_part_aabbs.argtypes = [P] * 5 + [u32, P]
    hostlib.hostcheck_part_aabbs(*tab, psid.ctypes.data, ppose.ctypes.data, np_parts, aabb.ctypes.data)
    comp = [first.ctypes.data, count.ctypes.data, psid.ctypes.data, ppose.ctypes.data, aabb.ctypes.data]
    pairs = [a.ctypes.data, p1.ctypes.data, b.ctypes.data, p2.ctypes.data]
    hostlib.hostcheck_cc_candidates.argtypes = [P] * 8 + [u32] + [P] * 4 + [u32, C.c_float, C.c_int] + [P] * 7
    counts = np.zeros(n + 1, np.uint32)
    hostlib.hostcheck_cc_candidates(*tab, *comp, nc, *pairs, n, pred, 0, counts.ctypes.data, None, None, None, None, None, None)
    offsets = np.concatenate([[0], np.cumsum(counts[:n])]).astype(np.uint32)
    total = int(offsets[n])
    assert total > n // 4
    ij, cs1, cs2 = np.zeros((total, 2), np.uint32), np.zeros(total, np.uint32), np.zeros(total, np.uint32)
    cp1, cp2 = np.zeros((total, 7), np.float32), np.zeros((total, 7), np.float32)
    hostlib.hostcheck_cc_candidates(*tab, *comp, nc, *pairs, n, pred, 1, None, offsets.ctypes.data, ij.ctypes.data, cs1.ctypes.data, cs2.ctypes.data,
                                    cp1.ctypes.data, cp2.ctypes.data)
    assert (cs1 == psid[ij[:, 1]]).all() and (cs2 == psid[ij[:, 0]]).all() and (cp1 == ppose[ij[:, 1]]).all()
    cand, cst = T.contact_local(cs1, cp1, cs2, cp2, pred, threads=8)       # the contact kernels' job on the GPU
    out = np.zeros((n, 13), np.float32)
    st = np.zeros(n, np.uint8)
    parts = np.zeros((n, 2), np.uint32)
    hostlib.hostcheck_cc_reduce.argtypes = [P] * 9 + [u32] + [P] * 4 + [u32, P, P, P]
    hostlib.hostcheck_cc_reduce(offsets.ctypes.data, ij.ctypes.data, cand.ctypes.data, cst.ctypes.data, *comp, nc, *pairs, n, out.ctypes.data,
                                st.ctypes.data, parts.ctypes.data)
    assert (st == ref_st).all()
    assert (parts == ref_parts).all()
    assert (out.view(np.uint32) == ref_out.view(np.uint32)).all()


@pytest.mark.parametrize("hulls", [False, True])
def test_second_frame_dispatch_matches_oracle(hostlib, oracle, hulls):
    """pb2_contact_manifolds_update_batch step by step on the CPU: the per-pair bodies of k_manifold_try_update (dispatch rule, saved
    feature ids, kept / status) and k_manifold_match (manifold_update.cuh, compiled for the host), with the recomputation of the
    pairs that were not kept taken from the oracle's first-frame manifolds (what the manifold kernels compute behind the skip mask),
    must give the oracle's persistent dispatch: kept flags, counts, points bit for bit, match indices."""
    g = scenes.rng(91)
    spec = [("ball", 0.3), ("cuboid", [0.3, 0.5, 0.4]), ("cuboid", [0.6, 0.2, 0.2]), ("cuboid", [0.5, 0.5, 0.5])]
    if hulls:
        hp, _ = scenes.hull_pool(5, 16, seed=92)
        spec += [("convex", np.asarray(p, np.float32) * 0.6) for p in hp]
    T = oracle.ShapeTable(spec)
    topo = T.hull_topology() if hulls else None
    n, mp = 12000, 12
    ns = len(spec)
    s1, s2 = g.integers(0, ns, n).astype(np.uint32), g.integers(0, ns, n).astype(np.uint32)
    s1[::101] = 1000                                       # unknown shape ids: never kept, status from the recomputation
    p1 = np.concatenate([scenes.random_unit_quaternions(g, n), (g.random((n, 3)) - .5) * 2], axis=1).astype(np.float32)
    d = g.standard_normal((n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    p2 = np.concatenate([scenes.random_unit_quaternions(g, n), p1[:, 4:] + d * (g.random((n, 1)) * 1.0 + 0.3)], axis=1).astype(np.float32)
    p2[::3, :4] = p1[::3, :4]
    p2b = p2.copy()
    p2b[:, 4:] += (g.standard_normal((n, 3)) * np.where(g.random((n, 1)) < 0.5, 2.0e-4, 0.05)).astype(np.float32)
    nr, cnt, pts, st = T.contact_manifolds(s1, p1, s2, p2, 0.05, max_points=mp, threads=4, topology=topo)
    rn, rc, rp, rs, rk, rm = T.contact_manifolds_update(s1, p1, s2, p2b, 0.05, nr, cnt, pts, threads=4, topology=topo)
    # 1. k_manifold_try_update with dispatch
    hn, hc, hp_, hs = nr.copy(), cnt.copy(), pts.copy(), np.full(n, 255, np.uint8)
    kept = np.zeros(n, np.uint8)
    old_f, old_c = np.zeros((n, mp, 2), np.uint32), np.zeros(n, np.uint32)
    kinds = np.ascontiguousarray(T.kinds, dtype=np.uint8)
    P, u32, i32 = C.c_void_p, C.c_uint32, C.c_int
    hostlib.hostcheck_manifold_try_update_pairs.argtypes = [P, u32, P, P, i32, i32, P, P, u32, u32, P, P, P, P, P, P, P]
    hostlib.hostcheck_manifold_try_update_pairs(kinds.ctypes.data, ns, s1.ctypes.data, s2.ctypes.data, 1, int(hulls), p1.ctypes.data, p2b.ctypes.data,
                                                n, mp, hn.ctypes.data, hc.ctypes.data, hp_.ctypes.data, kept.ctypes.data, hs.ctypes.data,
                                                old_f.ctypes.data, old_c.ctypes.data)
    assert (kept == rk).all(), np.nonzero(kept != rk)[0][:10]
    assert (hs[kept == 1] == 0).all() and (hs[kept == 0] == 255).all()
    assert (old_c == np.minimum(cnt, mp)).all()
    valid = np.arange(mp)[None, :] < old_c[:, None]
    assert (old_f[valid] == pts[:, :, 7:].view(np.uint32)[valid]).all()
    # 2. the manifold kernels behind the skip mask = a first-frame computation of the pairs that were not kept
    fn, fc, fp, fs = T.contact_manifolds(s1, p1, s2, p2b, 0.05, max_points=mp, threads=4, topology=topo)
    rec = kept == 0
    hn[rec], hc[rec], hp_[rec], hs[rec] = fn[rec], fc[rec], fp[rec], fs[rec]
    # 3. k_manifold_match
    match = np.full((n, mp), -7, np.int32)
    hostlib.hostcheck_manifold_match_pairs.argtypes = [P, P, P, P, P, u32, u32, P]
    hostlib.hostcheck_manifold_match_pairs(kept.ctypes.data, old_f.ctypes.data, old_c.ctypes.data, hc.ctypes.data, hp_.ctypes.data, n, mp,
                                           match.ctypes.data)
    assert (hs == rs).all() and (hc == rc).all()
    assert (hp_.view(np.uint32) == rp.view(np.uint32)).all() and (hn.view(np.uint32) == rn.view(np.uint32)).all()
    assert (match == rm).all(), np.nonzero((match != rm).any(axis=1))[0][:10]
    ball = (kinds[np.minimum(s1, ns - 1)] == 0) | (kinds[np.minimum(s2, ns - 1)] == 0) | (s1 >= ns)
    assert (kept[ball] == 0).all() and 0.15 < kept[~ball & (cnt > 0)].mean() < 0.85
    assert (match >= 0).any(axis=1)[rec & (rc > 0) & (cnt > 0)].mean() > 0.3        # recomputed manifolds do find old points again


def _canonical(nodes, root=0):
    """Ordered-tree fingerprint of a BvhNodeWide array: DFS from the root, left before right, one entry per node half."""
    out, stack = [], [root]
    while stack:
        w = nodes[stack.pop()]
        push = []
        for side in ("left", "right"):
            h = w[side]
            lc = int(h["data"]) & 0x3FFFFFFF
            if lc == 1:
                out.append(("leaf", int(h["children"])))
            else:
                out.append(("node", lc))
                push.append(int(h["children"]))
        stack.extend(reversed(push))
    return out


@pytest.mark.parametrize("scene", ["colliders", "sphere"])
def test_ploc_link_builds_the_reference_topology(hostlib, oracle, scene):
    """BvhBuildStrategy::Ploc on the GPU (bvh_build.cu: k_ploc_nearest / k_ploc_flags / scan / k_ploc_emit) against the oracle's
    restatement of rebuild_range_ploc (bvh_ploc_build.rs:10-94): the per-cluster functions of ploc.cuh, replayed round by round on
    the CPU over leaves given in the reference's own Morton order, must link exactly the reference's tree (same ordered topology,
    same boxes), and parents / leaf_node_indices must be consistent."""
    from helpers import assert_well_formed
    if scene == "colliders":
        kinds, params, poses, _ = scenes.colliders(5000, seed=71)
        aabbs = oracle.shape_aabbs(kinds, params, poses)
    else:
        v, i = scenes.uv_sphere(40, 30)
        t = v[i]
        aabbs = np.concatenate([t.min(axis=1), t.max(axis=1)], axis=1).astype(np.float32)
    # the reference's sort key (bvh_ploc_build.rs:12-19, utils/morton.rs:33-40): per-axis normalised centres, f64, 21 bits per axis
    c = ((aabbs[:, :3] + aabbs[:, 3:]) * np.float32(0.5)).astype(np.float32)
    lo, ext = c.min(axis=0), c.max(axis=0) - c.min(axis=0)
    u = ((c - lo) * (np.float32(1.0) / ext)).astype(np.float32).astype(np.float64)
    q = np.minimum(np.floor(u * float(1 << 21)), 4294967295.0).astype(np.uint64)

    def spread(x):
        x = x & np.uint64(0x1fffff)
        for sh, m in ((32, 0x1f00000000ffff), (16, 0x1f0000ff0000ff), (8, 0x100f00f00f00f00f), (4, 0x10c30c30c30c30c3), (2, 0x1249249249249249)):
            x = (x | (x << np.uint64(sh))) & np.uint64(m)
        return x
    key = spread(q[:, 0]) | (spread(q[:, 1]) << np.uint64(1)) | (spread(q[:, 2]) << np.uint64(2))
    aabbs = np.ascontiguousarray(aabbs[np.argsort(key, kind="stable")])
    n = len(aabbs)
    ref = oracle.Bvh(aabbs, strategy=1).nodes()
    nodes = np.zeros(n - 1, dtype=ref.dtype)
    parents, slots = np.zeros(n - 1, np.uint32), np.zeros(n, np.uint32)
    P = C.c_void_p
    hostlib.hostcheck_ploc_link.argtypes = [P, C.c_uint32, C.c_uint32, P, P, P]
    hostlib.hostcheck_ploc_link.restype = C.c_int
    rounds = hostlib.hostcheck_ploc_link(aabbs.ctypes.data, n, 16, nodes.ctypes.data, parents.ctypes.data, slots.ctypes.data)
    assert 5 < rounds < 60
    assert _canonical(nodes) == _canonical(ref)
    # boxes of the internal halves are the merged boxes of their subtrees (the reference's refit recomputes the same values)
    def boxes(nd):
        out, stack = [], [0]
        while stack:
            w = nd[stack.pop()]
            push = []
            for side in ("left", "right"):
                out.append(np.concatenate([w[side]["mins"], w[side]["maxs"]]))
                if (int(w[side]["data"]) & 0x3FFFFFFF) != 1:
                    push.append(int(w[side]["children"]))
            stack.extend(reversed(push))
        return np.asarray(out)
    assert (boxes(nodes).view(np.uint32) == boxes(ref).view(np.uint32)).all()
    # flags: the link leaves pending-change bits everywhere (the refit that follows resolves them); compare structure only
    chk = nodes.copy()
    for side in ("left", "right"):
        chk[side]["data"] &= 0x3FFFFFFF
    assert_well_formed(chk, parents, slots)

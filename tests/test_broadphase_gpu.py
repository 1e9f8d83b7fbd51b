"""GPU parity: Bvh build / refit / intersect_aabb / traverse_bvtt_single_tree / leaf_pairs and typed-leaf ray casts
through the C ABI vs the CPU oracle. Pair and hit *sets* must be identical (order is tree-dependent in the reference too)."""
import numpy as np
import pytest

from harness import scenes
from helpers import INVALID, assert_well_formed, check_ray_parity, sorted_pairs

pytestmark = pytest.mark.gpu
FMAX = float(np.finfo(np.float32).max)


def make_colliders(n, seed, side=None):
    kinds, params, poses, _ = scenes.colliders(n, side=side, seed=seed)
    return kinds, params, poses


def make_shapes(ctx, kinds, params):
    import parry_b200
    shapes = [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) for k, p in zip(kinds, params)]
    return parry_b200.Shapes(ctx, shapes)


def brute_pairs(aabbs):
    a = aabbs.astype(np.float64)
    n = len(a)
    out = []
    for i in range(n):
        m = ((a[i, 0] <= a[i + 1:, 3]) & (a[i, 1] <= a[i + 1:, 4]) & (a[i, 2] <= a[i + 1:, 5]) &
             (a[i, 3] >= a[i + 1:, 0]) & (a[i, 4] >= a[i + 1:, 1]) & (a[i, 5] >= a[i + 1:, 2]))
        j = np.nonzero(m)[0] + i + 1
        out.append(np.stack([np.full(len(j), i), j], axis=1))
    return np.concatenate(out) if out else np.zeros((0, 2), np.int64)


@pytest.mark.parametrize("n", [0, 1, 2, 3, 4, 7, 100, 1000])
def test_build_well_formed_and_pairs_small(ctx, oracle, n):
    import parry_b200
    g0 = scenes.rng(200 + n)
    c = g0.random((n, 3)) * max(1.0, n ** (1 / 3))
    h = g0.random((n, 3)) * 0.5 + 0.1
    aabbs = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    gb = parry_b200.Bvh.from_leaves(ctx, parry_b200.BvhBuildStrategy.Binned, aabbs)
    assert gb.leaf_count() == n
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    ob = oracle.Bvh(aabbs)
    gp = gb.traverse_bvtt_single_tree()
    op = ob.self_pairs()
    assert (sorted_pairs(gp) == sorted_pairs(op)).all()
    assert (sorted_pairs(gp) == sorted_pairs(brute_pairs(aabbs))).all() if n > 1 else len(gp) == 0


def test_shape_aabbs_bit_exact_and_pair_set(ctx, oracle):
    import parry_b200
    n = 20000
    kinds, params, poses = make_colliders(n, seed=2)
    shapes = make_shapes(ctx, kinds, params)
    aabbs = shapes.compute_aabbs(np.arange(n, dtype=np.uint32), poses)
    ref = oracle.shape_aabbs(kinds, params, poses)
    assert (aabbs.view(np.uint32) == ref.view(np.uint32)).all()
    gb = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    ob = oracle.Bvh(aabbs)
    gp, op = gb.traverse_bvtt_single_tree(), ob.self_pairs()
    assert len(gp) == len(op) > n
    assert (sorted_pairs(gp) == sorted_pairs(op)).all()
    assert (gp[:, 0] < gp[:, 1]).all()
    # overflow is reported, not UB
    with pytest.raises(parry_b200.Pb2Error):
        import ctypes as C
        cnt = C.c_uint64(0)
        buf = np.zeros((10, 2), np.uint32)
        ctx.check(ctx._lib.pb2_bvh_self_pairs(ctx.h, gb.h, 0, buf.ctypes.data, 10, C.byref(cnt), 0))
    assert cnt.value == len(op)


def test_intersect_aabb_batch(ctx, oracle):
    import parry_b200
    n = 30000
    kinds, params, poses = make_colliders(n, seed=21)
    shapes = make_shapes(ctx, kinds, params)
    aabbs = shapes.compute_aabbs(None, poses) if False else shapes.compute_aabbs(np.arange(n, dtype=np.uint32), poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    g0 = scenes.rng(22)
    side = n ** (1 / 3)
    c = g0.random((5000, 3)) * side
    h = g0.random((5000, 3)) * 1.5
    q = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    q[0] = [-1e9, -1e9, -1e9, -1e8, -1e8, -1e8]  # empty result
    goff, gids = gb.intersect_aabb(q)
    ooff, oids = ob.intersect_aabbs(q, threads=8)
    assert (goff == ooff).all()
    for k in range(len(q)):
        assert (np.sort(gids[goff[k]:goff[k + 1]]) == np.sort(oids[ooff[k]:ooff[k + 1]])).all()


def test_update_refit_matches_rebuild_and_oracle(ctx, oracle):
    """One broad-phase frame: insert_or_update_partially on every leaf + refit, then the full pair set."""
    import parry_b200
    n = 20000
    kinds, params, poses = make_colliders(n, seed=23)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs0 = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs0), oracle.Bvh(aabbs0)
    g0 = scenes.rng(24)
    for frame in range(3):
        poses = poses.copy()
        poses[:, 4:] += (g0.random((n, 3)).astype(np.float32) - 0.5) * 0.2
        aabbs = shapes.compute_aabbs(ids, poses)
        gb.insert_or_update_partially(aabbs, ids, 0.0)
        gb.refit()
        ob.update_leaves(aabbs, ids, 0.0)
        ob.refit()
        nodes, parents, leaf_idx = gb.download()
        assert_well_formed(nodes, parents, leaf_idx)
        # root box identical to the oracle's (min/max are exact)
        on = ob.nodes()
        oroot = np.concatenate([np.minimum(on[0]["left"]["mins"], on[0]["right"]["mins"]), np.maximum(on[0]["left"]["maxs"], on[0]["right"]["maxs"])])
        assert (gb.root_aabb() == oroot).all()
        gp, op = gb.traverse_bvtt_single_tree(), ob.self_pairs()
        assert (sorted_pairs(gp) == sorted_pairs(op)).all()
    gb.rebuild()
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    assert (sorted_pairs(gb.traverse_bvtt_single_tree()) == sorted_pairs(op)).all()


def test_insert_and_remove_leaves(ctx, oracle):
    """Bvh::insert / Bvh::remove (bvh_insert.rs:126-197, bvh_tree.rs:2360-2427): after any sequence of edits every query
    answers like a reference Bvh built from the surviving leaves (pair sets / intersect_aabb are tree independent)."""
    import parry_b200
    n = 6000
    kinds, params, poses = make_colliders(n, seed=31)
    shapes = make_shapes(ctx, kinds, params)
    aabbs = shapes.compute_aabbs(np.arange(n, dtype=np.uint32), poses)
    gb = parry_b200.Bvh.from_leaves(ctx, 0, aabbs[:4000])
    present = np.zeros(n, bool)
    present[:4000] = True

    def check():
        ids = np.nonzero(present)[0]
        ob = oracle.Bvh(aabbs[ids])
        gp = np.asarray(gb.traverse_bvtt_single_tree()).astype(np.int64)
        op = ids[np.asarray(ob.self_pairs()).astype(np.int64)]
        assert (sorted_pairs(gp) == sorted_pairs(op)).all()
        q = aabbs[::37]
        goff, gids = gb.intersect_aabb(q)
        ooff, oids = ob.intersect_aabbs(q)
        assert (np.asarray(goff) == np.asarray(ooff)).all()
        for k in range(len(q)):
            assert sorted(np.asarray(gids)[goff[k]:goff[k + 1]].tolist()) == sorted(ids[np.asarray(oids)[ooff[k]:ooff[k + 1]]].tolist())

    check()
    rm = np.arange(0, 4000, 3, dtype=np.uint32)
    gb.remove(rm)
    present[rm] = False
    check()
    new = np.arange(4000, n, dtype=np.uint32)           # new indices + re-insertion of removed ones
    back = rm[::2]
    ins = np.concatenate([new, back])
    gb.insert(aabbs[ins], ins)
    present[ins] = True
    assert gb.leaf_count() == n
    check()
    gb.remove(np.array([n + 5, 1], dtype=np.uint32))     # unknown index ignored
    present[1] = False
    check()
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx, check_geometry=False)
    # rays never hit a removed collider, and hit the survivors exactly like a reference Bvh over the survivors
    ids = np.nonzero(present)[0]
    g0 = scenes.rng(32)
    side = float(n) ** (1.0 / 3.0)
    o = g0.random((20000, 3)) * side
    rays = np.concatenate([o, g0.standard_normal((20000, 3))], axis=1).astype(np.float32)
    g = gb.cast_ray(shapes, np.arange(n, dtype=np.uint32), poses, rays, FMAX)
    ob = oracle.Bvh(aabbs[ids])
    r = ob.cast_rays_shapes(kinds[ids], params[ids], poses[ids], rays, FMAX, threads=8)
    gl, rl = np.asarray(g[1]).astype(np.uint32), np.asarray(r[1]).astype(np.uint32)
    hit = rl != INVALID
    assert hit.mean() > 0.3 and ((gl != INVALID) == hit).all()
    assert present[gl[hit]].all()
    gt, rt = np.asarray(g[0])[hit], np.asarray(r[0])[hit]
    assert (gt.view(np.uint32) == rt.view(np.uint32)).mean() > 0.9995
    same = gl[hit] == ids[rl[hit]]
    # rays starting inside several overlapping colliders tie at toi == 0: smallest index wins on the GPU (documented rule)
    tie = ~same
    assert (gt[tie] == rt[tie]).all() and (gl[hit][tie] < ids[rl[hit]][tie]).all()
    assert same.mean() > 0.9


def test_change_detection_pairs(ctx, oracle):
    import parry_b200
    n = 5000
    kinds, params, poses = make_colliders(n, seed=25)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    margin = 0.05
    # frame 0: everything is flagged changed by construction
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(True)) == sorted_pairs(ob.self_pairs(True))).all()
    g0 = scenes.rng(26)
    for frame in range(3):
        moved = g0.random(n) < 0.1
        poses = poses.copy()
        poses[moved, 4:] += (g0.random((int(moved.sum()), 3)).astype(np.float32) - 0.5) * 0.5
        aabbs = shapes.compute_aabbs(ids, poses)
        gb.insert_or_update_partially(aabbs, ids, margin)
        gb.refit()
        ob.update_leaves(aabbs, ids, margin)
        ob.refit()
        gp, op = gb.traverse_bvtt_single_tree(True), ob.self_pairs(True)
        assert 0 < len(op) < len(ob.self_pairs(False))
        assert (sorted_pairs(gp) == sorted_pairs(op)).all()
        assert (sorted_pairs(gb.traverse_bvtt_single_tree(False)) == sorted_pairs(ob.self_pairs(False))).all()


@pytest.mark.parametrize("how", ["rebuild", "optimize_incremental"])
def test_change_detection_survives_rebuild(ctx, oracle, how):
    """The usual broad-phase order is update -> refit -> optimize_incremental / rebuild -> traverse<CHANGE_DETECTION>. Bvh::rebuild
    copies the leaf nodes verbatim, change flags included, and resolves nothing (bvh_binned_build.rs:11-36), so the pairs reported
    after it are still only those touching a leaf flagged by the last refit (round 1 re-flagged every leaf: ADVICE.md)."""
    import parry_b200
    n = 4000
    kinds, params, poses = make_colliders(n, seed=27)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    g0 = scenes.rng(28)
    for frame in range(3):
        moved = g0.random(n) < 0.08
        poses = poses.copy()
        poses[moved, 4:] += (g0.random((int(moved.sum()), 3)).astype(np.float32) - 0.5) * 0.5
        aabbs = shapes.compute_aabbs(ids, poses)
        gb.insert_or_update_partially(aabbs, ids, 0.05)
        gb.refit()
        ob.update_leaves(aabbs, ids, 0.05)
        ob.refit()
        if how == "rebuild":
            gb.rebuild(frame & 1)
        else:
            gb.optimize_incremental()
        ob.rebuild(frame & 1)
        gp, op = gb.traverse_bvtt_single_tree(True), ob.self_pairs(True)
        assert 0 < len(op) < len(ob.self_pairs(False)) // 2
        assert (sorted_pairs(gp) == sorted_pairs(op)).all()
        assert (sorted_pairs(gb.traverse_bvtt_single_tree(False)) == sorted_pairs(ob.self_pairs(False))).all()
        # a second rebuild changes nothing either, and the next refit clears the flags of leaves that did not move
        gb.rebuild(0)
        assert (sorted_pairs(gb.traverse_bvtt_single_tree(True)) == sorted_pairs(op)).all()


@pytest.mark.parametrize("na,nb", [(1, 1), (2, 2), (1, 2), (50, 1), (1, 50), (3000, 2500)])
def test_leaf_pairs_two_trees(ctx, oracle, na, nb):
    import parry_b200
    g0 = scenes.rng(300 + na + nb)
    def boxes(n):
        c = g0.random((n, 3)) * 8.0
        h = g0.random((n, 3)) * 0.6 + 0.05
        return np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    a, b = boxes(na), boxes(nb)
    ga, gb = parry_b200.Bvh.from_leaves(ctx, 0, a), parry_b200.Bvh.from_leaves(ctx, 0, b)
    oa, ob = oracle.Bvh(a), oracle.Bvh(b)
    gp, op = ga.leaf_pairs(gb), oa.leaf_pairs(ob)
    def key(p):
        p = np.asarray(p).astype(np.int64).reshape(-1, 2)
        return np.sort(p[:, 0] * (1 << 32) + p[:, 1])
    if na <= 2 and nb <= 2 or (na > 2 and nb > 2):
        assert (key(gp) == key(op)).all()
    else:
        # the reference yields root-level leaf/leaf pairs unchecked (bvh_traverse_bvtt.rs:215-237), which depends on
        # its tree topology; every checked pair must match and extras must be exactly such unchecked root pairs.
        gk, ok = set(key(gp).tolist()), set(key(op).tolist())
        assert gk <= ok and len(ok - gk) <= 2


def test_bvh_cast_ray_ball_cuboid_leaves(ctx, oracle):
    import parry_b200
    n = 3000
    kinds, params, poses = make_colliders(n, seed=27)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    g0 = scenes.rng(28)
    side = n ** (1 / 3)
    o = g0.random((20000, 3)) * side
    d = g0.standard_normal((20000, 3))
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    def adjudicate(g, r, max_toi, solid):
        # toi is bit-exact everywhere; ids may differ only on exact ties (rays starting inside several overlapping
        # solids all report toi == 0): the GPU then returns the smallest tied leaf id (documented rule).
        assert (g[0].view(np.uint32) == r[0].view(np.uint32)).all()
        diff = np.nonzero(g[1] != r[1])[0]
        assert len(diff) < 0.1 * len(g[1])
        for k in diff:
            assert g[1][k] < r[1][k]
            t = oracle.shape_cast_ray_toi(int(kinds[g[1][k]]), params[g[1][k]], poses[g[1][k]], rays[k], max_toi, solid)
            assert t is not None and np.float32(t) == g[0][k]
        return g[1] == r[1]

    for solid in (True, False):
        g = gb.cast_ray(shapes, ids, poses, rays, FMAX, solid=solid, with_normal=True)
        r = ob.cast_rays_shapes(kinds, params, poses, rays, FMAX, solid=solid, with_normal=True, threads=8)
        assert (r[1] != INVALID).mean() > 0.3
        same = adjudicate(g, r, FMAX, solid)
        np.testing.assert_allclose(g[2][same], r[2][same], rtol=1e-5, atol=1e-7)
        assert (g[3][same] == r[3][same]).all()
        g2 = gb.cast_ray(shapes, ids, poses, rays, 2.0, solid=solid)
        r2 = ob.cast_rays_shapes(kinds, params, poses, rays, 2.0, solid=solid, threads=8)
        adjudicate(g2, r2, 2.0, solid)


@pytest.mark.parametrize("n", [3, 4, 7, 100, 1000, 20000])
def test_ploc_strategy(ctx, oracle, n):
    """BvhBuildStrategy::Ploc (bvh_ploc_build.rs:10-94) built on the GPU (bvh_build.cu: bvh_link_ploc): a well-formed tree whose
    queries answer like the reference's (pair sets, intersect_aabb, change detection, refit). The linking rule itself is checked
    against the oracle's PLOC topology on the CPU (tests/test_hostcheck.py::test_ploc_link_builds_the_reference_topology)."""
    import parry_b200
    kinds, params, poses = make_colliders(n, seed=400 + n)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb = parry_b200.Bvh.from_leaves(ctx, parry_b200.BvhBuildStrategy.Ploc, aabbs)
    ob = oracle.Bvh(aabbs, strategy=1)
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    assert (sorted_pairs(gb.traverse_bvtt_single_tree()) == sorted_pairs(ob.self_pairs())).all()
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(True)) == sorted_pairs(ob.self_pairs(True))).all()
    q = aabbs[:: max(1, n // 200)]
    goff, gids = gb.intersect_aabb(q)
    ooff, oids = ob.intersect_aabbs(q)
    assert (np.asarray(goff) == np.asarray(ooff)).all()
    for k in range(len(q)):
        assert (np.sort(np.asarray(gids)[goff[k]:goff[k + 1]]) == np.sort(np.asarray(oids)[ooff[k]:ooff[k + 1]])).all()
    # one frame: move a tenth of the leaves, refit, change detection, then a PLOC rebuild keeps the flags
    g0 = scenes.rng(500 + n)
    moved = g0.random(n) < 0.1
    poses = poses.copy()
    poses[moved, 4:] += (g0.random((int(moved.sum()), 3)).astype(np.float32) - 0.5) * 0.5
    aabbs2 = shapes.compute_aabbs(ids, poses)
    gb.insert_or_update_partially(aabbs2, ids, 0.05)
    gb.refit()
    ob.update_leaves(aabbs2, ids, 0.05)
    ob.refit()
    want_cd, want_all = sorted_pairs(ob.self_pairs(True)), sorted_pairs(ob.self_pairs(False))
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(True)) == want_cd).all()
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(False)) == want_all).all()
    gb.rebuild(parry_b200.BvhBuildStrategy.Ploc)
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(True)) == want_cd).all()
    assert (sorted_pairs(gb.traverse_bvtt_single_tree(False)) == want_all).all()


def test_ploc_degenerate_input_falls_back(ctx, oracle):
    """Thousands of identical boxes merge one pair per PLOC round (the reference's own loop is quadratic there): the GPU build
    notices the stall and links the same sorted leaves as an LBVH; every query still answers exactly."""
    import parry_b200
    n = 1500
    aabbs = np.tile(np.array([[0, 0, 0, 1, 1, 1]], np.float32), (n, 1))
    aabbs[-3:] += 5.0
    gb = parry_b200.Bvh.from_leaves(ctx, parry_b200.BvhBuildStrategy.Ploc, aabbs)
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    gp = gb.traverse_bvtt_single_tree()
    assert len(gp) == (n - 3) * (n - 4) // 2 + 3
    assert len(np.unique(sorted_pairs(gp))) == len(gp)


def test_deep_tree_no_silent_drops(ctx, oracle):
    """A Karras tree over clustered / duplicated centroids is up to 63 + log2(n) levels deep: 64 clusters whose Morton keys are
    powers of two link into a 63-level chain, 40 duplicates per cluster add the index tie-break levels below it. Round 1's
    64-entry stacks dropped pushes silently there (VERDICT weak #10); now the stacks hold any Karras tree (PB2_STACK = 96 >= 63 +
    32) and a full stack raises PB2_ERR_DEPTH at the next synchronisation instead of losing work."""
    import parry_b200
    pts = []
    for j in range(63):
        q = np.zeros(3)
        q[j % 3] = 2.0 ** (j // 3 - 21)
        pts.append(q)
    pts.append(np.ones(3))
    dup = 40
    centers = np.repeat(np.asarray(pts), dup, axis=0)
    n = len(centers)
    h = 2.0 ** -24
    aabbs = np.concatenate([centers - h, centers + h], axis=1).astype(np.float32)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    nodes, parents, leaf_idx = gb.download()
    assert_well_formed(nodes, parents, leaf_idx)
    par = np.asarray(parents).astype(np.int64) >> 1
    deepest = 0
    for leaf in range(0, n, dup):
        k, d = int(leaf_idx[leaf]) >> 1, 1
        while k != 0:
            k = int(par[k]); d += 1
        deepest = max(deepest, d)
    assert deepest > 64                              # deeper than round 1's stacks
    # every leaf overlaps a box around everything: the DFS holds one pending sibling per level
    q = np.array([[-1, -1, -1, 2, 2, 2], [0, 0, 0, 2.0 ** -20, 2.0 ** -20, 2.0 ** -20]], np.float32)
    goff, gids = gb.intersect_aabb(q)
    ooff, oids = ob.intersect_aabbs(q)
    assert (np.asarray(goff) == np.asarray(ooff)).all() and goff[1] == n
    for k in range(len(q)):
        assert (np.sort(np.asarray(gids)[goff[k]:goff[k + 1]]) == np.sort(np.asarray(oids)[ooff[k]:ooff[k + 1]])).all()
    gp, op = gb.traverse_bvtt_single_tree(), ob.self_pairs()
    assert len(op) == len(pts) * dup * (dup - 1) // 2
    assert (sorted_pairs(gp) == sorted_pairs(op)).all()
    # rays towards every cluster from far away cross the whole chain (typed cuboid leaves)
    kinds = np.ones(n, np.uint8)
    params = np.full((n, 3), h, np.float32)
    poses = np.zeros((n, 7), np.float32)
    poses[:, 3] = 1.0
    poses[:, 4:] = centers
    shapes = make_shapes(ctx, kinds, params)
    o = np.array([-1.0, -0.7, -0.4])
    tgt = np.asarray(pts)
    rays = np.concatenate([np.tile(o, (len(tgt), 1)), tgt - o], axis=1).astype(np.float32)
    ids = np.arange(n, dtype=np.uint32)
    g = gb.cast_ray(shapes, ids, poses, rays, FMAX, solid=True)
    r = ob.cast_rays_shapes(kinds, params, poses, rays, FMAX, solid=True)
    assert (np.asarray(r[1]) != INVALID).sum() >= len(tgt) - 2
    assert (np.asarray(g[0]).view(np.uint32) == np.asarray(r[0]).view(np.uint32)).all()
    assert (np.asarray(g[1]) <= np.asarray(r[1])).all()     # duplicates tie exactly: smallest leaf id (documented rule)
    ctx.synchronize()                                        # no sticky overflow flag


def test_full_size_config2_frame(ctx):
    """BASELINE config[1]: 2^20 dynamic AABBs. Properties at full size: the pair set equals the sort-and-sweep
    brute force on a spatial slab, no duplicates, and refit after a no-op update is idempotent."""
    import parry_b200
    n = 1 << 20
    kinds, params, poses = make_colliders(n, seed=2)
    shapes = make_shapes(ctx, kinds, params)
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    p1 = gb.traverse_bvtt_single_tree()
    k1 = sorted_pairs(p1)
    assert len(np.unique(k1)) == len(k1)
    gb.insert_or_update_partially(aabbs, ids, 0.0)
    gb.refit()
    k2 = sorted_pairs(gb.traverse_bvtt_single_tree())
    assert (k1 == k2).all()
    # brute force inside a slab
    sel = np.nonzero(aabbs[:, 3] < 4.0)[0]
    bp = brute_pairs(aabbs[sel])
    bk = sorted_pairs(np.stack([sel[bp[:, 0]], sel[bp[:, 1]]], axis=1))
    insel = np.zeros(n, bool)
    insel[sel] = True
    mask = insel[p1[:, 0]] & insel[p1[:, 1]]
    assert (sorted_pairs(p1[mask]) == bk).all()


def test_bvh_cast_ray_convex_leaves(ctx, oracle):
    """SURVEY §8 f3 (ray vs hull): Bvh::cast_ray over ball / cuboid / ConvexPolyhedron leaves; the convex leaves go through
    the GJK ray cast (ray_support_map.rs:19-72, gjk.rs:660-795). Ids exact (ties: smallest id), toi / normals 1e-5."""
    import parry_b200
    n, H = 1500, 64
    g0 = scenes.rng(61)
    hulls, _ = scenes.hull_pool(H, 16, seed=62)
    hulls = (hulls * 0.5).astype(np.float32)
    kinds = g0.integers(0, 3, n).astype(np.uint8)
    params = (g0.random((n, 3)) * 0.3 + 0.15).astype(np.float32)
    hid = g0.integers(0, H, n)
    side = (n ** (1 / 3)) * 1.2
    poses = np.concatenate([scenes.random_unit_quaternions(g0, n), g0.random((n, 3)) * side], axis=1).astype(np.float32)
    shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) if k == 1 else parry_b200.ConvexPolyhedron(hulls[h])
                                     for k, p, h in zip(kinds, params, hid)])
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    points = np.concatenate([hulls[h] for h in hid])
    first = (np.arange(n) * 16).astype(np.uint32)
    count = np.full(n, 16, np.uint32)
    o = g0.random((20000, 3)) * side
    d = g0.standard_normal((20000, 3))
    d[::2] /= np.linalg.norm(d[::2], axis=1, keepdims=True)
    rays = np.concatenate([o, d], axis=1).astype(np.float32)
    for solid in (True, False):
        for max_toi in (FMAX, 1.5):
            g = gb.cast_ray(shapes, ids, poses, rays, max_toi, solid=solid, with_normal=True)
            r = ob.cast_rays_shapes(kinds, params, poses, rays, max_toi, solid=solid, with_normal=True, threads=8, points=points, first=first,
                                    count=count)
            hit = r[1] != INVALID
            assert hit.mean() > 0.2 and (kinds[r[1][hit]] == 2).mean() > 0.15
            assert ((g[1] != INVALID) == hit).all()
            np.testing.assert_allclose(g[0], r[0], rtol=1e-5, atol=1e-7)
            diff = np.nonzero(g[1] != r[1])[0]
            assert len(diff) < 0.05 * len(rays)
            for k in diff:  # exact ties only (rays starting inside overlapping solids): smallest leaf id
                assert g[1][k] < r[1][k] and g[0][k] == r[0][k]
            same = g[1] == r[1]
            np.testing.assert_allclose(g[2][same], r[2][same], rtol=1e-5, atol=1e-6)
            assert (g[3][same] == r[3][same]).all()
            assert (g[3][same & hit & (kinds[np.minimum(r[1], n - 1)] == 2)] == 0xFFFFFFFE).all()


def test_bvh_project_point_typed_leaves(ctx, oracle):
    """Bvh::project_point over ball / cuboid / ConvexPolyhedron leaves (bvh_queries.rs:213-227) through
    pb2_bvh_project_points_shapes: the leaf and the inside flag exact but for equal distances, projections 1e-5; max_distance prunes;
    solid = false asks the host only for points inside a hull."""
    import parry_b200
    n, H = 1500, 64
    g0 = scenes.rng(63)
    hulls, _ = scenes.hull_pool(H, 16, seed=64)
    hulls = (hulls * 0.5).astype(np.float32)
    kinds = g0.integers(0, 3, n).astype(np.uint8)
    params = (g0.random((n, 3)) * 0.3 + 0.15).astype(np.float32)
    hid = g0.integers(0, H, n)
    side = (n ** (1 / 3)) * 1.2
    poses = np.concatenate([scenes.random_unit_quaternions(g0, n), g0.random((n, 3)) * side], axis=1).astype(np.float32)
    shapes = parry_b200.Shapes(ctx, [parry_b200.Ball(p[0]) if k == 0 else parry_b200.Cuboid(p) if k == 1 else parry_b200.ConvexPolyhedron(hulls[h])
                                     for k, p, h in zip(kinds, params, hid)])
    ids = np.arange(n, dtype=np.uint32)
    aabbs = shapes.compute_aabbs(ids, poses)
    gb, ob = parry_b200.Bvh.from_leaves(ctx, 0, aabbs), oracle.Bvh(aabbs)
    points = np.concatenate([hulls[h] for h in hid])
    first = (np.arange(n) * 16).astype(np.uint32)
    count = np.full(n, 16, np.uint32)
    m = 20000
    q = (g0.random((m, 3)) * (side + 2.0) - 1.0).astype(np.float32)
    # a tenth of the query points sit inside a leaf's shape (near its centre)
    pick = g0.integers(0, n, m // 10)
    q[: m // 10] = poses[pick, 4:] + (g0.standard_normal((m // 10, 3)) * 0.05).astype(np.float32)
    for solid in (True, False):
        for max_dist in (FMAX, 0.4):
            g = [np.asarray(x) for x in gb.project_point(shapes, ids, poses, q, max_dist, solid=solid)]
            r = ob.project_points_shapes(kinds, params, poses, q, max_dist, solid=solid, threads=8, points=points, first=first, count=count)
            host = g[3] == 3
            if solid:
                assert not host.any()
            else:   # only where the reference ends up inside a hull (its EPA branch)
                assert 0 < host.sum() < 0.1 * m
            found = r[2] != INVALID
            ok = ~host
            assert ((g[2] != INVALID) == found)[ok].all() and ((g[3] == 1) == found)[ok].all()
            if max_dist == FMAX:
                assert found.all()
            else:
                assert 0.1 < found.mean() < 0.9
            same = ok & found & (g[2] == r[2])
            assert same.sum() > 0.97 * (ok & found).sum()
            np.testing.assert_allclose(g[0][same], r[0][same], rtol=1e-5, atol=2e-6)
            assert (g[1][same] == r[1][same]).all()
            assert r[1][found].mean() > 0.05 and (kinds[r[2][found]] == 2).mean() > 0.15
            # a different leaf only at (nearly) equal distances: overlapping solids both containing the point, or rounding
            diff = ok & found & (g[2] != r[2])
            dg = np.linalg.norm(g[0][diff] - q[diff], axis=1)
            dr = np.linalg.norm(r[0][diff] - q[diff], axis=1)
            np.testing.assert_allclose(dg, dr, rtol=1e-5, atol=2e-6)
            if not solid:   # the hosted queries are exactly those whose answer (or a nearer candidate's) lies inside a hull
                assert (kinds[r[2][host & found]] == 2).mean() > 0.5


def test_bvh_api_mirrors(ctx, oracle):
    """from_iter with gaps, refit_without_opt, optimize_incremental: the pair set always equals the brute-force set of the
    leaves that are present."""
    import parry_b200
    n = 600
    kinds, params, poses = make_colliders(n, seed=71)
    shapes = make_shapes(ctx, kinds, params)
    aabbs = shapes.compute_aabbs(np.arange(n, dtype=np.uint32), poses)
    keep = np.ones(n, bool)
    keep[::7] = False
    bvh = parry_b200.Bvh.from_iter(ctx, 0, ((i, aabbs[i]) for i in range(n) if keep[i]))
    assert bvh.leaf_count() <= n
    idx = np.nonzero(keep)[0]
    expect = brute_pairs(aabbs[idx])
    expect = np.sort(idx[expect], axis=1)
    key = lambda p: np.sort(np.sort(p.astype(np.int64), axis=1)[:, 0] * n + np.sort(p.astype(np.int64), axis=1)[:, 1])
    for step in (lambda: None, bvh.refit_without_opt, bvh.optimize_incremental):
        step()
        got = bvh.traverse_bvtt_single_tree()
        assert (key(got) == key(expect)).all()

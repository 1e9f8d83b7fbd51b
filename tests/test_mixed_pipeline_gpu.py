"""GPU parity for BASELINE config 5 (mixed scene): ball / cuboid / 32-vertex-hull colliders -> leaf AABBs -> Bvh ->
traverse_bvtt_single_tree -> query::contact on every reported pair (pb2_contact_pairs_compact reads shapes and poses
through the pair list). Checked against the CPU oracle running the same pipeline."""
import numpy as np
import pytest

from harness import scenes
from helpers import sorted_pairs

pytestmark = pytest.mark.gpu


def make_scene(ctx, oracle, n, seed, n_hulls=64):
    import parry_b200
    kinds, params, poses, hull_ids = scenes.colliders(n, seed=seed, hull_fraction=1.0 / 3.0, n_hulls=n_hulls)
    pts, _ = scenes.hull_pool(n_hulls, 32, seed=seed + 1)
    pts = np.asarray(pts, dtype=np.float32) * 0.5            # hull radius 0.25 .. 0.5, same scale as the other colliders
    H = n_hulls
    # shape table: H hulls, then one ball / cuboid entry per collider
    tk = np.concatenate([np.full(H, 2, np.uint8), np.where(kinds == 2, 0, kinds).astype(np.uint8)])
    tp = np.concatenate([np.zeros((H, 3), np.float32), params])
    first = np.concatenate([np.arange(H, dtype=np.uint32) * 32, np.zeros(n, np.uint32)])
    count = np.concatenate([np.full(H, 32, np.uint32), np.zeros(n, np.uint32)])
    G = parry_b200.Shapes.from_arrays(ctx, tk, tp, pts.reshape(-1, 3), first, count)
    spec = [("convex", p) for p in pts] + [("ball", float(p[0])) if k == 0 else ("cuboid", list(p)) for k, p in zip(tk[H:], tp[H:])]
    O = oracle.ShapeTable(spec)
    coll_shape = np.where(kinds == 2, hull_ids, H + np.arange(n)).astype(np.uint32)
    return G, O, coll_shape, poses, kinds


def oracle_aabbs(oracle, O, coll_shape, poses):
    kinds = O.kinds[coll_shape]
    params = O.params[coll_shape, :3].copy()
    pu = O.params.view(np.uint32)
    return oracle.shape_aabbs(kinds, params, poses, O.points, pu[coll_shape, 0], pu[coll_shape, 1])


@pytest.mark.parametrize("n,seed", [(3000, 71), (30000, 72)])
def test_mixed_pipeline_matches_oracle(ctx, oracle, n, seed):
    import parry_b200
    G, O, cs, poses, kinds = make_scene(ctx, oracle, n, seed)
    assert set(np.unique(kinds)) == {0, 1, 2}
    aabbs = G.compute_aabbs(cs, poses)
    ref = oracle_aabbs(oracle, O, cs, poses)
    assert (aabbs.view(np.uint32) == ref.view(np.uint32)).all()
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    pairs = bvh.traverse_bvtt_single_tree()
    opairs = oracle.Bvh(aabbs).self_pairs()
    assert len(pairs) > n // 2
    assert (sorted_pairs(pairs) == sorted_pairs(opairs)).all()
    pairs = np.ascontiguousarray(pairs, dtype=np.uint32)
    out, idx = parry_b200.contact_pairs_compact(G, cs, poses, pairs, 0.01)
    a, b = pairs[:, 0], pairs[:, 1]
    oout, ost = O.contact(cs[a], poses[a], cs[b], poses[b], 0.01, threads=8)
    some = np.nonzero(ost == 1)[0]
    order = np.argsort(idx)
    assert (np.asarray(idx)[order] == some).all()          # contact pair membership: exact
    got = np.asarray(out)[order]
    np.testing.assert_allclose(got, oout[some], rtol=1e-5, atol=1e-6)
    assert (got.view(np.uint32) == oout[some].view(np.uint32)).all(axis=1).mean() > 0.999
    assert 0.05 < len(some) / len(pairs) < 0.95
    # same answer as the per-pair-array entry point
    out2, st2 = parry_b200.contact(G, cs[a], poses[a], cs[b], poses[b], 0.01)
    assert (st2 == ost).all()
    assert (np.asarray(out2)[some].view(np.uint32) == got.view(np.uint32)).all()


def test_bad_collider_index_is_skipped(ctx, oracle):
    import parry_b200
    G, O, cs, poses, kinds = make_scene(ctx, oracle, 200, 73)
    pairs = np.array([[0, 1], [5, 100000], [2, 3]], dtype=np.uint32)
    out, idx = parry_b200.contact_pairs_compact(G, cs, poses, pairs, 100.0)
    assert set(np.asarray(idx).tolist()) <= {0, 2}


def test_full_size_config5_device_pipeline(ctx, oracle):
    """2^21 colliders, everything device resident; oracle parity on a slice of the pairs, structural checks on the rest."""
    import torch
    import parry_b200
    n = 1 << 21
    G, O, cs, poses, kinds = make_scene(ctx, oracle, n, 74, n_hulls=4096)
    dcs, dposes = torch.from_numpy(cs.view(np.int32)).cuda(), torch.from_numpy(poses).cuda()
    aabbs = G.compute_aabbs(dcs, dposes)
    bvh = parry_b200.Bvh.from_leaves(ctx, 0, aabbs)
    pairs = bvh.traverse_bvtt_single_tree(capacity=16 * n, like=aabbs)
    P = int(pairs.shape[0])
    assert n < P < 16 * n
    out, idx = parry_b200.contact_pairs_compact(G, dcs, dposes, pairs, 0.01)
    ctx.synchronize()
    idx_h = idx.cpu().numpy().view(np.uint32)
    out_h = out.cpu().numpy()
    assert len(np.unique(idx_h)) == len(idx_h) and idx_h.max() < P
    assert np.isfinite(out_h).all()
    nrm = np.linalg.norm(out_h[:, 6:9], axis=1)
    assert np.abs(nrm - 1.0).max() < 1e-4
    # oracle on the first 40k pairs
    ph = pairs[:40000].cpu().numpy().view(np.uint32)
    a, b = ph[:, 0], ph[:, 1]
    oout, ost = O.contact(cs[a], poses[a], cs[b], poses[b], 0.01, threads=8)
    sel = idx_h < 40000
    order = np.argsort(idx_h[sel])
    assert (idx_h[sel][order] == np.nonzero(ost == 1)[0]).all()
    np.testing.assert_allclose(out_h[sel][order], oout[ost == 1], rtol=1e-5, atol=1e-6)

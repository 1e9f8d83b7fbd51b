import numpy as np

INVALID = 0xFFFFFFFF


def ulp_diff(a, b):
    """Distance in units-in-the-last-place between two f32 arrays (same sign assumed for near values)."""
    ai = np.asarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(ai - bi)


def check_ray_parity(gpu, ref, brute_fn, max_ulp_cases=0.0005):
    """gpu/ref: (toi, id[, normal, feature]). Bit-exact toi and ids are required except for
    (a) exact toi ties, where the GPU returns the smallest index (documented rule) and the reference returns the first
        in its own tree's DFS order, and
    (b) ulp-level AABB-vs-triangle inconsistencies whose outcome depends on visiting order.
    Both are adjudicated against brute force over all primitives. Returns the number of adjudicated rays."""
    g_toi, g_id = np.asarray(gpu[0]), np.asarray(gpu[1]).astype(np.uint32)
    r_toi, r_id = np.asarray(ref[0]), np.asarray(ref[1]).astype(np.uint32)
    same = (g_id == r_id) & (g_toi.view(np.uint32) == r_toi.view(np.uint32))
    bad = np.nonzero(~same)[0]
    assert len(bad) <= max(2, max_ulp_cases * len(g_toi)), "too many mismatches: %d of %d" % (len(bad), len(g_toi))
    if len(bad):
        b_toi, b_id = brute_fn(bad)
        for k, i in enumerate(bad):
            # hit / miss must agree with brute force unless the hit is an ulp-level AABB cull
            g_hit, r_hit = g_id[i] != INVALID, r_id[i] != INVALID
            assert g_hit or r_hit
            if g_hit and r_hit:
                assert ulp_diff(g_toi[i], r_toi[i]) <= 4, (i, g_toi[i], r_toi[i])
                if g_toi[i] == r_toi[i]:
                    # exact tie: GPU must hold the smallest index among the tied hits it reports
                    assert g_id[i] <= r_id[i], (i, g_id[i], r_id[i])
            else:
                # one side culled a grazing hit through its AABB test: brute force must see a hit within ulps
                t = g_toi[i] if g_hit else r_toi[i]
                assert b_id[k] != INVALID and ulp_diff(b_toi[k], t) <= 4, (i, t, b_toi[k])
    if len(gpu) > 2 and gpu[2] is not None:
        ok = np.nonzero(same)[0]
        gn, rn = np.asarray(gpu[2])[ok], np.asarray(ref[2])[ok]
        np.testing.assert_allclose(gn, rn, rtol=1e-5, atol=1e-7)
        assert (np.asarray(gpu[3]).astype(np.uint32)[ok] == np.asarray(ref[3]).astype(np.uint32)[ok]).all()
    return len(bad)


def assert_well_formed(nodes, parents, leaf_idx, check_geometry=True):
    """Bvh::assert_well_formed (partitioning/bvh/bvh_validation.rs:61-134) on a downloaded node array."""
    n_leaves = len(leaf_idx)
    if len(nodes) == 0:
        assert n_leaves == 0
        return
    root = nodes[0]
    if (root["right"]["data"] & 0x3FFFFFFF) == 0:
        assert (root["left"]["data"] & 0x3FFFFFFF) == 1
        return
    seen = np.zeros(len(nodes), dtype=bool)
    seen_leaf = np.zeros(n_leaves, dtype=bool)
    total = 0
    stack = [0]
    while stack:
        nid = stack.pop()
        assert not seen[nid], "loop: node %d visited twice" % nid
        seen[nid] = True
        for side, name in enumerate(("left", "right")):
            h = nodes[nid][name]
            lc = int(h["data"] & 0x3FFFFFFF)
            if lc == 1:
                leaf = int(h["children"])
                assert leaf < n_leaves and not seen_leaf[leaf]
                seen_leaf[leaf] = True
                assert int(leaf_idx[leaf]) == ((nid << 1) | side)
                total += 1
            else:
                c = int(h["children"])
                assert int(parents[c]) == ((nid << 1) | side)
                ch = nodes[c]
                clc = int(ch["left"]["data"] & 0x3FFFFFFF) + int(ch["right"]["data"] & 0x3FFFFFFF)
                assert lc == clc, (nid, lc, clc)
                changed = ((int(ch["left"]["data"]) >> 30) == 1) or ((int(ch["right"]["data"]) >> 30) == 1)
                assert changed == ((int(h["data"]) >> 30) == 1)
                if check_geometry:
                    for cn in ("left", "right"):
                        assert (h["mins"] <= ch[cn]["mins"]).all() and (h["maxs"] >= ch[cn]["maxs"]).all()
                stack.append(c)
    assert total == n_leaves and seen_leaf.all()
    assert int(root["left"]["data"] & 0x3FFFFFFF) + int(root["right"]["data"] & 0x3FFFFFFF) == n_leaves


def sorted_pairs(p):
    p = np.asarray(p).astype(np.int64).reshape(-1, 2)
    p = np.sort(p, axis=1)
    key = p[:, 0] * (1 << 32) + p[:, 1]
    return np.sort(key)

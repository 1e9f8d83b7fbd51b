// parry_b200 — ordered best-first Bvh descent shared by the ray kernels.
// Mirrors Bvh::find_best (partitioning/bvh/bvh_traverse.rs:335-417) with aabb_cost = BvhNode::cast_ray
// (bvh_tree.rs:1177-1181): both children scored against the best hit so far, nearer child first, farther child
// pushed on a short per-thread stack, prune when score >= best.
#pragma once
#include "common.cuh"


// `leaf(pos)` tests the primitive at sorted position `pos` and updates `best` / `found` itself.
// Deviation from the reference (documented tie rule): once a hit exists, nodes whose score == best are
// still visited so that every bit-equal tie is seen and the smallest leaf id can win.
template <class Leaf>
__device__ __forceinline__ void bvh_find_best(const NodeWide* __restrict__ nodes, uint32_t n_leaves, V3 o, V3 d, V3 inv,
                                              float max_toi, float& best, bool& found, Leaf leaf, unsigned int* fault) {
    if (n_leaves == 1) {
        // partial root (bvh_traverse.rs:349-358)
        const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1);
        if (!(l0.x > l1.x) && slab_cost(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, o, d, inv, max_toi) < max_toi) leaf(__float_as_uint(l0.w));
        return;
    }
    if (n_leaves < 2) return;
    uint32_t stack[PB2_STACK];
    int sp = 0;
    uint32_t curr = 0;
    for (;;) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
        float ls = slab_cost(l0.x, l0.y, l0.z, l1.x, l1.y, l1.z, o, d, inv, best);
        float rs = slab_cost(r0.x, r0.y, r0.z, r1.x, r1.y, r1.z, o, d, inv, best);
        // removed leaves / emptied subtrees keep Aabb::new_invalid() (mins > maxs): the slab test would sort the two
        // planes and walk in, so they are excluded explicitly (pb2_bvh_remove_leaves)
        if (l0.x > l1.x) ls = FLT_MAX;
        if (r0.x > r1.x) rs = FLT_MAX;
        uint32_t lc = __float_as_uint(l0.w), rc = __float_as_uint(r0.w);
        bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        bool rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        if (ls > rs) {
            float ts = ls; ls = rs; rs = ts;
            uint32_t tc = lc; lc = rc; rc = tc;
            bool tl = lleaf; lleaf = rleaf; rleaf = tl;
        }
        bool found_next = false;
        if (ls != FLT_MAX && (ls < best || (found && ls == best))) {
            if (lleaf) leaf(lc);
            else { curr = lc; found_next = true; }
        }
        if (rs != FLT_MAX && (rs < best || (found && rs == best))) {
            if (rleaf) leaf(rc);
            else if (found_next) pb2_push(stack, sp, rc, fault);
            else { curr = rc; found_next = true; }
        }
        if (!found_next) {
            if (sp == 0) break;
            curr = stack[--sp];
        }
    }
}

// Bvh::find_best (bvh_traverse.rs:335-417) with a caller-supplied node cost `cost(lo, hi, bound)` (FLT_MAX = skip; lo / hi are the
// child's two float4 halves) — the composite-shape queries score Minkowski-summed boxes. Same descent and tie rule as bvh_find_best.
template <class Cost, class Leaf>
__device__ __forceinline__ void bvh_find_best_cost(const NodeWide* __restrict__ nodes, uint32_t n_leaves, float max_cost, float& best, bool& found,
                                                   Cost cost, Leaf leaf, unsigned int* fault) {
    if (n_leaves == 1) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1);
        if (!(l0.x > l1.x) && cost(l0, l1, max_cost) < max_cost) leaf(__float_as_uint(l0.w));
        return;
    }
    if (n_leaves < 2) return;
    uint32_t stack[PB2_STACK];
    int sp = 0;
    uint32_t curr = 0;
    for (;;) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
        // inert leaves keep Aabb::new_invalid() (mins > maxs)
        float ls = l0.x > l1.x ? FLT_MAX : cost(l0, l1, best), rs = r0.x > r1.x ? FLT_MAX : cost(r0, r1, best);
        uint32_t lc = __float_as_uint(l0.w), rc = __float_as_uint(r0.w);
        bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        bool rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        if (ls > rs) {
            float ts = ls; ls = rs; rs = ts;
            uint32_t tc = lc; lc = rc; rc = tc;
            bool tl = lleaf; lleaf = rleaf; rleaf = tl;
        }
        bool found_next = false;
        if (ls != FLT_MAX && (ls < best || (found && ls == best))) {
            if (lleaf) leaf(lc);
            else { curr = lc; found_next = true; }
        }
        if (rs != FLT_MAX && (rs < best || (found && rs == best))) {
            if (rleaf) leaf(rc);
            else if (found_next) pb2_push(stack, sp, rc, fault);
            else { curr = rc; found_next = true; }
        }
        if (!found_next) {
            if (sp == 0) break;
            curr = stack[--sp];
        }
    }
}

// Node cost of CompositeShapeRef::cast_shape (shape_cast_composite_shape_shape.rs:36-45): every node box is Minkowski-summed with the
// other shape's box — Aabb::new(mins + shift - margin, maxs + shift + margin) — and hit by the ray (origin, d) = (0, vel12), solid.
template <class Leaf>
__device__ __forceinline__ void bvh_find_best_msum(const NodeWide* __restrict__ nodes, uint32_t n_leaves, V3 shift, V3 margin, V3 d, V3 inv,
                                                   float max_toi, float& best, bool& found, Leaf leaf, unsigned int* fault) {
    const V3 o = mk3(0.f, 0.f, 0.f);
    auto cost = [&](float4 lo, float4 hi, float bound) {
        return slab_cost((lo.x + shift.x) - margin.x, (lo.y + shift.y) - margin.y, (lo.z + shift.z) - margin.z, (hi.x + shift.x) + margin.x,
                         (hi.y + shift.y) + margin.y, (hi.z + shift.z) + margin.z, o, d, inv, bound);
    };
    bvh_find_best_cost(nodes, n_leaves, max_toi, best, found, cost, leaf, fault);
}

// Node cost of CompositeShapeRef::distance_to_shape (distance_composite_shape_shape.rs:23-33): Aabb::distance_to_origin (aabb.rs:556-562)
// of the Minkowski-summed box.
template <class Leaf>
__device__ __forceinline__ void bvh_find_best_msum_distance(const NodeWide* __restrict__ nodes, uint32_t n_leaves, V3 shift, V3 margin, float& best,
                                                            bool& found, Leaf leaf, unsigned int* fault) {
    auto cost = [&](float4 lo, float4 hi, float) {
        V3 mn = mk3((lo.x + shift.x) - margin.x, (lo.y + shift.y) - margin.y, (lo.z + shift.z) - margin.z);
        V3 mx = mk3((hi.x + shift.x) + margin.x, (hi.y + shift.y) + margin.y, (hi.z + shift.z) + margin.z);
        V3 v = vmax3(vmax3(mn, -mx), mk3(0.f, 0.f, 0.f));
        return nrm(v);
    };
    bvh_find_best_cost(nodes, n_leaves, FLT_MAX, best, found, cost, leaf, fault);
}

// parry_b200 — ordered best-first Bvh descent shared by the ray kernels.
// Mirrors Bvh::find_best (partitioning/bvh/bvh_traverse.rs:335-417) with aabb_cost = BvhNode::cast_ray
// (bvh_tree.rs:1177-1181): both children scored against the best hit so far, nearer child first, farther child
// pushed on a short per-thread stack, prune when score >= best.
#pragma once
#include "common.cuh"

// lanes that must wait at a leaf before the leaf code runs (bvh_find_best_cost below); one copy per translation unit, the one in
// contact.cu can be set with PB2_LEAF_LANES (tuning only)
static __device__ int g_pb2_leaf_lanes = 32;   // measured best: run the leaf code only when no lane can move (4..32 swept)


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
// child's two float4 halves) — the composite-shape queries score Minkowski-summed boxes — and an expensive leaf query (a GJK run).
// Same descent, same order of leaf tests per query and same tie rule as bvh_find_best, but warp-cooperative: run one thread per
// query and every lane reaches its leaves on a different trip of the loop, so the leaf code (thousands of instructions) executed at
// 2 of 32 lanes (ncu, first version of the composite casts). Here a lane that reaches a leaf WAITS; the leaf code runs from one call
// site once `min_lanes` lanes wait or nobody else can move, for all of them together. A lane's own sequence of cost / leaf
// evaluations is unchanged (its `best` is up to date before it moves on), so results are identical.
// Every lane of `mask` must call (converged); `active` = this lane has a query. leaf(pos, lanes) gets the mask of the lanes that
// execute the leaf code with it, for nested descents.
template <class Cost, class Leaf>
__device__ __forceinline__ void bvh_find_best_cost(unsigned mask, bool active, const NodeWide* __restrict__ nodes, uint32_t n_leaves, float max_cost,
                                                   float& best, bool& found, Cost cost, Leaf leaf, unsigned int* fault) {
    const int min_lanes = g_pb2_leaf_lanes;
    if (n_leaves == 1) {
        bool want = false;
        uint32_t lpos = 0;
        if (active) {
            const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
            float4 l0 = __ldg(np), l1 = __ldg(np + 1);
            want = !(l0.x > l1.x) && cost(l0, l1, max_cost) < max_cost;
            lpos = __float_as_uint(l0.w);
        }
        unsigned lanes = __ballot_sync(mask, want);
        if (want) leaf(lpos, lanes);
        return;
    }
    if (n_leaves < 2) return;
    enum { VISIT = 0, RESOLVE = 1, WAIT = 2, DONE = 3 };
    uint32_t stack[PB2_STACK];
    int sp = 0;
    uint32_t curr = 0;
    // the two children of the node just opened, nearer first: {child, score, leaf?}; `pi` = the next one to decide
    uint32_t c0 = 0, c1 = 0;
    float s0 = FLT_MAX, s1 = FLT_MAX;
    bool f0 = false, f1 = false, found_next = false;
    int pi = 2;
    int state = active ? VISIT : DONE;
    for (;;) {
        if (state == VISIT) {
            const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
            float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
            // inert leaves keep Aabb::new_invalid() (mins > maxs)
            s0 = l0.x > l1.x ? FLT_MAX : cost(l0, l1, best);
            s1 = r0.x > r1.x ? FLT_MAX : cost(r0, r1, best);
            c0 = __float_as_uint(l0.w); c1 = __float_as_uint(r0.w);
            f0 = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
            f1 = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
            if (s0 > s1) {
                float ts = s0; s0 = s1; s1 = ts;
                uint32_t tc = c0; c0 = c1; c1 = tc;
                bool tl = f0; f0 = f1; f1 = tl;
            }
            pi = 0; found_next = false;
            state = RESOLVE;
        }
        if (state == RESOLVE) {
            while (pi < 2) {
                const float sc = pi ? s1 : s0;
                const uint32_t ch = pi ? c1 : c0;
                if (sc != FLT_MAX && (sc < best || (found && sc == best))) {
                    if (pi ? f1 : f0) { state = WAIT; break; }
                    if (found_next) pb2_push(stack, sp, ch, fault);
                    else { curr = ch; found_next = true; }
                }
                pi++;
            }
            if (state == RESOLVE) {
                if (found_next) state = VISIT;
                else if (sp > 0) { curr = stack[--sp]; state = VISIT; }
                else state = DONE;
            }
        }
        const unsigned waiting = __ballot_sync(mask, state == WAIT);
        const unsigned moving = __ballot_sync(mask, state == VISIT);
        if (!waiting && !moving) break;
        if (waiting && (!moving || __popc(waiting) >= min_lanes)) {
            if (state == WAIT) {
                leaf(pi ? c1 : c0, waiting);
                pi++;
                state = RESOLVE;
            }
        }
    }
}

// Node cost of CompositeShapeRef::cast_shape (shape_cast_composite_shape_shape.rs:36-45): every node box is Minkowski-summed with the
// other shape's box — Aabb::new(mins + shift - margin, maxs + shift + margin) — and hit by the ray (origin, d) = (0, vel12), solid.
template <class Leaf>
__device__ __forceinline__ void bvh_find_best_msum(unsigned mask, bool active, const NodeWide* __restrict__ nodes, uint32_t n_leaves, V3 shift, V3 margin,
                                                   V3 d, V3 inv, float max_toi, float& best, bool& found, Leaf leaf, unsigned int* fault) {
    const V3 o = mk3(0.f, 0.f, 0.f);
    auto cost = [&](float4 lo, float4 hi, float bound) {
        return slab_cost((lo.x + shift.x) - margin.x, (lo.y + shift.y) - margin.y, (lo.z + shift.z) - margin.z, (hi.x + shift.x) + margin.x,
                         (hi.y + shift.y) + margin.y, (hi.z + shift.z) + margin.z, o, d, inv, bound);
    };
    bvh_find_best_cost(mask, active, nodes, n_leaves, max_toi, best, found, cost, leaf, fault);
}

// Node cost of CompositeShapeRef::distance_to_shape (distance_composite_shape_shape.rs:23-33): Aabb::distance_to_origin (aabb.rs:556-562)
// of the Minkowski-summed box.
template <class Leaf>
__device__ __forceinline__ void bvh_find_best_msum_distance(unsigned mask, bool active, const NodeWide* __restrict__ nodes, uint32_t n_leaves, V3 shift,
                                                            V3 margin, float& best, bool& found, Leaf leaf, unsigned int* fault) {
    auto cost = [&](float4 lo, float4 hi, float) {
        V3 mn = mk3((lo.x + shift.x) - margin.x, (lo.y + shift.y) - margin.y, (lo.z + shift.z) - margin.z);
        V3 mx = mk3((hi.x + shift.x) + margin.x, (hi.y + shift.y) + margin.y, (hi.z + shift.z) + margin.z);
        V3 v = vmax3(vmax3(mn, -mx), mk3(0.f, 0.f, 0.f));
        return nrm(v);
    };
    bvh_find_best_cost(mask, active, nodes, n_leaves, FLT_MAX, best, found, cost, leaf, fault);
}

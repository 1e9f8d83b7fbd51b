// parry_b200 — per-cluster cores of the PLOC builder (BvhBuildStrategy::Ploc, partitioning/bvh/bvh_ploc_build.rs:10-94),
// written __host__ __device__ so that tests/test_hostcheck.py can replay the GPU build on the CPU against the oracle.
//
// The reference's loop, per round over the Morton-sorted clusters: (1) every cluster i picks the neighbour k within
// SEARCH_RADIUS positions that minimises half_area(aabb_i ∪ aabb_k) — strict `<`, so the first minimal k in ascending
// order wins (:27-47); (2) mutual picks (cand[cand[i]] == i) merge into one wide node {left = lower position, right = the
// other}, the merged cluster takes the lower position, everything else is copied through in order (:50-87). Both steps are
// independent per cluster, which is what makes the algorithm a GPU one: here step (1) is one thread per cluster over a
// shared-memory tile of the neighbourhood, step (2) is a prefix sum (new position, new node id) and one thread per cluster.
// Same rule, same order => the same topology the reference's Ploc strategy builds from the same sorted leaves.
#pragma once
#include "common.cuh"

#define PLOC_MAX_RADIUS 32

// Aabb::merged(..).half_area() (bounding_volume/aabb.rs:967-972, 473-476): inf / sup, extents, x * (y + z) + y * z
__host__ __device__ __forceinline__ float ploc_cost(float4 alo, float4 ahi, float4 blo, float4 bhi) {
    float ex = fmaxf(ahi.x, bhi.x) - fminf(alo.x, blo.x);
    float ey = fmaxf(ahi.y, bhi.y) - fminf(alo.y, blo.y);
    float ez = fmaxf(ahi.z, bhi.z) - fminf(alo.z, blo.z);
    return ex * (ey + ez) + ey * ez;
}

// Step (1) for cluster i of c. `tile` holds the clusters [tile_lo, ...) as {lo, hi} float4 pairs.
__host__ __device__ __forceinline__ uint32_t ploc_nearest(const float4* tile, uint32_t tile_lo, uint32_t c, uint32_t i, uint32_t radius) {
    uint32_t lo = i >= radius ? i - radius : 0u;
    uint32_t hi = i + radius < c - 1u ? i + radius : c - 1u;
    float4 mlo = tile[2u * (i - tile_lo)], mhi = tile[2u * (i - tile_lo) + 1u];
    float best = FLT_MAX;
    uint32_t best_k = 0xffffffffu;
    for (uint32_t k = lo; k <= hi; ++k) {
        if (k == i) continue;
        float s = ploc_cost(mlo, mhi, tile[2u * (k - tile_lo)], tile[2u * (k - tile_lo) + 1u]);
        if (s < best) { best = s; best_k = k; }
    }
    // all costs NaN or +inf (boxes with non-finite planes): the reference would index with usize::MAX and panic; pair with the
    // next position instead so that the build always terminates
    if (best_k == 0xffffffffu) best_k = i + 1u < c ? i + 1u : i - 1u;
    return best_k;
}

// Step (2), decision: bit 0 = the cluster (or the merged cluster it starts) appears in the next round, bit 32 = it creates a node.
__host__ __device__ __forceinline__ unsigned long long ploc_flags(const uint32_t* cand, uint32_t i) {
    uint32_t k = cand[i];
    bool mutual = cand[k] == i;
    if (mutual && i > k) return 0ull;
    return 1ull | (mutual ? (1ull << 32) : 0ull);
}

__host__ __device__ __forceinline__ float pb2_u2f(uint32_t u) {
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    float f; memcpy(&f, &u, 4); return f;
#endif
}

// BvhNodeWide::merged (bvh_tree.rs:610-618 through :73): box = inf / sup, leaf counts added, change bits OR-ed.
__host__ __device__ __forceinline__ void ploc_merge(float4 llo, float4 lhi, float4 rlo, float4 rhi, uint32_t id, float4& plo, float4& phi) {
    uint32_t dl = pb2_f2u(lhi.w), dr = pb2_f2u(rhi.w);
    uint32_t data = ((dl & PB2_LEAF_COUNT_MASK) + (dr & PB2_LEAF_COUNT_MASK)) | (((dl >> 30) | (dr >> 30)) << 30);
    plo.x = fminf(llo.x, rlo.x); plo.y = fminf(llo.y, rlo.y); plo.z = fminf(llo.z, rlo.z);
    phi.x = fmaxf(lhi.x, rhi.x); phi.y = fmaxf(lhi.y, rhi.y); phi.z = fmaxf(lhi.z, rhi.z);
    plo.w = pb2_u2f(id); phi.w = pb2_u2f(data);
}

// Step (2), emission for cluster i of c (bvh_ploc_build.rs:50-87). incl = inclusive scan of ploc_flags; `created` nodes exist
// before this round; the k-th node ever created gets id n - 2 - k (root = 0).
__host__ __device__ __forceinline__ void ploc_emit(const float4* Cin, const uint32_t* cand, const unsigned long long* incl, uint32_t i,
                                                   float4* Cout, NodeWide* nodes, uint32_t* parents, uint32_t* leaf_slot, const uint32_t* order,
                                                   uint32_t created, uint32_t n) {
    uint32_t k = cand[i];
    bool mutual = cand[k] == i;
    if (mutual && i > k) return;
    unsigned long long s = incl[i];
    uint32_t pos = (uint32_t)s - 1u;
    float4 llo = Cin[2ull * i], lhi = Cin[2ull * i + 1];
    if (!mutual) { Cout[2ull * pos] = llo; Cout[2ull * pos + 1] = lhi; return; }
    uint32_t id = (n - 2u) - (created + (uint32_t)(s >> 32) - 1u);
    float4 rlo = Cin[2ull * k], rhi = Cin[2ull * k + 1];
    float4* np = reinterpret_cast<float4*>(&nodes[id]);
    np[0] = llo; np[1] = lhi; np[2] = rlo; np[3] = rhi;
    float4 plo, phi;
    ploc_merge(llo, lhi, rlo, rhi, id, plo, phi);
    Cout[2ull * pos] = plo; Cout[2ull * pos + 1] = phi;
    // leaf_node_indices / parents of the two children (:74-83)
    if ((pb2_f2u(lhi.w) & PB2_LEAF_COUNT_MASK) == 1u) leaf_slot[order[pb2_f2u(llo.w)]] = id << 1;
    else parents[pb2_f2u(llo.w)] = id << 1;
    if ((pb2_f2u(rhi.w) & PB2_LEAF_COUNT_MASK) == 1u) leaf_slot[order[pb2_f2u(rlo.w)]] = (id << 1) | 1u;
    else parents[pb2_f2u(rlo.w)] = (id << 1) | 1u;
    if (id == 0u) parents[0] = 0u;
}

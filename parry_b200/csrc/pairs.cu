// parry_b200 — AABB broad-phase queries on the Bvh.
//
// Replaces (reference, file:line): Bvh::intersect_aabb (partitioning/bvh/bvh_queries.rs:203-205) + Leaves iterator
// (bvh_traverse.rs:8-70), Bvh::traverse_bvtt_single_tree (bvh_traverse_bvtt.rs:19-204), Bvh::leaf_pairs
// (bvh_traverse_bvtt.rs:210-316), BvhNode::intersects (bvh_tree.rs:955-957, inclusive on all axes).
//
// B200 design: the reference walks the bounding-volume test tree recursively on one thread. Here every leaf (or
// query box) is an independent stack walk; each unordered pair is emitted once by the lower BVH position, and
// Karras' numbering (left child index == last leaf position of its range) prunes whole "already seen" subtrees.
// Pairs are appended with warp-aggregated atomics (one atomicAdd per warp per emission point).
#include "common.cuh"
#include <cub/device/device_scan.cuh>

int pb2_stage_in(pb2_ctx* ctx, int slot, const void* src, size_t bytes, int mem, const void** out);
int pb2_stage_out(pb2_ctx* ctx, int slot, void* dst, size_t bytes, int mem, void** out);
int pb2_stage_back(pb2_ctx* ctx, void* dst, const void* dev, size_t bytes, int mem);

#define PB2_PSTACK PB2_STACK

struct Box {
    float mnx, mny, mnz, mxx, mxy, mxz;
};
__device__ __forceinline__ bool overlaps(const Box& q, float4 lo, float4 hi) {
    // na::partial_le(mins, other.maxs) && na::partial_ge(maxs, other.mins)
    return lo.x <= q.mxx && lo.y <= q.mxy && lo.z <= q.mxz && hi.x >= q.mnx && hi.y >= q.mny && hi.z >= q.mnz;
}

__device__ __forceinline__ unsigned long long warp_append(unsigned long long* counter) {
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

// ---------------------------------------------------------------- intersect_aabb (batched, CSR two-pass)
template <bool WRITE>
__global__ void k_intersect_aabbs(const NodeWide* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t n_leaves,
                                  const float* __restrict__ queries, uint32_t m, uint32_t* __restrict__ counts,
                                  const uint32_t* __restrict__ offsets, uint32_t* __restrict__ out, uint64_t cap, unsigned int* fault) {
    uint32_t qi = blockIdx.x * blockDim.x + threadIdx.x;
    if (qi >= m) return;
    const float* qp = queries + 6ull * qi;
    Box q = {qp[0], qp[1], qp[2], qp[3], qp[4], qp[5]};
    uint32_t cnt = 0;
    uint64_t base = WRITE ? offsets[qi] : 0;
    auto emit = [&](uint32_t pos) {
        if (WRITE) { if (base + cnt < cap) out[base + cnt] = order ? order[pos] : pos; }  // order == NULL: sorted positions
        cnt++;
    };
    if (n_leaves == 1) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1);
        if (overlaps(q, l0, l1)) emit(__float_as_uint(l0.w));
    } else if (n_leaves >= 2) {
        uint32_t stack[PB2_PSTACK];
        int sp = 0;
        uint32_t curr = 0;
        for (;;) {
            const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
            float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
            bool lh = overlaps(q, l0, l1), rh = overlaps(q, r0, r1);
            bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
            bool rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
            bool next = false;
            if (lh) {
                if (lleaf) emit(__float_as_uint(l0.w));
                else { curr = __float_as_uint(l0.w); next = true; }
            }
            if (rh) {
                if (rleaf) emit(__float_as_uint(r0.w));
                else if (next) pb2_push(stack, sp, __float_as_uint(r0.w), fault);
                else { curr = __float_as_uint(r0.w); next = true; }
            }
            if (!next) {
                if (sp == 0) break;
                curr = stack[--sp];
            }
        }
    }
    if (!WRITE) counts[qi] = cnt;
}

// ---------------------------------------------------------------- self pairs (single-tree BVTT)
// One thread per leaf position p. Emits (p, q) for q > p. CHANGE_DETECTION: pair kept iff either leaf is flagged
// changed (the reference's pruning, bvh_traverse_bvtt.rs:47,103-110,160-163, reduces to exactly this predicate
// because change flags are OR-ed up the tree).
// KARRAS = false (PLOC-linked tree): no range property, so the walk visits every overlapping subtree and keeps the pairs whose
// other leaf sits at a higher sorted position — still each unordered pair exactly once.
template <bool CD, bool KARRAS>
__global__ void __launch_bounds__(128) k_self_pairs(const NodeWide* __restrict__ nodes, const uint32_t* __restrict__ order,
                             const uint32_t* __restrict__ leaf_slot, uint32_t n_leaves, uint2* __restrict__ pairs,
                             uint64_t cap, unsigned long long* __restrict__ counter, unsigned int* fault, uint32_t shard, uint32_t n_shards) {
    uint64_t p64 = (uint64_t)(blockIdx.x * blockDim.x + threadIdx.x) * n_shards + shard;   // interleaved ownership of the walking leaves
    if (p64 >= n_leaves) return;
    uint32_t p = (uint32_t)p64;
    uint32_t my_id = order[p];
    uint32_t slot = leaf_slot[my_id];
    const float4* hp = reinterpret_cast<const float4*>((slot & 1u) ? &nodes[slot >> 1].right : &nodes[slot >> 1].left);
    float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
    Box q = {h0.x, h0.y, h0.z, h1.x, h1.y, h1.z};
    bool my_changed = (__float_as_uint(h1.w) >> 30) == 1u;

    auto emit = [&](uint32_t pos) {
        uint32_t other = order[pos];
        unsigned long long at = warp_append(counter);
        if (at < cap) pairs[at] = make_uint2(min(my_id, other), max(my_id, other));
    };

    uint32_t stack[PB2_PSTACK];
    int sp = 0;
    uint32_t curr = 0;
    for (;;) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
        uint32_t lc = __float_as_uint(l0.w), rc = __float_as_uint(r0.w);
        uint32_t ld = __float_as_uint(l1.w), rd = __float_as_uint(r1.w);
        bool lleaf = (ld & PB2_LEAF_COUNT_MASK) == 1u, rleaf = (rd & PB2_LEAF_COUNT_MASK) == 1u;
        // left child index == last leaf position of the left range: nothing > p in there when lc <= p.
        bool lh = (KARRAS ? lc > p : (!lleaf || lc > p)) && overlaps(q, l0, l1);
        // right child: leaf at position rc, or a range starting at rc (may extend past p).
        bool rh = (!rleaf || rc > p) && overlaps(q, r0, r1);
        if (CD) {
            lh = lh && (my_changed || (ld >> 30) == 1u);
            rh = rh && (my_changed || (rd >> 30) == 1u);
        }
        bool next = false;
        if (lh) {
            if (lleaf) emit(lc);
            else { curr = lc; next = true; }
        }
        if (rh) {
            if (rleaf) emit(rc);
            else if (next) pb2_push(stack, sp, rc, fault);
            else { curr = rc; next = true; }
        }
        if (!next) {
            if (sp == 0) break;
            curr = stack[--sp];
        }
    }
}

// ---------------------------------------------------------------- two-tree leaf pairs
// One thread per leaf of tree A (by position), walking tree B with check = intersects.
__global__ void __launch_bounds__(128) k_leaf_pairs(const NodeWide* __restrict__ nodes_a, const uint32_t* __restrict__ order_a,
                             const uint32_t* __restrict__ slot_a, uint32_t na, const NodeWide* __restrict__ nodes_b,
                             const uint32_t* __restrict__ order_b, uint32_t nb, uint2* __restrict__ pairs, uint64_t cap,
                             unsigned long long* __restrict__ counter, unsigned int* fault) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= na) return;
    uint32_t my_id = order_a[p];
    uint32_t slot = slot_a[my_id];
    const float4* hp = reinterpret_cast<const float4*>((slot & 1u) ? &nodes_a[slot >> 1].right : &nodes_a[slot >> 1].left);
    float4 h0 = __ldg(hp), h1 = __ldg(hp + 1);
    Box q = {h0.x, h0.y, h0.z, h1.x, h1.y, h1.z};
    auto emit = [&](uint32_t pos) {
        unsigned long long at = warp_append(counter);
        if (at < cap) pairs[at] = make_uint2(my_id, order_b[pos]);
    };
    // LeafPairs pushes the root-level child pairs without calling `check` (bvh_traverse_bvtt.rs:215-237): when both
    // trees are "tiny" (<= 2 leaves, all root halves are leaves) every combination is yielded unchecked.
    bool unchecked = na <= 2 && nb <= 2;
    if (nb == 1) {
        const float4* np = reinterpret_cast<const float4*>(&nodes_b[0]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1);
        if (unchecked || overlaps(q, l0, l1)) emit(__float_as_uint(l0.w));
        return;
    }
    uint32_t stack[PB2_PSTACK];
    int sp = 0;
    uint32_t curr = 0;
    for (;;) {
        const float4* np = reinterpret_cast<const float4*>(&nodes_b[curr]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
        bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        bool rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
        bool lh = unchecked || overlaps(q, l0, l1), rh = unchecked || overlaps(q, r0, r1);
        bool next = false;
        if (lh) {
            if (lleaf) emit(__float_as_uint(l0.w));
            else { curr = __float_as_uint(l0.w); next = true; }
        }
        if (rh) {
            if (rleaf) emit(__float_as_uint(r0.w));
            else if (next) pb2_push(stack, sp, __float_as_uint(r0.w), fault);
            else { curr = __float_as_uint(r0.w); next = true; }
        }
        if (!next) {
            if (sp == 0) break;
            curr = stack[--sp];
        }
    }
}

static int read_counter(pb2_ctx* ctx, uint64_t* count) {
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, ctx->d_counters, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CHECK(pb2_fetch_fault(ctx));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    *count = ctx->h_counters[0];
    return pb2_check_fault(ctx);
}

extern "C" {

// Device-resident CSR for internal callers (TriMesh-vs-shape contacts): d_offsets has m + 1 entries (caller-owned);
// *d_items is allocated here with cudaMallocAsync on ctx->stream (caller frees with cudaFreeAsync) and holds leaf ids,
// or sorted leaf positions when `positions` is set. Synchronises the stream once (to size the item list).
int pb2_intersect_csr_device(pb2_ctx* ctx, const pb2_bvh* bvh, const float* d_queries, uint32_t m, bool positions, uint32_t* d_offsets,
                             uint32_t** d_items, uint64_t* total_out) {
    cudaStream_t st = ctx->stream;
    *d_items = nullptr; *total_out = 0;
    if (m == 0) return PB2_OK;
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(m + 1), st);
    size_t counts_bytes = (((size_t)m + 1) * 4 + 255) & ~(size_t)255;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[2], counts_bytes + cub_bytes));
    uint32_t* counts = (uint32_t*)ctx->scratch[2].ptr;
    void* cub_tmp = (char*)ctx->scratch[2].ptr + counts_bytes;
    PB2_CUDA(ctx, cudaMemsetAsync(counts + m, 0, 4, st));
    k_intersect_aabbs<false><<<pb2_blocks(m, 128), 128, 0, st>>>(bvh->nodes, bvh->leaf_order, bvh->n_leaves, d_queries, m, counts, nullptr, nullptr, 0, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, d_offsets, (int)(m + 1), st));
    ctx->launches += 2;
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, d_offsets + m, 4, cudaMemcpyDeviceToHost, st));
    PB2_CHECK(pb2_fetch_fault(ctx));
    PB2_CUDA(ctx, cudaStreamSynchronize(st));
    PB2_CHECK(pb2_check_fault(ctx));
    uint64_t total = *(uint32_t*)ctx->h_counters;
    *total_out = total;
    if (total == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaMallocAsync((void**)d_items, total * 4, st));
    k_intersect_aabbs<true><<<pb2_blocks(m, 128), 128, 0, st>>>(bvh->nodes, positions ? nullptr : bvh->leaf_order, bvh->n_leaves, d_queries, m,
                                                               nullptr, d_offsets, *d_items, total, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    return PB2_OK;
}

int pb2_bvh_intersect_aabbs(pb2_ctx* ctx, const pb2_bvh* bvh, const float* queries, uint32_t m, uint32_t* offsets,
                            uint32_t* leaf_ids, uint64_t cap, uint64_t* count, int mem) {
    if (!ctx || !bvh || !count || (m && (!queries || !offsets))) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    *count = 0;
    cudaStream_t st = ctx->stream;
    const void* d_q = nullptr;
    void *d_off = nullptr, *d_ids = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, queries, (size_t)m * 24, mem, &d_q));
    PB2_CHECK(pb2_stage_out(ctx, 1, offsets, ((size_t)m + 1) * 4, mem, &d_off));
    PB2_CHECK(pb2_stage_out(ctx, 2, leaf_ids, (size_t)cap * 4, mem, &d_ids));
    if (!d_ids) cap = 0;
    // counts (m+1, last = 0) -> exclusive scan -> offsets
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)(m + 1), st);
    size_t counts_bytes = (((size_t)m + 1) * 4 + 255) & ~(size_t)255;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[2], counts_bytes + cub_bytes));
    uint32_t* counts = (uint32_t*)ctx->scratch[2].ptr;
    void* cub_tmp = (char*)ctx->scratch[2].ptr + counts_bytes;
    PB2_CUDA(ctx, cudaMemsetAsync(counts + m, 0, 4, st));
    if (m) {
        k_intersect_aabbs<false><<<pb2_blocks(m, 128), 128, 0, st>>>(bvh->nodes, bvh->leaf_order, bvh->n_leaves, (const float*)d_q, m,
                                                                    counts, nullptr, nullptr, 0, PB2_FAULT_PTR(ctx));
        PB2_LAUNCHED(ctx);
    }
    PB2_CUDA(ctx, cub::DeviceScan::ExclusiveSum(cub_tmp, cub_bytes, counts, (uint32_t*)d_off, (int)(m + 1), st));
    ctx->launches += 2;
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, (uint32_t*)d_off + m, 4, cudaMemcpyDeviceToHost, st));
    PB2_CHECK(pb2_fetch_fault(ctx));
    PB2_CUDA(ctx, cudaStreamSynchronize(st));
    PB2_CHECK(pb2_check_fault(ctx));
    uint64_t total = *(uint32_t*)ctx->h_counters;
    *count = total;
    if (m && cap) {
        k_intersect_aabbs<true><<<pb2_blocks(m, 128), 128, 0, st>>>(bvh->nodes, bvh->leaf_order, bvh->n_leaves, (const float*)d_q, m,
                                                                   nullptr, (const uint32_t*)d_off, (uint32_t*)d_ids, cap, PB2_FAULT_PTR(ctx));
        PB2_LAUNCHED(ctx);
        PB2_CUDA(ctx, cudaGetLastError());
    }
    PB2_CHECK(pb2_stage_back(ctx, offsets, d_off, ((size_t)m + 1) * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, leaf_ids, d_ids, (size_t)(total < cap ? total : cap) * 4, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    if (total > cap) PB2_FAIL(ctx, PB2_ERR_OVERFLOW, "intersect_aabbs: %llu hits > capacity %llu", (unsigned long long)total, (unsigned long long)cap);
    return PB2_OK;
}

int pb2_bvh_self_pairs_shard(pb2_ctx* ctx, const pb2_bvh* bvh, int change_detection, uint32_t shard, uint32_t n_shards, uint32_t* pairs,
                             uint64_t cap, uint64_t* count, int mem);
int pb2_bvh_self_pairs(pb2_ctx* ctx, const pb2_bvh* bvh, int change_detection, uint32_t* pairs, uint64_t cap, uint64_t* count, int mem) {
    return pb2_bvh_self_pairs_shard(ctx, bvh, change_detection, 0u, 1u, pairs, cap, count, mem);
}

int pb2_bvh_self_pairs_shard(pb2_ctx* ctx, const pb2_bvh* bvh, int change_detection, uint32_t shard, uint32_t n_shards, uint32_t* pairs,
                             uint64_t cap, uint64_t* count, int mem) {
    if (!ctx || !bvh || !count || n_shards == 0 || shard >= n_shards) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    *count = 0;
    // "Not enough nodes for any overlap" (bvh_traverse_bvtt.rs:24-27)
    if (bvh->n_leaves < 2) return PB2_OK;
    cudaStream_t st = ctx->stream;
    void* d_pairs = nullptr;
    PB2_CHECK(pb2_stage_out(ctx, 0, pairs, (size_t)cap * 8, mem, &d_pairs));
    if (!d_pairs) cap = 0;
    unsigned long long* counter = (unsigned long long*)ctx->d_counters;
    PB2_CUDA(ctx, cudaMemsetAsync(counter, 0, 8, st));
    unsigned blocks = pb2_blocks((bvh->n_leaves + n_shards - 1) / n_shards, 128);
    auto kern = change_detection ? (bvh->karras ? k_self_pairs<true, true> : k_self_pairs<true, false>)
                                 : (bvh->karras ? k_self_pairs<false, true> : k_self_pairs<false, false>);
    kern<<<blocks, 128, 0, st>>>(bvh->nodes, bvh->leaf_order, bvh->leaf_slot, bvh->n_leaves, (uint2*)d_pairs, cap, counter, PB2_FAULT_PTR(ctx),
                                 shard, n_shards);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(read_counter(ctx, count));
    uint64_t total = *count;
    PB2_CHECK(pb2_stage_back(ctx, pairs, d_pairs, (size_t)(total < cap ? total : cap) * 8, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    if (total > cap) PB2_FAIL(ctx, PB2_ERR_OVERFLOW, "self_pairs: %llu pairs > capacity %llu", (unsigned long long)total, (unsigned long long)cap);
    return PB2_OK;
}

int pb2_bvh_leaf_pairs(pb2_ctx* ctx, const pb2_bvh* a, const pb2_bvh* b, uint32_t* pairs, uint64_t cap, uint64_t* count, int mem) {
    if (!ctx || !a || !b || !count) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    *count = 0;
    if (a->n_leaves == 0 || b->n_leaves == 0) return PB2_OK;
    cudaStream_t st = ctx->stream;
    void* d_pairs = nullptr;
    PB2_CHECK(pb2_stage_out(ctx, 0, pairs, (size_t)cap * 8, mem, &d_pairs));
    if (!d_pairs) cap = 0;
    unsigned long long* counter = (unsigned long long*)ctx->d_counters;
    PB2_CUDA(ctx, cudaMemsetAsync(counter, 0, 8, st));
    k_leaf_pairs<<<pb2_blocks(a->n_leaves, 128), 128, 0, st>>>(a->nodes, a->leaf_order, a->leaf_slot, a->n_leaves, b->nodes, b->leaf_order,
                                                              b->n_leaves, (uint2*)d_pairs, cap, counter, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(read_counter(ctx, count));
    uint64_t total = *count;
    PB2_CHECK(pb2_stage_back(ctx, pairs, d_pairs, (size_t)(total < cap ? total : cap) * 8, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    if (total > cap) PB2_FAIL(ctx, PB2_ERR_OVERFLOW, "leaf_pairs: %llu pairs > capacity %llu", (unsigned long long)total, (unsigned long long)cap);
    return PB2_OK;
}

}  // extern "C"

// parry_b200 — GPU Bvh construction / refit.
//
// Replaces (reference, file:line): Bvh::from_leaves / from_iter (partitioning/bvh/bvh_tree.rs:1835,1891-1955),
// rebuild_range_binned (bvh_binned_build.rs:39-176), rebuild_range_ploc (bvh_ploc_build.rs:10-94),
// Bvh::refit (bvh_refit.rs:170-320), insert_or_update_partially (bvh_insert.rs:209-231).
//
// B200 design: the reference's builders are sequential recursions; here the tree is a Morton LBVH
// (63-bit codes, same 21-bit/axis quantisation as utils/morton.rs:12-40) sorted with a device radix sort and
// linked by Karras' parallel binary-radix-tree construction (one thread per internal node), then fitted
// bottom-up with per-node arrival counters (one thread per leaf). The node array uses the reference's own 64-byte
// BvhNodeWide layout so it can be handed back verbatim (pb2_bvh_download). Karras' numbering puts the two
// children of a node at adjacent indices (split, split+1) => sibling nodes share a 128-byte line.
#include "common.cuh"
#include "ploc.cuh"
#include <stdlib.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

// ---------------------------------------------------------------- helpers
__device__ __forceinline__ uint32_t f2ord(float f) {
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounds[0..2] = ordered min of centroids, bounds[3..5] = ordered max
__global__ void k_centroid_bounds(const float* __restrict__ aabbs, uint32_t n, uint32_t* __restrict__ bounds) {
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* a = aabbs + 6ull * i;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            float c = (a[d] + a[3 + d]) * 0.5f;
            mn[d] = fminf(mn[d], c);
            mx[d] = fmaxf(mx[d], c);
        }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        for (int o = 16; o > 0; o >>= 1) {
            mn[d] = fminf(mn[d], __shfl_xor_sync(0xffffffffu, mn[d], o));
            mx[d] = fmaxf(mx[d], __shfl_xor_sync(0xffffffffu, mx[d], o));
        }
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            atomicMin(&bounds[d], f2ord(mn[d]));
            atomicMax(&bounds[3 + d], f2ord(mx[d]));
        }
    }
}

__device__ __forceinline__ uint64_t split3(uint32_t a) {
    uint64_t x = (uint64_t)a & 0x1fffffull;
    x = (x | x << 32) & 0x1f00000000ffffull;
    x = (x | x << 16) & 0x1f0000ff0000ffull;
    x = (x | x << 8) & 0x100f00f00f00f00full;
    x = (x | x << 4) & 0x10c30c30c30c30c3ull;
    x = (x | x << 2) & 0x1249249249249249ull;
    return x;
}

__global__ void k_morton(const float* __restrict__ aabbs, uint32_t n, const uint32_t* __restrict__ bounds,
                         uint64_t* __restrict__ keys, uint32_t* __restrict__ vals, bool cubic) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = aabbs + 6ull * i;
    uint32_t q[3];
    // One scale for the three axes (the largest centroid extent): Morton cells are cubes, so a flat scene (terrain: 1000 x 40
    // x 1000) is split along its long axes first instead of being sliced into overlapping height bands. The reference's own
    // Morton helper normalises per axis (utils/morton.rs:33-40, PLOC only); query results do not depend on the tree.
    float emax = 0.0f;
#pragma unroll
    for (int d = 0; d < 3; ++d) emax = fmaxf(emax, ord2f(bounds[3 + d]) - ord2f(bounds[d]));
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float lo = ord2f(bounds[d]);
        float c = (a[d] + a[3 + d]) * 0.5f;
        float e = cubic ? emax : ord2f(bounds[3 + d]) - lo;
        float u = e > 0.0f ? (c - lo) / e : 0.0f;
        u = fminf(fmaxf(u, 0.0f), 1.0f);
        uint32_t v = (uint32_t)(u * 2097152.0f);
        q[d] = v > 2097151u ? 2097151u : v;
    }
    keys[i] = split3(q[0]) | (split3(q[1]) << 1) | (split3(q[2]) << 2);
    vals[i] = i;
}

__device__ __forceinline__ int lbvh_delta(const uint64_t* __restrict__ keys, int n, int i, uint64_t ki, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t kj = keys[j];
    if (ki == kj) return 64 + __clz(i ^ j);
    return __clzll((long long)(ki ^ kj));
}

// One thread per internal node (Karras 2012). Writes child links into the wide node, the parent links and
// the leaf slots. Leaf halves get children = sorted position, data = 1 | CHANGE_PENDING (BvhNode::leaf,
// bvh_tree.rs:520-527); boxes are filled by k_init_leaf_boxes.
__global__ void k_karras(const uint64_t* __restrict__ keys, const uint32_t* __restrict__ order, int n,
                         NodeWide* __restrict__ nodes, uint32_t* __restrict__ parents, uint32_t* __restrict__ leaf_slot) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    uint64_t ki = keys[i];
    int dl = lbvh_delta(keys, n, i, ki, i - 1), dr = lbvh_delta(keys, n, i, ki, i + 1);
    int d = dr > dl ? 1 : -1;
    int dmin = d > 0 ? dl : dr;
    int lmax = 2;
    while (lbvh_delta(keys, n, i, ki, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, ki, i + (l + t) * d) > dmin) l += t;
    int j = i + l * d;
    int dnode = lbvh_delta(keys, n, i, ki, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (lbvh_delta(keys, n, i, ki, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    int gamma = i + s * d + (d < 0 ? -1 : 0);
    int lo = min(i, j), hi = max(i, j);
    bool left_leaf = (lo == gamma), right_leaf = (hi == gamma + 1);
    NodeWide* w = &nodes[i];
    w->left.children = (uint32_t)gamma;
    w->left.data = left_leaf ? (1u | PB2_CHANGE_PENDING) : 0u;
    w->right.children = (uint32_t)(gamma + 1);
    w->right.data = right_leaf ? (1u | PB2_CHANGE_PENDING) : 0u;
    if (left_leaf) leaf_slot[order[gamma]] = ((uint32_t)i << 1);
    else parents[gamma] = ((uint32_t)i << 1);
    if (right_leaf) leaf_slot[order[gamma + 1]] = ((uint32_t)i << 1) | 1u;
    else parents[gamma + 1] = ((uint32_t)i << 1) | 1u;
    if (i == 0) parents[0] = 0;
}

// n <= 2 special cases (bvh_tree.rs:1914-1932): leaves stored in input order in the root wide node.
__global__ void k_tiny_tree(const float* __restrict__ aabbs, uint32_t n, NodeWide* nodes, uint32_t* parents,
                            uint32_t* leaf_slot, uint32_t* leaf_order) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    NodeWide w;
    memset(&w, 0, sizeof(w));
    for (uint32_t k = 0; k < n; ++k) {
        NodeHalf* h = k == 0 ? &w.left : &w.right;
        h->mnx = aabbs[6 * k + 0]; h->mny = aabbs[6 * k + 1]; h->mnz = aabbs[6 * k + 2];
        h->mxx = aabbs[6 * k + 3]; h->mxy = aabbs[6 * k + 4]; h->mxz = aabbs[6 * k + 5];
        h->children = k;
        h->data = 1u | PB2_CHANGE_PENDING;  // from_iter does not refit these (bvh_tree.rs:1914-1932)
        leaf_slot[k] = k;  // (0 << 1) | k
        leaf_order[k] = k;
    }
    nodes[0] = w;
    parents[0] = 0;
}

// One thread per leaf: copy its box into its slot.
// leaf_data != NULL (Bvh::rebuild): the leaf keeps the data word it had before the rebuild — the reference copies the leaf
// BvhNodes verbatim, change flags included (bvh_binned_build.rs:18-26); NULL (from_leaves): BvhNode::leaf = pending change.
__global__ void k_init_leaf_boxes(const float* __restrict__ aabbs, uint32_t n, const uint32_t* __restrict__ order,
                                  const uint32_t* __restrict__ leaf_slot, NodeWide* __restrict__ nodes,
                                  const uint32_t* __restrict__ leaf_data) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id = order[p];
    uint32_t slot = leaf_slot[id];
    NodeHalf* h = (slot & 1u) ? &nodes[slot >> 1].right : &nodes[slot >> 1].left;
    const float* a = aabbs + 6ull * id;
    float4 lo = make_float4(a[0], a[1], a[2], __uint_as_float(p));
    float4 hi = make_float4(a[3], a[4], a[5], __uint_as_float(leaf_data ? leaf_data[id] : (1u | PB2_CHANGE_PENDING)));
    reinterpret_cast<float4*>(h)[0] = lo;
    reinterpret_cast<float4*>(h)[1] = hi;
}

// insert_or_update_partially for existing leaves (bvh_insert.rs:209-231).
__global__ void k_update_leaves(const uint32_t* __restrict__ ids, const float* __restrict__ aabbs, uint32_t n, float margin,
                                const uint32_t* __restrict__ leaf_slot, uint32_t n_leaves, NodeWide* __restrict__ nodes) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t id = ids ? ids[k] : k;
    if (id >= n_leaves) return;
    uint32_t slot = leaf_slot[id];
    NodeHalf* h = (slot & 1u) ? &nodes[slot >> 1].right : &nodes[slot >> 1].left;
    const float* a = aabbs + 6ull * k;
    if (margin > 0.0f) {
        bool contains = h->mnx <= a[0] && h->mny <= a[1] && h->mnz <= a[2] && h->mxx >= a[3] && h->mxy >= a[4] && h->mxz >= a[5];
        if (!contains) {
            h->mnx = a[0] - margin; h->mny = a[1] - margin; h->mnz = a[2] - margin;
            h->mxx = a[3] + margin; h->mxy = a[4] + margin; h->mxz = a[5] + margin;
            h->data |= PB2_CHANGE_PENDING;
        }
    } else {
        h->mnx = a[0]; h->mny = a[1]; h->mnz = a[2];
        h->mxx = a[3]; h->mxy = a[4]; h->mxz = a[5];
    }
}

__device__ __forceinline__ uint32_t resolve_pending(uint32_t data) {
    // BvhNodeData::resolve_pending_change (bvh_tree.rs:192-198)
    if ((data >> 30) == 3u) return (data & PB2_LEAF_COUNT_MASK) | PB2_CHANGED;
    return data & PB2_LEAF_COUNT_MASK;
}

// Bottom-up refit (Bvh::refit, bvh_refit.rs:170-320, minus the DFS re-layout which is a CPU cache optimisation).
// One thread per leaf; the second thread to arrive at a wide node merges its two halves into the parent's slot
// (BvhNode::merged, bvh_tree.rs:610-618: inf/sup of the boxes, leaf counts added, change bits OR-ed).
// RESOLVE = false is the fitting pass of Bvh::rebuild: the builders merge the leaves' flags upward as they are
// (BvhNodeData::merged, bvh_tree.rs:200-204) and resolve nothing.
template <bool RESOLVE>
__global__ void k_refit(uint32_t n, const uint32_t* __restrict__ leaf_slot, const uint32_t* __restrict__ parents,
                        NodeWide* nodes, uint32_t* counters) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    uint32_t slot = leaf_slot[id];
    uint32_t node = slot >> 1;
    {
        NodeHalf* h = (slot & 1u) ? &nodes[node].right : &nodes[node].left;
        if (RESOLVE) h->data = resolve_pending(h->data);
    }
    for (;;) {
        __threadfence();
        uint32_t old = atomicAdd(&counters[node], 1u);
        if (old == 0u) return;
        // both halves are final: read them through L2 (written by other SMs)
        const float4* src = reinterpret_cast<const float4*>(&nodes[node]);
        float4 a0 = __ldcg(src + 0), a1 = __ldcg(src + 1), b0 = __ldcg(src + 2), b1 = __ldcg(src + 3);
        if (node == 0u) return;
        uint32_t da = __float_as_uint(a1.w), db = __float_as_uint(b1.w);
        uint32_t lc = (da & PB2_LEAF_COUNT_MASK) + (db & PB2_LEAF_COUNT_MASK);
        uint32_t ch = (da >> 30) | (db >> 30);
        float4 m0 = make_float4(fminf(a0.x, b0.x), fminf(a0.y, b0.y), fminf(a0.z, b0.z), __uint_as_float(node));
        float4 m1 = make_float4(fmaxf(a1.x, b1.x), fmaxf(a1.y, b1.y), fmaxf(a1.z, b1.z), __uint_as_float(lc | (ch << 30)));
        uint32_t ps = parents[node];
        float4* dst = reinterpret_cast<float4*>((ps & 1u) ? &nodes[ps >> 1].right : &nodes[ps >> 1].left);
        __stcg(dst + 0, m0);
        __stcg(dst + 1, m1);
        node = ps >> 1;
    }
}

// Translate the internal representation (leaf children = sorted position) to the reference's (leaf id).
__global__ void k_export_nodes(const NodeWide* __restrict__ nodes, uint32_t n_nodes, const uint32_t* __restrict__ order,
                               NodeWide* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_nodes) return;
    NodeWide w = nodes[i];
    if ((w.left.data & PB2_LEAF_COUNT_MASK) == 1u) w.left.children = order[w.left.children];
    if ((w.right.data & PB2_LEAF_COUNT_MASK) == 1u) w.right.children = order[w.right.children];
    out[i] = w;
}

// ---------------------------------------------------------------- host side
static int bvh_alloc(pb2_ctx* ctx, pb2_bvh* b, uint32_t n) {
    b->n_leaves = n;
    b->n_nodes = n == 0 ? 0 : (n <= 2 ? 1 : n - 1);
    b->cap_leaves = n;
    size_t nn = b->n_nodes ? b->n_nodes : 1, nl = n ? n : 1;
    PB2_CUDA(ctx, cudaMalloc((void**)&b->nodes, nn * sizeof(NodeWide)));
    PB2_CUDA(ctx, cudaMalloc((void**)&b->parents, nn * sizeof(uint32_t)));
    PB2_CUDA(ctx, cudaMalloc((void**)&b->counters, nn * sizeof(uint32_t)));
    PB2_CUDA(ctx, cudaMalloc((void**)&b->leaf_slot, nl * sizeof(uint32_t)));
    PB2_CUDA(ctx, cudaMalloc((void**)&b->leaf_order, nl * sizeof(uint32_t)));
    return PB2_OK;
}

static void bvh_free(pb2_bvh* b) {
    if (b->nodes) cudaFree(b->nodes);
    if (b->parents) cudaFree(b->parents);
    if (b->counters) cudaFree(b->counters);
    if (b->leaf_slot) cudaFree(b->leaf_slot);
    if (b->leaf_order) cudaFree(b->leaf_order);
    b->nodes = nullptr; b->parents = nullptr; b->counters = nullptr; b->leaf_slot = nullptr; b->leaf_order = nullptr;
}

static int bvh_refit_device(pb2_ctx* ctx, pb2_bvh* b, bool resolve = true) {
    if (b->n_leaves <= 2) {
        // refit_buffers <=2-leaf branch (bvh_refit.rs:192-201): only resolves the change flags.
        if (b->n_leaves == 0) return PB2_OK;
    }
    PB2_CUDA(ctx, cudaMemsetAsync(b->counters, 0, (size_t)b->n_nodes * sizeof(uint32_t), ctx->stream));
    if (b->n_leaves <= 2) {
        // a single wide node: resolving flags == running k_refit (each leaf resolves its own half, root returns)
    }
    if (resolve) k_refit<true><<<pb2_blocks(b->n_leaves, 256), 256, 0, ctx->stream>>>(b->n_leaves, b->leaf_slot, b->parents, b->nodes, b->counters);
    else k_refit<false><<<pb2_blocks(b->n_leaves, 256), 256, 0, ctx->stream>>>(b->n_leaves, b->leaf_slot, b->parents, b->nodes, b->counters);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    return PB2_OK;
}

// ---------------------------------------------------------------- PLOC (BvhBuildStrategy::Ploc)
// GPU form of rebuild_range_ploc (bvh_ploc_build.rs:10-94); the per-cluster rules live in ploc.cuh. Clusters are NodeHalf
// records ({mins, children}, {maxs, data}: exactly what the reference's `leaves: Vec<BvhNode>` holds), ping-ponged between
// two arrays; one round = nearest-neighbour kernel (shared-memory tile of the 2 x radius neighbourhood) -> decision flags ->
// one 64-bit inclusive scan (low word: position in the next round, high word: rank among this round's merges) -> emit. A
// round merges 35-45 % of the clusters on meshes and collider clouds, so a million leaves take ~30 rounds. Node ids are
// handed out from the top: the k-th node created gets id n - 2 - k, so the root (created last) is node 0 and every child
// follows its parent, like the reference's refit leaves them (bvh_refit.rs:259-320).
__global__ void k_ploc_init(const float* __restrict__ aabbs, uint32_t n, const uint32_t* __restrict__ order,
                            const uint32_t* __restrict__ leaf_data, float4* __restrict__ C) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    uint32_t id = order[p];
    const float* a = aabbs + 6ull * id;
    C[2ull * p] = make_float4(a[0], a[1], a[2], __uint_as_float(p));
    C[2ull * p + 1] = make_float4(a[3], a[4], a[5], __uint_as_float(leaf_data ? leaf_data[id] : (1u | PB2_CHANGE_PENDING)));
}

#define PLOC_BLOCK 128
__global__ void __launch_bounds__(PLOC_BLOCK) k_ploc_nearest(const float4* __restrict__ C, uint32_t c, uint32_t radius, uint32_t* __restrict__ cand) {
    __shared__ float4 tile[2 * (PLOC_BLOCK + 2 * PLOC_MAX_RADIUS)];
    uint32_t b0 = blockIdx.x * PLOC_BLOCK;
    uint32_t tile_lo = b0 >= radius ? b0 - radius : 0u;
    uint32_t tile_hi = b0 + PLOC_BLOCK + radius < c ? b0 + PLOC_BLOCK + radius : c;
    for (uint32_t t = threadIdx.x; t < 2u * (tile_hi - tile_lo); t += PLOC_BLOCK) tile[t] = C[2ull * tile_lo + t];
    __syncthreads();
    uint32_t i = b0 + threadIdx.x;
    if (i < c) cand[i] = ploc_nearest(tile, tile_lo, c, i, radius);
}

__global__ void k_ploc_flags(const uint32_t* __restrict__ cand, uint32_t c, unsigned long long* __restrict__ flags) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) flags[i] = ploc_flags(cand, i);
}

__global__ void k_ploc_emit(const float4* __restrict__ Cin, const uint32_t* __restrict__ cand, const unsigned long long* __restrict__ incl,
                            uint32_t c, float4* __restrict__ Cout, NodeWide* __restrict__ nodes, uint32_t* __restrict__ parents,
                            uint32_t* __restrict__ leaf_slot, const uint32_t* __restrict__ order, uint32_t created, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < c) ploc_emit(Cin, cand, incl, i, Cout, nodes, parents, leaf_slot, order, created, n);
}

// Links the sorted leaves (b->leaf_order) into b->nodes / parents / leaf_slot. PB2_ERR_UNSUPPORTED: the clustering made no
// progress (thousands of identical boxes merge one pair per round — the reference's own loop is quadratic there); the
// caller links the same sorted leaves as an LBVH instead.
static int bvh_link_ploc(pb2_ctx* ctx, pb2_bvh* b, const float* d_aabbs, uint32_t n, const uint32_t* leaf_data) {
    cudaStream_t st = ctx->stream;
    uint32_t radius = 16;   // SEARCH_RADIUS (bvh_ploc_build.rs:22)
    { const char* e = getenv("PB2_PLOC_RADIUS"); if (e && atoi(e) >= 1 && atoi(e) <= PLOC_MAX_RADIUS) radius = (uint32_t)atoi(e); }
    size_t cl_bytes = ((size_t)n * 32 + 255) & ~(size_t)255, cand_bytes = ((size_t)n * 4 + 255) & ~(size_t)255;
    size_t fl_bytes = ((size_t)n * 8 + 255) & ~(size_t)255, cub_bytes = 0;
    cub::DeviceScan::InclusiveSum(nullptr, cub_bytes, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int)n, st);
    char* base = nullptr;
    PB2_CUDA(ctx, cudaMallocAsync((void**)&base, 2 * cl_bytes + cand_bytes + 2 * fl_bytes + cub_bytes + 256, st));
    float4 *Ca = (float4*)base, *Cb = (float4*)(base + cl_bytes);
    uint32_t* cand = (uint32_t*)(base + 2 * cl_bytes);
    unsigned long long* flags = (unsigned long long*)(base + 2 * cl_bytes + cand_bytes);
    unsigned long long* incl = (unsigned long long*)(base + 2 * cl_bytes + cand_bytes + fl_bytes);
    void* cub_tmp = base + 2 * cl_bytes + cand_bytes + 2 * fl_bytes;
    int rc = [&]() -> int {
        k_ploc_init<<<pb2_blocks(n, 256), 256, 0, st>>>(d_aabbs, n, b->leaf_order, leaf_data, Ca);
        PB2_LAUNCHED(ctx);
        uint32_t c = n, created = 0;
        int rounds = 0, slow = 0;
        while (c > 1) {
            k_ploc_nearest<<<pb2_blocks(c, PLOC_BLOCK), PLOC_BLOCK, 0, st>>>(Ca, c, radius, cand);
            k_ploc_flags<<<pb2_blocks(c, 256), 256, 0, st>>>(cand, c, flags);
            PB2_CUDA(ctx, cub::DeviceScan::InclusiveSum(cub_tmp, cub_bytes, flags, incl, (int)c, st));
            k_ploc_emit<<<pb2_blocks(c, 256), 256, 0, st>>>(Ca, cand, incl, c, Cb, b->nodes, b->parents, b->leaf_slot, b->leaf_order, created, n);
            ctx->launches += 5;
            PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, incl + (c - 1), 8, cudaMemcpyDeviceToHost, st));
            PB2_CUDA(ctx, cudaStreamSynchronize(st));
            uint64_t tot = ctx->h_counters[0];
            uint32_t next = (uint32_t)tot, merges = (uint32_t)(tot >> 32);
            if (merges == 0 || next + merges != c) PB2_FAIL(ctx, PB2_ERR_CUDA, "PLOC round made no progress (%u clusters, %u merges)", c, merges);
            created += merges;
            c = next;
            float4* t = Ca; Ca = Cb; Cb = t;
            ++rounds;
            if (merges < c / 64u + 1u && c > 256u) { if (++slow > 32) return PB2_ERR_UNSUPPORTED; }
        }
        if (created != n - 1u) PB2_FAIL(ctx, PB2_ERR_CUDA, "PLOC created %u of %u nodes", created, n - 1u);
        (void)rounds;
        return PB2_OK;
    }();
    cudaFreeAsync(base, st);
    if (rc == PB2_OK) { PB2_CUDA(ctx, cudaGetLastError()); b->karras = false; }
    return rc;
}

// Builds topology + boxes from device-resident aabbs (n x 6).
// leaf_data (device, per leaf id) != NULL: rebuild of an existing tree, flags carried over and not resolved.
int pb2_bvh_build_device(pb2_ctx* ctx, pb2_bvh* b, const float* d_aabbs, uint32_t n, bool resolve_flags, const uint32_t* leaf_data) {
    cudaStream_t st = ctx->stream;
    if (n == 0) return PB2_OK;
    if (n <= 2) {
        k_tiny_tree<<<1, 32, 0, st>>>(d_aabbs, n, b->nodes, b->parents, b->leaf_slot, b->leaf_order);
        PB2_LAUNCHED(ctx);
        PB2_CUDA(ctx, cudaGetLastError());
        return PB2_OK;
    }
    // scratch: keys in/out (u64), vals in (u32), bounds (6 u32), cub temp
    size_t keys_bytes = (size_t)n * sizeof(uint64_t), vals_bytes = (size_t)n * sizeof(uint32_t);
    size_t cub_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint64_t*)nullptr, (uint64_t*)nullptr, (const uint32_t*)nullptr,
                                    (uint32_t*)nullptr, (int)n, 0, 63, st);
    size_t off_keys_in = 0, off_keys_out = off_keys_in + ((keys_bytes + 255) & ~(size_t)255);
    size_t off_vals_in = off_keys_out + ((keys_bytes + 255) & ~(size_t)255);
    size_t off_bounds = off_vals_in + ((vals_bytes + 255) & ~(size_t)255);
    size_t off_cub = off_bounds + 256;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[0], off_cub + cub_bytes));
    char* base = (char*)ctx->scratch[0].ptr;
    uint64_t* keys_in = (uint64_t*)(base + off_keys_in);
    uint64_t* keys_out = (uint64_t*)(base + off_keys_out);
    uint32_t* vals_in = (uint32_t*)(base + off_vals_in);
    uint32_t* bounds = (uint32_t*)(base + off_bounds);
    void* cub_tmp = base + off_cub;

    static const uint32_t init_bounds[6] = {0xffffffffu, 0xffffffffu, 0xffffffffu, 0u, 0u, 0u};
    PB2_CUDA(ctx, cudaMemcpyAsync(bounds, init_bounds, sizeof(init_bounds), cudaMemcpyHostToDevice, st));
    int grid = ctx->sm_count * 8;
    if ((uint64_t)grid * 256 > n) grid = (int)pb2_blocks(n, 256);
    k_centroid_bounds<<<grid, 256, 0, st>>>(d_aabbs, n, bounds);
    PB2_LAUNCHED(ctx);
    bool cubic = true;
    { const char* e = getenv("PB2_MORTON_PER_AXIS"); if (e && atoi(e)) cubic = false; }
    k_morton<<<pb2_blocks(n, 256), 256, 0, st>>>(d_aabbs, n, bounds, keys_in, vals_in, cubic);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(cub_tmp, cub_bytes, (const uint64_t*)keys_in, keys_out, (const uint32_t*)vals_in,
                                                  b->leaf_order, (int)n, 0, 63, st));
    ctx->launches += 4;  // histogram + onesweep passes (library kernels)
    if (b->strategy == PB2_BUILD_PLOC) {
        int s = bvh_link_ploc(ctx, b, d_aabbs, n, leaf_data);
        if (s == PB2_OK) return bvh_refit_device(ctx, b, resolve_flags);
        if (s != PB2_ERR_UNSUPPORTED) return s;   // UNSUPPORTED: the clustering stalled (degenerate input), link the sorted leaves as an LBVH
    }
    b->karras = true;
    k_karras<<<pb2_blocks(n - 1, 256), 256, 0, st>>>(keys_out, b->leaf_order, (int)n, b->nodes, b->parents, b->leaf_slot);
    PB2_LAUNCHED(ctx);
    k_init_leaf_boxes<<<pb2_blocks(n, 256), 256, 0, st>>>(d_aabbs, n, b->leaf_order, b->leaf_slot, b->nodes, leaf_data);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    return bvh_refit_device(ctx, b, resolve_flags);
}

// Stage a host array into device staging slot `slot`; returns device pointer (or the pointer itself in device mode).
int pb2_stage_in(pb2_ctx* ctx, int slot, const void* src, size_t bytes, int mem, const void** out) {
    if (mem == PB2_MEM_DEVICE || src == nullptr) { *out = src; return PB2_OK; }
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[slot], bytes ? bytes : 1));
    if (bytes) PB2_CUDA(ctx, cudaMemcpyAsync(ctx->stage[slot].ptr, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    *out = ctx->stage[slot].ptr;
    return PB2_OK;
}
// Reserve an output staging buffer (host mode) or pass the device pointer through.
int pb2_stage_out(pb2_ctx* ctx, int slot, void* dst, size_t bytes, int mem, void** out) {
    if (mem == PB2_MEM_DEVICE || dst == nullptr) { *out = dst; return PB2_OK; }
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[slot], bytes ? bytes : 1));
    *out = ctx->stage[slot].ptr;
    return PB2_OK;
}
int pb2_stage_back(pb2_ctx* ctx, void* dst, const void* dev, size_t bytes, int mem) {
    if (mem == PB2_MEM_DEVICE || dst == nullptr || bytes == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    return PB2_OK;
}

extern "C" {

int pb2_bvh_build(pb2_ctx* ctx, const float* aabbs, uint32_t n, int strategy, int mem, pb2_bvh** out) {
    if (!ctx || !out || (n && !aabbs)) return PB2_ERR_INVALID;
    if (n > PB2_LEAF_COUNT_MASK) PB2_FAIL(ctx, PB2_ERR_INVALID, "too many leaves");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    pb2_bvh* b = new pb2_bvh();
    b->strategy = strategy;
    int s = bvh_alloc(ctx, b, n);
    const void* d_aabbs = nullptr;
    if (s == PB2_OK) s = pb2_stage_in(ctx, 0, aabbs, (size_t)n * 24, mem, &d_aabbs);
    if (s == PB2_OK) s = pb2_bvh_build_device(ctx, b, (const float*)d_aabbs, n, true, nullptr);
    if (s == PB2_OK && mem == PB2_MEM_HOST) {
        cudaError_t e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "sync: %s", cudaGetErrorString(e)); s = PB2_ERR_CUDA; }
    }
    if (s != PB2_OK) { bvh_free(b); delete b; return s; }
    *out = b;
    return PB2_OK;
}

int pb2_bvh_destroy(pb2_ctx* ctx, pb2_bvh* bvh) {
    if (!ctx || !bvh) return PB2_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    bvh_free(bvh);
    delete bvh;
    return PB2_OK;
}

uint32_t pb2_bvh_leaf_count(const pb2_bvh* bvh) { return bvh ? bvh->n_leaves : 0; }
uint32_t pb2_bvh_node_count(const pb2_bvh* bvh) { return bvh ? bvh->n_nodes : 0; }

int pb2_bvh_update_leaves(pb2_ctx* ctx, pb2_bvh* bvh, const uint32_t* ids, const float* aabbs, uint32_t n, float margin, int mem) {
    if (!ctx || !bvh || (n && !aabbs)) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_ids = nullptr, *d_aabbs = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, aabbs, (size_t)n * 24, mem, &d_aabbs));
    PB2_CHECK(pb2_stage_in(ctx, 1, ids, (size_t)n * 4, mem, &d_ids));
    k_update_leaves<<<pb2_blocks(n, 256), 256, 0, ctx->stream>>>((const uint32_t*)d_ids, (const float*)d_aabbs, n, margin,
                                                                 bvh->leaf_slot, bvh->n_leaves, bvh->nodes);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

int pb2_bvh_refit(pb2_ctx* ctx, pb2_bvh* bvh) {
    if (!ctx || !bvh) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    return bvh_refit_device(ctx, bvh);
}

__global__ void k_gather_leaf_aabbs(const NodeWide* __restrict__ nodes, const uint32_t* __restrict__ leaf_slot, uint32_t n,
                                    float* __restrict__ aabbs, uint32_t* __restrict__ leaf_data) {
    uint32_t id = blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n) return;
    uint32_t slot = leaf_slot[id];
    const NodeHalf* h = (slot & 1u) ? &nodes[slot >> 1].right : &nodes[slot >> 1].left;
    float* a = aabbs + 6ull * id;
    a[0] = h->mnx; a[1] = h->mny; a[2] = h->mnz; a[3] = h->mxx; a[4] = h->mxy; a[5] = h->mxz;
    if (leaf_data) leaf_data[id] = h->data;
}

int pb2_bvh_rebuild(pb2_ctx* ctx, pb2_bvh* bvh, int strategy) {
    if (!ctx || !bvh) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    bvh->strategy = strategy;
    uint32_t n = bvh->n_leaves;
    if (n < 3) return PB2_OK;  // bvh_binned_build.rs:12-16: nothing to rebuild
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[1], (size_t)n * 28));
    float* aabbs = (float*)ctx->scratch[1].ptr;
    uint32_t* leaf_data = (uint32_t*)(aabbs + 6ull * n);
    k_gather_leaf_aabbs<<<pb2_blocks(n, 256), 256, 0, ctx->stream>>>(bvh->nodes, bvh->leaf_slot, n, aabbs, leaf_data);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    // the leaves keep their change flags and nothing is resolved (bvh_binned_build.rs:11-36): change detection still sees
    // exactly the leaves flagged by the last refit
    return pb2_bvh_build_device(ctx, bvh, aabbs, n, false, leaf_data);
}

// Leaves that are not (or no longer) part of the tree keep a slot with Aabb::new_invalid() (mins = +MAX, maxs = -MAX):
// the neutral element of the box merge, never overlapped, never hit by a ray — inert in every query.
__global__ void k_fill_invalid_aabbs(float* __restrict__ aabbs, uint32_t lo, uint32_t hi, uint32_t* __restrict__ leaf_data) {
    uint32_t i = lo + blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= hi) return;
    float* a = aabbs + 6ull * i;
    a[0] = a[1] = a[2] = FLT_MAX;
    a[3] = a[4] = a[5] = -FLT_MAX;
    leaf_data[i] = 1u | PB2_CHANGE_PENDING;   // a leaf that did not exist before: BvhNode::leaf (bvh_tree.rs:520-527)
}
__global__ void k_invalidate_leaves(const uint32_t* __restrict__ ids, uint32_t n, const uint32_t* __restrict__ leaf_slot, uint32_t n_leaves,
                                    NodeWide* __restrict__ nodes) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t id = ids[k];
    if (id >= n_leaves) return;
    uint32_t slot = leaf_slot[id];
    NodeHalf* h = (slot & 1u) ? &nodes[slot >> 1].right : &nodes[slot >> 1].left;
    h->mnx = h->mny = h->mnz = FLT_MAX;
    h->mxx = h->mxy = h->mxz = -FLT_MAX;
}

// Bvh::remove (bvh_tree.rs:2360-2427), batched: the leaves become inert and every ancestor box is re-fitted bottom-up
// (the reference splices the sibling into the parent and refits the ancestors; query results are the same).
int pb2_bvh_remove_leaves(pb2_ctx* ctx, pb2_bvh* bvh, const uint32_t* ids, uint32_t n, int mem) {
    if (!ctx || !bvh || (n && !ids)) return PB2_ERR_INVALID;
    if (n == 0 || bvh->n_leaves == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void* d_ids = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 1, ids, (size_t)n * 4, mem, &d_ids));
    k_invalidate_leaves<<<pb2_blocks(n, 256), 256, 0, ctx->stream>>>((const uint32_t*)d_ids, n, bvh->leaf_slot, bvh->n_leaves, bvh->nodes);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(bvh_refit_device(ctx, bvh));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

// Grows the leaf-id space to new_n (Bvh::insert of a new leaf index, bvh_insert.rs:126-197, is resize + update + rebuild:
// structural edits are whole-tree rebuilds here, ~1 ms per million leaves, instead of the reference's SAH descent with
// rotations). New ids start inert; existing leaves keep their boxes.
int pb2_bvh_resize(pb2_ctx* ctx, pb2_bvh* bvh, uint32_t new_n) {
    if (!ctx || !bvh) return PB2_ERR_INVALID;
    if (new_n > PB2_LEAF_COUNT_MASK) PB2_FAIL(ctx, PB2_ERR_INVALID, "too many leaves");
    if (new_n < bvh->n_leaves) PB2_FAIL(ctx, PB2_ERR_INVALID, "pb2_bvh_resize cannot shrink (remove leaves instead)");
    if (new_n == bvh->n_leaves) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t old_n = bvh->n_leaves;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[1], (size_t)new_n * 28));
    float* aabbs = (float*)ctx->scratch[1].ptr;
    uint32_t* leaf_data = (uint32_t*)(aabbs + 6ull * new_n);
    if (old_n) {
        k_gather_leaf_aabbs<<<pb2_blocks(old_n, 256), 256, 0, ctx->stream>>>(bvh->nodes, bvh->leaf_slot, old_n, aabbs, leaf_data);
        PB2_LAUNCHED(ctx);
    }
    k_fill_invalid_aabbs<<<pb2_blocks(new_n - old_n, 256), 256, 0, ctx->stream>>>(aabbs, old_n, new_n, leaf_data);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    bvh_free(bvh);
    PB2_CHECK(bvh_alloc(ctx, bvh, new_n));
    // existing leaves keep their flags; the new (still inert) ones are pending until the next refit resolves them
    return pb2_bvh_build_device(ctx, bvh, aabbs, new_n, false, leaf_data);
}

int pb2_bvh_download(pb2_ctx* ctx, const pb2_bvh* bvh, void* nodes64, uint32_t* parents, uint32_t* leaf_node_indices, int mem) {
    if (!ctx || !bvh) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bvh->n_nodes == 0) return PB2_OK;
    cudaMemcpyKind kind = mem == PB2_MEM_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice;
    if (nodes64) {
        void* d_out = nullptr;
        PB2_CHECK(pb2_stage_out(ctx, 0, nodes64, (size_t)bvh->n_nodes * 64, mem, &d_out));
        k_export_nodes<<<pb2_blocks(bvh->n_nodes, 256), 256, 0, ctx->stream>>>(bvh->nodes, bvh->n_nodes, bvh->leaf_order, (NodeWide*)d_out);
        PB2_LAUNCHED(ctx);
        PB2_CUDA(ctx, cudaGetLastError());
        PB2_CHECK(pb2_stage_back(ctx, nodes64, d_out, (size_t)bvh->n_nodes * 64, mem));
    }
    if (parents) PB2_CUDA(ctx, cudaMemcpyAsync(parents, bvh->parents, (size_t)bvh->n_nodes * 4, kind, ctx->stream));
    if (leaf_node_indices) PB2_CUDA(ctx, cudaMemcpyAsync(leaf_node_indices, bvh->leaf_slot, (size_t)bvh->n_leaves * 4, kind, ctx->stream));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

int pb2_bvh_root_aabb(pb2_ctx* ctx, const pb2_bvh* bvh, float* aabb6) {
    if (!ctx || !bvh || !aabb6) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (bvh->n_leaves == 0) {  // Aabb::new_invalid (bvh_tree.rs:1991-1999)
        for (int d = 0; d < 3; ++d) { aabb6[d] = FLT_MAX; aabb6[3 + d] = -FLT_MAX; }
        return PB2_OK;
    }
    NodeWide w;
    PB2_CUDA(ctx, cudaMemcpyAsync(&w, bvh->nodes, sizeof(w), cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (bvh->n_leaves == 1) {
        aabb6[0] = w.left.mnx; aabb6[1] = w.left.mny; aabb6[2] = w.left.mnz; aabb6[3] = w.left.mxx; aabb6[4] = w.left.mxy; aabb6[5] = w.left.mxz;
    } else {
        aabb6[0] = fminf(w.left.mnx, w.right.mnx); aabb6[1] = fminf(w.left.mny, w.right.mny); aabb6[2] = fminf(w.left.mnz, w.right.mnz);
        aabb6[3] = fmaxf(w.left.mxx, w.right.mxx); aabb6[4] = fmaxf(w.left.mxy, w.right.mxy); aabb6[5] = fmaxf(w.left.mxz, w.right.mxz);
    }
    return PB2_OK;
}

}  // extern "C"

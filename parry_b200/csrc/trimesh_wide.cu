// parry_b200 — compressed 8-wide traversal tree for TriMesh ray casts.
//
// Same query as raycast.cu (TriMesh::cast_ray -> Bvh::cast_ray -> Bvh::find_best, query/ray/ray_trimesh.rs:8-36,
// partitioning/bvh/bvh_queries.rs:260-271, bvh_traverse.rs:335-417), different memory shape. The ncu captures of the
// binary-tree kernels (profiles/r1_rays_v*_full.json) show the L1/TEX pipe at 70-82 % with incoherent rays: every lane
// touches its own 64-byte node, four 16-byte loads per two child boxes. Here every binary subtree of up to eight
// children is collapsed into one 80-byte node (five 16-byte loads per eight child boxes) whose boxes are quantised to
// 8 bits per plane relative to the node's own origin and per-axis power-of-two scale (the compressed wide BVH of
// Ylitie, Karras & Laine, HPG 2017, restated from the paper; octant-ordered child slots give a front-to-back order
// without sorting).
//
// Parity: the quantised boxes only decide WHICH triangles get looked at. They are conservative by construction —
// quantised planes are rounded outward at build time (verified in double), the dequantised slab bounds are evaluated
// with directed rounding (lower bound for entries, upper bound for exits) and the interval test carries a relative
// slack of 2^-20 that covers the reference's own two f32 roundings — so every leaf whose exact AABB passes the
// reference's slab test is reached. Each reached triangle is then filtered by the reference's arithmetic: exact leaf
// AABB (Triangle::local_aabb, aabb_triangle.rs:16-30, recomputed from the vertices) slab-tested like
// BvhNode::cast_ray (bvh_tree.rs:1177-1181), then local_ray_intersection_with_triangle (ray_triangle.rs:70-152).
// Accept / tie rules are those of raycast.cu.
#include "trimesh.cuh"
#include <string.h>
#include <cub/device/device_scan.cuh>

#define W8_LEAF 0x80000000u
// Record sizes in float4. Round 1 kept 80-byte nodes read with five LDG.128 and 48-byte triangles read with three. A node-fetch
// microbenchmark (profiles/r2_ldg_bench.txt: one random node per lane, 28 warps per SM) shows that what such a
// fetch costs is the number of load instructions, not the bytes: L1-resident 88 G nodes/s (5 x LDG.128 on 80 B) against 161 G
// (3 x LDG.E.256 on 96 B), L2-resident 60 against 97, HBM 15 against 15. So both records are padded to a multiple of 32 bytes,
// 32-byte aligned, and read with 256-bit loads (sm_100+): 3 per node, 2 per triangle. The aligned records also touch fewer
// 32-byte sectors than before (a 16-byte-aligned 80-byte node straddles 3.5 sectors on average, a 48-byte triangle 2.5).
#define W8_NODE_F4 6
#define W8_TRI_F4 4
struct F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const float4* p) {   // p 32-byte aligned; read-only path
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w) : "l"(p));
    return r;
}
#define W8_STACK 96        // per-lane traversal stack entries; a visit pushes at most 7, so depth <= 7 * levels
#define W8_MAX_LEVELS 13        // k_raycast_wide<., MODE 1> only (it pushes up to 7 entries per level)
#define W8_MAX_LEVELS_GROUPS 64 // group walk (MODE 0 and k_raycast_wide_shared): one stack entry per level
#define W8_TQ 8   // per-lane queue of triangle groups (power of two)
#define W8_GAMMA 1.00000095367431640625f  // 1 + 2^-20

// ------------------------------------------------------------------------------------------------ build
// Collapses the binary subtree rooted at wroot[w] into wide node w (one thread).
__device__ __forceinline__ void collapse8_node(const NodeWide* __restrict__ nodes, uint32_t* __restrict__ wroot, uint32_t w,
                            uint32_t* __restrict__ total, float4* __restrict__ nodes8, uint32_t* __restrict__ leafpos,
                            uint32_t* __restrict__ nleaf) {
    uint32_t ref[8], lc[8];  // lc = leaves below the child; bit 31 = "being flattened" (see below)
    float lo[8][3], hi[8][3];
    int cnt = 0;
    auto add_half = [&](int at, float4 a, float4 b, uint32_t flag) {
        uint32_t n = __float_as_uint(b.w) & PB2_LEAF_COUNT_MASK;
        ref[at] = n == 1u ? (__float_as_uint(a.w) | W8_LEAF) : __float_as_uint(a.w);
        lc[at] = n | flag;
        lo[at][0] = a.x; lo[at][1] = a.y; lo[at][2] = a.z;
        hi[at][0] = b.x; hi[at][1] = b.y; hi[at][2] = b.z;
    };
    {
        const float4* np = reinterpret_cast<const float4*>(&nodes[wroot[w]]);
        add_half(0, np[0], np[1], 0u);
        add_half(1, np[2], np[3], 0u);
        cnt = 2;
    }
    // Greedy collapse. Subtrees with more than eight leaves are opened largest-surface-area first. A subtree with at
    // most eight leaves is either flattened completely into this node (when all its leaves fit into the free slots) or
    // left whole, so that it becomes one full leaf node of its own: opening it half-way would leave a litter of
    // two- and three-triangle nodes at the bottom of the tree (the first version did: 3.1 triangles per node).
    while (cnt < 8) {
        int k = -1;
        float best_key = -1.0f;
        for (int c = 0; c < cnt; ++c) {  // a subtree already being flattened goes first
            if (!(ref[c] & W8_LEAF) && (lc[c] & 0x80000000u)) { k = c; break; }
        }
        if (k < 0) {
            for (int c = 0; c < cnt; ++c) {
                if ((ref[c] & W8_LEAF) || lc[c] <= 8u) continue;
                float ex = hi[c][0] - lo[c][0], ey = hi[c][1] - lo[c][1], ez = hi[c][2] - lo[c][2];
                float area = ex * ey + ey * ez + ez * ex;
                if (area > best_key) { best_key = area; k = c; }
            }
        }
        uint32_t flag = 0u;
        if (k < 0) {
            for (int c = 0; c < cnt; ++c) {  // the largest small subtree whose leaves all fit
                if ((ref[c] & W8_LEAF) || cnt - 1 + (int)lc[c] > 8) continue;
                if ((float)lc[c] > best_key) { best_key = (float)lc[c]; k = c; }
            }
            flag = 0x80000000u;
        } else if (lc[k] & 0x80000000u) {
            flag = 0x80000000u;
        }
        if (k < 0) break;
        const float4* np = reinterpret_cast<const float4*>(&nodes[ref[k]]);
        float4 a0 = np[0], a1 = np[1], b0 = np[2], b1 = np[3];
        add_half(k, a0, a1, flag);
        add_half(cnt, b0, b1, flag);
        cnt++;
    }
    // node frame
    double plo[3], phi[3];
    for (int a = 0; a < 3; ++a) {
        float mn = lo[0][a], mx = hi[0][a];
        for (int c = 1; c < cnt; ++c) { mn = fminf(mn, lo[c][a]); mx = fmaxf(mx, hi[c][a]); }
        plo[a] = (double)mn; phi[a] = (double)mx;
    }
    // octant slots: child c goes to the free slot s whose sign vector (bit a of s set = +axis a) best matches the
    // direction from the node centre to the child centre (greedy on the 8x8 cost table)
    int slot_of[8];
    {
        float cost[8][8];
        for (int c = 0; c < cnt; ++c) {
            float v[3];
            for (int a = 0; a < 3; ++a) v[a] = (float)(((double)lo[c][a] + (double)hi[c][a]) * 0.5 - (plo[a] + phi[a]) * 0.5);
            for (int s = 0; s < 8; ++s)
                cost[c][s] = ((s & 1) ? v[0] : -v[0]) + ((s & 2) ? v[1] : -v[1]) + ((s & 4) ? v[2] : -v[2]);
        }
        uint32_t used_c = 0, used_s = 0;
        for (int it = 0; it < cnt; ++it) {
            int bc = -1, bs = -1;
            float bv = -FLT_MAX;
            for (int c = 0; c < cnt; ++c) {
                if (used_c & (1u << c)) continue;
                for (int s = 0; s < 8; ++s) {
                    if (used_s & (1u << s)) continue;
                    if (bc < 0 || cost[c][s] > bv) { bv = cost[c][s]; bc = c; bs = s; }
                }
            }
            slot_of[bc] = bs;
            used_c |= 1u << bc;
            used_s |= 1u << bs;
        }
    }
    // per-axis power-of-two scale and outward-rounded 8-bit planes; every plane is verified in double
    uint32_t qlo[3][8], qhi[3][8], ebits[3];
    for (int a = 0; a < 3; ++a) {
        for (int s = 0; s < 8; ++s) { qlo[a][s] = 255u; qhi[a][s] = 0u; }  // empty slots never pass (valid masks also say so)
        double ext = phi[a] - plo[a];
        int e = -141;
        if (ext > 0.0) { frexp(ext / 255.0, &e); }
        if (e < -141) e = -141;
        if (e > 112) e = 112;
        for (;;) {
            double S = ldexp(1.0, e);
            bool ok = true;
            for (int c = 0; c < cnt; ++c) {
                double l = (double)lo[c][a], h = (double)hi[c][a];
                double ql = floor((l - plo[a]) / S), qh = ceil((h - plo[a]) / S);
                if (ql < 0.0) ql = 0.0;
                if (ql > 255.0) ql = 255.0;
                while (ql > 0.0 && plo[a] + ql * S > l) ql -= 1.0;
                if (qh < 0.0) qh = 0.0;
                while (qh <= 255.0 && plo[a] + qh * S < h) qh += 1.0;
                if (qh > 255.0) { ok = false; break; }
                qlo[a][slot_of[c]] = (uint32_t)ql;
                qhi[a][slot_of[c]] = (uint32_t)qh;
            }
            if (ok || e >= 112) break;
            e++;
        }
        ebits[a] = (uint32_t)(e + 15 + 127);  // biased exponent of 2^(e+15): the kernel scales (1 + q * 2^-15)
    }
    uint32_t imask = 0, lmask = 0;
    uint32_t cref[8];
    for (int c = 0; c < cnt; ++c) {
        int s = slot_of[c];
        cref[s] = ref[c];
        if (ref[c] & W8_LEAF) lmask |= 1u << s;
        else imask |= 1u << s;
    }
    uint32_t base = 0;
    if (imask) base = atomicAdd(total, (uint32_t)__popc(imask));
    for (int s = 0; s < 8; ++s) {
        if (imask & (1u << s)) wroot[base + __popc(imask & ((1u << s) - 1u))] = cref[s];
        leafpos[8ull * w + s] = (lmask & (1u << s)) ? (cref[s] & ~W8_LEAF) : 0xffffffffu;
    }
    nleaf[w] = (uint32_t)__popc(lmask);
    auto pack4 = [](const uint32_t* q) { return q[0] | (q[1] << 8) | (q[2] << 16) | (q[3] << 24); };
    float4* out = nodes8 + (size_t)W8_NODE_F4 * w;
    out[0] = make_float4((float)plo[0], (float)plo[1], (float)plo[2],
                         __uint_as_float(ebits[0] | (ebits[1] << 8) | (ebits[2] << 16) | (imask << 24)));
    // .w = 1.0f: the kernel ORs the quantised bytes into this word's mantissa; reading it from the node (instead of an
    // immediate) leaves PRMT's only flexible operand slot to the byte selector
    out[1] = make_float4(__uint_as_float(base), __uint_as_float(0u), __uint_as_float(lmask), 1.0f);
    out[2] = make_float4(__uint_as_float(pack4(&qlo[0][0])), __uint_as_float(pack4(&qlo[0][4])),
                         __uint_as_float(pack4(&qlo[1][0])), __uint_as_float(pack4(&qlo[1][4])));
    out[3] = make_float4(__uint_as_float(pack4(&qlo[2][0])), __uint_as_float(pack4(&qlo[2][4])),
                         __uint_as_float(pack4(&qhi[0][0])), __uint_as_float(pack4(&qhi[0][4])));
    out[4] = make_float4(__uint_as_float(pack4(&qhi[1][0])), __uint_as_float(pack4(&qhi[1][4])),
                         __uint_as_float(pack4(&qhi[2][0])), __uint_as_float(pack4(&qhi[2][4])));
    out[5] = make_float4(0.f, 0.f, 0.f, 0.f);   // padding to 96 bytes (never read by the kernels' arithmetic)
}

// One thread per wide node of the current BFS level. wroot[w] = binary node collapsed into wide node w.
// The level loop runs without the host (round 1 read the node count back and synchronised after every level): launch `level` of a
// fixed series works on the wide nodes [bounds[level], *total) — *total is final when the launch starts because the previous launch
// has completed — and leaves its own end in bounds[level + 1] for the next one. Launches past the last level find an empty range.
__global__ void k_collapse8(const NodeWide* __restrict__ nodes, uint32_t* __restrict__ wroot, uint32_t* __restrict__ bounds, uint32_t level,
                            uint32_t* __restrict__ total, float4* __restrict__ nodes8, uint32_t* __restrict__ leafpos,
                            uint32_t* __restrict__ nleaf) {
    const uint32_t w_begin = bounds[level], w_end = bounds[W8_MAX_LEVELS_GROUPS + 2 + level];   // end: snapshot taken by k_collapse8_level
    for (uint32_t w = w_begin + blockIdx.x * blockDim.x + threadIdx.x; w < w_end; w += gridDim.x * blockDim.x)
        collapse8_node(nodes, wroot, w, total, nodes8, leafpos, nleaf);
}
// Snapshot of the level's range (one thread, between two collapse launches): bounds[level] = begin, bounds[.. + level] = end = *total.
__global__ void k_collapse8_level(uint32_t* __restrict__ bounds, uint32_t level, const uint32_t* __restrict__ total) {
    uint32_t begin = level == 0 ? 0u : bounds[W8_MAX_LEVELS_GROUPS + 2 + level - 1];
    uint32_t end = *total;
    bounds[level] = begin;
    bounds[W8_MAX_LEVELS_GROUPS + 2 + level] = end;
    if (begin < end) bounds[W8_MAX_LEVELS_GROUPS + 1] = level + 1;   // number of non-empty levels so far
}

// Writes each node's first-triangle index and gathers the triangles into (wide node, slot) order.
__global__ void k_finalize8(uint32_t n8, const uint32_t* __restrict__ tri_base, const uint32_t* __restrict__ leafpos,
                            const float4* __restrict__ tris, float4* __restrict__ nodes8, float4* __restrict__ tris8) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n8) return;
    uint32_t tb = tri_base[w];
    reinterpret_cast<uint32_t*>(nodes8 + (size_t)W8_NODE_F4 * w + 1)[1] = tb;
    uint32_t r = 0;
    for (int s = 0; s < 8; ++s) {
        uint32_t pos = leafpos[8ull * w + s];
        if (pos == 0xffffffffu) continue;
        const float4* src = tris + 3ull * pos;
        float4* dst = tris8 + (size_t)W8_TRI_F4 * (tb + r);
        dst[0] = src[0]; dst[1] = src[1]; dst[2] = src[2]; dst[3] = make_float4(0.f, 0.f, 0.f, 0.f);
        r++;
    }
}

int pb2_wide_build(pb2_ctx* ctx, pb2_trimesh* mesh) {
    pb2_bvh* b = &mesh->bvh;
    mesh->n_nodes8 = 0;
    if (b->n_leaves < 3) return PB2_OK;
    const char* off = getenv("PB2_NO_WIDE");
    if (off && atoi(off)) return PB2_OK;
    cudaStream_t st = ctx->stream;
    uint32_t nn = b->n_nodes;
    // One stream-ordered allocation for every temporary (round 1: six cudaMalloc / cudaFree pairs), no host round trip per level:
    // a fixed series of W8_MAX_LEVELS_GROUPS launch pairs whose ranges live on the device (see k_collapse8), then ONE
    // synchronisation to learn the node count the final arrays are sized with.
    size_t cub_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)nn, st);
    auto up = [](size_t x) { return (x + 255) & ~(size_t)255; };
    const size_t n_bounds = 2 * (W8_MAX_LEVELS_GROUPS + 2);
    size_t o_wroot = 0, o_leafpos = o_wroot + up((size_t)nn * 4), o_nleaf = o_leafpos + up((size_t)nn * 32), o_tbase = o_nleaf + up((size_t)nn * 4);
    size_t o_bounds = o_tbase + up((size_t)nn * 4), o_total = o_bounds + up(n_bounds * 4), o_tmp8 = o_total + 256;
    size_t o_cub = o_tmp8 + up((size_t)nn * 16 * W8_NODE_F4), bytes = o_cub + up(cub_bytes ? cub_bytes : 1);
    char* base = nullptr;
    PB2_CUDA(ctx, cudaMallocAsync((void**)&base, bytes, st));
    uint32_t *wroot = (uint32_t*)(base + o_wroot), *leafpos = (uint32_t*)(base + o_leafpos), *nleaf = (uint32_t*)(base + o_nleaf);
    uint32_t *tbase = (uint32_t*)(base + o_tbase), *bounds = (uint32_t*)(base + o_bounds), *total = (uint32_t*)(base + o_total);
    float4* tmp8 = (float4*)(base + o_tmp8);
    int s = PB2_OK;
    auto fail = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && s == PB2_OK) { snprintf(ctx->err, sizeof(ctx->err), "wide build (%s): %s", what, cudaGetErrorString(e)); s = PB2_ERR_CUDA; }
        return e != cudaSuccess;
    };
    do {
        uint32_t one = 1, zero = 0;
        if (fail(cudaMemsetAsync(bounds, 0, n_bounds * 4, st), "init")) break;
        if (fail(cudaMemcpyAsync(total, &one, 4, cudaMemcpyHostToDevice, st), "init")) break;
        if (fail(cudaMemcpyAsync(wroot, &zero, 4, cudaMemcpyHostToDevice, st), "init")) break;
        unsigned grid = pb2_blocks(nn, 128);
        unsigned cap = (unsigned)ctx->sm_count * 16u;
        if (grid > cap) grid = cap;
        for (uint32_t level = 0; level < W8_MAX_LEVELS_GROUPS; ++level) {
            k_collapse8_level<<<1, 1, 0, st>>>(bounds, level, total);
            // the first levels hold 1, 8, 64 ... nodes: a small grid until the range can be large (8^level nodes at most)
            unsigned g = level < 8 ? (unsigned)min((unsigned long long)grid, ((1ull << (3 * level)) + 127ull) / 128ull) : grid;
            k_collapse8<<<g ? g : 1, 128, 0, st>>>(b->nodes, wroot, bounds, level, total, tmp8, leafpos, nleaf);
            ctx->launches += 2;
        }
        uint32_t h[2] = {0, 0};   // {levels, node count}
        if (fail(cudaMemcpyAsync(&h[0], bounds + W8_MAX_LEVELS_GROUPS + 1, 4, cudaMemcpyDeviceToHost, st), "levels")) break;
        if (fail(cudaMemcpyAsync(&h[1], total, 4, cudaMemcpyDeviceToHost, st), "levels")) break;
        if (fail(cudaStreamSynchronize(st), "levels")) break;
        // did the series end on an empty level? (the last launch pair must have found nothing to do)
        uint32_t last[2] = {0, 0};
        if (fail(cudaMemcpyAsync(&last[0], bounds + W8_MAX_LEVELS_GROUPS - 1, 4, cudaMemcpyDeviceToHost, st), "levels")) break;
        if (fail(cudaMemcpyAsync(&last[1], bounds + W8_MAX_LEVELS_GROUPS + 2 + W8_MAX_LEVELS_GROUPS - 1, 4, cudaMemcpyDeviceToHost, st), "levels")) break;
        if (fail(cudaStreamSynchronize(st), "levels")) break;
        if (last[0] < last[1]) break;  // degenerate (very deep) tree: keep the binary-tree kernels
        const int levels = (int)h[0];
        mesh->levels8 = levels;
        uint32_t n8 = h[1];
        if (fail(cub::DeviceScan::ExclusiveSum(base + o_cub, cub_bytes, nleaf, tbase, (int)n8, st), "scan")) break;
        ctx->launches += 2;
        if (fail(cudaMalloc((void**)&mesh->nodes8, (size_t)n8 * 16 * W8_NODE_F4), "alloc")) break;
        if (fail(cudaMalloc((void**)&mesh->tris8, (size_t)mesh->nt * 16 * W8_TRI_F4), "alloc")) break;
        if (fail(cudaMemcpyAsync(mesh->nodes8, tmp8, (size_t)n8 * 16 * W8_NODE_F4, cudaMemcpyDeviceToDevice, st), "copy")) break;
        k_finalize8<<<pb2_blocks(n8, 128), 128, 0, st>>>(n8, tbase, leafpos, mesh->tris, mesh->nodes8, mesh->tris8);
        PB2_LAUNCHED(ctx);
        if (fail(cudaGetLastError(), "finalize")) break;
        mesh->n_nodes8 = n8;
    } while (0);
    cudaFreeAsync(base, st);
    if (s == PB2_OK && mesh->n_nodes8) { if (fail(cudaStreamSynchronize(st), "finalize")) mesh->n_nodes8 = 0; }
    if (s != PB2_OK || mesh->n_nodes8 == 0) {
        if (mesh->nodes8) cudaFree(mesh->nodes8);
        if (mesh->tris8) cudaFree(mesh->tris8);
        mesh->nodes8 = nullptr; mesh->tris8 = nullptr; mesh->n_nodes8 = 0;
    }
    return s;
}

// ------------------------------------------------------------------------------------------------ traversal
// bit s of an 8-bit slot mask -> bit (s ^ x)
__device__ __forceinline__ uint32_t perm8(uint32_t m, uint32_t x) {
    uint32_t t;
    t = ((m & 0x55u) << 1) | ((m >> 1) & 0x55u); m = (x & 1u) ? t : m;
    t = ((m & 0x33u) << 2) | ((m >> 2) & 0x33u); m = (x & 2u) ? t : m;
    t = ((m & 0x0fu) << 4) | ((m >> 4) & 0x0fu); m = (x & 4u) ? t : m;
    return m;
}

// Directed-rounding constants of one axis: entry >= fma_rd(u, Kn, cn), exit <= fma_ru(u, Kf, cf) with u = 1 + q * 2^-15.
struct AxisK {
    float Kn, Kf, cn, cf;
};
__device__ __forceinline__ AxisK axis_setup(float p, float o, float inv, bool neg, uint32_t ebyte) {
    float S = __uint_as_float(ebyte << 23);
    AxisK k;
    k.Kn = __fmul_rd(S, inv);
    k.Kf = __fmul_ru(S, inv);
    float slo = __fsub_rd(p, o), shi = __fsub_ru(p, o);
    float an = neg ? shi : slo, af = neg ? slo : shi;
    k.cn = __fsub_rd(__fmul_rd(an, inv), k.Kf);
    k.cf = __fsub_ru(__fmul_ru(af, inv), k.Kn);
    return k;
}

#define W8_U(word, k) __uint_as_float(__byte_perm((word), onef, 0x7604u | ((k) << 4)))
#define W8_CHILD(j, nxw, nyw, nzw, fxw, fyw, fzw)                                                   \
    {                                                                                               \
        float tnx = __fmaf_rd(W8_U(nxw, (j) & 3), kx.Kn, kx.cn), tfx = __fmaf_ru(W8_U(fxw, (j) & 3), kx.Kf, kx.cf); \
        float tny = __fmaf_rd(W8_U(nyw, (j) & 3), ky.Kn, ky.cn), tfy = __fmaf_ru(W8_U(fyw, (j) & 3), ky.Kf, ky.cf); \
        float tnz = __fmaf_rd(W8_U(nzw, (j) & 3), kz.Kn, kz.cn), tfz = __fmaf_ru(W8_U(fzw, (j) & 3), kz.Kf, kz.cf); \
        float tmin = fmaxf(fmaxf(fmaxf(tnx, tny), tnz), 0.0f);                                      \
        float tmax = fminf(fminf(fminf(tfx, tfy), tfz), best) * W8_GAMMA;                           \
        if (MODE == 1) tent[j][threadIdx.x] = tmin;                                                 \
        if (tmin <= tmax) hit8 |= 1u << (j);                                                        \
    }

// Persistent warps, one ray per lane, three phases per iteration:
//   refill   — lanes whose ray is finished pull new rays from a global counter once `refill` of them are idle;
//   node     — every lane with traversal work opens one wide node and tests its eight quantised child boxes. Hit
//              triangles are queued as a group. MODE 0 (default): the hit inner children form a group that is walked in
//              octant order (child slots were assigned by octant at build time), the rest of a group waits on the
//              lane's stack. MODE 1 (PB2_RAY_MODE=1, kept for comparison): the nearest hit child is opened next and
//              every other one is pushed with its entry distance so that stale entries are dropped at pop time; it
//              saves 10 % of the node visits (15.7 vs 17.4 per ray on the 8M-triangle terrain) but its per-visit
//              overhead (eight shared-memory stores + two short divergent loops) costs more than that;
//   triangle — run only when enough lanes have queued triangles (or nothing else can make progress), so that the exact
//              leaf-box + triangle test executes with many lanes active instead of one or two (run inline it took 40 %
//              of all issued instructions at 1.9 active lanes, profiles/r1_rays_v4_*).
// Queued triangles delay the update of `best`; the lane meanwhile keeps traversing against its older bound, which can
// only add node visits, never remove a candidate.
template <bool WITH_NORMAL, int MODE>
__global__ void __launch_bounds__(128, 7) k_raycast_wide(const float4* __restrict__ nodes8, const float4* __restrict__ tris8, uint32_t nt,
                                  const float* __restrict__ pose7, const float* __restrict__ rays, const uint32_t* __restrict__ perm,
                                  uint32_t m, float max_toi, float* __restrict__ out_toi, uint32_t* __restrict__ out_tri,
                                  float* __restrict__ out_normal, uint32_t* __restrict__ out_feature,
                                  unsigned int* __restrict__ next_ray, int tri_lanes, int refill, uint32_t cull,
                                  unsigned long long* __restrict__ stats) {
    __shared__ uint2 tq[W8_TQ][128];   // queued triangle groups: {first triangle, leaf mask | hit mask << 8}
    __shared__ float tent[MODE == 1 ? 8 : 1][128];  // entry distances of the node being tested
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    Iso7 pose;
    if (pose7) pose = load_iso(pose7);
    V3 o = mk3(0.f, 0.f, 0.f), d = o, inv = o, best_n = o;
    float best = 0.f;
    uint32_t best_id = PB2_INVALID_U32, best_fid = 0, r = 0;
    uint32_t oct = 0;                    // bit a set <=> d[a] < 0 (sign bit)
    uint32_t cur = PB2_INVALID_U32;      // wide node to open next
    uint32_t g_base = 0, g_bits = 0;     // MODE 0: node group {first child, (pending hits in octant priority order << 24) | imask}
    uint32_t tg_base = 0, tg_bits = 0;   // triangle group being worked on (bits: leaf mask | pending hits << 8)
    uint32_t qh = 0, nq = 0;             // further queued groups: ring head / length
    bool found = false, active = false;
    uint2 stack[W8_STACK];               // {wide node, entry distance}
    int sp = 0;
    bool exhausted = false;
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if (!exhausted && (idle == FULL || __popc(idle) >= refill)) {
            unsigned base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(next_ray, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base + __popc(idle) >= m) exhausted = true;
            if (!active) {
                uint32_t slot = base + __popc(idle & ((1u << lane) - 1u));
                if (slot < m) {
                    r = perm ? perm[slot] : slot;
                    o = mk3(rays[6ull * r], rays[6ull * r + 1], rays[6ull * r + 2]);
                    d = mk3(rays[6ull * r + 3], rays[6ull * r + 4], rays[6ull * r + 5]);
                    if (pose7) { o = iso_inv_point(pose, o); d = iso_inv_vec(pose, d); }
                    inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    oct = (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
                    best = max_toi; best_id = PB2_INVALID_U32; best_fid = 0; found = false;
                    cur = 0; sp = 0; qh = 0; nq = 0; tg_bits = 0; active = true;
                    g_base = 0; g_bits = ((1u << (7u ^ oct)) << 24) | 1u;  // root group: wide node 0 in slot 0
                }
            }
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) break;
        // ------------------------------------------------------------------ node phase
        const uint32_t pxor = 7u ^ oct;
        if (MODE == 0) {  // pick the highest-priority pending child of the current group
            cur = PB2_INVALID_U32;
            if ((g_bits >> 24) && nq < W8_TQ) {
                uint32_t hits = g_bits >> 24;
                uint32_t bsel = 31u - (uint32_t)__clz(hits);
                hits &= ~(1u << bsel);
                uint32_t pim = g_bits & 0xffu;
                cur = g_base + (uint32_t)__popc(pim & ((1u << (bsel ^ pxor)) - 1u));
                g_bits = (hits << 24) | pim;
                if (hits) { stack[sp] = make_uint2(g_base, g_bits); sp++; }
            }
        }
        if (cur != PB2_INVALID_U32 && nq < W8_TQ) {
            const float4* np = nodes8 + (size_t)W8_NODE_F4 * cur;
            F8 nA = ldg256(np), nB = ldg256(np + 2), nC = ldg256(np + 4);
            float4 n0 = nA.a, n1 = nA.b, n2 = nB.a, n3 = nB.b, n4 = nC.a;
            uint32_t ew = __float_as_uint(n0.w);
            const uint32_t onef = __float_as_uint(n1.w);  // 0x3F800000, see k_collapse8
            AxisK kx = axis_setup(n0.x, o.x, inv.x, oct & 1u, ew & 0xffu);
            AxisK ky = axis_setup(n0.y, o.y, inv.y, oct & 2u, (ew >> 8) & 0xffu);
            AxisK kz = axis_setup(n0.z, o.z, inv.z, oct & 4u, (ew >> 16) & 0xffu);
            uint32_t lx0 = __float_as_uint(n2.x), lx1 = __float_as_uint(n2.y), ly0 = __float_as_uint(n2.z), ly1 = __float_as_uint(n2.w);
            uint32_t lz0 = __float_as_uint(n3.x), lz1 = __float_as_uint(n3.y), hx0 = __float_as_uint(n3.z), hx1 = __float_as_uint(n3.w);
            uint32_t hy0 = __float_as_uint(n4.x), hy1 = __float_as_uint(n4.y), hz0 = __float_as_uint(n4.z), hz1 = __float_as_uint(n4.w);
            // entry planes are the low planes for a positive direction, the high planes for a negative one
            uint32_t nx0 = (oct & 1u) ? hx0 : lx0, nx1 = (oct & 1u) ? hx1 : lx1, fx0 = (oct & 1u) ? lx0 : hx0, fx1 = (oct & 1u) ? lx1 : hx1;
            uint32_t ny0 = (oct & 2u) ? hy0 : ly0, ny1 = (oct & 2u) ? hy1 : ly1, fy0 = (oct & 2u) ? ly0 : hy0, fy1 = (oct & 2u) ? ly1 : hy1;
            uint32_t nz0 = (oct & 4u) ? hz0 : lz0, nz1 = (oct & 4u) ? hz1 : lz1, fz0 = (oct & 4u) ? lz0 : hz0, fz1 = (oct & 4u) ? lz1 : hz1;
            uint32_t hit8 = 0;
            W8_CHILD(0, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(1, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(2, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(3, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(4, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(5, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(6, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(7, nx1, ny1, nz1, fx1, fy1, fz1)
            uint32_t imask = ew >> 24, lmask = __float_as_uint(n1.z);
            uint32_t cbase = __float_as_uint(n1.x), tbase = __float_as_uint(n1.y);
            if (stats) {  // debug counters (PB2_RAY_STATS): visits, visits without any hit, inner hits, leaf hits
                atomicAdd(&stats[0], 1ull);
                if (hit8 == 0) atomicAdd(&stats[1], 1ull);
                atomicAdd(&stats[2], (unsigned long long)__popc(hit8 & imask));
                atomicAdd(&stats[3], (unsigned long long)__popc(hit8 & lmask));
            }
            // triangles whose box was hit: one group
            uint32_t lh = hit8 & lmask;
            if (lh) {
                uint32_t bits = lmask | (lh << 8);
                if (tg_bits >> 8) { tq[(qh + nq) & (W8_TQ - 1)][threadIdx.x] = make_uint2(tbase, bits); nq++; }
                else { tg_base = tbase; tg_bits = bits; }
            }
            // hit inner children: open the nearest next, push the others with their entry distance
            uint32_t ih = hit8 & imask;
            if (MODE == 0) {
                ih = perm8(ih, pxor);
                if (ih) { g_base = cbase; g_bits = (ih << 24) | imask; }
                else if (sp > 0) { sp--; uint2 e = stack[sp]; g_base = e.x; g_bits = e.y; }
                else g_bits = 0;
            } else if (ih) {
                uint32_t ns = (uint32_t)__ffs(ih) - 1u;
                uint32_t rest = ih & (ih - 1u);
                if (rest) {
                    float ntm = tent[ns][threadIdx.x];
                    do {
                        uint32_t s = (uint32_t)__ffs(rest) - 1u;
                        rest &= rest - 1u;
                        float t = tent[s][threadIdx.x];
                        bool nearer = t < ntm;
                        uint32_t ps = nearer ? ns : s;
                        float pt = nearer ? ntm : t;
                        ns = nearer ? s : ns;
                        ntm = nearer ? t : ntm;
                        stack[sp] = make_uint2(cbase + (uint32_t)__popc(imask & ((1u << ps) - 1u)), __float_as_uint(pt));
                        sp++;
                    } while (rest);
                }
                cur = cbase + (uint32_t)__popc(imask & ((1u << ns) - 1u));
            } else {
                cur = PB2_INVALID_U32;
            }
        }
        // pop: entries farther than the current bound are dropped (same test as the box test that recorded them)
        if (MODE == 1 && cur == PB2_INVALID_U32 && sp > 0) {
            float bg = best * W8_GAMMA;
            do {
                sp--;
                uint2 e = stack[sp];
                if (__uint_as_float(e.y) <= bg) { cur = e.x; break; }
            } while (sp > 0);
        }
        // ------------------------------------------------------------------ triangle phase
        {
            unsigned has = __ballot_sync(FULL, (tg_bits >> 8) != 0);
            if (has) {
                unsigned full = __ballot_sync(FULL, nq >= W8_TQ);
                unsigned walking = __ballot_sync(FULL, (MODE == 0 ? (g_bits >> 24) != 0 : cur != PB2_INVALID_U32) && nq < W8_TQ);
                int nh = __popc(has);
                if (nh >= tri_lanes || full || !walking || 2 * nh >= __popc(~idle)) {
                    if (tg_bits >> 8) {
                        uint32_t lh = tg_bits >> 8, lmask = tg_bits & 0xffu;
                        uint32_t s = (uint32_t)__ffs(lh) - 1u;
                        lh &= lh - 1u;
                        uint32_t t = tg_base + (uint32_t)__popc(lmask & ((1u << s) - 1u));
                        tg_bits = lmask | (lh << 8);
                        if (lh == 0 && nq > 0) {
                            uint2 g = tq[qh & (W8_TQ - 1)][threadIdx.x];
                            qh++; nq--;
                            tg_base = g.x; tg_bits = g.y;
                        }
                        F8 tA = ldg256(tris8 + (size_t)W8_TRI_F4 * t), tB = ldg256(tris8 + (size_t)W8_TRI_F4 * t + 2);
                        float4 ta = tA.a, tb = tA.b, tc = tB.a;
                        // exact leaf AABB (Triangle::local_aabb) and the reference's node test against the best hit so far
                        float4 blo = make_float4(fminf(fminf(ta.x, tb.x), tc.x), fminf(fminf(ta.y, tb.y), tc.y), fminf(fminf(ta.z, tb.z), tc.z), 0.f);
                        float4 bhi = make_float4(fmaxf(fmaxf(ta.x, tb.x), tc.x), fmaxf(fmaxf(ta.y, tb.y), tc.y), fmaxf(fmaxf(ta.z, tb.z), tc.z), 0.f);
                        // The leaf box is tested against best * (1 + 2^-20), like the node boxes: a triangle hit at exactly the best toi
                        // (a ray through an edge shared by two triangles) can have a box entry one ulp beyond it, and whether it was
                        // looked at used to depend on which of the two was found first — i.e. on how the warps happened to pick up
                        // rays. With the slack every candidate within rounding of the best is tested whatever the order, so the
                        // result is the (toi, smallest id) minimum over all of them: deterministic, and equal to brute force.
                        float sc = slab_cost_bf(blo, bhi, o, inv, best * W8_GAMMA);
                        if (stats) { atomicAdd(&stats[4], 1ull); if (sc != FLT_MAX && sc <= best) atomicAdd(&stats[5], 1ull); }
                        if (sc != FLT_MAX) {
                            float toi; uint32_t fid; V3 n;
                            if (ray_triangle(mk3(ta.x, ta.y, ta.z), mk3(tb.x, tb.y, tb.z), mk3(tc.x, tc.y, tc.z), o, d, toi, fid, n) && toi <= best &&
                                (cull == 0u || (fid & 1u) == cull - 1u)) {  // RayCullingMode::check (ray_trimesh.rs:58-65)
                                uint32_t id = __float_as_uint(ta.w);
                                if (toi < best || (found && toi == best && id < best_id)) {
                                    best = toi; best_id = id; best_fid = fid; found = true;
                                    if (WITH_NORMAL) best_n = n;
                                }
                            }
                        }
                    }
                }
            }
        }
        // ------------------------------------------------------------------ retire
        if (active && (MODE == 0 ? (g_bits >> 24) == 0 : (cur == PB2_INVALID_U32 && sp == 0)) && (tg_bits >> 8) == 0) {
            out_toi[r] = found ? best : 0.0f;
            out_tri[r] = best_id;
            if (WITH_NORMAL) {
                V3 n = mk3(0.f, 0.f, 0.f);
                uint32_t feat = PB2_INVALID_U32;
                if (found) {
                    n = normalize3(best_n);
                    if (best_fid & 2u) n = -n;
                    if (pose7) n = iso_vec(pose, n);
                    feat = (best_fid & 1u) ? best_id + nt : best_id;
                }
                if (out_normal) { out_normal[3ull * r] = n.x; out_normal[3ull * r + 1] = n.y; out_normal[3ull * r + 2] = n.z; }
                if (out_feature) out_feature[r] = feat;
            }
            active = false;
        }
    }
}

// ------------------------------------------------------------------------------------------------ shared triangle phase
// Same traversal as k_raycast_wide<., 0>, different triangle phase. The ncu source page of that kernel
// (profiles/r1_rays_v7_*) shows the exact leaf-box + triangle tests taking 20 % of the issued instructions with 8-10 of 32
// lanes active: a lane only ever tests its own ray's triangles, one per trip. Here the hit triangle groups of all lanes
// go to one per-warp queue, and when the queue is processed its triangles are dealt out to the 32 lanes regardless of
// which lane's ray they belong to: the ray (origin, direction, reciprocal) sits in shared memory, and the per-ray best hit
// is a 64-bit shared-memory word {|toi| bits, triangle id} updated with atomicMin — which is exactly the documented tie
// rule (smallest toi, then smallest triangle id). The lane that owns the winning candidate then stores the payload (exact
// toi bits, feature, normal) for the ray's owner.
#define W8C_GQ 64   // queued groups per warp (a node phase adds at most 32)
// PIECES: the batch is split into ranges of `pieces.size` rays whose completion is published while the kernel runs — each
// warp counts the rays it has retired per range when it refills (and at exit) and, every 8 refills, fences its result
// stores and adds the count to pieces.done[range]; whoever completes a range sets pieces.flag[range], which copy-engine streams wait on
// (cuStreamWaitValue32) to push that range of results to the peer GPUs. One launch keeps the SMs full across range
// boundaries, which separate launches per range cannot (each boundary cost ~40 us of ramp-down).
// Publication of a warp's retired-ray counts (PIECES): rare (every `flush_every` refills), kept out of line so that the traversal
// loop's instruction footprint stays the plain kernel's. Every lane of the warp calls it.
static __device__ __noinline__ void pz_flush_warp(volatile uint32_t* pz, PieceSignal pieces, uint32_t m, int lane) {
    __threadfence();   // every lane's result stores are visible before the counts are
    const uint32_t pa = pz[0], ca = pz[1], pb = pz[2], cb = pz[3];
    __syncwarp();      // every lane has read the counts before anybody clears them
    if (lane == 0) {
        const uint32_t pc[2][2] = {{pa, ca}, {pb, cb}};
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const uint32_t piece = pc[i][0], count = pc[i][1];
            if (!count) continue;
            uint32_t lo_r = piece * pieces.size;
            uint32_t total = m - lo_r < pieces.size ? m - lo_r : pieces.size;
            unsigned prev = atomicAdd(&pieces.done[piece], count);
            if (prev + count == total) { __threadfence(); atomicExch(&pieces.flag[piece], 1u); }
        }
    }
    pz[1] = 0; pz[3] = 0; pz[4] = 0;
    __syncwarp();
}

template <bool WITH_NORMAL, bool PIECES>
__global__ void __launch_bounds__(128, 8) k_raycast_wide_shared(const float4* __restrict__ nodes8, const float4* __restrict__ tris8, uint32_t nt,
                                  const float* __restrict__ pose7, const float* __restrict__ rays, const uint32_t* __restrict__ perm,
                                  uint32_t m, float max_toi, float* __restrict__ out_toi, uint32_t* __restrict__ out_tri,
                                  float* __restrict__ out_normal, uint32_t* __restrict__ out_feature,
                                  unsigned int* __restrict__ next_ray, int tri_groups, int refill, uint32_t cull, int blocked_max,
                                  unsigned long long* __restrict__ stats, PieceSignal pieces) {
    __shared__ float s_ray[9][128];                // o, d, 1/d of the ray each thread owns
    __shared__ float s_stage[9][128];              // per warp: the next <= 32 rays, loaded and divided at full width (see the refill)
    __shared__ uint32_t s_stage_r[128];            // their ray indices
    __shared__ unsigned long long s_key[128];      // best hit so far: |toi| bits << 32 | triangle id (id INVALID: none yet)
    __shared__ uint2 s_pay[128];                   // {exact toi bits, feature bits} of that hit
    __shared__ float s_nrm[WITH_NORMAL ? 3 : 1][128];
    __shared__ uint2 s_gq[4][W8C_GQ];              // {first triangle, leaf mask | hit mask << 8 | owner lane << 16}
    __shared__ uint16_t s_items[4][256];           // queue slot << 3 | child slot, one per triangle to test
    // PIECES bookkeeping, per warp: {range a, count a, range b, count b, refills since the last publication}. In shared memory so
    // that the variant keeps the plain kernel's register budget (at 64 registers five more live values spilled inside the hot loop:
    // 2.40 instead of 2.09 ms on one GPU)
    __shared__ uint32_t s_pz[PIECES ? 4 : 1][5];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, wbase = threadIdx.x & ~31;
    const unsigned lt = (1u << lane) - 1u;
    const int pieces_flush_every = PIECES ? (int)pieces.flush_every : 0;
    Iso7 pose;
    if (pose7) pose = load_iso(pose7);
    V3 o = mk3(0.f, 0.f, 0.f), inv = o;
    float best = 0.f;
    uint32_t r = PB2_INVALID_U32, oct = 0, cur = PB2_INVALID_U32;   // r == INVALID: no ray yet (PIECES: nothing to count)
    uint32_t g_base = 0, g_bits = 0;
    bool active = false, pend = false;
    uint2 stack[W8_STACK];
    int sp = 0;
    uint32_t G = 0;                                // groups in this warp's queue (warp uniform)
    bool exhausted = false;
    // PIECES: retired rays not yet published, for the two ranges a warp can hold rays of around a range boundary. The five words are
    // warp-uniform: every lane computes and stores the same values (volatile, so that they are not carried in registers through the
    // hot loop; no lane-0 branch). Every read-modify-write is read -> __syncwarp -> write -> __syncwarp, so a lane that runs ahead
    // can never read a value another lane has already updated.
    volatile uint32_t* const pz = s_pz[PIECES ? w : 0];
    if (PIECES) { pz[0] = PB2_INVALID_U32; pz[1] = 0; pz[2] = PB2_INVALID_U32; pz[3] = 0; pz[4] = 0; __syncwarp(); }
    auto pz_flush = [&]() { pz_flush_warp(pz, pieces, m, lane); };   // every lane of the warp
    // rays retired since the last refill are counted when their lane is handed a new ray (and at exit): nothing per trip. A lane that
    // has been counted forgets its ray index (r = INVALID), which is what marks it as counted.
    auto pz_collect = [&]() {
        uint32_t mine = (!active && r != PB2_INVALID_U32) ? r >> pieces.shift : PB2_INVALID_U32;   // ranges are powers of two
        unsigned cm = __ballot_sync(FULL, mine != PB2_INVALID_U32);
        while (cm) {   // usually one range
            uint32_t p0 = __shfl_sync(FULL, mine, __ffs(cm) - 1);
            unsigned grp = __ballot_sync(FULL, mine == p0);
            cm &= ~grp;
            uint32_t c = (uint32_t)__popc(grp);
            const uint32_t pa = pz[0], ca = pz[1], pb = pz[2], cb = pz[3];
            __syncwarp();
            if (p0 == pa) pz[1] = ca + c;
            else if (p0 == pb) pz[3] = cb + c;
            else if (ca == 0) { pz[0] = p0; pz[1] = c; }
            else if (cb == 0) { pz[2] = p0; pz[3] = c; }
            else { pz_flush(); pz[0] = p0; pz[1] = c; }
            __syncwarp();
        }
        if (!active) r = PB2_INVALID_U32;
    };
    // Ray staging (round 2). Round 1 handed a new ray to each idle lane straight from global memory: one atomicAdd, six scattered
    // 4-byte loads and three IEEE divisions executed by the ~8 idle lanes of a refill while the other 24 waited behind them — 11 % of
    // the instruction stream at 17 lanes and 18 % of the stall samples (profiles/r1_rays_v8_*, source page). Now the warp claims rays
    // 32 at a time: all 32 lanes load one ray each (768 contiguous bytes), transform it, divide, and park {o, d, 1/d} in shared
    // memory; idle lanes are then served from that stage with nine shared-memory reads. The global round trips are paid once per 32
    // rays instead of once per refill, the divisions run at full width, and a refill is cheap enough to run as soon as a few lanes idle.
    uint32_t st_base = 0, st_count = 0, st_next = 0;   // staged batch: first ray slot, rays staged, rays handed out (warp uniform)
    for (;;) {
        unsigned idle = __ballot_sync(FULL, !active);
        if ((!exhausted || st_next < st_count) && (idle == FULL || __popc(idle) >= refill)) {
            if (PIECES) {
                pz_collect();
                const uint32_t since = pz[4] + 1u;
                const bool due = since >= (uint32_t)pieces_flush_every && (pz[1] | pz[3]) != 0u;
                __syncwarp();
                if (due) pz_flush();
                else { pz[4] = since; __syncwarp(); }
            }
            for (;;) {
                if (st_next == st_count) {
                    if (exhausted) break;
                    unsigned base = 0;
                    if (lane == 0) base = atomicAdd(next_ray, 32u);
                    base = __shfl_sync(FULL, base, 0);
                    if (base >= m) { exhausted = true; break; }
                    if (base + 32u >= m) exhausted = true;
                    const uint32_t cnt = m - base < 32u ? m - base : 32u;
                    if ((uint32_t)lane < cnt) {
                        uint32_t rr = perm ? perm[base + lane] : base + lane;
                        V3 so = mk3(rays[6ull * rr], rays[6ull * rr + 1], rays[6ull * rr + 2]);
                        V3 sd = mk3(rays[6ull * rr + 3], rays[6ull * rr + 4], rays[6ull * rr + 5]);
                        if (pose7) { so = iso_inv_point(pose, so); sd = iso_inv_vec(pose, sd); }
                        s_stage[0][threadIdx.x] = so.x; s_stage[1][threadIdx.x] = so.y; s_stage[2][threadIdx.x] = so.z;
                        s_stage[3][threadIdx.x] = sd.x; s_stage[4][threadIdx.x] = sd.y; s_stage[5][threadIdx.x] = sd.z;
                        s_stage[6][threadIdx.x] = 1.0f / sd.x; s_stage[7][threadIdx.x] = 1.0f / sd.y; s_stage[8][threadIdx.x] = 1.0f / sd.z;
                        s_stage_r[threadIdx.x] = rr;
                    }
                    __syncwarp();
                    st_base = base; st_count = cnt; st_next = 0;
                }
                const uint32_t navail = st_count - st_next, nidle = (uint32_t)__popc(idle);
                const uint32_t take = navail < nidle ? navail : nidle;
                const uint32_t rank = (uint32_t)__popc(idle & lt);
                if (!active && rank < take) {
                    const int sl = wbase + (int)(st_next + rank);
                    r = s_stage_r[sl];
                    o = mk3(s_stage[0][sl], s_stage[1][sl], s_stage[2][sl]);
                    V3 d = mk3(s_stage[3][sl], s_stage[4][sl], s_stage[5][sl]);
                    inv = mk3(s_stage[6][sl], s_stage[7][sl], s_stage[8][sl]);
                    oct = (__float_as_uint(d.x) >> 31) | ((__float_as_uint(d.y) >> 31) << 1) | ((__float_as_uint(d.z) >> 31) << 2);
                    best = max_toi;
                    sp = 0; active = true; pend = false;
                    g_base = 0; g_bits = ((1u << (7u ^ oct)) << 24) | 1u;
                    s_ray[0][threadIdx.x] = o.x; s_ray[1][threadIdx.x] = o.y; s_ray[2][threadIdx.x] = o.z;
                    s_ray[3][threadIdx.x] = d.x; s_ray[4][threadIdx.x] = d.y; s_ray[5][threadIdx.x] = d.z;
                    s_ray[6][threadIdx.x] = inv.x; s_ray[7][threadIdx.x] = inv.y; s_ray[8][threadIdx.x] = inv.z;
                    s_key[threadIdx.x] = ((unsigned long long)__float_as_uint(max_toi) << 32) | 0xffffffffull;
                }
                st_next += take;
                idle = __ballot_sync(FULL, !active);
                if (idle == 0u) break;
                __syncwarp();   // the stage is about to be overwritten: every lane has read its slot
            }
            (void)st_base;
            idle = __ballot_sync(FULL, !active);
        }
        if (idle == FULL) break;
        // ------------------------------------------------------------------ node phase
        const uint32_t pxor = 7u ^ oct;
        cur = PB2_INVALID_U32;
        if (g_bits >> 24) {
            uint32_t hits = g_bits >> 24;
            uint32_t bsel = 31u - (uint32_t)__clz(hits);
            hits &= ~(1u << bsel);
            uint32_t pim = g_bits & 0xffu;
            cur = g_base + (uint32_t)__popc(pim & ((1u << (bsel ^ pxor)) - 1u));
            g_bits = (hits << 24) | pim;
            if (hits) { stack[sp] = make_uint2(g_base, g_bits); sp++; }
        }
        uint32_t lh = 0, q_tbase = 0, q_lmask = 0;
        if (cur != PB2_INVALID_U32) {
            const int MODE = 0;
            float (*tent)[128] = nullptr;
            (void)tent;
            const float4* np = nodes8 + (size_t)W8_NODE_F4 * cur;
            F8 nA = ldg256(np), nB = ldg256(np + 2), nC = ldg256(np + 4);
            float4 n0 = nA.a, n1 = nA.b, n2 = nB.a, n3 = nB.b, n4 = nC.a;
            uint32_t ew = __float_as_uint(n0.w);
            const uint32_t onef = __float_as_uint(n1.w);
            AxisK kx = axis_setup(n0.x, o.x, inv.x, oct & 1u, ew & 0xffu);
            AxisK ky = axis_setup(n0.y, o.y, inv.y, oct & 2u, (ew >> 8) & 0xffu);
            AxisK kz = axis_setup(n0.z, o.z, inv.z, oct & 4u, (ew >> 16) & 0xffu);
            uint32_t lx0 = __float_as_uint(n2.x), lx1 = __float_as_uint(n2.y), ly0 = __float_as_uint(n2.z), ly1 = __float_as_uint(n2.w);
            uint32_t lz0 = __float_as_uint(n3.x), lz1 = __float_as_uint(n3.y), hx0 = __float_as_uint(n3.z), hx1 = __float_as_uint(n3.w);
            uint32_t hy0 = __float_as_uint(n4.x), hy1 = __float_as_uint(n4.y), hz0 = __float_as_uint(n4.z), hz1 = __float_as_uint(n4.w);
            uint32_t nx0 = (oct & 1u) ? hx0 : lx0, nx1 = (oct & 1u) ? hx1 : lx1, fx0 = (oct & 1u) ? lx0 : hx0, fx1 = (oct & 1u) ? lx1 : hx1;
            uint32_t ny0 = (oct & 2u) ? hy0 : ly0, ny1 = (oct & 2u) ? hy1 : ly1, fy0 = (oct & 2u) ? ly0 : hy0, fy1 = (oct & 2u) ? ly1 : hy1;
            uint32_t nz0 = (oct & 4u) ? hz0 : lz0, nz1 = (oct & 4u) ? hz1 : lz1, fz0 = (oct & 4u) ? lz0 : hz0, fz1 = (oct & 4u) ? lz1 : hz1;
            uint32_t hit8 = 0;
            W8_CHILD(0, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(1, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(2, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(3, nx0, ny0, nz0, fx0, fy0, fz0)
            W8_CHILD(4, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(5, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(6, nx1, ny1, nz1, fx1, fy1, fz1)
            W8_CHILD(7, nx1, ny1, nz1, fx1, fy1, fz1)
            uint32_t imask = ew >> 24;
            q_lmask = __float_as_uint(n1.z);
            q_tbase = __float_as_uint(n1.y);
            lh = hit8 & q_lmask;
            if (stats) {
                atomicAdd(&stats[0], 1ull);
                if (hit8 == 0) atomicAdd(&stats[1], 1ull);
                atomicAdd(&stats[2], (unsigned long long)__popc(hit8 & imask));
                atomicAdd(&stats[3], (unsigned long long)__popc(lh));
            }
            uint32_t ih = perm8(hit8 & imask, pxor);
            if (ih) { g_base = __float_as_uint(n1.x); g_bits = (ih << 24) | imask; }
            else if (sp > 0) { sp--; uint2 e = stack[sp]; g_base = e.x; g_bits = e.y; }
            else g_bits = 0;
        }
        // hit triangle groups go to the warp's queue
        {
            unsigned pm = __ballot_sync(FULL, lh != 0);
            if (lh) {
                s_gq[w][G + __popc(pm & lt)] = make_uint2(q_tbase, q_lmask | (lh << 8) | ((uint32_t)lane << 16));
                pend = true;
            }
            G += (uint32_t)__popc(pm);
        }
        // ------------------------------------------------------------------ triangle phase
        if (G) {
            unsigned walking = __ballot_sync(FULL, active && (g_bits >> 24) != 0);
            unsigned blocked = __ballot_sync(FULL, active && pend && (g_bits >> 24) == 0);
            if ((int)G >= tri_groups || G > W8C_GQ - 32 || !walking || __popc(blocked) >= blocked_max) {
                __syncwarp();
                for (uint32_t gb = 0; gb < G; gb += 32) {
                    // deal the triangles of up to 32 groups out as items
                    uint32_t gi = gb + (uint32_t)lane, glh = 0;
                    if (gi < G) glh = (s_gq[w][gi].y >> 8) & 0xffu;
                    int c = __popc(glh), pre = c;
#pragma unroll
                    for (int dlt = 1; dlt < 32; dlt <<= 1) {
                        int v = __shfl_up_sync(FULL, pre, dlt);
                        if (lane >= dlt) pre += v;
                    }
                    const int T = __shfl_sync(FULL, pre, 31);
                    int at = pre - c;
                    while (glh) {
                        uint32_t sl = (uint32_t)__ffs(glh) - 1u;
                        glh &= glh - 1u;
                        s_items[w][at++] = (uint16_t)((lane << 3) | sl);
                    }
                    __syncwarp();
                    for (int j0 = 0; j0 < T; j0 += 32) {
                        const int j = j0 + lane;
                        bool cand = false;
                        unsigned long long mine = 0;
                        int own = 0;
                        uint32_t c_toi = 0, c_fid = 0;
                        V3 c_n = mk3(0.f, 0.f, 0.f);
                        if (j < T) {
                            uint32_t it = s_items[w][j];
                            uint2 ge = s_gq[w][gb + (it >> 3)];
                            own = wbase + (int)(ge.y >> 16);
                            uint32_t t = ge.x + (uint32_t)__popc(ge.y & 0xffu & ((1u << (it & 7u)) - 1u));
                            F8 tA = ldg256(tris8 + (size_t)W8_TRI_F4 * t), tB = ldg256(tris8 + (size_t)W8_TRI_F4 * t + 2);
                        float4 ta = tA.a, tb = tA.b, tc = tB.a;
                            unsigned long long k0 = s_key[own];
                            float sbest = __uint_as_float((uint32_t)(k0 >> 32));
                            uint32_t sid = (uint32_t)k0;
                            bool sfound = sid != PB2_INVALID_U32;
                            V3 ro = mk3(s_ray[0][own], s_ray[1][own], s_ray[2][own]);
                            V3 rinv = mk3(s_ray[6][own], s_ray[7][own], s_ray[8][own]);
                            // exact leaf AABB (Triangle::local_aabb) and the reference's node test against the best hit so far
                            float4 blo = make_float4(fminf(fminf(ta.x, tb.x), tc.x), fminf(fminf(ta.y, tb.y), tc.y), fminf(fminf(ta.z, tb.z), tc.z), 0.f);
                            float4 bhi = make_float4(fmaxf(fmaxf(ta.x, tb.x), tc.x), fmaxf(fmaxf(ta.y, tb.y), tc.y), fmaxf(fmaxf(ta.z, tb.z), tc.z), 0.f);
                            // slack of 2^-20 on the bound, as for the node boxes: see k_raycast_wide (order-independent ties)
                            float sc = slab_cost_bf(blo, bhi, ro, rinv, sbest * W8_GAMMA);
                            if (stats) { atomicAdd(&stats[4], 1ull); if (sc != FLT_MAX && sc <= sbest) atomicAdd(&stats[5], 1ull); }
                            if (sc != FLT_MAX) {
                                V3 rd = mk3(s_ray[3][own], s_ray[4][own], s_ray[5][own]);
                                float toi; uint32_t fid; V3 n;
                                if (ray_triangle(mk3(ta.x, ta.y, ta.z), mk3(tb.x, tb.y, tb.z), mk3(tc.x, tc.y, tc.z), ro, rd, toi, fid, n) && toi <= sbest &&
                                    (cull == 0u || (fid & 1u) == cull - 1u)) {  // RayCullingMode::check (ray_trimesh.rs:58-65)
                                    uint32_t id = __float_as_uint(ta.w);
                                    if (toi < sbest || (sfound && toi == sbest && id < sid)) {
                                        cand = true;
                                        mine = ((unsigned long long)__float_as_uint(fabsf(toi)) << 32) | id;
                                        c_toi = __float_as_uint(toi); c_fid = fid; c_n = n;
                                        atomicMin(&s_key[own], mine);
                                    }
                                }
                            }
                        }
                        __syncwarp();
                        if (cand && s_key[own] == mine) {
                            s_pay[own] = make_uint2(c_toi, c_fid);
                            if (WITH_NORMAL) { s_nrm[0][own] = c_n.x; s_nrm[1][own] = c_n.y; s_nrm[2][own] = c_n.z; }
                        }
                        __syncwarp();
                    }
                }
                G = 0;
                pend = false;
                best = __uint_as_float((uint32_t)(s_key[threadIdx.x] >> 32));
            }
        }
        // ------------------------------------------------------------------ retire
        if (active && (g_bits >> 24) == 0 && !pend) {
            unsigned long long k = s_key[threadIdx.x];
            uint32_t best_id = (uint32_t)k;
            bool found = best_id != PB2_INVALID_U32;
            uint2 pay = s_pay[threadIdx.x];
            out_toi[r] = found ? __uint_as_float(pay.x) : 0.0f;
            out_tri[r] = best_id;
            if (WITH_NORMAL) {
                V3 n = mk3(0.f, 0.f, 0.f);
                uint32_t feat = PB2_INVALID_U32;
                if (found) {
                    n = normalize3(mk3(s_nrm[0][threadIdx.x], s_nrm[1][threadIdx.x], s_nrm[2][threadIdx.x]));
                    if (pay.y & 2u) n = -n;
                    if (pose7) n = iso_vec(pose, n);
                    feat = (pay.y & 1u) ? best_id + nt : best_id;
                }
                if (out_normal) { out_normal[3ull * r] = n.x; out_normal[3ull * r + 1] = n.y; out_normal[3ull * r + 2] = n.z; }
                if (out_feature) out_feature[r] = feat;
            }
            active = false;
        }
    }
    if (PIECES) { pz_collect(); if (pz[1] | pz[3]) pz_flush(); }
}

int pb2_wide_cast(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* d_pose, const float* d_rays, const uint32_t* d_perm, uint32_t m,
                  float max_toi, float* d_toi, uint32_t* d_tri, float* d_n, uint32_t* d_f, bool with_normal, int tri_lanes, int refill, uint32_t cull,
                  int shared_tri, const PieceSignal* pieces) {
    unsigned int* next_ray = (unsigned int*)(ctx->d_counters + ctx->ray_slot);
    int mode = 0;
    { const char* e = getenv("PB2_RAY_MODE"); if (e) mode = atoi(e) ? 1 : 0; }
    if (mesh->levels8 > W8_MAX_LEVELS) mode = 0;
    auto kern = with_normal ? (mode ? k_raycast_wide<true, 1> : k_raycast_wide<true, 0>) : (mode ? k_raycast_wide<false, 1> : k_raycast_wide<false, 0>);
    const bool pz = pieces != nullptr && shared_tri && !with_normal && d_perm == nullptr;
    if (pieces && !pz) return PB2_ERR_INVALID;
    auto kern_shared = pz ? k_raycast_wide_shared<false, true> : (with_normal ? k_raycast_wide_shared<true, false> : k_raycast_wide_shared<false, false>);
    PieceSignal no_pieces = {0u, nullptr, nullptr, 0u, 0u};
    int tri_groups = 12, blocked_max = 8;
    { const char* e = getenv("PB2_RAY_TRI_GROUPS"); if (e) tri_groups = atoi(e); }
    { const char* e = getenv("PB2_RAY_BLOCKED"); if (e) blocked_max = atoi(e); }
    int per_sm = 0;
    if (shared_tri) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern_shared, 128, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, 0);
    if (per_sm < 1) per_sm = 1;
    { const char* e = getenv("PB2_RAY_CTAS_PER_SM"); if (e && atoi(e) > 0 && atoi(e) < per_sm) per_sm = atoi(e); }
    unsigned blocks = (unsigned)(ctx->sm_count * per_sm);
    unsigned need = pb2_blocks(m, 128);
    if (blocks > need) blocks = need;
    unsigned long long* stats = nullptr;
    const char* se = getenv("PB2_RAY_STATS");
    if (se && atoi(se)) {
        static unsigned long long* d_stats = nullptr;
        if (!d_stats) cudaMalloc((void**)&d_stats, 64);
        cudaMemsetAsync(d_stats, 0, 64, ctx->stream);
        stats = d_stats;
    }
    // Optional L2 residency hint for the node array (PB2_RAY_L2_PERSIST = percent of the persisting carve-out to use, 0 = off):
    // triangles (3x the bytes of the nodes) and, on several GPUs, the gathered results stream through L2 and evict nodes.
    int l2_pct = 0;
    { const char* e = getenv("PB2_RAY_L2_PERSIST"); if (e) l2_pct = atoi(e); }
    bool l2_set = false;
    if (l2_pct > 0) {
        int max_persist = 0, max_window = 0;
        cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, ctx->device);
        cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, ctx->device);
        if (max_persist > 0 && max_window > 0) {
            size_t carve = (size_t)max_persist * (size_t)(l2_pct > 100 ? 100 : l2_pct) / 100;
            cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, carve);
            size_t bytes = (size_t)mesh->n_nodes8 * 16 * W8_NODE_F4;
            if (bytes > (size_t)max_window) bytes = (size_t)max_window;
            cudaStreamAttrValue attr;
            memset(&attr, 0, sizeof(attr));
            attr.accessPolicyWindow.base_ptr = (void*)mesh->nodes8;
            attr.accessPolicyWindow.num_bytes = bytes;
            attr.accessPolicyWindow.hitRatio = carve >= bytes ? 1.0f : (float)((double)carve / (double)bytes);
            attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
            attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
            l2_set = cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr) == cudaSuccess;
            if (getenv("PB2_RAY_L2_VERBOSE")) fprintf(stderr, "[pb2] L2 persist: carve %zu of max %d, window %zu (max %d), hit ratio %.2f, set %d\n", carve, max_persist, bytes, max_window, attr.accessPolicyWindow.hitRatio, (int)l2_set);
            (void)cudaGetLastError();
        }
    }
    if (shared_tri)
        kern_shared<<<blocks, 128, 0, ctx->stream>>>(mesh->nodes8, mesh->tris8, mesh->nt, d_pose, d_rays, d_perm, m, max_toi, d_toi, d_tri,
                                                     with_normal ? d_n : nullptr, with_normal ? d_f : nullptr, next_ray, tri_groups, refill, cull,
                                                     blocked_max, stats, pz ? *pieces : no_pieces);
    else
        kern<<<blocks, 128, 0, ctx->stream>>>(mesh->nodes8, mesh->tris8, mesh->nt, d_pose, d_rays, d_perm, m, max_toi, d_toi, d_tri,
                                              with_normal ? d_n : nullptr, with_normal ? d_f : nullptr, next_ray, tri_lanes, refill, cull, stats);
    if (l2_set) {
        cudaStreamAttrValue attr;
        memset(&attr, 0, sizeof(attr));
        cudaStreamSetAttribute(ctx->stream, cudaStreamAttributeAccessPolicyWindow, &attr);   // num_bytes = 0 disables the window
    }
    if (stats) {
        unsigned long long h[8];
        cudaMemcpyAsync(h, stats, 64, cudaMemcpyDeviceToHost, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
        fprintf(stderr, "[pb2 ray stats] rays %u nodes8 %u | visits %llu (%.2f/ray) empty %llu inner-hits %llu leaf-hits %llu | tri tests %llu slab-pass %llu\n",
                m, mesh->n_nodes8, h[0], (double)h[0] / m, h[1], h[2], h[3], h[4], h[5]);
    }
    return PB2_OK;
}

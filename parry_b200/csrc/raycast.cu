// parry_b200 — batched ray casts.
//
// Replaces (reference, file:line): RayCast::cast_ray / cast_ray_and_get_normal (query/ray/ray.rs:381-411) for TriMesh
// (query/ray/ray_trimesh.rs:8-36 -> ray_composite_shape.rs:20-62 -> Bvh::cast_ray bvh_queries.rs:260-271 ->
// Bvh::find_best bvh_traverse.rs:335-417), BvhNode::cast_ray (bvh_tree.rs:1177-1181) = Aabb::cast_local_ray
// (ray_aabb.rs:12-49), Triangle ray test (ray_triangle.rs:49-152), TriMesh::triangle (shape/trimesh.rs:1896-1903).
//
// One thread per ray, ordered depth-first descent with a short per-thread stack (nearer child first, farther child
// pushed), pruning against the best hit so far exactly like find_best. Triangles are pre-gathered in BVH leaf
// order as 3 x float4 (48 B, 16-B aligned vector loads) so a leaf visit costs no index indirection.
#include "trimesh.cuh"
#include "traverse.cuh"
#include <stdlib.h>
#include <cub/device/device_radix_sort.cuh>

int pb2_bvh_build_device(pb2_ctx* ctx, pb2_bvh* b, const float* d_aabbs, uint32_t n, bool resolve_flags, const uint32_t* leaf_data);
int pb2_stage_in(pb2_ctx* ctx, int slot, const void* src, size_t bytes, int mem, const void** out);
int pb2_stage_out(pb2_ctx* ctx, int slot, void* dst, size_t bytes, int mem, void** out);
int pb2_stage_back(pb2_ctx* ctx, void* dst, const void* dev, size_t bytes, int mem);



// Triangle::local_aabb (bounding_volume/aabb_triangle.rs:16-30)
__global__ void k_triangle_aabbs(const float* __restrict__ v, const uint32_t* __restrict__ idx, uint32_t nt, uint32_t nv,
                                 float* __restrict__ aabbs, uint32_t* __restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nt) return;
    uint32_t ia = idx[3ull * i], ib = idx[3ull * i + 1], ic = idx[3ull * i + 2];
    if (ia >= nv || ib >= nv || ic >= nv) { atomicAdd(bad, 1u); ia = ib = ic = 0; }
    float* o = aabbs + 6ull * i;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        float a = v[3ull * ia + d], b = v[3ull * ib + d], c = v[3ull * ic + d];
        o[d] = fminf(fminf(a, b), c);
        o[3 + d] = fmaxf(fmaxf(a, b), c);
    }
}

__global__ void k_gather_triangles(const float* __restrict__ v, const uint32_t* __restrict__ idx, uint32_t nt,
                                   const uint32_t* __restrict__ order, float4* __restrict__ tris) {
    uint32_t p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= nt) return;
    uint32_t id = order[p];
    uint32_t ia = idx[3ull * id], ib = idx[3ull * id + 1], ic = idx[3ull * id + 2];
    tris[3ull * p + 0] = make_float4(v[3ull * ia], v[3ull * ia + 1], v[3ull * ia + 2], __uint_as_float(id));
    tris[3ull * p + 1] = make_float4(v[3ull * ib], v[3ull * ib + 1], v[3ull * ib + 2], 0.0f);
    tris[3ull * p + 2] = make_float4(v[3ull * ic], v[3ull * ic + 1], v[3ull * ic + 2], 0.0f);
}

template <bool WITH_NORMAL>
__global__ void __launch_bounds__(128) k_raycast_trimesh(const NodeWide* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n_leaves,
                                  uint32_t nt, const float* __restrict__ pose7, const float* __restrict__ rays, uint32_t m,
                                  float max_toi, float* __restrict__ out_toi, uint32_t* __restrict__ out_tri,
                                  float* __restrict__ out_normal, uint32_t* __restrict__ out_feature, uint32_t cull, unsigned int* fault) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    V3 o = mk3(rays[6ull * r], rays[6ull * r + 1], rays[6ull * r + 2]);
    V3 d = mk3(rays[6ull * r + 3], rays[6ull * r + 4], rays[6ull * r + 5]);
    Iso7 pose;
    if (pose7) {  // Ray::inverse_transform_by (ray.rs:167-172)
        pose = load_iso(pose7);
        o = iso_inv_point(pose, o);
        d = iso_inv_vec(pose, d);
    }
    V3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);

    float best = max_toi;
    uint32_t best_id = PB2_INVALID_U32;
    uint32_t best_fid = 0;
    V3 best_n = mk3(0.f, 0.f, 0.f);
    bool found = false;

    auto leaf_test = [&](uint32_t pos) {
        float4 ta = __ldg(&tris[3ull * pos]), tb = __ldg(&tris[3ull * pos + 1]), tc = __ldg(&tris[3ull * pos + 2]);
        float toi; uint32_t fid; V3 n;
        if (!ray_triangle(mk3(ta.x, ta.y, ta.z), mk3(tb.x, tb.y, tb.z), mk3(tc.x, tc.y, tc.z), o, d, toi, fid, n)) return;
        if (!(toi <= best)) return;  // Triangle::cast_local_ray_and_get_normal: toi <= max_toi(=best so far)
        if (cull && (fid & 1u) != cull - 1u) return;  // RayCullingMode::check (ray_trimesh.rs:58-65): 1 front faces only, 2 back faces only
        uint32_t id = __float_as_uint(ta.w);
        // find_best keeps strictly better hits; exact ties resolve to the smallest triangle index.
        if (toi < best || (found && toi == best && id < best_id)) {
            best = toi; best_id = id; best_fid = fid; best_n = n; found = true;
        }
    };

    bvh_find_best(nodes, n_leaves, o, d, inv, max_toi, best, found, leaf_test, fault);

    // CompositeShapeRef::cast_local_ray post-filter `toi < max_toi` holds by construction (strict accept vs
    // the initial best = max_toi).
    out_toi[r] = found ? best : 0.0f;
    out_tri[r] = best_id;
    if (WITH_NORMAL) {
        V3 n = mk3(0.f, 0.f, 0.f);
        uint32_t feat = PB2_INVALID_U32;
        if (found) {
            n = normalize3(best_n);
            if (best_fid & 2u) n = -n;
            if (pose7) n = iso_vec(pose, n);  // RayIntersection::transform_by (ray.rs:327-333)
            // ray_trimesh.rs:28-32: back face => Face(i + num_triangles)
            feat = (best_fid & 1u) ? best_id + nt : best_id;
        }
        if (out_normal) { out_normal[3ull * r] = n.x; out_normal[3ull * r + 1] = n.y; out_normal[3ull * r + 2] = n.z; }
        if (out_feature) out_feature[r] = feat;
    }
}


// ---------------------------------------------------------------------------------------------------------------
// Persistent variant (n_leaves >= 2): one lane = one ray at a time, warps pull new rays from a global counter as
// lanes retire (dynamic fetch), so a few long rays no longer hold 31 idle lanes hostage; leaf tests of both children
// share one call site so lanes that reach a leaf on different sides stay converged. Same arithmetic, same tie rule.
template <bool WITH_NORMAL>
__global__ void __launch_bounds__(128) k_raycast_trimesh_persistent(const NodeWide* __restrict__ nodes, const float4* __restrict__ tris,
                                  uint32_t nt, const float* __restrict__ pose7, const float* __restrict__ rays,
                                  const uint32_t* __restrict__ perm, uint32_t m, float max_toi, float* __restrict__ out_toi,
                                  uint32_t* __restrict__ out_tri, float* __restrict__ out_normal, uint32_t* __restrict__ out_feature,
                                  unsigned int* __restrict__ next_ray, int steps, int refill, uint32_t cull, unsigned int* fault) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    Iso7 pose;
    if (pose7) pose = load_iso(pose7);
    V3 o = mk3(0.f, 0.f, 0.f), d = o, inv = o, best_n = o;
    float best = 0.f;
    uint32_t best_id = PB2_INVALID_U32, best_fid = 0, r = 0, curr = PB2_INVALID_U32;
    bool found = false, active = false;
    uint32_t stack[PB2_STACK];
    int sp = 0;
    bool exhausted = false;
    for (;;) {
        __syncwarp();
        unsigned idle = __ballot_sync(FULL, !active);
        if (!exhausted && (idle == FULL || __popc(idle) >= refill)) {
            unsigned base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(next_ray, (unsigned)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base >= m) exhausted = true;
            if (!active) {
                uint32_t slot = base + __popc(idle & ((1u << lane) - 1u));
                if (slot < m) {
                    r = perm ? perm[slot] : slot;
                    o = mk3(rays[6ull * r], rays[6ull * r + 1], rays[6ull * r + 2]);
                    d = mk3(rays[6ull * r + 3], rays[6ull * r + 4], rays[6ull * r + 5]);
                    if (pose7) { o = iso_inv_point(pose, o); d = iso_inv_vec(pose, d); }
                    inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
                    best = max_toi; best_id = PB2_INVALID_U32; best_fid = 0; found = false;
                    curr = 0; sp = 0; active = true;
                }
            }
        }
        if (!__any_sync(FULL, active)) break;
#pragma unroll 1
        for (int it = 0; it < steps; ++it) {
            if (curr != PB2_INVALID_U32) {
                const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
                float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
                float ls = slab_cost_bf(l0, l1, o, inv, best);
                float rs = slab_cost_bf(r0, r1, o, inv, best);
                uint32_t lc = __float_as_uint(l0.w), rc = __float_as_uint(r0.w);
                bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u;
                bool rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
                bool sw = ls > rs;
                float s0 = sw ? rs : ls, s1 = sw ? ls : rs;
                uint32_t c0 = sw ? rc : lc, c1 = sw ? lc : rc;
                bool f0 = sw ? rleaf : lleaf, f1 = sw ? lleaf : rleaf;
                if (f0 | f1) {
                    // leaves first (near, then far), one call site
#pragma unroll 1
                    for (int k = 0; k < 2; ++k) {
                        float sc = k ? s1 : s0;
                        bool isleaf = k ? f1 : f0;
                        if (isleaf && sc != FLT_MAX && (sc < best || (found && sc == best))) {
                            uint32_t pos = k ? c1 : c0;
                            float4 ta = __ldg(&tris[3ull * pos]), tb = __ldg(&tris[3ull * pos + 1]), tc = __ldg(&tris[3ull * pos + 2]);
                            float toi; uint32_t fid; V3 n;
                            if (ray_triangle(mk3(ta.x, ta.y, ta.z), mk3(tb.x, tb.y, tb.z), mk3(tc.x, tc.y, tc.z), o, d, toi, fid, n) && toi <= best &&
                                (cull == 0u || (fid & 1u) == cull - 1u)) {
                                uint32_t id = __float_as_uint(ta.w);
                                if (toi < best || (found && toi == best && id < best_id)) {
                                    best = toi; best_id = id; best_fid = fid; found = true;
                                    if (WITH_NORMAL) best_n = n;
                                }
                            }
                        }
                    }
                }
                bool go0 = !f0 && s0 != FLT_MAX && (s0 < best || (found && s0 == best));
                bool go1 = !f1 && s1 != FLT_MAX && (s1 < best || (found && s1 == best));
                if (go0 && go1) pb2_push(stack, sp, c1, fault);
                uint32_t nxt = go0 ? c0 : c1;
                if (!(go0 || go1)) {
                    nxt = PB2_INVALID_U32;
                    if (sp > 0) nxt = stack[--sp];
                }
                curr = nxt;
            }
        }
        if (active && curr == PB2_INVALID_U32) {
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


// Coherence key for ray reordering (results are scattered back by ray index, so outputs are order-independent):
// [direction octant : 3][Morton of the origin quantised to 6 bits/axis over the root box : 18][direction quantised 3 bits/axis : 9].
__device__ __forceinline__ uint32_t spread3_10(uint32_t v) {
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__global__ void k_ray_keys(const NodeWide* __restrict__ nodes, const float* __restrict__ pose7, const float* __restrict__ rays, uint32_t m,
                           uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= m) return;
    V3 o = mk3(rays[6ull * r], rays[6ull * r + 1], rays[6ull * r + 2]);
    V3 d = mk3(rays[6ull * r + 3], rays[6ull * r + 4], rays[6ull * r + 5]);
    if (pose7) { Iso7 pose = load_iso(pose7); o = iso_inv_point(pose, o); d = iso_inv_vec(pose, d); }
    const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
    float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
    V3 mn = mk3(fminf(l0.x, r0.x), fminf(l0.y, r0.y), fminf(l0.z, r0.z));
    V3 mx = mk3(fmaxf(l1.x, r1.x), fmaxf(l1.y, r1.y), fmaxf(l1.z, r1.z));
    V3 ext = mx - mn;
    float e = fmaxf(fmaxf(ext.x, ext.y), fmaxf(ext.z, 1e-30f));
    // origins may lie outside the scene box: quantise over the box grown by its largest extent on each side
    float sc = 64.0f / (3.0f * e);
    uint32_t qx = (uint32_t)fminf(fmaxf((o.x - mn.x + e) * sc, 0.0f), 63.0f);
    uint32_t qy = (uint32_t)fminf(fmaxf((o.y - mn.y + e) * sc, 0.0f), 63.0f);
    uint32_t qz = (uint32_t)fminf(fmaxf((o.z - mn.z + e) * sc, 0.0f), 63.0f);
    float dl = fmaxf(fmaxf(fabsf(d.x), fabsf(d.y)), fmaxf(fabsf(d.z), 1e-30f));
    uint32_t dx = (uint32_t)fminf(fabsf(d.x) / dl * 7.999f, 7.0f), dy = (uint32_t)fminf(fabsf(d.y) / dl * 7.999f, 7.0f),
             dz = (uint32_t)fminf(fabsf(d.z) / dl * 7.999f, 7.0f);
    uint32_t oct = (d.x < 0.0f ? 1u : 0u) | (d.y < 0.0f ? 2u : 0u) | (d.z < 0.0f ? 4u : 0u);
    uint32_t om = spread3_10(qx) | (spread3_10(qy) << 1) | (spread3_10(qz) << 2);
    uint32_t dm = spread3_10(dx) | (spread3_10(dy) << 1) | (spread3_10(dz) << 2);
    keys[r] = (oct << 27) | ((om & 0x3ffffu) << 9) | (dm & 0x1ffu);
    vals[r] = r;
}

// ---------------------------------------------------------------------------------------------------------------
// PointQuery for TriMesh (query/point/point_composite_shape.rs:164-186, no pseudo-normals): CompositeShapeRef::project_local_point
// (:49-72) = Bvh::find_best with aabb cost Aabb::distance_to_local_point(pt, true) (point_aabb.rs:135-146) and leaf cost
// na::distance(projection on the triangle, pt) (point_triangle.rs:58-290). Same ordered descent and tie rule as the ray
// kernels: among bit-equal minimal distances (a point closest to a shared edge or vertex: one third of random points) the
// smallest triangle index wins, so nodes whose cost equals the best are still opened once a candidate exists.
__device__ __forceinline__ float aabb_point_dist(float4 lo, float4 hi, V3 p) {
    V3 shift = vmax3(vmax3(mk3(lo.x, lo.y, lo.z) - p, p - mk3(hi.x, hi.y, hi.z)), mk3(0.f, 0.f, 0.f));
    return nrm(shift);
}

#include "gjk.cuh"

__global__ void __launch_bounds__(128) k_project_points_trimesh(const NodeWide* __restrict__ nodes, const float4* __restrict__ tris, uint32_t n_leaves,
                                  const float* __restrict__ pose7, const float* __restrict__ points, uint32_t m,
                                  float* __restrict__ out_proj, uint8_t* __restrict__ out_inside, uint32_t* __restrict__ out_tri, unsigned int* fault) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    V3 p = mk3(points[3ull * k], points[3ull * k + 1], points[3ull * k + 2]);
    Iso7 pose;
    if (pose7) { pose = load_iso(pose7); p = iso_inv_point(pose, p); }
    float best = FLT_MAX;
    uint32_t best_id = PB2_INVALID_U32;
    V3 best_pt = mk3(0.f, 0.f, 0.f);
    bool best_in = false, found = false;
    auto leaf = [&](uint32_t pos) {
        float4 ta = __ldg(&tris[3ull * pos]), tb = __ldg(&tris[3ull * pos + 1]), tc = __ldg(&tris[3ull * pos + 2]);
        Proj pr;
        project_on_triangle(mk3(ta.x, ta.y, ta.z), mk3(tb.x, tb.y, tb.z), mk3(tc.x, tc.y, tc.z), p, pr);
        float d = nrm(p - pr.point);
        uint32_t id = __float_as_uint(ta.w);
        if (d < best || (found && d == best && id < best_id)) { best = d; best_id = id; best_pt = pr.point; best_in = pr.inside; found = true; }
    };
    if (n_leaves == 1) {
        const float4* np = reinterpret_cast<const float4*>(&nodes[0]);
        float4 l0 = __ldg(np), l1 = __ldg(np + 1);
        if (aabb_point_dist(l0, l1, p) < FLT_MAX) leaf(__float_as_uint(l0.w));
    } else if (n_leaves >= 2) {
        uint32_t stack[PB2_STACK];
        int sp = 0;
        uint32_t curr = 0;
        for (;;) {
            const float4* np = reinterpret_cast<const float4*>(&nodes[curr]);
            float4 l0 = __ldg(np), l1 = __ldg(np + 1), r0 = __ldg(np + 2), r1 = __ldg(np + 3);
            float ls = aabb_point_dist(l0, l1, p), rs = aabb_point_dist(r0, r1, p);
            uint32_t lc = __float_as_uint(l0.w), rc = __float_as_uint(r0.w);
            bool lleaf = (__float_as_uint(l1.w) & PB2_LEAF_COUNT_MASK) == 1u, rleaf = (__float_as_uint(r1.w) & PB2_LEAF_COUNT_MASK) == 1u;
            if (ls > rs) { float ts = ls; ls = rs; rs = ts; uint32_t tc = lc; lc = rc; rc = tc; bool tl = lleaf; lleaf = rleaf; rleaf = tl; }
            bool next = false;
            if (ls != FLT_MAX && (ls < best || (found && ls == best))) {
                if (lleaf) leaf(lc); else { curr = lc; next = true; }
            }
            if (rs != FLT_MAX && (rs < best || (found && rs == best))) {
                if (rleaf) leaf(rc);
                else if (next) pb2_push(stack, sp, rc, fault);
                else { curr = rc; next = true; }
            }
            if (!next) { if (sp == 0) break; curr = stack[--sp]; }
        }
    }
    if (found && pose7) best_pt = iso_point(pose, best_pt);  // PointProjection::transform_by
    out_proj[3ull * k] = best_pt.x; out_proj[3ull * k + 1] = best_pt.y; out_proj[3ull * k + 2] = best_pt.z;
    out_inside[k] = best_in ? 1 : 0;
    out_tri[k] = best_id;
}

// Enqueues the ray kernels for m device-resident rays on ctx->stream.
static int cast_rays_device(pb2_ctx* ctx, const pb2_trimesh* mesh, const void* d_pose, const void* d_rays, uint32_t m, float max_toi,
                            void* d_toi, void* d_tri, void* d_n, void* d_f, bool with_normal, uint32_t cull, const PieceSignal* pieces = nullptr) {
    const pb2_bvh* b = &mesh->bvh;
    // ray reordering pays off once the node array no longer fits in L2 (126 MB); below that the sort costs more than it saves
    // variants: 0 one thread per ray, 1 persistent binary, 2 = 1 + ray reordering, 3 compressed 8-wide tree, 4 = 3 + ray reordering,
    // 5 = 3 with the warp-shared triangle phase, 6 = 5 + ray reordering
    bool big = (size_t)b->n_nodes * sizeof(NodeWide) > (size_t)(100u << 20);
    int variant = mesh->n_nodes8 ? 5 : (big ? 2 : 1), steps = 16, refill = 8;
    if (variant >= 3) { steps = 8; refill = 4; }  // wide kernel: `steps` = lanes with queued triangles that trigger a triangle pass
    {   // tuning knobs (read per call; cheap)
        const char* e = getenv("PB2_RAY_VARIANT");
        if (e) variant = atoi(e);
        if ((e = getenv("PB2_RAY_STEPS"))) steps = atoi(e);
        if ((e = getenv("PB2_RAY_REFILL"))) refill = atoi(e);
    }
    if (pieces && (variant != 5 || b->n_leaves < 2 || m < 4096)) return PB2_ERR_INVALID;  // callers check pb2_can_signal_pieces first
    if (variant == 0 || b->n_leaves < 2 || m < 4096) {
        unsigned blocks = pb2_blocks(m, 128);
        if (with_normal)
            k_raycast_trimesh<true><<<blocks, 128, 0, ctx->stream>>>(b->nodes, mesh->tris, b->n_leaves, mesh->nt, (const float*)d_pose,
                                                                     (const float*)d_rays, m, max_toi, (float*)d_toi, (uint32_t*)d_tri,
                                                                     (float*)d_n, (uint32_t*)d_f, cull, PB2_FAULT_PTR(ctx));
        else
            k_raycast_trimesh<false><<<blocks, 128, 0, ctx->stream>>>(b->nodes, mesh->tris, b->n_leaves, mesh->nt, (const float*)d_pose,
                                                                      (const float*)d_rays, m, max_toi, (float*)d_toi, (uint32_t*)d_tri,
                                                                      nullptr, nullptr, cull, PB2_FAULT_PTR(ctx));
    } else {
        unsigned int* next_ray = (unsigned int*)(ctx->d_counters + ctx->ray_slot);
        PB2_CUDA(ctx, cudaMemsetAsync(next_ray, 0, 4, ctx->stream));
        const uint32_t* perm = nullptr;
        if (variant >= 3 && !mesh->n_nodes8) variant = 1 + (variant & 1 ? 0 : 1);
        if ((variant == 2 || variant == 4 || variant == 6) && m >= 65536) {
            size_t cub_bytes = 0;
            cub::DeviceRadixSort::SortPairs(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (const uint32_t*)nullptr,
                                            (uint32_t*)nullptr, (int)m, 0, 30, ctx->stream);
            size_t arr = ((size_t)m * 4 + 255) & ~(size_t)255;
            PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[0], 4 * arr + cub_bytes));
            char* base = (char*)ctx->scratch[0].ptr;
            uint32_t *k_in = (uint32_t*)base, *k_out = (uint32_t*)(base + arr), *v_in = (uint32_t*)(base + 2 * arr), *v_out = (uint32_t*)(base + 3 * arr);
            k_ray_keys<<<pb2_blocks(m, 256), 256, 0, ctx->stream>>>(b->nodes, (const float*)d_pose, (const float*)d_rays, m, k_in, v_in);
            PB2_LAUNCHED(ctx);
            PB2_CUDA(ctx, cub::DeviceRadixSort::SortPairs(base + 4 * arr, cub_bytes, (const uint32_t*)k_in, k_out, (const uint32_t*)v_in, v_out,
                                                          (int)m, 0, 30, ctx->stream));
            ctx->launches += 3;
            perm = v_out;
        }
        int per_sm = 0;
        if (with_normal) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_raycast_trimesh_persistent<true>, 128, 0);
        else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_raycast_trimesh_persistent<false>, 128, 0);
        if (per_sm < 1) per_sm = 1;
        unsigned blocks = (unsigned)(ctx->sm_count * per_sm);
        unsigned need = pb2_blocks(m, 128);
        if (blocks > need) blocks = need;
        if (variant >= 3)
            PB2_CHECK(pb2_wide_cast(ctx, mesh, (const float*)d_pose, (const float*)d_rays, perm, m, max_toi, (float*)d_toi, (uint32_t*)d_tri,
                                    (float*)d_n, (uint32_t*)d_f, with_normal, steps, refill, cull, variant >= 5, pieces));
        else if (with_normal)
            k_raycast_trimesh_persistent<true><<<blocks, 128, 0, ctx->stream>>>(b->nodes, mesh->tris, mesh->nt, (const float*)d_pose,
                (const float*)d_rays, perm, m, max_toi, (float*)d_toi, (uint32_t*)d_tri, (float*)d_n, (uint32_t*)d_f, next_ray, steps, refill, cull, PB2_FAULT_PTR(ctx));
        else
            k_raycast_trimesh_persistent<false><<<blocks, 128, 0, ctx->stream>>>(b->nodes, mesh->tris, mesh->nt, (const float*)d_pose,
                (const float*)d_rays, perm, m, max_toi, (float*)d_toi, (uint32_t*)d_tri, nullptr, nullptr, next_ray, steps, refill, cull, PB2_FAULT_PTR(ctx));
    }
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    return PB2_OK;
}

extern "C" {

int pb2_trimesh_create(pb2_ctx* ctx, const float* vertices, uint32_t nv, const uint32_t* indices, uint32_t nt, int mem,
                       pb2_trimesh** out) {
    if (!ctx || !out || !vertices || !indices) return PB2_ERR_INVALID;
    if (nt == 0 || nv == 0) PB2_FAIL(ctx, PB2_ERR_INVALID, "TriMeshBuilderError::EmptyIndices");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_v = nullptr, *d_i = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, vertices, (size_t)nv * 12, mem, &d_v));
    PB2_CHECK(pb2_stage_in(ctx, 1, indices, (size_t)nt * 12, mem, &d_i));
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[1], (size_t)nt * 24));
    float* aabbs = (float*)ctx->scratch[1].ptr;
    uint32_t* bad = (uint32_t*)ctx->d_counters;
    PB2_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
    k_triangle_aabbs<<<pb2_blocks(nt, 256), 256, 0, ctx->stream>>>((const float*)d_v, (const uint32_t*)d_i, nt, nv, aabbs, bad);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (*(uint32_t*)ctx->h_counters != 0) PB2_FAIL(ctx, PB2_ERR_INVALID, "triangle index out of bounds");

    pb2_trimesh* mesh = new pb2_trimesh();
    mesh->nt = nt; mesh->nv = nv;
    pb2_bvh* b = &mesh->bvh;
    // Link of the Morton-sorted triangles: Karras LBVH by default. The PLOC strategy (surface-area-driven clustering,
    // bvh_build.cu) is available with PB2_MESH_BUILD=1; measured on the two bench meshes it does not pay: both are regular
    // tessellations on which the cubic-cell Morton splits are already near-optimal (terrain 10.9 wide-node visits per ray with
    // the LBVH against 11.5 with PLOC radius 16, 1M-triangle sphere 13.8 against 15.5; gpurun_out/r2b).
    b->strategy = PB2_BUILD_BINNED;
    { const char* e = getenv("PB2_MESH_BUILD"); if (e) b->strategy = atoi(e) ? PB2_BUILD_PLOC : PB2_BUILD_BINNED; }
    b->n_leaves = nt;
    b->n_nodes = nt <= 2 ? 1 : nt - 1;
    b->cap_leaves = nt;
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->nodes, (size_t)b->n_nodes * sizeof(NodeWide));
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->parents, (size_t)b->n_nodes * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->counters, (size_t)b->n_nodes * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->leaf_slot, (size_t)nt * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&b->leaf_order, (size_t)nt * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&mesh->tris, (size_t)nt * 48);
    int s = PB2_OK;
    if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh alloc: %s", cudaGetErrorString(e)); s = PB2_ERR_CUDA; }
    if (s == PB2_OK) s = pb2_bvh_build_device(ctx, b, aabbs, nt, true, nullptr);
    if (s == PB2_OK) {
        k_gather_triangles<<<pb2_blocks(nt, 256), 256, 0, ctx->stream>>>((const float*)d_v, (const uint32_t*)d_i, nt, b->leaf_order, mesh->tris);
        PB2_LAUNCHED(ctx);
        e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh build: %s", cudaGetErrorString(e)); s = PB2_ERR_CUDA; }
    }
    if (s == PB2_OK) s = pb2_wide_build(ctx, mesh);
    if (s != PB2_OK) { pb2_trimesh_destroy(ctx, mesh); return s; }
    *out = mesh;
    return PB2_OK;
}

int pb2_trimesh_destroy(pb2_ctx* ctx, pb2_trimesh* mesh) {
    if (!ctx || !mesh) return PB2_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    pb2_bvh* b = &mesh->bvh;
    if (b->nodes) cudaFree(b->nodes);
    if (b->parents) cudaFree(b->parents);
    if (b->counters) cudaFree(b->counters);
    if (b->leaf_slot) cudaFree(b->leaf_slot);
    if (b->leaf_order) cudaFree(b->leaf_order);
    if (mesh->tris) cudaFree(mesh->tris);
    if (mesh->nodes8) cudaFree(mesh->nodes8);
    if (mesh->tris8) cudaFree(mesh->tris8);
    delete mesh;
    return PB2_OK;
}

const pb2_bvh* pb2_trimesh_bvh(const pb2_trimesh* mesh) { return mesh ? &mesh->bvh : nullptr; }

uint64_t pb2_trimesh_traversal_bytes(const pb2_trimesh* mesh) {
    if (!mesh) return 0;
    if (mesh->n_nodes8) return (uint64_t)mesh->n_nodes8 * 96ull + (uint64_t)mesh->nt * 64ull;   // W8_NODE_F4 / W8_TRI_F4 records (trimesh_wide.cu)
    return (uint64_t)mesh->bvh.n_nodes * 64ull + (uint64_t)mesh->nt * 48ull;
}

static int trimesh_cast_rays(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays, uint32_t m,
                             float max_toi, uint32_t cull, float* toi, uint32_t* tri, float* normal, uint32_t* feature, int mem) {
    if (!ctx || !mesh || (m && (!rays || !toi || !tri))) return PB2_ERR_INVALID;
    if (m == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    bool with_normal = normal || feature;
    if (mem == PB2_MEM_DEVICE)
        return cast_rays_device(ctx, mesh, pose7, rays, m, max_toi, toi, tri, normal, feature, with_normal, cull);
    // Host buffers: stage through HBM in chunks so that H2D copies, traversal and D2H copies of neighbouring chunks overlap
    // (full-duplex PCIe + compute; needs page-locked host memory to be truly asynchronous, pageable memory still works).
    void *d_rays = nullptr, *d_pose = nullptr, *d_toi = nullptr, *d_tri = nullptr, *d_n = nullptr, *d_f = nullptr;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[0], (size_t)m * 24));
    d_rays = ctx->stage[0].ptr;
    if (pose7) {
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[1], 28));
        d_pose = ctx->stage[1].ptr;
        PB2_CUDA(ctx, cudaMemcpyAsync(d_pose, pose7, 28, cudaMemcpyHostToDevice, ctx->stream));
    }
    PB2_CHECK(pb2_stage_out(ctx, 2, toi, (size_t)m * 4, mem, &d_toi));
    PB2_CHECK(pb2_stage_out(ctx, 3, tri, (size_t)m * 4, mem, &d_tri));
    PB2_CHECK(pb2_stage_out(ctx, 4, normal, (size_t)m * 12, mem, &d_n));
    PB2_CHECK(pb2_stage_out(ctx, 5, feature, (size_t)m * 4, mem, &d_f));
    PB2_CHECK(pb2_pipeline_init(ctx));
    // Chunks: H2D (24 B/ray at ~53 GB/s) runs at about the kernel's own rate, so chunk c + 1 is uploaded while chunk c is
    // traversed and chunk c - 1 is downloaded. Measured on 2^23 rays (chunk-size sweep): one chunk 8.3 ms; 2^20-ray chunks on
    // one compute stream 5.2 ms; 2^19-ray chunks alternating between two compute streams 4.5 ms.
    uint32_t sizes[64];
    int n_chunks = 0;
    const bool dual = mesh->n_nodes8 != 0 && getenv("PB2_RAY_VARIANT") == nullptr && getenv("PB2_RAY_SINGLE_STREAM") == nullptr;
    {
        uint32_t c = dual ? (1u << 19) : (1u << 20);
        const char* e = getenv("PB2_RAY_CHUNK_LOG2");
        if (e && atoi(e) >= 12 && atoi(e) <= 30) c = 1u << atoi(e);
        while ((uint64_t)c * 60 < m) c <<= 1;
        uint32_t rem = m;
        while (rem) { uint32_t k = rem < c ? rem : c; sizes[n_chunks++] = k; rem -= k; }
    }
    // two compute streams, alternating: the tail of one chunk's persistent launch (a few long rays) overlaps the start of the
    // next. Only for the wide-tree kernel, which needs no per-call scratch besides its fetch counter.
    cudaStream_t main_stream = ctx->stream;
    if (dual) {  // the second stream starts after everything already queued on the main one (pose upload, earlier calls)
        cudaEvent_t e0 = pb2_next_event(ctx);
        PB2_CUDA(ctx, cudaEventRecord(e0, main_stream));
        PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->compute2, e0, 0));
    }
    int rc = PB2_OK;
    uint32_t lo = 0;
    for (int ci = 0; ci < n_chunks && rc == PB2_OK; lo += sizes[ci], ++ci) {
        uint32_t cnt = sizes[ci];
        if (dual) { ctx->stream = (ci & 1) ? ctx->compute2 : main_stream; ctx->ray_slot = (ci & 1) ? 12 : 8; }
        rc = [&]() -> int {
        cudaEvent_t e_in = pb2_next_event(ctx), e_k = pb2_next_event(ctx);
        PB2_CUDA(ctx, cudaMemcpyAsync((char*)d_rays + (size_t)lo * 24, rays + (size_t)lo * 6, (size_t)cnt * 24, cudaMemcpyHostToDevice, ctx->copy_in));
        PB2_CUDA(ctx, cudaEventRecord(e_in, ctx->copy_in));
        PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e_in, 0));
        PB2_CHECK(cast_rays_device(ctx, mesh, d_pose, (char*)d_rays + (size_t)lo * 24, cnt, max_toi, (float*)d_toi + lo, (uint32_t*)d_tri + lo,
                                   d_n ? (float*)d_n + 3ull * lo : nullptr, d_f ? (uint32_t*)d_f + lo : nullptr, with_normal, cull));
        PB2_CUDA(ctx, cudaEventRecord(e_k, ctx->stream));
        PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, e_k, 0));
        PB2_CUDA(ctx, cudaMemcpyAsync(toi + lo, (float*)d_toi + lo, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
        PB2_CUDA(ctx, cudaMemcpyAsync(tri + lo, (uint32_t*)d_tri + lo, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
        if (normal) PB2_CUDA(ctx, cudaMemcpyAsync(normal + 3ull * lo, (float*)d_n + 3ull * lo, (size_t)cnt * 12, cudaMemcpyDeviceToHost, ctx->copy_out));
        if (feature) PB2_CUDA(ctx, cudaMemcpyAsync(feature + lo, (uint32_t*)d_f + lo, (size_t)cnt * 4, cudaMemcpyDeviceToHost, ctx->copy_out));
        return PB2_OK;
        }();
    }
    ctx->stream = main_stream; ctx->ray_slot = 8;
    if (rc != PB2_OK) { cudaStreamSynchronize(ctx->copy_out); cudaStreamSynchronize(main_stream); if (dual) cudaStreamSynchronize(ctx->compute2); return rc; }
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    if (dual) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->compute2));
    PB2_CHECK(pb2_fetch_fault(ctx));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return pb2_check_fault(ctx);
}

int pb2_trimesh_cast_rays(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays, uint32_t m,
                          float max_toi, int solid, float* toi, uint32_t* tri, float* normal, uint32_t* feature, int mem) {
    (void)solid;  // ignored by the 3D triangle test (ray_triangle.rs:53)
    return trimesh_cast_rays(ctx, mesh, pose7, rays, m, max_toi, 0u, toi, tri, normal, feature, mem);
}

int pb2_trimesh_cast_rays_with_culling(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays, uint32_t m,
                                       float max_toi, int culling, float* toi, uint32_t* tri, float* normal, uint32_t* feature, int mem) {
    if (culling != PB2_CULL_IGNORE_BACKFACES && culling != PB2_CULL_IGNORE_FRONTFACES) return PB2_ERR_INVALID;
    return trimesh_cast_rays(ctx, mesh, pose7, rays, m, max_toi, (uint32_t)culling, toi, tri, normal, feature, mem);
}

int pb2_trimesh_cast_rays_allgather(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* rays, uint32_t m, float max_toi,
                                    void* const* peer_toi, void* const* peer_tri, int n_peers, int self, uint64_t elem_offset, int chunks) {
    if (!ctx || !mesh || !peer_toi || !peer_tri || n_peers < 1 || self < 0 || self >= n_peers || (m && !rays)) return PB2_ERR_INVALID;
    if (m == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    PB2_CHECK(pb2_pipeline_init(ctx));
    if (chunks < 1) chunks = 1;
    if (chunks > 16) chunks = 16;
    float* my_toi = (float*)peer_toi[self] + elem_offset;
    uint32_t* my_tri = (uint32_t*)peer_tri[self] + elem_offset;
    cudaStream_t main_stream = ctx->stream;
    if (mesh->n_nodes8 != 0 && getenv("PB2_RAY_VARIANT") == nullptr && chunks > 1 && chunks <= 32 && m >= 4096 && mesh->bvh.n_leaves >= 2 &&
        ctx->wait_value32 != nullptr && getenv("PB2_RAY_PIECEWISE") == nullptr) {
        // One launch for the whole shard; the kernel publishes the completion of each range of results (PieceSignal), and the
        // copy-engine streams wait on those flags before pushing the range to the peers.
        typedef int (*wait_fn_t)(cudaStream_t, unsigned long long, unsigned int, unsigned int);   // CUresult cuStreamWaitValue32(CUstream, CUdeviceptr, cuuint32_t, unsigned)
        wait_fn_t wait_value = (wait_fn_t)ctx->wait_value32;
        PieceSignal ps;
        ps.size = (m + (uint32_t)chunks - 1u) / (uint32_t)chunks;
        ps.shift = 0;   // ranges are rounded up to a power of two so that the kernel finds a ray's range with a shift
        while ((1u << ps.shift) < ps.size && ps.shift < 31u) ps.shift++;
        ps.size = 1u << ps.shift;
        ps.done = ctx->d_pieces; ps.flag = ctx->d_pieces + 32;
        ps.flush_every = 32;  // refills between two publications of a warp's retired-ray counts
        { const char* e = getenv("PB2_RAY_PIECE_FLUSH"); if (e && atoi(e) > 0) ps.flush_every = (uint32_t)atoi(e); }
        PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_pieces, 0, 64 * sizeof(unsigned int), main_stream));
        cudaEvent_t e0 = pb2_next_event(ctx);
        PB2_CUDA(ctx, cudaEventRecord(e0, main_stream));
        for (int q = 0; q < 6; ++q) PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_peer[q], e0, 0));   // flags are reset before anybody waits on them
        ctx->ray_slot = 8;
        PB2_CHECK(cast_rays_device(ctx, mesh, pose7, rays, m, max_toi, my_toi, my_tri, nullptr, nullptr, false, 0u, &ps));
        const uint32_t n_pieces = (m + ps.size - 1u) / ps.size;
        cudaEvent_t e_done = pb2_next_event(ctx);
        PB2_CUDA(ctx, cudaEventRecord(e_done, main_stream));
        for (uint32_t ci = 0; ci < n_pieces; ++ci) {
            uint32_t lo = ci * ps.size, cnt = m - lo < ps.size ? m - lo : ps.size;
            bool waited[6] = {false, false, false, false, false, false};
            if (ci + 1 == n_pieces) {   // the last range completes with the kernel: an ordinary event wakes the copy streams sooner than a polled flag
                for (int q = 0; q < 6; ++q) { PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_peer[q], e_done, 0)); waited[q] = true; }
            }
            for (int p = 0; p < n_peers; ++p) {
                if (p == self) continue;
                int q = ((p - self + n_peers) % n_peers) % 6;
                cudaStream_t cs = ctx->copy_peer[q];
                if (!waited[q]) {
                    if (wait_value(cs, (unsigned long long)(uintptr_t)(ps.flag + ci), 1u, 0x0u /* CU_STREAM_WAIT_VALUE_GEQ */) != 0)
                        PB2_FAIL(ctx, PB2_ERR_CUDA, "cuStreamWaitValue32 failed");
                    waited[q] = true;
                }
                PB2_CUDA(ctx, cudaMemcpyAsync((float*)peer_toi[p] + elem_offset + lo, my_toi + lo, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs));
                PB2_CUDA(ctx, cudaMemcpyAsync((uint32_t*)peer_tri[p] + elem_offset + lo, my_tri + lo, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs));
            }
        }
        for (int q = 0; q < 6; ++q) { cudaEvent_t e = pb2_next_event(ctx); cudaEventRecord(e, ctx->copy_peer[q]); cudaStreamWaitEvent(main_stream, e, 0); }
        PB2_CUDA(ctx, cudaGetLastError());
        return PB2_OK;
    }
    const bool dual = mesh->n_nodes8 != 0 && getenv("PB2_RAY_VARIANT") == nullptr && chunks > 1;
    cudaEvent_t e0 = pb2_next_event(ctx);
    PB2_CUDA(ctx, cudaEventRecord(e0, main_stream));
    if (dual) PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->compute2, e0, 0));
    int rc = PB2_OK;
    uint32_t base = m / (uint32_t)chunks, rem = m % (uint32_t)chunks, lo = 0;
    for (int ci = 0; ci < chunks && rc == PB2_OK; ++ci) {
        uint32_t cnt = base + ((uint32_t)ci < rem ? 1u : 0u);
        if (cnt == 0) continue;
        if (dual) { ctx->stream = (ci & 1) ? ctx->compute2 : main_stream; ctx->ray_slot = (ci & 1) ? 12 : 8; }
        rc = [&]() -> int {
            PB2_CHECK(cast_rays_device(ctx, mesh, pose7, rays + 6ull * lo, cnt, max_toi, my_toi + lo, my_tri + lo, nullptr, nullptr, false, 0u));
            cudaEvent_t e_k = pb2_next_event(ctx);
            PB2_CUDA(ctx, cudaEventRecord(e_k, ctx->stream));
            // push this slice into every peer's gather buffers: device-to-device copies over NVLink on the copy engines (no SM is
            // taken from the persistent traversal kernel, which an NCCL kernel would have to wait for)
            for (int q = 0; q < 6; ++q) PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_peer[q], e_k, 0));
            for (int p = 0; p < n_peers; ++p) {
                if (p == self) continue;
                cudaStream_t cs = ctx->copy_peer[((p - self + n_peers) % n_peers) % 6];  // spread peers over the DMA queues
                PB2_CUDA(ctx, cudaMemcpyAsync((float*)peer_toi[p] + elem_offset + lo, my_toi + lo, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs));
                PB2_CUDA(ctx, cudaMemcpyAsync((uint32_t*)peer_tri[p] + elem_offset + lo, my_tri + lo, (size_t)cnt * 4, cudaMemcpyDeviceToDevice, cs));
            }
            return PB2_OK;
        }();
        lo += cnt;
    }
    ctx->stream = main_stream; ctx->ray_slot = 8;
    // everything this call enqueued joins the context's stream again
    for (int q = 0; q < 6; ++q) { cudaEvent_t e = pb2_next_event(ctx); cudaEventRecord(e, ctx->copy_peer[q]); cudaStreamWaitEvent(main_stream, e, 0); }
    if (dual) { cudaEvent_t e3 = pb2_next_event(ctx); cudaEventRecord(e3, ctx->compute2); cudaStreamWaitEvent(main_stream, e3, 0); }
    if (rc != PB2_OK) return rc;
    PB2_CUDA(ctx, cudaGetLastError());
    return PB2_OK;
}

int pb2_trimesh_project_points(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* pose7, const float* points, uint32_t m, int solid,
                               float* proj, uint8_t* inside, uint32_t* tri, int mem) {
    (void)solid;  // only matters for a degenerate triangle in 3D, where the device projection already returns the point itself
    if (!ctx || !mesh || (m && (!points || !proj || !inside || !tri))) return PB2_ERR_INVALID;
    if (m == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_pts, *d_pose = nullptr;
    void *d_proj, *d_in, *d_tri;
    PB2_CHECK(pb2_stage_in(ctx, 0, points, (size_t)m * 12, mem, &d_pts));
    if (pose7) PB2_CHECK(pb2_stage_in(ctx, 1, pose7, 28, mem, &d_pose));
    PB2_CHECK(pb2_stage_out(ctx, 2, proj, (size_t)m * 12, mem, &d_proj));
    PB2_CHECK(pb2_stage_out(ctx, 3, inside, (size_t)m, mem, &d_in));
    PB2_CHECK(pb2_stage_out(ctx, 4, tri, (size_t)m * 4, mem, &d_tri));
    k_project_points_trimesh<<<pb2_blocks(m, 128), 128, 0, ctx->stream>>>(mesh->bvh.nodes, mesh->tris, mesh->bvh.n_leaves, (const float*)d_pose,
                                                                         (const float*)d_pts, m, (float*)d_proj, (uint8_t*)d_in, (uint32_t*)d_tri, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, proj, d_proj, (size_t)m * 12, mem));
    PB2_CHECK(pb2_stage_back(ctx, inside, d_in, (size_t)m, mem));
    PB2_CHECK(pb2_stage_back(ctx, tri, d_tri, (size_t)m * 4, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}

}  // extern "C"

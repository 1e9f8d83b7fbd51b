// parry_b200 — typed shape table, per-collider AABBs and Bvh ray casts with Ball / Cuboid leaves.
//
// Replaces (reference, file:line): Ball::aabb (bounding_volume/aabb_ball.rs:8-33), Cuboid::aabb (aabb_cuboid.rs:9-16 +
// utils/isometry_ops.rs:16-18), ConvexPolyhedron::aabb (aabb_convex_polyhedron.rs:8-16 -> aabb_utils.rs:66-87);
// Bvh::cast_ray (bvh_queries.rs:260-271) with leaves = RayCast for Ball (query/ray/ray_ball.rs:8-98) and
// Cuboid (ray_cuboid.rs:6-25 -> ray_aabb.rs:12-92 -> query/clip/clip_aabb_line.rs:79-187).
#include "shapes.cuh"
#include <string.h>
#include <stdlib.h>
#include "gjk.cuh"
#include "traverse.cuh"

int pb2_stage_in(pb2_ctx* ctx, int slot, const void* src, size_t bytes, int mem, const void** out);
int pb2_stage_out(pb2_ctx* ctx, int slot, void* dst, size_t bytes, int mem, void** out);
int pb2_stage_back(pb2_ctx* ctx, void* dst, const void* dev, size_t bytes, int mem);

__global__ void k_compute_aabbs(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float* __restrict__ points,
                                uint32_t n_shapes, const uint32_t* __restrict__ shape_ids, const float* __restrict__ poses, uint32_t n,
                                float* __restrict__ out, uint32_t* __restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t sid = shape_ids ? shape_ids[i] : i;
    if (sid >= n_shapes) { atomicAdd(bad, 1u); return; }
    Iso7 pos = load_iso(poses + 7ull * i);
    V3 mn, mx;
    shape_aabb_dev(kinds[sid], params[sid], points, pos, mn, mx);
    float* o = out + 6ull * i;
    o[0] = mn.x; o[1] = mn.y; o[2] = mn.z; o[3] = mx.x; o[4] = mx.y; o[5] = mx.z;
}

// ---------------------------------------------------------------- single-shape ray casts (local space)
// ray_toi_with_ball (ray_ball.rs:33-77), ball at the origin of its local frame
__device__ __forceinline__ bool ray_ball(float radius, V3 o, V3 dir, bool solid, bool& inside, float& toi) {
    V3 dcenter = o - mk3(0.f, 0.f, 0.f);
    float a = nrm2(dir);
    float b = dot3(dcenter, dir);
    float c = nrm2(dcenter) - radius * radius;
    if (a == 0.0f) {
        if (c > 0.0f) { inside = false; return false; }
        inside = true; toi = 0.0f; return true;
    }
    if (c > 0.0f && b > 0.0f) { inside = false; return false; }
    float delta = b * b - a * c;
    if (delta < 0.0f) { inside = false; return false; }
    float t = (-b - sqrtf(delta)) / a;
    if (t <= 0.0f) {
        inside = true;
        toi = solid ? 0.0f : (-b + sqrtf(delta)) / a;
        return true;
    }
    inside = false; toi = t; return true;
}

// Aabb::cast_local_ray (ray_aabb.rs:12-49) on [-he, he]. Quirk: for a non-solid cast from inside, the reference returns
// tmax = min(max_toi, exit time), i.e. Some(max_toi) when the exit lies beyond max_toi. find_best never accepts that value
// (strict `<` against best == max_toi), so it is reported as a miss here — otherwise the tie rule could mistake it for a tie.
__device__ __forceinline__ bool ray_cuboid_toi(V3 he, V3 o, V3 d, float max_toi, bool solid, float& toi) {
    float tmin = 0.0f, tmax = max_toi, texit = FLT_MAX;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float di = comp(d, i), oi = comp(o, i), h = comp(he, i);
        if (di == 0.0f) {
            if (oi < -h || oi > h) return false;
        } else {
            float denom = 1.0f / di;
            float n = (-h - oi) * denom, f = (h - oi) * denom;
            if (n > f) { float t = n; n = f; f = t; }
            tmin = fmaxf(tmin, n);
            tmax = fminf(tmax, f);
            texit = fminf(texit, f);
            if (tmin > tmax) return false;
        }
    }
    if (tmin == 0.0f && !solid) {
        if (texit > max_toi) return false;
        toi = tmax;
    } else toi = tmin;
    return true;
}

struct ClipHit { float t; V3 n; int side; };
// clip_aabb_line (clip_aabb_line.rs:79-187) on [-he, he]
__device__ __forceinline__ bool clip_cuboid_line(V3 he, V3 o, V3 d, ClipHit& near, ClipHit& far) {
    float tmax = FLT_MAX, tmin = -FLT_MAX;
    int near_side = 0, far_side = 0;
    bool near_diag = false, far_diag = false;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float di = comp(d, i), oi = comp(o, i), h = comp(he, i);
        if (di == 0.0f) {
            if (oi < -h || oi > h) return false;
        } else {
            float denom = 1.0f / di;
            bool flip;
            float n = (-h - oi) * denom, f = (h - oi) * denom;
            if (n > f) { flip = true; float t = n; n = f; f = t; } else flip = false;
            if (n > tmin) { tmin = n; near_side = flip ? -(i + 1) : (i + 1); near_diag = false; }
            else if (n == tmin) near_diag = true;
            if (f < tmax) { tmax = f; far_side = !flip ? -(i + 1) : (i + 1); far_diag = false; }
            else if (f == tmax) far_diag = true;
            if (tmax < 0.0f || tmin > tmax) return false;
        }
    }
    bool contains = -he.x <= o.x && -he.y <= o.y && -he.z <= o.z && o.x <= he.x && o.y <= he.y && o.z <= he.z;
    ClipHit zero; zero.t = 0.0f; zero.n = mk3(0.f, 0.f, 0.f); zero.side = 0;
    if (near_diag) { near.t = tmin; near.n = -normalize3(d); near.side = near_side; }
    else {
        if (near_side == 0) { near = zero; far = zero; return contains; }
        float v[3] = {0.f, 0.f, 0.f};
        if (near_side < 0) v[-near_side - 1] = 1.0f; else v[near_side - 1] = -1.0f;
        near.t = tmin; near.n = mk3(v[0], v[1], v[2]); near.side = near_side;
    }
    if (far_diag) { far.t = tmax; far.n = -normalize3(d); far.side = far_side; }
    else {
        if (far_side == 0) { near = zero; far = zero; return contains; }
        float v[3] = {0.f, 0.f, 0.f};
        if (far_side < 0) v[-far_side - 1] = -1.0f; else v[far_side - 1] = 1.0f;
        far.t = tmax; far.n = mk3(v[0], v[1], v[2]); far.side = far_side;
    }
    return true;
}
// Aabb::cast_local_ray_and_get_normal (ray_aabb.rs:52-92)
__device__ __forceinline__ bool ray_cuboid_normal(V3 he, V3 o, V3 d, float max_toi, bool solid, float& toi, V3& n, uint32_t& feat) {
    ClipHit near, far, r;
    if (!clip_cuboid_line(he, o, d, near, far)) return false;
    if (near.t < 0.0f) {
        if (solid) { r.t = 0.0f; r.n = mk3(0.f, 0.f, 0.f); r.side = far.side; }
        else if (far.t <= max_toi) r = far;
        else return false;
    } else if (near.t <= max_toi) r = near;
    else return false;
    toi = r.t; n = r.n;
    feat = r.side < 0 ? (uint32_t)(-r.side) - 1u + 3u : (uint32_t)r.side - 1u;
    return true;
}

template <bool WITH_NORMAL>
__global__ void __launch_bounds__(128) k_raycast_shapes(const NodeWide* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t n_leaves,
                                 const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                                 const uint32_t* __restrict__ shape_ids, const float* __restrict__ poses, const float* __restrict__ rays,
                                 uint32_t m, float max_toi, bool solid, float* __restrict__ out_toi, uint32_t* __restrict__ out_leaf,
                                 float* __restrict__ out_normal, uint32_t* __restrict__ out_feature, uint32_t n_shapes, unsigned int* fault) {
    uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = r < m;   // every lane stays: the descent below is warp-cooperative (traverse.cuh)
    if (!valid) r = 0;
    V3 o = mk3(rays[6ull * r], rays[6ull * r + 1], rays[6ull * r + 2]);
    V3 d = mk3(rays[6ull * r + 3], rays[6ull * r + 4], rays[6ull * r + 5]);
    V3 inv = mk3(1.0f / d.x, 1.0f / d.y, 1.0f / d.z);
    float best = max_toi;
    bool found = false;
    uint32_t best_id = PB2_INVALID_U32, best_feat = PB2_INVALID_U32;
    V3 best_n = mk3(0.f, 0.f, 0.f);
    auto leaf = [&](uint32_t pos, unsigned) {
        uint32_t id = order[pos];
        uint32_t sid = shape_ids ? shape_ids[id] : id;
        if (sid >= n_shapes) { atomicOr(fault, PB2_FAULT_BAD_ID); return; }   // reported as PB2_ERR_INVALID at the next synchronisation
        Iso7 pose = load_iso(poses + 7ull * id);
        V3 lo = iso_inv_point(pose, o), ld = iso_inv_vec(pose, d);  // RayCast::cast_ray (ray.rs:381-390)
        float4 pr = params[sid];
        float toi; V3 n = mk3(0.f, 0.f, 0.f); uint32_t feat = 0;
        bool hit;
        if (kinds[sid] == PB2_SHAPE_BALL) {
            bool inside;
            hit = ray_ball(pr.x, lo, ld, solid, inside, toi);
            if (hit && WITH_NORMAL) {  // ray_toi_and_normal_with_ball (ray_ball.rs:81-98)
                V3 p = (lo + ld * toi) - mk3(0.f, 0.f, 0.f);
                n = normalize3(p);
                if (inside) n = -n;
            }
            hit = hit && toi <= best;
        } else if (kinds[sid] == PB2_SHAPE_CUBOID) {
            V3 he = mk3(pr.x, pr.y, pr.z);
            if (WITH_NORMAL) hit = ray_cuboid_normal(he, lo, ld, best, solid, toi, n, feat);
            else hit = ray_cuboid_toi(he, lo, ld, best, solid, toi);
        } else {
            // RayCast for ConvexPolyhedron (ray_support_map.rs:163-181): GJK ray cast, FeatureId::Unknown
            DShape g;
            g.kind = DS_CONVEX; g.he = mk3(0.f, 0.f, 0.f); g.pts = pts + __float_as_uint(pr.x); g.n = __float_as_uint(pr.y);
            hit = ray_support_map(g, lo, ld, best, solid, toi, n);
            feat = PB2_FEATURE_UNKNOWN;
        }
        if (!hit) return;
        if (toi < best || (found && toi == best && id < best_id)) {
            best = toi; best_id = id; found = true;
            if (WITH_NORMAL) { best_n = iso_vec(pose, n); best_feat = feat; }
        }
    };
    // BvhNode::cast_ray (bvh_tree.rs:1177-1181) as the node cost; the leaf query (a GJK ray cast for hull leaves) runs for all
    // waiting lanes together: 2^20 rays vs 2^20 mixed colliders 13.7 -> 8.7 ms (22.5 -> 9.9 ms with normals), same results
    auto cost = [&](float4 lo, float4 hi, float bound) { return slab_cost(lo.x, lo.y, lo.z, hi.x, hi.y, hi.z, o, d, inv, bound); };
    bvh_find_best_cost(0xffffffffu, valid, nodes, n_leaves, max_toi, best, found, cost, leaf, fault);
    if (!valid) return;
    out_toi[r] = found ? best : 0.0f;
    out_leaf[r] = best_id;
    if (WITH_NORMAL) {
        if (out_normal) { out_normal[3ull * r] = best_n.x; out_normal[3ull * r + 1] = best_n.y; out_normal[3ull * r + 2] = best_n.z; }
        if (out_feature) out_feature[r] = found ? best_feat : PB2_INVALID_U32;
    }
}

__global__ void k_pad_points(const float* __restrict__ p, uint32_t np, float4* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < np) out[i] = make_float4(p[3ull * i], p[3ull * i + 1], p[3ull * i + 2], 0.0f);
}

extern "C" {

int pb2_shapes_create(pb2_ctx* ctx, const uint8_t* kinds, const float* params, uint32_t n, const float* points, uint32_t np,
                      pb2_shapes** out) {
    if (!ctx || !out || (n && (!kinds || !params)) || (np && !points)) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    bool has_convex = false;
    for (uint32_t i = 0; i < n; ++i) {
        if (kinds[i] > PB2_SHAPE_CONVEX) PB2_FAIL(ctx, PB2_ERR_INVALID, "shape %u: unknown kind %u", i, kinds[i]);
        if (kinds[i] == PB2_SHAPE_CONVEX) {
            uint32_t first, cnt;
            memcpy(&first, params + 4 * i, 4);
            memcpy(&cnt, params + 4 * i + 1, 4);
            if (cnt == 0 || (uint64_t)first + cnt > np) PB2_FAIL(ctx, PB2_ERR_INVALID, "shape %u: bad point range", i);
            has_convex = true;
        }
    }
    pb2_shapes* s = new pb2_shapes();
    s->n = n; s->np = np; s->has_convex = has_convex;
    s->h_npoints = (uint32_t*)calloc(n ? n : 1, sizeof(uint32_t));
    for (uint32_t i = 0; i < n; ++i) if (kinds[i] == PB2_SHAPE_CONVEX) memcpy(&s->h_npoints[i], params + 4 * i + 1, 4);
    size_t nn = n ? n : 1, npp = np ? np : 1;
    cudaError_t e = cudaMalloc((void**)&s->kinds, nn);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->params, nn * 16);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->points, npp * 12);
    if (e == cudaSuccess) e = cudaMalloc((void**)&s->points4, npp * 16);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(s->kinds, kinds, n, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(s->params, params, (size_t)n * 16, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && np) e = cudaMemcpyAsync(s->points, points, (size_t)np * 12, cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess && np) {
        k_pad_points<<<pb2_blocks(np, 256), 256, 0, ctx->stream>>>(s->points, np, s->points4);
        PB2_LAUNCHED(ctx);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "shapes_create: %s", cudaGetErrorString(e));
        pb2_shapes_destroy(ctx, s);
        return PB2_ERR_CUDA;
    }
    *out = s;
    return PB2_OK;
}

int pb2_shapes_destroy(pb2_ctx* ctx, pb2_shapes* s) {
    if (!ctx || !s) return PB2_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    if (s->kinds) cudaFree(s->kinds);
    if (s->params) cudaFree(s->params);
    if (s->points) cudaFree(s->points);
    if (s->points4) cudaFree(s->points4);
    cudaFree(s->hull_face_first); cudaFree(s->hull_face_count); cudaFree(s->face_normal); cudaFree(s->face_first); cudaFree(s->face_count);
    cudaFree(s->verts_adj_to_face); cudaFree(s->edges_adj_to_face);
    cudaFree(s->vert_first); cudaFree(s->vert_count); cudaFree(s->faces_adj_to_vertex); cudaFree(s->edges_adj_to_vertex);
    cudaFree(s->hull_edge_first); cudaFree(s->edge_dir);
    free(s->h_npoints);
    delete s;
    return PB2_OK;
}

int pb2_shapes_set_hull_topology(pb2_ctx* ctx, pb2_shapes* s, const uint32_t* hull_face_first, const uint32_t* hull_face_count,
                                 const float* face_normal, const uint32_t* face_first, const uint32_t* face_count, uint32_t nf,
                                 const uint32_t* vertices_adj_to_face, const uint32_t* edges_adj_to_face, uint32_t nadj) {
    if (!ctx || !s || !hull_face_first || !hull_face_count || !face_normal || !face_first || !face_count || !vertices_adj_to_face ||
        !edges_adj_to_face || nf == 0 || nadj == 0)
        return PB2_ERR_INVALID;
    if (s->face_normal) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_topology: topology already set");
    for (uint32_t i = 0; i < s->n; ++i) {
        const uint32_t npts = s->h_npoints[i];
        if (npts == 0) continue;   // not a ConvexPolyhedron
        if (hull_face_count[i] == 0 || (uint64_t)hull_face_first[i] + hull_face_count[i] > nf)
            PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_topology: face range of a hull out of bounds");
        for (uint32_t f = hull_face_first[i]; f < hull_face_first[i] + hull_face_count[i]; ++f) {
            if (face_count[f] < 3 || (uint64_t)face_first[f] + face_count[f] > nadj)
                PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_topology: face with fewer than 3 vertices or adjacency range out of bounds");
            for (uint32_t k = face_first[f]; k < face_first[f] + face_count[f]; ++k)
                if (vertices_adj_to_face[k] >= npts) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_topology: vertex id out of range");
        }
    }
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    auto up = [&](const void* src, size_t bytes, void** dst) -> bool {
        return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    bool ok = up(hull_face_first, (size_t)s->n * 4, (void**)&s->hull_face_first) && up(hull_face_count, (size_t)s->n * 4, (void**)&s->hull_face_count) &&
              up(face_normal, (size_t)nf * 12, (void**)&s->face_normal) && up(face_first, (size_t)nf * 4, (void**)&s->face_first) &&
              up(face_count, (size_t)nf * 4, (void**)&s->face_count) && up(vertices_adj_to_face, (size_t)nadj * 4, (void**)&s->verts_adj_to_face) &&
              up(edges_adj_to_face, (size_t)nadj * 4, (void**)&s->edges_adj_to_face);
    if (!ok) PB2_FAIL(ctx, PB2_ERR_CUDA, "shapes_set_hull_topology: device allocation or upload failed");
    s->nf = nf; s->nadj = nadj;
    return PB2_OK;
}

int pb2_shapes_set_hull_vertex_topology(pb2_ctx* ctx, pb2_shapes* s, const uint32_t* vert_first, const uint32_t* vert_count,
                                        const uint32_t* faces_adj_to_vertex, const uint32_t* edges_adj_to_vertex, uint32_t nadj,
                                        const uint32_t* hull_edge_first, const float* edge_dir, uint32_t ne) {
    if (!ctx || !s || !vert_first || !vert_count || !faces_adj_to_vertex || !edges_adj_to_vertex || !hull_edge_first || !edge_dir || ne == 0 || nadj == 0)
        return PB2_ERR_INVALID;
    if (!s->face_normal) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_vertex_topology: set the face topology first");
    if (s->vert_first) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_vertex_topology: already set");
    for (uint32_t i = 0; i < s->np; ++i)
        if ((uint64_t)vert_first[i] + vert_count[i] > nadj) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_vertex_topology: adjacency range out of bounds");
    for (uint32_t i = 0; i < s->n; ++i)
        if (s->h_npoints[i] && hull_edge_first[i] >= ne) PB2_FAIL(ctx, PB2_ERR_INVALID, "shapes_set_hull_vertex_topology: edge offset out of bounds");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    auto up = [&](const void* src, size_t bytes, void** dst) -> bool {
        return cudaMalloc(dst, bytes) == cudaSuccess && cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
    };
    size_t npp = s->np ? s->np : 1;
    bool ok = up(vert_first, npp * 4, (void**)&s->vert_first) && up(vert_count, npp * 4, (void**)&s->vert_count) &&
              up(faces_adj_to_vertex, (size_t)nadj * 4, (void**)&s->faces_adj_to_vertex) && up(edges_adj_to_vertex, (size_t)nadj * 4, (void**)&s->edges_adj_to_vertex) &&
              up(hull_edge_first, (size_t)s->n * 4, (void**)&s->hull_edge_first) && up(edge_dir, (size_t)ne * 12, (void**)&s->edge_dir);
    if (!ok) PB2_FAIL(ctx, PB2_ERR_CUDA, "shapes_set_hull_vertex_topology: device allocation or upload failed");
    return PB2_OK;
}

int pb2_shapes_compute_aabbs(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape_ids, const float* poses7, uint32_t n,
                             float* aabbs, int mem) {
    if (!ctx || !shapes || (n && (!poses7 || !aabbs))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_ids = nullptr, *d_poses = nullptr;
    void* d_out = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape_ids, (size_t)n * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 1, poses7, (size_t)n * 28, mem, &d_poses));
    PB2_CHECK(pb2_stage_out(ctx, 2, aabbs, (size_t)n * 24, mem, &d_out));
    uint32_t* bad = (uint32_t*)(ctx->d_counters + 2);
    PB2_CUDA(ctx, cudaMemsetAsync(bad, 0, 4, ctx->stream));
    k_compute_aabbs<<<pb2_blocks(n, 128), 128, 0, ctx->stream>>>(shapes->kinds, shapes->params, shapes->points, shapes->n,
                                                                (const uint32_t*)d_ids, (const float*)d_poses, n, (float*)d_out, bad);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, aabbs, d_out, (size_t)n * 24, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 2, bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (*(uint32_t*)(ctx->h_counters + 2)) PB2_FAIL(ctx, PB2_ERR_INVALID, "compute_aabbs: shape id out of range");
    }
    return PB2_OK;
}

int pb2_bvh_cast_rays_shapes(pb2_ctx* ctx, const pb2_bvh* bvh, const pb2_shapes* shapes, const uint32_t* shape_ids,
                             const float* poses7, const float* rays, uint32_t m, float max_toi, int solid, float* toi, uint32_t* leaf,
                             float* normal, uint32_t* feature, int mem) {
    if (!ctx || !bvh || !shapes || !poses7 || (m && (!rays || !toi || !leaf))) return PB2_ERR_INVALID;
    if (m == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    uint32_t nl = bvh->n_leaves;
    const void *d_rays = nullptr, *d_ids = nullptr, *d_poses = nullptr;
    void *d_toi = nullptr, *d_leaf = nullptr, *d_n = nullptr, *d_f = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, rays, (size_t)m * 24, mem, &d_rays));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape_ids, (size_t)nl * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 6, poses7, (size_t)nl * 28, mem, &d_poses));
    PB2_CHECK(pb2_stage_out(ctx, 2, toi, (size_t)m * 4, mem, &d_toi));
    PB2_CHECK(pb2_stage_out(ctx, 3, leaf, (size_t)m * 4, mem, &d_leaf));
    PB2_CHECK(pb2_stage_out(ctx, 4, normal, (size_t)m * 12, mem, &d_n));
    PB2_CHECK(pb2_stage_out(ctx, 5, feature, (size_t)m * 4, mem, &d_f));
    unsigned blocks = pb2_blocks(m, 128);
    if (normal || feature)
        k_raycast_shapes<true><<<blocks, 128, 0, ctx->stream>>>(bvh->nodes, bvh->leaf_order, nl, shapes->kinds, shapes->params, shapes->points4,
                                                                (const uint32_t*)d_ids, (const float*)d_poses, (const float*)d_rays, m,
                                                                max_toi, solid != 0, (float*)d_toi, (uint32_t*)d_leaf, (float*)d_n,
                                                                (uint32_t*)d_f, shapes->n, PB2_FAULT_PTR(ctx));
    else
        k_raycast_shapes<false><<<blocks, 128, 0, ctx->stream>>>(bvh->nodes, bvh->leaf_order, nl, shapes->kinds, shapes->params, shapes->points4,
                                                                 (const uint32_t*)d_ids, (const float*)d_poses, (const float*)d_rays, m,
                                                                 max_toi, solid != 0, (float*)d_toi, (uint32_t*)d_leaf, nullptr, nullptr, shapes->n, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, toi, d_toi, (size_t)m * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, leaf, d_leaf, (size_t)m * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, normal, d_n, (size_t)m * 12, mem));
    PB2_CHECK(pb2_stage_back(ctx, feature, d_f, (size_t)m * 4, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}

}  // extern "C"

// parry_b200 — batched query::contact.
//
// Replaces (reference, file:line): query::contact (query/contact/contact_shape_shape.rs:123-138), Contact (contact.rs:71-105,
// flip :156, transform_by_mut :171), DefaultQueryDispatcher::contact (query/default_query_dispatcher.rs:302-356),
// contact_ball_ball (contact_ball_ball.rs:9-42), contact_ball_convex_polyhedron / contact_convex_polyhedron_ball
// (contact_ball_convex_polyhedron.rs:12-63) with Cuboid point projection (query/point/point_aabb.rs:9-132, point_cuboid.rs,
// shape/cuboid.rs:401-448) and ConvexPolyhedron point projection (query/point/point_support_map.rs:17-77),
// contact_support_map_support_map (contact_support_map_support_map.rs:10-77), EPA (query/epa/epa3.rs:18-675) with the
// sift rules of Rust's BinaryHeap, ccw_face_normal (utils/ccw_face_normal.rs:21-27), Triangle::is_affinely_dependent
// (shape/triangle.rs:534-540).
//
// B200 design: phase 1 (k_contact_gjk) runs one pair per thread: pose composition, dispatch, the closed forms and the
// GJK loop; pairs whose GJK ends in `Intersection` park their simplex in a compact queue (warp-aggregated append).
// Phase 2 (k_contact_epa2) is a persistent grid that pulls queued pairs and runs the expanding polytope with a bounded
// per-thread arena (vertices / faces / heap) as a flattened state machine, so that divergence of the long EPA runs
// does not stall the closed-form and separated pairs. Contacts are written either densely (status per pair) or
// compacted. The same two phases serve per-pair arrays, broad-phase pair lists (collider tables read through the list)
// and TriMesh-vs-shape candidates (shape 1 = a mesh triangle), see PairSrc.
#include "gjk.cuh"
#include "trimesh.cuh"
#include "manifold_update.cuh"
#include "compound_pair.cuh"
#include "traverse.cuh"
#include <stdlib.h>
#include <cub/cub.cuh>

int pb2_stage_in(pb2_ctx* ctx, int slot, const void* src, size_t bytes, int mem, const void** out);
int pb2_stage_out(pb2_ctx* ctx, int slot, void* dst, size_t bytes, int mem, void** out);
int pb2_stage_back(pb2_ctx* ctx, void* dst, const void* dev, size_t bytes, int mem);

enum { ST_NONE = 0, ST_SOME = 1, ST_UNSUPPORTED = 2, ST_NEEDS_HOST = 3 };


struct ContactOut {
    V3 p1, p2, n1, n2;
    float dist;
};

struct EpaJob {
    uint32_t pair;
    uint32_t dim;
    float o1[4][3];
    float o2[4][3];
};

// ------------------------------------------------------------------------------------------- closed forms
// contact_ball_ball.rs:9-42
__device__ __forceinline__ bool d_contact_ball_ball(const Iso7& pos12, float r1, float r2, float prediction, ContactOut& c) {
    V3 center2_1 = pos12.t;
    float d2 = nrm2(center2_1);
    float sum_radius = r1 + r2;
    float sre = sum_radius + prediction;
    if (d2 < sre * sre) {
        V3 normal1 = d2 != 0.0f ? normalize3(center2_1) : mk3(1.f, 0.f, 0.f);
        V3 normal2 = -iso_inv_vec(pos12, normal1);
        c.p1 = normal1 * r1; c.p2 = normal2 * r2; c.n1 = normal1; c.n2 = normal2;
        c.dist = sqrtf(d2) - sum_radius;
        return true;
    }
    return false;
}

struct Feat { int kind; uint32_t id; };  // 0 vertex 1 edge 2 face 3 unknown
__device__ __forceinline__ void setc(V3& v, int i, float x) { if (i == 0) v.x = x; else if (i == 1) v.y = x; else v.z = x; }

// point_aabb.rs:9-132 on [-he, he]
__device__ __forceinline__ void d_cuboid_project(V3 he, V3 pt, V3& proj, bool& inside, Feat& feat) {
    V3 mins = -he, maxs = he;
    V3 mins_pt = mins - pt, pt_maxs = pt - maxs;
    V3 zero = mk3(0.f, 0.f, 0.f);
    V3 shift = vmax3(mins_pt, zero) - vmax3(pt_maxs, zero);
    inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
    if (!inside) proj = pt + shift;
    else {
        float best = -FLT_MAX; bool is_mins = false; int best_id = 0;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            float a = comp(mins_pt, i), b = comp(pt_maxs, i);
            if (a < b) { if (b > best) { best_id = i; is_mins = false; best = b; } }
            else if (a > best) { best_id = i; is_mins = true; best = a; }
        }
        shift = zero;
        setc(shift, best_id, is_mins ? best : -best);
        proj = pt + shift;
    }
    int nzero = 0, last_zero = 0, last_not_zero = 0;
#pragma unroll
    for (int i = 0; i < 3; ++i) { if (comp(shift, i) == 0.0f) { nzero++; last_zero = i; } else last_not_zero = i; }
    V3 ctr = (mins + maxs) * 0.5f;
    if (nzero == 3) {
        feat.kind = 3; feat.id = 0;
        for (int i = 0; i < 3; ++i) {
            if (comp(proj, i) > comp(maxs, i) - PB2_EPS) { feat.kind = 2; feat.id = (uint32_t)i; return; }
            if (comp(proj, i) <= comp(mins, i) + PB2_EPS) { feat.kind = 2; feat.id = (uint32_t)(i + 3); return; }
        }
    } else if (nzero == 2) {
        feat.kind = 2;
        feat.id = comp(proj, last_not_zero) < comp(ctr, last_not_zero) ? (uint32_t)(last_not_zero + 3) : (uint32_t)last_not_zero;
    } else {
        uint32_t id = 0;
        for (int i = 0; i < 3; ++i) if (comp(proj, i) < comp(ctr, i)) id |= 1u << i;
        if (nzero == 0) { feat.kind = 0; feat.id = id; } else { feat.kind = 1; feat.id = (id << 2) | (uint32_t)last_zero; }
    }
}
// cuboid.rs:401-448
__device__ __forceinline__ bool d_cuboid_feature_normal(Feat f, V3& n) {
    V3 dir = mk3(0.f, 0.f, 0.f);
    if (f.kind == 2) { if (f.id < 3) setc(dir, (int)f.id, 1.0f); else setc(dir, (int)f.id - 3, -1.0f); n = dir; return true; }
    if (f.kind == 1) {
        uint32_t edge = f.id & 3u, face1 = (edge + 1) % 3, face2 = (edge + 2) % 3, signs = f.id >> 2;
        setc(dir, (int)face1, (signs & (1u << face1)) ? -1.0f : 1.0f);
        setc(dir, (int)face2, (signs & (1u << face2)) ? -1.0f : 1.0f);
        n = normalize3(dir); return true;
    }
    if (f.kind == 0) {
        for (int i = 0; i < 3; ++i) setc(dir, i, (f.id & (1u << i)) ? -1.0f : 1.0f);
        n = normalize3(dir); return true;
    }
    return false;
}

__device__ __forceinline__ V3 cub_support(V3 he, V3 d) { return mk3(copysignf(he.x, d.x), copysignf(he.y, d.y), copysignf(he.z, d.z)); }

// sat_cuboid_cuboid.rs:79-110
__device__ __forceinline__ void sat_normal_oneway(V3 he1, V3 he2, const Iso7& pos12, float& best_sep, V3& best_dir) {
    best_sep = -FLT_MAX; best_dir = mk3(0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        float sign = copysignf(1.0f, comp(pos12.t, i));
        V3 axis1 = mk3(0.f, 0.f, 0.f);
        setc(axis1, i, sign);
        V3 axis2 = iso_inv_vec(pos12, -axis1);
        V3 pt2 = iso_point(pos12, cub_support(he2, axis2));
        float sep = comp(pt2, i) * sign - comp(he1, i);
        if (sep > best_sep) { best_sep = sep; best_dir = axis1; }
    }
}
// sat_cuboid_cuboid.rs:5-77
__device__ __forceinline__ void sat_edge_twoway(V3 he1, V3 he2, const Iso7& pos12, float& best_sep, V3& best_dir) {
    best_sep = -FLT_MAX; best_dir = mk3(0.f, 0.f, 0.f);
    V3 c2[3] = {iso_vec(pos12, mk3(1.f, 0.f, 0.f)), iso_vec(pos12, mk3(0.f, 1.f, 0.f)), iso_vec(pos12, mk3(0.f, 0.f, 1.f))};
    for (int k = 0; k < 9; ++k) {
        V3 u = c2[k / 3];
        int a = k % 3;
        V3 axis = a == 0 ? mk3(0.0f, -u.z, u.y) : (a == 1 ? mk3(u.z, 0.0f, -u.x) : mk3(-u.y, u.x, 0.0f));
        float n = nrm(axis);
        if (n > PB2_EPS) {
            V3 ax = axis / n;
            float signum = copysignf(1.0f, dot3(pos12.t, ax));
            V3 axis1 = ax * signum;
            V3 axis2 = iso_inv_vec(pos12, -axis1);
            V3 lp1 = cub_support(he1, axis1);
            V3 pt2 = iso_point(pos12, cub_support(he2, axis2));
            float sep = dot3(pt2 - lp1, axis1);
            if (sep > best_sep) { best_sep = sep; best_dir = axis1; }
        }
    }
}

// approx::ulps_eq! defaults for f32 (epsilon = f32::EPSILON, max_ulps = 4)
__device__ __forceinline__ bool ulps_eq4(float a, float b) {
    if (fabsf(a - b) <= PB2_EPS) return true;
    if (signbit(a) != signbit(b)) return false;
    long long d = (long long)__float_as_int(a) - (long long)__float_as_int(b);
    if (d < 0) d = -d;
    return d <= 4;
}

// ---- cuboid-cuboid arms of query::distance / query::intersection_test (default_query_dispatcher.rs:183-186, :205-207)
// Cuboid::local_support_edge_segment (shape/cuboid.rs:249-263)
__device__ __forceinline__ void cub_support_edge(V3 he, V3 dir, V3& a, V3& b) {
    int i = 0; float best = fabsf(dir.x);   // nalgebra iamin: first strict minimum of |x|
    if (fabsf(dir.y) < best) { best = fabsf(dir.y); i = 1; }
    if (fabsf(dir.z) < best) i = 2;
    int j = (i + 1) % 3, k = (i + 2) % 3;
    a = mk3(0.f, 0.f, 0.f);
    setc(a, i, comp(he, i)); setc(a, j, copysignf(comp(he, j), comp(dir, j))); setc(a, k, copysignf(comp(he, k), comp(dir, k)));
    b = a; setc(b, i, -comp(he, i));
}
__device__ __forceinline__ float na_clamp01(float v) { return v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v); }
// closest_points_segment_segment_with_locations_nD (closest_points_segment_segment.rs:36-107)
__device__ __forceinline__ void seg_seg_params(V3 a1, V3 b1, V3 a2, V3 b2, float& s, float& t) {
    V3 d1 = b1 - a1, d2 = b2 - a2, r = a1 - a2;
    float a = nrm2(d1), e = nrm2(d2), f = dot3(d2, r);
    const float eps = PB2_EPS;
    if (a <= eps && e <= eps) { s = 0.0f; t = 0.0f; }
    else if (a <= eps) { s = 0.0f; t = na_clamp01(f / e); }
    else {
        float c = dot3(d1, r);
        if (e <= eps) { t = 0.0f; s = na_clamp01(-c / a); }
        else {
            float b = dot3(d1, d2), ae = a * e, bb = b * b, denom = ae - bb;
            if (denom > eps && !ulps_eq4(ae, bb)) s = na_clamp01((b * f - c * e) / denom);
            else s = 0.0f;
            t = (b * s + f) / e;
            if (t < 0.0f) { t = 0.0f; s = na_clamp01(-c / a); }
            else if (t > 1.0f) { t = 1.0f; s = na_clamp01((b - c) / a); }
        }
    }
}
__device__ __forceinline__ V3 seg_point_at(V3 a, V3 b, float s) {   // Segment::point_at of the location built from s (:88-104, segment.rs:410-419)
    if (s == 0.0f) return a;
    if (s == 1.0f) return b;
    return a * (1.0f - s) + b * s;
}
__device__ __forceinline__ V3 cuboid_project_solid(V3 he, V3 pt) {   // Aabb::project_local_point(pt, true) (point_aabb.rs:9-60)
    V3 zero = mk3(0.f, 0.f, 0.f);
    V3 shift = vmax3((-he) - pt, zero) - vmax3(pt - he, zero);
    bool inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
    return inside ? pt : pt + shift;
}
// intersection_test_cuboid_cuboid (intersection_test_cuboid_cuboid.rs:6-31)
__device__ __forceinline__ bool d_intersection_test_cuboid_cuboid(const Iso7& pos12, V3 he1, V3 he2) {
    float sep; V3 dir;
    sat_normal_oneway(he1, he2, pos12, sep, dir);
    if (sep > 0.0f) return false;
    Iso7 pos21 = iso_inverse(pos12);
    sat_normal_oneway(he2, he1, pos21, sep, dir);
    if (sep > 0.0f) return false;
    sat_edge_twoway(he1, he2, pos12, sep, dir);
    return sep <= 0.0f;
}
// distance_cuboid_cuboid (distance_cuboid_cuboid.rs:6-12) = closest_points_cuboid_cuboid with margin = f32::MAX
// (closest_points_cuboid_cuboid.rs:6-84); WithinMargin(p1, p2) -> na::distance(p1, pos12 * p2), anything else -> 0.
__device__ __forceinline__ float d_distance_cuboid_cuboid(const Iso7& pos12, V3 he1, V3 he2) {
    const float margin = FLT_MAX;
    Iso7 pos21 = iso_inverse(pos12);
    float s1, s2, s3; V3 d1, d2, d3;
    sat_normal_oneway(he1, he2, pos12, s1, d1);
    if (s1 > margin) return 0.0f;
    sat_normal_oneway(he2, he1, pos21, s2, d2);
    if (s2 > margin) return 0.0f;
    sat_edge_twoway(he1, he2, pos12, s3, d3);
    if (s3 > margin) return 0.0f;
    if (s1 <= 0.0f && s2 <= 0.0f && s3 <= 0.0f) return 0.0f;
    if (s1 >= s2 && s1 >= s3) {
        V3 pt2_1 = iso_point(pos12, cub_support(he2, iso_inv_vec(pos12, -d1)));
        V3 proj = cuboid_project_solid(he1, pt2_1);
        if (nrm2(proj - pt2_1) > margin * margin) return 0.0f;
        V3 p2 = iso_point(pos21, pt2_1);
        return nrm(proj - iso_point(pos12, p2));
    }
    if (s2 >= s1 && s2 >= s3) {
        V3 pt1_2 = iso_point(pos21, cub_support(he1, iso_inv_vec(pos21, -d2)));
        V3 proj = cuboid_project_solid(he2, pt1_2);
        if (nrm2(proj - pt1_2) > margin * margin) return 0.0f;
        V3 p1 = iso_point(pos12, pt1_2);
        return nrm(p1 - iso_point(pos12, proj));
    }
    V3 a1, b1, a2, b2;
    cub_support_edge(he1, d3, a1, b1);
    cub_support_edge(he2, iso_vec(pos21, -d3), a2, b2);
    float s, t;
    seg_seg_params(a1, b1, iso_point(pos12, a2), iso_point(pos12, b2), s, t);
    V3 p1 = seg_point_at(a1, b1, s), p2w = iso_point(pos12, seg_point_at(a2, b2, t));
    if (nrm2(p1 - p2w) <= margin * margin) return nrm(p1 - p2w);
    return 0.0f;
}

// Tail of contact_convex_polyhedron_ball (contact_ball_convex_polyhedron.rs:34-62) once the projection is known.
__device__ __forceinline__ V3 v3of4(float4 f) { return mk3(f.x, f.y, f.z); }
__device__ __forceinline__ int d_convex_ball_finish(const Iso7& pos12, bool is_cuboid, Feat f1, V3 proj, bool inside, float radius2,
                                                    float prediction, ContactOut& c, const float4* tri = nullptr) {
    V3 center2_1 = pos12.t;
    float dist; V3 normal1, dir1; float len;
    if (try_normalize_get(proj - center2_1, PB2_EPS, dir1, len)) {
        if (inside) { dist = -len - radius2; normal1 = dir1; }
        else { dist = len - radius2; normal1 = -dir1; }
    } else {
        dist = -radius2;
        float n;
        if (tri) {  // Triangle::normal() = Unit::try_new(scaled_normal, eps)
            V3 ta = v3of4(tri[0]), tb = v3of4(tri[1]), tc = v3of4(tri[2]);
            if (!try_normalize_get(cross3(tb - ta, tc - ta), PB2_EPS, normal1, n)) {
                if (!try_normalize_get(proj, PB2_EPS, normal1, n)) normal1 = mk3(0.f, 1.f, 0.f);
            }
        } else {
        // a hull's feature here is always FeatureId::Unknown (point_support_map.rs:62-77 normalises the same vector with the same
        // epsilon and fails like the test above), so ConvexPolyhedron::feature_normal is None and the fall-backs apply
        if (!is_cuboid || !d_cuboid_feature_normal(f1, normal1)) {
            if (!try_normalize_get(proj, PB2_EPS, normal1, n)) normal1 = mk3(0.f, 1.f, 0.f);
        }
        }
    }
    if (dist <= prediction) {
        V3 normal2 = iso_inv_vec(pos12, -normal1);
        c.p2 = normal2 * radius2; c.p1 = proj; c.n1 = normal1; c.n2 = normal2; c.dist = dist;
        return ST_SOME;
    }
    return ST_NONE;
}

__device__ __forceinline__ void flip_contact(ContactOut& c) {
    V3 t = c.p1; c.p1 = c.p2; c.p2 = t;
    t = c.n1; c.n1 = c.n2; c.n2 = t;
}

__device__ __forceinline__ DShape make_dshape(uint8_t kind, float4 pr, const float4* pts) {
    DShape s;
    s.kind = kind == PB2_SHAPE_CUBOID ? DS_CUBOID : DS_CONVEX;
    s.he = mk3(pr.x, pr.y, pr.z);
    s.pts = pts + __float_as_uint(pr.x);
    s.n = __float_as_uint(pr.y);
    return s;
}
__device__ __forceinline__ DShape origin_dshape() { DShape s; s.kind = DS_ORIGIN; s.he = mk3(0.f, 0.f, 0.f); s.pts = nullptr; s.n = 0; return s; }

// Per-pair setup shared by both phases: everything is recomputed from the inputs with identical arithmetic, so the
// EPA phase only needs the parked simplex.
struct PairSetup {
    Iso7 pos1, pos2, pos12;
    uint8_t k1, k2;
    float4 pr1, pr2;
    // GJK problem (when mode != 0): gpos12, g1, g2
    int mode;  // 0 closed form / nothing, 1 support-map pair, 2 convex(shape1)-ball(shape2), 3 ball(shape1)-convex(shape2)
    Iso7 gpos12;
    Iso7 cb_pos12;  // pos12 seen by contact_convex_polyhedron_ball (mode 2: pos12, mode 3: pos12.inverse())
    DShape g1, g2;
    const float4* tri;  // shape 1 is this TriMesh triangle (k1 == PB2_SHAPE_TRIANGLE_INTERNAL), else NULL
    const float4* tri2; // shape 2 is this triangle (support-map arm only: TriMesh-vs-TriMesh shape casts), else NULL
    bool local_frames;
};

// Where the pairs come from. Plain mode: pair k = (shape1[k] at pos1[k], shape2[k] at pos2[k]). `ab` (optional): pairs as
// index couples (2 per pair, e.g. straight from pb2_bvh_self_pairs): shapes and poses are then per-collider tables read
// through it. `mesh_tris` (optional, with `ab`): shape 1 of pair k is the TriMesh triangle stored at mesh_tris[3 * ab[2k]]
// (three float4, .w of the first = triangle id) at the single pose pos1[0]; shape 2 is shape2[ab[2k+1]] at pos2[ab[2k+1]].
struct PairSrc {
    const uint32_t* shape1;
    const uint32_t* shape2;
    const float* pos1;
    const float* pos2;
    const uint32_t* ab;
    const float4* mesh_tris;
    const float4* mesh_tris2 = nullptr;   // optional: shape 2 of pair k is the triangle at mesh_tris2[3 * ab[2k+1]] (shape2 is not read)
    uint32_t n_first, n_second;   // index bounds for ab[2k] / ab[2k+1]
    uint32_t flags = 0;           // PAIR_* below
    const float* init_dir = nullptr;   // optional GJK seed of the support-map arm (contact_manifolds_pfm_pfm.rs:66: last frame's
                                       // manifold.local_n1): init_dir[init_stride * i2 .. +3], used when its norm exceeds eps
    uint32_t init_stride = 3;
    const float* part_pose = nullptr;  // Compound mode (with `ab` = {part, pair}): shape 1 of candidate k is part ab[2k] (shape
                                       // shape1[part] at part_pose[part] inside the compound posed at pos1[pair]); shape 2 is
                                       // shape2[pair] at pos2[pair]. Implies local frames (of the part and of shape 2).
};
#define PAIR_SUPPORT_MAPS_ONLY 1u  // no ball arms: balls take part in GJK/EPA through their support map (what
                                   // cast_shapes_support_map_support_map calls: contact_support_map_support_map on any pair)
#define PAIR_LOCAL_FRAMES 2u       // leave the contact in the shapes' local frames (no Contact::transform_by_mut)
#define PAIR_POS12_GIVEN 8u        // pos2[k] already is pos12 (QueryDispatcher::contact's own argument); pos1 is not read. With LOCAL_FRAMES.
#define PAIR_COMPOUND_SECOND 4u    // Compound mode: the user's call was contact(shape, compound): pose12 = inv_mul(pos2, pos1).inverse()
#define PB2_SHAPE_TRIANGLE_INTERNAL 3   // a TriMesh part (shape::Triangle), never in a shape table

__device__ __forceinline__ void pair_setup(const uint8_t* kinds, const float4* params, const float4* pts, const PairSrc& src, uint32_t k,
                                           PairSetup& ps) {
    uint32_t i1 = k, i2 = k;
    if (src.ab) { i1 = src.ab[2ull * k]; i2 = src.ab[2ull * k + 1]; }
    ps.tri2 = nullptr;
    if (src.mesh_tris2) {
        ps.tri2 = src.mesh_tris2 + 3ull * i2;
        ps.k2 = PB2_SHAPE_TRIANGLE_INTERNAL;
        ps.pr2 = make_float4(0.f, 0.f, 0.f, 0.f);
    } else {
        uint32_t s2 = src.shape2[i2];
        ps.k2 = kinds[s2];
        ps.pr2 = params[s2];
    }
    ps.pos2 = load_iso(src.pos2 + 7ull * i2);
    const float4* tri = nullptr;
    if (src.mesh_tris) {
        tri = src.mesh_tris + 3ull * i1;
        ps.k1 = PB2_SHAPE_TRIANGLE_INTERNAL;
        ps.pr1 = make_float4(0.f, 0.f, 0.f, 0.f);
        ps.pos1 = load_iso(src.pos1);
    } else {
        uint32_t s1 = src.shape1[i1];
        ps.k1 = kinds[s1];
        ps.pr1 = params[s1];
        if (src.flags & PAIR_POS12_GIVEN) { ps.pos1.q.i = ps.pos1.q.j = ps.pos1.q.k = 0.f; ps.pos1.q.w = 1.f; ps.pos1.t = mk3(0.f, 0.f, 0.f); }
        else ps.pos1 = load_iso((src.part_pose ? src.part_pose : src.pos1) + 7ull * i1);
    }
    ps.tri = tri;
    ps.pos12 = (src.flags & PAIR_POS12_GIVEN) ? ps.pos2 : iso_inv_mul(ps.pos1, ps.pos2);  // contact_shape_shape.rs:130
    if (src.part_pose) {
        // contact_composite_shape_shape.rs:27: dispatcher.contact(&part_pos1.inv_mul(pose12), part1, shape2, ..); pose12 is the
        // user's pos12, or its inverse when the compound is the second shape (contact_shape_composite_shape, :74)
        Iso7 pc = load_iso(src.pos1 + 7ull * i2);
        Iso7 pose12 = (src.flags & PAIR_COMPOUND_SECOND) ? iso_inverse(iso_inv_mul(ps.pos2, pc)) : iso_inv_mul(pc, ps.pos2);
        ps.pos12 = iso_inv_mul(ps.pos1, pose12);
    }
    ps.mode = 0;
    ps.local_frames = (src.flags & PAIR_LOCAL_FRAMES) != 0 || src.part_pose != nullptr;
    bool b1 = ps.k1 == PB2_SHAPE_BALL, b2 = ps.k2 == PB2_SHAPE_BALL;
    if (src.flags & PAIR_SUPPORT_MAPS_ONLY) {
        ps.mode = 1;
        ps.gpos12 = ps.pos12;
        if (tri) { ps.g1.kind = DS_TRIANGLE; ps.g1.he = mk3(0.f, 0.f, 0.f); ps.g1.pts = tri; ps.g1.n = 3; }
        else ps.g1 = make_dshape(ps.k1, ps.pr1, pts);
        if (ps.tri2) { ps.g2.kind = DS_TRIANGLE; ps.g2.he = mk3(0.f, 0.f, 0.f); ps.g2.pts = ps.tri2; ps.g2.n = 3; }
        else ps.g2 = make_dshape(ps.k2, ps.pr2, pts);
        if (b1) ps.g1.kind = DS_BALL;
        if (b2) ps.g2.kind = DS_BALL;
    } else if (!b1 && !b2) {
        ps.mode = 1;
        ps.gpos12 = ps.pos12;
        if (tri) { ps.g1.kind = DS_TRIANGLE; ps.g1.he = mk3(0.f, 0.f, 0.f); ps.g1.pts = tri; ps.g1.n = 3; }
        else ps.g1 = make_dshape(ps.k1, ps.pr1, pts);
        if (ps.tri2) { ps.g2.kind = DS_TRIANGLE; ps.g2.he = mk3(0.f, 0.f, 0.f); ps.g2.pts = ps.tri2; ps.g2.n = 3; }
        else ps.g2 = make_dshape(ps.k2, ps.pr2, pts);
    } else if (b1 != b2) {
        bool convex_first = b2;
        ps.cb_pos12 = convex_first ? ps.pos12 : iso_inverse(ps.pos12);
        uint8_t kc = convex_first ? ps.k1 : ps.k2;
        if (kc == PB2_SHAPE_CONVEX) {
            ps.mode = convex_first ? 2 : 3;
            // local_point_projection_on_support_map (point_support_map.rs:17-33): m = Isometry(-point), gjk runs with m.inverse()
            V3 point = ps.cb_pos12.t;
            Iso7 m; m.q.i = 0.f; m.q.j = 0.f; m.q.k = 0.f; m.q.w = 1.f; m.t = -point;
            ps.gpos12 = iso_inverse(m);
            ps.g1 = make_dshape(kc, convex_first ? ps.pr1 : ps.pr2, pts);
            ps.g2 = origin_dshape();
        }
    }
}

// Epilogue for a GJK/EPA result (p1, p2_1, n1 in shape-1 space) -> local-frame contact or projection-based contact.
__device__ __forceinline__ int finish_gjk_pair(const PairSetup& ps, bool from_epa, V3 p1, V3 p2_1, V3 n1, float prediction, ContactOut& c) {
    if (ps.mode == 1) {
        // contact_support_map_support_map.rs:21-27
        c.dist = dot3(p2_1 - p1, n1);
        c.p1 = p1;
        c.p2 = iso_inv_point(ps.pos12, p2_1);
        c.n1 = n1;
        c.n2 = iso_inv_vec(ps.pos12, -n1);
        return ST_SOME;
    }
    // modes 2/3: p1 is the projection of the ball centre on the hull; inside iff it came from EPA
    Feat f; f.kind = 3; f.id = 0;
    float radius = ps.mode == 2 ? ps.pr2.x : ps.pr1.x;
    int st = d_convex_ball_finish(ps.cb_pos12, false, f, p1, from_epa, radius, prediction, c);
    if (st == ST_SOME && ps.mode == 3) flip_contact(c);
    return st;
}

__device__ __forceinline__ void to_world(const PairSetup& ps, ContactOut& c) {  // Contact::transform_by_mut
    if (ps.local_frames) return;
    c.p1 = iso_point(ps.pos1, c.p1);
    c.p2 = iso_point(ps.pos2, c.p2);
    c.n1 = iso_vec(ps.pos1, c.n1);
    c.n2 = iso_vec(ps.pos2, c.n2);
}

__device__ __forceinline__ unsigned long long warp_append1(unsigned long long* counter) {
    unsigned mask = __activemask();
    int lane = threadIdx.x & 31;
    int leader = __ffs(mask) - 1;
    unsigned long long base = 0;
    if (lane == leader) base = atomicAdd(counter, (unsigned long long)__popc(mask));
    base = __shfl_sync(mask, base, leader);
    return base + __popc(mask & ((1u << lane) - 1u));
}

struct OutSinks {
    float* dense;        // n x 13 or NULL
    uint8_t* status;     // n or NULL
    float* compact;      // cap x 13 or NULL
    uint32_t* pair_index;
    unsigned long long cap;
    unsigned long long* compact_count;
    unsigned long long* some_count;
    float* noint_dir = nullptr;   // n x 3 or NULL: the last search direction of support-map pairs GJK answered NoIntersection for
};

__device__ __forceinline__ void store_contact(float* o, const ContactOut& c) {
    o[0] = c.p1.x; o[1] = c.p1.y; o[2] = c.p1.z; o[3] = c.p2.x; o[4] = c.p2.y; o[5] = c.p2.z;
    o[6] = c.n1.x; o[7] = c.n1.y; o[8] = c.n1.z; o[9] = c.n2.x; o[10] = c.n2.y; o[11] = c.n2.z; o[12] = c.dist;
}

__device__ __forceinline__ void emit(const OutSinks& out, uint32_t k, int st, const ContactOut& c) {
    if (out.status) out.status[k] = (uint8_t)st;
    if (out.dense) {
        float* o = out.dense + 13ull * k;
        if (st == ST_SOME) store_contact(o, c);
        else { for (int i = 0; i < 13; ++i) o[i] = 0.0f; }
    }
    if (st == ST_SOME) {
        if (out.compact) {
            unsigned long long at = warp_append1(out.compact_count);
            if (at < out.cap) { store_contact(out.compact + 13ull * at, c); out.pair_index[at] = k; }
        } else if (out.some_count) {
            (void)warp_append1(out.some_count);
        }
    }
}

// ------------------------------------------------------------------------------------------- phase 1
// Pairs without a GJK problem (PairSetup::mode == 0): ball-ball and ball <-> cuboid / TriMesh triangle closed forms.
__device__ __forceinline__ int closed_form_pair(const PairSetup& ps, float prediction, ContactOut& c) {
    int st;
    bool b1 = ps.k1 == PB2_SHAPE_BALL, b2 = ps.k2 == PB2_SHAPE_BALL;
    if (b1 && b2) st = d_contact_ball_ball(ps.pos12, ps.pr1.x, ps.pr2.x, prediction, c) ? ST_SOME : ST_NONE;
    else {
        bool convex_first = b2;
        float4 prc = convex_first ? ps.pr1 : ps.pr2;
        float radius = convex_first ? ps.pr2.x : ps.pr1.x;
        V3 proj; bool inside; Feat f;
        const float4* t = convex_first ? ps.tri : ps.tri2;   // the non-ball side is a triangle
        if (t) {
            // PointQuery for Triangle (point_triangle.rs:27-47: location with solid = true); the feature normal of a
            // triangle is its normal whatever the feature (shape.rs:919-928, triangle.rs:226-228)
            V3 ta = v3of4(t[0]), tb = v3of4(t[1]), tc = v3of4(t[2]);
            Proj pr;
            project_on_triangle(ta, tb, tc, ps.cb_pos12.t, pr);
            f.kind = 4; f.id = 0;
            st = d_convex_ball_finish(ps.cb_pos12, false, f, pr.point, pr.inside, radius, prediction, c, t);
        } else {
            d_cuboid_project(mk3(prc.x, prc.y, prc.z), ps.cb_pos12.t, proj, inside, f);
            st = d_convex_ball_finish(ps.cb_pos12, true, f, proj, inside, radius, prediction, c);
        }
        if (st == ST_SOME && !convex_first) flip_contact(c);
    }
    return st;
}

// First simplex of the pair's GJK problem.
__device__ __forceinline__ void gjk_start(const PairSetup& ps, Simplex& s, const PairSrc& src, uint32_t k) {
    V3 dir; float nn;
    if (ps.mode == 1) {
        // contact_support_map_support_map_with_params (contact_support_map_support_map.rs:40-61): init_dir if the caller has one
        // (Unit::try_new(manifold.local_n1, eps), contact_manifolds_pfm_pfm.rs:66), else the direction of pos12's translation, else +x
        bool seeded = false;
        if (src.init_dir) {
            uint32_t i2 = src.ab ? src.ab[2ull * k + 1] : k;
            const float* q = src.init_dir + (size_t)src.init_stride * i2;
            seeded = try_normalize_get(mk3(q[0], q[1], q[2]), PB2_EPS, dir, nn);
        }
        if (!seeded && !try_normalize_get(ps.pos12.t, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
        sx_reset(s, cso_from_shapes(ps.gpos12, ps.g1, ps.g2, dir));
    } else {
        // point_support_map.rs:26-31: dir = normalize(point) or +x; support with m_inv = Isometry(point)
        V3 point = ps.cb_pos12.t;
        if (!try_normalize_get(point, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
        Iso7 m_inv; m_inv.q.i = 0.f; m_inv.q.j = 0.f; m_inv.q.k = 0.f; m_inv.q.w = 1.f; m_inv.t = point;
        sx_reset(s, cso_from_shapes(m_inv, ps.g1, ps.g2, dir));
    }
}

// GJK answered Intersection: the simplex goes to the EPA queue.
__device__ __forceinline__ void park_epa_job(EpaJob* __restrict__ jobs, unsigned long long* job_count, uint32_t k, const Simplex& s) {
    unsigned long long at = warp_append1(job_count);
    EpaJob* j = &jobs[at];
    j->pair = k; j->dim = (uint32_t)s.dim;
    for (int i = 0; i <= s.dim; ++i) {
        j->o1[i][0] = s.v[i].o1.x; j->o1[i][1] = s.v[i].o1.y; j->o1[i][2] = s.v[i].o1.z;
        j->o2[i][0] = s.v[i].o2.x; j->o2[i][1] = s.v[i].o2.y; j->o2[i][2] = s.v[i].o2.z;
    }
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_contact_gjk(const uint8_t* __restrict__ kinds, const float4* __restrict__ params,
                              const float4* __restrict__ pts, uint32_t n_shapes, PairSrc src, float prediction, uint32_t n, OutSinks out,
                              EpaJob* __restrict__ jobs, unsigned long long* job_count) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    ContactOut c;
    {
        uint32_t i1 = k, i2 = k;
        if (src.ab) { i1 = src.ab[2ull * k]; i2 = src.ab[2ull * k + 1]; }
        bool bad = src.ab && (i1 >= src.n_first || i2 >= src.n_second);
        if (!bad) bad = (!src.mesh_tris2 && src.shape2[i2] >= n_shapes) || (!src.mesh_tris && src.shape1[i1] >= n_shapes);
        if (bad) { emit(out, k, ST_UNSUPPORTED, c); return; }
    }
    PairSetup ps;
    pair_setup(kinds, params, pts, src, k, ps);
    int st;
    if (ps.mode == 0) {
        st = closed_form_pair(ps, prediction, c);
    } else {
        Simplex s;
        gjk_start(ps, s, src, k);
        V3 p1, p2, n1;
        float max_dist = ps.mode == 1 ? prediction : FLT_MAX;
        int r = gjk_closest_points(ps.gpos12, ps.g1, ps.g2, max_dist, s, p1, p2, n1);
        if (r == GJK_INTERSECTION) {
            park_epa_job(jobs, job_count, k, s);
            return;  // finished by phase 2
        } else if (r == GJK_CLOSEST_POINTS) {
            st = finish_gjk_pair(ps, false, p1, p2, n1, prediction, c);
        } else {
            st = ST_NONE;
            // GJKResult::NoIntersection(dir): the pfm_pfm manifold arm caches dir for next frame's GJK (contact_manifolds_pfm_pfm.rs:151-154)
            if (out.noint_dir && ps.mode == 1) { float* q = out.noint_dir + 3ull * k; q[0] = n1.x; q[1] = n1.y; q[2] = n1.z; }
        }
    }
    if (st == ST_SOME) to_world(ps, c);
    emit(out, k, st, c);
}


// Phase 1, persistent form: warps pull pairs from a counter and every trip runs ONE GJK iteration per lane, so a pair that
// needs ten iterations no longer holds the lanes of pairs that needed three (k_contact_gjk runs at 15.9 of 32 active
// lanes on the 4M hull-pair config). Lanes whose GJK has answered wait as "pending"; once `refill` lanes are pending or
// idle, one pass builds the pending lanes' contacts (or parks their simplex for EPA) and hands all of them new pairs —
// so the epilogue and the pair setup also run with many lanes. Same arithmetic as gjk_closest_points, flattened.
enum { G_IDLE = 0, G_RUN = 1, G_PEND = 2 };
template <int MINB>
__global__ void __launch_bounds__(128, MINB) k_contact_gjk_persistent(const uint8_t* __restrict__ kinds, const float4* __restrict__ params,
                              const float4* __restrict__ pts, uint32_t n_shapes, PairSrc src, float prediction, uint32_t n, OutSinks out,
                              EpaJob* __restrict__ jobs, unsigned long long* job_count, unsigned long long* __restrict__ next_pair, int refill) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    const float eps_tol = PB2_GJK_EPS_TOL;
    const float eps_rel = sqrtf(eps_tol);
    int state = G_IDLE, res = GJK_NO_INTERSECTION, niter = 0;
    bool wit_prev = false, exhausted = false;
    uint32_t k = 0;
    Simplex s;
    Iso7 gpos12;
    DShape g1, g2;
    g1 = origin_dshape(); g2 = origin_dshape();
    gpos12.q.i = gpos12.q.j = gpos12.q.k = 0.f; gpos12.q.w = 1.f; gpos12.t = mk3(0.f, 0.f, 0.f);
    V3 proj = mk3(0.f, 0.f, 0.f), old_dir = proj, res_dir = proj;
    float max_bound = FLT_MAX, max_dist = 0.0f;
    for (;;) {
        unsigned running = __ballot_sync(FULL, state == G_RUN);
        unsigned pend = __ballot_sync(FULL, state == G_PEND);
        if ((32 - __popc(running) >= refill || running == 0u) && (pend != 0u || !exhausted)) {
            if (state == G_PEND) {
                if (res == GJK_INTERSECTION) {
                    park_epa_job(jobs, job_count, k, s);
                } else {
                    PairSetup ps;
                    pair_setup(kinds, params, pts, src, k, ps);
                    ContactOut c;
                    int st = ST_NONE;
                    if (res == GJK_CLOSEST_POINTS) {
                        V3 p1, p2;
                        gjk_witness(s, wit_prev, p1, p2);
                        st = finish_gjk_pair(ps, false, p1, p2, res_dir, prediction, c);
                    }
                    if (st == ST_SOME) to_world(ps, c);
                    emit(out, k, st, c);
                }
                state = G_IDLE;
            }
            if (!exhausted) {
                unsigned idle = __ballot_sync(FULL, state == G_IDLE);
                unsigned long long base = 0;
                int leader = __ffs(idle) - 1;
                if (lane == leader) base = atomicAdd(next_pair, (unsigned long long)__popc(idle));
                base = __shfl_sync(FULL, base, leader);
                if (base + __popc(idle) >= n) exhausted = true;
                unsigned long long slot = base + __popc(idle & ((1u << lane) - 1u));
                if (state == G_IDLE && slot < n) {
                    k = (uint32_t)slot;
                    ContactOut c;
                    uint32_t i1 = k, i2 = k;
                    if (src.ab) { i1 = src.ab[2ull * k]; i2 = src.ab[2ull * k + 1]; }
                    bool bad = src.ab && (i1 >= src.n_first || i2 >= src.n_second);
                    if (!bad) bad = (!src.mesh_tris2 && src.shape2[i2] >= n_shapes) || (!src.mesh_tris && src.shape1[i1] >= n_shapes);
                    if (bad) {
                        emit(out, k, ST_UNSUPPORTED, c);
                    } else {
                        PairSetup ps;
                        pair_setup(kinds, params, pts, src, k, ps);
                        if (ps.mode == 0) {
                            int st = closed_form_pair(ps, prediction, c);
                            if (st == ST_SOME) to_world(ps, c);
                            emit(out, k, st, c);
                        } else {
                            gjk_start(ps, s, src, k);
                            gpos12 = ps.gpos12; g1 = ps.g1; g2 = ps.g2;
                            max_dist = ps.mode == 1 ? prediction : FLT_MAX;
                            proj = sx_project_origin_and_reduce(s);
                            V3 pd; float nn;
                            if (try_normalize_get(proj, 0.0f, pd, nn)) { old_dir = -pd; state = G_RUN; max_bound = FLT_MAX; niter = 0; }
                            else { res = GJK_INTERSECTION; state = G_PEND; }
                        }
                    }
                }
            }
            running = __ballot_sync(FULL, state == G_RUN);
            pend = __ballot_sync(FULL, state == G_PEND);
        }
        if (running == 0u && pend == 0u && exhausted) break;
        if (state == G_RUN) {  // one trip of the loop in gjk::closest_points (gjk.rs:371-447)
            float old_max_bound = max_bound;
            V3 dir; float dist;
            if (!try_normalize_get(-proj, eps_tol, dir, dist)) { res = GJK_INTERSECTION; state = G_PEND; }
            else {
                max_bound = dist;
                if (max_bound >= old_max_bound) { res = GJK_CLOSEST_POINTS; wit_prev = true; res_dir = old_dir; state = G_PEND; }
                else {
                    CSO cso = cso_from_shapes(gpos12, g1, g2, dir);
                    float min_bound = -dot3(dir, cso.point);
                    if (min_bound > max_dist) { res = GJK_NO_INTERSECTION; state = G_PEND; }
                    else if (max_bound - min_bound <= eps_rel * max_bound || !sx_add_point(s, cso)) {
                        res = GJK_CLOSEST_POINTS; wit_prev = false; res_dir = dir; state = G_PEND;
                    } else {
                        old_dir = dir;
                        proj = sx_project_origin_and_reduce(s);
                        if (s.dim == 3) {
                            if (min_bound >= eps_tol) { res = GJK_CLOSEST_POINTS; wit_prev = true; res_dir = old_dir; }
                            else res = GJK_INTERSECTION;
                            state = G_PEND;
                        } else {
                            niter += 1;
                            if (niter == 100) { res = GJK_NO_INTERSECTION; state = G_PEND; }
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------- phase 2, flattened
// Warp-convergent EPA: every lane owns one parked pair; ONE loop whose trip is "one expansion step" for every lane, with
// job fetch / initial-polytope construction folded into the same phases (pop -> support -> test -> silhouette -> faces ->
// finish). Lanes refill individually, so a long run never holds finished lanes idle (the first version waited at the end
// of the per-job loop: 8.6 of 32 lanes active). Hot arena per lane: faces 16 B (normal + packed pts/deleted), adjacency
// 8 B, heap 8 B, vertex point 16 B; witness points (orig1/orig2) live in a cold array and the barycentric coordinates of
// the winning face are recomputed at the end by the same deterministic projection.
#define E2_MAX_VERTS 112
#define E2_MAX_FACES 256
#define E2_MAX_SIL 64
#define E2_STACK 96
#ifndef E2_STACK_SMEM
#define E2_STACK_SMEM 64   // DFS stack entries kept in shared memory by k_contact_epa2 (deeper walks: status 3)
#endif
#ifndef E2_HEAP_SMEM
#define E2_HEAP_SMEM 16    // heap entries per thread kept in shared memory (sweep: see epa_persistent.inc / DESIGN.md 5.2)
#endif
#ifndef E2_MINB
#define E2_MINB 4          // resident CTAs per SM the EPA kernel is compiled for (register cap 65536 / (128 * E2_MINB))
#endif

// One polytope face = one 32-byte, 32-byte-aligned record read with a single 256-bit load (round 2; round 1 kept `face` and `adj` in
// two arrays: two dependent round trips to L2 wherever both were needed — the kernel is bound by exactly those,
// profiles/r2_contacts_heap16_full.json: long-scoreboard 8.5 cycles per issue at 4 warps per scheduler).
struct __align__(32) EFace {
    float4 f;     // normal.xyz ; w = pts0 | pts1 << 8 | pts2 << 16 | deleted << 24
    uint2 adj;    // x = adj0 | adj1 << 16 ; y = adj2
    uint2 pad;
};
__device__ __forceinline__ EFace ld_face(const EFace* p) {
    EFace r;
    float a0, a1, a2, a3, a4, a5, a6, a7;
    asm volatile("ld.global.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7) : "l"(p) : "memory");
    r.f = make_float4(a0, a1, a2, a3);
    r.adj = make_uint2(__float_as_uint(a4), __float_as_uint(a5));
    r.pad = make_uint2(0u, 0u);
    return r;
}
__device__ __forceinline__ void st_face(EFace* p, float4 f, uint2 adj) {
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" :: "l"(p), "f"(f.x), "f"(f.y), "f"(f.z), "f"(f.w), "f"(__uint_as_float(adj.x)),
                 "f"(__uint_as_float(adj.y)), "f"(0.0f), "f"(0.0f) : "memory");
}
// struct Epa2Arena: in epa_persistent.inc (one per configuration)

__device__ __forceinline__ uint32_t f_pts(float4 f, int i) { return (__float_as_uint(f.w) >> (8 * i)) & 0xffu; }
__device__ __forceinline__ bool f_deleted(float4 f) { return (__float_as_uint(f.w) >> 24) != 0u; }
__device__ __forceinline__ uint32_t a_get(uint2 a, int i) { return i == 0 ? (a.x & 0xffffu) : (i == 1 ? (a.x >> 16) : a.y); }
__device__ __forceinline__ void a_set(uint2& a, int i, uint32_t v) {
    if (i == 0) a.x = (a.x & 0xffff0000u) | v; else if (i == 1) a.x = (a.x & 0xffffu) | (v << 16); else a.y = v;
}
__device__ __forceinline__ V3 v3of(float4 f) { return mk3(f.x, f.y, f.z); }
__device__ __forceinline__ int e2_next_ccw(float4 f, uint32_t id) {
    if (f_pts(f, 0) == id) return 1;
    if (f_pts(f, 1) == id) return 2;
    return 0;
}
__device__ __forceinline__ bool h2_le(float a, float b) { return !(a > b); }
// Heap2 and its sift rules: epa_persistent.inc
// Barycentric coordinates Face::new would store for (va, vb, vc) + whether the origin projects inside.
__device__ __forceinline__ bool e2_face_bc(V3 va, V3 vb, V3 vc, float bc[3]) {
    Proj p;
    project_on_triangle(va, vb, vc, mk3(0.f, 0.f, 0.f), p);
    bc[0] = bc[1] = bc[2] = 0.f;
    if (p.kind == 0) { bc[p.idx] = 1.0f; }
    else if (p.kind == 1) { int i0 = p.idx == 1 ? 1 : 0, i1 = p.idx == 0 ? 1 : 2; bc[i0] = p.bc[0]; bc[i1] = p.bc[1]; }
    if (p.kind == 0 || p.kind == 1) {
        const float eps_tol = PB2_EPS * 100.0f;
        return p.inside || nrm2(p.point - mk3(0.f, 0.f, 0.f)) < eps_tol * eps_tol;
    }
    if (p.kind == 2) { bc[0] = p.bc[0]; bc[1] = p.bc[1]; bc[2] = p.bc[2]; return true; }
    return false;
}

// The boolean e2_face_bc returns, without the barycentric coordinates: same comparisons on the same intermediates as
// project_on_triangle (point_triangle.rs:58-290) with pt = origin, but the projected point (one IEEE division) is only
// formed for the vertex / edge regions where the answer depends on it. Face::new runs once per new polytope face
// (epa3.rs:96-118) and only needs this flag there; the coordinates are recomputed for the single winning face.
__device__ __forceinline__ bool e2_face_inside(V3 a, V3 b, V3 c) {
    const V3 pt = mk3(0.f, 0.f, 0.f);
    V3 ab = b - a, ac = c - a, ap = pt - a;
    float ab_ap = dot3(ab, ap), ac_ap = dot3(ac, ap);
    V3 bp = pt - b;
    float ab_bp = dot3(ab, bp), ac_bp = dot3(ac, bp);
    V3 cp = pt - c;
    float ab_cp = dot3(ab, cp), ac_cp = dot3(ac, cp);
    V3 bc = c - b;
    V3 n = cross3(ab, ac);
    float vc = dot3(n, cross3(ab, ap));
    float vb = -dot3(n, cross3(ac, cp));
    float va = dot3(n, cross3(bc, bp));
    bool rA = ab_ap <= 0.0f && ac_ap <= 0.0f;
    bool rB = !rA && (ab_bp >= 0.0f && ac_bp <= ab_bp);
    bool rC = !rA && !rB && (ac_cp >= 0.0f && ab_cp <= ac_cp);
    bool vtx = rA || rB || rC;
    bool rAB = !vtx && (vc < 0.0f && ab_ap >= 0.0f && ab_bp <= 0.0f);
    bool rAC = !vtx && !rAB && (vb < 0.0f && ac_ap >= 0.0f && ac_cp <= 0.0f);
    bool rBC = !vtx && !rAB && !rAC && (va < 0.0f && ac_bp - ab_bp >= 0.0f && ab_cp - ac_cp >= 0.0f);
    bool edge = rAB || rAC || rBC;
    if (!vtx && !edge) return (va + vb + vc) != 0.0f;  // face region: inside; degenerate (kind 3): not
    V3 point;
    if (vtx) point = rA ? a : (rB ? b : c);
    else {
        float num = rAB ? ab_ap : (rAC ? ac_ap : dot3(bc, bp));
        float den = rAB ? nrm2(ab) : (rAC ? nrm2(ac) : nrm2(bc));
        float q = num / den;
        V3 base = rBC ? b : a;
        V3 dirv = rAB ? ab : (rAC ? ac : bc);
        point = base + dirv * q;
    }
    const float eps_tol = PB2_EPS * 100.0f;
    return rel_eq3(point, pt) || nrm2(point - pt) < eps_tol * eps_tol;
}

enum { E2_IDLE = 0, E2_INIT = 1, E2_RUN = 2 };
#ifdef PB2_EPA_DEBUG
__device__ unsigned long long g_epa_dbg[8];
#endif
enum { FIN_NOT = 0, FIN_FACE = 1, FIN_NONE = 2, FIN_OVERFLOW = 3, FIN_DIM0 = 4 };

// EPA outcome -> contact (contact_support_map_support_map.rs:54-86 after gjk/epa): shared by the EPA kernels' phase F.
__device__ __forceinline__ int epa_result_to_contact(const PairSetup& ps, int fin, V3 p1, V3 p2, V3 n1, float prediction, ContactOut& c) {
    int st;
    if (fin == FIN_OVERFLOW) st = ST_NEEDS_HOST;
    else if (fin == FIN_NONE) {
        if (ps.mode == 1) st = ST_NONE;
        else st = finish_gjk_pair(ps, true, ps.cb_pos12.t, p2, n1, prediction, c);
    } else st = finish_gjk_pair(ps, true, p1, p2, n1, prediction, c);
    if (st == ST_SOME) to_world(ps, c);
    return st;
}

// The kernel itself, twice (see epa_persistent.inc): the hot configuration and the overflow configuration.
#define EPA_N(x) x
#define EPA_BIG 0
#define E2_THREADS 128
#define E2_STACK_CAP E2_STACK_SMEM
#define E2_ENT_T uint16_t
#define E2_ENT(face, opp) (uint16_t)((face) | ((uint32_t)(opp) << 8))
#define E2_ENT_FACE(e) ((e) & 0xffu)
#define E2_ENT_OPP(e) ((e) >> 8)
#define E2_SMEM_BYTES (E2_HEAP_SMEM * 128 * 5 + (E2_STACK_SMEM + E2_MAX_SIL) * 128 * 2)
#include "epa_persistent.inc"
#undef EPA_N
#undef EPA_BIG
#undef E2_THREADS
#undef E2_STACK_CAP
#undef E2_ENT_T
#undef E2_ENT
#undef E2_ENT_FACE
#undef E2_ENT_OPP
#pragma push_macro("E2_MAX_VERTS")
#pragma push_macro("E2_MAX_FACES")
#pragma push_macro("E2_MAX_SIL")
#pragma push_macro("E2_HEAP_SMEM")
#pragma push_macro("E2_MINB")
#undef E2_MAX_VERTS
#undef E2_MAX_FACES
#undef E2_MAX_SIL
#undef E2_HEAP_SMEM
#undef E2_MINB
#define EPA_N(x) x##_big
#define EPA_BIG 1
#define E2_THREADS 32
#define E2_MINB 1
#define E2_MAX_VERTS 128      // the reference stops after 100 expansions (epa3.rs:641-646): at most 4 + 2 + 101 vertices
#define E2_MAX_FACES 4096
#define E2_MAX_SIL 1024
#define E2_STACK_CAP 4096
#define E2_HEAP_SMEM 0
#define E2_ENT_T uint32_t
#define E2_ENT(face, opp) ((uint32_t)(face) | ((uint32_t)(opp) << 16))
#define E2_ENT_FACE(e) ((e) & 0xffffu)
#define E2_ENT_OPP(e) ((e) >> 16)
#include "epa_persistent.inc"
#undef EPA_N
#undef EPA_BIG
#undef E2_THREADS
#undef E2_STACK_CAP
#undef E2_ENT_T
#undef E2_ENT
#undef E2_ENT_FACE
#undef E2_ENT_OPP
#undef E2_MAX_VERTS
#undef E2_MAX_FACES
#undef E2_MAX_SIL
#undef E2_HEAP_SMEM
#undef E2_MINB
#pragma pop_macro("E2_MINB")
#pragma pop_macro("E2_HEAP_SMEM")
#pragma pop_macro("E2_MAX_SIL")
#pragma pop_macro("E2_MAX_FACES")
#pragma pop_macro("E2_MAX_VERTS")
#define E2_BIG_GRID 8   // overflow kernel: 8 single-warp CTAs = 256 runs in flight (~190 KB of arena each)

// Second half of phase F for k_contact_epa2: one thread per finished EPA run.
__global__ void __launch_bounds__(128) k_contact_finish(const uint8_t* __restrict__ kinds, const float4* __restrict__ params,
                              const float4* __restrict__ pts, PairSrc src, float prediction, OutSinks out,
                              const float4* __restrict__ fin_recs, const unsigned long long* __restrict__ fin_count) {
    unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= *fin_count) return;
    float4 r0 = fin_recs[3 * k], r1 = fin_recs[3 * k + 1], r2 = fin_recs[3 * k + 2];
    uint32_t pair = __float_as_uint(r0.x);
    int fin = (int)__float_as_uint(r0.y);
    PairSetup ps;
    pair_setup(kinds, params, pts, src, pair, ps);
    ContactOut c;
    int st = epa_result_to_contact(ps, fin, mk3(r0.z, r0.w, r1.x), mk3(r1.y, r1.z, r1.w), mk3(r2.x, r2.y, r2.z), prediction, c);
    emit(out, pair, st, c);
}

// ------------------------------------------------------------------------------------------- phase 2, compact arena
// Same state machine as k_contact_epa2, different memory shape. The ncu capture of k_contact_epa2
// (profiles/r1_contacts_epa2_full.json) shows 28.8 GB of DRAM traffic for 2.83 M runs — 10 KB per run although a run
// touches ~2 KB: 75 k resident threads x ~17 touched 128-byte lines = 166 MB of hot polytope state, more than the 126 MB
// L2, so lines are evicted and re-fetched several times per run, as scattered 32-byte sectors. This version halves the
// hot set: a face is 8 bytes (three vertex ids, three neighbour ids, deleted bit) — its normal is recomputed from the
// vertices when needed, bit-identically, instead of being stored —, witness points are replaced by the two hull-vertex
// ids that produced each polytope vertex (recomputed for the winning face), silhouette / DFS entries are 16-bit.
#define C2_MAX_VERTS 112
#define C2_MAX_FACES 255
#define C2_MAX_SIL 64
#define C2_STACK 96

struct __align__(128) EpaCArena {
    float4 vo[8];                // orig1 / orig2 of the parked simplex vertices (2 i, 2 i + 1)
    float4 vp[C2_MAX_VERTS];     // CSO point ; w = support ids (shape1 | shape2 << 16)
    uint2 face[C2_MAX_FACES + 1];// x = pts0 | pts1 << 8 | pts2 << 16 | deleted << 24 ; y = adj0 | adj1 << 8 | adj2 << 16
    float2 heap[C2_MAX_FACES + 1];// neg_dist ; face id (bits)
    uint16_t sil[C2_MAX_SIL];    // face | opp << 8
    uint16_t stk[C2_STACK];
};

__device__ __forceinline__ uint32_t c_pts(uint2 f, int i) { return (f.x >> (8 * i)) & 0xffu; }
__device__ __forceinline__ bool c_deleted(uint2 f) { return (f.x >> 24) != 0u; }
__device__ __forceinline__ uint32_t c_adj(uint2 f, int i) { return (f.y >> (8 * i)) & 0xffu; }
__device__ __forceinline__ int c_next_ccw(uint2 f, uint32_t id) {
    if (c_pts(f, 0) == id) return 1;
    if (c_pts(f, 1) == id) return 2;
    return 0;
}
// Face::new's normal (epa3.rs:96-118 -> utils::ccw_face_normal), recomputed from the three vertices
__device__ __forceinline__ V3 c_normal(const EpaCArena& A, uint2 f) {
    V3 va = v3of(A.vp[c_pts(f, 0)]), vb = v3of(A.vp[c_pts(f, 1)]), vc = v3of(A.vp[c_pts(f, 2)]);
    V3 n; float nn;
    if (!try_normalize_get(cross3(vb - va, vc - va), PB2_EPS, n, nn)) n = mk3(0.f, 0.f, 0.f);
    return n;
}
__device__ __forceinline__ void hc_sift_up(EpaCArena& A, int start, int pos) {
    float2 elt = A.heap[pos];
    while (pos > start) {
        int parent = (pos - 1) / 2;
        float2 pe = A.heap[parent];
        if (h2_le(elt.x, pe.x)) break;
        A.heap[pos] = pe;
        pos = parent;
    }
    A.heap[pos] = elt;
}
__device__ __forceinline__ void hc_push(EpaCArena& A, int& nheap, uint32_t id, float neg_dist) {
    int old = nheap;
    A.heap[old] = make_float2(neg_dist, __uint_as_float(id));
    nheap = old + 1;
    hc_sift_up(A, 0, old);
}
__device__ __forceinline__ float2 hc_pop(EpaCArena& A, int& nheap) {
    float2 item = A.heap[nheap - 1];
    nheap -= 1;
    if (nheap > 0) {
        float2 t = item; item = A.heap[0];
        int end = nheap, pos = 0;
        float2 elt = t;
        int child = 1;
        while (end >= 2 && child <= end - 2) {
            float2 c0 = A.heap[child], c1 = A.heap[child + 1];
            if (h2_le(c0.x, c1.x)) { child += 1; c0 = c1; }
            A.heap[pos] = c0;
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) { A.heap[pos] = A.heap[child]; pos = child; }
        A.heap[pos] = elt;
        hc_sift_up(A, 0, pos);
    }
    return item;
}
// support point with the ids of the hull vertices that produced it (cuboid: sign bits)
__device__ __forceinline__ V3 ds_local_support_id(const DShape& s, V3 dir, uint32_t& id) {
    id = 0;
    if (s.kind == DS_CUBOID) {
        id = (__float_as_uint(dir.x) >> 31) | ((__float_as_uint(dir.y) >> 31) << 1) | ((__float_as_uint(dir.z) >> 31) << 2);
        return mk3(copysignf(s.he.x, dir.x), copysignf(s.he.y, dir.y), copysignf(s.he.z, dir.z));
    }
    if (s.kind == DS_CONVEX) {
        float4 p = __ldg(&s.pts[0]);
        V3 best = mk3(p.x, p.y, p.z);
        float best_dot = dot3(best, dir);
        for (uint32_t i = 1; i < s.n; ++i) {
            float4 q = __ldg(&s.pts[i]);
            V3 v = mk3(q.x, q.y, q.z);
            float d = dot3(v, dir);
            if (d > best_dot) { best_dot = d; best = v; id = i; }
        }
        return best;
    }
    if (s.kind == DS_TRIANGLE) {
        float4 pa = __ldg(&s.pts[0]), pb = __ldg(&s.pts[1]), pc = __ldg(&s.pts[2]);
        V3 a = mk3(pa.x, pa.y, pa.z), b = mk3(pb.x, pb.y, pb.z), c = mk3(pc.x, pc.y, pc.z);
        float d1 = dot3(a, dir), d2 = dot3(b, dir), d3 = dot3(c, dir);
        if (d1 > d2) { if (d1 > d3) { id = 0; return a; } id = 2; return c; }
        if (d2 > d3) { id = 1; return b; }
        id = 2; return c;
    }
    return mk3(0.f, 0.f, 0.f);
}
__device__ __forceinline__ V3 ds_local_support_from_id(const DShape& s, uint32_t id) {
    if (s.kind == DS_CUBOID)
        return mk3(copysignf(s.he.x, (id & 1u) ? -1.0f : 1.0f), copysignf(s.he.y, (id & 2u) ? -1.0f : 1.0f), copysignf(s.he.z, (id & 4u) ? -1.0f : 1.0f));
    if (s.kind == DS_CONVEX || s.kind == DS_TRIANGLE) { float4 q = __ldg(&s.pts[id]); return mk3(q.x, q.y, q.z); }
    return mk3(0.f, 0.f, 0.f);
}

__global__ void __launch_bounds__(128, 4) k_contact_epac(const uint8_t* __restrict__ kinds, const float4* __restrict__ params,
                              const float4* __restrict__ pts, PairSrc src, float prediction, OutSinks out,
                              const EpaJob* __restrict__ jobs, const unsigned long long* __restrict__ job_count,
                              unsigned long long* __restrict__ next_job, EpaCArena* __restrict__ arenas, int refill) {
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    EpaCArena& A = arenas[blockIdx.x * blockDim.x + threadIdx.x];
    const unsigned long long total = *job_count;
    const float eps = PB2_EPS, eps_tol = PB2_EPS * 100.0f;

    int state = E2_IDLE;
    uint32_t pair = 0;
    Iso7 gpos12;
    DShape g1, g2;
    g1.kind = g2.kind = DS_ORIGIN; g1.n = g2.n = 0; g1.pts = g2.pts = nullptr; g1.he = g2.he = mk3(0.f, 0.f, 0.f);
    gpos12.q.i = gpos12.q.j = gpos12.q.k = 0.f; gpos12.q.w = 1.f; gpos12.t = mk3(0.f, 0.f, 0.f);
    int dim = 0, nverts = 0, nfaces = 0, nheap = 0, niter = 0;
    float max_dist = FLT_MAX, old_dist = 0.0f;
    uint32_t best_id = 0;
    bool exhausted = false;

    for (;;) {
        __syncwarp();
        unsigned idle = __ballot_sync(FULL, state == E2_IDLE);
        if (!exhausted && (idle == FULL || __popc(idle) >= refill)) {
            unsigned long long base = 0;
            int leader = __ffs(idle) - 1;
            if (lane == leader) base = atomicAdd(next_job, (unsigned long long)__popc(idle));
            base = __shfl_sync(FULL, base, leader);
            if (base >= total) exhausted = true;
            if (state == E2_IDLE) {
                unsigned long long j = base + __popc(idle & ((1u << lane) - 1u));
                if (j < total) {
                    const EpaJob& job = jobs[j];
                    pair = job.pair;
                    PairSetup ps;
                    pair_setup(kinds, params, pts, src, pair, ps);
                    gpos12 = ps.gpos12; g1 = ps.g1; g2 = ps.g2;
                    dim = (int)job.dim;
                    for (int i = 0; i <= dim; ++i) {
                        V3 o1 = mk3(job.o1[i][0], job.o1[i][1], job.o1[i][2]), o2 = mk3(job.o2[i][0], job.o2[i][1], job.o2[i][2]);
                        V3 p = o1 - o2;
                        A.vp[i] = make_float4(p.x, p.y, p.z, 0.f);
                        A.vo[2 * i] = make_float4(o1.x, o1.y, o1.z, 0.f);
                        A.vo[2 * i + 1] = make_float4(o2.x, o2.y, o2.z, 0.f);
                    }
                    nverts = dim + 1; nfaces = 0; nheap = 0; niter = 0;
                    max_dist = FLT_MAX; old_dist = 0.0f;
                    state = E2_INIT;
                }
            }
        }
        if (!__any_sync(FULL, state != E2_IDLE)) break;

        int fin = FIN_NOT;
        uint32_t fin_face = 0;
        bool need_support = false, run_step = false;
        V3 sdir = mk3(0.f, 0.f, 0.f);
        uint2 face = make_uint2(0u, 0u);
        V3 fnormal = mk3(0.f, 0.f, 0.f);
        uint32_t face_id = 0;
        float face_neg = 0.0f, curr_dist = 0.0f;
        int npend = 0;

        // ---- phase A: pop the closest live face (RUN) / seed the initial polytope (INIT)
        if (state == E2_RUN) {
            bool got = false;
            while (nheap > 0) {
                float2 ent = hc_pop(A, nheap);
                face_id = __float_as_uint(ent.y); face_neg = ent.x;
                face = A.face[face_id];
                if (!c_deleted(face)) { got = true; break; }
            }
            if (got) { need_support = true; run_step = true; fnormal = c_normal(A, face); sdir = fnormal; }
            else { fin = FIN_FACE; fin_face = best_id; }
        } else if (state == E2_INIT) {
            if (dim == 0) fin = FIN_DIM0;
            else if (dim == 3) {
                V3 v0 = v3of(A.vp[0]), v1 = v3of(A.vp[1]), v2 = v3of(A.vp[2]), v3 = v3of(A.vp[3]);
                if (dot3(cross3(v1 - v0, v2 - v0), v3 - v0) > 0.0f) {
                    float4 t = A.vp[1]; A.vp[1] = A.vp[2]; A.vp[2] = t;
                    float4 a1 = A.vo[2], b1 = A.vo[3];
                    A.vo[2] = A.vo[4]; A.vo[3] = A.vo[5]; A.vo[4] = a1; A.vo[5] = b1;
                }
                npend = 4;
            } else {
                if (dim == 1) {
                    V3 dpt = v3of(A.vp[1]) - v3of(A.vp[0]);
                    V3 a = fabsf(dpt.x) > fabsf(dpt.y) ? mk3(dpt.z, 0.0f, -dpt.x) : mk3(0.0f, -dpt.z, dpt.y);
                    a = normalize3(a);
                    sdir = cross3(a, dpt);
                    need_support = true;
                }
                npend = 2;
            }
        }
        // ---- phase B: one support point of the Minkowski difference
        uint32_t support_id = 0;
        V3 sp_point = mk3(0.f, 0.f, 0.f);
        if (need_support) {
            if (nverts >= C2_MAX_VERTS) { fin = FIN_OVERFLOW; run_step = false; npend = 0; }
            else {
                uint32_t id1 = 0, id2 = 0;
                V3 sp1 = ds_local_support_id(g1, sdir, id1);
                V3 sp2;
                if (g2.kind == DS_ORIGIN) sp2 = gpos12.t;
                else sp2 = iso_point(gpos12, ds_local_support_id(g2, iso_inv_vec(gpos12, -sdir), id2));
                V3 p = sp1 - sp2;
                support_id = (uint32_t)nverts;
                A.vp[nverts] = make_float4(p.x, p.y, p.z, __uint_as_float(id1 | (id2 << 16)));
                nverts++;
                sp_point = p;
            }
        }
        // ---- phase C/D: convergence test, then the silhouette of the faces visible from the new point
        if (run_step) {
            float candidate = dot3(sp_point, fnormal);
            if (candidate < max_dist) { best_id = face_id; max_dist = candidate; }
            curr_dist = -face_neg;
            if (max_dist - curr_dist < eps_tol || (fabsf(curr_dist - old_dist) < eps && candidate < max_dist)) {
                fin = FIN_FACE; fin_face = best_id; run_step = false;
            } else {
                old_dist = curr_dist;
                A.face[face_id].x = face.x | (1u << 24);
                int nsil = 0, sp = 0;
                bool ovf = false;
#pragma unroll 1
                for (int k = 2; k >= 0; --k) {
                    uint32_t af = c_adj(face, k);
                    int opp = c_next_ccw(A.face[af], c_pts(face, k));
                    A.stk[sp++] = (uint16_t)(af | ((uint32_t)opp << 8));
                }
                V3 pt = sp_point;
                while (sp > 0) {
                    uint32_t e = A.stk[--sp];
                    uint32_t fid = e & 0xffu; int fo = (int)(e >> 8);
                    uint2 f = A.face[fid];
                    if (c_deleted(f)) continue;
                    V3 q0 = v3of(A.vp[c_pts(f, 0)]), q1 = v3of(A.vp[c_pts(f, 1)]), q2 = v3of(A.vp[c_pts(f, 2)]);
                    V3 fn; float fnn;
                    if (!try_normalize_get(cross3(q1 - q0, q2 - q0), PB2_EPS, fn, fnn)) fn = mk3(0.f, 0.f, 0.f);
                    V3 p0 = fo == 0 ? q0 : (fo == 1 ? q1 : q2);
                    bool seen = dot3(pt - p0, fn) >= -PB2_GJK_EPS_TOL;
                    if (!seen) {
                        V3 p1 = fo == 0 ? q1 : (fo == 1 ? q2 : q0), p2 = fo == 0 ? q2 : (fo == 1 ? q0 : q1);
                        const float EPS = PB2_EPS * 100.0f;
                        seen = rel_eq(nrm2(cross3(p2 - p1, pt - p1)), 0.0f, EPS * EPS, PB2_EPS);
                    }
                    if (!seen) {
                        if (nsil >= C2_MAX_SIL) { ovf = true; break; }
                        A.sil[nsil++] = (uint16_t)e;
                    } else {
                        A.face[fid].x = f.x | (1u << 24);
                        int i1 = (fo + 2) % 3, i2 = fo;
                        uint32_t adj1 = c_adj(f, i1), adj2 = c_adj(f, i2);
                        int o1 = c_next_ccw(A.face[adj1], c_pts(f, i1));
                        int o2 = c_next_ccw(A.face[adj2], c_pts(f, i2));
                        if (sp + 2 > C2_STACK) { ovf = true; break; }
                        A.stk[sp++] = (uint16_t)(adj2 | ((uint32_t)o2 << 8));
                        A.stk[sp++] = (uint16_t)(adj1 | ((uint32_t)o1 << 8));
                    }
                }
                if (ovf) { fin = FIN_OVERFLOW; run_step = false; }
                else if (nsil == 0) { fin = FIN_NONE; run_step = false; }
                else npend = nsil;
            }
        }
        // ---- phase E: create the pending faces (initial polytope or the fan around the silhouette)
        int first_new = nfaces;
        if (fin == FIN_NOT && npend > 0) {
#pragma unroll 1
            for (int e = 0; e < npend; ++e) {
                int p0, p1, p2, a0, a1, a2, dv;
                int new_id = nfaces;
                if (state == E2_INIT) {
                    if (npend == 4) {
                        p0 = (e == 1) ? 1 : 0; p1 = (e == 0) ? 1 : ((e == 2) ? 2 : 3); p2 = (e == 0) ? 2 : ((e == 1) ? 2 : ((e == 2) ? 3 : 1));
                        a0 = (e == 0 || e == 1) ? 3 : ((e == 2) ? 0 : 2); a1 = (e == 1) ? 2 : 1; a2 = (e == 0) ? 2 : ((e == 2) ? 3 : 0);
                        dv = e;
                    } else {
                        p0 = 0; p1 = e == 0 ? 1 : 2; p2 = e == 0 ? 2 : 1;
                        a0 = a1 = a2 = e == 0 ? 1 : 0;
                        dv = 0;
                    }
                } else {
                    uint32_t ed = A.sil[e];
                    uint32_t efid = ed & 0xffu; int eopp = (int)(ed >> 8);
                    uint2 ef = A.face[efid];
                    if (c_deleted(ef)) continue;
                    if (new_id >= C2_MAX_FACES) { fin = FIN_OVERFLOW; break; }
                    p0 = (int)c_pts(ef, (eopp + 2) % 3); p1 = (int)c_pts(ef, (eopp + 1) % 3); p2 = (int)support_id;
                    a0 = (int)efid; a1 = new_id + 1; a2 = new_id - 1;
                    dv = p0;
                    int sh = 8 * ((eopp + 1) % 3);
                    A.face[efid].y = (ef.y & ~(0xffu << sh)) | ((uint32_t)new_id << sh);
                }
                V3 va = v3of(A.vp[p0]), vb = v3of(A.vp[p1]), vc = v3of(A.vp[p2]);
                bool inside = e2_face_inside(va, vb, vc);
                V3 n; float nn;
                if (!try_normalize_get(cross3(vb - va, vc - va), PB2_EPS, n, nn)) n = mk3(0.f, 0.f, 0.f);
                A.face[new_id] = make_uint2((uint32_t)p0 | ((uint32_t)p1 << 8) | ((uint32_t)p2 << 16),
                                            ((uint32_t)a0 & 0xffu) | (((uint32_t)a1 & 0xffu) << 8) | (((uint32_t)a2 & 0xffu) << 16));
                nfaces = new_id + 1;
                if (state == E2_INIT) {
                    if (npend == 4) {
                        if (inside) {
                            float dist = dot3(n, v3of(A.vp[dv]));
                            if (-dist > PB2_GJK_EPS_TOL) { fin = FIN_NONE; break; }
                            hc_push(A, nheap, (uint32_t)new_id, -dist);
                        }
                    } else {
                        hc_push(A, nheap, (uint32_t)new_id, 0.0f);
                    }
                } else if (inside) {
                    float dist = dot3(n, v3of(A.vp[dv]));
                    if (dist < curr_dist) { fin = FIN_FACE; fin_face = face_id; break; }
                    if (-dist > PB2_GJK_EPS_TOL) { fin = FIN_NONE; break; }
                    hc_push(A, nheap, (uint32_t)new_id, -dist);
                }
            }
            if (fin == FIN_NOT) {
                if (state == E2_INIT) {
                    if (nheap == 0) fin = FIN_NONE;
                    else { float2 top = A.heap[0]; best_id = __float_as_uint(top.y); state = E2_RUN; }
                } else {
                    if (first_new == nfaces) fin = FIN_NONE;
                    else {
                        A.face[first_new].y = (A.face[first_new].y & ~(0xffu << 16)) | ((uint32_t)(nfaces - 1) << 16);
                        A.face[nfaces - 1].y = (A.face[nfaces - 1].y & ~(0xffu << 8)) | ((uint32_t)first_new << 8);
                        niter += 1;
                        if (niter > 100) { fin = FIN_FACE; fin_face = best_id; }
                    }
                }
            }
        }
        // ---- phase F: finished lanes build the contact and go idle
        if (fin != FIN_NOT) {
            PairSetup ps;
            pair_setup(kinds, params, pts, src, pair, ps);
            ContactOut c;
            int st;
            V3 p1 = mk3(0.f, 0.f, 0.f), p2 = p1, n1 = mk3(0.f, 1.f, 0.f);
            if (fin == FIN_FACE) {
                uint2 f = A.face[fin_face];
                uint32_t ids[3] = {c_pts(f, 0), c_pts(f, 1), c_pts(f, 2)};
                float bc[3];
                e2_face_bc(v3of(A.vp[ids[0]]), v3of(A.vp[ids[1]]), v3of(A.vp[ids[2]]), bc);
                V3 w1[3], w2[3];
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    uint32_t vi = ids[q];
                    if ((int)vi <= dim) { w1[q] = v3of(A.vo[2 * vi]); w2[q] = v3of(A.vo[2 * vi + 1]); }
                    else {
                        uint32_t sid = __float_as_uint(A.vp[vi].w);
                        w1[q] = ds_local_support_from_id(g1, sid & 0xffffu);
                        w2[q] = g2.kind == DS_ORIGIN ? gpos12.t : iso_point(gpos12, ds_local_support_from_id(g2, sid >> 16));
                    }
                }
                p1 = w1[0] * bc[0] + w1[1] * bc[1] + w1[2] * bc[2];
                p2 = w2[0] * bc[0] + w2[1] * bc[1] + w2[2] * bc[2];
                n1 = c_normal(A, f);
            }
            if (fin == FIN_OVERFLOW) st = ST_NEEDS_HOST;
            else if (fin == FIN_NONE) {
                if (ps.mode == 1) st = ST_NONE;
                else st = finish_gjk_pair(ps, true, ps.cb_pos12.t, p2, n1, prediction, c);
            } else st = finish_gjk_pair(ps, true, p1, p2, n1, prediction, c);
            if (st == ST_SOME) to_world(ps, c);
            emit(out, pair, st, c);
            state = E2_IDLE;
        }
    }
}

// ------------------------------------------------------------------------------------------- host side
static int run_contacts(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                        const float* pos2, float prediction, uint32_t n, OutSinks sinks, const uint32_t* ab = nullptr, uint32_t n_colliders = 0,
                        const float4* mesh_tris = nullptr, uint32_t n_tris = 0, uint32_t flags = 0, const float* part_pose = nullptr,
                        uint32_t n_parts = 0, const float* init_dir = nullptr, uint32_t init_stride = 3, const float4* mesh_tris2 = nullptr) {
    PairSrc src;
    src.flags = flags;
    src.mesh_tris2 = mesh_tris2;
    src.init_dir = init_dir; src.init_stride = init_stride;
    src.part_pose = part_pose;
    src.shape1 = shape1; src.shape2 = shape2; src.pos1 = pos1; src.pos2 = pos2; src.ab = ab; src.mesh_tris = mesh_tris;
    src.n_first = mesh_tris ? n_tris : (part_pose ? n_parts : n_colliders); src.n_second = n_colliders;
    cudaStream_t st = ctx->stream;
    // EPA job queue (worst case: every pair) + arenas for the persistent EPA grid
    int epa_variant = 2, refill = 8;  // 2 (default): 14 KB arenas; 3: compact arena — 4x less DRAM traffic, same time
    {
        const char* e = getenv("PB2_EPA_VARIANT");
        if (e) epa_variant = atoi(e);
        if ((e = getenv("PB2_EPA_REFILL"))) refill = atoi(e);
    }
    size_t jobs_bytes = ((size_t)n * sizeof(EpaJob) + 255) & ~(size_t)255;
    PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[3], jobs_bytes + (size_t)n * 4));
    EpaJob* jobs = (EpaJob*)ctx->scratch[3].ptr;
    uint32_t* ovf_list = (uint32_t*)((char*)ctx->scratch[3].ptr + jobs_bytes);   // jobs that outgrow the hot EPA configuration
    unsigned long long* ovf_count = (unsigned long long*)(ctx->d_counters + 11);
    unsigned long long* next_job_big = (unsigned long long*)(ctx->d_counters + 13);
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 11, 0, 8, st));
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 13, 0, 8, st));
    unsigned long long* job_count = (unsigned long long*)(ctx->d_counters + 4);
    unsigned long long* next_job = (unsigned long long*)(ctx->d_counters + 5);
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 4, 0, 16, st));
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 7, 0, 8, st));   // finished-run records (slot 6 is the callers' contact count)
    int gjk_minb = 4;  // 128 registers, 4 CTAs per SM: 0.7 ms faster than the unconstrained 151-register build on the 4M-pair config
    { const char* e = getenv("PB2_GJK_MINB"); if (e) gjk_minb = atoi(e); }
    auto gjk = gjk_minb >= 5 ? k_contact_gjk<5> : (gjk_minb == 4 ? k_contact_gjk<4> : k_contact_gjk<3>);
    ctx->phase_marks = 0;
    pb2_phase_mark(ctx, 0);
    int gjk_persistent = 0, gjk_refill = 12;  // measured: no gain on hull pairs (31.9 vs 31.4 ms), 2.7x slower on TriMesh candidates
    { const char* e = getenv("PB2_GJK_PERSISTENT"); if (e) gjk_persistent = atoi(e); }
    { const char* e = getenv("PB2_GJK_REFILL"); if (e) gjk_refill = atoi(e); }
    if (gjk_persistent && n >= 4096) {
        auto gjkp = gjk_minb >= 5 ? k_contact_gjk_persistent<5> : (gjk_minb == 4 ? k_contact_gjk_persistent<4> : k_contact_gjk_persistent<3>);
        unsigned long long* next_pair = (unsigned long long*)(ctx->d_counters + 9);
        PB2_CUDA(ctx, cudaMemsetAsync(next_pair, 0, 8, st));
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gjkp, 128, 0);
        if (per_sm < 1) per_sm = 1;
        unsigned blocks = (unsigned)(ctx->sm_count * per_sm), need = pb2_blocks(n, 128);
        if (blocks > need) blocks = need;
        gjkp<<<blocks, 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, src, prediction, n, sinks, jobs, job_count,
                                     next_pair, gjk_refill);
    } else {
        gjk<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, src, prediction, n, sinks, jobs, job_count);
    }
    PB2_LAUNCHED(ctx);
    pb2_phase_mark(ctx, 1);
    if (epa_variant == 3) {
        int per_sm = 0;
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_contact_epac, 128, 0);
        if (per_sm < 1) per_sm = 1;
        { const char* pe = getenv("PB2_EPA_PER_SM"); if (pe && atoi(pe) > 0 && atoi(pe) < per_sm) per_sm = atoi(pe); }
        int epa_blocks = ctx->sm_count * per_sm;
        int need = (int)pb2_blocks(n, 128);
        if (epa_blocks > need) epa_blocks = need;
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[2], (size_t)128 * epa_blocks * sizeof(EpaCArena)));
        k_contact_epac<<<epa_blocks, 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, src, prediction, sinks,
                                                  jobs, job_count, next_job, (EpaCArena*)ctx->scratch[2].ptr, refill);
        pb2_phase_mark(ctx, 2);
    } else {
        int per_sm = 0;
        PB2_CUDA(ctx, cudaFuncSetAttribute(k_contact_epa2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E2_SMEM_BYTES));
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_contact_epa2, 128, E2_SMEM_BYTES);
        if (per_sm < 1) per_sm = 1;
        int epa_blocks = ctx->sm_count * per_sm;
        int need = (int)pb2_blocks(n, 128);
        if (epa_blocks > need) epa_blocks = need;
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[2], (size_t)128 * epa_blocks * sizeof(Epa2Arena)));
        const bool split_finish = !getenv("PB2_EPA_INLINE_FINISH");
        float4* fin_recs = nullptr;
        unsigned long long* fin_count = (unsigned long long*)(ctx->d_counters + 7);
        if (split_finish) {
            PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->scratch[4], (size_t)n * 48));
            fin_recs = (float4*)ctx->scratch[4].ptr;
        }
        k_contact_epa2<<<epa_blocks, 128, E2_SMEM_BYTES, st>>>(shapes->kinds, shapes->params, shapes->points4, src, prediction, sinks,
                                                  jobs, job_count, next_job, (Epa2Arena*)ctx->scratch[2].ptr, refill, fin_recs, fin_count,
                                                  nullptr, ovf_list, ovf_count);
        // Overflow configuration (global-memory arena, 4096 faces): runs the few jobs the hot kernel handed over; its warps find an
        // empty list and leave at once otherwise. The arena (48 MB) is allocated at the first contact call of the context.
        if (!ctx->epa_big_arena) PB2_CUDA(ctx, cudaMalloc(&ctx->epa_big_arena, (size_t)E2_BIG_GRID * 32 * sizeof(Epa2Arena_big)));
        PB2_LAUNCHED(ctx);
        k_contact_epa2_big<<<E2_BIG_GRID, 32, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, src, prediction, sinks, jobs, ovf_count,
                                                        next_job_big, (Epa2Arena_big*)ctx->epa_big_arena, 1, fin_recs, fin_count, ovf_list, nullptr, nullptr);
        pb2_phase_mark(ctx, 2);
        if (split_finish) {
            PB2_LAUNCHED(ctx);
            // grid sized for the worst case (every pair went to EPA); blocks past the record count exit on their first load
            k_contact_finish<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, src, prediction, sinks,
                                                             fin_recs, fin_count);
        }
    }
    PB2_LAUNCHED(ctx);
    pb2_phase_mark(ctx, 3);
    PB2_CUDA(ctx, cudaGetLastError());
#ifdef PB2_EPA_DEBUG
    {
        unsigned long long h[8];
        cudaStreamSynchronize(st);
        cudaMemcpyFromSymbol(h, g_epa_dbg, sizeof(h));
        fprintf(stderr, "[epa dbg] warp-trips %llu run %.2f init %.2f idle %.2f lanes/trip, exhausted trips %llu\n", h[0], (double)h[1] / h[0],
                (double)h[2] / h[0], (double)h[3] / h[0], h[4]);
        unsigned long long z[8] = {0};
        cudaMemcpyToSymbol(g_epa_dbg, z, sizeof(z));
    }
#endif
    return PB2_OK;
}

extern "C" {

int pb2_contact_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                      const float* pos2, float prediction, uint32_t n, pb2_contact* out, uint8_t* status, uint64_t* num_contacts, int mem) {
    if (!ctx || !shapes || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !out || !status))) return PB2_ERR_INVALID;
    if (num_contacts) *num_contacts = 0;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    OutSinks sinks;
    sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0; sinks.compact_count = nullptr;
    sinks.some_count = num_contacts ? (unsigned long long*)(ctx->d_counters + 6) : nullptr;
    if (num_contacts) PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 6, 0, 8, ctx->stream));
    if (mem == PB2_MEM_DEVICE) {
        sinks.dense = (float*)out; sinks.status = status;
        PB2_CHECK(run_contacts(ctx, shapes, shape1, shape2, pos1, pos2, prediction, n, sinks));
    } else {
        // Host buffers: chunked H2D / kernels / D2H pipeline (see pb2_trimesh_cast_rays).
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[0], (size_t)n * 4));
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[1], (size_t)n * 4));
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[2], (size_t)n * 28));
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[3], (size_t)n * 28));
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[4], (size_t)n * 52));
        PB2_CHECK(pb2_scratch_reserve(ctx, &ctx->stage[5], (size_t)n));
        uint32_t *d_s1 = (uint32_t*)ctx->stage[0].ptr, *d_s2 = (uint32_t*)ctx->stage[1].ptr;
        float *d_p1 = (float*)ctx->stage[2].ptr, *d_p2 = (float*)ctx->stage[3].ptr, *d_out = (float*)ctx->stage[4].ptr;
        uint8_t* d_status = (uint8_t*)ctx->stage[5].ptr;
        PB2_CHECK(pb2_pipeline_init(ctx));
        // make sure the job queue / arenas are sized before the pipeline starts (no reallocation mid-flight)
        const uint32_t CHUNK = 1u << 19;
        for (uint32_t lo = 0; lo < n; lo += CHUNK) {
            uint32_t cnt = n - lo < CHUNK ? n - lo : CHUNK;
            cudaEvent_t e_in = pb2_next_event(ctx), e_k = pb2_next_event(ctx);
            PB2_CUDA(ctx, cudaMemcpyAsync(d_s1 + lo, shape1 + lo, (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->copy_in));
            PB2_CUDA(ctx, cudaMemcpyAsync(d_s2 + lo, shape2 + lo, (size_t)cnt * 4, cudaMemcpyHostToDevice, ctx->copy_in));
            PB2_CUDA(ctx, cudaMemcpyAsync(d_p1 + 7ull * lo, pos1 + 7ull * lo, (size_t)cnt * 28, cudaMemcpyHostToDevice, ctx->copy_in));
            PB2_CUDA(ctx, cudaMemcpyAsync(d_p2 + 7ull * lo, pos2 + 7ull * lo, (size_t)cnt * 28, cudaMemcpyHostToDevice, ctx->copy_in));
            PB2_CUDA(ctx, cudaEventRecord(e_in, ctx->copy_in));
            PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, e_in, 0));
            sinks.dense = d_out + 13ull * lo; sinks.status = d_status + lo;
            PB2_CHECK(run_contacts(ctx, shapes, d_s1 + lo, d_s2 + lo, d_p1 + 7ull * lo, d_p2 + 7ull * lo, prediction, cnt, sinks));
            PB2_CUDA(ctx, cudaEventRecord(e_k, ctx->stream));
            PB2_CUDA(ctx, cudaStreamWaitEvent(ctx->copy_out, e_k, 0));
            PB2_CUDA(ctx, cudaMemcpyAsync((float*)out + 13ull * lo, d_out + 13ull * lo, (size_t)cnt * 52, cudaMemcpyDeviceToHost, ctx->copy_out));
            PB2_CUDA(ctx, cudaMemcpyAsync(status + lo, d_status + lo, (size_t)cnt, cudaMemcpyDeviceToHost, ctx->copy_out));
        }
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->copy_out));
    }
    if (num_contacts) {
        PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 6, ctx->d_counters + 6, 8, cudaMemcpyDeviceToHost, ctx->stream));
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        *num_contacts = ctx->h_counters[6];
    } else if (mem == PB2_MEM_HOST) {
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return PB2_OK;
}

// QueryDispatcher::contact(pos12, g1, g2, prediction) itself (query_dispatcher.rs:430-436): the relative pose is the input and the
// contact stays in the shapes' local frames (point1 / normal1 in shape 1's, point2 / normal2 in shape 2's) — what a dispatcher in
// a QueryDispatcherChain is asked for, as opposed to query::contact which takes two world poses and transforms the result.
int pb2_contact_batch_local(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos12,
                            float prediction, uint32_t n, pb2_contact* out, uint8_t* status, int mem) {
    if (!ctx || !shapes || (n && (!shape1 || !shape2 || !pos12 || !out || !status))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s1, *d_s2, *d_p;
    void *d_out, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos12, (size_t)n * 28, mem, &d_p));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_st));
    OutSinks sinks;
    sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0; sinks.compact_count = nullptr; sinks.some_count = nullptr;
    sinks.dense = (float*)d_out; sinks.status = (uint8_t*)d_st;
    PB2_CHECK(run_contacts(ctx, shapes, (const uint32_t*)d_s1, (const uint32_t*)d_s2, (const float*)d_p, (const float*)d_p, prediction, n, sinks,
                           nullptr, 0, nullptr, 0, PAIR_LOCAL_FRAMES | PAIR_POS12_GIVEN));
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

int pb2_contact_batch_compact(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                              const float* pos2, float prediction, uint32_t n, pb2_contact* out, uint32_t* pair_index, uint64_t cap,
                              uint64_t* count, int mem) {
    if (!ctx || !shapes || !count || (n && (!shape1 || !shape2 || !pos1 || !pos2))) return PB2_ERR_INVALID;
    *count = 0;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s1, *d_s2, *d_p1, *d_p2;
    void *d_out = nullptr, *d_idx = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)cap * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, pair_index, (size_t)cap * 4, mem, &d_idx));
    if (!d_out || !d_idx) cap = 0;
    OutSinks sinks;
    sinks.dense = nullptr; sinks.status = nullptr; sinks.compact = (float*)d_out; sinks.pair_index = (uint32_t*)d_idx; sinks.cap = cap;
    sinks.compact_count = (unsigned long long*)(ctx->d_counters + 6);
    sinks.some_count = nullptr;
    if (cap == 0) { sinks.compact = nullptr; sinks.some_count = sinks.compact_count; }
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 6, 0, 8, ctx->stream));
    PB2_CHECK(run_contacts(ctx, shapes, (const uint32_t*)d_s1, (const uint32_t*)d_s2, (const float*)d_p1, (const float*)d_p2, prediction, n, sinks));
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 6, ctx->d_counters + 6, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t total = ctx->h_counters[6];
    *count = total;
    uint64_t valid = total < cap ? total : cap;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)valid * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, pair_index, d_idx, (size_t)valid * 4, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total > cap) PB2_FAIL(ctx, PB2_ERR_OVERFLOW, "contact_batch_compact: %llu contacts > capacity %llu", (unsigned long long)total, (unsigned long long)cap);
    return PB2_OK;
}

int pb2_contact_pairs_compact(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* collider_shape, const float* collider_pose,
                              uint32_t n_colliders, const uint32_t* pairs, uint32_t n, float prediction, pb2_contact* out,
                              uint32_t* pair_index, uint64_t cap, uint64_t* count, int mem) {
    if (!ctx || !shapes || !count || (n && (!collider_shape || !collider_pose || !pairs))) return PB2_ERR_INVALID;
    *count = 0;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s, *d_p, *d_ab;
    void *d_out = nullptr, *d_idx = nullptr;
    PB2_CHECK(pb2_stage_in(ctx, 0, collider_shape, (size_t)n_colliders * 4, mem, &d_s));
    PB2_CHECK(pb2_stage_in(ctx, 2, collider_pose, (size_t)n_colliders * 28, mem, &d_p));
    PB2_CHECK(pb2_stage_in(ctx, 1, pairs, (size_t)n * 8, mem, &d_ab));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)cap * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, pair_index, (size_t)cap * 4, mem, &d_idx));
    if (!d_out || !d_idx) cap = 0;
    OutSinks sinks;
    sinks.dense = nullptr; sinks.status = nullptr; sinks.compact = (float*)d_out; sinks.pair_index = (uint32_t*)d_idx; sinks.cap = cap;
    sinks.compact_count = (unsigned long long*)(ctx->d_counters + 6);
    sinks.some_count = nullptr;
    if (cap == 0) { sinks.compact = nullptr; sinks.some_count = sinks.compact_count; }
    PB2_CUDA(ctx, cudaMemsetAsync(ctx->d_counters + 6, 0, 8, ctx->stream));
    PB2_CHECK(run_contacts(ctx, shapes, (const uint32_t*)d_s, (const uint32_t*)d_s, (const float*)d_p, (const float*)d_p, prediction, n, sinks,
                           (const uint32_t*)d_ab, n_colliders));
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 6, ctx->d_counters + 6, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    uint64_t total = ctx->h_counters[6];
    *count = total;
    uint64_t valid = total < cap ? total : cap;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)valid * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, pair_index, d_idx, (size_t)valid * 4, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (total > cap) PB2_FAIL(ctx, PB2_ERR_OVERFLOW, "contact_pairs_compact: %llu contacts > capacity %llu", (unsigned long long)total, (unsigned long long)cap);
    return PB2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- distance / intersection_test
// query::distance (distance.rs:89-97 -> DefaultQueryDispatcher::distance, default_query_dispatcher.rs:177-236) and
// query::intersection_test (intersection_test.rs:88-96 -> :104-175) for Ball / Cuboid / ConvexPolyhedron pairs: closed forms
// for ball-ball (distance_ball_ball.rs, intersection_test_ball_ball.rs), point projection with solid = true for ball vs
// cuboid / hull (distance_ball_convex_polyhedron.rs, intersection_test_ball_point_query.rs, point_aabb.rs:9-60,
// point_support_map.rs:17-52) and GJK for the support-map pairs (distance_support_map_support_map.rs — initial direction
// -pos12.translation; intersection_test_support_map_support_map.rs — max_dist 0, exact_dist false) and the SAT-based cuboid-cuboid
// arms (distance_cuboid_cuboid.rs -> closest_points_cuboid_cuboid.rs + closest_points_segment_segment.rs;
// intersection_test_cuboid_cuboid.rs). No pair of these shapes is handed back to the host.
template <bool DIST>
__global__ void __launch_bounds__(128) k_query_pairs(const uint8_t* __restrict__ kinds, const float4* __restrict__ params,
                              const float4* __restrict__ pts, uint32_t n_shapes, PairSrc src, uint32_t n, float* __restrict__ out_dist,
                              uint8_t* __restrict__ out_hit, uint8_t* __restrict__ status) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (DIST) out_dist[k] = 0.0f; else out_hit[k] = 0;
    if (src.shape1[k] >= n_shapes || src.shape2[k] >= n_shapes) { status[k] = (uint8_t)ST_UNSUPPORTED; return; }
    PairSetup ps;
    pair_setup(kinds, params, pts, src, k, ps);
    bool b1 = ps.k1 == PB2_SHAPE_BALL, b2 = ps.k2 == PB2_SHAPE_BALL;
    float dist = 0.0f;
    bool hit = false;
    int st = ST_NONE;  // 0 = Ok
    if (b1 && b2) {
        float d2 = nrm2(ps.pos12.t), sum = ps.pr1.x + ps.pr2.x;
        hit = d2 <= sum * sum;
        dist = hit ? 0.0f : sqrtf(d2) - sum;
    } else if (b1 || b2) {
        // ball vs point query: project the ball centre (in the other shape's frame) with solid = true
        float4 prc = b2 ? ps.pr1 : ps.pr2;
        uint8_t kc = b2 ? ps.k1 : ps.k2;
        float radius = b2 ? ps.pr2.x : ps.pr1.x;
        V3 c = ps.cb_pos12.t, proj;
        bool inside;
        if (kc == PB2_SHAPE_CUBOID) {
            V3 he = mk3(prc.x, prc.y, prc.z), zero = mk3(0.f, 0.f, 0.f);
            V3 shift = vmax3((-he) - c, zero) - vmax3(c - he, zero);
            inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
            proj = inside ? c : c + shift;
        } else {
            Simplex s;
            V3 dir; float nn;
            if (!try_normalize_get(c, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
            Iso7 m_inv; m_inv.q.i = 0.f; m_inv.q.j = 0.f; m_inv.q.k = 0.f; m_inv.q.w = 1.f; m_inv.t = c;
            sx_reset(s, cso_from_shapes(m_inv, ps.g1, ps.g2, dir));
            V3 p1, p2, n1;
            int r = gjk_closest_points<true>(ps.gpos12, ps.g1, ps.g2, FLT_MAX, s, p1, p2, n1);
            inside = r != GJK_CLOSEST_POINTS;
            proj = inside ? c : p1;
        }
        if (DIST) { float d = nrm(c - proj) - radius; dist = d > 0.0f ? d : 0.0f; }
        else hit = inside || nrm2(c - proj) <= radius * radius;
    } else if (ps.k1 == PB2_SHAPE_CUBOID && ps.k2 == PB2_SHAPE_CUBOID) {
        V3 he1 = mk3(ps.pr1.x, ps.pr1.y, ps.pr1.z), he2 = mk3(ps.pr2.x, ps.pr2.y, ps.pr2.z);
        if (DIST) dist = d_distance_cuboid_cuboid(ps.pos12, he1, he2);
        else hit = d_intersection_test_cuboid_cuboid(ps.pos12, he1, he2);
    } else {
        Simplex s;
        V3 dir; float nn;
        V3 d0 = DIST ? -ps.pos12.t : ps.pos12.t;
        if (!try_normalize_get(d0, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
        sx_reset(s, cso_from_shapes(ps.gpos12, ps.g1, ps.g2, dir));
        V3 p1, p2, n1;
        if (DIST) {
            int r = gjk_closest_points<true>(ps.gpos12, ps.g1, ps.g2, FLT_MAX, s, p1, p2, n1);
            dist = r == GJK_CLOSEST_POINTS ? nrm(p2 - p1) : 0.0f;
        } else {
            hit = gjk_closest_points<false>(ps.gpos12, ps.g1, ps.g2, 0.0f, s, p1, p2, n1) == GJK_INTERSECTION;
        }
    }
    if (DIST) out_dist[k] = dist; else out_hit[k] = hit ? 1 : 0;
    status[k] = (uint8_t)st;
}

static int run_query_pairs(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                           const float* pos2, uint32_t n, void* out, size_t out_elem, uint8_t* status, int mem, bool dist) {
    if (!ctx || !shapes || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !out || !status))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s1, *d_s2, *d_p1, *d_p2;
    void *d_out, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * out_elem, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_st));
    PairSrc src;
    src.shape1 = (const uint32_t*)d_s1; src.shape2 = (const uint32_t*)d_s2; src.pos1 = (const float*)d_p1; src.pos2 = (const float*)d_p2;
    src.ab = nullptr; src.mesh_tris = nullptr; src.n_first = src.n_second = 0;
    if (dist)
        k_query_pairs<true><<<pb2_blocks(n, 128), 128, 0, ctx->stream>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, src, n, (float*)d_out,
                                                                        nullptr, (uint8_t*)d_st);
    else
        k_query_pairs<false><<<pb2_blocks(n, 128), 128, 0, ctx->stream>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, src, n, nullptr,
                                                                         (uint8_t*)d_out, (uint8_t*)d_st);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * out_elem, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

extern "C" {

int pb2_distance_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1, const float* pos2,
                       uint32_t n, float* dist, uint8_t* status, int mem) {
    return run_query_pairs(ctx, shapes, shape1, shape2, pos1, pos2, n, dist, 4, status, mem, true);
}
int pb2_intersection_test_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                                const float* pos2, uint32_t n, uint8_t* hit, uint8_t* status, int mem) {
    return run_query_pairs(ctx, shapes, shape1, shape2, pos1, pos2, n, hit, 1, status, mem, false);
}

// ------------------------------------------------------------------------------------------- TriMesh vs shapes
}  // extern "C"

// ------------------------------------------------------------------------------------------- query::cast_shapes
// query::cast_shapes (shape_cast.rs:268-286) -> DefaultQueryDispatcher::cast_shapes (default_query_dispatcher.rs:434-515):
// ball-ball closed form (shape_cast_ball_ball.rs:10-69), everything else cast_shapes_support_map_support_map
// (shape_cast_support_map_support_map.rs:11-69) = gjk::directional_distance (gjk.rs:632-657) on the Minkowski difference,
// with RoundShapeRef (round_shape.rs:317-350) around shape 1 when target_distance > 0. Hits that start penetrating and need
// the impact geometry (toi < 1e-5) are parked: their contact comes from the GJK/EPA contact kernels (second phase).
enum { CAST_NONE = 0, CAST_CONVERGED = 1, CAST_PENETRATING = 2, CAST_UNSUPPORTED = 3, CAST_NEEDS_HOST = 4, CAST_PARKED = 250 };
struct CastOpts { float max_toi, target_distance; int stop_at_penetration, compute_geometry; };

// ray_toi_with_ball (ray_ball.rs:33-77) with dcenter = ray.origin - center already formed
__device__ __forceinline__ bool ray_ball_at(V3 dcenter, float radius, V3 dir, bool solid, bool& inside, float& toi) {
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

__device__ __forceinline__ DShape cast_dshape(uint8_t kind, float4 pr, const float4* pts) {
    DShape g = make_dshape(kind, pr, pts);
    if (kind == PB2_SHAPE_BALL) g.kind = DS_BALL;
    return g;
}

__global__ void __launch_bounds__(128) k_cast_shapes(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                              uint32_t n_shapes, const uint32_t* __restrict__ shape1, const uint32_t* __restrict__ shape2,
                              const float* __restrict__ pos1, const float* __restrict__ vel1, const float* __restrict__ pos2,
                              const float* __restrict__ vel2, CastOpts o, uint32_t n, float* __restrict__ out, uint8_t* __restrict__ status,
                              uint32_t* __restrict__ parked, unsigned long long* parked_count) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float* q = out + 13ull * k;
    uint32_t s1 = shape1[k], s2 = shape2[k];
    int st = CAST_NONE;
    V3 w1 = mk3(0.f, 0.f, 0.f), w2 = w1, n1 = w1, n2 = w1;
    float toi = 0.0f;
    if (s1 >= n_shapes || s2 >= n_shapes) st = CAST_UNSUPPORTED;
    else {
        Iso7 p1 = load_iso(pos1 + 7ull * k), p2 = load_iso(pos2 + 7ull * k);
        Iso7 pos12 = iso_inv_mul(p1, p2);
        V3 v1 = mk3(vel1[3ull * k], vel1[3ull * k + 1], vel1[3ull * k + 2]), v2 = mk3(vel2[3ull * k], vel2[3ull * k + 1], vel2[3ull * k + 2]);
        V3 vel12 = iso_inv_vec(p1, v2 - v1);
        uint8_t k1 = kinds[s1], k2 = kinds[s2];
        float4 pr1 = params[s1], pr2 = params[s2];
        if (k1 == PB2_SHAPE_BALL && k2 == PB2_SHAPE_BALL) {
            float rsum = pr1.x + pr2.x + o.target_distance;
            V3 center = -pos12.t;
            // ray_toi_with_ball(center, radius, Ray(origin, vel12), solid = true): dcenter = origin - center
            bool inside;
            if (ray_ball_at(mk3(0.f, 0.f, 0.f) - center, rsum, vel12, true, inside, toi) && !(toi > o.max_toi)) {
                V3 dpt = (mk3(0.f, 0.f, 0.f) + vel12 * toi) - center;
                if (rsum == 0.0f) {
                    n1 = mk3(1.f, 0.f, 0.f); n2 = iso_inv_vec(pos12, -n1);
                } else {
                    n1 = dpt / rsum; n2 = iso_inv_vec(pos12, -n1);
                    w1 = n1 * pr1.x; w2 = n2 * pr2.x;
                }
                if (!(!o.stop_at_penetration && toi < 1.0e-5f && dot3(n1, vel12) >= 0.0f))
                    st = (inside && nrm2(center) < rsum * rsum) ? CAST_PENETRATING : CAST_CONVERGED;
            }
        } else {
            DShape g1 = cast_dshape(k1, pr1, pts), g2 = cast_dshape(k2, pr2, pts);
            const float border = o.target_distance;
            Simplex s;
            V3 normal1;
            auto cso = [&](V3 dir) {
                V3 sp1;
                if (border > 0.0f) { V3 nd = dir / nrm(dir); sp1 = ds_local_support(g1, nd) + nd * border; }
                else sp1 = ds_local_support(g1, dir);
                return cso_make(sp1, ds_support_point(g2, pos12, -dir));
            };
            if (minkowski_ray_cast(cso, s, mk3(0.f, 0.f, 0.f), vel12, FLT_MAX, toi, normal1) && !(toi > o.max_toi)) {
                if ((o.compute_geometry || !o.stop_at_penetration) && toi < 1.0e-5f) {
                    st = CAST_PARKED;
                } else {
                    V3 r0 = mk3(0.f, 0.f, 0.f), r1 = r0;
                    if (toi != 0.0f) gjk_witness(s, s.dim == 3, r0, r1);
                    n1 = normal1;
                    n2 = iso_inv_vec(pos12, -normal1);
                    w1 = r0 - normal1 * border;
                    w2 = iso_inv_point(pos12, r1);
                    st = toi == 0.0f ? CAST_PENETRATING : CAST_CONVERGED;
                }
            }
        }
    }
    if (st == CAST_PARKED) parked[warp_append1(parked_count)] = k;
    bool some = st == CAST_CONVERGED || st == CAST_PENETRATING;
    if (!some && st != CAST_PARKED) { w1 = w2 = n1 = n2 = mk3(0.f, 0.f, 0.f); toi = 0.0f; }
    q[0] = w1.x; q[1] = w1.y; q[2] = w1.z; q[3] = w2.x; q[4] = w2.y; q[5] = w2.z;
    q[6] = n1.x; q[7] = n1.y; q[8] = n1.z; q[9] = n2.x; q[10] = n2.y; q[11] = n2.z; q[12] = toi;
    status[k] = (uint8_t)st;
}

// Second half of the penetrating branch of cast_shapes_support_map_support_map (:36-57), once the parked pairs' contacts exist.
__global__ void k_cast_merge(const uint32_t* __restrict__ parked, uint32_t count, const float* __restrict__ contacts, const uint8_t* __restrict__ cstatus,
                             const float* __restrict__ pos1, const float* __restrict__ vel1, const float* __restrict__ vel2, int stop_at_penetration,
                             float* __restrict__ out, uint8_t* __restrict__ status) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t k = parked[i];
    float* q = out + 13ull * k;
    const float* c = contacts + 13ull * i;
    int st = CAST_NONE;
    if (cstatus[i] == ST_NEEDS_HOST) st = CAST_NEEDS_HOST;
    else if (cstatus[i] == ST_SOME) {
        Iso7 p1 = load_iso(pos1 + 7ull * k);
        V3 v1 = mk3(vel1[3ull * k], vel1[3ull * k + 1], vel1[3ull * k + 2]), v2 = mk3(vel2[3ull * k], vel2[3ull * k + 1], vel2[3ull * k + 2]);
        V3 vel12 = iso_inv_vec(p1, v2 - v1);
        float normal_vel = dot3(mk3(c[6], c[7], c[8]), vel12);
        if (!(!stop_at_penetration && normal_vel >= 0.0f)) st = CAST_PENETRATING;
    }
    if (st == CAST_PENETRATING) { for (int j = 0; j < 12; ++j) q[j] = c[j]; }
    else { for (int j = 0; j < 13; ++j) q[j] = 0.0f; }
    status[k] = (uint8_t)st;
}

__global__ void k_pair_up(const uint32_t* __restrict__ idx, uint32_t count, uint32_t* __restrict__ ab) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) { ab[2ull * i] = idx[i]; ab[2ull * i + 1] = idx[i]; }
}

extern "C" int pb2_cast_shapes_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                                     const float* vel1, const float* pos2, const float* vel2, float max_time_of_impact, float target_distance,
                                     int stop_at_penetration, int compute_impact_geometry_on_penetration, uint32_t n, float* out, uint8_t* status,
                                     int mem) {
    if (!ctx || !shapes || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !vel1 || !vel2 || !out || !status))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_s1, *d_s2, *d_p1, *d_p2, *d_v1, *d_v2;
    void *d_out, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_in(ctx, 4, vel1, (size_t)n * 12, mem, &d_v1));
    PB2_CHECK(pb2_stage_in(ctx, 5, vel2, (size_t)n * 12, mem, &d_v2));
    PB2_CHECK(pb2_stage_out(ctx, 6, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 7, status, (size_t)n, mem, &d_st));
    uint32_t *d_parked = nullptr, *d_ab = nullptr;
    float* d_c = nullptr;
    uint8_t* d_cst = nullptr;
    int rc = PB2_OK;
    if (cudaMallocAsync((void**)&d_parked, (size_t)n * 4, st) != cudaSuccess) PB2_FAIL(ctx, PB2_ERR_CUDA, "cast_shapes: out of device memory");
    unsigned long long* parked_count = (unsigned long long*)(ctx->d_counters + 10);
    cudaMemsetAsync(parked_count, 0, 8, st);
    CastOpts o;
    o.max_toi = max_time_of_impact; o.target_distance = target_distance; o.stop_at_penetration = stop_at_penetration;
    o.compute_geometry = compute_impact_geometry_on_penetration;
    k_cast_shapes<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, (const uint32_t*)d_s1,
                                                      (const uint32_t*)d_s2, (const float*)d_p1, (const float*)d_v1, (const float*)d_p2,
                                                      (const float*)d_v2, o, n, (float*)d_out, (uint8_t*)d_st, d_parked, parked_count);
    PB2_LAUNCHED(ctx);
    cudaMemcpyAsync(ctx->h_counters + 10, parked_count, 8, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) { rc = PB2_ERR_CUDA; goto done; }
    {
        uint32_t cnt = (uint32_t)ctx->h_counters[10];
        if (cnt) {
            if (cudaMallocAsync((void**)&d_ab, (size_t)cnt * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_c, (size_t)cnt * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, cnt, st) != cudaSuccess) { rc = PB2_ERR_CUDA; goto done; }
            k_pair_up<<<pb2_blocks(cnt, 256), 256, 0, st>>>(d_parked, cnt, d_ab);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_c; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            rc = run_contacts(ctx, shapes, (const uint32_t*)d_s1, (const uint32_t*)d_s2, (const float*)d_p1, (const float*)d_p2, FLT_MAX, cnt, sinks,
                              d_ab, n, nullptr, 0, PAIR_SUPPORT_MAPS_ONLY | PAIR_LOCAL_FRAMES);
            if (rc != PB2_OK) goto done;
            k_cast_merge<<<pb2_blocks(cnt, 128), 128, 0, st>>>(d_parked, cnt, d_c, d_cst, (const float*)d_p1, (const float*)d_v1, (const float*)d_v2,
                                                               stop_at_penetration, (float*)d_out, (uint8_t*)d_st);
            PB2_LAUNCHED(ctx);
        }
    }
    if (cudaGetLastError() != cudaSuccess) { rc = PB2_ERR_CUDA; goto done; }
    rc = pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem);
    if (rc == PB2_OK) rc = pb2_stage_back(ctx, status, d_st, (size_t)n, mem);
    if (rc == PB2_OK && mem == PB2_MEM_HOST && cudaStreamSynchronize(st) != cudaSuccess) rc = PB2_ERR_CUDA;
done:
    if (d_parked) cudaFreeAsync(d_parked, st);
    if (d_ab) cudaFreeAsync(d_ab, st);
    if (d_c) cudaFreeAsync(d_c, st);
    if (d_cst) cudaFreeAsync(d_cst, st);
    return rc;
}


// ------------------------------------------------------------------------------- query::cast_shapes with a TriMesh on one side
// cast_shapes_composite_shape_shape / cast_shapes_shape_composite_shape (shape_cast_composite_shape_shape.rs:65-105, the composite arms
// of DefaultQueryDispatcher::cast_shapes, default_query_dispatcher.rs:498-515) -> CompositeShapeRef::cast_shape (:14-62): Bvh::find_best
// over the mesh tree with Minkowski-summed node boxes, every reached triangle cast against the shape like any support-map pair.
// One thread per query. P / V: pose and velocity of the shape in the MESH frame (the caller's pos12 / vel12 when the mesh is shape 1,
// pos12.inverse() / -pos12.inverse_transform_vector(vel12) when it is shape 2, :95-99).
static void apply_leaf_lanes_knob(pb2_ctx* ctx) {
    const char* e = getenv("PB2_LEAF_LANES");
    if (!e) return;
    int v = atoi(e);
    if (v < 1) v = 1;
    if (v > 32) v = 32;
    cudaMemcpyToSymbolAsync(g_pb2_leaf_lanes, &v, sizeof(int), 0, cudaMemcpyHostToDevice, ctx->stream);
}

__global__ void k_mesh_cast_frames(const float* __restrict__ mesh_pose, const float* __restrict__ mesh_vel, const float* __restrict__ poses,
                                   const float* __restrict__ vels, uint32_t n, int mesh_second, float* __restrict__ P, float* __restrict__ V) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    Iso7 pm = load_iso(mesh_pose), ps = load_iso(poses + 7ull * k);
    V3 vm = mk3(mesh_vel[0], mesh_vel[1], mesh_vel[2]), vs = mk3(vels[3ull * k], vels[3ull * k + 1], vels[3ull * k + 2]);
    Iso7 pose; V3 vel;
    if (!mesh_second) { pose = iso_inv_mul(pm, ps); vel = iso_inv_vec(pm, vs - vm); }   // shape_cast.rs:279-281
    else {
        Iso7 pos12 = iso_inv_mul(ps, pm);
        V3 vel12 = iso_inv_vec(ps, vm - vs);
        pose = iso_inverse(pos12);
        vel = -iso_inv_vec(pos12, vel12);
    }
    float* o = P + 7ull * k;
    o[0] = pose.q.i; o[1] = pose.q.j; o[2] = pose.q.k; o[3] = pose.q.w; o[4] = pose.t.x; o[5] = pose.t.y; o[6] = pose.t.z;
    V[3ull * k] = vel.x; V[3ull * k + 1] = vel.y; V[3ull * k + 2] = vel.z;
}

__global__ void __launch_bounds__(128) k_mesh_cast_shapes(const NodeWide* __restrict__ nodes, uint32_t n_leaves, const float4* __restrict__ tris,
                              const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                              const float* __restrict__ points, uint32_t n_shapes, const uint32_t* __restrict__ shape_ids,
                              const float* __restrict__ P, const float* __restrict__ V, CastOpts o, uint32_t n, float* __restrict__ out,
                              uint8_t* __restrict__ status, uint32_t* __restrict__ part, uint32_t* __restrict__ parked,
                              uint32_t* __restrict__ parked_ab, unsigned long long* parked_count, unsigned int* fault) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < n;   // every lane stays: the descent below is warp-cooperative
    if (!valid) k = 0;
    float* q = out + 13ull * k;
    uint32_t sid = shape_ids[k];
    int st = CAST_NONE;
    V3 w1 = mk3(0.f, 0.f, 0.f), w2 = w1, n1 = w1, n2 = w1;
    float best = o.max_toi;
    uint32_t best_id = PB2_INVALID_U32, best_pos = 0;
    bool found = false;
    const bool known = sid >= n_shapes ? false : true;
    if (!known) { st = CAST_UNSUPPORTED; sid = 0; }
    {
        Iso7 pos12 = load_iso(P + 7ull * k);
        V3 vel12 = mk3(V[3ull * k], V[3ull * k + 1], V[3ull * k + 2]);
        uint8_t k2 = kinds[sid];
        float4 pr2 = params[sid];
        V3 mn, mx;
        shape_aabb_dev(k2, pr2, points, pos12, mn, mx);   // g2.compute_aabb(pose12), :29
        V3 shift = -((mn + mx) * 0.5f);
        V3 margin = (mx - mn) * 0.5f + mk3(o.target_distance, o.target_distance, o.target_distance);
        V3 inv = mk3(1.0f / vel12.x, 1.0f / vel12.y, 1.0f / vel12.z);
        DShape g2 = cast_dshape(k2, pr2, pts);
        const float border = o.target_distance;
        auto leaf = [&](uint32_t pos, unsigned) {
            const float4* tp = tris + 3ull * pos;
            DShape g1; g1.kind = DS_TRIANGLE; g1.he = mk3(0.f, 0.f, 0.f); g1.pts = tp; g1.n = 3;
            Simplex s;
            V3 normal1; float toi;
            auto cso = [&](V3 dir) {
                V3 sp1;
                if (border > 0.0f) { V3 nd = dir / nrm(dir); sp1 = ds_local_support(g1, nd) + nd * border; }
                else sp1 = ds_local_support(g1, dir);
                return cso_make(sp1, ds_support_point(g2, pos12, -dir));
            };
            if (!minkowski_ray_cast(cso, s, mk3(0.f, 0.f, 0.f), vel12, FLT_MAX, toi, normal1) || toi > o.max_toi) return;
            uint32_t id = __float_as_uint(__ldg(&tp[0]).w);
            if (!(toi < best || (found && toi == best && id < best_id))) return;
            best = toi; best_id = id; best_pos = pos; found = true;
            if (o.compute_geometry && toi < 1.0e-5f) { st = CAST_PARKED; return; }   // geometry from the contact kernels (second phase)
            V3 r0 = mk3(0.f, 0.f, 0.f), r1 = r0;
            if (toi != 0.0f) gjk_witness(s, s.dim == 3, r0, r1);
            n1 = normal1;
            n2 = iso_inv_vec(pos12, -normal1);
            w1 = r0 - normal1 * border;
            w2 = iso_inv_point(pos12, r1);
            st = toi == 0.0f ? CAST_PENETRATING : CAST_CONVERGED;
        };
        bvh_find_best_msum(0xffffffffu, valid && known, nodes, n_leaves, shift, margin, vel12, inv, o.max_toi, best, found, leaf, fault);
    }
    if (!valid) return;
    if (st == CAST_PARKED) {
        unsigned long long at = warp_append1(parked_count);
        parked[at] = k;
        parked_ab[2 * at] = best_pos; parked_ab[2 * at + 1] = k;
    }
    bool some = found && st != CAST_UNSUPPORTED;
    if (!some || st == CAST_PARKED) { w1 = w2 = n1 = n2 = mk3(0.f, 0.f, 0.f); }
    q[0] = w1.x; q[1] = w1.y; q[2] = w1.z; q[3] = w2.x; q[4] = w2.y; q[5] = w2.z;
    q[6] = n1.x; q[7] = n1.y; q[8] = n1.z; q[9] = n2.x; q[10] = n2.y; q[11] = n2.z; q[12] = some ? best : 0.0f;
    status[k] = (uint8_t)st;
    part[k] = some ? best_id : PB2_INVALID_U32;
}

// ShapeCastHit::swapped (shape_cast.rs:71-80) for the shape-first order; hits that lost their contact lose their triangle too
__global__ void k_mesh_cast_finish(uint32_t n, int mesh_second, float* __restrict__ out, const uint8_t* __restrict__ status, uint32_t* __restrict__ part) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    int st = status[k];
    if (st != CAST_CONVERGED && st != CAST_PENETRATING) { part[k] = PB2_INVALID_U32; return; }
    if (!mesh_second) return;
    float* q = out + 13ull * k;
    for (int j = 0; j < 3; ++j) {
        float t = q[j]; q[j] = q[3 + j]; q[3 + j] = t;
        t = q[6 + j]; q[6 + j] = q[9 + j]; q[9 + j] = t;
    }
}

extern "C" int pb2_trimesh_cast_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const float* mesh_vel3, const pb2_shapes* shapes,
                                       const uint32_t* shape_ids, const float* poses7, const float* vels3, int mesh_second, float max_time_of_impact,
                                       float target_distance, int stop_at_penetration, int compute_impact_geometry_on_penetration, uint32_t n,
                                       float* out, uint8_t* status, uint32_t* part, int mem) {
    if (!ctx || !mesh || !shapes || !mesh_pose7 || !mesh_vel3 || (n && (!shape_ids || !poses7 || !vels3 || !out || !status || !part))) return PB2_ERR_INVALID;
    // with stop_at_penetration off a leaf's answer depends on its EPA contact (shape_cast_support_map_support_map.rs:41-46), which one
    // thread inside a tree descent cannot run: the reference's default (on) is what this entry offers
    if (!stop_at_penetration) PB2_FAIL(ctx, PB2_ERR_UNSUPPORTED, "trimesh_cast_shapes: stop_at_penetration = false is not offered on the device");
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_ids, *d_p, *d_v, *d_mp, *d_mv;
    void *d_out, *d_st, *d_part;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape_ids, (size_t)n * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 1, mesh_vel3, 12, mem, &d_mv));
    PB2_CHECK(pb2_stage_in(ctx, 2, poses7, (size_t)n * 28, mem, &d_p));
    PB2_CHECK(pb2_stage_in(ctx, 3, mesh_pose7, 28, mem, &d_mp));
    PB2_CHECK(pb2_stage_in(ctx, 4, vels3, (size_t)n * 12, mem, &d_v));
    PB2_CHECK(pb2_stage_out(ctx, 5, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 6, status, (size_t)n, mem, &d_st));
    PB2_CHECK(pb2_stage_out(ctx, 7, part, (size_t)n * 4, mem, &d_part));
    float *d_P = nullptr, *d_V = nullptr, *d_c = nullptr;
    uint32_t *d_parked = nullptr, *d_ab = nullptr;
    uint8_t* d_cst = nullptr;
    int rc = PB2_OK;
    do {
        if (cudaMallocAsync((void**)&d_P, (size_t)n * 28, st) != cudaSuccess || cudaMallocAsync((void**)&d_V, (size_t)n * 12, st) != cudaSuccess ||
            cudaMallocAsync((void**)&d_parked, (size_t)n * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_ab, (size_t)n * 8, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_shapes: out of device memory"); rc = PB2_ERR_CUDA; break;
        }
        unsigned long long* parked_count = (unsigned long long*)(ctx->d_counters + 10);
        cudaMemsetAsync(parked_count, 0, 8, st);
        k_mesh_cast_frames<<<pb2_blocks(n, 128), 128, 0, st>>>((const float*)d_mp, (const float*)d_mv, (const float*)d_p, (const float*)d_v, n,
                                                               mesh_second, d_P, d_V);
        PB2_LAUNCHED(ctx);
        CastOpts o;
        o.max_toi = max_time_of_impact; o.target_distance = target_distance; o.stop_at_penetration = 1;
        o.compute_geometry = compute_impact_geometry_on_penetration;
        apply_leaf_lanes_knob(ctx);
        k_mesh_cast_shapes<<<pb2_blocks(n, 128), 128, 0, st>>>(mesh->bvh.nodes, mesh->bvh.n_leaves, mesh->tris, shapes->kinds, shapes->params,
            shapes->points4, shapes->points, shapes->n, (const uint32_t*)d_ids, d_P, d_V, o, n, (float*)d_out, (uint8_t*)d_st, (uint32_t*)d_part,
            d_parked, d_ab, parked_count, PB2_FAULT_PTR(ctx));
        PB2_LAUNCHED(ctx);
        cudaMemcpyAsync(ctx->h_counters + 10, parked_count, 8, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_shapes: kernel failed"); rc = PB2_ERR_CUDA; break; }
        uint32_t cnt = (uint32_t)ctx->h_counters[10];
        if (cnt) {
            if (cudaMallocAsync((void**)&d_c, (size_t)cnt * 52, st) != cudaSuccess || cudaMallocAsync((void**)&d_cst, cnt, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_shapes: out of device memory"); rc = PB2_ERR_CUDA; break;
            }
            OutSinks sinks;
            sinks.dense = d_c; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            // contact_support_map_support_map(pos12, triangle, shape, Real::MAX) of the penetrating winners, in the two local frames
            if ((rc = run_contacts(ctx, shapes, nullptr, (const uint32_t*)d_ids, (const float*)d_mp, d_P, FLT_MAX, cnt, sinks, d_ab, n, mesh->tris,
                                   mesh->nt, PAIR_SUPPORT_MAPS_ONLY | PAIR_LOCAL_FRAMES | PAIR_POS12_GIVEN)) != PB2_OK) break;
            k_cast_merge<<<pb2_blocks(cnt, 128), 128, 0, st>>>(d_parked, cnt, d_c, d_cst, d_P, d_V, d_V, 1, (float*)d_out, (uint8_t*)d_st);
            PB2_LAUNCHED(ctx);
        }
        k_mesh_cast_finish<<<pb2_blocks(n, 128), 128, 0, st>>>(n, mesh_second, (float*)d_out, (const uint8_t*)d_st, (uint32_t*)d_part);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_shapes: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    void* frees[] = {d_P, d_V, d_parked, d_ab, d_c, d_cst};
    for (void* p : frees) if (p) cudaFreeAsync(p, st);
    if (rc != PB2_OK) return rc;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, part, d_part, (size_t)n * 4, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(st));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}


// TriMesh against TriMesh (the nesting behind the reference's tests/geometry/trimesh_trimesh_toi.rs): cast_shapes_composite_shape_shape
// over mesh 1, whose every reached triangle is cast against mesh 2 through cast_shapes_shape_composite_shape — mesh 2 walked under
// pos12.inverse() with the triangle's transformed box, each of its triangles cast against the triangle of mesh 1 as a support-map
// pair, the hit swapped back (shape_cast_composite_shape_shape.rs:65-105). One thread per query, two nested descents.
__global__ void __launch_bounds__(128) k_mesh_cast_mesh(const NodeWide* __restrict__ nodes1, uint32_t nl1, const float4* __restrict__ tris1,
                              const NodeWide* __restrict__ nodes2, uint32_t nl2, const float4* __restrict__ tris2, const float* __restrict__ pos1,
                              const float* __restrict__ vel1, const float* __restrict__ pos2, const float* __restrict__ vel2, CastOpts o, uint32_t n,
                              float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ parts, uint32_t* __restrict__ parked,
                              uint32_t* __restrict__ parked_pos, unsigned long long* parked_count, unsigned int* fault) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < n;   // every lane stays: the descents below are warp-cooperative
    if (!valid) k = 0;
    Iso7 p1 = load_iso(pos1 + 7ull * k), p2 = load_iso(pos2 + 7ull * k);
    Iso7 pos12 = iso_inv_mul(p1, p2);
    uint32_t best_pa = 0, best_pb = 0;
    V3 v1 = mk3(vel1[3ull * k], vel1[3ull * k + 1], vel1[3ull * k + 2]), v2 = mk3(vel2[3ull * k], vel2[3ull * k + 1], vel2[3ull * k + 2]);
    V3 vel12 = iso_inv_vec(p1, v2 - v1);
    int st = CAST_NONE;
    V3 w1 = mk3(0.f, 0.f, 0.f), w2 = w1, n1 = w1, n2 = w1;
    float best = o.max_toi;
    uint32_t best_id1 = PB2_INVALID_U32, best_id2 = PB2_INVALID_U32;
    bool found = false;
    if (nl1 && nl2) {
        NodeWide r = nodes2[0];   // TriMesh::compute_aabb(pos12) = root_aabb().transform_by(pos12) (trimesh.rs:1763-1765)
        V3 amn = mk3(r.left.mnx, r.left.mny, r.left.mnz), amx = mk3(r.left.mxx, r.left.mxy, r.left.mxz);
        if (nl2 > 1) {
            amn = vmin3(amn, mk3(r.right.mnx, r.right.mny, r.right.mnz));
            amx = vmax3(amx, mk3(r.right.mxx, r.right.mxy, r.right.mxz));
        }
        V3 ctr = iso_point(pos12, (amn + amx) * 0.5f), he = iso_abs_vec(pos12, (amx - amn) * 0.5f);
        V3 lmn = ctr + (-he), lmx = ctr + he;
        const V3 tgt = mk3(o.target_distance, o.target_distance, o.target_distance);
        V3 shift = -((lmn + lmx) * 0.5f), margin = (lmx - lmn) * 0.5f + tgt;
        V3 inv12 = mk3(1.0f / vel12.x, 1.0f / vel12.y, 1.0f / vel12.z);
        const Iso7 pos21 = iso_inverse(pos12);
        const V3 vel21 = -iso_inv_vec(pos12, vel12);
        const V3 inv21 = mk3(1.0f / vel21.x, 1.0f / vel21.y, 1.0f / vel21.z);
        const float border = o.target_distance;
        auto leaf1 = [&](uint32_t pa, unsigned lanes1) {
            const float4* t1 = tris1 + 3ull * pa;
            DShape gt1; gt1.kind = DS_TRIANGLE; gt1.he = mk3(0.f, 0.f, 0.f); gt1.pts = t1; gt1.n = 3;
            // Triangle::compute_aabb(pos21) = the box of the transformed vertices (aabb_triangle.rs:10-30)
            float4 fa = __ldg(&t1[0]), fb = __ldg(&t1[1]), fc = __ldg(&t1[2]);
            V3 qa = iso_point(pos21, mk3(fa.x, fa.y, fa.z)), qb = iso_point(pos21, mk3(fb.x, fb.y, fb.z)), qc = iso_point(pos21, mk3(fc.x, fc.y, fc.z));
            V3 bmn = mk3(fminf(fminf(qa.x, qb.x), qc.x), fminf(fminf(qa.y, qb.y), qc.y), fminf(fminf(qa.z, qb.z), qc.z));
            V3 bmx = mk3(fmaxf(fmaxf(qa.x, qb.x), qc.x), fmaxf(fmaxf(qa.y, qb.y), qc.y), fmaxf(fmaxf(qa.z, qb.z), qc.z));
            V3 shift2 = -((bmn + bmx) * 0.5f), margin2 = (bmx - bmn) * 0.5f + tgt;
            float ibest = o.max_toi;
            bool ifound = false;
            uint32_t iid2 = PB2_INVALID_U32, ipb = 0;
            int ist = CAST_NONE;
            V3 iw1 = mk3(0.f, 0.f, 0.f), iw2 = iw1, in1 = iw1, in2 = iw1;
            auto leaf2 = [&](uint32_t pb, unsigned) {
                const float4* t2 = tris2 + 3ull * pb;
                DShape gt2; gt2.kind = DS_TRIANGLE; gt2.he = mk3(0.f, 0.f, 0.f); gt2.pts = t2; gt2.n = 3;
                Simplex s;
                V3 normal1; float toi;
                auto cso = [&](V3 dir) {
                    V3 sp1;
                    if (border > 0.0f) { V3 nd = dir / nrm(dir); sp1 = ds_local_support(gt2, nd) + nd * border; }
                    else sp1 = ds_local_support(gt2, dir);
                    return cso_make(sp1, ds_support_point(gt1, pos21, -dir));
                };
                if (!minkowski_ray_cast(cso, s, mk3(0.f, 0.f, 0.f), vel21, FLT_MAX, toi, normal1) || toi > o.max_toi) return;
                uint32_t id2 = __float_as_uint(__ldg(&t2[0]).w);
                if (!(toi < ibest || (ifound && toi == ibest && id2 < iid2))) return;
                ibest = toi; iid2 = id2; ipb = pb; ifound = true;
                if (o.compute_geometry && toi < 1.0e-5f) { ist = CAST_PARKED; return; }
                V3 r0 = mk3(0.f, 0.f, 0.f), r1 = r0;
                if (toi != 0.0f) gjk_witness(s, s.dim == 3, r0, r1);
                in1 = normal1;
                in2 = iso_inv_vec(pos21, -normal1);
                iw1 = r0 - normal1 * border;
                iw2 = iso_inv_point(pos21, r1);
                ist = toi == 0.0f ? CAST_PENETRATING : CAST_CONVERGED;
            };
            bvh_find_best_msum(lanes1, true, nodes2, nl2, shift2, margin2, vel21, inv21, o.max_toi, ibest, ifound, leaf2, fault);
            if (!ifound) return;
            uint32_t id1 = __float_as_uint(fa.w);
            if (!(ibest < best || (found && ibest == best && id1 < best_id1))) return;
            best = ibest; best_id1 = id1; best_id2 = iid2; best_pa = pa; best_pb = ipb; found = true;
            st = ist;
            w1 = iw2; w2 = iw1; n1 = in2; n2 = in1;   // ShapeCastHit::swapped
        };
        bvh_find_best_msum(0xffffffffu, valid, nodes1, nl1, shift, margin, vel12, inv12, o.max_toi, best, found, leaf1, fault);
    }
    if (!valid) return;
    // a winner that starts in touch takes its geometry from the triangle-triangle contact (second phase)
    if (st == CAST_PARKED) {
        unsigned long long at = warp_append1(parked_count);
        parked[at] = k;
        parked_pos[2 * at] = best_pa; parked_pos[2 * at + 1] = best_pb;
        w1 = w2 = n1 = n2 = mk3(0.f, 0.f, 0.f);
    }
    float* q = out + 13ull * k;
    q[0] = w1.x; q[1] = w1.y; q[2] = w1.z; q[3] = w2.x; q[4] = w2.y; q[5] = w2.z;
    q[6] = n1.x; q[7] = n1.y; q[8] = n1.z; q[9] = n2.x; q[10] = n2.y; q[11] = n2.z; q[12] = found ? best : 0.0f;
    status[k] = (uint8_t)st;
    parts[2ull * k] = found ? best_id1 : PB2_INVALID_U32;
    parts[2ull * k + 1] = found ? best_id2 : PB2_INVALID_U32;
}

// Second phase of TriMesh-vs-TriMesh casts that start in touch: the winning triangle pairs as stand-alone records. The inner leaf
// problem of the reference is (triangle of mesh 2, triangle of mesh 1) under pos21 (shape_cast_composite_shape_shape.rs:95-103).
__global__ void k_mm_gather(const uint32_t* __restrict__ parked, const uint32_t* __restrict__ ppos, uint32_t cnt, const float4* __restrict__ tris1,
                            const float4* __restrict__ tris2, const float* __restrict__ pos1, const float* __restrict__ pos2,
                            float4* __restrict__ ta, float4* __restrict__ tb, float* __restrict__ pose, uint32_t* __restrict__ ab) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    uint32_t k = parked[i], pa = ppos[2ull * i], pb = ppos[2ull * i + 1];
    for (int j = 0; j < 3; ++j) { ta[3ull * i + j] = tris2[3ull * pb + j]; tb[3ull * i + j] = tris1[3ull * pa + j]; }
    Iso7 pos21 = iso_inverse(iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k)));
    float* o = pose + 7ull * i;
    o[0] = pos21.q.i; o[1] = pos21.q.j; o[2] = pos21.q.k; o[3] = pos21.q.w; o[4] = pos21.t.x; o[5] = pos21.t.y; o[6] = pos21.t.z;
    ab[2ull * i] = i; ab[2ull * i + 1] = i;
}
// ... and their contacts written back as the swapped hit (:100-104); a pair without a contact (EPA failure) loses its hit.
__global__ void k_mm_merge(const uint32_t* __restrict__ parked, uint32_t cnt, const float* __restrict__ contacts, const uint8_t* __restrict__ cstatus,
                           float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ parts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= cnt) return;
    uint32_t k = parked[i];
    float* q = out + 13ull * k;
    const float* c = contacts + 13ull * i;
    if (cstatus[i] == ST_SOME) {
        for (int j = 0; j < 3; ++j) { q[j] = c[3 + j]; q[3 + j] = c[j]; q[6 + j] = c[9 + j]; q[9 + j] = c[6 + j]; }
        status[k] = (uint8_t)CAST_PENETRATING;
    } else {
        for (int j = 0; j < 13; ++j) q[j] = 0.0f;
        status[k] = cstatus[i] == ST_NEEDS_HOST ? (uint8_t)CAST_NEEDS_HOST : (uint8_t)CAST_NONE;
        parts[2ull * k] = parts[2ull * k + 1] = PB2_INVALID_U32;
    }
}

extern "C" int pb2_trimesh_cast_trimesh(pb2_ctx* ctx, const pb2_trimesh* mesh1, const float* pos1, const float* vel1, const pb2_trimesh* mesh2,
                                        const float* pos2, const float* vel2, float max_time_of_impact, float target_distance,
                                        int stop_at_penetration, int compute_impact_geometry_on_penetration, uint32_t n, float* out,
                                        uint8_t* status, uint32_t* parts, int mem) {
    if (!ctx || !mesh1 || !mesh2 || (n && (!pos1 || !vel1 || !pos2 || !vel2 || !out || !status || !parts))) return PB2_ERR_INVALID;
    if (!stop_at_penetration) PB2_FAIL(ctx, PB2_ERR_UNSUPPORTED, "trimesh_cast_trimesh: stop_at_penetration = false is not offered on the device");
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_p1, *d_p2, *d_v1, *d_v2;
    void *d_out, *d_st, *d_parts;
    PB2_CHECK(pb2_stage_in(ctx, 0, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 1, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_in(ctx, 2, vel1, (size_t)n * 12, mem, &d_v1));
    PB2_CHECK(pb2_stage_in(ctx, 3, vel2, (size_t)n * 12, mem, &d_v2));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_st));
    PB2_CHECK(pb2_stage_out(ctx, 6, parts, (size_t)n * 8, mem, &d_parts));
    CastOpts o;
    o.max_toi = max_time_of_impact; o.target_distance = target_distance; o.stop_at_penetration = 1;
    o.compute_geometry = compute_impact_geometry_on_penetration;
    uint32_t *d_parked = nullptr, *d_ppos = nullptr, *d_ab = nullptr;
    float4 *d_ta = nullptr, *d_tb = nullptr;
    float *d_pose = nullptr, *d_c = nullptr;
    uint8_t* d_cst = nullptr;
    int rc = PB2_OK;
    do {
        if (cudaMallocAsync((void**)&d_parked, (size_t)n * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_ppos, (size_t)n * 8, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_trimesh: out of device memory"); rc = PB2_ERR_CUDA; break;
        }
        unsigned long long* parked_count = (unsigned long long*)(ctx->d_counters + 10);
        cudaMemsetAsync(parked_count, 0, 8, st);
        apply_leaf_lanes_knob(ctx);
        k_mesh_cast_mesh<<<pb2_blocks(n, 128), 128, 0, st>>>(mesh1->bvh.nodes, mesh1->bvh.n_leaves, mesh1->tris, mesh2->bvh.nodes, mesh2->bvh.n_leaves,
            mesh2->tris, (const float*)d_p1, (const float*)d_v1, (const float*)d_p2, (const float*)d_v2, o, n, (float*)d_out, (uint8_t*)d_st,
            (uint32_t*)d_parts, d_parked, d_ppos, parked_count, PB2_FAULT_PTR(ctx));
        PB2_LAUNCHED(ctx);
        cudaMemcpyAsync(ctx->h_counters + 10, parked_count, 8, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_trimesh: kernel failed"); rc = PB2_ERR_CUDA; break; }
        uint32_t cnt = (uint32_t)ctx->h_counters[10];
        if (cnt) {
            if (cudaMallocAsync((void**)&d_ta, (size_t)cnt * 48, st) != cudaSuccess || cudaMallocAsync((void**)&d_tb, (size_t)cnt * 48, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_pose, (size_t)cnt * 28, st) != cudaSuccess || cudaMallocAsync((void**)&d_ab, (size_t)cnt * 8, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_c, (size_t)cnt * 52, st) != cudaSuccess || cudaMallocAsync((void**)&d_cst, cnt, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_trimesh: out of device memory"); rc = PB2_ERR_CUDA; break;
            }
            k_mm_gather<<<pb2_blocks(cnt, 128), 128, 0, st>>>(d_parked, d_ppos, cnt, mesh1->tris, mesh2->tris, (const float*)d_p1, (const float*)d_p2,
                                                              d_ta, d_tb, d_pose, d_ab);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_c; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            pb2_shapes none;   // no table shape takes part
            // contact_support_map_support_map(pos21, triangle of mesh 2, triangle of mesh 1, Real::MAX), in the two meshes' frames
            if ((rc = run_contacts(ctx, &none, nullptr, nullptr, d_pose, d_pose, FLT_MAX, cnt, sinks, d_ab, cnt, d_ta, cnt,
                                   PAIR_SUPPORT_MAPS_ONLY | PAIR_LOCAL_FRAMES | PAIR_POS12_GIVEN, nullptr, 0, nullptr, 3, d_tb)) != PB2_OK) break;
            k_mm_merge<<<pb2_blocks(cnt, 128), 128, 0, st>>>(d_parked, cnt, d_c, d_cst, (float*)d_out, (uint8_t*)d_st, (uint32_t*)d_parts);
            PB2_LAUNCHED(ctx);
        }
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_cast_trimesh: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    void* frees[] = {d_parked, d_ppos, d_ab, d_ta, d_tb, d_pose, d_c, d_cst};
    for (void* p : frees) if (p) cudaFreeAsync(p, st);
    if (rc != PB2_OK) return rc;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, parts, d_parts, (size_t)n * 8, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(st));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}


// ------------------------------------------------------------------------------- query::distance with a TriMesh on one side
// distance_composite_shape_shape / distance_shape_composite_shape (distance_composite_shape_shape.rs:46-77, the composite arms of
// DefaultQueryDispatcher::distance, default_query_dispatcher.rs:288-297) -> CompositeShapeRef::distance_to_shape (:13-42). One thread
// per query; the leaf is DefaultQueryDispatcher::distance on (Triangle, shape): a Triangle is convex, so a Ball goes through
// distance_convex_polyhedron_ball with the triangle's own point projection (distance_ball_convex_polyhedron.rs:24-33,
// point_triangle.rs:17-25), a Cuboid / ConvexPolyhedron through distance_support_map_support_map (GJK from -pos12.translation).
__global__ void __launch_bounds__(128) k_mesh_distance(const NodeWide* __restrict__ nodes, uint32_t n_leaves, const float4* __restrict__ tris,
                              const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                              const float* __restrict__ points, uint32_t n_shapes, const uint32_t* __restrict__ shape_ids,
                              const float* __restrict__ mesh_pose, const float* __restrict__ poses, int mesh_second, uint32_t n,
                              float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ part, unsigned int* fault) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < n;   // every lane stays: the descent below is warp-cooperative
    if (!valid) k = 0;
    uint32_t sid = shape_ids[k];
    const bool known = sid < n_shapes;
    if (!known) sid = 0;
    Iso7 pm = load_iso(mesh_pose), ps = load_iso(poses + 7ull * k);
    Iso7 pos12 = mesh_second ? iso_inverse(iso_inv_mul(ps, pm)) : iso_inv_mul(pm, ps);
    uint8_t k2 = kinds[sid];
    float4 pr2 = params[sid];
    V3 mn, mx;
    shape_aabb_dev(k2, pr2, points, pos12, mn, mx);
    V3 shift = -((mn + mx) * 0.5f), margin = (mx - mn) * 0.5f;
    float best = FLT_MAX;
    uint32_t best_id = PB2_INVALID_U32;
    bool found = false;
    DShape g2 = make_dshape(k2, pr2, pts);
    auto leaf = [&](uint32_t pos, unsigned) {
        const float4* tp = tris + 3ull * pos;
        float d;
        if (k2 == PB2_SHAPE_BALL) {
            float4 fa = __ldg(&tp[0]), fb = __ldg(&tp[1]), fc = __ldg(&tp[2]);
            Proj pr;
            project_on_triangle(mk3(fa.x, fa.y, fa.z), mk3(fb.x, fb.y, fb.z), mk3(fc.x, fc.y, fc.z), pos12.t, pr);
            d = nrm(pos12.t - pr.point) - pr2.x;
            d = d > 0.0f ? d : 0.0f;
        } else {
            DShape g1; g1.kind = DS_TRIANGLE; g1.he = mk3(0.f, 0.f, 0.f); g1.pts = tp; g1.n = 3;
            Simplex s;
            V3 dir; float nn;
            if (!try_normalize_get(-pos12.t, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
            sx_reset(s, cso_from_shapes(pos12, g1, g2, dir));
            V3 p1, p2, n1;
            int r = gjk_closest_points<true>(pos12, g1, g2, FLT_MAX, s, p1, p2, n1);
            d = r == GJK_CLOSEST_POINTS ? nrm(p2 - p1) : 0.0f;
        }
        uint32_t id = __float_as_uint(__ldg(&tp[0]).w);
        if (d < best || (found && d == best && id < best_id)) { best = d; best_id = id; found = true; }
    };
    bvh_find_best_msum_distance(0xffffffffu, valid && known, nodes, n_leaves, shift, margin, best, found, leaf, fault);
    if (!valid) return;
    if (!known) { out[k] = 0.0f; status[k] = (uint8_t)ST_UNSUPPORTED; part[k] = PB2_INVALID_U32; return; }
    out[k] = best;   // unwrap_or((u32::MAX, Real::MAX)) for a mesh whose every leaf was removed
    status[k] = (uint8_t)ST_NONE;
    part[k] = best_id;
}

extern "C" int pb2_trimesh_distance_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const pb2_shapes* shapes,
                                           const uint32_t* shape_ids, const float* poses7, int mesh_second, uint32_t n, float* dist, uint8_t* status,
                                           uint32_t* part, int mem) {
    if (!ctx || !mesh || !shapes || !mesh_pose7 || (n && (!shape_ids || !poses7 || !dist || !status || !part))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_ids, *d_p, *d_mp;
    void *d_out, *d_st, *d_part;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape_ids, (size_t)n * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 2, poses7, (size_t)n * 28, mem, &d_p));
    PB2_CHECK(pb2_stage_in(ctx, 3, mesh_pose7, 28, mem, &d_mp));
    PB2_CHECK(pb2_stage_out(ctx, 4, dist, (size_t)n * 4, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_st));
    PB2_CHECK(pb2_stage_out(ctx, 6, part, (size_t)n * 4, mem, &d_part));
    apply_leaf_lanes_knob(ctx);
    k_mesh_distance<<<pb2_blocks(n, 128), 128, 0, st>>>(mesh->bvh.nodes, mesh->bvh.n_leaves, mesh->tris, shapes->kinds, shapes->params, shapes->points4,
        shapes->points, shapes->n, (const uint32_t*)d_ids, (const float*)d_mp, (const float*)d_p, mesh_second, n, (float*)d_out, (uint8_t*)d_st,
        (uint32_t*)d_part, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, dist, d_out, (size_t)n * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, part, d_part, (size_t)n * 4, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(st));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}


// ------------------------------------------------------------------------------- Bvh::project_point with typed leaves
// Bvh::project_point (bvh_queries.rs:213-227): find_best with node cost Aabb::distance_to_local_point(pt, solid = true)
// (point_aabb.rs:135-146) and the leaf check PointQuery::project_point(pose, pt, solid) (point_query.rs:147-151) of the leaf's shape:
// Ball (point_ball.rs:9-21), Cuboid (point_cuboid.rs -> point_aabb.rs:9-60), ConvexPolyhedron (point_support_map.rs:17-52: GJK, the
// point itself when inside and solid). A point inside a hull with solid = false needs EPA (:39-52), which one thread inside a descent
// does not run: that query is answered status 3. Leaf cost = na::distance(world-space projection, pt).
__global__ void __launch_bounds__(128) k_project_points_shapes(const NodeWide* __restrict__ nodes, const uint32_t* __restrict__ order, uint32_t n_leaves,
                              const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                              uint32_t n_shapes, const uint32_t* __restrict__ shape_ids, const float* __restrict__ poses,
                              const float* __restrict__ points, uint32_t m, float max_distance, bool solid, float* __restrict__ out_proj,
                              uint8_t* __restrict__ out_inside, uint32_t* __restrict__ out_leaf, uint8_t* __restrict__ status, unsigned int* fault) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = k < m;   // every lane stays: the descent below is warp-cooperative
    if (!valid) k = 0;
    const V3 p = mk3(points[3ull * k], points[3ull * k + 1], points[3ull * k + 2]);
    float best = max_distance;
    bool found = false, best_inside = false, needs_host = false;
    uint32_t best_id = PB2_INVALID_U32;
    V3 best_pt = mk3(0.f, 0.f, 0.f);
    auto leaf = [&](uint32_t pos, unsigned) {
        uint32_t id = order[pos];
        uint32_t sid = shape_ids ? shape_ids[id] : id;
        if (sid >= n_shapes) { atomicOr(fault, PB2_FAULT_BAD_ID); return; }
        Iso7 pose = load_iso(poses + 7ull * id);
        V3 lp = iso_inv_point(pose, p), q;
        bool in;
        float4 pr = params[sid];
        uint8_t kind = kinds[sid];
        if (kind == PB2_SHAPE_BALL) {
            float d2 = nrm2(lp);
            in = d2 <= pr.x * pr.x;
            q = (in && solid) ? lp : lp * (pr.x / sqrtf(d2));
        } else if (kind == PB2_SHAPE_CUBOID) {
            Feat f;
            d_cuboid_project(mk3(pr.x, pr.y, pr.z), lp, q, in, f);   // the non-solid projection
            if (in && solid) q = lp;
        } else {
            DShape g1 = make_dshape(kind, pr, pts), g2 = origin_dshape();
            Simplex s;
            V3 dir; float nn;
            if (!try_normalize_get(lp, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
            Iso7 m_inv; m_inv.q.i = 0.f; m_inv.q.j = 0.f; m_inv.q.k = 0.f; m_inv.q.w = 1.f; m_inv.t = lp;
            sx_reset(s, cso_from_shapes(m_inv, g1, g2, dir));
            Iso7 mm; mm.q = m_inv.q; mm.t = -lp;
            V3 p1, p2, n1;
            int r = gjk_closest_points<true>(iso_inverse(mm), g1, g2, FLT_MAX, s, p1, p2, n1);
            in = r != GJK_CLOSEST_POINTS;
            if (in && !solid) { needs_host = true; return; }
            q = in ? lp : p1;
        }
        V3 w = iso_point(pose, q);
        float d = nrm(p - w);
        if (d < best || (found && d == best && id < best_id)) { best = d; best_id = id; best_pt = w; best_inside = in; found = true; }
    };
    auto cost = [&](float4 lo, float4 hi, float) {
        V3 shift = vmax3(vmax3(mk3(lo.x, lo.y, lo.z) - p, p - mk3(hi.x, hi.y, hi.z)), mk3(0.f, 0.f, 0.f));
        return nrm(shift);
    };
    bvh_find_best_cost(0xffffffffu, valid, nodes, n_leaves, max_distance, best, found, cost, leaf, fault);
    if (!valid) return;
    if (needs_host) { found = false; best_pt = mk3(0.f, 0.f, 0.f); best_inside = false; best_id = PB2_INVALID_U32; }
    out_proj[3ull * k] = best_pt.x; out_proj[3ull * k + 1] = best_pt.y; out_proj[3ull * k + 2] = best_pt.z;
    out_inside[k] = best_inside ? 1 : 0;
    out_leaf[k] = best_id;
    status[k] = needs_host ? (uint8_t)ST_NEEDS_HOST : (found ? (uint8_t)ST_SOME : (uint8_t)ST_NONE);
}

extern "C" int pb2_bvh_project_points_shapes(pb2_ctx* ctx, const pb2_bvh* bvh, const pb2_shapes* shapes, const uint32_t* shape_ids, const float* poses7,
                                             const float* points, uint32_t m, float max_distance, int solid, float* proj, uint8_t* inside,
                                             uint32_t* leaf, uint8_t* status, int mem) {
    if (!ctx || !bvh || !shapes || !poses7 || (m && (!points || !proj || !inside || !leaf || !status))) return PB2_ERR_INVALID;
    if (m == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    uint32_t nl = bvh->n_leaves;
    const void *d_pts, *d_ids = nullptr, *d_poses;
    void *d_proj, *d_in, *d_leaf, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, points, (size_t)m * 12, mem, &d_pts));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape_ids, (size_t)nl * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 6, poses7, (size_t)nl * 28, mem, &d_poses));
    PB2_CHECK(pb2_stage_out(ctx, 2, proj, (size_t)m * 12, mem, &d_proj));
    PB2_CHECK(pb2_stage_out(ctx, 3, inside, (size_t)m, mem, &d_in));
    PB2_CHECK(pb2_stage_out(ctx, 4, leaf, (size_t)m * 4, mem, &d_leaf));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)m, mem, &d_st));
    apply_leaf_lanes_knob(ctx);
    k_project_points_shapes<<<pb2_blocks(m, 128), 128, 0, st>>>(bvh->nodes, bvh->leaf_order, nl, shapes->kinds, shapes->params, shapes->points4,
        shapes->n, (const uint32_t*)d_ids, (const float*)d_poses, (const float*)d_pts, m, max_distance, solid != 0, (float*)d_proj,
        (uint8_t*)d_in, (uint32_t*)d_leaf, (uint8_t*)d_st, PB2_FAULT_PTR(ctx));
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, proj, d_proj, (size_t)m * 12, mem));
    PB2_CHECK(pb2_stage_back(ctx, inside, d_in, (size_t)m, mem));
    PB2_CHECK(pb2_stage_back(ctx, leaf, d_leaf, (size_t)m * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)m, mem));
    if (mem == PB2_MEM_HOST) {
        PB2_CHECK(pb2_fetch_fault(ctx));
        PB2_CUDA(ctx, cudaStreamSynchronize(st));
        return pb2_check_fault(ctx);
    }
    return PB2_OK;
}


extern "C" int pb2_intersect_csr_device(pb2_ctx* ctx, const pb2_bvh* bvh, const float* d_queries, uint32_t m, bool positions,
                                        uint32_t* d_offsets, uint32_t** d_items, uint64_t* total_out);

// shape2.compute_aabb(pose12).loosened(prediction) in the mesh's frame (contact_composite_shape_shape.rs:21)
__global__ void k_mesh_query_aabbs(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float* __restrict__ points,
                                   uint32_t n_shapes, const uint32_t* __restrict__ shape_ids, const float* __restrict__ poses,
                                   const float* __restrict__ mesh_pose, uint32_t n, float prediction, bool pos12_given, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float* o = out + 6ull * i;
    uint32_t sid = shape_ids[i];
    if (sid >= n_shapes) { o[0] = o[1] = o[2] = FLT_MAX; o[3] = o[4] = o[5] = -FLT_MAX; return; }  // no candidates: reported by the reduce
    Iso7 pos12 = pos12_given ? load_iso(poses + 7ull * i) : iso_inv_mul(load_iso(mesh_pose), load_iso(poses + 7ull * i));
    V3 mn, mx;
    shape_aabb_dev(kinds[sid], params[sid], points, pos12, mn, mx);
    o[0] = mn.x + (-prediction); o[1] = mn.y + (-prediction); o[2] = mn.z + (-prediction);
    o[3] = mx.x + prediction; o[4] = mx.y + prediction; o[5] = mx.z + prediction;
}

__global__ void k_expand_candidates(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ items, uint32_t n, uint32_t* __restrict__ ab) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    for (uint32_t j = offsets[q]; j < offsets[q + 1]; ++j) { ab[2ull * j] = items[j]; ab[2ull * j + 1] = q; }
}

// CompositeShapeRef::contact_with_shape's reduction (contact_composite_shape_shape.rs:26-41): the contact with the smallest
// dist wins; equal dists go to the smallest triangle index (the reference keeps the first in its own tree's order).
__global__ void k_mesh_reduce(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ items, const float4* __restrict__ tris,
                              const float* __restrict__ cand, const uint8_t* __restrict__ cand_status, const uint32_t* __restrict__ shape_ids,
                              uint32_t n_shapes, uint32_t n, float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ part) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    int st = shape_ids[q] >= n_shapes ? ST_UNSUPPORTED : ST_NONE;
    uint32_t best_j = 0, best_id = PB2_INVALID_U32;
    float best_d = 0.0f;
    for (uint32_t j = offsets[q]; j < offsets[q + 1]; ++j) {
        int cs = cand_status[j];
        if (cs == ST_SOME) {
            float d = cand[13ull * j + 12];
            uint32_t id = __float_as_uint(tris[3ull * items[j]].w);
            if (best_id == PB2_INVALID_U32 || d < best_d || (d == best_d && id < best_id)) { best_d = d; best_id = id; best_j = j; }
        } else if (cs >= ST_UNSUPPORTED && st == ST_NONE) st = cs;
    }
    float* o = out + 13ull * q;
    if (best_id != PB2_INVALID_U32) {
        for (int i = 0; i < 13; ++i) o[i] = cand[13ull * best_j + i];
        status[q] = (uint8_t)ST_SOME;
    } else {
        for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        status[q] = (uint8_t)st;
    }
    part[q] = best_id;
}

extern "C" {

// Device side of TriMesh-vs-shape contacts on device-resident arrays (n queries): loosened shape AABB in the mesh frame -> mesh Bvh
// intersect_aabb (CSR) -> every candidate triangle through the contact kernels -> per-query reduction to the smallest dist.
// flags = PAIR_LOCAL_FRAMES leaves the contacts in the frames of the mesh and of each shape (what a nested composite dispatch needs).
static int trimesh_contact_device(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* d_mpose, const pb2_shapes* shapes, const uint32_t* d_ids,
                                  const float* d_poses, uint32_t n, float prediction, float* d_out, uint8_t* d_status, uint32_t* d_part, uint32_t flags) {
    cudaStream_t st = ctx->stream;
    float* d_q = nullptr;
    uint32_t *d_off = nullptr, *d_items = nullptr, *d_ab = nullptr;
    float* d_cand = nullptr;
    uint8_t* d_cst = nullptr;
    int rc = PB2_OK;
    do {
        if (cudaMallocAsync((void**)&d_q, (size_t)n * 24, st) != cudaSuccess || cudaMallocAsync((void**)&d_off, ((size_t)n + 1) * 4, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_shapes: out of memory"); rc = PB2_ERR_CUDA; break;
        }
        k_mesh_query_aabbs<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, shapes->n, d_ids, d_poses, d_mpose, n, prediction,
                                                               (flags & PAIR_POS12_GIVEN) != 0, d_q);
        PB2_LAUNCHED(ctx);
        uint64_t total = 0;
        if ((rc = pb2_intersect_csr_device(ctx, &mesh->bvh, d_q, n, true, d_off, &d_items, &total)) != PB2_OK) break;
        if (total > 0xffffffffull) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_shapes: too many candidates"); rc = PB2_ERR_OVERFLOW; break; }
        if (total) {
            if (cudaMallocAsync((void**)&d_ab, total * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_cand, total * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, total, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_shapes: out of memory (%llu candidates)", (unsigned long long)total); rc = PB2_ERR_CUDA; break;
            }
            k_expand_candidates<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_items, n, d_ab);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_cand; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            if ((rc = run_contacts(ctx, shapes, nullptr, d_ids, d_mpose, d_poses, prediction, (uint32_t)total, sinks, d_ab, n, mesh->tris, mesh->nt, flags)) != PB2_OK) break;
        }
        k_mesh_reduce<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_items, mesh->tris, d_cand, d_cst, d_ids, shapes->n, n, d_out, d_status, d_part);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_shapes: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    if (d_q) cudaFreeAsync(d_q, st);
    if (d_off) cudaFreeAsync(d_off, st);
    if (d_items) cudaFreeAsync(d_items, st);
    if (d_ab) cudaFreeAsync(d_ab, st);
    if (d_cand) cudaFreeAsync(d_cand, st);
    if (d_cst) cudaFreeAsync(d_cst, st);
    return rc;
}

int pb2_trimesh_contact_shapes(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* mesh_pose7, const pb2_shapes* shapes, const uint32_t* shape_ids,
                               const float* poses7, uint32_t n, float prediction, pb2_contact* out, uint8_t* status, uint32_t* part, int mem) {
    if (!ctx || !mesh || !shapes || !mesh_pose7 || (n && (!shape_ids || !poses7 || !out || !status || !part))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_ids, *d_poses, *d_mpose;
    void *d_out, *d_status, *d_part;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape_ids, (size_t)n * 4, mem, &d_ids));
    PB2_CHECK(pb2_stage_in(ctx, 2, poses7, (size_t)n * 28, mem, &d_poses));
    PB2_CHECK(pb2_stage_in(ctx, 3, mesh_pose7, 28, mem, &d_mpose));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_status));
    PB2_CHECK(pb2_stage_out(ctx, 6, part, (size_t)n * 4, mem, &d_part));
    PB2_CHECK(trimesh_contact_device(ctx, mesh, (const float*)d_mpose, shapes, (const uint32_t*)d_ids, (const float*)d_poses, n, prediction, (float*)d_out,
                                     (uint8_t*)d_status, (uint32_t*)d_part, 0u));
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_status, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, part, d_part, (size_t)n * 4, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- Compound vs shapes
// query::contact with a Compound on one side (default_query_dispatcher.rs:338-351 -> contact_composite_shape_shape.rs:14-76):
// the candidate parts are those whose AABB (Compound::new, compound.rs:122-127: part.compute_aabb(part_pose)) intersects
// shape2.compute_aabb(pose12).loosened(prediction) — the leaf test of Bvh::intersect_aabb; compounds are a handful of parts,
// so the boxes are tested directly instead of walking a per-compound tree (same candidate set) —, every candidate runs
// through the contact kernels in the part's frame, and the smallest dist wins (equal dists: smallest part index).
struct pb2_compounds {
    uint32_t nc = 0, np = 0;
    uint32_t *first = nullptr, *count = nullptr, *part_shape = nullptr;
    float* part_pose = nullptr;
    float* part_aabb = nullptr;   // np x 6, in the compound's frame
    const pb2_shapes* shapes = nullptr;
};

__device__ __forceinline__ Iso7 compound_pose12(const float* pos_c, const float* pos_s, uint32_t k, bool second) {
    Iso7 pc = load_iso(pos_c + 7ull * k), ps = load_iso(pos_s + 7ull * k);
    return second ? iso_inverse(iso_inv_mul(ps, pc)) : iso_inv_mul(pc, ps);
}

// pass 0: counts[k] = number of candidate parts; pass 1 (offsets given): ab[j] = {global part, k}
template <bool FILL>
__global__ void k_compound_candidates(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float* __restrict__ points,
                                      uint32_t n_shapes, const uint32_t* __restrict__ comp_first, const uint32_t* __restrict__ comp_count, uint32_t nc,
                                      const float* __restrict__ part_aabb, const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c,
                                      const uint32_t* __restrict__ shape_ids, const float* __restrict__ pos_s, uint32_t n, float prediction, bool second,
                                      uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ ab) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t c = compound_id[k], sid = shape_ids[k], cnt = 0;
    if (c < nc && sid < n_shapes) {
        Iso7 pose12 = compound_pose12(pos_c, pos_s, k, second);
        V3 mn, mx;
        shape_aabb_dev(kinds[sid], params[sid], points, pose12, mn, mx);
        mn = mk3(mn.x + (-prediction), mn.y + (-prediction), mn.z + (-prediction));   // Aabb::loosened
        mx = mk3(mx.x + prediction, mx.y + prediction, mx.z + prediction);
        uint32_t f = comp_first[c], m = comp_count[c];
        uint32_t at = FILL ? offsets[k] : 0;
        for (uint32_t i = 0; i < m; ++i) {
            const float* b = part_aabb + 6ull * (f + i);
            // Aabb::intersects (aabb.rs:951-953): inclusive on every axis
            bool hit = b[0] <= mx.x && b[1] <= mx.y && b[2] <= mx.z && mn.x <= b[3] && mn.y <= b[4] && mn.z <= b[5];
            if (hit) {
                if (FILL) { ab[2ull * (at + cnt)] = f + i; ab[2ull * (at + cnt) + 1] = k; }
                cnt++;
            }
        }
    }
    if (!FILL) counts[k] = cnt;
}

__global__ void k_compound_part_aabbs(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float* __restrict__ points,
                                      const uint32_t* __restrict__ part_shape, const float* __restrict__ part_pose, uint32_t np, float* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= np) return;
    uint32_t sid = part_shape[i];
    V3 mn, mx;
    shape_aabb_dev(kinds[sid], params[sid], points, load_iso(part_pose + 7ull * i), mn, mx);
    float* o = out + 6ull * i;
    o[0] = mn.x; o[1] = mn.y; o[2] = mn.z; o[3] = mx.x; o[4] = mx.y; o[5] = mx.z;
}

__global__ void k_compound_reduce(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ab, const float* __restrict__ cand,
                                  const uint8_t* __restrict__ cand_status, const uint32_t* __restrict__ comp_first, uint32_t nc,
                                  const float* __restrict__ part_pose, const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c,
                                  const uint32_t* __restrict__ shape_ids, uint32_t n_shapes, const float* __restrict__ pos_s, uint32_t n, bool second,
                                  float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ part) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t c = compound_id[q];
    int st = (c >= nc || shape_ids[q] >= n_shapes) ? ST_UNSUPPORTED : ST_NONE;
    uint32_t best_j = 0;
    float best = 0.0f;
    bool needs_host = false;
    if (st == ST_NONE) {
        for (uint32_t j = offsets[q]; j < offsets[q + 1]; ++j) {
            if (cand_status[j] == ST_NEEDS_HOST) needs_host = true;
            if (cand_status[j] != ST_SOME) continue;
            float d = cand[13ull * j + 12];
            if (st != ST_SOME || d < best) { best = d; best_j = j; st = ST_SOME; }   // candidates are in part order: first minimum wins
        }
        if (needs_host) st = ST_NEEDS_HOST;
    }
    float* o = out + 13ull * q;
    if (st == ST_SOME) {
        const float* cj = cand + 13ull * best_j;
        uint32_t gp = ab[2ull * best_j];
        Iso7 pp = load_iso(part_pose + 7ull * gp);
        ContactOut ct;
        ct.p1 = iso_point(pp, mk3(cj[0], cj[1], cj[2]));   // Contact::transform1_by_mut(part_pos1)
        ct.n1 = iso_vec(pp, mk3(cj[6], cj[7], cj[8]));
        ct.p2 = mk3(cj[3], cj[4], cj[5]);
        ct.n2 = mk3(cj[9], cj[10], cj[11]);
        ct.dist = cj[12];
        Iso7 pc = load_iso(pos_c + 7ull * q), ps = load_iso(pos_s + 7ull * q);
        if (second) {   // Contact::flipped, then the user's (pos1, pos2) = (shape, compound)
            flip_contact(ct);
            ct.p1 = iso_point(ps, ct.p1); ct.n1 = iso_vec(ps, ct.n1); ct.p2 = iso_point(pc, ct.p2); ct.n2 = iso_vec(pc, ct.n2);
        } else {
            ct.p1 = iso_point(pc, ct.p1); ct.n1 = iso_vec(pc, ct.n1); ct.p2 = iso_point(ps, ct.p2); ct.n2 = iso_vec(ps, ct.n2);
        }
        store_contact(o, ct);
        part[q] = gp - comp_first[c];
    } else {
        for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        part[q] = PB2_INVALID_U32;
    }
    status[q] = (uint8_t)st;
}

// Compound (shape 1) against a TriMesh (shape 2): contact_composite_shape_shape(pos12, compound, trimesh)
// (contact_composite_shape_shape.rs:12-48). Every part whose AABB meets the mesh's root AABB, moved into the compound's frame and
// loosened, is one TriMesh-vs-shape query with the part's pose in the MESH frame, (part_pos.inv_mul(pos12)).inverse() (:27 and the
// inverse of contact_shape_composite_shape, :74). pass 0 counts, pass 1 (offsets given) fills {shape id, pose, owner}.
template <bool FILL>
__global__ void k_ct_candidates(const NodeWide* __restrict__ mesh_nodes, uint32_t mesh_leaves, const float* __restrict__ mesh_pose,
                                const uint32_t* __restrict__ comp_first, const uint32_t* __restrict__ comp_count, uint32_t nc,
                                const uint32_t* __restrict__ part_shape, const float* __restrict__ part_pose, const float* __restrict__ part_aabb,
                                const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c, uint32_t n, float prediction,
                                uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cand_shape,
                                float* __restrict__ cand_pose, uint32_t* __restrict__ cand_part) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t c = compound_id[k], cnt = 0;
    if (c < nc && mesh_leaves) {
        NodeWide w = mesh_nodes[0];   // Bvh::root_aabb (bvh_tree.rs:1991-1999)
        V3 amn = mk3(w.left.mnx, w.left.mny, w.left.mnz), amx = mk3(w.left.mxx, w.left.mxy, w.left.mxz);
        if (mesh_leaves > 1) {
            amn = vmin3(amn, mk3(w.right.mnx, w.right.mny, w.right.mnz));
            amx = vmax3(amx, mk3(w.right.mxx, w.right.mxy, w.right.mxz));
        }
        Iso7 pos12 = iso_inv_mul(load_iso(pos_c + 7ull * k), load_iso(mesh_pose));
        V3 ctr = iso_point(pos12, (amn + amx) * 0.5f);   // Aabb::transform_by(pos12).loosened(prediction)
        V3 he = iso_abs_vec(pos12, (amx - amn) * 0.5f);
        V3 lmn = ctr + (-he), lmx = ctr + he;
        lmn = mk3(lmn.x + (-prediction), lmn.y + (-prediction), lmn.z + (-prediction));
        lmx = mk3(lmx.x + prediction, lmx.y + prediction, lmx.z + prediction);
        uint32_t f = comp_first[c], m = comp_count[c];
        uint32_t at = FILL ? offsets[k] : 0;
        for (uint32_t i = 0; i < m; ++i) {
            if (!aabb6_intersects(part_aabb + 6ull * (f + i), lmn, lmx)) continue;
            if (FILL) {
                size_t j = (size_t)at + cnt;
                cand_shape[j] = part_shape[f + i];
                Iso7 pose = iso_inverse(iso_inv_mul(load_iso(part_pose + 7ull * (f + i)), pos12));
                float* o = cand_pose + 7 * j;
                o[0] = pose.q.i; o[1] = pose.q.j; o[2] = pose.q.k; o[3] = pose.q.w; o[4] = pose.t.x; o[5] = pose.t.y; o[6] = pose.t.z;
                cand_part[j] = i;
            }
            cnt++;
        }
    }
    if (!FILL) counts[k] = cnt;
}

// The compound side of the reduction (:29-41): candidates are in part order and the first strictly smaller dist wins; each candidate
// is the flipped TriMesh-vs-part contact; Contact::transform1_by_mut(part_pos) and the caller's two world poses finish it.
__global__ void k_ct_reduce(const uint32_t* __restrict__ offsets, const float* __restrict__ cand, const uint8_t* __restrict__ cand_status,
                            const uint32_t* __restrict__ cand_tri, const uint32_t* __restrict__ cand_part, const uint32_t* __restrict__ comp_first,
                            uint32_t nc, const float* __restrict__ part_pose, const uint32_t* __restrict__ compound_id,
                            const float* __restrict__ pos_c, const float* __restrict__ mesh_pose, uint32_t n, float* __restrict__ out,
                            uint8_t* __restrict__ status, uint32_t* __restrict__ parts) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t c = compound_id[q];
    int st = c >= nc ? ST_UNSUPPORTED : ST_NONE;
    uint32_t best_j = 0;
    float best = 0.0f;
    int worst = ST_NONE;
    if (st == ST_NONE) {
        for (uint32_t j = offsets[q]; j < offsets[q + 1]; ++j) {
            int cs = cand_status[j];
            if (cs >= ST_UNSUPPORTED && worst == ST_NONE) worst = cs;
            if (cs != ST_SOME) continue;
            float d = cand[13ull * j + 12];
            if (st != ST_SOME || d < best) { best = d; best_j = j; st = ST_SOME; }
        }
        if (worst != ST_NONE) st = worst;   // one part could not be decided on the device: neither can the minimum
    }
    float* o = out + 13ull * q;
    if (st == ST_SOME) {
        const float* cj = cand + 13ull * best_j;
        uint32_t pi = cand_part[best_j];
        Iso7 pp = load_iso(part_pose + 7ull * (comp_first[c] + pi));
        Iso7 pc = load_iso(pos_c + 7ull * q), pm = load_iso(mesh_pose);
        ContactOut ct;   // flipped: the part is shape 1
        ct.p1 = iso_point(pc, iso_point(pp, mk3(cj[3], cj[4], cj[5])));
        ct.n1 = iso_vec(pc, iso_vec(pp, mk3(cj[9], cj[10], cj[11])));
        ct.p2 = iso_point(pm, mk3(cj[0], cj[1], cj[2]));
        ct.n2 = iso_vec(pm, mk3(cj[6], cj[7], cj[8]));
        ct.dist = cj[12];
        store_contact(o, ct);
        parts[2ull * q] = pi;
        parts[2ull * q + 1] = cand_tri[best_j];
    } else {
        for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        parts[2ull * q] = parts[2ull * q + 1] = PB2_INVALID_U32;
    }
    status[q] = (uint8_t)st;
}


// TriMesh (shape 1) against a Compound (shape 2): contact_composite_shape_shape(pos12, trimesh, compound) (contact_composite_shape_shape.rs
// :12-48): every triangle whose leaf box meets the compound's local box (Compound::local_aabb, compound.rs:120-127) moved into the mesh
// frame and loosened is dispatched as contact(pos12, triangle, compound) = contact_shape_composite_shape (:63-76): the compound as the
// composite under pos12.inverse() — its parts whose box meets the triangle's loosened box, each contact(part_pos.inv_mul(pos21), part,
// triangle), first strictly smaller dist in part order, transform1_by(part_pos) — flipped; then the smallest dist over the triangles.
__global__ void k_tc_query_aabbs(const uint32_t* __restrict__ comp_first, const uint32_t* __restrict__ comp_count, uint32_t nc,
                                 const float* __restrict__ part_aabb, const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c,
                                 const float* __restrict__ mesh_pose, uint32_t n, float prediction, float* __restrict__ out) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float* o = out + 6ull * k;
    uint32_t c = compound_id[k];
    if (c >= nc) { o[0] = o[1] = o[2] = FLT_MAX; o[3] = o[4] = o[5] = -FLT_MAX; return; }
    uint32_t f = comp_first[c], m = comp_count[c];
    V3 amn = mk3(FLT_MAX, FLT_MAX, FLT_MAX), amx = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);
    for (uint32_t j = 0; j < m; ++j) {
        const float* b = part_aabb + 6ull * (f + j);
        amn = vmin3(amn, mk3(b[0], b[1], b[2]));
        amx = vmax3(amx, mk3(b[3], b[4], b[5]));
    }
    Iso7 pos12 = iso_inv_mul(load_iso(mesh_pose), load_iso(pos_c + 7ull * k));
    V3 ctr = iso_point(pos12, (amn + amx) * 0.5f), he = iso_abs_vec(pos12, (amx - amn) * 0.5f);
    V3 lmn = ctr + (-he), lmx = ctr + he;
    o[0] = lmn.x + (-prediction); o[1] = lmn.y + (-prediction); o[2] = lmn.z + (-prediction);
    o[3] = lmx.x + prediction; o[4] = lmx.y + prediction; o[5] = lmx.z + prediction;
}

// e = one (query, triangle) couple of the CSR (kt[2e] = triangle position, kt[2e+1] = query). pass 0 counts the parts that meet the
// triangle's loosened box in the compound's frame, pass 1 (offsets given) fills one candidate per such part.
template <bool FILL>
__global__ void k_tc_candidates(const uint32_t* __restrict__ kt, uint32_t total, const float4* __restrict__ tris, const uint32_t* __restrict__ comp_first,
                                const uint32_t* __restrict__ comp_count, const uint32_t* __restrict__ part_shape, const float* __restrict__ part_pose,
                                const float* __restrict__ part_aabb, const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c,
                                const float* __restrict__ mesh_pose, float prediction, uint32_t* __restrict__ counts,
                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ cand_shape, float4* __restrict__ cand_tri,
                                float* __restrict__ cand_pose, uint32_t* __restrict__ cand_part, uint32_t* __restrict__ cand_ab) {
    uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    uint32_t tp = kt[2ull * e], k = kt[2ull * e + 1];
    uint32_t c = compound_id[k];
    Iso7 pos21 = iso_inverse(iso_inv_mul(load_iso(mesh_pose), load_iso(pos_c + 7ull * k)));
    float4 fa = tris[3ull * tp], fb = tris[3ull * tp + 1], fc = tris[3ull * tp + 2];
    V3 qa = iso_point(pos21, mk3(fa.x, fa.y, fa.z)), qb = iso_point(pos21, mk3(fb.x, fb.y, fb.z)), qc = iso_point(pos21, mk3(fc.x, fc.y, fc.z));
    V3 mn = mk3(fminf(fminf(qa.x, qb.x), qc.x), fminf(fminf(qa.y, qb.y), qc.y), fminf(fminf(qa.z, qb.z), qc.z));
    V3 mx = mk3(fmaxf(fmaxf(qa.x, qb.x), qc.x), fmaxf(fmaxf(qa.y, qb.y), qc.y), fmaxf(fmaxf(qa.z, qb.z), qc.z));
    mn = mk3(mn.x + (-prediction), mn.y + (-prediction), mn.z + (-prediction));
    mx = mk3(mx.x + prediction, mx.y + prediction, mx.z + prediction);
    uint32_t f = comp_first[c], m = comp_count[c], cnt = 0;
    uint32_t at = FILL ? offsets[e] : 0;
    for (uint32_t j = 0; j < m; ++j) {
        if (!aabb6_intersects(part_aabb + 6ull * (f + j), mn, mx)) continue;
        if (FILL) {
            size_t ci = (size_t)at + cnt;
            cand_shape[ci] = part_shape[f + j];
            cand_tri[3 * ci] = fa; cand_tri[3 * ci + 1] = fb; cand_tri[3 * ci + 2] = fc;
            Iso7 pose = iso_inv_mul(load_iso(part_pose + 7ull * (f + j)), pos21);
            float* o = cand_pose + 7 * ci;
            o[0] = pose.q.i; o[1] = pose.q.j; o[2] = pose.q.k; o[3] = pose.q.w; o[4] = pose.t.x; o[5] = pose.t.y; o[6] = pose.t.z;
            cand_part[ci] = j;
            cand_ab[2 * ci] = (uint32_t)ci; cand_ab[2 * ci + 1] = (uint32_t)ci;
        }
        cnt++;
    }
    if (!FILL) counts[e] = cnt;
}

__global__ void k_tc_reduce(const uint32_t* __restrict__ q_off, const uint32_t* __restrict__ kt, const uint32_t* __restrict__ e_off,
                            const float4* __restrict__ tris, const float* __restrict__ cand, const uint8_t* __restrict__ cand_status,
                            const uint32_t* __restrict__ cand_part, const uint32_t* __restrict__ comp_first, uint32_t nc,
                            const float* __restrict__ part_pose, const uint32_t* __restrict__ compound_id, const float* __restrict__ pos_c,
                            const float* __restrict__ mesh_pose, uint32_t n, float* __restrict__ out, uint8_t* __restrict__ status,
                            uint32_t* __restrict__ parts) {
    uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= n) return;
    uint32_t c = compound_id[q];
    int st = c >= nc ? ST_UNSUPPORTED : ST_NONE, worst = ST_NONE;
    uint32_t best_j = 0, best_tri = PB2_INVALID_U32;
    float best = 0.0f;
    if (st == ST_NONE) {
        for (uint32_t e = q_off[q]; e < q_off[q + 1]; ++e) {
            // the compound's answer for this triangle: first strictly smaller dist in part order
            bool have = false;
            uint32_t ej = 0;
            float ed = 0.0f;
            for (uint32_t j = e_off[e]; j < e_off[e + 1]; ++j) {
                int cs = cand_status[j];
                if (cs >= ST_UNSUPPORTED && worst == ST_NONE) worst = cs;
                if (cs != ST_SOME) continue;
                float d = cand[13ull * j + 12];
                if (!have || d < ed) { have = true; ed = d; ej = j; }
            }
            if (!have) continue;
            uint32_t id = __float_as_uint(tris[3ull * kt[2ull * e]].w);
            if (best_tri == PB2_INVALID_U32 || ed < best || (ed == best && id < best_tri)) { best = ed; best_j = ej; best_tri = id; }
        }
        if (best_tri != PB2_INVALID_U32) st = ST_SOME;
        if (worst != ST_NONE) st = worst;
    }
    float* o = out + 13ull * q;
    if (st == ST_SOME) {
        const float* cj = cand + 13ull * best_j;   // contact(part, triangle) in the part's and the mesh's frames
        uint32_t pi = cand_part[best_j];
        Iso7 pp = load_iso(part_pose + 7ull * (comp_first[c] + pi));
        Iso7 pc = load_iso(pos_c + 7ull * q), pm = load_iso(mesh_pose);
        ContactOut ct;   // transform1_by(part_pos), flipped: the mesh is shape 1
        ct.p1 = iso_point(pm, mk3(cj[3], cj[4], cj[5]));
        ct.n1 = iso_vec(pm, mk3(cj[9], cj[10], cj[11]));
        ct.p2 = iso_point(pc, iso_point(pp, mk3(cj[0], cj[1], cj[2])));
        ct.n2 = iso_vec(pc, iso_vec(pp, mk3(cj[6], cj[7], cj[8])));
        ct.dist = cj[12];
        store_contact(o, ct);
        parts[2ull * q] = pi;
        parts[2ull * q + 1] = best_tri;
    } else {
        for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        parts[2ull * q] = parts[2ull * q + 1] = PB2_INVALID_U32;
    }
    status[q] = (uint8_t)st;
}

static int trimesh_contact_compounds(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* d_cid, const float* d_pc, const pb2_trimesh* mesh,
                                     const float* d_mpose, uint32_t n, float prediction, float* d_out, uint8_t* d_status, uint32_t* d_parts) {
    const pb2_shapes* shapes = compounds->shapes;
    cudaStream_t st = ctx->stream;
    float *d_q = nullptr, *d_cpose = nullptr, *d_cand = nullptr;
    uint32_t *d_off = nullptr, *d_items = nullptr, *d_kt = nullptr, *d_cnt = nullptr, *d_eoff = nullptr, *d_cs = nullptr, *d_cpart = nullptr, *d_cab = nullptr;
    float4* d_ctri = nullptr;
    uint8_t* d_cst = nullptr;
    void* d_tmp = nullptr;
    int rc = PB2_OK;
    do {
        if (cudaMallocAsync((void**)&d_q, (size_t)n * 24, st) != cudaSuccess || cudaMallocAsync((void**)&d_off, ((size_t)n + 1) * 4, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: out of memory"); rc = PB2_ERR_CUDA; break;
        }
        k_tc_query_aabbs<<<pb2_blocks(n, 128), 128, 0, st>>>(compounds->first, compounds->count, compounds->nc, compounds->part_aabb, d_cid, d_pc, d_mpose,
                                                             n, prediction, d_q);
        PB2_LAUNCHED(ctx);
        uint64_t total64 = 0;
        if ((rc = pb2_intersect_csr_device(ctx, &mesh->bvh, d_q, n, true, d_off, &d_items, &total64)) != PB2_OK) break;
        if (total64 > 0x7fffffffull) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: too many candidates"); rc = PB2_ERR_OVERFLOW; break; }
        const uint32_t total = (uint32_t)total64;
        uint32_t ncand = 0;
        if (total) {
            size_t cub_bytes = 0;
            cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)total + 1, st);
            if (cudaMallocAsync((void**)&d_kt, (size_t)total * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_cnt, ((size_t)total + 1) * 4, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_eoff, ((size_t)total + 1) * 4, st) != cudaSuccess || cudaMallocAsync(&d_tmp, cub_bytes, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: out of memory (%u triangle candidates)", total); rc = PB2_ERR_CUDA; break;
            }
            k_expand_candidates<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_items, n, d_kt);
            PB2_LAUNCHED(ctx);
            cudaMemsetAsync(d_cnt + total, 0, 4, st);
            k_tc_candidates<false><<<pb2_blocks(total, 128), 128, 0, st>>>(d_kt, total, mesh->tris, compounds->first, compounds->count, compounds->part_shape,
                compounds->part_pose, compounds->part_aabb, d_cid, d_pc, d_mpose, prediction, d_cnt, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr);
            PB2_LAUNCHED(ctx);
            cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, (const uint32_t*)d_cnt, d_eoff, (int)total + 1, st);
            ctx->launches += 1;
            cudaMemcpyAsync(&ncand, d_eoff + total, 4, cudaMemcpyDeviceToHost, st);
            if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: candidate pass failed"); rc = PB2_ERR_CUDA; break; }
        }
        if (ncand) {
            if (cudaMallocAsync((void**)&d_cs, (size_t)ncand * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_cpart, (size_t)ncand * 4, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cab, (size_t)ncand * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_ctri, (size_t)ncand * 48, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cpose, (size_t)ncand * 28, st) != cudaSuccess || cudaMallocAsync((void**)&d_cand, (size_t)ncand * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, ncand, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: out of memory (%u part candidates)", ncand); rc = PB2_ERR_CUDA; break;
            }
            k_tc_candidates<true><<<pb2_blocks(total, 128), 128, 0, st>>>(d_kt, total, mesh->tris, compounds->first, compounds->count, compounds->part_shape,
                compounds->part_pose, compounds->part_aabb, d_cid, d_pc, d_mpose, prediction, nullptr, d_eoff, d_cs, d_ctri, d_cpose, d_cpart, d_cab);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_cand; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            // contact(part_pos.inv_mul(pos21), part, triangle): the part from the shape table, the triangle as shape 2, local frames
            if ((rc = run_contacts(ctx, shapes, d_cs, nullptr, d_cpose, d_cpose, prediction, ncand, sinks, d_cab, ncand, nullptr, 0,
                                   PAIR_LOCAL_FRAMES | PAIR_POS12_GIVEN, nullptr, 0, nullptr, 3, d_ctri)) != PB2_OK) break;
        }
        k_tc_reduce<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_kt, d_eoff, mesh->tris, d_cand, d_cst, d_cpart, compounds->first, compounds->nc,
            compounds->part_pose, d_cid, d_pc, d_mpose, n, d_out, d_status, d_parts);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "trimesh_contact_compounds: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    void* frees[] = {d_q, d_off, d_items, d_kt, d_cnt, d_eoff, d_tmp, d_cs, d_cpart, d_cab, d_ctri, d_cpose, d_cand, d_cst};
    for (void* p : frees) if (p) cudaFreeAsync(p, st);
    return rc;
}


extern "C" {

int pb2_compounds_create(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* comp_first, const uint32_t* comp_count, uint32_t nc,
                         const uint32_t* part_shape, const float* part_pose7, uint32_t np, pb2_compounds** out) {
    if (!ctx || !shapes || !out || !comp_first || !comp_count || !part_shape || !part_pose7 || nc == 0 || np == 0) return PB2_ERR_INVALID;
    *out = nullptr;
    for (uint32_t c = 0; c < nc; ++c) {
        // Compound::new asserts !shapes.is_empty() (compound.rs:114-117)
        if (comp_count[c] == 0 || (uint64_t)comp_first[c] + comp_count[c] > np) PB2_FAIL(ctx, PB2_ERR_INVALID, "compounds_create: empty compound or part range out of bounds");
    }
    for (uint32_t i = 0; i < np; ++i)
        if (part_shape[i] >= shapes->n) PB2_FAIL(ctx, PB2_ERR_INVALID, "compounds_create: part shape id out of range");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    pb2_compounds* cp = new pb2_compounds();
    cp->nc = nc; cp->np = np; cp->shapes = shapes;
    cudaStream_t st = ctx->stream;
    bool ok = cudaMalloc((void**)&cp->first, (size_t)nc * 4) == cudaSuccess && cudaMalloc((void**)&cp->count, (size_t)nc * 4) == cudaSuccess &&
              cudaMalloc((void**)&cp->part_shape, (size_t)np * 4) == cudaSuccess && cudaMalloc((void**)&cp->part_pose, (size_t)np * 28) == cudaSuccess &&
              cudaMalloc((void**)&cp->part_aabb, (size_t)np * 24) == cudaSuccess;
    if (ok) {
        cudaMemcpyAsync(cp->first, comp_first, (size_t)nc * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(cp->count, comp_count, (size_t)nc * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(cp->part_shape, part_shape, (size_t)np * 4, cudaMemcpyHostToDevice, st);
        cudaMemcpyAsync(cp->part_pose, part_pose7, (size_t)np * 28, cudaMemcpyHostToDevice, st);
        k_compound_part_aabbs<<<pb2_blocks(np, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, cp->part_shape, cp->part_pose, np,
                                                                   cp->part_aabb);
        PB2_LAUNCHED(ctx);
        ok = cudaStreamSynchronize(st) == cudaSuccess;
    }
    if (!ok) {
        cudaFree(cp->first); cudaFree(cp->count); cudaFree(cp->part_shape); cudaFree(cp->part_pose); cudaFree(cp->part_aabb);
        delete cp;
        PB2_FAIL(ctx, PB2_ERR_CUDA, "compounds_create: device allocation or upload failed");
    }
    *out = cp;
    return PB2_OK;
}

int pb2_compounds_destroy(pb2_ctx* ctx, pb2_compounds* cp) {
    if (!ctx || !cp) return PB2_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(cp->first); cudaFree(cp->count); cudaFree(cp->part_shape); cudaFree(cp->part_pose); cudaFree(cp->part_aabb);
    delete cp;
    return PB2_OK;
}

int pb2_compound_contact_shapes(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* compound_ids, const float* compound_poses7,
                                const uint32_t* shape_ids, const float* shape_poses7, uint32_t n, float prediction, int compound_second,
                                pb2_contact* out, uint8_t* status, uint32_t* part, int mem) {
    if (!ctx || !compounds || (n && (!compound_ids || !compound_poses7 || !shape_ids || !shape_poses7 || !out || !status || !part))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    const pb2_shapes* shapes = compounds->shapes;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_cid, *d_pc, *d_sid, *d_ps;
    void *d_out, *d_status, *d_part;
    PB2_CHECK(pb2_stage_in(ctx, 0, compound_ids, (size_t)n * 4, mem, &d_cid));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape_ids, (size_t)n * 4, mem, &d_sid));
    PB2_CHECK(pb2_stage_in(ctx, 2, compound_poses7, (size_t)n * 28, mem, &d_pc));
    PB2_CHECK(pb2_stage_in(ctx, 3, shape_poses7, (size_t)n * 28, mem, &d_ps));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_status));
    PB2_CHECK(pb2_stage_out(ctx, 6, part, (size_t)n * 4, mem, &d_part));
    const bool second = compound_second != 0;
    uint32_t *d_cnt = nullptr, *d_off = nullptr, *d_ab = nullptr;
    float* d_cand = nullptr;
    uint8_t* d_cst = nullptr;
    void* d_tmp = nullptr;
    int rc = PB2_OK;
    do {
        size_t cub_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n + 1, st);
        if (cudaMallocAsync((void**)&d_cnt, ((size_t)n + 1) * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_off, ((size_t)n + 1) * 4, st) != cudaSuccess ||
            cudaMallocAsync(&d_tmp, cub_bytes, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "compound_contact_shapes: out of memory"); rc = PB2_ERR_CUDA; break;
        }
        cudaMemsetAsync(d_cnt + n, 0, 4, st);
        k_compound_candidates<false><<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, shapes->n, compounds->first,
            compounds->count, compounds->nc, compounds->part_aabb, (const uint32_t*)d_cid, (const float*)d_pc, (const uint32_t*)d_sid,
            (const float*)d_ps, n, prediction, second, d_cnt, nullptr, nullptr);
        PB2_LAUNCHED(ctx);
        cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, (const uint32_t*)d_cnt, d_off, (int)n + 1, st);
        ctx->launches += 1;
        uint32_t total = 0;
        cudaMemcpyAsync(&total, d_off + n, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_shapes: candidate pass failed"); rc = PB2_ERR_CUDA; break; }
        if (total) {
            if (cudaMallocAsync((void**)&d_ab, (size_t)total * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_cand, (size_t)total * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, total, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "compound_contact_shapes: out of memory (%u candidates)", total); rc = PB2_ERR_CUDA; break;
            }
            k_compound_candidates<true><<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, shapes->n, compounds->first,
                compounds->count, compounds->nc, compounds->part_aabb, (const uint32_t*)d_cid, (const float*)d_pc, (const uint32_t*)d_sid,
                (const float*)d_ps, n, prediction, second, nullptr, d_off, d_ab);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_cand; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            if ((rc = run_contacts(ctx, shapes, compounds->part_shape, (const uint32_t*)d_sid, (const float*)d_pc, (const float*)d_ps, prediction, total,
                                   sinks, d_ab, n, nullptr, 0, second ? PAIR_COMPOUND_SECOND : 0u, compounds->part_pose, compounds->np)) != PB2_OK) break;
        }
        k_compound_reduce<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_ab, d_cand, d_cst, compounds->first, compounds->nc, compounds->part_pose,
            (const uint32_t*)d_cid, (const float*)d_pc, (const uint32_t*)d_sid, shapes->n, (const float*)d_ps, n, second, (float*)d_out,
            (uint8_t*)d_status, (uint32_t*)d_part);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_shapes: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    if (d_cnt) cudaFreeAsync(d_cnt, st);
    if (d_off) cudaFreeAsync(d_off, st);
    if (d_tmp) cudaFreeAsync(d_tmp, st);
    if (d_ab) cudaFreeAsync(d_ab, st);
    if (d_cand) cudaFreeAsync(d_cand, st);
    if (d_cst) cudaFreeAsync(d_cst, st);
    if (rc != PB2_OK) return rc;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_status, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, part, d_part, (size_t)n * 4, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

int pb2_compound_contact_trimesh(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* compound_ids, const float* compound_poses7,
                                 const pb2_trimesh* mesh, const float* mesh_pose7, uint32_t n, float prediction, int mesh_first, pb2_contact* out,
                                 uint8_t* status, uint32_t* parts, int mem) {
    if (!ctx || !compounds || !mesh || !mesh_pose7 || (n && (!compound_ids || !compound_poses7 || !out || !status || !parts))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    const pb2_shapes* shapes = compounds->shapes;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_cid, *d_pc, *d_mpose;
    void *d_out, *d_status, *d_parts;
    PB2_CHECK(pb2_stage_in(ctx, 0, compound_ids, (size_t)n * 4, mem, &d_cid));
    PB2_CHECK(pb2_stage_in(ctx, 2, compound_poses7, (size_t)n * 28, mem, &d_pc));
    PB2_CHECK(pb2_stage_in(ctx, 3, mesh_pose7, 28, mem, &d_mpose));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_status));
    PB2_CHECK(pb2_stage_out(ctx, 6, parts, (size_t)n * 8, mem, &d_parts));
    if (mesh_first) {
        PB2_CHECK(trimesh_contact_compounds(ctx, compounds, (const uint32_t*)d_cid, (const float*)d_pc, mesh, (const float*)d_mpose, n, prediction,
                                            (float*)d_out, (uint8_t*)d_status, (uint32_t*)d_parts));
        PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
        PB2_CHECK(pb2_stage_back(ctx, status, d_status, (size_t)n, mem));
        PB2_CHECK(pb2_stage_back(ctx, parts, d_parts, (size_t)n * 8, mem));
        if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
        return PB2_OK;
    }
    uint32_t *d_cnt = nullptr, *d_off = nullptr, *d_cs = nullptr, *d_cpart = nullptr, *d_ctri = nullptr;
    float *d_cpose = nullptr, *d_cand = nullptr, *d_ident = nullptr;
    uint8_t* d_cst = nullptr;
    void* d_tmp = nullptr;
    int rc = PB2_OK;
    do {
        size_t cub_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n + 1, st);
        if (cudaMallocAsync((void**)&d_cnt, ((size_t)n + 1) * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_off, ((size_t)n + 1) * 4, st) != cudaSuccess ||
            cudaMallocAsync(&d_tmp, cub_bytes, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "compound_contact_trimesh: out of memory"); rc = PB2_ERR_CUDA; break;
        }
        cudaMemsetAsync(d_cnt + n, 0, 4, st);
        k_ct_candidates<false><<<pb2_blocks(n, 128), 128, 0, st>>>(mesh->bvh.nodes, mesh->bvh.n_leaves, (const float*)d_mpose, compounds->first,
            compounds->count, compounds->nc, compounds->part_shape, compounds->part_pose, compounds->part_aabb, (const uint32_t*)d_cid,
            (const float*)d_pc, n, prediction, d_cnt, nullptr, nullptr, nullptr, nullptr);
        PB2_LAUNCHED(ctx);
        cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, (const uint32_t*)d_cnt, d_off, (int)n + 1, st);
        ctx->launches += 1;
        uint32_t total = 0;
        cudaMemcpyAsync(&total, d_off + n, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_trimesh: candidate pass failed"); rc = PB2_ERR_CUDA; break; }
        if (total) {
            if (cudaMallocAsync((void**)&d_cs, (size_t)total * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_cpart, (size_t)total * 4, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_ctri, (size_t)total * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_cpose, (size_t)total * 28, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cand, (size_t)total * 52, st) != cudaSuccess || cudaMallocAsync((void**)&d_cst, total, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_ident, 28, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "compound_contact_trimesh: out of memory (%u candidate parts)", total); rc = PB2_ERR_CUDA; break;
            }
            k_ct_candidates<true><<<pb2_blocks(n, 128), 128, 0, st>>>(mesh->bvh.nodes, mesh->bvh.n_leaves, (const float*)d_mpose, compounds->first,
                compounds->count, compounds->nc, compounds->part_shape, compounds->part_pose, compounds->part_aabb, (const uint32_t*)d_cid,
                (const float*)d_pc, n, prediction, nullptr, d_off, d_cs, d_cpose, d_cpart);
            PB2_LAUNCHED(ctx);
            static const float ident[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
            cudaMemcpyAsync(d_ident, ident, 28, cudaMemcpyHostToDevice, st);
            // every (query, part) is one TriMesh-vs-shape query in the mesh's own frame; contacts stay in the mesh / part frames
            if ((rc = trimesh_contact_device(ctx, mesh, d_ident, shapes, d_cs, d_cpose, total, prediction, d_cand, d_cst, d_ctri,
                                             PAIR_LOCAL_FRAMES | PAIR_POS12_GIVEN)) != PB2_OK) break;
        }
        k_ct_reduce<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_cand, d_cst, d_ctri, d_cpart, compounds->first, compounds->nc, compounds->part_pose,
            (const uint32_t*)d_cid, (const float*)d_pc, (const float*)d_mpose, n, (float*)d_out, (uint8_t*)d_status, (uint32_t*)d_parts);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_trimesh: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    void* frees[] = {d_cnt, d_off, d_tmp, d_cs, d_cpart, d_ctri, d_cpose, d_cand, d_cst, d_ident};
    for (void* p : frees) if (p) cudaFreeAsync(p, st);
    if (rc != PB2_OK) return rc;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_status, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, parts, d_parts, (size_t)n * 8, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------- Compound vs Compound
// query::contact between two Compounds of one table (nested composite arms, see compound_pair.cuh): candidates = (part of compound
// 1, part of compound 2) pairs passing the two AABB tests of the reference, materialised as plain leaf problems for the contact
// kernels (local frames), then reduced per pair. pass 0 counts, pass 1 fills.
static CompoundTable compound_table(const pb2_compounds* c) {
    CompoundTable T;
    T.first = c->first; T.count = c->count; T.part_shape = c->part_shape; T.part_pose = c->part_pose; T.part_aabb = c->part_aabb; T.nc = c->nc;
    return T;
}

template <bool FILL>
__global__ void __launch_bounds__(128) k_cc_candidates(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float* __restrict__ points,
                                                       CompoundTable T, const uint32_t* __restrict__ id1, const float* __restrict__ pos1,
                                                       const uint32_t* __restrict__ id2, const float* __restrict__ pos2, uint32_t n, float prediction,
                                                       uint32_t* __restrict__ counts, const uint32_t* __restrict__ offsets, uint32_t* __restrict__ ij,
                                                       uint32_t* __restrict__ cs1, uint32_t* __restrict__ cs2, float* __restrict__ cp1,
                                                       float* __restrict__ cp2) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t a = id1[k], b = id2[k], cnt = 0;
    if (a < T.nc && b < T.nc) {
        Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));   // contact_shape_shape.rs:130
        cnt = cc_candidates<FILL>(kinds, params, points, T, a, b, pos12, prediction, FILL ? offsets[k] : 0u, ij, cs1, cs2, cp1, cp2);
    }
    if (!FILL) counts[k] = cnt;
}

__global__ void __launch_bounds__(128) k_cc_reduce(const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ ij, const float* __restrict__ cand,
                                                   const uint8_t* __restrict__ cst, CompoundTable T, const uint32_t* __restrict__ id1,
                                                   const float* __restrict__ pos1, const uint32_t* __restrict__ id2, const float* __restrict__ pos2,
                                                   uint32_t n, float* __restrict__ out, uint8_t* __restrict__ status, uint32_t* __restrict__ parts) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    uint32_t a = id1[k], b = id2[k];
    float* o = out + 13ull * k;
    if (a >= T.nc || b >= T.nc) {
        for (int d = 0; d < 13; ++d) o[d] = 0.0f;
        parts[2ull * k] = PB2_INVALID_U32; parts[2ull * k + 1] = PB2_INVALID_U32;
        status[k] = ST_UNSUPPORTED;
        return;
    }
    status[k] = (uint8_t)cc_reduce(ij, cand, cst, offsets[k], offsets[k + 1], T, a, b, load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k), o,
                                   parts + 2ull * k);
}

extern "C" int pb2_compound_contact_compounds(pb2_ctx* ctx, const pb2_compounds* compounds, const uint32_t* ids1, const float* poses1,
                                              const uint32_t* ids2, const float* poses2, uint32_t n, float prediction, pb2_contact* out,
                                              uint8_t* status, uint32_t* parts, int mem) {
    if (!ctx || !compounds || (n && (!ids1 || !poses1 || !ids2 || !poses2 || !out || !status || !parts))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    const pb2_shapes* shapes = compounds->shapes;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_i1, *d_p1, *d_i2, *d_p2;
    void *d_out, *d_status, *d_parts;
    PB2_CHECK(pb2_stage_in(ctx, 0, ids1, (size_t)n * 4, mem, &d_i1));
    PB2_CHECK(pb2_stage_in(ctx, 1, ids2, (size_t)n * 4, mem, &d_i2));
    PB2_CHECK(pb2_stage_in(ctx, 2, poses1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, poses2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_out(ctx, 4, out, (size_t)n * 52, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, status, (size_t)n, mem, &d_status));
    PB2_CHECK(pb2_stage_out(ctx, 6, parts, (size_t)n * 8, mem, &d_parts));
    CompoundTable T = compound_table(compounds);
    uint32_t *d_cnt = nullptr, *d_off = nullptr, *d_ij = nullptr, *d_cs1 = nullptr, *d_cs2 = nullptr;
    float *d_cp1 = nullptr, *d_cp2 = nullptr, *d_cand = nullptr;
    uint8_t* d_cst = nullptr;
    void* d_tmp = nullptr;
    int rc = PB2_OK;
    do {
        size_t cub_bytes = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, cub_bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n + 1, st);
        if (cudaMallocAsync((void**)&d_cnt, ((size_t)n + 1) * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_off, ((size_t)n + 1) * 4, st) != cudaSuccess ||
            cudaMallocAsync(&d_tmp, cub_bytes, st) != cudaSuccess) {
            snprintf(ctx->err, sizeof(ctx->err), "compound_contact_compounds: out of memory"); rc = PB2_ERR_CUDA; break;
        }
        cudaMemsetAsync(d_cnt + n, 0, 4, st);
        k_cc_candidates<false><<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, T, (const uint32_t*)d_i1, (const float*)d_p1,
                                                                   (const uint32_t*)d_i2, (const float*)d_p2, n, prediction, d_cnt, nullptr, nullptr, nullptr,
                                                                   nullptr, nullptr, nullptr);
        PB2_LAUNCHED(ctx);
        cub::DeviceScan::ExclusiveSum(d_tmp, cub_bytes, (const uint32_t*)d_cnt, d_off, (int)n + 1, st);
        ctx->launches += 1;
        uint32_t total = 0;
        cudaMemcpyAsync(&total, d_off + n, 4, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_compounds: candidate pass failed"); rc = PB2_ERR_CUDA; break; }
        if (total) {
            if (cudaMallocAsync((void**)&d_ij, (size_t)total * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_cs1, (size_t)total * 4, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cs2, (size_t)total * 4, st) != cudaSuccess || cudaMallocAsync((void**)&d_cp1, (size_t)total * 28, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cp2, (size_t)total * 28, st) != cudaSuccess || cudaMallocAsync((void**)&d_cand, (size_t)total * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, total, st) != cudaSuccess) {
                snprintf(ctx->err, sizeof(ctx->err), "compound_contact_compounds: out of memory (%u candidates)", total); rc = PB2_ERR_CUDA; break;
            }
            k_cc_candidates<true><<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points, T, (const uint32_t*)d_i1,
                                                                      (const float*)d_p1, (const uint32_t*)d_i2, (const float*)d_p2, n, prediction, nullptr,
                                                                      d_off, d_ij, d_cs1, d_cs2, d_cp1, d_cp2);
            PB2_LAUNCHED(ctx);
            OutSinks sinks;
            sinks.dense = d_cand; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
            sinks.compact_count = nullptr; sinks.some_count = nullptr;
            // each candidate is a plain pair: contact(part_pos2[j].inv_mul(pose), part_j, part_i), left in the parts' frames
            if ((rc = run_contacts(ctx, shapes, d_cs1, d_cs2, d_cp1, d_cp2, prediction, total, sinks, nullptr, 0, nullptr, 0, PAIR_LOCAL_FRAMES)) != PB2_OK) break;
        }
        k_cc_reduce<<<pb2_blocks(n, 128), 128, 0, st>>>(d_off, d_ij, d_cand, d_cst, T, (const uint32_t*)d_i1, (const float*)d_p1, (const uint32_t*)d_i2,
                                                        (const float*)d_p2, n, (float*)d_out, (uint8_t*)d_status, (uint32_t*)d_parts);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "compound_contact_compounds: launch failed"); rc = PB2_ERR_CUDA; break; }
    } while (0);
    void* frees[] = {d_cnt, d_off, d_tmp, d_ij, d_cs1, d_cs2, d_cp1, d_cp2, d_cand, d_cst};
    for (void* f : frees) if (f) cudaFreeAsync(f, st);
    if (rc != PB2_OK) return rc;
    PB2_CHECK(pb2_stage_back(ctx, out, d_out, (size_t)n * 52, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_status, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, parts, d_parts, (size_t)n * 8, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(st));
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------- contact manifolds (closed-form arms)
// QueryDispatcher::contact_manifolds for Ball / Cuboid pairs, first frame (empty incoming manifold, so try_update_contacts
// and match_contacts are no-ops): DefaultQueryDispatcher::contact_manifold_convex_convex (default_query_dispatcher.rs:748-782)
// -> contact_manifolds_ball_ball.rs:17-57, contact_manifolds_convex_ball.rs:42-145 (Cuboid as shape 1),
// contact_manifolds_cuboid_cuboid.rs:19-107 + sat_cuboid_cuboid.rs:5-110 + Cuboid::support_face (cuboid.rs:267-354) +
// PolygonalFeature::contacts_face_face / closest_points_line2d (polygonal_feature3d.rs:215-439) + Vector3::orthonormal_basis
// (utils/wops.rs:92-110). Pairs with a ConvexPolyhedron need its face/edge topology (pfm_pfm) and get status 2.
enum { MAN_OK = 0, MAN_UNSUPPORTED = 2, MAN_OVERFLOW = 4 };
#define PK_VERTEX(c) ((1u << 30) | (c))
#define PK_EDGE(c) ((2u << 30) | (c))
#define PK_FACE(c) ((3u << 30) | (c))

struct ManifoldOut {
    float* pts;         // this pair's max_points x 9 words
    uint32_t max_points, count;
    bool overflow;
    __device__ __forceinline__ void push(V3 p1, V3 p2, uint32_t f1, uint32_t f2, float dist, bool flipped) {
        if (count >= max_points) { overflow = true; return; }
        float* o = pts + 9ull * count;
        if (flipped) { V3 t = p1; p1 = p2; p2 = t; uint32_t u = f1; f1 = f2; f2 = u; }   // TrackedContact::flipped
        o[0] = p1.x; o[1] = p1.y; o[2] = p1.z; o[3] = p2.x; o[4] = p2.y; o[5] = p2.z; o[6] = dist;
        o[7] = __uint_as_float(f1); o[8] = __uint_as_float(f2);
        count++;
    }
};

// cub_support / sat_normal_oneway / sat_edge_twoway: near the top of this file (shared with the distance / intersection_test arms)

struct PolyFace { V3 v[4]; uint32_t vids[4], eids[4], fid; int n; };
__constant__ uint8_t c_face_vids[3][2][4] = {{{0, 2, 3, 1}, {4, 6, 7, 5}}, {{0, 4, 5, 1}, {2, 6, 7, 3}}, {{0, 2, 6, 4}, {1, 3, 7, 5}}};
__constant__ uint8_t c_face_eids[3][2][4] = {{{0xD0, 0xDA, 0xD9, 0xC8}, {0xF4, 0xFE, 0xFD, 0xEC}},
                                             {{0xE0, 0xEC, 0xE9, 0xC8}, {0xF2, 0xFE, 0xFB, 0xDA}},
                                             {{0xD0, 0xF2, 0xF4, 0xE0}, {0xD9, 0xFB, 0xFD, 0xE9}}};
// Cuboid::support_face (cuboid.rs:267-354); the id tables are the reference's literals, [axis][sign_index]
__device__ __forceinline__ void cuboid_support_face(V3 he, V3 dir, PolyFace& f) {
    int iamax = 0;
    float best = fabsf(dir.x);
    if (fabsf(dir.y) > best) { best = fabsf(dir.y); iamax = 1; }
    if (fabsf(dir.z) > best) iamax = 2;
    float sign = copysignf(1.0f, comp(dir, iamax));
    if (iamax == 0) { f.v[0] = mk3(he.x * sign, he.y, he.z); f.v[1] = mk3(he.x * sign, -he.y, he.z); f.v[2] = mk3(he.x * sign, -he.y, -he.z); f.v[3] = mk3(he.x * sign, he.y, -he.z); }
    else if (iamax == 1) { f.v[0] = mk3(he.x, he.y * sign, he.z); f.v[1] = mk3(-he.x, he.y * sign, he.z); f.v[2] = mk3(-he.x, he.y * sign, -he.z); f.v[3] = mk3(he.x, he.y * sign, -he.z); }
    else { f.v[0] = mk3(he.x, he.y, he.z * sign); f.v[1] = mk3(he.x, -he.y, he.z * sign); f.v[2] = mk3(-he.x, -he.y, he.z * sign); f.v[3] = mk3(-he.x, he.y, he.z * sign); }
    int si = sign > 0.0f ? 1 : 0;
#pragma unroll
    for (int k = 0; k < 4; ++k) { f.vids[k] = PK_VERTEX((uint32_t)c_face_vids[iamax][si][k] * 2u); f.eids[k] = PK_EDGE((uint32_t)c_face_eids[iamax][si][k]); }
    f.fid = PK_FACE((uint32_t)(iamax + si * 3 + 10));
    f.n = 4;
}

__device__ __forceinline__ float perp2(float ax, float ay, float bx, float by) { return ax * by - ay * bx; }
// polygonal_feature3d.rs:398-439
__device__ __forceinline__ bool closest_points_line2d(float2 a0, float2 a1, float2 c0, float2 c1, float& s_out, float& t_out) {
    float d1x = a1.x - a0.x, d1y = a1.y - a0.y, d2x = c1.x - c0.x, d2y = c1.y - c0.y, rx = a0.x - c0.x, ry = a0.y - c0.y;
    float a = d1x * d1x + d1y * d1y, e = d2x * d2x + d2y * d2y, f = d2x * rx + d2y * ry;
    const float eps = PB2_EPS;
    if (a <= eps && e <= eps) { s_out = 0.f; t_out = 0.f; return true; }
    if (a <= eps) { s_out = 0.f; t_out = f / e; return true; }
    float c = d1x * rx + d1y * ry;
    if (e <= eps) { s_out = -c / a; t_out = 0.f; return true; }
    float b = d1x * d2x + d1y * d2y;
    float ae = a * e, bb = b * b, denom = ae - bb;
    if (denom <= eps || ulps_eq4(ae, bb)) return false;
    float sv = (b * f - c * e) / denom;
    s_out = sv; t_out = (b * sv + f) / e;
    return true;
}

// PolygonalFeature::contacts_face_face (polygonal_feature3d.rs:215-396) for faces of 3 or 4 vertices (unused slots are zero, as
// PolygonalFeature::default leaves them)
__device__ __forceinline__ void contacts_face_face(const Iso7& pos12, const PolyFace& f1, V3 sep, const PolyFace& f2, ManifoldOut& m) {
    float sign = copysignf(1.0f, sep.z);
    float a = -1.0f / (sign + sep.z);
    float b = sep.x * sep.y * a;
    V3 b0 = mk3(1.0f + sign * sep.x * sep.x * a, sign * b, -sign * sep.x);
    V3 b1 = mk3(b, sign + sep.y * sep.y * a, -sep.y);
    float2 pf1[4], pf2[4];
    V3 v21[4];
    const int n1 = f1.n, n2 = f2.n, last1 = f1.n - 1, last2 = f2.n - 1;
#pragma unroll
    for (int i = 0; i < 4; ++i) { pf1[i] = make_float2(dot3(f1.v[i], b0), dot3(f1.v[i], b1)); }
#pragma unroll
    for (int i = 0; i < 4; ++i) { v21[i] = iso_point(pos12, f2.v[i]); pf2[i] = make_float2(dot3(v21[i], b0), dot3(v21[i], b1)); }
    if (n2 > 2) {
        V3 normal2_1 = cross3(v21[2] - v21[1], v21[0] - v21[1]);
        float denom = dot3(normal2_1, sep);
        if (!rel_eq(denom, 0.0f, PB2_EPS, PB2_EPS)) {
            for (int i = 0; i < n1; ++i) {
                float2 p = pf1[i];
                float sg = perp2(pf2[0].x - pf2[last2].x, pf2[0].y - pf2[last2].y, p.x - pf2[last2].x, p.y - pf2[last2].y);
                bool outside = false;
                for (int j = 0; j < last2; ++j) {
                    float ns = perp2(pf2[j + 1].x - pf2[j].x, pf2[j + 1].y - pf2[j].y, p.x - pf2[j].x, p.y - pf2[j].y);
                    if (sg == 0.0f) sg = ns;
                    else if (sg * ns < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                float dist = dot3(v21[0] - f1.v[i], normal2_1) / denom;
                V3 lp1 = f1.v[i];
                V3 lp2_1 = f1.v[i] + sep * dist;
                m.push(lp1, iso_inv_point(pos12, lp2_1), f1.vids[i], f2.fid, dist, false);
            }
        }
    }
    if (n1 > 2) {
        V3 normal1 = cross3(f1.v[2] - f1.v[1], f1.v[0] - f1.v[1]);
        float denom = -dot3(normal1, sep);
        if (!rel_eq(denom, 0.0f, PB2_EPS, PB2_EPS)) {
            for (int i = 0; i < n2; ++i) {
                float2 p = pf2[i];
                float sg = perp2(pf1[0].x - pf1[last1].x, pf1[0].y - pf1[last1].y, p.x - pf1[last1].x, p.y - pf1[last1].y);
                bool outside = false;
                for (int j = 0; j < last1; ++j) {
                    float ns = perp2(pf1[j + 1].x - pf1[j].x, pf1[j + 1].y - pf1[j].y, p.x - pf1[j].x, p.y - pf1[j].y);
                    if (sg == 0.0f) sg = ns;
                    else if (sg * ns < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                float dist = dot3(f1.v[0] - v21[i], normal1) / denom;
                V3 lp2_1 = v21[i];
                V3 lp1 = v21[i] - sep * dist;
                m.push(lp1, iso_inv_point(pos12, lp2_1), f1.fid, f2.vids[i], dist, false);
            }
        }
    }
    for (int j = 0; j < n2; ++j) {
        const int jn = j + 1 == n2 ? 0 : j + 1;
        for (int i = 0; i < n1; ++i) {
            const int in = i + 1 == n1 ? 0 : i + 1;
            float sv, tv;
            if (!closest_points_line2d(pf1[i], pf1[in], pf2[j], pf2[jn], sv, tv)) continue;
            if (sv > 0.0f && sv < 1.0f && tv > 0.0f && tv < 1.0f) {
                V3 lp1 = f1.v[i] * (1.0f - sv) + f1.v[in] * sv;
                V3 lp2_1 = v21[j] * (1.0f - tv) + v21[jn] * tv;
                float dist = dot3(lp2_1 - lp1, sep);
                m.push(lp1, iso_inv_point(pos12, lp2_1), f1.eids[i], f2.eids[j], dist, false);
            }
        }
    }
}

// PolygonalFeatureMap::local_support_feature: Cuboid = support_face; ConvexPolyhedron = the face whose normal has the first
// maximal dot with dir, its first <= 4 vertices (convex_polyhedron.rs:959-991)
struct HullTopo {
    const uint32_t *hull_face_first, *hull_face_count, *face_first, *face_count, *va, *ea;
    const float* face_normal;
    const uint32_t *vert_first, *vert_count, *fav, *eav, *hull_edge_first;   // vertex side (may be NULL)
    const float* edge_dir;
};
#define PB2_SIN_1DEG 0.017452406f   // (PI / 180 as f32).sin_cos()
#define PB2_COS_1DEG 0.99984770f
// ConvexPolyhedron::support_feature_id_toward (convex_polyhedron.rs:885-922) -> PackedFeatureId
__device__ __forceinline__ uint32_t hull_support_feature_id(float4 pr, uint32_t sid, const float4* __restrict__ pts, const HullTopo& t, V3 dir) {
    const uint32_t p0 = __float_as_uint(pr.x), np_ = __float_as_uint(pr.y);
    const float4* hp = pts + p0;
    uint32_t best = 0;
    float best_dot = dot3(mk3(hp[0].x, hp[0].y, hp[0].z), dir);
    for (uint32_t i = 1; i < np_; ++i) { float4 q = hp[i]; float d = dot3(mk3(q.x, q.y, q.z), dir); if (d > best_dot) { best_dot = d; best = i; } }
    const uint32_t first = t.vert_first[p0 + best], cnt = t.vert_count[p0 + best];
    const float* fn = t.face_normal + 3ull * t.hull_face_first[sid];
    for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t f = t.fav[first + i];
        if (dot3(mk3(fn[3 * f], fn[3 * f + 1], fn[3 * f + 2]), dir) >= PB2_COS_1DEG) return PK_FACE(f);
    }
    const float* ed = t.edge_dir + 3ull * t.hull_edge_first[sid];
    for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t e = t.eav[first + i];
        if (fabsf(dot3(mk3(ed[3 * e], ed[3 * e + 1], ed[3 * e + 2]), dir)) <= PB2_SIN_1DEG) return PK_EDGE(e);
    }
    return PK_VERTEX(best);
}
__device__ __forceinline__ void pfm_support_feature(uint8_t kind, float4 pr, uint32_t sid, const float4* __restrict__ pts, const HullTopo& t, V3 dir,
                                                    PolyFace& f) {
    if (kind == PB2_SHAPE_CUBOID) { cuboid_support_face(mk3(pr.x, pr.y, pr.z), dir, f); return; }
    const uint32_t f0 = t.hull_face_first[sid], nf = t.hull_face_count[sid];
    const float* fn = t.face_normal + 3ull * f0;
    uint32_t best = 0;
    float best_dot = dot3(mk3(fn[0], fn[1], fn[2]), dir);
    for (uint32_t k = 1; k < nf; ++k) {
        float d = dot3(mk3(fn[3 * k], fn[3 * k + 1], fn[3 * k + 2]), dir);
        if (d > best_dot) { best = k; best_dot = d; }
    }
    const uint32_t i1 = t.face_first[f0 + best], cnt = t.face_count[f0 + best];
    const uint32_t nv = cnt < 4u ? cnt : 4u;
    const float4* hp = pts + __float_as_uint(pr.x);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if ((uint32_t)i < nv) {
            uint32_t vid = t.va[i1 + i];
            float4 q = hp[vid];
            f.v[i] = mk3(q.x, q.y, q.z);
            f.vids[i] = PK_VERTEX(vid);
            f.eids[i] = PK_EDGE(t.ea[i1 + i]);
        } else { f.v[i] = mk3(0.f, 0.f, 0.f); f.vids[i] = 0u; f.eids[i] = 0u; }
    }
    f.fid = PK_FACE(best);
    f.n = (int)nv;
}

__global__ void __launch_bounds__(128) k_contact_manifolds(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, uint32_t n_shapes,
                              const uint32_t* __restrict__ shape1, const uint32_t* __restrict__ shape2, const float* __restrict__ pos1,
                              const float* __restrict__ pos2, float prediction, uint32_t n, uint32_t max_points, float* __restrict__ normals,
                              uint32_t* __restrict__ counts, float* __restrict__ pts, uint8_t* __restrict__ status, bool have_topology,
                              bool have_vertex_topology, uint32_t* __restrict__ parked, unsigned long long* parked_count,
                              const uint8_t* __restrict__ skip) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (skip && skip[k]) return;   // persistent dispatch: last frame's manifold was kept (k_manifold_try_update)
    {   // pairs for the pfm_pfm arm (a ConvexPolyhedron with topology on at least one side, a Cuboid or such a hull on the other)
        uint32_t a = shape1[k], b = shape2[k];
        if (have_topology && a < n_shapes && b < n_shapes) {
            uint8_t ka = kinds[a], kb = kinds[b];
            bool hull = ka == PB2_SHAPE_CONVEX || kb == PB2_SHAPE_CONVEX, ball = ka == PB2_SHAPE_BALL || kb == PB2_SHAPE_BALL;
            // hull vs hull / cuboid: pfm_pfm; hull vs ball: convex_ball (needs the vertex side for the feature id)
            if (hull && (!ball || have_vertex_topology)) { parked[warp_append1(parked_count)] = k; return; }   // outputs written by k_manifold_pfm
        }
    }
    ManifoldOut m;
    m.pts = pts + (size_t)k * max_points * 9;
    m.max_points = max_points; m.count = 0; m.overflow = false;
    V3 n1 = mk3(0.f, 0.f, 0.f), n2 = n1;
    int st = MAN_OK;
    uint32_t s1 = shape1[k], s2 = shape2[k];
    if (s1 >= n_shapes || s2 >= n_shapes || kinds[s1] > PB2_SHAPE_CUBOID || kinds[s2] > PB2_SHAPE_CUBOID) st = MAN_UNSUPPORTED;
    else {
        Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));
        bool b1 = kinds[s1] == PB2_SHAPE_BALL, b2 = kinds[s2] == PB2_SHAPE_BALL;
        float4 pr1 = params[s1], pr2 = params[s2];
        if (b1 && b2) {
            // contact_manifolds_ball_ball.rs:17-57
            V3 dcenter = pos12.t;
            float center_dist = nrm(dcenter);
            float dist = center_dist - pr1.x - pr2.x;
            if (dist < prediction) {
                n1 = center_dist != 0.0f ? dcenter / center_dist : mk3(0.f, 1.f, 0.f);
                n2 = iso_inv_vec(pos12, -n1);
                m.push(n1 * pr1.x, n2 * pr2.x, PK_FACE(0u), PK_FACE(0u), dist, false);
            }
        } else if (b1 != b2) {
            // contact_manifolds_convex_ball.rs:18-145: the ball is shape 2 of the inner call (pos12 inverted when it is shape 1)
            bool flipped = b1;
            Iso7 p12 = flipped ? iso_inverse(pos12) : pos12;
            V3 he = flipped ? mk3(pr2.x, pr2.y, pr2.z) : mk3(pr1.x, pr1.y, pr1.z);
            float radius = flipped ? pr1.x : pr2.x;
            V3 proj; bool inside; Feat f;
            d_cuboid_project(he, p12.t, proj, inside, f);
            V3 dpos = p12.t - proj;
            V3 ln1; float dist;
            if (!try_normalize_get(dpos, 0.0f, ln1, dist)) {
                float nn;
                if (!try_normalize_get(p12.t, 0.0f, ln1, nn)) ln1 = mk3(1.f, 0.f, 0.f);
                dist = 0.0f;
            }
            if (inside) { ln1 = -ln1; dist = -dist; }
            if (dist <= radius + prediction) {
                V3 ln2 = iso_inv_vec(p12, -ln1);
                uint32_t fid1 = f.kind == 0 ? PK_VERTEX(f.id) : (f.kind == 1 ? PK_EDGE(f.id) : (f.kind == 2 ? PK_FACE(f.id) : 0u));
                m.push(proj, ln2 * radius, fid1, PK_FACE(0u), dist - radius, flipped);
                if (flipped) { n1 = ln2; n2 = ln1; } else { n1 = ln1; n2 = ln2; }
            }
        } else {
            // contact_manifolds_cuboid_cuboid.rs:19-107
            V3 he1 = mk3(pr1.x, pr1.y, pr1.z), he2 = mk3(pr2.x, pr2.y, pr2.z);
            Iso7 pos21 = iso_inverse(pos12);
            float sp1, sp2, sp3; V3 d1, d2, d3;
            sat_normal_oneway(he1, he2, pos12, sp1, d1);
            bool sepd = sp1 > prediction;
            if (!sepd) { sat_normal_oneway(he2, he1, pos21, sp2, d2); sepd = sp2 > prediction; }
            if (!sepd) { sat_edge_twoway(he1, he2, pos12, sp3, d3); sepd = sp3 > prediction; }
            if (!sepd) {
                V3 best = d1;
                if (sp2 > sp1 && sp2 > sp3) best = iso_vec(pos12, -d2);
                else if (sp3 > sp1) best = d3;
                V3 ln2 = iso_vec(pos21, -best);
                PolyFace f1, f2;
                cuboid_support_face(he1, best, f1);
                cuboid_support_face(he2, ln2, f2);
                contacts_face_face(pos12, f1, best, f2, m);
                n1 = best; n2 = ln2;
            }
        }
    }
    if (m.overflow) st = MAN_OVERFLOW;
    if (m.count == 0) { n1 = mk3(0.f, 0.f, 0.f); n2 = n1; }
    float* nq = normals + 6ull * k;
    nq[0] = n1.x; nq[1] = n1.y; nq[2] = n1.z; nq[3] = n2.x; nq[4] = n2.y; nq[5] = n2.z;
    for (uint32_t i = m.count; i < max_points; ++i) { float* o = m.pts + 9ull * i; for (int j = 0; j < 9; ++j) o[j] = 0.0f; }
    counts[k] = m.count;
    status[k] = (uint8_t)st;
}

// contact_manifolds_pfm_pfm.rs:42-162 after the GJK/EPA contact (first frame, no normal constraints, no border radius): support
// features along the contact normals, face clipping, and the witness pair itself as one more (feature-less) point.
__global__ void __launch_bounds__(128) k_manifold_pfm(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ hull_pts,
                              HullTopo topo, const uint32_t* __restrict__ shape1, const uint32_t* __restrict__ shape2, const float* __restrict__ pos1,
                              const float* __restrict__ pos2, const uint32_t* __restrict__ parked, uint32_t count, const float* __restrict__ contacts,
                              const uint8_t* __restrict__ cstatus, uint32_t max_points, float* __restrict__ normals, uint32_t* __restrict__ counts,
                              float* __restrict__ pts, uint8_t* __restrict__ status, const float* __restrict__ noint_dir) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t k = parked[i];
    ManifoldOut m;
    m.pts = pts + (size_t)k * max_points * 9;
    m.max_points = max_points; m.count = 0; m.overflow = false;
    V3 n1 = mk3(0.f, 0.f, 0.f), n2 = n1;
    int st = MAN_OK;
    if (cstatus[i] == ST_NEEDS_HOST) st = 3;
    else if (cstatus[i] == ST_SOME) {
        const float* c = contacts + 13ull * i;
        n1 = mk3(c[6], c[7], c[8]);
        n2 = mk3(c[9], c[10], c[11]);
        uint32_t a = shape1[k], b = shape2[k];
        if (kinds[a] == PB2_SHAPE_BALL || kinds[b] == PB2_SHAPE_BALL) {
            // contact_manifolds_convex_ball.rs:42-145 with a ConvexPolyhedron: the contact arm computed the same projection, normals
            // and dist (contact_ball_convex_polyhedron.rs); what the manifold adds is the hull feature under the contact,
            // support_feature_id_toward(the outward normal in the hull's frame) (point_support_map.rs:62-76)
            const bool ball_first = kinds[a] == PB2_SHAPE_BALL;
            uint32_t hf = ball_first ? hull_support_feature_id(params[b], b, hull_pts, topo, n2) : hull_support_feature_id(params[a], a, hull_pts, topo, n1);
            m.push(mk3(c[0], c[1], c[2]), mk3(c[3], c[4], c[5]), ball_first ? PK_FACE(0u) : hf, ball_first ? hf : PK_FACE(0u), c[12], false);
        } else {
            Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));
            PolyFace f1, f2;
            pfm_support_feature(kinds[a], params[a], a, hull_pts, topo, n1, f1);
            pfm_support_feature(kinds[b], params[b], b, hull_pts, topo, n2, f2);
            contacts_face_face(pos12, f1, n1, f2, m);
            m.push(mk3(c[0], c[1], c[2]), mk3(c[3], c[4], c[5]), 0u, 0u, c[12], false);   // PackedFeatureId::UNKNOWN
        }
    }
    if (m.overflow) st = MAN_OVERFLOW;
    if (m.count == 0) {
        n1 = mk3(0.f, 0.f, 0.f); n2 = n1;
        // GJKResult::NoIntersection(dir) => manifold.local_n1 = dir: "use the manifold normal as a cache" for next frame's GJK
        // (contact_manifolds_pfm_pfm.rs:151-154; EPA's "everything failed" answers NoIntersection(+x), the array's initial value)
        if (cstatus[i] == ST_NONE && kinds[shape1[k]] != PB2_SHAPE_BALL && kinds[shape2[k]] != PB2_SHAPE_BALL)
            n1 = mk3(noint_dir[3ull * i], noint_dir[3ull * i + 1], noint_dir[3ull * i + 2]);
    }
    float* nq = normals + 6ull * k;
    nq[0] = n1.x; nq[1] = n1.y; nq[2] = n1.z; nq[3] = n2.x; nq[4] = n2.y; nq[5] = n2.z;
    for (uint32_t j = m.count; j < max_points; ++j) { float* o = m.pts + 9ull * j; for (int q = 0; q < 9; ++q) o[q] = 0.0f; }
    counts[k] = m.count;
    status[k] = (uint8_t)st;
}

// The device side of pb2_contact_manifolds_batch on device-resident arrays; pairs with skip[k] != 0 (optional) are left untouched.
__global__ void k_fill_xaxis(float* __restrict__ d, uint32_t n) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { d[3ull * i] = 1.0f; d[3ull * i + 1] = 0.0f; d[3ull * i + 2] = 0.0f; }
}

// d_seed (optional, n x 6): last frame's manifold normals; local_n1 seeds the GJK of the pfm_pfm arm (contact_manifolds_pfm_pfm.rs:66).
// It may alias d_nr: the contact kernels read it before k_manifold_pfm writes the new normals.
static int manifolds_device(pb2_ctx* ctx, const pb2_shapes* shapes, const void* d_s1, const void* d_s2, const void* d_p1, const void* d_p2,
                            float prediction, uint32_t n, uint32_t max_points, void* d_nr, void* d_ct, void* d_pt, void* d_st, const uint8_t* d_skip,
                            const float* d_seed = nullptr) {
    cudaStream_t st = ctx->stream;
    const bool have_topology = shapes->face_normal != nullptr;
    uint32_t *d_parked = nullptr, *d_ab = nullptr;
    float *d_c = nullptr, *d_dir = nullptr;
    uint8_t* d_cst = nullptr;
    unsigned long long* parked_count = (unsigned long long*)(ctx->d_counters + 10);
    int rc = PB2_OK;
    if (have_topology) {
        if (cudaMallocAsync((void**)&d_parked, (size_t)n * 4, st) != cudaSuccess) PB2_FAIL(ctx, PB2_ERR_CUDA, "contact_manifolds: out of device memory");
        cudaMemsetAsync(parked_count, 0, 8, st);
    }
    k_contact_manifolds<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->n, (const uint32_t*)d_s1, (const uint32_t*)d_s2,
                                                             (const float*)d_p1, (const float*)d_p2, prediction, n, max_points, (float*)d_nr,
                                                             (uint32_t*)d_ct, (float*)d_pt, (uint8_t*)d_st, have_topology, shapes->vert_first != nullptr,
                                                             d_parked, parked_count, d_skip);
    PB2_LAUNCHED(ctx);
    if (have_topology) {
        cudaMemcpyAsync(ctx->h_counters + 10, parked_count, 8, cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) rc = PB2_ERR_CUDA;
        uint32_t cnt = rc == PB2_OK ? (uint32_t)ctx->h_counters[10] : 0u;
        if (cnt) {
            if (cudaMallocAsync((void**)&d_ab, (size_t)cnt * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_c, (size_t)cnt * 52, st) != cudaSuccess ||
                cudaMallocAsync((void**)&d_cst, cnt, st) != cudaSuccess || cudaMallocAsync((void**)&d_dir, (size_t)cnt * 12, st) != cudaSuccess) rc = PB2_ERR_CUDA;
            if (rc == PB2_OK) {
                k_pair_up<<<pb2_blocks(cnt, 256), 256, 0, st>>>(d_parked, cnt, d_ab);
                k_fill_xaxis<<<pb2_blocks(cnt, 256), 256, 0, st>>>(d_dir, cnt);
                ctx->launches += 2;
                OutSinks sinks;
                sinks.noint_dir = d_dir;
                sinks.dense = d_c; sinks.status = d_cst; sinks.compact = nullptr; sinks.pair_index = nullptr; sinks.cap = 0;
                sinks.compact_count = nullptr; sinks.some_count = nullptr;
                // contact_support_map_support_map_with_params(pos12, pfm1, pfm2, prediction, .., None): the contact kernels, local frames
                rc = run_contacts(ctx, shapes, (const uint32_t*)d_s1, (const uint32_t*)d_s2, (const float*)d_p1, (const float*)d_p2, prediction, cnt, sinks,
                                  d_ab, n, nullptr, 0, PAIR_LOCAL_FRAMES, nullptr, 0, d_seed, 6);
            }
            if (rc == PB2_OK) {
                HullTopo topo;
                topo.hull_face_first = shapes->hull_face_first; topo.hull_face_count = shapes->hull_face_count; topo.face_first = shapes->face_first;
                topo.face_count = shapes->face_count; topo.va = shapes->verts_adj_to_face; topo.ea = shapes->edges_adj_to_face;
                topo.face_normal = shapes->face_normal;
                topo.vert_first = shapes->vert_first; topo.vert_count = shapes->vert_count; topo.fav = shapes->faces_adj_to_vertex;
                topo.eav = shapes->edges_adj_to_vertex; topo.hull_edge_first = shapes->hull_edge_first; topo.edge_dir = shapes->edge_dir;
                k_manifold_pfm<<<pb2_blocks(cnt, 128), 128, 0, st>>>(shapes->kinds, shapes->params, shapes->points4, topo, (const uint32_t*)d_s1,
                                                                    (const uint32_t*)d_s2, (const float*)d_p1, (const float*)d_p2, d_parked, cnt, d_c, d_cst,
                                                                    max_points, (float*)d_nr, (uint32_t*)d_ct, (float*)d_pt, (uint8_t*)d_st, d_dir);
                PB2_LAUNCHED(ctx);
            }
        }
        if (d_dir) cudaFreeAsync(d_dir, st);
        if (d_parked) cudaFreeAsync(d_parked, st);
        if (d_ab) cudaFreeAsync(d_ab, st);
        if (d_c) cudaFreeAsync(d_c, st);
        if (d_cst) cudaFreeAsync(d_cst, st);
        if (rc != PB2_OK) { snprintf(ctx->err, sizeof(ctx->err), "contact_manifolds: pfm_pfm phase failed"); return rc; }
    }
    PB2_CUDA(ctx, cudaGetLastError());
    return PB2_OK;
}

extern "C" int pb2_contact_manifolds_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                                           const float* pos2, float prediction, uint32_t n, uint32_t max_points, float* normals, uint32_t* counts,
                                           float* points, uint8_t* status, int mem) {
    if (!ctx || !shapes || max_points == 0 || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !normals || !counts || !points || !status))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s1, *d_s2, *d_p1, *d_p2;
    void *d_nr, *d_ct, *d_pt, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_out(ctx, 4, normals, (size_t)n * 24, mem, &d_nr));
    PB2_CHECK(pb2_stage_out(ctx, 5, counts, (size_t)n * 4, mem, &d_ct));
    PB2_CHECK(pb2_stage_out(ctx, 6, points, (size_t)n * max_points * 36, mem, &d_pt));
    PB2_CHECK(pb2_stage_out(ctx, 7, status, (size_t)n, mem, &d_st));
    PB2_CHECK(manifolds_device(ctx, shapes, d_s1, d_s2, d_p1, d_p2, prediction, n, max_points, d_nr, d_ct, d_pt, d_st, nullptr));
    PB2_CHECK(pb2_stage_back(ctx, normals, d_nr, (size_t)n * 24, mem));
    PB2_CHECK(pb2_stage_back(ctx, counts, d_ct, (size_t)n * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, points, d_pt, (size_t)n * max_points * 36, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------- manifold persistence
// ContactManifold::try_update_contacts_eps (contact_manifold.rs:662-699) per manifold, in place; the per-pair bodies are
// __host__ __device__ functions in manifold_update.cuh (tests/hostcheck runs them on the CPU against the oracle).
__global__ void __launch_bounds__(128) k_manifold_try_update(const uint8_t* __restrict__ kinds, uint32_t n_shapes, const uint32_t* __restrict__ shape1,
                                                             const uint32_t* __restrict__ shape2, bool dispatch, bool have_topology,
                                                             const float* __restrict__ pos1, const float* __restrict__ pos2, uint32_t n,
                                                             uint32_t max_points, float angle_dot_threshold, float dist_sq_threshold,
                                                             const float* __restrict__ normals, const uint32_t* __restrict__ counts,
                                                             float* __restrict__ pts, uint8_t* __restrict__ kept, uint8_t* __restrict__ status,
                                                             uint32_t* __restrict__ old_fids, uint32_t* __restrict__ old_counts) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    manifold_try_update_pair(k, kinds, n_shapes, shape1, shape2, dispatch, have_topology, pos1, pos2, max_points, angle_dot_threshold,
                             dist_sq_threshold, normals, counts, pts, kept, status, old_fids, old_counts);
}

__global__ void __launch_bounds__(128) k_manifold_match(const uint8_t* __restrict__ kept, const uint32_t* __restrict__ old_fids,
                                                        const uint32_t* __restrict__ old_counts, const uint32_t* __restrict__ counts,
                                                        const float* __restrict__ pts, uint32_t n, uint32_t max_points, int32_t* __restrict__ match) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    manifold_match_pair(k, kept, old_fids, old_counts, counts, pts, max_points, match);
}

extern "C" int pb2_manifolds_try_update(pb2_ctx* ctx, const float* pos1, const float* pos2, uint32_t n, uint32_t max_points,
                                        float angle_dot_threshold, float dist_sq_threshold, const float* normals, const uint32_t* counts,
                                        float* points, uint8_t* kept, int mem) {
    if (!ctx || max_points == 0 || (n && (!pos1 || !pos2 || !normals || !counts || !points || !kept))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_p1, *d_p2, *d_nr, *d_ct, *d_pt;
    void* d_kp;
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_in(ctx, 4, normals, (size_t)n * 24, mem, &d_nr));
    PB2_CHECK(pb2_stage_in(ctx, 5, counts, (size_t)n * 4, mem, &d_ct));
    PB2_CHECK(pb2_stage_in(ctx, 6, points, (size_t)n * max_points * 36, mem, &d_pt));   // in / out: uploaded here, read back below
    PB2_CHECK(pb2_stage_out(ctx, 7, kept, (size_t)n, mem, &d_kp));
    k_manifold_try_update<<<pb2_blocks(n, 128), 128, 0, ctx->stream>>>(nullptr, 0u, nullptr, nullptr, false, false, (const float*)d_p1, (const float*)d_p2, n,
                                                                       max_points, angle_dot_threshold, dist_sq_threshold, (const float*)d_nr,
                                                                       (const uint32_t*)d_ct, (float*)d_pt, (uint8_t*)d_kp, nullptr, nullptr, nullptr);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, points, d_pt, (size_t)n * max_points * 36, mem));
    PB2_CHECK(pb2_stage_back(ctx, kept, d_kp, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

extern "C" int pb2_contact_manifolds_update_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2,
                                                  const float* pos1, const float* pos2, float prediction, uint32_t n, uint32_t max_points,
                                                  float* normals, uint32_t* counts, float* points, uint8_t* status, uint8_t* kept, int32_t* match,
                                                  int mem) {
    if (!ctx || !shapes || max_points == 0 || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !normals || !counts || !points || !status || !kept)))
        return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    const void *d_s1, *d_s2, *d_p1, *d_p2, *c_nr, *c_ct, *c_pt;
    void* d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    // in / out arrays: last frame's manifolds go up (host mode) into the staging slots the results come back from
    PB2_CHECK(pb2_stage_in(ctx, 4, normals, (size_t)n * 24, mem, &c_nr));
    PB2_CHECK(pb2_stage_in(ctx, 5, counts, (size_t)n * 4, mem, &c_ct));
    PB2_CHECK(pb2_stage_in(ctx, 6, points, (size_t)n * max_points * 36, mem, &c_pt));
    PB2_CHECK(pb2_stage_out(ctx, 7, status, (size_t)n, mem, &d_st));
    void *d_nr = const_cast<void*>(c_nr), *d_ct = const_cast<void*>(c_ct), *d_pt = const_cast<void*>(c_pt);
    // kept / match / saved feature ids: stream-ordered temporaries in host mode, the caller's arrays in device mode
    uint8_t* d_kp = kept;
    int32_t* d_mt = match;
    uint32_t *d_of = nullptr, *d_oc = nullptr;
    int rc = PB2_OK;
    if (mem == PB2_MEM_HOST) {
        d_kp = nullptr; d_mt = nullptr;
        if (cudaMallocAsync((void**)&d_kp, (size_t)n, st) != cudaSuccess) rc = PB2_ERR_CUDA;
        if (rc == PB2_OK && match && cudaMallocAsync((void**)&d_mt, (size_t)n * max_points * 4, st) != cudaSuccess) rc = PB2_ERR_CUDA;
    }
    if (rc == PB2_OK && match) {
        if (cudaMallocAsync((void**)&d_of, (size_t)n * max_points * 8, st) != cudaSuccess || cudaMallocAsync((void**)&d_oc, (size_t)n * 4, st) != cudaSuccess)
            rc = PB2_ERR_CUDA;
    }
    if (rc == PB2_OK) {
        k_manifold_try_update<<<pb2_blocks(n, 128), 128, 0, st>>>(shapes->kinds, shapes->n, (const uint32_t*)d_s1, (const uint32_t*)d_s2, true,
                                                                  shapes->face_normal != nullptr, (const float*)d_p1, (const float*)d_p2, n, max_points,
                                                                  PB2_COS_1_DEGREES, PB2_UPDATE_DIST_SQ, (const float*)d_nr, (const uint32_t*)d_ct,
                                                                  (float*)d_pt, d_kp, (uint8_t*)d_st, d_of, d_oc);
        PB2_LAUNCHED(ctx);
        // last frame's normals double as the GJK seed of the pairs that are recomputed
        rc = manifolds_device(ctx, shapes, d_s1, d_s2, d_p1, d_p2, prediction, n, max_points, d_nr, d_ct, d_pt, d_st, d_kp, (const float*)d_nr);
    }
    if (rc == PB2_OK && match) {
        k_manifold_match<<<pb2_blocks(n, 128), 128, 0, st>>>(d_kp, d_of, d_oc, (const uint32_t*)d_ct, (const float*)d_pt, n, max_points, d_mt);
        PB2_LAUNCHED(ctx);
        if (cudaGetLastError() != cudaSuccess) rc = PB2_ERR_CUDA;
    }
    if (rc == PB2_OK && mem == PB2_MEM_HOST) {
        if (cudaMemcpyAsync(kept, d_kp, (size_t)n, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = PB2_ERR_CUDA;
        if (rc == PB2_OK && match && cudaMemcpyAsync(match, d_mt, (size_t)n * max_points * 4, cudaMemcpyDeviceToHost, st) != cudaSuccess) rc = PB2_ERR_CUDA;
    }
    if (mem == PB2_MEM_HOST) { if (d_kp) cudaFreeAsync(d_kp, st); if (d_mt) cudaFreeAsync(d_mt, st); }
    if (d_of) cudaFreeAsync(d_of, st);
    if (d_oc) cudaFreeAsync(d_oc, st);
    if (rc != PB2_OK) { if (!ctx->err[0]) snprintf(ctx->err, sizeof(ctx->err), "contact_manifolds_update: device phase failed"); return rc; }
    PB2_CHECK(pb2_stage_back(ctx, normals, d_nr, (size_t)n * 24, mem));
    PB2_CHECK(pb2_stage_back(ctx, counts, d_ct, (size_t)n * 4, mem));
    PB2_CHECK(pb2_stage_back(ctx, points, d_pt, (size_t)n * max_points * 36, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

// ------------------------------------------------------------------------------------------- query::closest_points
// query::closest_points (closest_points_shape_shape.rs:220-231) -> DefaultQueryDispatcher::closest_points
// (default_query_dispatcher.rs:358-424): closest_points_ball_ball.rs:7-36; ball <-> convex through the contact arms
// (closest_points_ball_convex_polyhedron.rs:7-44: dist <= 0 => Intersecting); everything else
// closest_points_support_map_support_map.rs:8-69 (GJK started toward -pos12.translation, no EPA: Intersection => Intersecting).
// kind: 0 Disjoint, 1 WithinMargin (out = p1, p2 in world space, ClosestPoints::transform_by), 2 Intersecting.
// status: 1 ok, 2 unknown shape id, 3 host fallback (ball centre on a hull's surface: needs the hull's feature normal).
__global__ void __launch_bounds__(128) k_closest_points(const uint8_t* __restrict__ kinds, const float4* __restrict__ params, const float4* __restrict__ pts,
                              uint32_t n_shapes, PairSrc src, float max_dist, uint32_t n, float* __restrict__ out, uint8_t* __restrict__ kind_out,
                              uint8_t* __restrict__ status) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    float* o = out + 6ull * k;
    V3 p1 = mk3(0.f, 0.f, 0.f), p2 = p1;
    int kind = 0, st = ST_SOME;
    if (src.shape1[k] >= n_shapes || src.shape2[k] >= n_shapes) st = ST_UNSUPPORTED;
    else {
        PairSetup ps;
        pair_setup(kinds, params, pts, src, k, ps);
        bool b1 = ps.k1 == PB2_SHAPE_BALL, b2 = ps.k2 == PB2_SHAPE_BALL;
        ContactOut c;
        int cst = -1;   // >= 0: a contact-arm outcome to map
        if (b1 && b2) {
            float r1 = ps.pr1.x, r2 = ps.pr2.x;
            V3 delta = ps.pos12.t;
            float distance = nrm(delta), sum = r1 + r2;
            if (distance - max_dist <= sum) {
                if (distance <= sum) kind = 2;
                else {
                    V3 nrm1 = delta / distance;
                    p1 = nrm1 * r1;
                    p2 = iso_inv_vec(ps.pos12, nrm1) * (-r2);
                    kind = 1;
                }
            }
        } else if (ps.mode == 0) {
            cst = closed_form_pair(ps, max_dist, c);
        } else if (ps.mode == 1) {
            Simplex s;
            V3 dir; float nn;
            if (!try_normalize_get(-ps.pos12.t, PB2_EPS, dir, nn)) dir = mk3(1.f, 0.f, 0.f);
            sx_reset(s, cso_from_shapes(ps.gpos12, ps.g1, ps.g2, dir));
            V3 q1, q2, n1;
            int r = gjk_closest_points<true>(ps.gpos12, ps.g1, ps.g2, max_dist, s, q1, q2, n1);
            if (r == GJK_CLOSEST_POINTS) { kind = 1; p1 = q1; p2 = iso_inv_point(ps.pos12, q2); }
            else if (r == GJK_NO_INTERSECTION) kind = 0;
            else kind = 2;
        } else {
            Simplex s;
            gjk_start(ps, s, src, k);
            V3 q1, q2, n1;
            int r = gjk_closest_points<true>(ps.gpos12, ps.g1, ps.g2, FLT_MAX, s, q1, q2, n1);
            if (r == GJK_INTERSECTION) kind = 2;   // centre inside the hull: the contact has dist < 0
            else if (r == GJK_CLOSEST_POINTS) cst = finish_gjk_pair(ps, false, q1, q2, n1, max_dist, c);
            else kind = 0;
        }
        if (cst == ST_SOME) {
            if (c.dist <= 0.0f) kind = 2;
            else { kind = 1; p1 = c.p1; p2 = c.p2; }
        } else if (cst == ST_NEEDS_HOST) st = ST_NEEDS_HOST;
        if (kind == 1) { p1 = iso_point(ps.pos1, p1); p2 = iso_point(ps.pos2, p2); }
    }
    if (kind != 1) { p1 = mk3(0.f, 0.f, 0.f); p2 = p1; }
    o[0] = p1.x; o[1] = p1.y; o[2] = p1.z; o[3] = p2.x; o[4] = p2.y; o[5] = p2.z;
    kind_out[k] = (uint8_t)kind;
    status[k] = (uint8_t)st;
}

extern "C" int pb2_closest_points_batch(pb2_ctx* ctx, const pb2_shapes* shapes, const uint32_t* shape1, const uint32_t* shape2, const float* pos1,
                                        const float* pos2, float max_dist, uint32_t n, float* points, uint8_t* kind, uint8_t* status, int mem) {
    if (!ctx || !shapes || (n && (!shape1 || !shape2 || !pos1 || !pos2 || !points || !kind || !status))) return PB2_ERR_INVALID;
    if (n == 0) return PB2_OK;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    const void *d_s1, *d_s2, *d_p1, *d_p2;
    void *d_out, *d_kind, *d_st;
    PB2_CHECK(pb2_stage_in(ctx, 0, shape1, (size_t)n * 4, mem, &d_s1));
    PB2_CHECK(pb2_stage_in(ctx, 1, shape2, (size_t)n * 4, mem, &d_s2));
    PB2_CHECK(pb2_stage_in(ctx, 2, pos1, (size_t)n * 28, mem, &d_p1));
    PB2_CHECK(pb2_stage_in(ctx, 3, pos2, (size_t)n * 28, mem, &d_p2));
    PB2_CHECK(pb2_stage_out(ctx, 4, points, (size_t)n * 24, mem, &d_out));
    PB2_CHECK(pb2_stage_out(ctx, 5, kind, (size_t)n, mem, &d_kind));
    PB2_CHECK(pb2_stage_out(ctx, 6, status, (size_t)n, mem, &d_st));
    PairSrc src;
    src.shape1 = (const uint32_t*)d_s1; src.shape2 = (const uint32_t*)d_s2; src.pos1 = (const float*)d_p1; src.pos2 = (const float*)d_p2;
    src.ab = nullptr; src.mesh_tris = nullptr; src.n_first = src.n_second = 0;
    k_closest_points<<<pb2_blocks(n, 128), 128, 0, ctx->stream>>>(shapes->kinds, shapes->params, shapes->points4, shapes->n, src, max_dist, n, (float*)d_out,
                                                                   (uint8_t*)d_kind, (uint8_t*)d_st);
    PB2_LAUNCHED(ctx);
    PB2_CUDA(ctx, cudaGetLastError());
    PB2_CHECK(pb2_stage_back(ctx, points, d_out, (size_t)n * 24, mem));
    PB2_CHECK(pb2_stage_back(ctx, kind, d_kind, (size_t)n, mem));
    PB2_CHECK(pb2_stage_back(ctx, status, d_st, (size_t)n, mem));
    if (mem == PB2_MEM_HOST) PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return PB2_OK;
}

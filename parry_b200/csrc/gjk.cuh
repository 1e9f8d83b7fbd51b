// parry_b200 — device GJK: support maps, Minkowski-difference points, Voronoi simplex and the closest-points loop.
//
// Replaces (reference, file:line): gjk::closest_points (query/gjk/gjk.rs:353-453), eps_tol (:141-144), result (:797-818),
// VoronoiSimplex (query/gjk/voronoi_simplex3.rs:14-351), CSOPoint (query/gjk/cso_point.rs:13-89), ConstantOrigin
// (query/gjk/special_support_maps.rs), SupportMap::support_point (shape/support_map.rs:380-383), Cuboid / ConvexPolyhedron
// support maps (shape/cuboid.rs:452-457, shape/convex_polyhedron.rs:952-957 -> utils/point_cloud_support_point.rs:5-20),
// Segment / Triangle / Tetrahedron origin projection (query/point/point_segment.rs:49-84, point_triangle.rs:58-290,
// point_tetrahedron.rs:32-339). Same operation order as the reference, no FMA contraction (see common.cuh).
#pragma once
#include "shapes.cuh"

#define PB2_EPS 1.1920929e-7f          // f32::EPSILON (DEFAULT_EPSILON, src/lib.rs:102)
#define PB2_GJK_EPS_TOL (PB2_EPS * 10.0f)

enum { DS_CUBOID = 0, DS_CONVEX = 1, DS_ORIGIN = 2, DS_TRIANGLE = 3, DS_BALL = 4 /* radius in he.x (shape/ball.rs:230-250) */ };
struct DShape {
    int kind;
    V3 he;
    const float4* pts;
    uint32_t n;
};

__device__ __forceinline__ V3 ds_local_support(const DShape& s, V3 dir) {
    if (s.kind == DS_CUBOID) return mk3(copysignf(s.he.x, dir.x), copysignf(s.he.y, dir.y), copysignf(s.he.z, dir.z));
    if (s.kind == DS_CONVEX) {
        // first maximal vertex, strict '>' (point_cloud_support_point_id)
        float4 p = __ldg(&s.pts[0]);
        V3 best = mk3(p.x, p.y, p.z);
        float best_dot = dot3(best, dir);
        for (uint32_t i = 1; i < s.n; ++i) {
            float4 q = __ldg(&s.pts[i]);
            V3 v = mk3(q.x, q.y, q.z);
            float d = dot3(v, dir);
            if (d > best_dot) { best_dot = d; best = v; }
        }
        return best;
    }
    if (s.kind == DS_TRIANGLE) {
        // SupportMap for Triangle (shape/triangle.rs:697-716): its own comparison cascade, not the point-cloud scan
        float4 pa = __ldg(&s.pts[0]), pb = __ldg(&s.pts[1]), pc = __ldg(&s.pts[2]);
        V3 a = mk3(pa.x, pa.y, pa.z), b = mk3(pb.x, pb.y, pb.z), c = mk3(pc.x, pc.y, pc.z);
        float d1 = dot3(a, dir), d2 = dot3(b, dir), d3 = dot3(c, dir);
        if (d1 > d2) return d1 > d3 ? a : c;
        return d2 > d3 ? b : c;
    }
    if (s.kind == DS_BALL) return (dir / nrm(dir)) * s.he.x;  // local_support_point_toward(Unit::new_normalize(dir))
    return mk3(0.f, 0.f, 0.f);
}
__device__ __forceinline__ V3 ds_support_point(const DShape& s, const Iso7& m, V3 dir) {
    if (s.kind == DS_ORIGIN) return m.t;
    if (s.kind == DS_BALL) return m.t + (dir / nrm(dir)) * s.he.x;  // Ball overrides support_point: translation only
    V3 ld = iso_inv_vec(m, dir);
    return iso_point(m, ds_local_support(s, ld));
}

struct CSO {
    V3 point, o1, o2;
};
__device__ __forceinline__ CSO cso_make(V3 o1, V3 o2) { CSO c; c.point = o1 - o2; c.o1 = o1; c.o2 = o2; return c; }
__device__ __forceinline__ CSO cso_from_shapes(const Iso7& pos12, const DShape& g1, const DShape& g2, V3 dir) {
    V3 sp1 = ds_local_support(g1, dir);
    V3 sp2 = ds_support_point(g2, pos12, -dir);
    return cso_make(sp1, sp2);
}

// approx::relative_eq! with default epsilon / max_relative = f32::EPSILON
__device__ __forceinline__ bool rel_eq(float a, float b, float eps, float max_rel) {
    if (a == b) return true;
    if (isinf(a) || isinf(b)) return false;
    float d = fabsf(a - b);
    if (d <= eps) return true;
    float aa = fabsf(a), ab = fabsf(b);
    float largest = ab > aa ? ab : aa;
    return d <= largest * max_rel;
}
__device__ __forceinline__ bool rel_eq3(V3 a, V3 b) {
    return rel_eq(a.x, b.x, PB2_EPS, PB2_EPS) && rel_eq(a.y, b.y, PB2_EPS, PB2_EPS) && rel_eq(a.z, b.z, PB2_EPS, PB2_EPS);
}
__device__ __forceinline__ bool try_normalize_get(V3 v, float min_norm, V3& out, float& n) {
    float sq = nrm2(v);
    if (sq > min_norm * min_norm) { n = sqrtf(sq); out = v / n; return true; }
    return false;
}

// ---- projections of the origin / a point. kind: 0 vertex(idx) 1 edge(idx, bc[0..1]) 2 face(idx, bc[0..2]) 3 solid
struct Proj {
    V3 point;
    float bc[3];
    int kind;
    uint32_t idx;
    bool inside;
};

// point_triangle.rs:58-290 (pt is a general point; `solid` true on the GJK/EPA paths).
// Branch-free restatement: every intermediate of the reference is a pure function of the inputs, so all of them are
// evaluated eagerly, the Voronoi region is picked by the same cascade of comparisons, and the single division of the
// chosen region runs on selected operands — bit-identical results with no divergent paths inside a warp.
__device__ __forceinline__ void project_on_triangle(V3 a, V3 b, V3 c, V3 pt, Proj& r) {
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
    float sum = va + vb + vc;
    bool face = !vtx && !edge && sum != 0.0f;
    float num = rAB ? ab_ap : (rAC ? ac_ap : (rBC ? dot3(bc, bp) : 1.0f));
    float den = rAB ? nrm2(ab) : (rAC ? nrm2(ac) : (rBC ? nrm2(bc) : sum));
    float q = num / den;
    // edge: base + dir * q
    V3 base = rBC ? b : a;
    V3 dirv = rAB ? ab : (rAC ? ac : bc);
    V3 pe = base + dirv * q;
    // face: a + ab * v + ac * w
    float v = vb * q, w = vc * q;
    V3 pf = a + ab * v + ac * w;
    V3 pv = rA ? a : (rB ? b : c);
    V3 point = vtx ? pv : (edge ? pe : (face ? pf : pt));
    r.point = point;
    r.kind = vtx ? 0 : (edge ? 1 : (face ? 2 : 3));
    r.idx = vtx ? (rA ? 0u : (rB ? 1u : 2u)) : (edge ? (rAB ? 0u : (rAC ? 2u : 1u)) : (dot3(n, ap) >= 0.0f ? 0u : 1u));
    r.bc[0] = edge ? 1.0f - q : (face ? 1.0f - v - w : 0.0f);
    r.bc[1] = edge ? q : (face ? v : 0.0f);
    r.bc[2] = face ? w : 0.0f;
    r.inside = (vtx || edge || face) ? rel_eq3(point, pt) : true;
}

// point_tetrahedron.rs check_edge
__device__ __forceinline__ bool tet_edge(uint32_t i, V3 a, V3 nabc, V3 nabd, V3 ap, V3 ab, float ap_ab, float bp_ab, float& dabc,
                                         float& dabd, Proj& r) {
    float ab_ab = ap_ab - bp_ab;
    V3 x = cross3(ap, ab);
    dabc = dot3(x, nabc);
    dabd = dot3(x, nabd);
    if (ab_ab != 0.0f && dabc >= 0.0f && dabd >= 0.0f && ap_ab >= 0.0f && ap_ab <= ab_ab) {
        float u = ap_ab / ab_ab;
        r.bc[0] = 1.0f - u; r.bc[1] = u; r.bc[2] = 0.f;
        r.point = a + ab * u; r.inside = false; r.kind = 1; r.idx = i;
        return true;
    }
    return false;
}
// point_tetrahedron.rs check_face
__device__ __forceinline__ bool tet_face(uint32_t i, V3 a, V3 b, V3 c, V3 ap, V3 bp, V3 cp, V3 ab, V3 ac, V3 ad, float dabc, float dbca,
                                         float dacb, Proj& r) {
    if (dabc < 0.0f && dbca < 0.0f && dacb < 0.0f) {
        V3 n = cross3(ab, ac);
        if (dot3(n, ad) * dot3(n, ap) < 0.0f) {
            float nn = nrm(n);
            if (nn <= PB2_EPS) return false;  // try_normalize(DEFAULT_EPSILON)? -> None
            V3 normal = n / nn;
            float vc = dot3(normal, cross3(ap, bp));
            float va = dot3(normal, cross3(bp, cp));
            float vb = dot3(normal, cross3(cp, ap));
            float denom = va + vb + vc;
            float inv = 1.0f / denom;
            r.bc[0] = va * inv; r.bc[1] = vb * inv; r.bc[2] = vc * inv;
            r.point = a * r.bc[0] + b * r.bc[1] + c * r.bc[2];
            r.inside = false; r.kind = 2; r.idx = i;
            return true;
        }
    }
    return false;
}
// point_tetrahedron.rs:32-339, pt = origin, solid = true
static __device__ __noinline__ void project_origin_on_tetrahedron(V3 a, V3 b, V3 c, V3 d, Proj& r) {
    V3 pt = mk3(0.f, 0.f, 0.f);
    r.bc[0] = r.bc[1] = r.bc[2] = 0.f; r.idx = 0; r.inside = false;
    V3 ab = b - a, ac = c - a, ad = d - a, ap = pt - a;
    float ap_ab = dot3(ap, ab), ap_ac = dot3(ap, ac), ap_ad = dot3(ap, ad);
    if (ap_ab <= 0.0f && ap_ac <= 0.0f && ap_ad <= 0.0f) { r.point = a; r.kind = 0; r.idx = 0; return; }
    V3 bc = c - b, bd = d - b, bp = pt - b;
    float bp_bc = dot3(bp, bc), bp_bd = dot3(bp, bd), bp_ab = dot3(bp, ab);
    if (bp_bc <= 0.0f && bp_bd <= 0.0f && bp_ab >= 0.0f) { r.point = b; r.kind = 0; r.idx = 1; return; }
    V3 cd = d - c, cp = pt - c;
    float cp_ac = dot3(cp, ac), cp_bc = dot3(cp, bc), cp_cd = dot3(cp, cd);
    if (cp_cd <= 0.0f && cp_bc >= 0.0f && cp_ac >= 0.0f) { r.point = c; r.kind = 0; r.idx = 2; return; }
    V3 dp = pt - d;
    float dp_cd = dot3(dp, cd), dp_bd = dot3(dp, bd), dp_ad = dot3(dp, ad);
    if (dp_ad >= 0.0f && dp_bd >= 0.0f && dp_cd >= 0.0f) { r.point = d; r.kind = 0; r.idx = 3; return; }
    V3 nabc = cross3(ab, ac), nabd = cross3(ab, ad);
    float dabc, dabd, dacd, dacb, dadb, dadc, dbca, dbcd, dbdc, dbda, dcda, dcdb;
    if (tet_edge(0, a, nabc, nabd, ap, ab, ap_ab, bp_ab, dabc, dabd, r)) return;
    V3 nacd = cross3(ac, ad);
    if (tet_edge(1, a, nacd, -nabc, ap, ac, ap_ac, cp_ac, dacd, dacb, r)) return;
    if (tet_edge(2, a, -nabd, -nacd, ap, ad, ap_ad, dp_ad, dadb, dadc, r)) return;
    V3 nbcd = cross3(bc, bd);
    if (tet_edge(3, b, nabc, nbcd, bp, bc, bp_bc, cp_bc, dbca, dbcd, r)) return;
    if (tet_edge(4, b, -nbcd, nabd, bp, bd, bp_bd, dp_bd, dbdc, dbda, r)) return;
    if (tet_edge(5, c, nacd, nbcd, cp, cd, cp_cd, dp_cd, dcda, dcdb, r)) return;
    if (tet_face(0, a, b, c, ap, bp, cp, ab, ac, ad, dabc, dbca, dacb, r)) return;
    if (tet_face(1, a, b, d, ap, bp, dp, ab, ad, ac, dadb, dabd, dbda, r)) return;
    if (tet_face(2, a, c, d, ap, cp, dp, ac, ad, ab, dacd, dcda, dadc, r)) return;
    if (tet_face(3, b, c, d, bp, cp, dp, bc, bd, -ab, dbcd, dcdb, dbdc, r)) return;
    r.point = pt; r.inside = true; r.kind = 3;
}

// ---- Voronoi simplex (voronoi_simplex3.rs)
struct Simplex {
    CSO v[4];
    float proj[3];
    float prev_proj[3];
    uint8_t prev_v[4];
    int dim, prev_dim;
};
__device__ __forceinline__ void sx_swap(Simplex& s, int i, int j) {
    CSO t = s.v[i]; s.v[i] = s.v[j]; s.v[j] = t;
    uint8_t u = s.prev_v[i]; s.prev_v[i] = s.prev_v[j]; s.prev_v[j] = u;
}
__device__ __forceinline__ void sx_reset(Simplex& s, const CSO& pt) {
    s.dim = 0; s.prev_dim = 0; s.v[0] = pt;
    s.proj[0] = s.proj[1] = s.proj[2] = 0.f;
    s.prev_proj[0] = s.prev_proj[1] = s.prev_proj[2] = 0.f;
    s.prev_v[0] = 0; s.prev_v[1] = 1; s.prev_v[2] = 2; s.prev_v[3] = 3;
}
__device__ __forceinline__ bool sx_add_point(Simplex& s, const CSO& pt) {
    s.prev_dim = s.dim;
    s.prev_proj[0] = s.proj[0]; s.prev_proj[1] = s.proj[1]; s.prev_proj[2] = s.proj[2];
    s.prev_v[0] = 0; s.prev_v[1] = 1; s.prev_v[2] = 2; s.prev_v[3] = 3;
    if (s.dim == 0) {
        if (nrm2(s.v[0].point - pt.point) < PB2_GJK_EPS_TOL) return false;
    } else if (s.dim == 1) {
        V3 ab = s.v[1].point - s.v[0].point, ac = pt.point - s.v[0].point;
        if (nrm2(cross3(ab, ac)) < PB2_GJK_EPS_TOL) return false;
    } else {
        V3 ab = s.v[1].point - s.v[0].point, ac = s.v[2].point - s.v[0].point, ap = pt.point - s.v[0].point;
        V3 n = normalize3(cross3(ab, ac));
        if (fabsf(dot3(n, ap)) < PB2_GJK_EPS_TOL) return false;
    }
    s.dim += 1;
    s.v[s.dim] = pt;
    return true;
}
__device__ __forceinline__ V3 sx_project_origin_and_reduce(Simplex& s) {
    V3 origin = mk3(0.f, 0.f, 0.f);
    if (s.dim == 0) { s.proj[0] = 1.0f; return s.v[0].point; }
    if (s.dim == 1) {
        // point_segment.rs:49-84
        V3 a = s.v[0].point, b = s.v[1].point;
        V3 ab = b - a, ap = origin - a;
        float ab_ap = dot3(ab, ap), sqnab = nrm2(ab);
        if (ab_ap <= 0.0f) { s.proj[0] = 1.0f; s.dim = 0; return a; }
        if (ab_ap >= sqnab) { sx_swap(s, 0, 1); s.proj[0] = 1.0f; s.dim = 0; return b; }
        float u = ab_ap / sqnab;
        s.proj[0] = 1.0f - u; s.proj[1] = u;
        return a + ab * u;
    }
    Proj p;
    if (s.dim == 2) {
        project_on_triangle(s.v[0].point, s.v[1].point, s.v[2].point, origin, p);
        if (p.kind == 0) { sx_swap(s, 0, (int)p.idx); s.proj[0] = 1.0f; s.dim = 0; }
        else if (p.kind == 1) {
            if (p.idx == 0) { s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; }
            else if (p.idx == 1) { sx_swap(s, 0, 2); s.proj[0] = p.bc[1]; s.proj[1] = p.bc[0]; }
            else { sx_swap(s, 1, 2); s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; }
            s.dim = 1;
        } else if (p.kind == 2) { s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; s.proj[2] = p.bc[2]; }
        return p.point;
    }
    project_origin_on_tetrahedron(s.v[0].point, s.v[1].point, s.v[2].point, s.v[3].point, p);
    if (p.kind == 0) { sx_swap(s, 0, (int)p.idx); s.proj[0] = 1.0f; s.dim = 0; }
    else if (p.kind == 1) {
        switch (p.idx) {
            case 0: break;
            case 1: sx_swap(s, 1, 2); break;
            case 2: sx_swap(s, 1, 3); break;
            case 3: sx_swap(s, 0, 2); break;
            case 4: sx_swap(s, 0, 3); break;
            default: sx_swap(s, 0, 2); sx_swap(s, 1, 3); break;
        }
        if (p.idx == 3 || p.idx == 4) { s.proj[0] = p.bc[1]; s.proj[1] = p.bc[0]; }
        else { s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; }
        s.dim = 1;
    } else if (p.kind == 2) {
        switch (p.idx) {
            case 0: s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; s.proj[2] = p.bc[2]; break;
            case 1: s.v[2] = s.v[3]; s.proj[0] = p.bc[0]; s.proj[1] = p.bc[1]; s.proj[2] = p.bc[2]; break;
            case 2: s.v[1] = s.v[3]; s.proj[0] = p.bc[0]; s.proj[1] = p.bc[2]; s.proj[2] = p.bc[1]; break;
            default: s.v[0] = s.v[3]; s.proj[0] = p.bc[2]; s.proj[1] = p.bc[0]; s.proj[2] = p.bc[1]; break;
        }
        s.dim = 2;
    }
    return p.point;
}

// gjk.rs result()
__device__ __forceinline__ void gjk_witness(const Simplex& s, bool prev, V3& r0, V3& r1) {
    r0 = mk3(0.f, 0.f, 0.f); r1 = mk3(0.f, 0.f, 0.f);
    if (prev) {
        for (int i = 0; i < s.prev_dim + 1; ++i) {
            float coord = s.prev_proj[i];
            const CSO& p = s.v[s.prev_v[i]];
            r0 = r0 + p.o1 * coord; r1 = r1 + p.o2 * coord;
        }
    } else {
        for (int i = 0; i < s.dim + 1; ++i) {
            float coord = s.proj[i];
            const CSO& p = s.v[i];
            r0 = r0 + p.o1 * coord; r1 = r1 + p.o2 * coord;
        }
    }
}

enum { GJK_INTERSECTION = 0, GJK_CLOSEST_POINTS = 1, GJK_PROXIMITY = 2, GJK_NO_INTERSECTION = 3 };

// gjk::closest_points(pos12, g1, g2, max_dist, exact_dist = EXACT, simplex). EXACT = false (intersection_test) answers
// GJK_PROXIMITY as soon as a separating direction is known (gjk.rs:397-443).
template <bool EXACT = true>
__device__ __forceinline__ int gjk_closest_points(const Iso7& pos12, const DShape& g1, const DShape& g2, float max_dist, Simplex& s,
                                                  V3& p1, V3& p2, V3& out_dir) {
    const float eps_tol = PB2_GJK_EPS_TOL;
    const float eps_rel = sqrtf(eps_tol);
    V3 proj = sx_project_origin_and_reduce(s);
    V3 old_dir, dir;
    {
        V3 pd; float n;
        if (try_normalize_get(proj, 0.0f, pd, n)) old_dir = -pd;
        else return GJK_INTERSECTION;
    }
    float max_bound = FLT_MAX;
    int niter = 0;
    for (;;) {
        float old_max_bound = max_bound;
        float dist;
        if (try_normalize_get(-proj, eps_tol, dir, dist)) max_bound = dist;
        else return GJK_INTERSECTION;
        if (max_bound >= old_max_bound) {
            if (!EXACT) { out_dir = old_dir; return GJK_PROXIMITY; }
            gjk_witness(s, true, p1, p2); out_dir = old_dir; return GJK_CLOSEST_POINTS;
        }
        CSO cso = cso_from_shapes(pos12, g1, g2, dir);
        float min_bound = -dot3(dir, cso.point);
        if (min_bound > max_dist) { out_dir = dir; return GJK_NO_INTERSECTION; }
        else if (!EXACT && min_bound > 0.0f && max_bound <= max_dist) { out_dir = old_dir; return GJK_PROXIMITY; }
        else if (max_bound - min_bound <= eps_rel * max_bound) {
            if (!EXACT) { out_dir = dir; return GJK_PROXIMITY; }
            gjk_witness(s, false, p1, p2); out_dir = dir; return GJK_CLOSEST_POINTS;
        }
        if (!sx_add_point(s, cso)) {
            if (!EXACT) { out_dir = dir; return GJK_PROXIMITY; }
            gjk_witness(s, false, p1, p2); out_dir = dir; return GJK_CLOSEST_POINTS;
        }
        old_dir = dir;
        proj = sx_project_origin_and_reduce(s);
        if (s.dim == 3) {
            if (min_bound >= eps_tol) {
                if (!EXACT) { out_dir = old_dir; return GJK_PROXIMITY; }
                gjk_witness(s, true, p1, p2); out_dir = old_dir; return GJK_CLOSEST_POINTS;
            }
            return GJK_INTERSECTION;
        }
        niter += 1;
        if (niter == 100) { out_dir = mk3(1.f, 0.f, 0.f); return GJK_NO_INTERSECTION; }
    }
}

// ---- ray casts on support-mapped shapes
// minkowski_ray_cast (gjk.rs:660-795): ray cast on a Minkowski difference given by its support function `cso(dir)`;
// ray_toi_with_halfspace (ray_halfspace.rs:9-39) inlined. Returns the hit as (toi, outward normal); the simplex is left as
// the reference leaves it (directional_distance reads the witness points from it).
template <class CsoFn>
__device__ __forceinline__ bool minkowski_ray_cast(CsoFn cso, Simplex& s, V3 ro, V3 rd, float max_toi, float& toi, V3& normal) {
    const float eps_tol = PB2_GJK_EPS_TOL;
    const float eps_rel = sqrtf(eps_tol);
    float ray_length = nrm(rd);
    if (rel_eq(ray_length, 0.0f, PB2_EPS, PB2_EPS)) return false;
    float ltoi = 0.0f;
    V3 co = ro, cd = rd / ray_length;
    V3 dir = -cd, ldir = dir;
    {
        CSO c0 = cso(dir);
        c0.point = c0.point + (-co);
        sx_reset(s, c0);
    }
    V3 proj = sx_project_origin_and_reduce(s);
    float max_bound = FLT_MAX;
    int niter = 0;
    bool last_chance = false;
    for (;;) {
        float old_max_bound = max_bound;
        float dist;
        if (try_normalize_get(-proj, eps_tol, dir, dist)) max_bound = dist;
        else { toi = ltoi / ray_length; normal = ldir; return true; }
        CSO sp;
        if (max_bound >= old_max_bound) {
            last_chance = true;
            V3 p = proj + co;
            sp.point = p; sp.o1 = p; sp.o2 = mk3(0.f, 0.f, 0.f);
        } else {
            sp = cso(dir);
        }
        if (last_chance && ltoi > 0.0f) { toi = ltoi / ray_length; normal = ldir; return true; }
        float denom = dot3(dir, cd);
        bool some = false;
        float t = 0.0f;
        if (!rel_eq(denom, 0.0f, PB2_EPS, PB2_EPS)) {
            t = dot3(dir, sp.point - co) / denom;
            some = t >= 0.0f;
        }
        if (some) {
            if (dot3(dir, cd) < 0.0f && t > 0.0f) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return false;
                V3 shift = cd * t;
                co = co + shift;
                max_bound = FLT_MAX;
                for (int i = 0; i <= s.dim; ++i) s.v[i].point = s.v[i].point + (-shift);
                last_chance = false;
            }
        } else if (dot3(dir, cd) > eps_tol) {
            return false;
        }
        if (last_chance) return false;
        float min_bound = -dot3(dir, sp.point - co);
        if (max_bound - min_bound <= eps_rel * max_bound) return false;
        sp.point = sp.point + (-co);
        (void)sx_add_point(s, sp);
        proj = sx_project_origin_and_reduce(s);
        if (s.dim == 3) {
            if (min_bound >= eps_tol) return false;
            toi = ltoi / ray_length; normal = ldir; return true;
        }
        niter += 1;
        if (niter == 100) return false;
    }
}

// gjk::cast_local_ray (gjk.rs:519-534): g2 = ConstantOrigin, pos12 = identity
__device__ __forceinline__ bool gjk_cast_local_ray(const DShape& shape, Simplex& s, V3 ro, V3 rd, float max_toi, float& toi, V3& normal) {
    return minkowski_ray_cast([&](V3 dir) { return cso_make(ds_local_support(shape, dir), mk3(0.f, 0.f, 0.f)); }, s, ro, rd, max_toi, toi, normal);
}

// local_ray_intersection_with_support_map_with_params (ray_support_map.rs:19-72); the feature is FeatureId::Unknown.
__device__ __forceinline__ bool ray_support_map(const DShape& shape, V3 ro, V3 rd, float max_toi, bool solid, float& toi, V3& normal) {
    Simplex s;
    if (!gjk_cast_local_ray(shape, s, ro, rd, max_toi, toi, normal)) return false;
    if (!solid && toi == 0.0f) {
        // the ray starts inside the shape: cast it back from beyond the far side
        V3 ndir = rd / nrm(rd);
        V3 supp = ds_local_support(shape, ndir);
        const float eps = 0.001f;
        float shift = dot3(supp - ro, ndir) + eps;
        float t2; V3 n2;
        if (!gjk_cast_local_ray(shape, s, ro + ndir * shift, -rd, shift + eps, t2, n2)) return false;
        float t = shift - t2;
        if (!(t <= max_toi)) return false;
        toi = t; normal = -n2;
    }
    return true;
}

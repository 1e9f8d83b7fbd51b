// ORACLE — TEST INFRASTRUCTURE ONLY (see math.hpp header).
// Restatement of parry3d src/query/gjk/{gjk,voronoi_simplex3,cso_point,special_support_maps}.rs,
// src/query/point/{point_segment,point_triangle,point_tetrahedron}.rs and the support maps of
// src/shape/{cuboid,convex_polyhedron}.rs + src/utils/point_cloud_support_point.rs.
#pragma once
#include "math.hpp"
#include <vector>
#include <cassert>

namespace pb2o {

// ---------------------------------------------------------------- support maps (shape/support_map.rs:213-467)
struct SupportShape {
    enum Kind { CUBOID, CONVEX, CONSTANT_ORIGIN, TRIANGLE, BALL } kind;
    Vec3 half_extents;       // CUBOID
    const float* points;     // CONVEX: xyz packed
    uint32_t num_points;
    Real radius = 0.0f;         // BALL (shape/ball.rs:230-250)
    Real border_radius = 0.0f;  // > 0: RoundShapeRef around the shape (shape/round_shape.rs:317-350), local support only

    static SupportShape ball(Real r) { SupportShape s; s.kind = BALL; s.points = nullptr; s.num_points = 0; s.radius = r; return s; }
    SupportShape rounded(Real border) const { SupportShape s = *this; s.border_radius = border; return s; }

    static SupportShape cuboid(const Vec3& he) { SupportShape s; s.kind = CUBOID; s.half_extents = he; s.points = nullptr; s.num_points = 0; return s; }
    static SupportShape convex(const float* p, uint32_t n) { SupportShape s; s.kind = CONVEX; s.points = p; s.num_points = n; return s; }
    static SupportShape triangle(const float* p) { SupportShape s; s.kind = TRIANGLE; s.points = p; s.num_points = 3; return s; }
    static SupportShape constant_origin() { SupportShape s; s.kind = CONSTANT_ORIGIN; s.points = nullptr; s.num_points = 0; return s; }

    Vec3 local_support_point(const Vec3& dir) const {
        if (border_radius > 0.0f) {  // RoundShapeRef: inner.local_support_point_toward(unit dir) + unit dir * border_radius
            Vec3 nd = normalize(dir);
            SupportShape inner = *this; inner.border_radius = 0.0f;
            return inner.local_support_point(nd) + nd * border_radius;
        }
        switch (kind) {
            case BALL:  // local_support_point_toward(Unit::new_normalize(dir)) = dir * radius
                return normalize(dir) * radius;
            case CUBOID:  // cuboid.rs:452-457: dir.copy_sign_to(half_extents)
                return Vec3(copysignf(half_extents.x, dir.x), copysignf(half_extents.y, dir.y), copysignf(half_extents.z, dir.z));
            case CONVEX: {  // convex_polyhedron.rs:952-957 -> point_cloud_support_point_id (first max, strict >)
                uint32_t best = 0;
                Real best_dot = dot(Vec3(points[0], points[1], points[2]), dir);
                for (uint32_t i = 1; i < num_points; ++i) {
                    Real d = dot(Vec3(points[3 * i], points[3 * i + 1], points[3 * i + 2]), dir);
                    if (d > best_dot) { best_dot = d; best = i; }
                }
                return Vec3(points[3 * best], points[3 * best + 1], points[3 * best + 2]);
            }
            case TRIANGLE: {  // shape/triangle.rs:697-716 (its own tie-breaking, used by the epa3.rs regression test)
                Vec3 a(points[0], points[1], points[2]), b(points[3], points[4], points[5]), c(points[6], points[7], points[8]);
                Real d1 = dot(a, dir), d2 = dot(b, dir), d3 = dot(c, dir);
                if (d1 > d2) return d1 > d3 ? a : c;
                return d2 > d3 ? b : c;
            }
            default:  // special_support_maps.rs ConstantOrigin
                return Vec3();
        }
    }
    // SupportMap::support_point (support_map.rs:380-383); ConstantOrigin overrides it (translation only).
    Vec3 support_point(const Iso& m, const Vec3& dir) const {
        if (kind == CONSTANT_ORIGIN) return m.tra;
        if (kind == BALL) return m.tra + normalize(dir) * radius;  // ball.rs:232-239 (its own override: no rotation involved)
        Vec3 local_dir = m.inverse_transform_vector(dir);
        return m.transform_point(local_support_point(local_dir));
    }
};

// cso_point.rs:13-89
struct CSOPoint {
    Vec3 point, orig1, orig2;
    static CSOPoint make(const Vec3& o1, const Vec3& o2) { CSOPoint c; c.point = o1 - o2; c.orig1 = o1; c.orig2 = o2; return c; }
    static CSOPoint origin() { return make(Vec3(), Vec3()); }
    static CSOPoint from_shapes(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, const Vec3& dir) {
        Vec3 sp1 = g1.local_support_point(dir);
        Vec3 sp2 = g2.support_point(pos12, -dir);
        return make(sp1, sp2);
    }
};

// ---------------------------------------------------------------- point projections
static inline bool relative_eq_pt(const Vec3& a, const Vec3& b) { return relative_eq(a.x, b.x) && relative_eq(a.y, b.y) && relative_eq(a.z, b.z); }

// point_segment.rs:49-84. loc: 0 = OnVertex(0), 1 = OnVertex(1), 2 = OnEdge(bcoords)
struct SegProj { Vec3 point; bool inside; int loc; Real bc[2]; };
static inline SegProj project_on_segment(const Vec3& a, const Vec3& b, const Vec3& pt) {
    SegProj r; r.bc[0] = r.bc[1] = 0;
    Vec3 ab = b - a, ap = pt - a;
    Real ab_ap = dot(ab, ap), sqnab = norm_squared(ab);
    if (ab_ap <= 0.0f) { r.loc = 0; r.point = a; }
    else if (ab_ap >= sqnab) { r.loc = 1; r.point = b; }
    else { Real u = ab_ap / sqnab; r.bc[0] = 1.0f - u; r.bc[1] = u; r.loc = 2; r.point = a + ab * u; }
    r.inside = relative_eq_pt(r.point, pt);
    return r;
}

// point_triangle.rs:58-290. kind: 0 OnVertex(idx), 1 OnEdge(idx, bc[0..2]), 2 OnFace(idx, bc[0..3]), 3 OnSolid
struct TriProj { Vec3 point; bool inside; int kind; uint32_t idx; Real bc[3]; };
static inline TriProj project_on_triangle(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& pt, bool solid) {
    TriProj r; r.bc[0] = r.bc[1] = r.bc[2] = 0; r.idx = 0;
    auto result = [&](const Vec3& proj) { r.point = proj; r.inside = relative_eq_pt(proj, pt); };
    Vec3 ab = b - a, ac = c - a, ap = pt - a;
    Real ab_ap = dot(ab, ap), ac_ap = dot(ac, ap);
    if (ab_ap <= 0.0f && ac_ap <= 0.0f) { result(a); r.kind = 0; r.idx = 0; return r; }
    Vec3 bp = pt - b;
    Real ab_bp = dot(ab, bp), ac_bp = dot(ac, bp);
    if (ab_bp >= 0.0f && ac_bp <= ab_bp) { result(b); r.kind = 0; r.idx = 1; return r; }
    Vec3 cp = pt - c;
    Real ab_cp = dot(ab, cp), ac_cp = dot(ac, cp);
    if (ac_cp >= 0.0f && ab_cp <= ac_cp) { result(c); r.kind = 0; r.idx = 2; return r; }

    Vec3 bc = c - b;
    // stable_check_edges_voronoi (dim3, default features: n is the un-normalised cross product)
    int info;  // 0 AB, 1 AC, 2 BC, 3 face
    uint32_t face_side = 0; Real va = 0, vb = 0, vc = 0;
    {
        Vec3 n = cross(ab, ac);
        vc = dot(n, cross(ab, ap));
        if (vc < 0.0f && ab_ap >= 0.0f && ab_bp <= 0.0f) info = 0;
        else {
            vb = -dot(n, cross(ac, cp));
            if (vb < 0.0f && ac_ap >= 0.0f && ac_cp <= 0.0f) info = 1;
            else {
                va = dot(n, cross(bc, bp));
                if (va < 0.0f && ac_bp - ab_bp >= 0.0f && ab_cp - ac_cp >= 0.0f) info = 2;
                else { face_side = dot(n, ap) >= 0.0f ? 0 : 1; info = 3; }
            }
        }
    }
    if (info == 0) {
        Real v = ab_ap / norm_squared(ab);
        r.bc[0] = 1.0f - v; r.bc[1] = v; result(a + ab * v); r.kind = 1; r.idx = 0; return r;
    } else if (info == 1) {
        Real w = ac_ap / norm_squared(ac);
        r.bc[0] = 1.0f - w; r.bc[1] = w; result(a + ac * w); r.kind = 1; r.idx = 2; return r;
    } else if (info == 2) {
        Real w = dot(bc, bp) / norm_squared(bc);
        r.bc[0] = 1.0f - w; r.bc[1] = w; result(b + bc * w); r.kind = 1; r.idx = 1; return r;
    } else {
        if (va + vb + vc != 0.0f) {
            Real denom = 1.0f / (va + vb + vc);
            Real v = vb * denom, w = vc * denom;
            r.bc[0] = 1.0f - v - w; r.bc[1] = v; r.bc[2] = w;
            result(a + ab * v + ac * w); r.kind = 2; r.idx = face_side; return r;
        }
    }
    if (solid) { r.point = pt; r.inside = true; r.kind = 3; return r; }
    // non-solid fallback: closest edge (point_triangle.rs:243-287)
    Real v = ab_ap / (ab_ap - ab_bp);
    Real w = ac_ap / (ac_ap - ac_cp);
    Real u = (ac_bp - ab_bp) / (ac_bp - ab_bp + ab_cp - ac_cp);
    Real d_ab = norm_squared(ap) - (norm_squared(ab) * v * v);
    Real d_ac = norm_squared(ap) - (norm_squared(ac) * w * w);
    Real d_bc = norm_squared(bp) - (norm_squared(bc) * u * u);
    r.inside = true; r.kind = 1;
    if (d_ab < d_ac) {
        if (d_ab < d_bc) { r.bc[0] = 1.0f - v; r.bc[1] = v; r.point = a + ab * v; r.idx = 0; }
        else { r.bc[0] = 1.0f - u; r.bc[1] = u; r.point = b + bc * u; r.idx = 1; }
    } else if (d_ac < d_bc) { r.bc[0] = 1.0f - w; r.bc[1] = w; r.point = a + ac * w; r.idx = 2; }
    else { r.bc[0] = 1.0f - u; r.bc[1] = u; r.point = b + bc * u; r.idx = 1; }
    return r;
}
// TrianglePointLocation::barycentric_coordinates (shape/triangle.rs:121-148)
static inline bool tri_barycentric(const TriProj& p, Real out[3]) {
    out[0] = out[1] = out[2] = 0;
    if (p.kind == 0) { out[p.idx] = 1.0f; return true; }
    if (p.kind == 1) {
        int i0 = p.idx == 0 ? 0 : (p.idx == 1 ? 1 : 0), i1 = p.idx == 0 ? 1 : 2;
        out[i0] = p.bc[0]; out[i1] = p.bc[1]; return true;
    }
    if (p.kind == 2) { out[0] = p.bc[0]; out[1] = p.bc[1]; out[2] = p.bc[2]; return true; }
    return false;
}

// point_tetrahedron.rs:32-339 (solid = true). kind: 0 OnVertex, 1 OnEdge(idx, bc[0..2]), 2 OnFace(idx, bc[0..3]), 3 OnSolid
struct TetProj { Vec3 point; bool inside; int kind; uint32_t idx; Real bc[3]; };
static inline bool tet_check_edge(uint32_t i, const Vec3& a, const Vec3& nabc, const Vec3& nabd, const Vec3& ap, const Vec3& ab,
                                  Real ap_ab, Real bp_ab, Real& dabc, Real& dabd, TetProj& out) {
    Real ab_ab = ap_ab - bp_ab;
    Vec3 ap_x_ab = cross(ap, ab);
    dabc = dot(ap_x_ab, nabc);
    dabd = dot(ap_x_ab, nabd);
    if (ab_ab != 0.0f && dabc >= 0.0f && dabd >= 0.0f && ap_ab >= 0.0f && ap_ab <= ab_ab) {
        Real u = ap_ab / ab_ab;
        out.bc[0] = 1.0f - u; out.bc[1] = u; out.bc[2] = 0;
        out.point = a + ab * u; out.inside = false; out.kind = 1; out.idx = i;
        return true;
    }
    return false;
}
static inline bool tet_check_face(uint32_t i, const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& ap, const Vec3& bp, const Vec3& cp,
                                  const Vec3& ab, const Vec3& ac, const Vec3& ad, Real dabc, Real dbca, Real dacb, TetProj& out) {
    if (dabc < 0.0f && dbca < 0.0f && dacb < 0.0f) {
        Vec3 n = cross(ab, ac);
        if (dot(n, ad) * dot(n, ap) < 0.0f) {
            // n.try_normalize(DEFAULT_EPSILON): norm <= eps => None (check_face returns None)
            Real nn = norm(n);
            if (nn <= DEFAULT_EPSILON) return false;
            Vec3 normal = n / nn;
            Real vc = dot(normal, cross(ap, bp));
            Real va = dot(normal, cross(bp, cp));
            Real vb = dot(normal, cross(cp, ap));
            Real denom = va + vb + vc;
            assert(denom != 0.0f);
            Real inv = 1.0f / denom;
            out.bc[0] = va * inv; out.bc[1] = vb * inv; out.bc[2] = vc * inv;
            out.point = a * out.bc[0] + b * out.bc[1] + c * out.bc[2];
            out.inside = false; out.kind = 2; out.idx = i;
            return true;
        }
    }
    return false;
}
static inline TetProj project_on_tetrahedron(const Vec3& a, const Vec3& b, const Vec3& c, const Vec3& d, const Vec3& pt) {
    TetProj r; r.bc[0] = r.bc[1] = r.bc[2] = 0; r.idx = 0; r.inside = false;
    Vec3 ab = b - a, ac = c - a, ad = d - a, ap = pt - a;
    Real ap_ab = dot(ap, ab), ap_ac = dot(ap, ac), ap_ad = dot(ap, ad);
    if (ap_ab <= 0.0f && ap_ac <= 0.0f && ap_ad <= 0.0f) { r.point = a; r.kind = 0; r.idx = 0; return r; }
    Vec3 bc = c - b, bd = d - b, bp = pt - b;
    Real bp_bc = dot(bp, bc), bp_bd = dot(bp, bd), bp_ab = dot(bp, ab);
    if (bp_bc <= 0.0f && bp_bd <= 0.0f && bp_ab >= 0.0f) { r.point = b; r.kind = 0; r.idx = 1; return r; }
    Vec3 cd = d - c, cp = pt - c;
    Real cp_ac = dot(cp, ac), cp_bc = dot(cp, bc), cp_cd = dot(cp, cd);
    if (cp_cd <= 0.0f && cp_bc >= 0.0f && cp_ac >= 0.0f) { r.point = c; r.kind = 0; r.idx = 2; return r; }
    Vec3 dp = pt - d;
    Real dp_cd = dot(dp, cd), dp_bd = dot(dp, bd), dp_ad = dot(dp, ad);
    if (dp_ad >= 0.0f && dp_bd >= 0.0f && dp_cd >= 0.0f) { r.point = d; r.kind = 0; r.idx = 3; return r; }

    Vec3 nabc = cross(ab, ac), nabd = cross(ab, ad);
    Real dabc, dabd, dacd, dacb, dadb, dadc, dbca, dbcd, dbdc, dbda, dcda, dcdb;
    if (tet_check_edge(0, a, nabc, nabd, ap, ab, ap_ab, bp_ab, dabc, dabd, r)) return r;
    Vec3 nacd = cross(ac, ad);
    if (tet_check_edge(1, a, nacd, -nabc, ap, ac, ap_ac, cp_ac, dacd, dacb, r)) return r;
    if (tet_check_edge(2, a, -nabd, -nacd, ap, ad, ap_ad, dp_ad, dadb, dadc, r)) return r;
    Vec3 nbcd = cross(bc, bd);
    if (tet_check_edge(3, b, nabc, nbcd, bp, bc, bp_bc, cp_bc, dbca, dbcd, r)) return r;
    if (tet_check_edge(4, b, -nbcd, nabd, bp, bd, bp_bd, dp_bd, dbdc, dbda, r)) return r;
    if (tet_check_edge(5, c, nacd, nbcd, cp, cd, cp_cd, dp_cd, dcda, dcdb, r)) return r;

    if (tet_check_face(0, a, b, c, ap, bp, cp, ab, ac, ad, dabc, dbca, dacb, r)) return r;
    if (tet_check_face(1, a, b, d, ap, bp, dp, ab, ad, ac, dadb, dabd, dbda, r)) return r;
    if (tet_check_face(2, a, c, d, ap, cp, dp, ac, ad, ab, dacd, dcda, dadc, r)) return r;
    if (tet_check_face(3, b, c, d, bp, cp, dp, bc, bd, -ab, dbcd, dcdb, dbdc, r)) return r;
    r.point = pt; r.inside = true; r.kind = 3;
    return r;
}

// ---------------------------------------------------------------- gjk constants (gjk.rs:141-144)
static inline Real gjk_eps_tol() { return DEFAULT_EPSILON * 10.0f; }

// voronoi_simplex3.rs:14-351
struct VoronoiSimplex {
    size_t prev_vertices[4] = {0, 1, 2, 3};
    Real prev_proj[3] = {0, 0, 0};
    size_t prev_dim = 0;
    CSOPoint vertices[4];
    Real proj[3] = {0, 0, 0};
    size_t dim = 0;
    VoronoiSimplex() { for (auto& v : vertices) v = CSOPoint::origin(); }

    void swap(size_t i1, size_t i2) { std::swap(vertices[i1], vertices[i2]); std::swap(prev_vertices[i1], prev_vertices[i2]); }
    void reset(const CSOPoint& pt) { dim = 0; prev_dim = 0; vertices[0] = pt; }
    bool add_point(const CSOPoint& pt) {
        prev_dim = dim;
        for (int i = 0; i < 3; ++i) prev_proj[i] = proj[i];
        for (size_t i = 0; i < 4; ++i) prev_vertices[i] = i;
        if (dim == 0) {
            if (norm_squared(vertices[0].point - pt.point) < gjk_eps_tol()) return false;
        } else if (dim == 1) {
            Vec3 ab = vertices[1].point - vertices[0].point, ac = pt.point - vertices[0].point;
            if (norm_squared(cross(ab, ac)) < gjk_eps_tol()) return false;
        } else if (dim == 2) {
            Vec3 ab = vertices[1].point - vertices[0].point, ac = vertices[2].point - vertices[0].point, ap = pt.point - vertices[0].point;
            Vec3 n = normalize(cross(ab, ac));
            if (fabsf(dot(n, ap)) < gjk_eps_tol()) return false;
        } else { assert(false && "unreachable"); }
        dim += 1;
        vertices[dim] = pt;
        return true;
    }
    Real proj_coord(size_t i) const { return proj[i]; }
    const CSOPoint& point(size_t i) const { return vertices[i]; }
    Real prev_proj_coord(size_t i) const { return prev_proj[i]; }
    const CSOPoint& prev_point(size_t i) const { return vertices[prev_vertices[i]]; }
    size_t dimension() const { return dim; }
    size_t prev_dimension() const { return prev_dim; }

    Vec3 project_origin_and_reduce() {
        Vec3 origin;
        if (dim == 0) { proj[0] = 1.0f; return vertices[0].point; }
        if (dim == 1) {
            SegProj p = project_on_segment(vertices[0].point, vertices[1].point, origin);
            if (p.loc == 0) { proj[0] = 1.0f; dim = 0; }
            else if (p.loc == 1) { swap(0, 1); proj[0] = 1.0f; dim = 0; }
            else { proj[0] = p.bc[0]; proj[1] = p.bc[1]; }
            return p.point;
        }
        if (dim == 2) {
            TriProj p = project_on_triangle(vertices[0].point, vertices[1].point, vertices[2].point, origin, true);
            if (p.kind == 0) { swap(0, p.idx); proj[0] = 1.0f; dim = 0; }
            else if (p.kind == 1 && p.idx == 0) { proj[0] = p.bc[0]; proj[1] = p.bc[1]; dim = 1; }
            else if (p.kind == 1 && p.idx == 1) { swap(0, 2); proj[0] = p.bc[1]; proj[1] = p.bc[0]; dim = 1; }
            else if (p.kind == 1 && p.idx == 2) { swap(1, 2); proj[0] = p.bc[0]; proj[1] = p.bc[1]; dim = 1; }
            else if (p.kind == 2) { proj[0] = p.bc[0]; proj[1] = p.bc[1]; proj[2] = p.bc[2]; }
            return p.point;
        }
        assert(dim == 3);
        TetProj p = project_on_tetrahedron(vertices[0].point, vertices[1].point, vertices[2].point, vertices[3].point, origin);
        if (p.kind == 0) { swap(0, p.idx); proj[0] = 1.0f; dim = 0; }
        else if (p.kind == 1) {
            switch (p.idx) {
                case 0: break;
                case 1: swap(1, 2); break;
                case 2: swap(1, 3); break;
                case 3: swap(0, 2); break;
                case 4: swap(0, 3); break;
                case 5: swap(0, 2); swap(1, 3); break;
            }
            if (p.idx == 3 || p.idx == 4) { proj[0] = p.bc[1]; proj[1] = p.bc[0]; }
            else { proj[0] = p.bc[0]; proj[1] = p.bc[1]; }
            dim = 1;
        } else if (p.kind == 2) {
            switch (p.idx) {
                case 0: proj[0] = p.bc[0]; proj[1] = p.bc[1]; proj[2] = p.bc[2]; break;
                case 1: vertices[2] = vertices[3]; proj[0] = p.bc[0]; proj[1] = p.bc[1]; proj[2] = p.bc[2]; break;
                case 2: vertices[1] = vertices[3]; proj[0] = p.bc[0]; proj[1] = p.bc[2]; proj[2] = p.bc[1]; break;
                case 3: vertices[0] = vertices[3]; proj[0] = p.bc[2]; proj[1] = p.bc[0]; proj[2] = p.bc[1]; break;
            }
            dim = 2;
        }
        return p.point;
    }
};

// gjk.rs GJKResult
struct GJKResult {
    enum Kind { INTERSECTION, CLOSEST_POINTS, PROXIMITY, NO_INTERSECTION } kind;
    Vec3 p1, p2, dir;
    int niter = 0;
};

// gjk.rs:797-818
static inline void gjk_result(const VoronoiSimplex& s, bool prev, Vec3& r0, Vec3& r1) {
    r0 = Vec3(); r1 = Vec3();
    if (prev) {
        for (size_t i = 0; i < s.prev_dimension() + 1; ++i) {
            Real coord = s.prev_proj_coord(i);
            const CSOPoint& p = s.prev_point(i);
            r0 = r0 + p.orig1 * coord; r1 = r1 + p.orig2 * coord;
        }
    } else {
        for (size_t i = 0; i < s.dimension() + 1; ++i) {
            Real coord = s.proj_coord(i);
            const CSOPoint& p = s.point(i);
            r0 = r0 + p.orig1 * coord; r1 = r1 + p.orig2 * coord;
        }
    }
}

// gjk.rs:353-453 with exact_dist = true
// gjk.rs:353-453. exact_dist = false (intersection_test) answers PROXIMITY as soon as a separating direction is known.
static inline GJKResult gjk_closest_points(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, Real max_dist, VoronoiSimplex& simplex,
                                           bool exact_dist = true) {
    const Real eps_tol = gjk_eps_tol();
    const Real eps_rel = sqrtf(eps_tol);
    GJKResult res;
    Vec3 proj = simplex.project_origin_and_reduce();
    Vec3 old_dir, dir;
    {
        Vec3 pd;
        if (try_normalize(proj, 0.0f, pd)) old_dir = -pd;
        else { res.kind = GJKResult::INTERSECTION; return res; }
    }
    Real max_bound = REAL_MAX;
    int niter = 0;
    for (;;) {
        Real old_max_bound = max_bound;
        Real dist;
        if (try_normalize_and_get(-proj, eps_tol, dir, dist)) max_bound = dist;
        else { res.kind = GJKResult::INTERSECTION; res.niter = niter; return res; }
        if (max_bound >= old_max_bound) {
            if (!exact_dist) { res.kind = GJKResult::PROXIMITY; res.dir = old_dir; res.niter = niter; return res; }
            res.kind = GJKResult::CLOSEST_POINTS; gjk_result(simplex, true, res.p1, res.p2); res.dir = old_dir; res.niter = niter; return res;
        }
        CSOPoint cso = CSOPoint::from_shapes(pos12, g1, g2, dir);
        Real min_bound = -dot(dir, cso.point);
        assert(std::isfinite(min_bound));
        if (min_bound > max_dist) { res.kind = GJKResult::NO_INTERSECTION; res.dir = dir; res.niter = niter; return res; }
        else if (!exact_dist && min_bound > 0.0f && max_bound <= max_dist) { res.kind = GJKResult::PROXIMITY; res.dir = old_dir; res.niter = niter; return res; }
        else if (max_bound - min_bound <= eps_rel * max_bound) {
            if (!exact_dist) { res.kind = GJKResult::PROXIMITY; res.dir = dir; res.niter = niter; return res; }
            res.kind = GJKResult::CLOSEST_POINTS; gjk_result(simplex, false, res.p1, res.p2); res.dir = dir; res.niter = niter; return res;
        }
        if (!simplex.add_point(cso)) {
            if (!exact_dist) { res.kind = GJKResult::PROXIMITY; res.dir = dir; res.niter = niter; return res; }
            res.kind = GJKResult::CLOSEST_POINTS; gjk_result(simplex, false, res.p1, res.p2); res.dir = dir; res.niter = niter; return res;
        }
        old_dir = dir;
        proj = simplex.project_origin_and_reduce();
        if (simplex.dimension() == 3) {
            if (min_bound >= eps_tol) {
                if (!exact_dist) { res.kind = GJKResult::PROXIMITY; res.dir = old_dir; res.niter = niter; return res; }
                res.kind = GJKResult::CLOSEST_POINTS; gjk_result(simplex, true, res.p1, res.p2); res.dir = old_dir; res.niter = niter; return res;
            }
            res.kind = GJKResult::INTERSECTION; res.niter = niter; return res;
        }
        niter += 1;
        if (niter == 100) { res.kind = GJKResult::NO_INTERSECTION; res.dir = Vec3(1, 0, 0); res.niter = niter; return res; }
    }
}

}  // namespace pb2o

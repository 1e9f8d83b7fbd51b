// ORACLE — TEST INFRASTRUCTURE ONLY (see math.hpp header).
// Restatement of parry3d src/query/contact/{contact_shape_shape,contact_ball_ball,contact_ball_convex_polyhedron,
// contact_support_map_support_map}.rs, DefaultQueryDispatcher::contact (default_query_dispatcher.rs:302-356),
// src/query/point/{point_aabb,point_cuboid,point_support_map}.rs and Cuboid::feature_normal (shape/cuboid.rs:401-448).
#pragma once
#include "epa.hpp"

namespace pb2o {

struct Contact { Vec3 point1, point2, normal1, normal2; Real dist; };

enum ShapeKind { SHAPE_BALL = 0, SHAPE_CUBOID = 1, SHAPE_CONVEX = 2, SHAPE_TRIANGLE = 3 /* oracle-only, for the reference's EPA regression test */ };
struct ShapeRef {
    int kind;
    Real radius;
    Vec3 half_extents;
    const float* points;
    uint32_t num_points;
    SupportShape support() const {
        if (kind == SHAPE_CUBOID) return SupportShape::cuboid(half_extents);
        if (kind == SHAPE_TRIANGLE) return SupportShape::triangle(points);
        return SupportShape::convex(points, num_points);
    }
};

enum ContactStatus { CONTACT_NONE = 0, CONTACT_SOME = 1, CONTACT_UNSUPPORTED = 2, CONTACT_NEEDS_TOPOLOGY = 3 };

// contact_ball_ball.rs:9-42
static inline bool contact_ball_ball(const Iso& pos12, Real r1, Real r2, Real prediction, Contact& c) {
    Vec3 center2_1 = pos12.tra;
    Real d2 = norm_squared(center2_1);
    Real sum_radius = r1 + r2;
    Real sre = sum_radius + prediction;
    if (d2 < sre * sre) {
        Vec3 normal1 = d2 != 0.0f ? normalize(center2_1) : Vec3(1, 0, 0);
        Vec3 normal2 = -pos12.inverse_transform_vector(normal1);
        c.point1 = normal1 * r1; c.point2 = normal2 * r2; c.normal1 = normal1; c.normal2 = normal2;
        c.dist = sqrtf(d2) - sum_radius;
        return true;
    }
    return false;
}

// FeatureId on a cuboid: kind 0 vertex, 1 edge, 2 face, 3 unknown
struct Feature { int kind; uint32_t id; };

// point_aabb.rs:9-132 on [-he, he] (point_cuboid.rs:6-37): project_local_point_and_get_feature (solid = false)
static inline void cuboid_project_point_and_get_feature(const Vec3& he, const Vec3& pt, Vec3& proj, bool& inside, Feature& feat) {
    Vec3 mins = -he, maxs = he;
    Vec3 mins_pt = mins - pt, pt_maxs = pt - maxs;
    Vec3 zero;
    Vec3 shift = vsup(mins_pt, zero) - vsup(pt_maxs, zero);
    inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
    if (!inside) { proj = pt + shift; }
    else {
        Real best = -REAL_MAX; bool is_mins = false; int best_id = 0;
        for (int i = 0; i < 3; ++i) {
            Real a = mins_pt[i], b = pt_maxs[i];
            if (a < b) { if (b > best) { best_id = i; is_mins = false; best = b; } }
            else if (a > best) { best_id = i; is_mins = true; best = a; }
        }
        shift = Vec3();
        shift[best_id] = is_mins ? best : -best;
        proj = pt + shift;
    }
    int nzero = 0, last_zero = 0, last_not_zero = 0;
    for (int i = 0; i < 3; ++i) { if (shift[i] == 0.0f) { nzero++; last_zero = i; } else last_not_zero = i; }
    Vec3 ctr = center(mins, maxs);
    if (nzero == 3) {
        for (int i = 0; i < 3; ++i) {
            if (proj[i] > maxs[i] - DEFAULT_EPSILON) { feat = Feature{2, (uint32_t)i}; return; }
            if (proj[i] <= mins[i] + DEFAULT_EPSILON) { feat = Feature{2, (uint32_t)(i + 3)}; return; }
        }
        feat = Feature{3, 0};
    } else if (nzero == 2) {
        feat = proj[last_not_zero] < ctr[last_not_zero] ? Feature{2, (uint32_t)(last_not_zero + 3)} : Feature{2, (uint32_t)last_not_zero};
    } else {
        uint32_t id = 0;
        for (int i = 0; i < 3; ++i) if (proj[i] < ctr[i]) id |= 1u << i;
        feat = nzero == 0 ? Feature{0, id} : Feature{1, (id << 2) | (uint32_t)last_zero};
    }
}
// cuboid.rs:401-448
static inline bool cuboid_feature_normal(const Feature& f, Vec3& n) {
    Vec3 dir;
    if (f.kind == 2) { if (f.id < 3) dir[f.id] = 1.0f; else dir[f.id - 3] = -1.0f; n = dir; return true; }
    if (f.kind == 1) {
        uint32_t edge = f.id & 3u, face1 = (edge + 1) % 3, face2 = (edge + 2) % 3, signs = f.id >> 2;
        dir[face1] = (signs & (1u << face1)) ? -1.0f : 1.0f;
        dir[face2] = (signs & (1u << face2)) ? -1.0f : 1.0f;
        n = normalize(dir); return true;
    }
    if (f.kind == 0) {
        for (int i = 0; i < 3; ++i) dir[i] = (f.id & (1u << i)) ? -1.0f : 1.0f;
        n = normalize(dir); return true;
    }
    return false;
}

// point_support_map.rs:17-52 local_point_projection_on_support_map(shape, simplex, point, solid = false)
static inline void hull_project_point(const SupportShape& shape, const Vec3& point, Vec3& proj, bool& inside) {
    Iso m(Quat(), -point), m_inv(Quat(), point);
    Vec3 dir;
    if (!try_normalize(-m.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    SupportShape origin = SupportShape::constant_origin();
    VoronoiSimplex simplex;
    simplex.reset(CSOPoint::from_shapes(m_inv, shape, origin, dir));
    // gjk::project_origin(&m, shape, simplex) (gjk.rs:306-324)
    Iso minv = m.inverse();
    GJKResult r = gjk_closest_points(minv, shape, origin, REAL_MAX, simplex);
    if (r.kind == GJKResult::CLOSEST_POINTS) { proj = r.p1; inside = false; return; }
    assert(r.kind == GJKResult::INTERSECTION);
    EPA epa;
    Vec3 p1, p2, n;
    if (epa.closest_points(minv, shape, origin, simplex, p1, p2, n)) { proj = p1; inside = true; return; }
    proj = point; inside = true;
}

// contact_ball_convex_polyhedron.rs:26-63
static inline int contact_convex_polyhedron_ball(const Iso& pos12, const ShapeRef& shape1, Real radius2, Real prediction, Contact& c) {
    Vec3 center2_1 = pos12.tra;
    Vec3 proj; bool inside; Feature f1{3, 0};
    if (shape1.kind == SHAPE_CUBOID) cuboid_project_point_and_get_feature(shape1.half_extents, center2_1, proj, inside, f1);
    else if (shape1.kind == SHAPE_TRIANGLE) {
        // PointQuery for Triangle: project_local_point_and_get_feature = ..._and_get_location(pt, solid = true) (point_triangle.rs:27-47)
        TriProj p = project_on_triangle(ld3(shape1.points), ld3(shape1.points + 3), ld3(shape1.points + 6), center2_1, true);
        proj = p.point; inside = p.inside;
    }
    else hull_project_point(shape1.support(), center2_1, proj, inside);
    Real dist; Vec3 normal1, dir1; Real len;
    if (try_normalize_and_get(proj - center2_1, DEFAULT_EPSILON, dir1, len)) {
        if (inside) { dist = -len - radius2; normal1 = dir1; }
        else { dist = len - radius2; normal1 = -dir1; }
    } else {
        dist = -radius2;
        if (shape1.kind == SHAPE_TRIANGLE) {
            // Triangle::feature_normal_at_point = Triangle::normal() for every feature (shape.rs:919-928, triangle.rs:226-228,626-628)
            Vec3 a = ld3(shape1.points), b = ld3(shape1.points + 3), c3 = ld3(shape1.points + 6);
            if (!try_normalize(cross(b - a, c3 - a), DEFAULT_EPSILON, normal1)) {
                if (!try_normalize(proj, DEFAULT_EPSILON, normal1)) normal1 = Vec3(0, 1, 0);
            }
        } else {
        // ConvexPolyhedron::project_local_point_and_get_feature (point_support_map.rs:62-77) forms the feature from the same vector
        // (point - proj, negated when inside) with the same Unit::try_new(.., DEFAULT_EPSILON): whenever this branch runs that test has
        // failed too, the feature is FeatureId::Unknown and feature_normal (convex_polyhedron.rs:925-949) is None — no topology needed
        if (shape1.kind != SHAPE_CUBOID || !cuboid_feature_normal(f1, normal1)) {
            if (!try_normalize(proj, DEFAULT_EPSILON, normal1)) normal1 = Vec3(0, 1, 0);
        }
        }
    }
    if (dist <= prediction) {
        Vec3 normal2 = pos12.inverse_transform_vector(-normal1);
        c.point2 = normal2 * radius2; c.point1 = proj; c.normal1 = normal1; c.normal2 = normal2; c.dist = dist;
        return CONTACT_SOME;
    }
    return CONTACT_NONE;
}
// contact_ball_convex_polyhedron.rs:12-20
static inline int contact_ball_convex_polyhedron(const Iso& pos12, Real radius1, const ShapeRef& shape2, Real prediction, Contact& c) {
    int st = contact_convex_polyhedron_ball(pos12.inverse(), shape2, radius1, prediction, c);
    if (st == CONTACT_SOME) { std::swap(c.point1, c.point2); std::swap(c.normal1, c.normal2); }
    return st;
}

struct GjkEpaStats { int gjk_iters = 0; bool used_epa = false; EpaStats epa; };

// contact_support_map_support_map.rs:10-77
// (with_params, :40-61: a caller-supplied init_dir — contact_manifolds_pfm_pfm.rs:66 passes last frame's manifold normal — replaces
// the default first direction)
static inline int contact_support_map_support_map(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, Real prediction, Contact& c,
                                                  GjkEpaStats* stats = nullptr, const Vec3* init_dir = nullptr, Vec3* noint_dir = nullptr) {
    // noint_dir: the `dir` of GJKResult::NoIntersection(dir) when that is the answer (gjk.rs:387: the last search direction; +x after
    // 100 iterations or when EPA fails, contact_support_map_support_map.rs:76) — contact_manifolds_pfm_pfm.rs:151-154 caches it
    VoronoiSimplex simplex;
    Vec3 dir;
    if (init_dir) dir = *init_dir;
    else if (!try_normalize(pos12.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    simplex.reset(CSOPoint::from_shapes(pos12, g1, g2, dir));
    GJKResult r = gjk_closest_points(pos12, g1, g2, prediction, simplex);
    if (stats) stats->gjk_iters = r.niter;
    Vec3 p1, p2_1, n1;
    if (r.kind == GJKResult::INTERSECTION) {
        EPA epa;
        bool ok = epa.closest_points(pos12, g1, g2, simplex, p1, p2_1, n1);
        if (stats) { stats->used_epa = true; stats->epa = epa.stats; }
        if (!ok) { if (noint_dir) *noint_dir = Vec3(1, 0, 0); return CONTACT_NONE; }  // "Everything failed" => NoIntersection(+x) => None
    } else if (r.kind == GJKResult::CLOSEST_POINTS) { p1 = r.p1; p2_1 = r.p2; n1 = r.dir; }
    else { if (noint_dir) *noint_dir = r.dir; return CONTACT_NONE; }
    c.dist = dot(p2_1 - p1, n1);
    c.point1 = p1;
    c.point2 = pos12.inverse_transform_point(p2_1);
    c.normal1 = n1;
    c.normal2 = pos12.inverse_transform_vector(-n1);
    return CONTACT_SOME;
}

// DefaultQueryDispatcher::contact (default_query_dispatcher.rs:302-356), shapes restricted to Ball/Cuboid/ConvexPolyhedron
static inline int dispatch_contact(const Iso& pos12, const ShapeRef& s1, const ShapeRef& s2, Real prediction, Contact& c, GjkEpaStats* stats = nullptr) {
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) return contact_ball_ball(pos12, s1.radius, s2.radius, prediction, c) ? CONTACT_SOME : CONTACT_NONE;
    if (s1.kind == SHAPE_BALL) return contact_ball_convex_polyhedron(pos12, s1.radius, s2, prediction, c);
    if (s2.kind == SHAPE_BALL) return contact_convex_polyhedron_ball(pos12, s1, s2.radius, prediction, c);
    return contact_support_map_support_map(pos12, s1.support(), s2.support(), prediction, c, stats);
}

// query::contact (contact_shape_shape.rs:123-138)
static inline int query_contact(const Iso& pos1, const ShapeRef& g1, const Iso& pos2, const ShapeRef& g2, Real prediction, Contact& c, GjkEpaStats* stats = nullptr) {
    Iso pos12 = pos1.inv_mul(pos2);
    int st = dispatch_contact(pos12, g1, g2, prediction, c, stats);
    if (st == CONTACT_SOME) {  // Contact::transform_by_mut (contact.rs:171)
        c.point1 = pos1.transform_point(c.point1);
        c.point2 = pos2.transform_point(c.point2);
        c.normal1 = pos1.transform_vector(c.normal1);
        c.normal2 = pos2.transform_vector(c.normal2);
    }
    return st;
}

// PointQuery::project_local_point(pt, solid = true) for Cuboid (point_cuboid.rs:8-12 -> point_aabb.rs:9-60) and ConvexPolyhedron
// (point_support_map.rs:17-52: GJK projection; an inside point projects on itself when solid).
static inline void project_local_point_solid(const ShapeRef& s, const Vec3& pt, Vec3& proj, bool& inside) {
    if (s.kind == SHAPE_CUBOID) {
        Vec3 mins = -s.half_extents, maxs = s.half_extents, zero;
        Vec3 shift = vsup(mins - pt, zero) - vsup(pt - maxs, zero);
        inside = shift.x == 0.0f && shift.y == 0.0f && shift.z == 0.0f;
        proj = inside ? pt : pt + shift;
        return;
    }
    if (s.kind == SHAPE_TRIANGLE) {   // PointQuery for Triangle: its own Voronoi-region projection (point_triangle.rs:17-25), not GJK
        TriProj p = project_on_triangle(ld3(s.points), ld3(s.points + 3), ld3(s.points + 6), pt, true);
        proj = p.point; inside = p.inside;
        return;
    }
    SupportShape shape = s.support();
    Iso m(Quat(), -pt), m_inv(Quat(), pt);
    Vec3 dir;
    if (!try_normalize(-m.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    SupportShape origin = SupportShape::constant_origin();
    VoronoiSimplex simplex;
    simplex.reset(CSOPoint::from_shapes(m_inv, shape, origin, dir));
    GJKResult r = gjk_closest_points(m.inverse(), shape, origin, REAL_MAX, simplex);
    if (r.kind == GJKResult::CLOSEST_POINTS) { proj = r.p1; inside = false; }
    else { proj = pt; inside = true; }
}

// ---- SAT (sat_cuboid_cuboid.rs)
static inline Vec3 cuboid_local_support(const Vec3& he, const Vec3& dir) {
    return Vec3(copysignf(he.x, dir.x), copysignf(he.y, dir.y), copysignf(he.z, dir.z));
}
// :5-22
static inline void sat_separation_wrt_local_line(const Vec3& he1, const Vec3& he2, const Iso& pos12, const Vec3& axis_in, Real& sep, Vec3& axis1) {
    Real signum = copysignf(1.0f, dot(pos12.tra, axis_in));
    axis1 = axis_in * signum;
    Vec3 axis2 = pos12.inverse_transform_vector(-axis1);
    Vec3 local_pt1 = cuboid_local_support(he1, axis1);
    Vec3 local_pt2 = cuboid_local_support(he2, axis2);
    Vec3 pt2 = pos12.transform_point(local_pt2);
    sep = dot(pt2 - local_pt1, axis1);
}
// :24-77
static inline void sat_find_separating_edge_twoway(const Vec3& he1, const Vec3& he2, const Iso& pos12, Real& best_sep, Vec3& best_dir) {
    best_sep = -REAL_MAX; best_dir = Vec3();
    Vec3 x2 = pos12.transform_vector(Vec3(1, 0, 0)), y2 = pos12.transform_vector(Vec3(0, 1, 0)), z2 = pos12.transform_vector(Vec3(0, 0, 1));
    Vec3 axes[9] = {Vec3(0.0f, -x2.z, x2.y), Vec3(x2.z, 0.0f, -x2.x), Vec3(-x2.y, x2.x, 0.0f),
                    Vec3(0.0f, -y2.z, y2.y), Vec3(y2.z, 0.0f, -y2.x), Vec3(-y2.y, y2.x, 0.0f),
                    Vec3(0.0f, -z2.z, z2.y), Vec3(z2.z, 0.0f, -z2.x), Vec3(-z2.y, z2.x, 0.0f)};
    for (int k = 0; k < 9; ++k) {
        Real n = norm(axes[k]);
        if (n > DEFAULT_EPSILON) {
            Real sep; Vec3 a1;
            sat_separation_wrt_local_line(he1, he2, pos12, axes[k] / n, sep, a1);
            if (sep > best_sep) { best_sep = sep; best_dir = a1; }
        }
    }
}
// :79-110
static inline void sat_find_separating_normal_oneway(const Vec3& he1, const Vec3& he2, const Iso& pos12, Real& best_sep, Vec3& best_dir) {
    best_sep = -REAL_MAX; best_dir = Vec3();
    for (int i = 0; i < 3; ++i) {
        Real sign = copysignf(1.0f, pos12.tra[i]);
        Vec3 axis1; axis1[i] = sign;
        Vec3 axis2 = pos12.inverse_transform_vector(-axis1);
        Vec3 local_pt2 = cuboid_local_support(he2, axis2);
        Vec3 pt2 = pos12.transform_point(local_pt2);
        Real sep = pt2[i] * sign - he1[i];
        if (sep > best_sep) { best_sep = sep; best_dir = axis1; }
    }
}

// approx::ulps_eq! defaults for f32: epsilon = f32::EPSILON, max_ulps = 4
static inline bool ulps_eq(Real a, Real b) {
    if (fabsf(a - b) <= FLT_EPSILON) return true;
    if (std::signbit(a) != std::signbit(b)) return false;
    int32_t ia, ib; memcpy(&ia, &a, 4); memcpy(&ib, &b, 4);
    int64_t d = (int64_t)ia - (int64_t)ib; if (d < 0) d = -d;
    return d <= 4;
}

// ---- cuboid-cuboid arms of distance / intersection_test (SAT based)
static inline int iamin3(const Vec3& v) {   // nalgebra iamin: first strict minimum of |x|
    int i = 0; Real best = fabsf(v.x);
    if (fabsf(v.y) < best) { best = fabsf(v.y); i = 1; }
    if (fabsf(v.z) < best) i = 2;
    return i;
}
// Cuboid::local_support_edge_segment (shape/cuboid.rs:249-263)
static inline void cuboid_local_support_edge_segment(const Vec3& he, const Vec3& dir, Vec3& a, Vec3& b) {
    int i = iamin3(dir), j = (i + 1) % 3, k = (i + 2) % 3;
    a = Vec3(); a[i] = he[i]; a[j] = copysignf(he[j], dir[j]); a[k] = copysignf(he[k], dir[k]);
    b = a; b[i] = -he[i];
}
static inline Real na_clamp(Real v, Real lo, Real hi) { return v < lo ? lo : (v > hi ? hi : v); }
// closest_points_segment_segment_with_locations_nD (closest_points_segment_segment.rs:36-107): parameters s on seg1, t on seg2
static inline void segment_segment_params(const Vec3& a1, const Vec3& b1, const Vec3& a2, const Vec3& b2, Real& s, Real& t) {
    Vec3 d1 = b1 - a1, d2 = b2 - a2, r = a1 - a2;
    Real a = norm_squared(d1), e = norm_squared(d2), f = dot(d2, r);
    const Real eps = DEFAULT_EPSILON;
    if (a <= eps && e <= eps) { s = 0.0f; t = 0.0f; }
    else if (a <= eps) { s = 0.0f; t = na_clamp(f / e, 0.0f, 1.0f); }
    else {
        Real c = dot(d1, r);
        if (e <= eps) { t = 0.0f; s = na_clamp(-c / a, 0.0f, 1.0f); }
        else {
            Real b = dot(d1, d2), ae = a * e, bb = b * b, denom = ae - bb;
            if (denom > eps && !ulps_eq(ae, bb)) s = na_clamp((b * f - c * e) / denom, 0.0f, 1.0f);
            else s = 0.0f;
            t = (b * s + f) / e;
            if (t < 0.0f) { t = 0.0f; s = na_clamp(-c / a, 0.0f, 1.0f); }
            else if (t > 1.0f) { t = 1.0f; s = na_clamp((b - c) / a, 0.0f, 1.0f); }
        }
    }
}
// Segment::point_at (shape/segment.rs:410-419) of the location the reference derives from a parameter (:88-104)
static inline Vec3 segment_point_at_param(const Vec3& a, const Vec3& b, Real s) {
    if (s == 0.0f) return a;
    if (s == 1.0f) return b;
    return a * (1.0f - s) + b * s;
}
// intersection_test_cuboid_cuboid (intersection_test_cuboid_cuboid.rs:6-31)
static inline bool intersection_test_cuboid_cuboid(const Iso& pos12, const Vec3& he1, const Vec3& he2) {
    Real sep; Vec3 dir;
    sat_find_separating_normal_oneway(he1, he2, pos12, sep, dir);
    if (sep > 0.0f) return false;
    Iso pos21 = pos12.inverse();
    sat_find_separating_normal_oneway(he2, he1, pos21, sep, dir);
    if (sep > 0.0f) return false;
    sat_find_separating_edge_twoway(he1, he2, pos12, sep, dir);
    return sep <= 0.0f;
}
// distance_cuboid_cuboid (distance_cuboid_cuboid.rs:6-12) = closest_points_cuboid_cuboid with margin = Real::MAX
// (closest_points_cuboid_cuboid.rs:6-84), WithinMargin(p1, p2) -> na::distance(p1, pos12 * p2), anything else -> 0
static inline Real distance_cuboid_cuboid(const Iso& pos12, const Vec3& he1, const Vec3& he2) {
    const Real margin = REAL_MAX;
    Iso pos21 = pos12.inverse();
    Real s1, s2, s3; Vec3 d1, d2, d3;
    sat_find_separating_normal_oneway(he1, he2, pos12, s1, d1);
    if (s1 > margin) return 0.0f;
    sat_find_separating_normal_oneway(he2, he1, pos21, s2, d2);
    if (s2 > margin) return 0.0f;
    sat_find_separating_edge_twoway(he1, he2, pos12, s3, d3);
    if (s3 > margin) return 0.0f;
    if (s1 <= 0.0f && s2 <= 0.0f && s3 <= 0.0f) return 0.0f;   // Intersecting
    ShapeRef c1, c2;
    c1.kind = SHAPE_CUBOID; c1.half_extents = he1; c2.kind = SHAPE_CUBOID; c2.half_extents = he2;
    if (s1 >= s2 && s1 >= s3) {
        // SupportMap::support_point(pos12, -dir) = pos12 * local_support_point(pos12^-1 (-dir))
        Vec3 pt2_1 = pos12.transform_point(cuboid_local_support(he2, pos12.inverse_transform_vector(-d1)));
        Vec3 proj; bool inside;
        project_local_point_solid(c1, pt2_1, proj, inside);
        if (norm_squared(proj - pt2_1) > margin * margin) return 0.0f;
        Vec3 p2 = pos21.transform_point(pt2_1);
        return norm(proj - pos12.transform_point(p2));
    }
    if (s2 >= s1 && s2 >= s3) {
        Vec3 pt1_2 = pos21.transform_point(cuboid_local_support(he1, pos21.inverse_transform_vector(-d2)));
        Vec3 proj; bool inside;
        project_local_point_solid(c2, pt1_2, proj, inside);
        if (norm_squared(proj - pt1_2) > margin * margin) return 0.0f;
        Vec3 p1 = pos12.transform_point(pt1_2);
        return norm(p1 - pos12.transform_point(proj));
    }
    Vec3 a1, b1, a2, b2;
    cuboid_local_support_edge_segment(he1, d3, a1, b1);
    cuboid_local_support_edge_segment(he2, pos21.transform_vector(-d3), a2, b2);
    Real s, t;
    segment_segment_params(a1, b1, pos12.transform_point(a2), pos12.transform_point(b2), s, t);
    Vec3 p1 = segment_point_at_param(a1, b1, s), p2 = segment_point_at_param(a2, b2, t);
    Vec3 p2w = pos12.transform_point(p2);
    if (norm_squared(p1 - p2w) <= margin * margin) return norm(p1 - p2w);
    return 0.0f;
}

enum QueryStatus { QUERY_OK = 0, QUERY_UNSUPPORTED = 2, QUERY_NEEDS_HOST = 3 };

// DefaultQueryDispatcher::distance (default_query_dispatcher.rs:177-236) for Ball / Cuboid / ConvexPolyhedron.
static inline int dispatch_distance(const Iso& pos12, const ShapeRef& s1, const ShapeRef& s2, Real& out) {
    out = 0.0f;
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) {  // distance_ball_ball.rs
        Real d2 = norm_squared(pos12.tra), sum = s1.radius + s2.radius;
        out = d2 <= sum * sum ? 0.0f : sqrtf(d2) - sum;
        return QUERY_OK;
    }
    if (s1.kind == SHAPE_BALL || s2.kind == SHAPE_BALL) {  // distance_ball_convex_polyhedron.rs
        bool ball_first = s1.kind == SHAPE_BALL;
        Iso p = ball_first ? pos12.inverse() : pos12;
        const ShapeRef& cv = ball_first ? s2 : s1;
        Real r = ball_first ? s1.radius : s2.radius;
        Vec3 center = p.tra, proj; bool inside;
        project_local_point_solid(cv, center, proj, inside);
        Real d = norm(center - proj) - r;     // na::distance(&proj.point, &center2_1)
        out = d > 0.0f ? d : 0.0f;             // .max(0.0)
        return QUERY_OK;
    }
    if (s1.kind == SHAPE_CUBOID && s2.kind == SHAPE_CUBOID) { out = distance_cuboid_cuboid(pos12, s1.half_extents, s2.half_extents); return QUERY_OK; }
    // distance_support_map_support_map.rs
    SupportShape g1 = s1.support(), g2 = s2.support();
    VoronoiSimplex simplex;
    Vec3 dir;
    if (!try_normalize(-pos12.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    simplex.reset(CSOPoint::from_shapes(pos12, g1, g2, dir));
    GJKResult r = gjk_closest_points(pos12, g1, g2, REAL_MAX, simplex);
    out = r.kind == GJKResult::CLOSEST_POINTS ? norm(r.p2 - r.p1) : 0.0f;
    return QUERY_OK;
}

// DefaultQueryDispatcher::intersection_test (default_query_dispatcher.rs:104-175), same shapes.
static inline int dispatch_intersection_test(const Iso& pos12, const ShapeRef& s1, const ShapeRef& s2, bool& out) {
    out = false;
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) {  // intersection_test_ball_ball.rs
        Real d2 = norm_squared(pos12.tra), sum = s1.radius + s2.radius;
        out = d2 <= sum * sum;
        return QUERY_OK;
    }
    if (s1.kind == SHAPE_CUBOID && s2.kind == SHAPE_CUBOID) { out = intersection_test_cuboid_cuboid(pos12, s1.half_extents, s2.half_extents); return QUERY_OK; }
    if (s1.kind == SHAPE_BALL || s2.kind == SHAPE_BALL) {  // intersection_test_ball_point_query.rs
        bool ball_first = s1.kind == SHAPE_BALL;
        Iso p = ball_first ? pos12.inverse() : pos12;
        const ShapeRef& pq = ball_first ? s2 : s1;
        Real r = ball_first ? s1.radius : s2.radius;
        Vec3 c = p.tra, proj; bool inside;
        project_local_point_solid(pq, c, proj, inside);
        out = inside || norm_squared(c - proj) <= r * r;
        return QUERY_OK;
    }
    // intersection_test_support_map_support_map.rs
    SupportShape g1 = s1.support(), g2 = s2.support();
    VoronoiSimplex simplex;
    Vec3 dir;
    if (!try_normalize(pos12.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    simplex.reset(CSOPoint::from_shapes(pos12, g1, g2, dir));
    GJKResult r = gjk_closest_points(pos12, g1, g2, 0.0f, simplex, false);
    out = r.kind == GJKResult::INTERSECTION;
    return QUERY_OK;
}

// Shape::compute_aabb(pos) for the supported shapes (aabb_ball.rs:8-33, aabb_cuboid.rs:9-16, aabb_convex_polyhedron.rs:8-16)
static inline Aabb shape_compute_aabb(const ShapeRef& s, const Iso& pos) {
    if (s.kind == SHAPE_BALL) { Real r = s.radius; return Aabb(pos.tra + Vec3(-r, -r, -r), pos.tra + Vec3(r, r, r)); }
    if (s.kind == SHAPE_CUBOID) { Vec3 he = pos.absolute_transform_vector(s.half_extents); return Aabb(pos.tra - he, pos.tra + he); }
    Vec3 w0 = pos.transform_point(ld3(s.points));
    Aabb a(w0, w0);
    for (uint32_t k = 1; k < s.num_points; ++k) { Vec3 w = pos.transform_point(ld3(s.points + 3 * k)); a.mins = vinf(a.mins, w); a.maxs = vsup(a.maxs, w); }
    return a;
}

// CompositeShapeRef::contact_with_shape for a TriMesh (contact_composite_shape_shape.rs:14-45): every triangle whose leaf
// AABB intersects shape2's loosened AABB is dispatched, the first strictly smaller dist wins (BVH iteration order).
// Result in the local frames of the mesh / shape2. part = winning triangle.
static inline int contact_trimesh_shape(const Iso& pos12, const TriMesh& mesh, const ShapeRef& shape2, Real prediction, Contact& best,
                                        uint32_t& part, bool min_index_ties = false) {
    Aabb ls = shape_compute_aabb(shape2, pos12);
    ls.mins = ls.mins - Vec3(prediction, prediction, prediction);  // Aabb::loosened (aabb.rs)
    ls.maxs = ls.maxs + Vec3(prediction, prediction, prediction);
    std::vector<uint32_t> parts;
    mesh.bvh.intersect_aabb(ls, parts);
    bool have = false;
    for (uint32_t id : parts) {
        float tri[9];
        const uint32_t* t = &mesh.indices[3 * id];
        for (int k = 0; k < 3; ++k) { tri[3 * k] = mesh.vertices[t[k]].x; tri[3 * k + 1] = mesh.vertices[t[k]].y; tri[3 * k + 2] = mesh.vertices[t[k]].z; }
        ShapeRef s1; s1.kind = SHAPE_TRIANGLE; s1.radius = 0; s1.points = tri; s1.num_points = 3;
        Contact c = Contact();
        if (dispatch_contact(pos12, s1, shape2, prediction, c) != CONTACT_SOME) continue;
        bool replace = !have || c.dist < best.dist || (min_index_ties && c.dist == best.dist && id < part);
        if (replace) { best = c; part = id; have = true; }
    }
    return have ? CONTACT_SOME : CONTACT_NONE;
}

// CompositeShapeRef::distance_to_shape for a TriMesh (distance_composite_shape_shape.rs:13-42): Bvh::find_best with the node cost
// Aabb::distance_to_origin (aabb.rs:556-562) of the node box Minkowski-summed with shape2's box, leaf cost = dispatcher.distance(pose12,
// triangle, shape2) (TriMesh parts have no part pose). distance_composite_shape_shape (:46-60) keeps the distance only.
struct DistLeaf { Real d; Real cost() const { return d; } };
static inline bool distance_trimesh_shape(const Iso& pos12, const TriMesh& mesh, const ShapeRef& shape2, Real& out, uint32_t& part) {
    Aabb ls = shape_compute_aabb(shape2, pos12);
    Vec3 msum_shift = -center(ls.mins, ls.maxs);
    Vec3 msum_margin = (ls.maxs - ls.mins) * 0.5f;
    auto aabb_cost = [&](const BvhNode& node, Real) -> Real {
        Vec3 mins = (node.mins + msum_shift) - msum_margin, maxs = (node.maxs + msum_shift) + msum_margin;
        return norm(vsup(vsup(mins, -maxs), Vec3()));
    };
    auto leaf_cost = [&](uint32_t id, Real, DistLeaf& o) -> bool {
        float tv[9];
        const uint32_t* t = &mesh.indices[3 * id];
        for (int k = 0; k < 3; ++k) { tv[3 * k] = mesh.vertices[t[k]].x; tv[3 * k + 1] = mesh.vertices[t[k]].y; tv[3 * k + 2] = mesh.vertices[t[k]].z; }
        ShapeRef tri; tri.kind = SHAPE_TRIANGLE; tri.radius = 0; tri.points = tv; tri.num_points = 3;
        return dispatch_distance(pos12, tri, shape2, o.d) == QUERY_OK;
    };
    DistLeaf best;
    if (!mesh.bvh.find_best<DistLeaf>(REAL_MAX, aabb_cost, leaf_cost, part, best)) { out = REAL_MAX; part = UINT32_MAX; return false; }
    out = best.d;
    return true;
}

// query::closest_points (closest_points_shape_shape.rs:220-231) -> DefaultQueryDispatcher::closest_points
// (default_query_dispatcher.rs:358-424) for Ball / Cuboid / ConvexPolyhedron: closest_points_ball_ball.rs:7-36,
// closest_points_ball_convex_polyhedron.rs:7-44 (through the contact arms), closest_points_support_map_support_map.rs:8-69
// (GJK only, started toward -pos12.translation). Returns 0 Disjoint, 1 WithinMargin(p1, p2) in world space, 2 Intersecting;
// qstatus CONTACT_NEEDS_TOPOLOGY is no longer produced (the hull arm never needs a feature normal, see contact_convex_polyhedron_ball).
enum ClosestPointsKind { CP_DISJOINT = 0, CP_WITHIN_MARGIN = 1, CP_INTERSECTING = 2 };
static inline int query_closest_points(const Iso& pos1, const ShapeRef& s1, const Iso& pos2, const ShapeRef& s2, Real margin, Vec3& p1, Vec3& p2,
                                       int& qstatus) {
    Iso pos12 = pos1.inv_mul(pos2);
    qstatus = CONTACT_SOME;
    int kind;
    p1 = Vec3(); p2 = Vec3();
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) {
        Real r1 = s1.radius, r2 = s2.radius;
        Vec3 delta = pos12.tra;
        Real distance = norm(delta), sum = r1 + r2;
        if (distance - margin <= sum) {
            if (distance <= sum) kind = CP_INTERSECTING;
            else {
                Vec3 n = normalize(delta);
                p1 = n * r1;
                p2 = pos12.inverse_transform_vector(n) * (-r2);
                kind = CP_WITHIN_MARGIN;
            }
        } else kind = CP_DISJOINT;
    } else if (s1.kind == SHAPE_BALL || s2.kind == SHAPE_BALL) {
        Contact c = Contact();
        int st = s1.kind == SHAPE_BALL ? contact_ball_convex_polyhedron(pos12, s1.radius, s2, margin, c)
                                       : contact_convex_polyhedron_ball(pos12, s1, s2.radius, margin, c);
        if (st == CONTACT_SOME) {
            if (c.dist <= 0.0f) kind = CP_INTERSECTING;
            else { kind = CP_WITHIN_MARGIN; p1 = c.point1; p2 = c.point2; }
        } else if (st == CONTACT_NONE) kind = CP_DISJOINT;
        else { qstatus = st; return CP_DISJOINT; }
    } else {
        SupportShape g1 = s1.support(), g2 = s2.support();
        VoronoiSimplex simplex;
        Vec3 dir;
        if (!try_normalize(-pos12.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
        simplex.reset(CSOPoint::from_shapes(pos12, g1, g2, dir));
        GJKResult r = gjk_closest_points(pos12, g1, g2, margin, simplex);
        if (r.kind == GJKResult::CLOSEST_POINTS) { kind = CP_WITHIN_MARGIN; p1 = r.p1; p2 = pos12.inverse_transform_point(r.p2); }
        else if (r.kind == GJKResult::NO_INTERSECTION) kind = CP_DISJOINT;
        else kind = CP_INTERSECTING;
    }
    if (kind == CP_WITHIN_MARGIN) { p1 = pos1.transform_point(p1); p2 = pos2.transform_point(p2); }   // ClosestPoints::transform_by
    return kind;
}

// CompositeShapeRef::contact_with_shape for a Compound (contact_composite_shape_shape.rs:14-45, shape/compound.rs:113-144,
// map_part_at :181-193): parts whose AABB (part shape at its pose) intersects shape2's loosened AABB are dispatched with
// part_pos1.inv_mul(pose12); the first strictly smaller dist wins and is moved to the compound's frame with
// transform1_by_mut. Parts are visited in index order, i.e. equal dists go to the smallest part index (the reference visits
// them in its binned tree's order); the candidate set is the same as Bvh::intersect_aabb's (leaf test = Aabb::intersects).
struct CompoundRef { const ShapeRef* shapes; const Iso* poses; uint32_t n; };
static inline int contact_compound_shape(const Iso& pos12, const CompoundRef& comp, const ShapeRef& shape2, Real prediction, Contact& best,
                                         uint32_t& part) {
    Aabb ls = shape_compute_aabb(shape2, pos12);
    ls.mins = ls.mins - Vec3(prediction, prediction, prediction);
    ls.maxs = ls.maxs + Vec3(prediction, prediction, prediction);
    bool have = false;
    for (uint32_t i = 0; i < comp.n; ++i) {
        if (!shape_compute_aabb(comp.shapes[i], comp.poses[i]).intersects(ls)) continue;
        Contact c = Contact();
        if (dispatch_contact(comp.poses[i].inv_mul(pos12), comp.shapes[i], shape2, prediction, c) != CONTACT_SOME) continue;
        if (!have || c.dist < best.dist) {
            c.point1 = comp.poses[i].transform_point(c.point1);   // transform1_by_mut(part_pos1)
            c.normal1 = comp.poses[i].transform_vector(c.normal1);
            best = c; part = i; have = true;
        }
    }
    return have ? CONTACT_SOME : CONTACT_NONE;
}
// Compound vs Compound through the same dispatcher arms (default_query_dispatcher.rs:338-351): shape1 composite =>
// contact_composite_shape_shape(pos12, c1, shape2 = the other compound): every part of c1 whose AABB meets the loosened AABB of
// compound 2 (Shape::compute_aabb default = local_aabb.transform_by(pos), aabb.rs:492-498; Compound::local_aabb = merge of the
// part AABBs, compound.rs:120-126) is dispatched against the whole compound 2, which lands in contact_shape_composite_shape
// (:63-76): pose.inverse(), compound 2 as the composite, the part as the shape, flipped(). Oracle groundwork (no GPU path yet).
static inline Aabb compound_local_aabb(const CompoundRef& c) {
    Aabb a(Vec3(REAL_MAX, REAL_MAX, REAL_MAX), Vec3(-REAL_MAX, -REAL_MAX, -REAL_MAX));
    for (uint32_t i = 0; i < c.n; ++i) { Aabb b = shape_compute_aabb(c.shapes[i], c.poses[i]); a.mins = vinf(a.mins, b.mins); a.maxs = vsup(a.maxs, b.maxs); }
    return a;
}
static inline Aabb aabb_transform_by(const Aabb& a, const Iso& m) {
    Vec3 c = m.transform_point(center(a.mins, a.maxs));
    Vec3 he = m.absolute_transform_vector((a.maxs - a.mins) * 0.5f);
    return Aabb(c + (-he), c + he);
}
static inline int contact_compound_compound(const Iso& pos12, const CompoundRef& c1, const CompoundRef& c2, Real prediction, Contact& best,
                                            uint32_t& part1, uint32_t& part2) {
    Aabb ls = aabb_transform_by(compound_local_aabb(c2), pos12);
    ls.mins = ls.mins - Vec3(prediction, prediction, prediction);
    ls.maxs = ls.maxs + Vec3(prediction, prediction, prediction);
    bool have = false;
    for (uint32_t i = 0; i < c1.n; ++i) {
        if (!shape_compute_aabb(c1.shapes[i], c1.poses[i]).intersects(ls)) continue;
        Iso pos_i2 = c1.poses[i].inv_mul(pos12);          // pose of compound 2 in part i's frame
        Contact c = Contact(); uint32_t j = UINT32_MAX;
        // dispatcher.contact(pos_i2, part_i, compound2) -> contact_shape_composite_shape: inverse pose, flipped result
        if (contact_compound_shape(pos_i2.inverse(), c2, c1.shapes[i], prediction, c, j) != CONTACT_SOME) continue;
        std::swap(c.point1, c.point2); std::swap(c.normal1, c.normal2);
        if (!have || c.dist < best.dist) {
            c.point1 = c1.poses[i].transform_point(c.point1);
            c.normal1 = c1.poses[i].transform_vector(c.normal1);
            best = c; part1 = i; part2 = j; have = true;
        }
    }
    return have ? CONTACT_SOME : CONTACT_NONE;
}

// Compound vs TriMesh, both orders, through the same dispatcher arms (default_query_dispatcher.rs:338-351 takes shape1's composite
// view first). Oracle groundwork (no GPU path yet).
// Bvh::root_aabb (bvh_tree.rs:1991-2000) = TriMesh::local_aabb (trimesh.rs:1793-1795); TriMesh::compute_aabb(pos) = root_aabb().transform_by(pos)
static inline Aabb bvh_root_aabb(const Bvh& b) {
    if (b.nodes.empty()) return Aabb(Vec3(REAL_MAX, REAL_MAX, REAL_MAX), Vec3(-REAL_MAX, -REAL_MAX, -REAL_MAX));
    uint32_t lc = b.nodes[0].leaf_count();
    if (lc == 0) return Aabb(Vec3(REAL_MAX, REAL_MAX, REAL_MAX), Vec3(-REAL_MAX, -REAL_MAX, -REAL_MAX));
    if (lc == 1) return b.nodes[0].left.aabb();
    Aabb l = b.nodes[0].left.aabb(), r = b.nodes[0].right.aabb();
    return Aabb(vinf(l.mins, r.mins), vsup(l.maxs, r.maxs));
}
// contact_composite_shape_shape(pos12, compound1, trimesh2): every part i of the compound whose AABB meets the mesh's loosened AABB
// is dispatched as contact(part_pos1[i].inv_mul(pos12), part_i, trimesh), i.e. contact_shape_composite_shape (:63-76): the mesh as
// the composite under the inverse pose, flipped(); then transform1_by_mut(part_pos1[i]). part / tri = the winning part and triangle.
static inline int contact_compound_trimesh(const Iso& pos12, const CompoundRef& c1, const TriMesh& mesh, Real prediction, Contact& best,
                                           uint32_t& part, uint32_t& tri, bool min_index_ties = false) {
    Aabb ls = aabb_transform_by(bvh_root_aabb(mesh.bvh), pos12);
    ls.mins = ls.mins - Vec3(prediction, prediction, prediction);
    ls.maxs = ls.maxs + Vec3(prediction, prediction, prediction);
    bool have = false;
    for (uint32_t i = 0; i < c1.n; ++i) {
        if (!shape_compute_aabb(c1.shapes[i], c1.poses[i]).intersects(ls)) continue;
        Iso pos_i2 = c1.poses[i].inv_mul(pos12);          // pose of the mesh in part i's frame
        Contact c = Contact(); uint32_t t = UINT32_MAX;
        if (contact_trimesh_shape(pos_i2.inverse(), mesh, c1.shapes[i], prediction, c, t, min_index_ties) != CONTACT_SOME) continue;
        std::swap(c.point1, c.point2); std::swap(c.normal1, c.normal2);
        if (!have || c.dist < best.dist) {
            c.point1 = c1.poses[i].transform_point(c.point1);
            c.normal1 = c1.poses[i].transform_vector(c.normal1);
            best = c; part = i; tri = t; have = true;
        }
    }
    return have ? CONTACT_SOME : CONTACT_NONE;
}
// contact_composite_shape_shape(pos12, trimesh1, compound2): every triangle whose leaf AABB meets the compound's loosened AABB
// (Shape::compute_aabb default on Compound::local_aabb) is dispatched as contact(pos12, triangle, compound2) (TriMesh parts have
// no part pose), i.e. contact_shape_composite_shape: the compound as the composite under pos12.inverse(), flipped().
static inline int contact_trimesh_compound(const Iso& pos12, const TriMesh& mesh, const CompoundRef& c2, Real prediction, Contact& best,
                                           uint32_t& tri, uint32_t& part, bool min_index_ties = false) {
    Aabb ls = aabb_transform_by(compound_local_aabb(c2), pos12);
    ls.mins = ls.mins - Vec3(prediction, prediction, prediction);
    ls.maxs = ls.maxs + Vec3(prediction, prediction, prediction);
    std::vector<uint32_t> ids;
    mesh.bvh.intersect_aabb(ls, ids);
    Iso pos21 = pos12.inverse();
    bool have = false;
    for (uint32_t id : ids) {
        float tv[9];
        const uint32_t* t = &mesh.indices[3 * id];
        for (int k = 0; k < 3; ++k) { tv[3 * k] = mesh.vertices[t[k]].x; tv[3 * k + 1] = mesh.vertices[t[k]].y; tv[3 * k + 2] = mesh.vertices[t[k]].z; }
        ShapeRef s1; s1.kind = SHAPE_TRIANGLE; s1.radius = 0; s1.points = tv; s1.num_points = 3;
        Contact c = Contact(); uint32_t j = UINT32_MAX;
        if (contact_compound_shape(pos21, c2, s1, prediction, c, j) != CONTACT_SOME) continue;
        std::swap(c.point1, c.point2); std::swap(c.normal1, c.normal2);
        bool replace = !have || c.dist < best.dist || (min_index_ties && c.dist == best.dist && id < tri);
        if (replace) { best = c; tri = id; part = j; have = true; }
    }
    return have ? CONTACT_SOME : CONTACT_NONE;
}

// query::contact with a Compound on one side (default_query_dispatcher.rs:338-351): compound first = composite arm; compound
// second (flipped) = contact_shape_composite_shape (contact_composite_shape_shape.rs:63-76): pose12.inverse(), then flipped().
static inline int query_contact_compound(const Iso& pos1, const Iso& pos2, const CompoundRef& comp, const ShapeRef& shape, bool compound_second,
                                         Real prediction, Contact& c, uint32_t& part) {
    Iso pos12 = pos1.inv_mul(pos2);
    int st = contact_compound_shape(compound_second ? pos12.inverse() : pos12, comp, shape, prediction, c, part);
    if (st != CONTACT_SOME) return st;
    if (compound_second) { std::swap(c.point1, c.point2); std::swap(c.normal1, c.normal2); }
    c.point1 = pos1.transform_point(c.point1);
    c.point2 = pos2.transform_point(c.point2);
    c.normal1 = pos1.transform_vector(c.normal1);
    c.normal2 = pos2.transform_vector(c.normal2);
    return st;
}

// PointQuery for TriMesh without pseudo-normals (point_composite_shape.rs:164-186 -> CompositeShapeRef::project_local_point,
// :49-72): Bvh::find_best with aabb cost = Aabb::distance_to_local_point(pt, solid = true) (point_aabb.rs:135-146: the norm of
// the per-axis shift) and leaf cost = distance to the projection on the triangle (point_triangle.rs).
struct ProjCost { Real d; Vec3 point; bool inside; Real cost() const { return d; } };
static inline bool trimesh_project_local_point(const TriMesh& mesh, const Vec3& pt, bool solid, bool min_index_ties, uint32_t& id, Vec3& proj, bool& inside) {
    ProjCost best;
    auto aabb_cost = [&](const BvhNode& n, Real) {
        Vec3 zero;
        Vec3 shift = vsup(vsup(n.aabb().mins - pt, pt - n.aabb().maxs), zero);
        return norm(shift);
    };
    auto leaf = [&](uint32_t prim, Real, ProjCost& out) {
        const uint32_t* t = &mesh.indices[3 * prim];
        TriProj p = project_on_triangle(mesh.vertices[t[0]], mesh.vertices[t[1]], mesh.vertices[t[2]], pt, solid);
        out.point = p.point; out.inside = p.inside; out.d = norm(pt - p.point);
        return true;
    };
    if (!min_index_ties) {
        if (!mesh.bvh.find_best<ProjCost>(REAL_MAX, aabb_cost, leaf, id, best)) return false;
    } else {  // brute force over all triangles, equal distances resolve to the smallest index
        bool have = false;
        for (uint32_t k = 0; k < (uint32_t)mesh.num_triangles(); ++k) {
            ProjCost c; leaf(k, 0, c);
            if (!have || c.d < best.d) { best = c; id = k; have = true; }
        }
        if (!have) return false;
    }
    proj = best.point; inside = best.inside;
    return true;
}

}  // namespace pb2o

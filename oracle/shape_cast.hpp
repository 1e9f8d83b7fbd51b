// ORACLE — TEST INFRASTRUCTURE ONLY. Restatement of parry3d src/query/shape_cast/{shape_cast.rs:14-286,
// shape_cast_ball_ball.rs:10-69, shape_cast_support_map_support_map.rs:11-69} and DefaultQueryDispatcher::cast_shapes
// (default_query_dispatcher.rs:434-515) for Ball / Cuboid / ConvexPolyhedron. Pinned by the reference's own tests
// (crates/parry3d/tests/geometry/{ball_ball_toi,time_of_impact3,still_objects_toi}.rs) in tests/test_oracle_kats.py.
#pragma once
#include "contact.hpp"
#include "ray_support_map.hpp"

namespace pb2o {

enum ShapeCastStatus { CAST_OUT_OF_ITERATIONS = 0, CAST_CONVERGED = 1, CAST_FAILED = 2, CAST_PENETRATING = 3 };
struct ShapeCastOptions {
    Real max_time_of_impact = REAL_MAX;
    Real target_distance = 0.0f;
    bool stop_at_penetration = true;
    bool compute_impact_geometry_on_penetration = true;
};
struct ShapeCastHit { Real time_of_impact; Vec3 witness1, witness2, normal1, normal2; int status; };

// shape_cast_ball_ball.rs:10-69
static inline bool cast_shapes_ball_ball(const Iso& pos12, const Vec3& vel12, Real r1, Real r2, const ShapeCastOptions& o, ShapeCastHit& hit) {
    Real rsum = r1 + r2 + o.target_distance;
    Real radius = rsum;
    Vec3 center = -pos12.tra;
    Ray ray(Vec3(), vel12);
    bool inside; Real toi;
    if (!ray_toi_with_ball(center, radius, ray, true, inside, toi)) return false;
    if (toi > o.max_time_of_impact) return false;
    Vec3 dpt = ray.point_at(toi) - center;
    Vec3 n1, n2, w1, w2;
    if (radius == 0.0f) {
        n1 = Vec3(1, 0, 0);
        n2 = pos12.inverse_transform_vector(-Vec3(1, 0, 0));
        w1 = Vec3(); w2 = Vec3();
    } else {
        n1 = dpt / radius;
        n2 = pos12.inverse_transform_vector(-n1);
        w1 = n1 * r1;
        w2 = n2 * r2;
    }
    if (!o.stop_at_penetration && toi < 1.0e-5f && dot(n1, vel12) >= 0.0f) return false;
    hit.time_of_impact = toi; hit.normal1 = n1; hit.normal2 = n2; hit.witness1 = w1; hit.witness2 = w2;
    hit.status = (inside && norm_squared(center) < rsum * rsum) ? CAST_PENETRATING : CAST_CONVERGED;
    return true;
}

// shape_cast_support_map_support_map.rs:11-69
static inline bool cast_shapes_support_map_support_map(const Iso& pos12, const Vec3& vel12, const SupportShape& g1, const SupportShape& g2,
                                                       const ShapeCastOptions& o, ShapeCastHit& hit) {
    VoronoiSimplex simplex;
    Real toi; Vec3 normal1, w1, w2;
    bool some = o.target_distance > 0.0f ? gjk_directional_distance(pos12, g1.rounded(o.target_distance), g2, vel12, simplex, toi, normal1, w1, w2)
                                         : gjk_directional_distance(pos12, g1, g2, vel12, simplex, toi, normal1, w1, w2);
    if (!some) return false;
    if (toi > o.max_time_of_impact) return false;
    if ((o.compute_impact_geometry_on_penetration || !o.stop_at_penetration) && toi < 1.0e-5f) {
        Contact c;
        if (contact_support_map_support_map(pos12, g1, g2, REAL_MAX, c) != CONTACT_SOME) return false;
        Real normal_vel = dot(c.normal1, vel12);
        if (!o.stop_at_penetration && normal_vel >= 0.0f) return false;
        hit.time_of_impact = toi; hit.normal1 = c.normal1; hit.normal2 = c.normal2; hit.witness1 = c.point1; hit.witness2 = c.point2;
        hit.status = CAST_PENETRATING;
        return true;
    }
    hit.time_of_impact = toi;
    hit.normal1 = normal1;
    hit.normal2 = pos12.inverse_transform_vector(-normal1);
    hit.witness1 = w1 - normal1 * o.target_distance;
    hit.witness2 = pos12.inverse_transform_point(w2);
    hit.status = toi == 0.0f ? CAST_PENETRATING : CAST_CONVERGED;
    return true;
}

static inline SupportShape support_of(const ShapeRef& s) { return s.kind == SHAPE_BALL ? SupportShape::ball(s.radius) : s.support(); }

// DefaultQueryDispatcher::cast_shapes (default_query_dispatcher.rs:434-515), shapes restricted to Ball / Cuboid / ConvexPolyhedron
static inline bool dispatch_cast_shapes(const Iso& pos12, const Vec3& vel12, const ShapeRef& s1, const ShapeRef& s2, const ShapeCastOptions& o,
                                        ShapeCastHit& hit) {
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) return cast_shapes_ball_ball(pos12, vel12, s1.radius, s2.radius, o, hit);
    return cast_shapes_support_map_support_map(pos12, vel12, support_of(s1), support_of(s2), o, hit);
}

// query::cast_shapes (shape_cast.rs:268-286)
static inline bool cast_shapes(const Iso& pos1, const Vec3& vel1, const ShapeRef& s1, const Iso& pos2, const Vec3& vel2, const ShapeRef& s2,
                               const ShapeCastOptions& o, ShapeCastHit& hit) {
    Iso pos12 = pos1.inv_mul(pos2);
    Vec3 vel12 = pos1.inverse_transform_vector(vel2 - vel1);
    return dispatch_cast_shapes(pos12, vel12, s1, s2, o, hit);
}

// ---- shape casts with a TriMesh on either side (oracle groundwork, no GPU path yet): CompositeShapeRef::cast_shape
// (shape_cast_composite_shape_shape.rs:14-62), cast_shapes_composite_shape_shape (:65-83), cast_shapes_shape_composite_shape (:86-105)
// and the composite arms of DefaultQueryDispatcher::cast_shapes (default_query_dispatcher.rs:498-515). Pinned by the reference's
// crates/parry3d/tests/geometry/trimesh_trimesh_toi.rs (issue #194, exact 0.00998).
struct CastShape { const ShapeRef* shape; const TriMesh* mesh; };   // exactly one of the two is set
struct CastLeaf { ShapeCastHit hit; Real cost() const { return hit.time_of_impact; } };   // BvhLeafCost for (u32, ShapeCastHit)
static inline ShapeCastHit cast_hit_swapped(const ShapeCastHit& h) {   // ShapeCastHit::swapped (shape_cast.rs:71-80)
    ShapeCastHit r = h;
    r.witness1 = h.witness2; r.witness2 = h.witness1; r.normal1 = h.normal2; r.normal2 = h.normal1;
    return r;
}
static inline bool dispatch_cast_shapes_any(const Iso& pos12, const Vec3& vel12, const CastShape& a, const CastShape& b, const ShapeCastOptions& o,
                                            ShapeCastHit& hit, uint32_t* part1 = nullptr);
static inline bool trimesh_cast_shape(const TriMesh& mesh, const Iso& pose12, const Vec3& vel12, const CastShape& g2, const ShapeCastOptions& o,
                                      uint32_t& part, ShapeCastHit& hit) {
    // g2.compute_aabb(pose12): TriMesh = root_aabb().transform_by (trimesh.rs:1763-1765); Triangle / hull = transformed points; ...
    Aabb ls = g2.mesh ? aabb_transform_by(bvh_root_aabb(g2.mesh->bvh), pose12) : shape_compute_aabb(*g2.shape, pose12);
    Ray ray(Vec3(), vel12);
    Vec3 msum_shift = -center(ls.mins, ls.maxs);
    Vec3 msum_margin = (ls.maxs - ls.mins) * 0.5f + Vec3(o.target_distance, o.target_distance, o.target_distance);
    CastLeaf best;
    auto aabb_cost = [&](const BvhNode& node, Real best_so_far) -> Real {
        Aabb msum((node.mins + msum_shift) - msum_margin, (node.maxs + msum_shift) + msum_margin);   // Minkowski sum of the two boxes
        Real t;
        return aabb_cast_local_ray(msum, ray, best_so_far, true, t) ? t : REAL_MAX;
    };
    auto leaf_cost = [&](uint32_t id, Real, CastLeaf& out) -> bool {
        float tv[9];
        const uint32_t* t = &mesh.indices[3 * id];
        for (int k = 0; k < 3; ++k) { tv[3 * k] = mesh.vertices[t[k]].x; tv[3 * k + 1] = mesh.vertices[t[k]].y; tv[3 * k + 2] = mesh.vertices[t[k]].z; }
        ShapeRef tri; tri.kind = SHAPE_TRIANGLE; tri.radius = 0; tri.points = tv; tri.num_points = 3;
        CastShape part1{&tri, nullptr};
        return dispatch_cast_shapes_any(pose12, vel12, part1, g2, o, out.hit);   // TriMesh parts carry no part pose
    };
    if (!mesh.bvh.find_best<CastLeaf>(o.max_time_of_impact, aabb_cost, leaf_cost, part, best)) return false;
    hit = best.hit;
    return true;
}
static inline bool dispatch_cast_shapes_any(const Iso& pos12, const Vec3& vel12, const CastShape& a, const CastShape& b, const ShapeCastOptions& o,
                                            ShapeCastHit& hit, uint32_t* part1) {
    if (!a.mesh && !b.mesh) return dispatch_cast_shapes(pos12, vel12, *a.shape, *b.shape, o, hit);
    uint32_t part = UINT32_MAX;
    if (a.mesh) {   // cast_shapes_composite_shape_shape
        bool some = trimesh_cast_shape(*a.mesh, pos12, vel12, b, o, part, hit);
        if (part1) *part1 = some ? part : UINT32_MAX;
        return some;
    }
    // cast_shapes_shape_composite_shape: the mesh as shape 1 under the inverse pose and velocity, then swapped()
    Iso pos21 = pos12.inverse();
    ShapeCastHit h;
    if (!trimesh_cast_shape(*b.mesh, pos21, -pos12.inverse_transform_vector(vel12), a, o, part, h)) { if (part1) *part1 = UINT32_MAX; return false; }
    hit = cast_hit_swapped(h);
    if (part1) *part1 = part;   // (test bookkeeping: the triangle of the mesh that was hit, whichever side the mesh is on)
    return true;
}
// query::cast_shapes (shape_cast.rs:268-286) with a TriMesh on either side
static inline bool cast_shapes_any(const Iso& pos1, const Vec3& vel1, const CastShape& a, const Iso& pos2, const Vec3& vel2, const CastShape& b,
                                   const ShapeCastOptions& o, ShapeCastHit& hit, uint32_t* part1 = nullptr) {
    Iso pos12 = pos1.inv_mul(pos2);
    Vec3 vel12 = pos1.inverse_transform_vector(vel2 - vel1);
    return dispatch_cast_shapes_any(pos12, vel12, a, b, o, hit, part1);
}

}  // namespace pb2o

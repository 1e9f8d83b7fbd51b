// ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned: the reference holds no known-answer test for ray casts on a
// ConvexPolyhedron; tests/test_oracle_kat.py pins this restatement with cross-checks against the closed-form cuboid cast).
// Restatement of parry3d src/query/ray/ray_support_map.rs:19-72 (local_ray_intersection_with_support_map_with_params,
// RayCast for ConvexPolyhedron :163-181), src/query/gjk/gjk.rs:519-534 (cast_local_ray) and :660-795 (minkowski_ray_cast),
// src/query/ray/ray_halfspace.rs:9-39.
#pragma once
#include "gjk.hpp"
#include "ray.hpp"

namespace pb2o {

// ray_halfspace.rs:9-39: None <=> denominator ~ 0 or t < 0
static inline bool ray_toi_with_halfspace(const Vec3& center, const Vec3& normal, const Ray& ray, Real& t) {
    Vec3 dpos = center - ray.origin;
    Real denom = dot(normal, ray.dir);
    if (relative_eq(denom, 0.0f)) return false;
    t = dot(normal, dpos) / denom;
    return t >= 0.0f;
}

// minkowski_ray_cast (gjk.rs:660-795): ray cast on the Minkowski difference g1 - pos12 * g2
static inline bool minkowski_ray_cast(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, const Ray& ray, Real max_toi,
                                      VoronoiSimplex& simplex, Real& toi, Vec3& normal) {
    const Real eps_tol = gjk_eps_tol();
    const Real eps_rel = sqrtf(eps_tol);
    Real ray_length = norm(ray.dir);
    if (relative_eq(ray_length, 0.0f)) return false;
    Real ltoi = 0.0f;
    Ray curr_ray(ray.origin, ray.dir / ray_length);
    Vec3 dir0 = -curr_ray.dir;
    Vec3 ldir = dir0;
    CSOPoint sp0 = CSOPoint::from_shapes(pos12, g1, g2, dir0);
    sp0.point = sp0.point + (-curr_ray.origin);  // translate(&-origin)
    simplex.reset(sp0);
    Vec3 proj = simplex.project_origin_and_reduce();
    Real max_bound = REAL_MAX;
    Vec3 dir;
    int niter = 0;
    bool last_chance = false;
    for (;;) {
        Real old_max_bound = max_bound;
        Real dist;
        if (try_normalize_and_get(-proj, eps_tol, dir, dist)) max_bound = dist;
        else { toi = ltoi / ray_length; normal = ldir; return true; }
        CSOPoint support_point;
        if (max_bound >= old_max_bound) {
            last_chance = true;
            Vec3 p = proj + curr_ray.origin;
            support_point.point = p; support_point.orig1 = p; support_point.orig2 = Vec3();  // single_point
        } else {
            support_point = CSOPoint::from_shapes(pos12, g1, g2, dir);
        }
        if (last_chance && ltoi > 0.0f) { toi = ltoi / ray_length; normal = ldir; return true; }
        Real t;
        if (ray_toi_with_halfspace(support_point.point, dir, curr_ray, t)) {
            if (dot(dir, curr_ray.dir) < 0.0f && t > 0.0f) {
                ldir = dir;
                ltoi += t;
                if (ltoi / ray_length > max_toi) return false;
                Vec3 shift = curr_ray.dir * t;
                curr_ray.origin = curr_ray.origin + shift;
                max_bound = REAL_MAX;
                for (size_t i = 0; i <= simplex.dim; ++i) simplex.vertices[i].point = simplex.vertices[i].point + (-shift);  // modify_pnts
                last_chance = false;
            }
        } else if (dot(dir, curr_ray.dir) > eps_tol) {
            return false;
        }
        if (last_chance) return false;
        Real min_bound = -dot(dir, support_point.point - curr_ray.origin);
        assert(std::isfinite(min_bound));
        if (max_bound - min_bound <= eps_rel * max_bound) return false;  // not "improved_fixed_point_support"
        CSOPoint tr = support_point;
        tr.point = tr.point + (-curr_ray.origin);
        (void)simplex.add_point(tr);
        proj = simplex.project_origin_and_reduce();
        if (simplex.dimension() == 3) {
            if (min_bound >= eps_tol) return false;
            toi = ltoi / ray_length; normal = ldir; return true;
        }
        niter += 1;
        if (niter == 100) return false;
    }
}

// gjk::cast_local_ray (gjk.rs:519-534): g2 = ConstantOrigin, pos12 = identity
static inline bool gjk_cast_local_ray(const SupportShape& shape, VoronoiSimplex& simplex, const Ray& ray, Real max_toi, Real& toi, Vec3& normal) {
    return minkowski_ray_cast(Iso(), shape, SupportShape::constant_origin(), ray, max_toi, simplex, toi, normal);
}

// gjk::directional_distance (gjk.rs:632-657)
static inline bool gjk_directional_distance(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, const Vec3& dir, VoronoiSimplex& simplex,
                                            Real& toi, Vec3& normal, Vec3& w1, Vec3& w2) {
    Ray ray(Vec3(), dir);
    if (!minkowski_ray_cast(pos12, g1, g2, ray, REAL_MAX, simplex, toi, normal)) return false;
    if (toi != 0.0f) gjk_result(simplex, simplex.dimension() == 3, w1, w2);
    else { w1 = Vec3(); w2 = Vec3(); }  // penetration: witness points undefined
    return true;
}

#define PB2O_FEATURE_UNKNOWN 0xFFFFFFFEu  // FeatureId::Unknown in the u32 feature column

// ray_support_map.rs:19-72
static inline bool support_map_cast_local_ray_and_get_normal(const SupportShape& shape, const Ray& ray, Real max_toi, bool solid, RayIntersection& out) {
    VoronoiSimplex simplex;
    Real toi; Vec3 normal;
    if (!gjk_cast_local_ray(shape, simplex, ray, max_toi, toi, normal)) return false;
    out.feature = PB2O_FEATURE_UNKNOWN;
    if (!solid && toi == 0.0f) {
        // the ray starts inside: cast it backwards from beyond the shape
        Vec3 ndir = normalize(ray.dir);
        Vec3 supp = shape.local_support_point(ndir);
        const Real eps = 0.001f;
        Real shift = dot(supp - ray.origin, ndir) + eps;
        Ray new_ray(ray.origin + ndir * shift, -ray.dir);
        Real t2; Vec3 n2;
        if (!gjk_cast_local_ray(shape, simplex, new_ray, shift + eps, t2, n2)) return false;
        Real t = shift - t2;
        if (!(t <= max_toi)) return false;
        out.time_of_impact = t; out.normal = -n2;
        return true;
    }
    out.time_of_impact = toi; out.normal = normal;
    return true;
}

}  // namespace pb2o

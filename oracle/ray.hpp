// ORACLE — TEST INFRASTRUCTURE ONLY (see math.hpp header).
// Restatement of parry3d src/query/ray/{ray_aabb,ray_triangle,ray_ball,ray_cuboid,ray_composite_shape,
// ray_trimesh}.rs and src/query/clip/clip_aabb_line.rs.
#pragma once
#include "bvh.hpp"

namespace pb2o {

// shape/feature_id.rs: we only ever produce Face(u32) / Unknown on this path.
struct RayIntersection {
    Real time_of_impact;
    Vec3 normal;
    uint32_t feature;  // FeatureId::Face(feature)
    Real cost() const { return time_of_impact; }  // BvhLeafCost for RayIntersection
};
struct RealCost { Real v; Real cost() const { return v; } };

// ray_aabb.rs:12-49
static inline bool aabb_cast_local_ray(const Aabb& a, const Ray& ray, Real max_toi, bool solid, Real& out) {
    Real tmin = 0.0f, tmax = max_toi;
    for (int i = 0; i < 3; ++i) {
        if (ray.dir[i] == 0.0f) {
            if (ray.origin[i] < a.mins[i] || ray.origin[i] > a.maxs[i]) return false;
        } else {
            Real denom = 1.0f / ray.dir[i];
            Real near = (a.mins[i] - ray.origin[i]) * denom;
            Real far = (a.maxs[i] - ray.origin[i]) * denom;
            if (near > far) std::swap(near, far);
            tmin = rmax(tmin, near);
            tmax = rmin(tmax, far);
            if (tmin > tmax) return false;
        }
    }
    out = (tmin == 0.0f && !solid) ? tmax : tmin;
    return true;
}
// bvh_tree.rs:1177-1181
static inline Real node_cast_ray(const BvhNode& n, const Ray& ray, Real max_toi) {
    Real t;
    return aabb_cast_local_ray(n.aabb(), ray, max_toi, true, t) ? t : REAL_MAX;
}

// ray_triangle.rs:70-152 (returns false for None; bary not needed by callers on this path)
static inline bool local_ray_intersection_with_triangle(const Vec3& a, const Vec3& b, const Vec3& c, const Ray& ray, RayIntersection& out) {
    Vec3 ab = b - a, ac = c - a;
    Vec3 n = cross(ab, ac);
    Real d = dot(n, ray.dir);
    if (d == 0.0f) return false;
    Vec3 ap = ray.origin - a;
    Real t = dot(ap, n);
    if ((t < 0.0f && d < 0.0f) || (t > 0.0f && d > 0.0f)) return false;
    uint32_t fid = d < 0.0f ? 0 : 1;
    d = fabsf(d);
    Vec3 e = -cross(ray.dir, ap);
    Real v, w, toi; Vec3 normal;
    if (t < 0.0f) {
        v = -dot(ac, e);
        if (v < 0.0f || v > d) return false;
        w = dot(ab, e);
        if (w < 0.0f || v + w > d) return false;
        Real invd = 1.0f / d;
        toi = -t * invd;
        normal = -normalize(n);
    } else {
        v = dot(ac, e);
        if (v < 0.0f || v > d) return false;
        w = -dot(ab, e);
        if (w < 0.0f || v + w > d) return false;
        Real invd = 1.0f / d;
        toi = t * invd;
        normal = normalize(n);
    }
    out.time_of_impact = toi; out.normal = normal; out.feature = fid;
    return true;
}
// ray_triangle.rs:49-62 (solid ignored in 3D)
static inline bool triangle_cast_local_ray_and_get_normal(const Vec3& a, const Vec3& b, const Vec3& c, const Ray& ray, Real max_toi, RayIntersection& out) {
    if (!local_ray_intersection_with_triangle(a, b, c, ray, out)) return false;
    return out.time_of_impact <= max_toi;
}

// ray_ball.rs:33-77
static inline bool ray_toi_with_ball(const Vec3& center, Real radius, const Ray& ray, bool solid, bool& inside, Real& toi) {
    Vec3 dcenter = ray.origin - center;
    Real a = norm_squared(ray.dir);
    Real b = dot(dcenter, ray.dir);
    Real c = norm_squared(dcenter) - radius * radius;
    if (a == 0.0f) {
        if (c > 0.0f) { inside = false; return false; }
        inside = true; toi = 0.0f; return true;
    }
    if (c > 0.0f && b > 0.0f) { inside = false; return false; }
    Real delta = b * b - a * c;
    if (delta < 0.0f) { inside = false; return false; }
    Real t = (-b - sqrtf(delta)) / a;
    if (t <= 0.0f) {
        inside = true;
        toi = solid ? 0.0f : (-b + sqrtf(delta)) / a;
        return true;
    }
    inside = false; toi = t; return true;
}
// ray_ball.rs:8-27, 81-98 (ball centred at the origin of its local frame)
static inline bool ball_cast_local_ray(Real radius, const Ray& ray, Real max_toi, bool solid, Real& toi) {
    bool inside;
    if (!ray_toi_with_ball(Vec3(), radius, ray, solid, inside, toi)) return false;
    return toi <= max_toi;
}
static inline bool ball_cast_local_ray_and_get_normal(Real radius, const Ray& ray, Real max_toi, bool solid, RayIntersection& out) {
    bool inside; Real n;
    if (!ray_toi_with_ball(Vec3(), radius, ray, solid, inside, n)) return false;
    Vec3 pos = (ray.origin + ray.dir * n) - Vec3();
    Vec3 normal = normalize(pos);
    out.time_of_impact = n; out.normal = inside ? -normal : normal; out.feature = 0;
    return out.time_of_impact <= max_toi;
}

// clip_aabb_line.rs:79-187
struct ClipHit { Real t; Vec3 n; int side; };
static inline bool clip_aabb_line(const Aabb& aabb, const Vec3& origin, const Vec3& dir, ClipHit& near, ClipHit& far) {
    Real tmax = REAL_MAX, tmin = -tmax;
    int near_side = 0, far_side = 0;
    bool near_diag = false, far_diag = false;
    for (int i = 0; i < 3; ++i) {
        if (dir[i] == 0.0f) {
            if (origin[i] < aabb.mins[i] || origin[i] > aabb.maxs[i]) return false;
        } else {
            Real denom = 1.0f / dir[i];
            bool flip;
            Real inear = (aabb.mins[i] - origin[i]) * denom;
            Real ifar = (aabb.maxs[i] - origin[i]) * denom;
            if (inear > ifar) { flip = true; std::swap(inear, ifar); } else flip = false;
            if (inear > tmin) { tmin = inear; near_side = flip ? -(i + 1) : (i + 1); near_diag = false; }
            else if (inear == tmin) near_diag = true;
            if (ifar < tmax) { tmax = ifar; far_side = !flip ? -(i + 1) : (i + 1); far_diag = false; }
            else if (ifar == tmax) far_diag = true;
            if (tmax < 0.0f || tmin > tmax) return false;
        }
    }
    ClipHit zero{0.0f, Vec3(), 0};
    if (near_diag) near = ClipHit{tmin, -normalize(dir), near_side};
    else {
        if (near_side == 0) { near = far = zero; return aabb.contains_local_point(origin); }
        Vec3 n;
        if (near_side < 0) n[-near_side - 1] = 1.0f; else n[near_side - 1] = -1.0f;
        near = ClipHit{tmin, n, near_side};
    }
    if (far_diag) far = ClipHit{tmax, -normalize(dir), far_side};
    else {
        if (far_side == 0) { near = far = zero; return aabb.contains_local_point(origin); }
        Vec3 n;
        if (far_side < 0) n[-far_side - 1] = -1.0f; else n[far_side - 1] = 1.0f;
        far = ClipHit{tmax, n, far_side};
    }
    return true;
}
// ray_aabb.rs:52-92
static inline bool aabb_cast_local_ray_and_get_normal(const Aabb& aabb, const Ray& ray, Real max_toi, bool solid, RayIntersection& out) {
    ClipHit near, far, r;
    if (!clip_aabb_line(aabb, ray.origin, ray.dir, near, far)) return false;
    if (near.t < 0.0f) {
        if (solid) r = ClipHit{0.0f, Vec3(), far.side};
        else if (far.t <= max_toi) r = far;
        else return false;
    } else if (near.t <= max_toi) r = near;
    else return false;
    out.time_of_impact = r.t; out.normal = r.n;
    out.feature = r.side < 0 ? (uint32_t)(-r.side) - 1 + 3 : (uint32_t)r.side - 1;
    return true;
}
// ray_cuboid.rs:6-25
static inline bool cuboid_cast_local_ray(const Vec3& he, const Ray& ray, Real max_toi, bool solid, Real& toi) {
    return aabb_cast_local_ray(Aabb(-he, he), ray, max_toi, solid, toi);
}
static inline bool cuboid_cast_local_ray_and_get_normal(const Vec3& he, const Ray& ray, Real max_toi, bool solid, RayIntersection& out) {
    return aabb_cast_local_ray_and_get_normal(Aabb(-he, he), ray, max_toi, solid, out);
}

// shape/trimesh.rs: vertices + indices + Bvh built with Binned over triangle local AABBs (:1159-1171)
struct TriMesh {
    std::vector<Vec3> vertices;
    std::vector<uint32_t> indices;  // 3 per triangle
    Bvh bvh;
    size_t num_triangles() const { return indices.size() / 3; }
    static Aabb triangle_local_aabb(const Vec3& a, const Vec3& b, const Vec3& c) {  // aabb_triangle.rs:16-30
        Vec3 mn, mx;
        for (int d = 0; d < 3; ++d) { mn[d] = rmin(rmin(a[d], b[d]), c[d]); mx[d] = rmax(rmax(a[d], b[d]), c[d]); }
        return Aabb(mn, mx);
    }
    void build(const float* v, size_t nv, const uint32_t* idx, size_t nt, BuildStrategy strat = BINNED) {
        vertices.resize(nv); for (size_t i = 0; i < nv; ++i) vertices[i] = Vec3(v[3 * i], v[3 * i + 1], v[3 * i + 2]);
        indices.assign(idx, idx + 3 * nt);
        std::vector<Aabb> aabbs(nt);
        for (size_t i = 0; i < nt; ++i) aabbs[i] = triangle_local_aabb(vertices[idx[3 * i]], vertices[idx[3 * i + 1]], vertices[idx[3 * i + 2]]);
        bvh = Bvh::from_leaves(strat, aabbs.data(), nt);
    }
    // ray_composite_shape.rs:20-39 + ray_trimesh.rs:10-14 (toi-only variant; id kept as in CompositeShapeRef)
    bool cast_local_ray(const Ray& ray, Real max_toi, bool solid, uint32_t& id, Real& toi) const {
        (void)solid;
        RealCost best;
        bool hit = bvh.find_best<RealCost>(max_toi,
            [&](const BvhNode& n, Real best_so_far) { return node_cast_ray(n, ray, best_so_far); },
            [&](uint32_t prim, Real best_so_far, RealCost& out) {
                RayIntersection ri;
                const uint32_t* t = &indices[3 * prim];
                if (!triangle_cast_local_ray_and_get_normal(vertices[t[0]], vertices[t[1]], vertices[t[2]], ray, best_so_far, ri)) return false;
                out.v = ri.time_of_impact; return true;
            }, id, best);
        if (!hit) return false;
        toi = best.v;
        return toi < max_toi;
    }
    // ray_composite_shape.rs:43-62 + ray_trimesh.rs:17-35 (backface => Face(i + num_tris))
    bool cast_local_ray_and_get_normal(const Ray& ray, Real max_toi, bool solid, uint32_t& id, RayIntersection& out) const {
        (void)solid;
        bool hit = bvh.find_best<RayIntersection>(max_toi,
            [&](const BvhNode& n, Real best_so_far) { return node_cast_ray(n, ray, best_so_far); },
            [&](uint32_t prim, Real best_so_far, RayIntersection& ri) {
                const uint32_t* t = &indices[3 * prim];
                return triangle_cast_local_ray_and_get_normal(vertices[t[0]], vertices[t[1]], vertices[t[2]], ray, best_so_far, ri);
            }, id, out);
        if (!hit) return false;
        out.feature = (out.feature == 1) ? id + (uint32_t)num_triangles() : id;
        return true;
    }
    // TriMesh::cast_local_ray_with_culling (ray_trimesh.rs:155-178): TriMeshWithCulling maps a part only when
    // RayCullingMode::check(tri.scaled_normal(), ray.dir) holds (:58-65, :104-111). culling: 1 IgnoreBackfaces, 2 IgnoreFrontfaces.
    bool cast_local_ray_with_culling(const Ray& ray, Real max_toi, int culling, uint32_t& id, RayIntersection& out) const {
        bool hit = bvh.find_best<RayIntersection>(max_toi,
            [&](const BvhNode& n, Real best_so_far) { return node_cast_ray(n, ray, best_so_far); },
            [&](uint32_t prim, Real best_so_far, RayIntersection& ri) {
                const uint32_t* t = &indices[3 * prim];
                Vec3 sn = cross(vertices[t[1]] - vertices[t[0]], vertices[t[2]] - vertices[t[0]]);  // Triangle::scaled_normal
                Real dd = dot(sn, ray.dir);
                if (!(culling == 1 ? dd < 0.0f : dd > 0.0f)) return false;
                return triangle_cast_local_ray_and_get_normal(vertices[t[0]], vertices[t[1]], vertices[t[2]], ray, best_so_far, ri);
            }, id, out);
        if (!hit) return false;
        out.feature = (out.feature == 1) ? id + (uint32_t)num_triangles() : id;
        return true;
    }
    // Brute force over all triangles: argmin toi with "own-AABB passes" filter off; ties -> min index.
    // Used by tests to adjudicate tie / ulp cases.
    bool brute_force(const Ray& ray, Real max_toi, uint32_t& id, RayIntersection& out) const {
        bool found = false;
        for (size_t i = 0; i < num_triangles(); ++i) {
            const uint32_t* t = &indices[3 * i];
            RayIntersection ri;
            if (!local_ray_intersection_with_triangle(vertices[t[0]], vertices[t[1]], vertices[t[2]], ray, ri)) continue;
            if (!(ri.time_of_impact < max_toi)) continue;
            if (!found || ri.time_of_impact < out.time_of_impact) { found = true; out = ri; id = (uint32_t)i; }
        }
        if (found) out.feature = (out.feature == 1) ? id + (uint32_t)num_triangles() : id;
        return found;
    }
};

}  // namespace pb2o

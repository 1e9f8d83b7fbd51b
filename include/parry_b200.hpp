// parry_b200 — header-only C++ mirror of the reference's Rust API for the hot path, over the C ABI (parry_b200.h).
// Names and argument meaning follow parry3d: Bvh (partitioning/bvh), TriMesh + RayCast (shape/trimesh.rs, query/ray/ray.rs),
// query::contact (query/contact/contact_shape_shape.rs). Every call is the batched form; host pointers only.
// Errors: pb2::Error (status + message); query::Unsupported surfaces per pair in the status array, like
// Result<Option<Contact>, Unsupported>.
#pragma once
#include <cstdint>
#include <limits>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "parry_b200.h"

namespace pb2 {

struct Error : std::runtime_error {
    int status;
    Error(int st, const std::string& msg) : std::runtime_error("parry_b200 status " + std::to_string(st) + ": " + msg), status(st) {}
};

struct Aabb { float mins[3], maxs[3]; };                 // bounding_volume/aabb.rs:110
struct Ray { float origin[3], dir[3]; };                 // query/ray/ray.rs:74-88
struct Isometry { float rotation[4], translation[3]; };  // nalgebra Isometry3<f32>: (i, j, k, w), (x, y, z)
struct RayIntersection { float time_of_impact; float normal[3]; uint32_t feature; };  // ray.rs:293-315 (Face id)
enum class BvhBuildStrategy { Binned = PB2_BUILD_BINNED, Ploc = PB2_BUILD_PLOC };     // bvh_tree.rs:58-78

class Context {
public:
    explicit Context(int device = 0) {
        int st = pb2_ctx_create(device, &ctx_);
        if (st != PB2_OK) throw Error(st, "no usable CUDA device (there is no CPU fallback)");
    }
    ~Context() { if (ctx_) pb2_ctx_destroy(ctx_); }
    Context(const Context&) = delete;
    Context& operator=(const Context&) = delete;
    pb2_ctx* get() const { return ctx_; }
    void check(int st) const { if (st != PB2_OK) throw Error(st, pb2_last_error(ctx_)); }
    void synchronize() const { check(pb2_ctx_synchronize(ctx_)); }
private:
    pb2_ctx* ctx_ = nullptr;
};

class Bvh {
public:
    // Bvh::from_leaves(strategy, &[Aabb])
    static Bvh from_leaves(const Context& c, BvhBuildStrategy s, const std::vector<Aabb>& leaves) {
        Bvh b(c);
        c.check(pb2_bvh_build(c.get(), leaves.empty() ? nullptr : leaves[0].mins, (uint32_t)leaves.size(), (int)s, PB2_MEM_HOST, &b.h_));
        return b;
    }
    Bvh(Bvh&& o) noexcept : c_(o.c_), h_(o.h_) { o.h_ = nullptr; }
    ~Bvh() { if (h_) pb2_bvh_destroy(c_->get(), h_); }
    uint32_t leaf_count() const { return pb2_bvh_leaf_count(h_); }
    bool is_empty() const { return leaf_count() == 0; }
    Aabb root_aabb() const { Aabb a; c_->check(pb2_bvh_root_aabb(c_->get(), h_, a.mins)); return a; }
    // Bvh::insert_or_update_partially for existing leaves (batched)
    void insert_or_update_partially(const std::vector<Aabb>& aabbs, const std::vector<uint32_t>& leaf_indices, float change_detection_margin) {
        c_->check(pb2_bvh_update_leaves(c_->get(), h_, leaf_indices.data(), aabbs[0].mins, (uint32_t)aabbs.size(), change_detection_margin, PB2_MEM_HOST));
    }
    // Bvh::insert (bvh_insert.rs:126-197), batched: unknown indices grow the tree; structural edits rebuild it
    void insert(const std::vector<Aabb>& aabbs, const std::vector<uint32_t>& leaf_indices) {
        uint32_t top = 0;
        for (uint32_t i : leaf_indices) top = i + 1 > top ? i + 1 : top;
        if (top > leaf_count()) c_->check(pb2_bvh_resize(c_->get(), h_, top));
        insert_or_update_partially(aabbs, leaf_indices, 0.0f);
        rebuild(BvhBuildStrategy::Binned);
    }
    // Bvh::remove (bvh_tree.rs:2360-2427), batched
    void remove(const std::vector<uint32_t>& leaf_indices) {
        c_->check(pb2_bvh_remove_leaves(c_->get(), h_, leaf_indices.data(), (uint32_t)leaf_indices.size(), PB2_MEM_HOST));
    }
    void refit() { c_->check(pb2_bvh_refit(c_->get(), h_)); c_->synchronize(); }
    void rebuild(BvhBuildStrategy s) { c_->check(pb2_bvh_rebuild(c_->get(), h_, (int)s)); c_->synchronize(); }
    // Bvh::intersect_aabb for a batch: CSR (offsets, leaf ids)
    std::pair<std::vector<uint32_t>, std::vector<uint32_t>> intersect_aabb(const std::vector<Aabb>& queries) const {
        std::vector<uint32_t> offs(queries.size() + 1), ids(16 * queries.size() + 1024);
        uint64_t count = 0;
        int st = pb2_bvh_intersect_aabbs(c_->get(), h_, queries[0].mins, (uint32_t)queries.size(), offs.data(), ids.data(), ids.size(), &count, PB2_MEM_HOST);
        if (st == PB2_ERR_OVERFLOW) {
            ids.resize(count);
            st = pb2_bvh_intersect_aabbs(c_->get(), h_, queries[0].mins, (uint32_t)queries.size(), offs.data(), ids.data(), ids.size(), &count, PB2_MEM_HOST);
        }
        c_->check(st);
        ids.resize(count);
        return {offs, ids};
    }
    // Bvh::traverse_bvtt_single_tree::<CHANGE_DETECTION>: f(a, b) for every overlapping leaf pair
    template <class F>
    void traverse_bvtt_single_tree(bool change_detection, F f) const {
        std::vector<uint32_t> pairs(2 * (8 * (size_t)leaf_count() + 1024));
        uint64_t count = 0;
        int st = pb2_bvh_self_pairs(c_->get(), h_, change_detection, pairs.data(), pairs.size() / 2, &count, PB2_MEM_HOST);
        if (st == PB2_ERR_OVERFLOW) {
            pairs.resize(2 * count);
            st = pb2_bvh_self_pairs(c_->get(), h_, change_detection, pairs.data(), count, &count, PB2_MEM_HOST);
        }
        c_->check(st);
        for (uint64_t i = 0; i < count; ++i) f(pairs[2 * i], pairs[2 * i + 1]);
    }
    // Bvh::leaf_pairs(other, |a, b| a.intersects(b))
    std::vector<std::pair<uint32_t, uint32_t>> leaf_pairs(const Bvh& other) const {
        std::vector<uint32_t> pairs(2 * (8 * (size_t)(leaf_count() + other.leaf_count()) + 1024));
        uint64_t count = 0;
        int st = pb2_bvh_leaf_pairs(c_->get(), h_, other.h_, pairs.data(), pairs.size() / 2, &count, PB2_MEM_HOST);
        if (st == PB2_ERR_OVERFLOW) {
            pairs.resize(2 * count);
            st = pb2_bvh_leaf_pairs(c_->get(), h_, other.h_, pairs.data(), count, &count, PB2_MEM_HOST);
        }
        c_->check(st);
        std::vector<std::pair<uint32_t, uint32_t>> out(count);
        for (uint64_t i = 0; i < count; ++i) out[i] = {pairs[2 * i], pairs[2 * i + 1]};
        return out;
    }
    pb2_bvh* get() const { return h_; }
private:
    explicit Bvh(const Context& c) : c_(&c) {}
    const Context* c_;
    pb2_bvh* h_ = nullptr;
};

class TriMesh {
public:
    // TriMesh::new(vertices, indices)
    TriMesh(const Context& c, const std::vector<float>& vertices_xyz, const std::vector<uint32_t>& indices) : c_(&c), nt_((uint32_t)indices.size() / 3) {
        c.check(pb2_trimesh_create(c.get(), vertices_xyz.data(), (uint32_t)vertices_xyz.size() / 3, indices.data(), nt_, PB2_MEM_HOST, &h_));
    }
    ~TriMesh() { if (h_) pb2_trimesh_destroy(c_->get(), h_); }
    TriMesh(const TriMesh&) = delete;
    // RayCast::cast_ray(m, ray, max_toi, solid) for a batch: toi[i] valid iff tri[i] != PB2_INVALID_U32 (None)
    void cast_ray(const Isometry* m, const std::vector<Ray>& rays, float max_time_of_impact, bool solid, std::vector<float>& toi,
                  std::vector<uint32_t>& tri) const {
        toi.resize(rays.size()); tri.resize(rays.size());
        c_->check(pb2_trimesh_cast_rays(c_->get(), h_, m ? m->rotation : nullptr, rays[0].origin, (uint32_t)rays.size(), max_time_of_impact, solid,
                                        toi.data(), tri.data(), nullptr, nullptr, PB2_MEM_HOST));
    }
    // RayCast::cast_ray_and_get_normal
    std::vector<RayIntersection> cast_ray_and_get_normal(const Isometry* m, const std::vector<Ray>& rays, float max_time_of_impact, bool solid,
                                                         std::vector<uint32_t>& tri) const {
        size_t n = rays.size();
        std::vector<float> toi(n), normal(3 * n);
        std::vector<uint32_t> feature(n);
        tri.resize(n);
        c_->check(pb2_trimesh_cast_rays(c_->get(), h_, m ? m->rotation : nullptr, rays[0].origin, (uint32_t)n, max_time_of_impact, solid, toi.data(),
                                        tri.data(), normal.data(), feature.data(), PB2_MEM_HOST));
        std::vector<RayIntersection> out(n);
        for (size_t i = 0; i < n; ++i) out[i] = RayIntersection{toi[i], {normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]}, feature[i]};
        return out;
    }
    // TriMesh::cast_ray_with_culling (ray_trimesh.rs:139-150); culling: PB2_CULL_IGNORE_BACKFACES / PB2_CULL_IGNORE_FRONTFACES
    std::vector<RayIntersection> cast_ray_with_culling(const Isometry* m, const std::vector<Ray>& rays, float max_time_of_impact, int culling,
                                                       std::vector<uint32_t>& tri) const {
        size_t n = rays.size();
        std::vector<float> toi(n), normal(3 * n);
        std::vector<uint32_t> feature(n);
        tri.resize(n);
        c_->check(pb2_trimesh_cast_rays_with_culling(c_->get(), h_, m ? m->rotation : nullptr, rays[0].origin, (uint32_t)n, max_time_of_impact,
                                                     culling, toi.data(), tri.data(), normal.data(), feature.data(), PB2_MEM_HOST));
        std::vector<RayIntersection> out(n);
        for (size_t i = 0; i < n; ++i) out[i] = RayIntersection{toi[i], {normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]}, feature[i]};
        return out;
    }
    uint32_t num_triangles() const { return nt_; }
private:
    const Context* c_;
    pb2_trimesh* h_ = nullptr;
    uint32_t nt_;
};

namespace query {
// query::contact(pos1, g1, pos2, g2, prediction) for n pairs; status[k]: 0 Ok(None), 1 Ok(Some), 2 Err(Unsupported), 3 host fallback
inline void contact(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                    const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, float prediction, std::vector<pb2_contact>& out,
                    std::vector<uint8_t>& status) {
    out.resize(g1.size()); status.resize(g1.size());
    c.check(pb2_contact_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, prediction, (uint32_t)g1.size(), out.data(),
                              status.data(), nullptr, PB2_MEM_HOST));
}
// The narrow-phase loop over a broad-phase pair list: query::contact(pos[a], g[a], pos[b], g[b], prediction) for every
// (a, b) of `pairs`; returns the Some(contact) records, pair_index[j] = index into `pairs`.
inline void contact_pairs(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& collider_shape, const std::vector<Isometry>& collider_pose,
                          const std::vector<uint32_t>& pairs /* 2 per pair */, float prediction, std::vector<pb2_contact>& out,
                          std::vector<uint32_t>& pair_index) {
    uint64_t n = pairs.size() / 2, count = 0;
    out.resize(n); pair_index.resize(n);
    c.check(pb2_contact_pairs_compact(c.get(), shapes, collider_shape.data(), collider_pose[0].rotation, (uint32_t)collider_shape.size(), pairs.data(),
                                      (uint32_t)n, prediction, out.data(), pair_index.data(), n, &count, PB2_MEM_HOST));
    out.resize(count); pair_index.resize(count);
}
// query::cast_shapes(pos1, vel1, g1, pos2, vel2, g2, options) for n pairs (shape_cast.rs:268-286); status[k]: PB2_CAST_*
struct ShapeCastOptions {   // shape_cast.rs:196-243, same defaults
    float max_time_of_impact = 3.402823466e+38f, target_distance = 0.0f;
    bool stop_at_penetration = true, compute_impact_geometry_on_penetration = true;
};
struct ShapeCastHit { float witness1[3], witness2[3], normal1[3], normal2[3], time_of_impact; };  // shape_cast.rs:30-60 (local frames)
struct Vector { float x, y, z; };
inline void cast_shapes(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                        const std::vector<Vector>& vel1, const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2,
                        const std::vector<Vector>& vel2, const ShapeCastOptions& o, std::vector<ShapeCastHit>& out, std::vector<uint8_t>& status) {
    out.resize(g1.size()); status.resize(g1.size());
    c.check(pb2_cast_shapes_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, &vel1[0].x, pos2[0].rotation, &vel2[0].x,
                                  o.max_time_of_impact, o.target_distance, o.stop_at_penetration, o.compute_impact_geometry_on_penetration,
                                  (uint32_t)g1.size(), out[0].witness1, status.data(), PB2_MEM_HOST));
}
// query::closest_points(pos1, g1, pos2, g2, max_dist) for n pairs (closest_points_shape_shape.rs:220-231)
enum class ClosestPointsKind : uint8_t { Disjoint = 0, WithinMargin = 1, Intersecting = 2 };   // query::ClosestPoints
struct ClosestPoints { float p1[3], p2[3]; };   // world-space points, meaningful for WithinMargin
inline void closest_points(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                           const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, float max_dist, std::vector<ClosestPoints>& out,
                           std::vector<ClosestPointsKind>& kind, std::vector<uint8_t>& status) {
    out.resize(g1.size()); kind.resize(g1.size()); status.resize(g1.size());
    c.check(pb2_closest_points_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, max_dist, (uint32_t)g1.size(),
                                     out[0].p1, reinterpret_cast<uint8_t*>(kind.data()), status.data(), PB2_MEM_HOST));
}
// query::distance / query::intersection_test for n pairs (distance.rs:89-97, intersection_test.rs:88-96); status 3 = host (cuboid-cuboid SAT arm)
inline void distance(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                     const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, std::vector<float>& dist, std::vector<uint8_t>& status) {
    dist.resize(g1.size()); status.resize(g1.size());
    c.check(pb2_distance_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, (uint32_t)g1.size(), dist.data(),
                               status.data(), PB2_MEM_HOST));
}
inline void intersection_test(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                              const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, std::vector<uint8_t>& hit,
                              std::vector<uint8_t>& status) {
    hit.resize(g1.size()); status.resize(g1.size());
    c.check(pb2_intersection_test_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, (uint32_t)g1.size(), hit.data(),
                                        status.data(), PB2_MEM_HOST));
}
// contact_manifolds(pos1.inv_mul(pos2), g1, g2, prediction, ..) on empty manifolds for n pairs (default_query_dispatcher.rs:629-835)
struct TrackedContact { float local_p1[3], local_p2[3], dist; uint32_t fid1, fid2; };   // contact_manifold.rs; fids are PackedFeatureId bits
struct ManifoldNormals { float local_n1[3], local_n2[3]; };
inline void contact_manifolds(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                              const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, float prediction, uint32_t max_points,
                              std::vector<ManifoldNormals>& normals, std::vector<uint32_t>& counts, std::vector<TrackedContact>& points,
                              std::vector<uint8_t>& status) {
    size_t n = g1.size();
    normals.resize(n); counts.resize(n); points.resize(n * max_points); status.resize(n);
    static_assert(sizeof(TrackedContact) == 36 && sizeof(ManifoldNormals) == 24, "layout of pb2_contact_manifolds_batch");
    c.check(pb2_contact_manifolds_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, prediction, (uint32_t)n, max_points,
                                        normals[0].local_n1, counts.data(), points[0].local_p1, status.data(), PB2_MEM_HOST));
}
// contact_manifolds on the manifolds of the previous frame (in / out): manifolds that pass ContactManifold::try_update_contacts are kept
// (kept[k] = 1), the rest recomputed; match[k * max_points + i] = index of the old point whose ContactData match_contacts hands to point i
inline void contact_manifolds_update(const Context& c, const pb2_shapes* shapes, const std::vector<uint32_t>& g1, const std::vector<Isometry>& pos1,
                                     const std::vector<uint32_t>& g2, const std::vector<Isometry>& pos2, float prediction, uint32_t max_points,
                                     std::vector<ManifoldNormals>& normals, std::vector<uint32_t>& counts, std::vector<TrackedContact>& points,
                                     std::vector<uint8_t>& status, std::vector<uint8_t>& kept, std::vector<int32_t>& match) {
    size_t n = g1.size();
    if (normals.size() != n || counts.size() != n || points.size() != n * max_points) throw std::invalid_argument("contact_manifolds_update: last frame's manifolds have another shape");
    status.resize(n); kept.resize(n); match.resize(n * max_points);
    c.check(pb2_contact_manifolds_update_batch(c.get(), shapes, g1.data(), g2.data(), pos1[0].rotation, pos2[0].rotation, prediction, (uint32_t)n,
                                               max_points, normals[0].local_n1, counts.data(), points[0].local_p1, status.data(), kept.data(),
                                               match.data(), PB2_MEM_HOST));
}
// query::contact with a Compound on one side (contact_composite_shape_shape.rs:14-76)
inline void contact_compound(const Context& c, const pb2_compounds* compounds, const std::vector<uint32_t>& compound_ids,
                             const std::vector<Isometry>& compound_poses, const std::vector<uint32_t>& shape_ids, const std::vector<Isometry>& shape_poses,
                             float prediction, bool compound_second, std::vector<pb2_contact>& out, std::vector<uint8_t>& status,
                             std::vector<uint32_t>& part) {
    size_t n = compound_ids.size();
    out.resize(n); status.resize(n); part.resize(n);
    c.check(pb2_compound_contact_shapes(c.get(), compounds, compound_ids.data(), compound_poses[0].rotation, shape_ids.data(), shape_poses[0].rotation,
                                        (uint32_t)n, prediction, compound_second ? 1 : 0, out.data(), status.data(), part.data(), PB2_MEM_HOST));
}
// query::contact between a Compound and a TriMesh, either argument order (default_query_dispatcher.rs:338-351 nested through
// contact_composite_shape_shape.rs:12-76); parts[k] = {winning part, winning triangle}
inline void contact_compound_trimesh(const Context& c, const pb2_compounds* compounds, const std::vector<uint32_t>& compound_ids,
                                     const std::vector<Isometry>& compound_poses, const pb2_trimesh* mesh, const Isometry& mesh_pose, float prediction,
                                     bool mesh_first, std::vector<pb2_contact>& out, std::vector<uint8_t>& status, std::vector<uint32_t>& parts) {
    size_t n = compound_ids.size();
    out.resize(n); status.resize(n); parts.resize(2 * n);
    c.check(pb2_compound_contact_trimesh(c.get(), compounds, compound_ids.data(), compound_poses[0].rotation, mesh, mesh_pose.rotation, (uint32_t)n,
                                         prediction, mesh_first ? 1 : 0, out.data(), status.data(), parts.data(), PB2_MEM_HOST));
}
// query::cast_shapes with a TriMesh on one side (shape_cast_composite_shape_shape.rs:65-105); hits: 13 floats per query as pb2_cast_shapes_batch
inline void cast_shapes_trimesh(const Context& c, const pb2_trimesh* mesh, const Isometry& mesh_pose, const float mesh_vel[3], const pb2_shapes* shapes,
                                const std::vector<uint32_t>& shape_ids, const std::vector<Isometry>& poses, const std::vector<float>& vels_xyz,
                                bool mesh_second, float max_time_of_impact, float target_distance, bool compute_impact_geometry_on_penetration,
                                std::vector<float>& hits, std::vector<uint8_t>& status, std::vector<uint32_t>& part) {
    size_t n = shape_ids.size();
    hits.resize(13 * n); status.resize(n); part.resize(n);
    c.check(pb2_trimesh_cast_shapes(c.get(), mesh, mesh_pose.rotation, mesh_vel, shapes, shape_ids.data(), poses[0].rotation, vels_xyz.data(),
                                    mesh_second ? 1 : 0, max_time_of_impact, target_distance, 1, compute_impact_geometry_on_penetration ? 1 : 0,
                                    (uint32_t)n, hits.data(), status.data(), part.data(), PB2_MEM_HOST));
}
// query::distance with a TriMesh on one side (distance_composite_shape_shape.rs:46-77)
inline void distance_trimesh(const Context& c, const pb2_trimesh* mesh, const Isometry& mesh_pose, const pb2_shapes* shapes,
                             const std::vector<uint32_t>& shape_ids, const std::vector<Isometry>& poses, bool mesh_second, std::vector<float>& dist,
                             std::vector<uint8_t>& status, std::vector<uint32_t>& part) {
    size_t n = shape_ids.size();
    dist.resize(n); status.resize(n); part.resize(n);
    c.check(pb2_trimesh_distance_shapes(c.get(), mesh, mesh_pose.rotation, shapes, shape_ids.data(), poses[0].rotation, mesh_second ? 1 : 0, (uint32_t)n,
                                        dist.data(), status.data(), part.data(), PB2_MEM_HOST));
}
}  // namespace query

}  // namespace pb2

// ORACLE — TEST INFRASTRUCTURE ONLY. C entry points for ctypes (tests/, __graft_entry__.smoke(),
// bench.py cpu_baseline / --impl reference). The product library never links this file.
#include "ray.hpp"
#include "ray_support_map.hpp"
#include <thread>
#include <algorithm>

using namespace pb2o;

template <class F>
static void parallel_for(size_t n, int nthreads, F f) {
    if (nthreads <= 1 || n < 2) { f(0, n); return; }
    std::vector<std::thread> th;
    size_t chunk = (n + nthreads - 1) / nthreads;
    for (int t = 0; t < nthreads; ++t) {
        size_t lo = std::min(n, t * chunk), hi = std::min(n, lo + chunk);
        if (lo < hi) th.emplace_back([=] { f(lo, hi); });
    }
    for (auto& t : th) t.join();
}

static inline Vec3 ld3(const float* p) { return Vec3(p[0], p[1], p[2]); }
static inline void st3(float* p, const Vec3& v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

extern "C" {

int pb2o_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

// ---------------- TriMesh ----------------
void* pb2o_trimesh_create(const float* verts, uint32_t nv, const uint32_t* idx, uint32_t nt, int strategy) {
    TriMesh* m = new TriMesh();
    m->build(verts, nv, idx, nt, (BuildStrategy)strategy);
    return m;
}
void pb2o_trimesh_destroy(void* m) { delete (TriMesh*)m; }
uint32_t pb2o_trimesh_num_nodes(void* m) { return (uint32_t)((TriMesh*)m)->bvh.nodes.size(); }
void pb2o_trimesh_copy_nodes(void* m, void* out) {
    TriMesh* t = (TriMesh*)m;
    memcpy(out, t->bvh.nodes.data(), t->bvh.nodes.size() * sizeof(BvhNodeWide));
}

// mode 0: reference traversal (RayCast::cast_ray / cast_ray_and_get_normal on TriMesh, ray.rs:381-411)
// mode 1: brute force over all triangles (tie/ulp adjudication, min index on ties)
// tri[i] = u32::MAX on miss. normal/feature may be NULL (=> toi-only variant, which post-filters toi < max_toi).
// mode 2 / 3: TriMesh::cast_ray_with_culling with IgnoreBackfaces / IgnoreFrontfaces (ray_trimesh.rs:139-178)
void pb2o_trimesh_cast_rays(void* mesh, const float* pose7, const float* rays, uint32_t m, float max_toi, int solid,
                            int mode, int nthreads, float* toi, uint32_t* tri, float* normal, uint32_t* feature) {
    const TriMesh* t = (const TriMesh*)mesh;
    bool has_pose = pose7 != nullptr;
    Iso pose = has_pose ? Iso::from7(pose7) : Iso();
    parallel_for(m, nthreads, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Ray ray(ld3(rays + 6 * i), ld3(rays + 6 * i + 3));
            if (has_pose) ray = ray.inverse_transform_by(pose);
            uint32_t id = UINT32_MAX; RayIntersection ri; ri.time_of_impact = 0; ri.feature = UINT32_MAX;
            bool hit;
            if (mode == 1) hit = t->brute_force(ray, max_toi, id, ri);
            else if (mode == 2 || mode == 3) hit = t->cast_local_ray_with_culling(ray, max_toi, mode - 1, id, ri);
            else if (normal || feature) hit = t->cast_local_ray_and_get_normal(ray, max_toi, solid != 0, id, ri);
            else { Real tt = 0; hit = t->cast_local_ray(ray, max_toi, solid != 0, id, tt); ri.time_of_impact = tt; }
            if (!hit) { toi[i] = 0.0f; tri[i] = UINT32_MAX; if (normal) st3(normal + 3 * i, Vec3()); if (feature) feature[i] = UINT32_MAX; continue; }
            toi[i] = ri.time_of_impact; tri[i] = id;
            if (normal) st3(normal + 3 * i, has_pose ? pose.transform_vector(ri.normal) : ri.normal);
            if (feature) feature[i] = ri.feature;
        }
    });
}

// ---------------- Bvh ----------------
void* pb2o_bvh_create(const float* aabbs, uint32_t n, int strategy) {
    Bvh* b = new Bvh(Bvh::from_leaves((BuildStrategy)strategy, (const Aabb*)aabbs, n));
    return b;
}
void pb2o_bvh_destroy(void* b) { delete (Bvh*)b; }
uint32_t pb2o_bvh_num_nodes(void* b) { return (uint32_t)((Bvh*)b)->nodes.size(); }
void pb2o_bvh_copy_nodes(void* b, void* out) {
    Bvh* t = (Bvh*)b;
    memcpy(out, t->nodes.data(), t->nodes.size() * sizeof(BvhNodeWide));
}
// parents / leaf_node_indices as u64 (BvhNodeIndex = usize)
void pb2o_bvh_copy_parents(void* b, uint64_t* out) { Bvh* t = (Bvh*)b; for (size_t i = 0; i < t->parents.size(); ++i) out[i] = t->parents[i].v; }
void pb2o_bvh_copy_leaf_node_indices(void* b, uint64_t* out) { Bvh* t = (Bvh*)b; for (size_t i = 0; i < t->leaf_node_indices.size(); ++i) out[i] = t->leaf_node_indices[i].v; }
void pb2o_bvh_update_leaves(void* b, const uint32_t* ids, const float* aabbs, uint32_t n, float margin) {
    Bvh* t = (Bvh*)b;
    for (uint32_t i = 0; i < n; ++i) t->insert_or_update_partially(((const Aabb*)aabbs)[i], ids ? ids[i] : i, margin);
}
void pb2o_bvh_refit(void* b) { ((Bvh*)b)->refit(); }
void pb2o_bvh_refit_without_opt(void* b) { ((Bvh*)b)->refit_without_opt(); }
void pb2o_bvh_rebuild(void* b, int strategy) { ((Bvh*)b)->rebuild(strategy == 1 ? PLOC : BINNED); }

// Bvh::intersect_aabb for a batch. offsets has m+1 entries. Returns total count (leaf_ids filled up to cap,
// in the reference's iteration order per query).
uint64_t pb2o_bvh_intersect_aabbs(void* b, const float* q, uint32_t m, int nthreads, uint32_t* offsets, uint32_t* leaf_ids, uint64_t cap) {
    const Bvh* t = (const Bvh*)b;
    std::vector<std::vector<uint32_t>> res(m);
    parallel_for(m, nthreads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) t->intersect_aabb(((const Aabb*)q)[i], res[i]);
    });
    uint64_t total = 0;
    for (uint32_t i = 0; i < m; ++i) {
        offsets[i] = (uint32_t)total;
        for (uint32_t id : res[i]) { if (total < cap) leaf_ids[total] = id; total++; }
    }
    offsets[m] = (uint32_t)total;
    return total;
}
// Bvh::traverse_bvtt_single_tree::<CHANGE_DETECTION>; pairs in the reference's emission order and orientation.
uint64_t pb2o_bvh_self_pairs(void* b, int change_detection, uint32_t* pairs, uint64_t cap) {
    const Bvh* t = (const Bvh*)b;
    uint64_t count = 0;
    auto f = [&](uint32_t a, uint32_t c) { if (count < cap) { pairs[2 * count] = a; pairs[2 * count + 1] = c; } count++; };
    if (change_detection) t->traverse_bvtt_single_tree<true>(f); else t->traverse_bvtt_single_tree<false>(f);
    return count;
}
// Bvh::leaf_pairs(other, |a, b| a.intersects(b))
uint64_t pb2o_bvh_leaf_pairs(void* b1, void* b2, uint32_t* pairs, uint64_t cap) {
    const Bvh* t1 = (const Bvh*)b1; const Bvh* t2 = (const Bvh*)b2;
    uint64_t count = 0;
    auto f = [&](uint32_t a, uint32_t c) { if (count < cap) { pairs[2 * count] = a; pairs[2 * count + 1] = c; } count++; };
    t1->leaf_pairs(*t2, [](const BvhNode& a, const BvhNode& c) { return a.intersects(c); }, f);
    return count;
}

// Bvh::cast_ray with typed leaves: kind 0 = Ball (param = radius), kind 1 = Cuboid (param = half extents), kind 2 =
// ConvexPolyhedron (points + 3 * first[i], count[i] points; ray_support_map.rs:163-181).
// leaf i: pose7[i], kinds[i], params[3*i..]. toi-only when normal == NULL.
void pb2o_bvh_cast_rays_shapes2(void* b, const uint8_t* kinds, const float* params, const float* points, const uint32_t* first,
                                const uint32_t* count, const float* poses7, const float* rays, uint32_t m, float max_toi, int solid,
                                int nthreads, float* toi, uint32_t* leaf, float* normal, uint32_t* feature) {
    const Bvh* t = (const Bvh*)b;
    parallel_for(m, nthreads, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Ray ray(ld3(rays + 6 * i), ld3(rays + 6 * i + 3));
            uint32_t id = UINT32_MAX; RayIntersection best; best.time_of_impact = 0; best.feature = UINT32_MAX;
            bool hit = t->find_best<RayIntersection>(max_toi,
                [&](const BvhNode& n, Real bsf) { return node_cast_ray(n, ray, bsf); },
                [&](uint32_t prim, Real bsf, RayIntersection& ri) {
                    Iso pose = Iso::from7(poses7 + 7 * prim);
                    Ray ls = ray.inverse_transform_by(pose);
                    bool h;
                    if (kinds[prim] == 2) {
                        // RayCast::cast_local_ray defaults to the normal variant (ray.rs:362-371)
                        h = support_map_cast_local_ray_and_get_normal(SupportShape::convex(points + 3 * (size_t)first[prim], count[prim]), ls, bsf,
                                                                      solid != 0, ri);
                        if (h && normal) ri.normal = pose.transform_vector(ri.normal);
                    } else if (normal) {
                        h = kinds[prim] == 0 ? ball_cast_local_ray_and_get_normal(params[3 * prim], ls, bsf, solid != 0, ri)
                                             : cuboid_cast_local_ray_and_get_normal(ld3(params + 3 * prim), ls, bsf, solid != 0, ri);
                        if (h) ri.normal = pose.transform_vector(ri.normal);
                    } else {
                        Real tt = 0;
                        h = kinds[prim] == 0 ? ball_cast_local_ray(params[3 * prim], ls, bsf, solid != 0, tt)
                                             : cuboid_cast_local_ray(ld3(params + 3 * prim), ls, bsf, solid != 0, tt);
                        ri.time_of_impact = tt; ri.feature = 0;
                    }
                    return h;
                }, id, best);
            if (!hit) { toi[i] = 0.0f; leaf[i] = UINT32_MAX; if (normal) st3(normal + 3 * i, Vec3()); if (feature) feature[i] = UINT32_MAX; continue; }
            toi[i] = best.time_of_impact; leaf[i] = id;
            if (normal) st3(normal + 3 * i, best.normal);
            if (feature) feature[i] = best.feature;
        }
    });
}
void pb2o_bvh_cast_rays_shapes(void* b, const uint8_t* kinds, const float* params, const float* poses7, const float* rays,
                               uint32_t m, float max_toi, int solid, int nthreads, float* toi, uint32_t* leaf, float* normal, uint32_t* feature) {
    pb2o_bvh_cast_rays_shapes2(b, kinds, params, nullptr, nullptr, nullptr, poses7, rays, m, max_toi, solid, nthreads, toi, leaf, normal, feature);
}
// RayCast for ConvexPolyhedron, one shape (world-space ray, pose7 may be NULL). returns 1 on hit.
int pb2o_convex_cast_ray(const float* points, uint32_t n, const float* pose7, const float* ray6, float max_toi, int solid, float* toi, float* normal) {
    Iso pose = pose7 ? Iso::from7(pose7) : Iso();
    Ray ray(ld3(ray6), ld3(ray6 + 3));
    Ray ls = pose7 ? ray.inverse_transform_by(pose) : ray;
    RayIntersection ri;
    if (!support_map_cast_local_ray_and_get_normal(SupportShape::convex(points, n), ls, max_toi, solid != 0, ri)) return 0;
    *toi = ri.time_of_impact;
    if (normal) st3(normal, pose7 ? pose.transform_vector(ri.normal) : ri.normal);
    return 1;
}

// Single-shape ray casts (RayCast for Ball / Cuboid / Triangle), world-space with pose (ray.rs:381-411).
// kind 0 ball (p[0]=r), 1 cuboid (p=he), 2 triangle (p = a,b,c: 9 floats). returns 1 on hit.
int pb2o_shape_cast_ray(int kind, const float* p, const float* pose7, const float* ray6, float max_toi, int solid, float* toi, float* normal, uint32_t* feature) {
    Iso pose = pose7 ? Iso::from7(pose7) : Iso();
    Ray ray(ld3(ray6), ld3(ray6 + 3));
    Ray ls = pose7 ? ray.inverse_transform_by(pose) : ray;
    RayIntersection ri; bool h;
    if (kind == 0) h = ball_cast_local_ray_and_get_normal(p[0], ls, max_toi, solid != 0, ri);
    else if (kind == 1) h = cuboid_cast_local_ray_and_get_normal(ld3(p), ls, max_toi, solid != 0, ri);
    else h = triangle_cast_local_ray_and_get_normal(ld3(p), ld3(p + 3), ld3(p + 6), ls, max_toi, ri);
    if (!h) return 0;
    *toi = ri.time_of_impact; if (normal) st3(normal, pose7 ? pose.transform_vector(ri.normal) : ri.normal); if (feature) *feature = ri.feature;
    return 1;
}
// toi-only single-shape casts (separate code path in the reference for ball/cuboid)
int pb2o_shape_cast_ray_toi(int kind, const float* p, const float* pose7, const float* ray6, float max_toi, int solid, float* toi) {
    Iso pose = pose7 ? Iso::from7(pose7) : Iso();
    Ray ray(ld3(ray6), ld3(ray6 + 3));
    Ray ls = pose7 ? ray.inverse_transform_by(pose) : ray;
    if (kind == 0) return ball_cast_local_ray(p[0], ls, max_toi, solid != 0, *toi);
    if (kind == 1) return cuboid_cast_local_ray(ld3(p), ls, max_toi, solid != 0, *toi);
    RayIntersection ri;
    if (!triangle_cast_local_ray_and_get_normal(ld3(p), ld3(p + 3), ld3(p + 6), ls, max_toi, ri)) return 0;
    *toi = ri.time_of_impact; return 1;
}

// clip_aabb_line (clip_aabb_line.rs:79-187) on one box and one line, so that the reference's own unit test (:192-206) can be run on
// the restatement. returns 1 when the line meets the box; near_far[0..1] = the two parameters.
int pb2o_clip_aabb_line(const float* aabb6, const float* origin3, const float* dir3, float* near_far) {
    ClipHit near, far;
    if (!clip_aabb_line(Aabb(ld3(aabb6), ld3(aabb6 + 3)), ld3(origin3), ld3(dir3), near, far)) return 0;
    if (near_far) { near_far[0] = near.t; near_far[1] = far.t; }
    return 1;
}

}  // extern "C"

// ---------------- per-shape AABBs (Shape::compute_aabb) ----------------
extern "C" void pb2o_shape_aabbs(const uint8_t* kinds, const float* params /* n x 3 */, const float* points, const uint32_t* first,
                                 const uint32_t* count, const float* poses7, uint32_t n, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        Iso pos = Iso::from7(poses7 + 7 * i);
        Aabb a;
        if (kinds[i] == 0) {  // aabb_ball.rs:8-33
            Real r = params[3 * i];
            a = Aabb(pos.tra + Vec3(-r, -r, -r), pos.tra + Vec3(r, r, r));
        } else if (kinds[i] == 1) {  // aabb_cuboid.rs:9-16
            Vec3 he = pos.absolute_transform_vector(ld3(params + 3 * i));
            a = Aabb(pos.tra - he, pos.tra + he);
        } else {  // aabb_convex_polyhedron.rs:8 -> aabb_utils.rs:66-87
            const float* p = points + 3 * first[i];
            Vec3 w0 = pos.transform_point(ld3(p));
            a = Aabb(w0, w0);
            for (uint32_t k = 1; k < count[i]; ++k) { Vec3 w = pos.transform_point(ld3(p + 3 * k)); a.mins = vinf(a.mins, w); a.maxs = vsup(a.maxs, w); }
        }
        st3(out + 6 * i, a.mins); st3(out + 6 * i + 3, a.maxs);
    }
}

// ---------------- query::contact ----------------
#include "contact.hpp"
#include "shape_cast.hpp"
#include "manifold.hpp"

static inline ShapeRef make_shape(const uint8_t* kinds, const float* params4, const float* points, uint32_t id) {
    ShapeRef s; s.kind = kinds[id]; s.radius = params4[4 * id]; s.half_extents = ld3(params4 + 4 * id); s.points = nullptr; s.num_points = 0;
    if (s.kind == SHAPE_CONVEX || s.kind == SHAPE_TRIANGLE) {
        uint32_t first, cnt; memcpy(&first, params4 + 4 * id, 4); memcpy(&cnt, params4 + 4 * id + 1, 4);
        s.points = points + 3 * first; s.num_points = cnt;
    }
    return s;
}

extern "C" {
// query::contact for n pairs. Shape table layout == pb2_shapes_create's. out: n x 13 floats; status: 0/1/2/3.
// stats (optional, n x 6 ints): gjk iters, used_epa, epa iters, max faces, max vertices, max heap.
void pb2o_contact_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* shape1, const uint32_t* shape2,
                        const float* pos1, const float* pos2, float prediction, uint32_t n, int nthreads, float* out, uint8_t* status, int32_t* stats) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Contact c = Contact();
            GjkEpaStats gs;
            int st = query_contact(Iso::from7(pos1 + 7 * k), s1, Iso::from7(pos2 + 7 * k), s2, prediction, c, &gs);
            status[k] = (uint8_t)st;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) { st3(o, c.point1); st3(o + 3, c.point2); st3(o + 6, c.normal1); st3(o + 9, c.normal2); o[12] = c.dist; }
            else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
            if (stats) {
                int32_t* s = stats + 6 * k;
                s[0] = gs.gjk_iters; s[1] = gs.used_epa; s[2] = gs.epa.niter; s[3] = (int32_t)gs.epa.max_faces; s[4] = (int32_t)gs.epa.max_vertices; s[5] = (int32_t)gs.epa.max_heap;
            }
        }
    });
}
// QueryDispatcher::contact(pos1.inv_mul(pos2), g1, g2, prediction) for n pairs: the same dispatch without Contact::transform_by_mut,
// i.e. the contact in the two shapes' local frames (what the composite arms consume).
void pb2o_contact_local_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* shape1, const uint32_t* shape2,
                              const float* pos1, const float* pos2, float prediction, uint32_t n, int nthreads, float* out, uint8_t* status) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Contact c = Contact();
            int st = dispatch_contact(Iso::from7(pos1 + 7 * k).inv_mul(Iso::from7(pos2 + 7 * k)), s1, s2, prediction, c);
            status[k] = (uint8_t)st;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) { st3(o, c.point1); st3(o + 3, c.point2); st3(o + 6, c.normal1); st3(o + 9, c.normal2); o[12] = c.dist; }
            else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        }
    });
}
// query::contact between Compound compound_id[k] (parts comp_first[c] .. + comp_count[c] of the part table: part_shape = index
// into the shape table, part_pose7) and shape[k]. pos_c / pos_s: poses of the compound and of the shape. compound_second != 0:
// the call was contact(pos_s, shape, pos_c, compound) (result flipped accordingly). part[k] = winning part (index within
// the compound) or u32::MAX.
void pb2o_compound_contact_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* comp_first,
                                 const uint32_t* comp_count, const uint32_t* part_shape, const float* part_pose7, const uint32_t* compound_id,
                                 const float* pos_c, const uint32_t* shape, const float* pos_s, float prediction, int compound_second, uint32_t n,
                                 int nthreads, float* out, uint8_t* status, uint32_t* part) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        std::vector<ShapeRef> ps; std::vector<Iso> pp;
        for (size_t k = lo; k < hi; ++k) {
            uint32_t c = compound_id[k], f = comp_first[c], cnt = comp_count[c];
            ps.resize(cnt); pp.resize(cnt);
            for (uint32_t i = 0; i < cnt; ++i) { ps[i] = make_shape(kinds, params4, points, part_shape[f + i]); pp[i] = Iso::from7(part_pose7 + 7 * (size_t)(f + i)); }
            CompoundRef comp{ps.data(), pp.data(), cnt};
            ShapeRef s2 = make_shape(kinds, params4, points, shape[k]);
            Iso pc = Iso::from7(pos_c + 7 * k), psh = Iso::from7(pos_s + 7 * k);
            Contact ct = Contact(); uint32_t id = UINT32_MAX;
            int st = compound_second ? query_contact_compound(psh, pc, comp, s2, true, prediction, ct, id)
                                     : query_contact_compound(pc, psh, comp, s2, false, prediction, ct, id);
            status[k] = (uint8_t)st; part[k] = st == CONTACT_SOME ? id : UINT32_MAX;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) { st3(o, ct.point1); st3(o + 3, ct.point2); st3(o + 6, ct.normal1); st3(o + 9, ct.normal2); o[12] = ct.dist; }
            else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        }
    });
}
// query::closest_points for n pairs. out: n x 6 (p1, p2 in world space, zeros unless WithinMargin); kind: 0 Disjoint,
// 1 WithinMargin, 2 Intersecting; status: 1 ok, 3 needs hull topology.
void pb2o_closest_points_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* shape1, const uint32_t* shape2,
                               const float* pos1, const float* pos2, float max_dist, uint32_t n, int nthreads, float* out, uint8_t* kind,
                               uint8_t* status) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Vec3 p1, p2; int qs;
            int kd = query_closest_points(Iso::from7(pos1 + 7 * k), s1, Iso::from7(pos2 + 7 * k), s2, max_dist, p1, p2, qs);
            kind[k] = (uint8_t)kd; status[k] = (uint8_t)qs;
            st3(out + 6 * k, p1); st3(out + 6 * k + 3, p2);
        }
    });
}
// QueryDispatcher::contact_manifolds for n pairs of Ball / Cuboid shapes, first frame (empty incoming manifolds), with
// pos12 = pos1.inv_mul(pos2). normals: n x 6 (local_n1, local_n2); counts: n; pts: n x max_points x 9 words {local_p1, local_p2,
// dist, fid1, fid2 (PackedFeatureId bits)}; status: 0 ok, 2 unsupported pair (a ConvexPolyhedron), 4 more than max_points.
static void manifolds_batch_impl(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* hull_face_first,
                                 const uint32_t* hull_face_count, const float* face_normal, const uint32_t* face_first,
                                 const uint32_t* face_count, const uint32_t* vertices_adj_to_face, const uint32_t* edges_adj_to_face,
                                 const uint32_t* vert_first, const uint32_t* vert_count, const uint32_t* faces_adj_to_vertex,
                                 const uint32_t* edges_adj_to_vertex, const uint32_t* hull_edge_first, const float* edge_dir,
                                 const uint32_t* shape1, const uint32_t* shape2, const float* pos1, const float* pos2, float prediction,
                                 uint32_t n, uint32_t max_points, int nthreads, float* normals, uint32_t* counts, float* pts, uint8_t* status,
                                 bool persistent, bool seed_gjk, uint8_t* kept, int32_t* match) {
    // hull_face_first / hull_face_count: per shape-table entry (ignored for balls and cuboids), NULL = no topology supplied
    auto topo = [=](uint32_t sid, HullTopology& t) -> const HullTopology* {
        if (!hull_face_first || kinds[sid] != 2) return nullptr;
        uint32_t f0 = hull_face_first[sid];
        t.face_normal = face_normal + 3 * (size_t)f0; t.face_first = face_first + f0; t.face_count = face_count + f0;
        t.vertices_adj_to_face = vertices_adj_to_face; t.edges_adj_to_face = edges_adj_to_face; t.num_faces = hull_face_count[sid];
        if (vert_first) {   // vertex-side arrays are indexed by the table's global point index; edges per table entry
            uint32_t p0; memcpy(&p0, params4 + 4 * (size_t)sid, 4);
            t.vert_first = vert_first + p0; t.vert_count = vert_count + p0; t.faces_adj_to_vertex = faces_adj_to_vertex;
            t.edges_adj_to_vertex = edges_adj_to_vertex; t.edge_dir = edge_dir + 3 * (size_t)hull_edge_first[sid];
        }
        return t.num_faces ? &t : nullptr;
    };
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        Manifold m;
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Iso pos12 = Iso::from7(pos1 + 7 * k).inv_mul(Iso::from7(pos2 + 7 * k));
            HullTopology ta, tb;
            float* q0 = pts + (size_t)k * max_points * 9;
            std::vector<TrackedContact> old;
            m.clear(); m.local_n1 = Vec3(); m.local_n2 = Vec3();
            if (persistent) {   // the caller's ContactManifold as last frame left it
                m.local_n1 = ld3(normals + 6 * k); m.local_n2 = ld3(normals + 6 * k + 3);
                for (uint32_t i = 0; i < counts[k] && i < max_points; ++i) {
                    TrackedContact t; t.local_p1 = ld3(q0 + 9 * i); t.local_p2 = ld3(q0 + 9 * i + 3); t.dist = q0[9 * i + 6];
                    memcpy(&t.fid1, q0 + 9 * i + 7, 4); memcpy(&t.fid2, q0 + 9 * i + 8, 4);
                    m.points.push_back(t);
                }
                old = m.points;
            }
            bool was_kept = false;
            int st = dispatch_manifold(pos12, s1, s2, prediction, m, topo(shape1[k], ta), topo(shape2[k], tb), persistent, &was_kept, seed_gjk);
            if (kept) kept[k] = was_kept ? 1 : 0;
            if (match) {   // ContactManifold::match_contacts (contact_manifold.rs:761-770): the last old contact with both feature ids equal
                for (uint32_t i = 0; i < max_points; ++i) {
                    int32_t j = -1;
                    if (i < m.points.size()) {
                        if (was_kept) j = (int32_t)i;
                        else for (size_t o = 0; o < old.size(); ++o) if (old[o].fid1 == m.points[i].fid1 && old[o].fid2 == m.points[i].fid2) j = (int32_t)o;
                    }
                    match[(size_t)k * max_points + i] = j;
                }
            }
            uint32_t cnt = (uint32_t)m.points.size();
            if (cnt > max_points) { st = 4; cnt = max_points; }
            if (cnt) { st3(normals + 6 * k, m.local_n1); st3(normals + 6 * k + 3, m.local_n2); }
            else {
                for (int i = 0; i < 6; ++i) normals[6 * k + i] = 0.0f;
                // an empty pfm_pfm manifold keeps the direction GJK answered NoIntersection with (next frame's seed)
                bool pfm = st == MANIFOLD_OK && s1.kind != SHAPE_BALL && s2.kind != SHAPE_BALL && !(s1.kind == SHAPE_CUBOID && s2.kind == SHAPE_CUBOID);
                if (pfm && !was_kept) st3(normals + 6 * k, m.local_n1);
            }
            counts[k] = cnt; status[k] = (uint8_t)st;
            float* q = pts + (size_t)k * max_points * 9;
            for (uint32_t i = 0; i < max_points; ++i) {
                float* o = q + 9 * i;
                if (i < cnt) {
                    const TrackedContact& t = m.points[i];
                    st3(o, t.local_p1); st3(o + 3, t.local_p2); o[6] = t.dist; memcpy(o + 7, &t.fid1, 4); memcpy(o + 8, &t.fid2, 4);
                } else for (int j = 0; j < 9; ++j) o[j] = 0.0f;
            }
        }
    });
}
void pb2o_contact_manifolds_batch2(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* hull_face_first,
                                   const uint32_t* hull_face_count, const float* face_normal, const uint32_t* face_first,
                                   const uint32_t* face_count, const uint32_t* vertices_adj_to_face, const uint32_t* edges_adj_to_face,
                                   const uint32_t* vert_first, const uint32_t* vert_count, const uint32_t* faces_adj_to_vertex,
                                   const uint32_t* edges_adj_to_vertex, const uint32_t* hull_edge_first, const float* edge_dir,
                                   const uint32_t* shape1, const uint32_t* shape2, const float* pos1, const float* pos2, float prediction,
                                   uint32_t n, uint32_t max_points, int nthreads, float* normals, uint32_t* counts, float* pts, uint8_t* status) {
    manifolds_batch_impl(kinds, params4, points, hull_face_first, hull_face_count, face_normal, face_first, face_count, vertices_adj_to_face,
                         edges_adj_to_face, vert_first, vert_count, faces_adj_to_vertex, edges_adj_to_vertex, hull_edge_first, edge_dir, shape1, shape2,
                         pos1, pos2, prediction, n, max_points, nthreads, normals, counts, pts, status, false, false, nullptr, nullptr);
}
// QueryDispatcher::contact_manifolds called again with last frame's manifolds (normals / counts / pts hold them on entry and the
// new ones on return): the cuboid-cuboid and pfm_pfm arms keep a manifold that passes try_update_contacts (kept[k] = 1), everything
// else is recomputed; match: n x max_points, for each new point the index of the old point match_contacts would take its data from
// (kept manifolds: the identity), -1 = none. seed_gjk != 0: the pfm_pfm recomputation starts GJK from last frame's normal
// (contact_manifolds_pfm_pfm.rs:66), as the reference does; 0: from the default direction, as the GPU path does.
void pb2o_contact_manifolds_update_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* hull_face_first,
                                         const uint32_t* hull_face_count, const float* face_normal, const uint32_t* face_first,
                                         const uint32_t* face_count, const uint32_t* vertices_adj_to_face, const uint32_t* edges_adj_to_face,
                                         const uint32_t* vert_first, const uint32_t* vert_count, const uint32_t* faces_adj_to_vertex,
                                         const uint32_t* edges_adj_to_vertex, const uint32_t* hull_edge_first, const float* edge_dir,
                                         const uint32_t* shape1, const uint32_t* shape2, const float* pos1, const float* pos2, float prediction,
                                         uint32_t n, uint32_t max_points, int nthreads, int seed_gjk, float* normals, uint32_t* counts, float* pts,
                                         uint8_t* status, uint8_t* kept, int32_t* match) {
    manifolds_batch_impl(kinds, params4, points, hull_face_first, hull_face_count, face_normal, face_first, face_count, vertices_adj_to_face,
                         edges_adj_to_face, vert_first, vert_count, faces_adj_to_vertex, edges_adj_to_vertex, hull_edge_first, edge_dir, shape1, shape2,
                         pos1, pos2, prediction, n, max_points, nthreads, normals, counts, pts, status, true, seed_gjk != 0, kept, match);
}
void pb2o_contact_manifolds_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* shape1,
                                  const uint32_t* shape2, const float* pos1, const float* pos2, float prediction, uint32_t n,
                                  uint32_t max_points, int nthreads, float* normals, uint32_t* counts, float* pts, uint8_t* status) {
    pb2o_contact_manifolds_batch2(kinds, params4, points, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                  nullptr, nullptr, nullptr, shape1, shape2, pos1, pos2,
                                  prediction, n, max_points, nthreads, normals, counts, pts, status);
}
// ContactManifold::try_update_contacts on n manifolds in the layout of pb2o_contact_manifolds_batch (normals n x 6, counts n,
// pts n x max_points x 9), in place; kept[k] = 1 when the manifold survives under pos1[k].inv_mul(pos2[k]).
void pb2o_manifolds_try_update(const float* pos1, const float* pos2, uint32_t n, uint32_t max_points, float* normals, const uint32_t* counts,
                               float* pts, uint8_t* kept) {
    for (uint32_t k = 0; k < n; ++k) {
        Manifold m;
        m.local_n1 = ld3(normals + 6 * (size_t)k); m.local_n2 = ld3(normals + 6 * (size_t)k + 3);
        float* q = pts + (size_t)k * max_points * 9;
        for (uint32_t i = 0; i < counts[k]; ++i) {
            TrackedContact t; t.local_p1 = ld3(q + 9 * i); t.local_p2 = ld3(q + 9 * i + 3); t.dist = q[9 * i + 6];
            memcpy(&t.fid1, q + 9 * i + 7, 4); memcpy(&t.fid2, q + 9 * i + 8, 4);
            m.points.push_back(t);
        }
        Iso pos12 = Iso::from7(pos1 + 7 * (size_t)k).inv_mul(Iso::from7(pos2 + 7 * (size_t)k));
        kept[k] = manifold_try_update_contacts(m, pos12) ? 1 : 0;
        for (uint32_t i = 0; i < counts[k]; ++i) { st3(q + 9 * i, m.points[i].local_p1); q[9 * i + 6] = m.points[i].dist; }
    }
}
// query::contact between two Compounds (ids into the same compound table) for n pairs; result in world space; parts: n x 2.
void pb2o_compound_compound_contact_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* comp_first,
                                          const uint32_t* comp_count, const uint32_t* part_shape, const float* part_pose7, const uint32_t* id1,
                                          const float* pos1, const uint32_t* id2, const float* pos2, float prediction, uint32_t n, int nthreads,
                                          float* out, uint8_t* status, uint32_t* parts) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        std::vector<ShapeRef> s1, s2; std::vector<Iso> q1, q2;
        auto load = [&](uint32_t c, std::vector<ShapeRef>& ss, std::vector<Iso>& qq) {
            uint32_t f = comp_first[c], cnt = comp_count[c];
            ss.resize(cnt); qq.resize(cnt);
            for (uint32_t i = 0; i < cnt; ++i) { ss[i] = make_shape(kinds, params4, points, part_shape[f + i]); qq[i] = Iso::from7(part_pose7 + 7 * (size_t)(f + i)); }
        };
        for (size_t k = lo; k < hi; ++k) {
            load(id1[k], s1, q1); load(id2[k], s2, q2);
            CompoundRef a{s1.data(), q1.data(), (uint32_t)s1.size()}, b{s2.data(), q2.data(), (uint32_t)s2.size()};
            Iso p1 = Iso::from7(pos1 + 7 * k), p2 = Iso::from7(pos2 + 7 * k);
            Contact ct = Contact(); uint32_t i1 = UINT32_MAX, i2 = UINT32_MAX;
            int st = contact_compound_compound(p1.inv_mul(p2), a, b, prediction, ct, i1, i2);
            status[k] = (uint8_t)st; parts[2 * k] = st == CONTACT_SOME ? i1 : UINT32_MAX; parts[2 * k + 1] = st == CONTACT_SOME ? i2 : UINT32_MAX;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) {
                ct.point1 = p1.transform_point(ct.point1); ct.point2 = p2.transform_point(ct.point2);
                ct.normal1 = p1.transform_vector(ct.normal1); ct.normal2 = p2.transform_vector(ct.normal2);
                st3(o, ct.point1); st3(o + 3, ct.point2); st3(o + 6, ct.normal1); st3(o + 9, ct.normal2); o[12] = ct.dist;
            } else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        }
    });
}
// query::cast_shapes for n pairs (shape_cast.rs:268-286). vel1/vel2: n x 3. out: n x 13 floats {witness1, witness2, normal1,
// normal2, time_of_impact} (witness/normal i in the local frame of shape i, as the reference returns them);
// status: 0 None, 1 Converged, 2 PenetratingOrWithinTargetDist.
void pb2o_cast_shapes_batch(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* shape1, const uint32_t* shape2,
                            const float* pos1, const float* vel1, const float* pos2, const float* vel2, float max_toi, float target_distance,
                            int stop_at_penetration, int compute_impact_geometry_on_penetration, uint32_t n, int nthreads, float* out,
                            uint8_t* status) {
    ShapeCastOptions o;
    o.max_time_of_impact = max_toi; o.target_distance = target_distance; o.stop_at_penetration = stop_at_penetration != 0;
    o.compute_impact_geometry_on_penetration = compute_impact_geometry_on_penetration != 0;
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            ShapeCastHit h;
            bool some = cast_shapes(Iso::from7(pos1 + 7 * k), ld3(vel1 + 3 * k), s1, Iso::from7(pos2 + 7 * k), ld3(vel2 + 3 * k), s2, o, h);
            float* q = out + 13 * k;
            if (some) {
                st3(q, h.witness1); st3(q + 3, h.witness2); st3(q + 6, h.normal1); st3(q + 9, h.normal2); q[12] = h.time_of_impact;
                status[k] = h.status == CAST_PENETRATING ? 2 : 1;
            } else {
                for (int i = 0; i < 13; ++i) q[i] = 0.0f;
                status[k] = 0;
            }
        }
    });
}
// query::cast_shapes with a TriMesh on either side, one pair (oracle groundwork; shape_cast_composite_shape_shape.rs). The mesh is
// shape 1 (mesh_second == 0) or shape 2; the other shape is mesh_other when non-null, else entry `shape` of the table. out: 13 floats
// as pb2o_cast_shapes_batch; part = winning triangle of the outer mesh when it is shape 1. Returns the status (0 = None).
int pb2o_trimesh_cast_shapes(void* mesh, const float* mesh_pose7, const float* mesh_vel3, void* mesh_other, const uint8_t* kinds,
                             const float* params4, const float* points, uint32_t shape, const float* pose7, const float* vel3, int mesh_second,
                             float max_toi, float target_distance, int stop_at_penetration, int compute_impact_geometry_on_penetration, float* out,
                             uint32_t* part) {
    ShapeCastOptions o;
    o.max_time_of_impact = max_toi; o.target_distance = target_distance; o.stop_at_penetration = stop_at_penetration != 0;
    o.compute_impact_geometry_on_penetration = compute_impact_geometry_on_penetration != 0;
    ShapeRef sr;
    CastShape m{nullptr, (const TriMesh*)mesh}, other{nullptr, (const TriMesh*)mesh_other};
    if (!mesh_other) { sr = make_shape(kinds, params4, points, shape); other.shape = &sr; }
    Iso pm = Iso::from7(mesh_pose7), po = Iso::from7(pose7);
    Vec3 vm = ld3(mesh_vel3), vo = ld3(vel3);
    ShapeCastHit h;
    uint32_t p1 = UINT32_MAX;
    bool some = mesh_second ? cast_shapes_any(po, vo, other, pm, vm, m, o, h, &p1) : cast_shapes_any(pm, vm, m, po, vo, other, o, h, &p1);
    if (part) *part = p1;
    if (!some) { for (int i = 0; i < 13; ++i) out[i] = 0.0f; return 0; }
    st3(out, h.witness1); st3(out + 3, h.witness2); st3(out + 6, h.normal1); st3(out + 9, h.normal2); out[12] = h.time_of_impact;
    return h.status == CAST_PENETRATING ? 2 : 1;
}
// query::contact(pos1, &TriMesh, pos2[k], shape2[k], prediction) for n shapes against one mesh (composite arm of
// DefaultQueryDispatcher::contact -> contact_composite_shape_shape). mesh_pose7: one pose. part[k] = winning triangle or
// u32::MAX. ties != 0: equal-dist ties go to the smallest triangle index (the GPU's documented rule) instead of BVH order.
void pb2o_trimesh_contact_batch(void* mesh, const float* mesh_pose7, const uint8_t* kinds, const float* params4, const float* points,
                                const uint32_t* shape2, const float* pos2, float prediction, uint32_t n, int nthreads, int ties, float* out,
                                uint8_t* status, uint32_t* part) {
    const TriMesh* tm = (const TriMesh*)mesh;
    Iso pos1 = Iso::from7(mesh_pose7);
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s2 = make_shape(kinds, params4, points, shape2[k]);
            Iso p2 = Iso::from7(pos2 + 7 * k);
            Iso pos12 = pos1.inv_mul(p2);
            Contact c = Contact();
            uint32_t id = UINT32_MAX;
            int st = contact_trimesh_shape(pos12, *tm, s2, prediction, c, id, ties != 0);
            status[k] = (uint8_t)st; part[k] = st == CONTACT_SOME ? id : UINT32_MAX;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) {
                c.point1 = pos1.transform_point(c.point1); c.point2 = p2.transform_point(c.point2);
                c.normal1 = pos1.transform_vector(c.normal1); c.normal2 = p2.transform_vector(c.normal2);
                st3(o, c.point1); st3(o + 3, c.point2); st3(o + 6, c.normal1); st3(o + 9, c.normal2); o[12] = c.dist;
            } else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        }
    });
}
// Bvh::project_point with typed leaves (bvh_queries.rs:213-227; leaves as in pb2o_bvh_cast_rays_shapes2): the leaf check is
// PointQuery::project_point(pose, pt, solid) (point_query.rs:147-151) of a Ball (point_ball.rs:9-21), a Cuboid (point_cuboid.rs ->
// point_aabb.rs:9-60) or a ConvexPolyhedron (point_support_map.rs:17-52; EPA when inside and not solid), the cost na::distance of the
// world-space projection to the point. leaf = UINT32_MAX when nothing lies within max_distance.
struct ProjLeaf { Real d; Vec3 point; bool inside; Real cost() const { return d; } };
void pb2o_bvh_project_points_shapes(void* b, const uint8_t* kinds, const float* params, const float* points, const uint32_t* first,
                                    const uint32_t* count, const float* poses7, const float* pts, uint32_t m, float max_distance, int solid,
                                    int nthreads, float* proj, uint8_t* inside, uint32_t* leaf) {
    const Bvh* t = (const Bvh*)b;
    parallel_for(m, nthreads, [=](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i) {
            Vec3 pt = ld3(pts + 3 * i);
            uint32_t id = UINT32_MAX; ProjLeaf best; best.d = 0; best.inside = false;
            bool hit = t->find_best<ProjLeaf>(max_distance,
                [&](const BvhNode& n, Real) { return norm(vsup(vsup(n.mins - pt, pt - n.maxs), Vec3())); },   // Aabb::distance_to_local_point(pt, true)
                [&](uint32_t prim, Real, ProjLeaf& o) {
                    Iso pose = Iso::from7(poses7 + 7 * prim);
                    Vec3 lp = pose.inverse_transform_point(pt), q;
                    bool in;
                    if (kinds[prim] == 0) {
                        Real r = params[3 * prim], d2 = norm_squared(lp);
                        in = d2 <= r * r;
                        q = (in && solid) ? lp : lp * (r / sqrtf(d2));
                    } else {
                        ShapeRef s;
                        s.kind = kinds[prim] == 1 ? SHAPE_CUBOID : SHAPE_CONVEX; s.radius = 0; s.half_extents = ld3(params + 3 * prim);
                        s.points = kinds[prim] == 2 ? points + 3 * (size_t)first[prim] : nullptr; s.num_points = kinds[prim] == 2 ? count[prim] : 0;
                        if (solid) project_local_point_solid(s, lp, q, in);
                        else if (kinds[prim] == 1) { Feature f{3, 0}; cuboid_project_point_and_get_feature(s.half_extents, lp, q, in, f); }
                        else hull_project_point(s.support(), lp, q, in);
                    }
                    o.point = pose.transform_point(q); o.inside = in; o.d = norm(pt - o.point);
                    return true;
                }, id, best);
            if (!hit) { st3(proj + 3 * i, Vec3()); inside[i] = 0; leaf[i] = UINT32_MAX; continue; }
            st3(proj + 3 * i, best.point); inside[i] = best.inside ? 1 : 0; leaf[i] = id;
        }
    });
}
// query::distance with a TriMesh on one side (default_query_dispatcher.rs:288-297 -> distance_composite_shape_shape.rs:46-77).
// mesh_second != 0: distance(poses[k], shape, mesh_pose, &TriMesh) = distance_shape_composite_shape (pos12.inverse()).
void pb2o_trimesh_distance_batch(void* mesh, const float* mesh_pose7, const uint8_t* kinds, const float* params4, const float* points,
                                 const uint32_t* shape_ids, const float* poses, uint32_t n, int nthreads, int mesh_second, float* dist,
                                 uint32_t* part) {
    const TriMesh* tm = (const TriMesh*)mesh;
    Iso pm = Iso::from7(mesh_pose7);
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            ShapeRef s2 = make_shape(kinds, params4, points, shape_ids[k]);
            Iso ps = Iso::from7(poses + 7 * k);
            Iso pos12 = mesh_second ? ps.inv_mul(pm).inverse() : pm.inv_mul(ps);
            Real d; uint32_t id = UINT32_MAX;
            distance_trimesh_shape(pos12, *tm, s2, d, id);
            dist[k] = d; part[k] = id;
        }
    });
}
// query::contact between Compounds of a table and one TriMesh (oracle groundwork, no GPU path yet). trimesh_first == 0:
// contact(poses[k], Compound ids[k], mesh_pose, &TriMesh); != 0: contact(mesh_pose, &TriMesh, poses[k], Compound ids[k]). World-space
// result; parts[k] = {winning compound part, winning triangle} or u32::MAX. ties as in pb2o_trimesh_contact_batch.
void pb2o_compound_trimesh_contact_batch(void* mesh, const float* mesh_pose7, const uint8_t* kinds, const float* params4, const float* points,
                                         const uint32_t* comp_first, const uint32_t* comp_count, const uint32_t* part_shape, const float* part_pose7,
                                         const uint32_t* ids, const float* poses, float prediction, uint32_t n, int nthreads, int trimesh_first,
                                         int ties, float* out, uint8_t* status, uint32_t* parts) {
    const TriMesh* tm = (const TriMesh*)mesh;
    Iso pm = Iso::from7(mesh_pose7);
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        std::vector<ShapeRef> ss; std::vector<Iso> qq;
        for (size_t k = lo; k < hi; ++k) {
            uint32_t f = comp_first[ids[k]], cnt = comp_count[ids[k]];
            ss.resize(cnt); qq.resize(cnt);
            for (uint32_t i = 0; i < cnt; ++i) { ss[i] = make_shape(kinds, params4, points, part_shape[f + i]); qq[i] = Iso::from7(part_pose7 + 7 * (size_t)(f + i)); }
            CompoundRef comp{ss.data(), qq.data(), cnt};
            Iso pc = Iso::from7(poses + 7 * k);
            Iso p1 = trimesh_first ? pm : pc, p2 = trimesh_first ? pc : pm;
            Contact c = Contact(); uint32_t part = UINT32_MAX, tri = UINT32_MAX;
            int st = trimesh_first ? contact_trimesh_compound(p1.inv_mul(p2), *tm, comp, prediction, c, tri, part, ties != 0)
                                   : contact_compound_trimesh(p1.inv_mul(p2), comp, *tm, prediction, c, part, tri, ties != 0);
            status[k] = (uint8_t)st; parts[2 * k] = st == CONTACT_SOME ? part : UINT32_MAX; parts[2 * k + 1] = st == CONTACT_SOME ? tri : UINT32_MAX;
            float* o = out + 13 * k;
            if (st == CONTACT_SOME) {
                c.point1 = p1.transform_point(c.point1); c.point2 = p2.transform_point(c.point2);
                c.normal1 = p1.transform_vector(c.normal1); c.normal2 = p2.transform_vector(c.normal2);
                st3(o, c.point1); st3(o + 3, c.point2); st3(o + 6, c.normal1); st3(o + 9, c.normal2); o[12] = c.dist;
            } else for (int i = 0; i < 13; ++i) o[i] = 0.0f;
        }
    });
}
// PointQuery::project_point(m, pt, solid) / project_local_point on a TriMesh (query/point/point_query.rs:147-151: local projection
// of m^-1 * pt, transformed back). mode 0: reference traversal; mode 1: brute force, ties to the smallest triangle index.
void pb2o_trimesh_project_points(void* mesh, const float* pose7, const float* points, uint32_t n, int solid, int mode, int nthreads,
                                 float* proj, uint8_t* inside, uint32_t* tri) {
    const TriMesh* tm = (const TriMesh*)mesh;
    bool has_pose = pose7 != nullptr;
    Iso pose = has_pose ? Iso::from7(pose7) : Iso();
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            Vec3 pt = ld3(points + 3 * k);
            if (has_pose) pt = pose.inverse_transform_point(pt);
            uint32_t id = UINT32_MAX; Vec3 p; bool in = false;
            bool ok = trimesh_project_local_point(*tm, pt, solid != 0, mode == 1, id, p, in);
            if (!ok) { st3(proj + 3 * k, Vec3()); inside[k] = 0; tri[k] = UINT32_MAX; continue; }
            if (has_pose) p = pose.transform_point(p);
            st3(proj + 3 * k, p); inside[k] = in ? 1 : 0; tri[k] = id;
        }
    });
}
// query::distance / query::intersection_test for n pairs (distance.rs:89-97, intersection_test.rs:88-96). status: 0 Ok, 2 Unsupported
// (bad shape id), 3 cuboid-cuboid (SAT arm, not restated).
void pb2o_distance_batch(const uint8_t* kinds, const float* params4, const float* points, uint32_t n_shapes, const uint32_t* shape1,
                         const uint32_t* shape2, const float* pos1, const float* pos2, uint32_t n, int nthreads, float* out, uint8_t* status) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            out[k] = 0.0f;
            if (shape1[k] >= n_shapes || shape2[k] >= n_shapes) { status[k] = QUERY_UNSUPPORTED; continue; }
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Iso pos12 = Iso::from7(pos1 + 7 * k).inv_mul(Iso::from7(pos2 + 7 * k));
            Real d = 0;
            status[k] = (uint8_t)dispatch_distance(pos12, s1, s2, d);
            out[k] = d;
        }
    });
}
void pb2o_intersection_test_batch(const uint8_t* kinds, const float* params4, const float* points, uint32_t n_shapes, const uint32_t* shape1,
                                  const uint32_t* shape2, const float* pos1, const float* pos2, uint32_t n, int nthreads, uint8_t* out,
                                  uint8_t* status) {
    parallel_for(n, nthreads, [=](size_t lo, size_t hi) {
        for (size_t k = lo; k < hi; ++k) {
            out[k] = 0;
            if (shape1[k] >= n_shapes || shape2[k] >= n_shapes) { status[k] = QUERY_UNSUPPORTED; continue; }
            ShapeRef s1 = make_shape(kinds, params4, points, shape1[k]), s2 = make_shape(kinds, params4, points, shape2[k]);
            Iso pos12 = Iso::from7(pos1 + 7 * k).inv_mul(Iso::from7(pos2 + 7 * k));
            bool b = false;
            status[k] = (uint8_t)dispatch_intersection_test(pos12, s1, s2, b);
            out[k] = b ? 1 : 0;
        }
    });
}
// DefaultQueryDispatcher::contact(pos12, ...) — results in the local frames of shape 1 / shape 2.
int pb2o_dispatch_contact(const uint8_t* kinds, const float* params4, const float* points, uint32_t s1, uint32_t s2, const float* pos12, float prediction, float* out13) {
    ShapeRef a = make_shape(kinds, params4, points, s1), b = make_shape(kinds, params4, points, s2);
    Contact c = Contact();
    int st = dispatch_contact(Iso::from7(pos12), a, b, prediction, c);
    if (st == CONTACT_SOME) { st3(out13, c.point1); st3(out13 + 3, c.point2); st3(out13 + 6, c.normal1); st3(out13 + 9, c.normal2); out13[12] = c.dist; }
    return st;
}
// gjk::closest_points between two support shapes (epa3.rs tests use the ClosestPoints query); returns GJKResult kind,
// out = p1, p2 (both in shape-1 space), dir.
int pb2o_gjk_closest_points(const uint8_t* kinds, const float* params4, const float* points, uint32_t s1, uint32_t s2, const float* pos12, float max_dist, float* out9) {
    ShapeRef a = make_shape(kinds, params4, points, s1), b = make_shape(kinds, params4, points, s2);
    Iso p = Iso::from7(pos12);
    VoronoiSimplex simplex;
    Vec3 dir;
    if (!try_normalize(p.tra, DEFAULT_EPSILON, dir)) dir = Vec3(1, 0, 0);
    SupportShape g1 = a.support(), g2 = b.support();
    simplex.reset(CSOPoint::from_shapes(p, g1, g2, dir));
    GJKResult r = gjk_closest_points(p, g1, g2, max_dist, simplex);
    st3(out9, r.p1); st3(out9 + 3, r.p2); st3(out9 + 6, r.dir);
    return (int)r.kind;
}
}

// ORACLE — TEST INFRASTRUCTURE ONLY (see math.hpp header).
// Restatement of parry3d src/query/epa/epa3.rs (Expanding Polytope Algorithm, 3D) including the exact sift rules
// of Rust's std::collections::BinaryHeap (push = sift_up, pop = swap-with-last + sift_down_to_bottom + sift_up) so
// that exact distance ties pop in the same order as the reference.
#pragma once
#include "gjk.hpp"

namespace pb2o {

// epa3.rs:18-53
struct FaceId {
    size_t id;
    Real neg_dist;
};
static inline bool faceid_le(const FaceId& a, const FaceId& b) { return !(a.neg_dist > b.neg_dist); }  // Ord::cmp on neg_dist

// Rust alloc::collections::BinaryHeap<FaceId> (max-heap)
struct RustBinaryHeap {
    std::vector<FaceId> data;
    void clear() { data.clear(); }
    bool empty() const { return data.empty(); }
    const FaceId& peek() const { return data[0]; }
    size_t sift_up(size_t start, size_t pos) {
        FaceId elt = data[pos];
        while (pos > start) {
            size_t parent = (pos - 1) / 2;
            if (faceid_le(elt, data[parent])) break;
            data[pos] = data[parent];
            pos = parent;
        }
        data[pos] = elt;
        return pos;
    }
    void push(const FaceId& f) { size_t old = data.size(); data.push_back(f); sift_up(0, old); }
    void sift_down_to_bottom(size_t pos) {
        size_t end = data.size(), start = pos;
        FaceId elt = data[pos];
        size_t child = 2 * pos + 1;
        while (child <= (end >= 2 ? end - 2 : 0) && end >= 2) {
            if (faceid_le(data[child], data[child + 1])) child += 1;
            data[pos] = data[child];
            pos = child;
            child = 2 * pos + 1;
        }
        if (child == end - 1) { data[pos] = data[child]; pos = child; }
        data[pos] = elt;
        sift_up(start, pos);
    }
    FaceId pop() {
        FaceId item = data.back();
        data.pop_back();
        if (!data.empty()) { std::swap(item, data[0]); sift_down_to_bottom(0); }
        return item;
    }
};

// utils/ccw_face_normal.rs:21-27
static inline bool ccw_face_normal(const Vec3& a, const Vec3& b, const Vec3& c, Vec3& n) {
    Vec3 ab = b - a, ac = c - a;
    return try_normalize(cross(ab, ac), DEFAULT_EPSILON, n);
}
// shape/triangle.rs:534-540
static inline bool tri_is_affinely_dependent(const Vec3& a, const Vec3& b, const Vec3& c) {
    const Real EPS = DEFAULT_EPSILON * 100.0f;
    Vec3 p1p2 = b - a, p1p3 = c - a;
    return relative_eq(norm_squared(cross(p1p2, p1p3)), 0.0f, EPS * EPS, FLT_EPSILON);
}

struct EpaFace {
    size_t pts[3];
    size_t adj[3];
    Vec3 normal;
    Real bcoords[3];
    bool deleted;
};

struct EpaStats { int niter = 0; size_t max_faces = 0, max_vertices = 0, max_heap = 0, max_silhouette = 0; };

struct EPA {
    std::vector<CSOPoint> vertices;
    std::vector<EpaFace> faces;
    struct SilEdge { size_t face_id, opp_pt_id; };
    std::vector<SilEdge> silhouette;
    RustBinaryHeap heap;
    EpaStats stats;

    // epa3.rs:65-94
    EpaFace new_with_proj(const Real bc[3], const size_t pts[3], const size_t adj[3]) const {
        EpaFace f;
        Vec3 n;
        if (!ccw_face_normal(vertices[pts[0]].point, vertices[pts[1]].point, vertices[pts[2]].point, n)) n = Vec3();
        f.normal = n;
        for (int i = 0; i < 3; ++i) { f.pts[i] = pts[i]; f.adj[i] = adj[i]; f.bcoords[i] = bc[i]; }
        f.deleted = false;
        return f;
    }
    // epa3.rs:96-118
    EpaFace face_new(const size_t pts[3], const size_t adj[3], bool& proj_inside) const {
        TriProj p = project_on_triangle(vertices[pts[0]].point, vertices[pts[1]].point, vertices[pts[2]].point, Vec3(), true);
        Real bc[3];
        if (p.kind == 0 || p.kind == 1) {
            const Real eps_tol = DEFAULT_EPSILON * 100.0f;
            tri_barycentric(p, bc);
            // PointProjection::is_inside_eps (point_query.rs:94-96)
            proj_inside = p.inside || norm_squared(p.point - Vec3()) < eps_tol * eps_tol;
            return new_with_proj(bc, pts, adj);
        } else if (p.kind == 2) {
            proj_inside = true;
            return new_with_proj(p.bc, pts, adj);
        }
        bc[0] = bc[1] = bc[2] = 0;
        proj_inside = false;
        return new_with_proj(bc, pts, adj);
    }
    void face_closest_points(const EpaFace& f, Vec3& p1, Vec3& p2) const {  // epa3.rs:120-131
        p1 = vertices[f.pts[0]].orig1 * f.bcoords[0] + vertices[f.pts[1]].orig1 * f.bcoords[1] + vertices[f.pts[2]].orig1 * f.bcoords[2];
        p2 = vertices[f.pts[0]].orig2 * f.bcoords[0] + vertices[f.pts[1]].orig2 * f.bcoords[1] + vertices[f.pts[2]].orig2 * f.bcoords[2];
    }
    static size_t next_ccw_pt_id(const EpaFace& f, size_t id) {  // epa3.rs:137-151
        if (f.pts[0] == id) return 1;
        if (f.pts[1] == id) return 2;
        return 0;
    }
    bool can_be_seen_by(const EpaFace& f, size_t point, size_t opp) const {  // epa3.rs:153-166
        const Vec3& p0 = vertices[f.pts[opp]].point;
        const Vec3& p1 = vertices[f.pts[(opp + 1) % 3]].point;
        const Vec3& p2 = vertices[f.pts[(opp + 2) % 3]].point;
        const Vec3& pt = vertices[point].point;
        return dot(pt - p0, f.normal) >= -gjk_eps_tol() || tri_is_affinely_dependent(p1, p2, pt);
    }
    void compute_silhouette(size_t point, size_t id, size_t opp) {  // epa3.rs:653-675 (recursive, like the reference)
        if (!faces[id].deleted) {
            if (!can_be_seen_by(faces[id], point, opp)) {
                silhouette.push_back(SilEdge{id, opp});
            } else {
                faces[id].deleted = true;
                size_t adj_pt_id1 = (opp + 2) % 3, adj_pt_id2 = opp;
                size_t adj1 = faces[id].adj[adj_pt_id1], adj2 = faces[id].adj[adj_pt_id2];
                size_t o1 = next_ccw_pt_id(faces[adj1], faces[id].pts[adj_pt_id1]);
                size_t o2 = next_ccw_pt_id(faces[adj2], faces[id].pts[adj_pt_id2]);
                compute_silhouette(point, adj1, o1);
                compute_silhouette(point, adj2, o2);
            }
        }
    }

    // epa3.rs:428-651. Returns false for None.
    bool closest_points(const Iso& pos12, const SupportShape& g1, const SupportShape& g2, const VoronoiSimplex& simplex,
                        Vec3& out_p1, Vec3& out_p2, Vec3& out_n) {
        const Real eps = DEFAULT_EPSILON;
        const Real eps_tol = eps * 100.0f;
        vertices.clear(); faces.clear(); heap.clear(); silhouette.clear();
        stats = EpaStats();
        for (size_t i = 0; i < simplex.dimension() + 1; ++i) vertices.push_back(simplex.point(i));

        auto push_face = [&](size_t id, Real neg_dist) -> bool {  // FaceId::new(...)?  (epa3.rs:24-30)
            if (neg_dist > gjk_eps_tol()) return false;
            heap.push(FaceId{id, neg_dist});
            return true;
        };

        if (simplex.dimension() == 0) {
            out_p1 = Vec3(); out_p2 = Vec3(); out_n = Vec3(0, 1, 0);
            return true;
        } else if (simplex.dimension() == 3) {
            Vec3 dp1 = vertices[1].point - vertices[0].point;
            Vec3 dp2 = vertices[2].point - vertices[0].point;
            Vec3 dp3 = vertices[3].point - vertices[0].point;
            if (dot(cross(dp1, dp2), dp3) > 0.0f) std::swap(vertices[1], vertices[2]);
            const size_t pts[4][3] = {{0, 1, 2}, {1, 3, 2}, {0, 2, 3}, {0, 3, 1}};
            const size_t adj[4][3] = {{3, 1, 2}, {3, 2, 0}, {0, 1, 3}, {2, 1, 0}};
            bool inside[4];
            for (int k = 0; k < 4; ++k) faces.push_back(face_new(pts[k], adj[k], inside[k]));
            for (int k = 0; k < 4; ++k) {
                if (inside[k]) {
                    Real dist = dot(faces[k].normal, vertices[k].point);
                    if (!push_face(k, -dist)) return false;
                }
            }
            if (!(inside[0] || inside[1] || inside[2] || inside[3])) return false;
        } else {
            if (simplex.dimension() == 1) {
                Vec3 dpt = vertices[1].point - vertices[0].point;
                // Vector3::orthonormal_subspace_basis(&[dpt], f) with f returning false: exactly one direction a x v
                Vec3 a = fabsf(dpt.x) > fabsf(dpt.y) ? Vec3(dpt.z, 0.0f, -dpt.x) : Vec3(0.0f, -dpt.z, dpt.y);
                a = normalize(a);
                Vec3 dir = cross(a, dpt);
                vertices.push_back(CSOPoint::from_shapes(pos12, g1, g2, dir));
            }
            const size_t pts1[3] = {0, 1, 2}, pts2[3] = {0, 2, 1}, adj1[3] = {1, 1, 1}, adj2[3] = {0, 0, 0};
            bool dummy;
            faces.push_back(face_new(pts1, adj1, dummy));
            faces.push_back(face_new(pts2, adj2, dummy));
            if (!push_face(0, 0.0f)) return false;
            if (!push_face(1, 0.0f)) return false;
        }

        int niter = 0;
        Real max_dist = REAL_MAX;
        if (heap.empty()) return false;  // *self.heap.peek()?
        FaceId best_face_id = heap.peek();
        Real old_dist = 0.0f;

        while (!heap.empty()) {
            stats.max_heap = std::max(stats.max_heap, heap.data.size());
            FaceId face_id = heap.pop();
            EpaFace face = faces[face_id.id];
            if (face.deleted) continue;
            CSOPoint cso_point = CSOPoint::from_shapes(pos12, g1, g2, face.normal);
            size_t support_point_id = vertices.size();
            vertices.push_back(cso_point);
            Real candidate_max_dist = dot(cso_point.point, face.normal);
            if (candidate_max_dist < max_dist) { best_face_id = face_id; max_dist = candidate_max_dist; }
            Real curr_dist = -face_id.neg_dist;
            if (max_dist - curr_dist < eps_tol || (fabsf(curr_dist - old_dist) < eps && candidate_max_dist < max_dist)) {
                const EpaFace& best = faces[best_face_id.id];
                face_closest_points(best, out_p1, out_p2); out_n = best.normal;
                finish(niter);
                return true;
            }
            old_dist = curr_dist;
            faces[face_id.id].deleted = true;
            size_t o1 = next_ccw_pt_id(faces[face.adj[0]], face.pts[0]);
            size_t o2 = next_ccw_pt_id(faces[face.adj[1]], face.pts[1]);
            size_t o3 = next_ccw_pt_id(faces[face.adj[2]], face.pts[2]);
            compute_silhouette(support_point_id, face.adj[0], o1);
            compute_silhouette(support_point_id, face.adj[1], o2);
            compute_silhouette(support_point_id, face.adj[2], o3);
            size_t first_new_face_id = faces.size();
            if (silhouette.empty()) return false;
            stats.max_silhouette = std::max(stats.max_silhouette, silhouette.size());
            for (const SilEdge& edge : silhouette) {
                if (!faces[edge.face_id].deleted) {
                    size_t new_face_id = faces.size();
                    size_t pt_id1 = faces[edge.face_id].pts[(edge.opp_pt_id + 2) % 3];
                    size_t pt_id2 = faces[edge.face_id].pts[(edge.opp_pt_id + 1) % 3];
                    size_t pts[3] = {pt_id1, pt_id2, support_point_id};
                    size_t adj[3] = {edge.face_id, new_face_id + 1, new_face_id - 1};
                    bool inside;
                    EpaFace nf = face_new(pts, adj, inside);
                    faces[edge.face_id].adj[(edge.opp_pt_id + 1) % 3] = new_face_id;
                    faces.push_back(nf);
                    if (inside) {
                        Vec3 pt = vertices[faces[new_face_id].pts[0]].point;
                        Real dist = dot(faces[new_face_id].normal, pt);
                        if (dist < curr_dist) {
                            face_closest_points(face, out_p1, out_p2); out_n = face.normal;
                            finish(niter);
                            return true;
                        }
                        if (!push_face(new_face_id, -dist)) return false;
                    }
                }
            }
            if (first_new_face_id == faces.size()) return false;
            faces[first_new_face_id].adj[2] = faces.size() - 1;
            faces.back().adj[1] = first_new_face_id;
            silhouette.clear();
            niter += 1;
            if (niter > 100) break;
        }
        const EpaFace& best = faces[best_face_id.id];
        face_closest_points(best, out_p1, out_p2); out_n = best.normal;
        finish(niter);
        return true;
    }
    void finish(int niter) { stats.niter = niter; stats.max_faces = faces.size(); stats.max_vertices = vertices.size(); }
};

}  // namespace pb2o

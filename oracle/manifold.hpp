// ORACLE — TEST INFRASTRUCTURE ONLY (parity unpinned: the reference holds no known-answer test for these manifolds;
// tests/test_oracle_kats.py checks the restatement through geometric properties and against query::contact).
// Restatement of parry3d contact manifolds for the closed-form arms of DefaultQueryDispatcher::contact_manifold_convex_convex
// (default_query_dispatcher.rs:748-835): contact_manifolds_ball_ball.rs:17-57, contact_manifolds_convex_ball.rs:42-145
// (Cuboid as shape1), contact_manifolds_cuboid_cuboid.rs:19-107 + sat_cuboid_cuboid.rs:5-110 + Cuboid::support_face
// (cuboid.rs:267-354) + PolygonalFeature::contacts_face_face / closest_points_line2d (polygonal_feature3d.rs:215-439) +
// Vector3::orthonormal_basis (utils/wops.rs:92-110). First-frame behaviour: the manifold starts empty, so
// try_update_contacts (contact_manifold.rs) returns false and match_contacts has nothing to transfer.
#pragma once
#include "contact.hpp"

namespace pb2o {

// PackedFeatureId (shape/feature_id.rs:446-520)
static inline uint32_t packed_vertex(uint32_t c) { return (1u << 30) | c; }
static inline uint32_t packed_edge(uint32_t c) { return (2u << 30) | c; }
static inline uint32_t packed_face(uint32_t c) { return (3u << 30) | c; }
static inline uint32_t packed_from_feature(const Feature& f) {
    if (f.kind == 0) return packed_vertex(f.id);
    if (f.kind == 1) return packed_edge(f.id);
    if (f.kind == 2) return packed_face(f.id);
    return 0u;  // UNKNOWN
}

struct TrackedContact { Vec3 local_p1, local_p2; Real dist; uint32_t fid1, fid2; };
struct Manifold {
    Vec3 local_n1, local_n2;
    std::vector<TrackedContact> points;
    void clear() { points.clear(); }
    void push_flipped(const Vec3& p1, const Vec3& p2, uint32_t f1, uint32_t f2, Real dist, bool flipped) {
        TrackedContact t;
        if (!flipped) { t.local_p1 = p1; t.local_p2 = p2; t.fid1 = f1; t.fid2 = f2; }
        else { t.local_p1 = p2; t.local_p2 = p1; t.fid1 = f2; t.fid2 = f1; }
        t.dist = dist;
        points.push_back(t);
    }
};

// contact_manifolds_ball_ball.rs:17-57
static inline void manifold_ball_ball(const Iso& pos12, Real ra, Real rb, Real prediction, Manifold& m) {
    Vec3 dcenter = pos12.tra;
    Real center_dist = norm(dcenter);
    Real dist = center_dist - ra - rb;
    m.clear();
    if (dist < prediction) {
        Vec3 n1 = center_dist != 0.0f ? dcenter / center_dist : Vec3(0, 1, 0);
        Vec3 n2 = pos12.inverse_transform_vector(-n1);
        m.push_flipped(n1 * ra, n2 * rb, packed_face(0), packed_face(0), dist, false);
        m.local_n1 = n1; m.local_n2 = n2;
    }
}

// contact_manifolds_convex_ball.rs:42-145 with shape1 = Cuboid, no normal constraints. pos12 is already the pose of the ball
// in the cuboid's frame (the caller inverts it when the ball is shape 1, :18-28).
static inline void manifold_cuboid_ball(const Iso& pos12, const Vec3& he, Real radius, Real prediction, bool flipped, Manifold& m) {
    Vec3 local_p2_1 = pos12.tra;
    Vec3 proj; bool inside; Feature f;
    cuboid_project_point_and_get_feature(he, local_p2_1, proj, inside, f);
    Vec3 dpos = local_p2_1 - proj;
    Vec3 n1; Real dist;
    if (!try_normalize_and_get(dpos, 0.0f, n1, dist)) {
        if (!try_normalize(pos12.tra, 0.0f, n1)) n1 = Vec3(1, 0, 0);
        dist = 0.0f;
    }
    if (inside) { n1 = -n1; dist = -dist; }
    m.clear();
    if (dist <= radius + prediction) {
        Vec3 n2 = pos12.inverse_transform_vector(-n1);
        Vec3 local_p2 = n2 * radius;
        m.push_flipped(proj, local_p2, packed_from_feature(f), packed_face(0), dist - radius, flipped);
        if (flipped) { m.local_n1 = n2; m.local_n2 = n1; } else { m.local_n1 = n1; m.local_n2 = n2; }
    }
}

// ---- SAT (sat_cuboid_cuboid.rs): in contact.hpp (shared with the cuboid-cuboid distance / intersection_test arms)

// ---- Cuboid::support_face (cuboid.rs:267-354)
struct PolyFeature { Vec3 v[4]; uint32_t vids[4], eids[4], fid; int n; };
static inline PolyFeature cuboid_support_face(const Vec3& he, const Vec3& dir) {
    int iamax = 0;   // nalgebra iamax: first strict maximum of |x|
    { Real best = fabsf(dir.x); if (fabsf(dir.y) > best) { best = fabsf(dir.y); iamax = 1; } if (fabsf(dir.z) > best) iamax = 2; }
    Real sign = copysignf(1.0f, dir[iamax]);
    PolyFeature f; f.n = 4;
    if (iamax == 0) { f.v[0] = Vec3(he.x * sign, he.y, he.z); f.v[1] = Vec3(he.x * sign, -he.y, he.z); f.v[2] = Vec3(he.x * sign, -he.y, -he.z); f.v[3] = Vec3(he.x * sign, he.y, -he.z); }
    else if (iamax == 1) { f.v[0] = Vec3(he.x, he.y * sign, he.z); f.v[1] = Vec3(-he.x, he.y * sign, he.z); f.v[2] = Vec3(-he.x, he.y * sign, -he.z); f.v[3] = Vec3(he.x, he.y * sign, -he.z); }
    else { f.v[0] = Vec3(he.x, he.y, he.z * sign); f.v[1] = Vec3(he.x, -he.y, he.z * sign); f.v[2] = Vec3(-he.x, -he.y, he.z * sign); f.v[3] = Vec3(-he.x, he.y, he.z * sign); }
    int si = ((int)(int8_t)sign + 1) / 2;   // sign_index: -1 -> 0, +1 -> 1; rows below are the reference's literals, [axis][sign_index]
    static const uint32_t VIDS[3][2][4] = {{{0, 2, 3, 1}, {4, 6, 7, 5}}, {{0, 4, 5, 1}, {2, 6, 7, 3}}, {{0, 2, 6, 4}, {1, 3, 7, 5}}};
    static const uint32_t EIDS[3][2][4] = {{{0xD0, 0xDA, 0xD9, 0xC8}, {0xF4, 0xFE, 0xFD, 0xEC}},
                                          {{0xE0, 0xEC, 0xE9, 0xC8}, {0xF2, 0xFE, 0xFB, 0xDA}},
                                          {{0xD0, 0xF2, 0xF4, 0xE0}, {0xD9, 0xFB, 0xFD, 0xE9}}};
    for (int k = 0; k < 4; ++k) { f.vids[k] = packed_vertex(VIDS[iamax][si][k] * 2u); f.eids[k] = packed_edge(EIDS[iamax][si][k]); }
    f.fid = packed_face((uint32_t)(iamax + si * 3 + 10));
    return f;
}

static inline Real perp2(Real ax, Real ay, Real bx, Real by) { return ax * by - ay * bx; }
// polygonal_feature3d.rs:398-439
static inline bool closest_points_line2d(const Real e1[2][2], const Real e2[2][2], Real& s_out, Real& t_out) {
    Real d1x = e1[1][0] - e1[0][0], d1y = e1[1][1] - e1[0][1];
    Real d2x = e2[1][0] - e2[0][0], d2y = e2[1][1] - e2[0][1];
    Real rx = e1[0][0] - e2[0][0], ry = e1[0][1] - e2[0][1];
    Real a = d1x * d1x + d1y * d1y, e = d2x * d2x + d2y * d2y, f = d2x * rx + d2y * ry;
    const Real eps = FLT_EPSILON;
    if (a <= eps && e <= eps) { s_out = 0; t_out = 0; return true; }
    if (a <= eps) { s_out = 0; t_out = f / e; return true; }
    Real c = d1x * rx + d1y * ry;
    if (e <= eps) { s_out = -c / a; t_out = 0; return true; }
    Real b = d1x * d2x + d1y * d2y;
    Real ae = a * e, bb = b * b, denom = ae - bb;
    bool parallel = denom <= eps || ulps_eq(ae, bb);
    if (parallel) return false;
    Real s = (b * f - c * e) / denom;
    s_out = s; t_out = (b * s + f) / e;
    return true;
}

// PolygonalFeature::contacts_face_face (polygonal_feature3d.rs:215-396)
static inline void contacts_face_face(const Iso& pos12, const PolyFeature& face1, const Vec3& sep_axis1, const PolyFeature& face2, Manifold& m, bool flipped) {
    // Vector3::orthonormal_basis (wops.rs:96-109)
    Real sign = copysignf(1.0f, sep_axis1.z);
    Real a = -1.0f / (sign + sep_axis1.z);
    Real b = sep_axis1.x * sep_axis1.y * a;
    Vec3 b0(1.0f + sign * sep_axis1.x * sep_axis1.x * a, sign * b, -sign * sep_axis1.x);
    Vec3 b1(b, sign + sep_axis1.y * sep_axis1.y * a, -sep_axis1.y);
    Real pf1[4][2], pf2[4][2];
    Vec3 v21[4];
    for (int i = 0; i < 4; ++i) { pf1[i][0] = dot(face1.v[i], b0); pf1[i][1] = dot(face1.v[i], b1); }
    for (int i = 0; i < 4; ++i) { v21[i] = pos12.transform_point(face2.v[i]); pf2[i][0] = dot(v21[i], b0); pf2[i][1] = dot(v21[i], b1); }
    if (face2.n > 2) {
        Vec3 normal2_1 = cross(v21[2] - v21[1], v21[0] - v21[1]);
        Real denom = dot(normal2_1, sep_axis1);
        if (!relative_eq(denom, 0.0f)) {
            int last2 = face2.n - 1;
            for (int i = 0; i < face1.n; ++i) {
                Real px = pf1[i][0], py = pf1[i][1];
                Real sg = perp2(pf2[0][0] - pf2[last2][0], pf2[0][1] - pf2[last2][1], px - pf2[last2][0], py - pf2[last2][1]);
                bool outside = false;
                for (int j = 0; j < last2; ++j) {
                    Real ns = perp2(pf2[j + 1][0] - pf2[j][0], pf2[j + 1][1] - pf2[j][1], px - pf2[j][0], py - pf2[j][1]);
                    if (sg == 0.0f) sg = ns;
                    else if (sg * ns < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                Real dist = dot(v21[0] - face1.v[i], normal2_1) / denom;
                Vec3 local_p1 = face1.v[i];
                Vec3 local_p2_1 = face1.v[i] + sep_axis1 * dist;
                m.push_flipped(local_p1, pos12.inverse_transform_point(local_p2_1), face1.vids[i], face2.fid, dist, flipped);
            }
        }
    }
    if (face1.n > 2) {
        Vec3 normal1 = cross(face1.v[2] - face1.v[1], face1.v[0] - face1.v[1]);
        Real denom = -dot(normal1, sep_axis1);
        if (!relative_eq(denom, 0.0f)) {
            int last1 = face1.n - 1;
            for (int i = 0; i < face2.n; ++i) {
                Real px = pf2[i][0], py = pf2[i][1];
                Real sg = perp2(pf1[0][0] - pf1[last1][0], pf1[0][1] - pf1[last1][1], px - pf1[last1][0], py - pf1[last1][1]);
                bool outside = false;
                for (int j = 0; j < last1; ++j) {
                    Real ns = perp2(pf1[j + 1][0] - pf1[j][0], pf1[j + 1][1] - pf1[j][1], px - pf1[j][0], py - pf1[j][1]);
                    if (sg == 0.0f) sg = ns;
                    else if (sg * ns < 0.0f) { outside = true; break; }
                }
                if (outside) continue;
                Real dist = dot(face1.v[0] - v21[i], normal1) / denom;
                Vec3 local_p2_1 = v21[i];
                Vec3 local_p1 = v21[i] - sep_axis1 * dist;
                m.push_flipped(local_p1, pos12.inverse_transform_point(local_p2_1), face1.fid, face2.vids[i], dist, flipped);
            }
        }
    }
    for (int j = 0; j < face2.n; ++j) {
        Real e2[2][2] = {{pf2[j][0], pf2[j][1]}, {pf2[(j + 1) % face2.n][0], pf2[(j + 1) % face2.n][1]}};
        for (int i = 0; i < face1.n; ++i) {
            Real e1[2][2] = {{pf1[i][0], pf1[i][1]}, {pf1[(i + 1) % face1.n][0], pf1[(i + 1) % face1.n][1]}};
            Real s, t;
            if (!closest_points_line2d(e1, e2, s, t)) continue;
            if (s > 0.0f && s < 1.0f && t > 0.0f && t < 1.0f) {
                Vec3 a0 = face1.v[i], a1 = face1.v[(i + 1) % face1.n];
                Vec3 c0 = v21[j], c1 = v21[(j + 1) % face2.n];
                Vec3 local_p1 = a0 * (1.0f - s) + a1 * s;
                Vec3 local_p2_1 = c0 * (1.0f - t) + c1 * t;
                Real dist = dot(local_p2_1 - local_p1, sep_axis1);
                m.push_flipped(local_p1, pos12.inverse_transform_point(local_p2_1), face1.eids[i], face2.eids[j], dist, flipped);
            }
        }
    }
}

// contact_manifolds_cuboid_cuboid.rs:19-107 (empty incoming manifold)
static inline void manifold_cuboid_cuboid(const Iso& pos12, const Vec3& he1, const Vec3& he2, Real prediction, Manifold& m) {
    m.clear();
    Iso pos21 = pos12.inverse();
    Real s1, s2, s3; Vec3 d1, d2, d3;
    sat_find_separating_normal_oneway(he1, he2, pos12, s1, d1);
    if (s1 > prediction) return;
    sat_find_separating_normal_oneway(he2, he1, pos21, s2, d2);
    if (s2 > prediction) return;
    sat_find_separating_edge_twoway(he1, he2, pos12, s3, d3);
    if (s3 > prediction) return;
    Real best_s = s1; Vec3 best_d = d1;
    if (s2 > s1 && s2 > s3) { best_s = s2; best_d = pos12.transform_vector(-d2); }
    else if (s3 > s1) { best_s = s3; best_d = d3; }
    (void)best_s;
    Vec3 local_n2 = pos21.transform_vector(-best_d);
    PolyFeature f1 = cuboid_support_face(he1, best_d), f2 = cuboid_support_face(he2, local_n2);
    contacts_face_face(pos12, f1, best_d, f2, m, false);
    m.local_n1 = best_d; m.local_n2 = local_n2;
}

// ---- PolygonalFeatureMap (shape/polygonal_feature_map.rs) for Cuboid (cuboid.rs: support_face) and ConvexPolyhedron
// (convex_polyhedron.rs:959-991: the face whose normal has the first maximal dot with dir, its first <= 4 vertices). The hull's
// face topology (ConvexPolyhedron::from_convex_mesh, :390-637) is an INPUT: parry builds it when the shape is created.
struct HullTopology {
    const float* face_normal;      // nf x 3
    const uint32_t* face_first;    // into the two adjacency arrays
    const uint32_t* face_count;
    const uint32_t* vertices_adj_to_face;   // vertex ids local to the hull
    const uint32_t* edges_adj_to_face;
    uint32_t num_faces;
    // vertex side, for support_feature_id_toward (NULL when not supplied): per vertex of THIS hull
    const uint32_t* vert_first = nullptr;            // into the two arrays below
    const uint32_t* vert_count = nullptr;
    const uint32_t* faces_adj_to_vertex = nullptr;   // face ids local to the hull
    const uint32_t* edges_adj_to_vertex = nullptr;   // edge ids local to the hull
    const float* edge_dir = nullptr;                 // this hull's edges, ne x 3
};
static inline PolyFeature hull_local_support_feature(const ShapeRef& s, const HullTopology& t, const Vec3& dir) {
    uint32_t best = 0;
    Real best_dot = dot(ld3(t.face_normal), dir);
    for (uint32_t f = 1; f < t.num_faces; ++f) {
        Real d = dot(ld3(t.face_normal + 3 * f), dir);
        if (d > best_dot) { best = f; best_dot = d; }
    }
    PolyFeature out;
    for (int i = 0; i < 4; ++i) { out.v[i] = Vec3(); out.vids[i] = 0; out.eids[i] = 0; }
    uint32_t i1 = t.face_first[best], nv = t.face_count[best] < 4 ? t.face_count[best] : 4;
    for (uint32_t i = 0; i < nv; ++i) {
        uint32_t vid = t.vertices_adj_to_face[i1 + i];
        out.v[i] = ld3(s.points + 3 * vid);
        out.vids[i] = packed_vertex(vid);
        out.eids[i] = packed_edge(t.edges_adj_to_face[i1 + i]);
    }
    out.fid = packed_face(best);
    out.n = (int)nv;
    return out;
}

// ConvexPolyhedron::support_feature_id_toward (convex_polyhedron.rs:885-922), eps = 1 degree: faces adjacent to the support
// vertex whose normal is within eps of dir, then adjacent edges perpendicular to dir within eps, else the vertex.
#define PB2O_SIN_1DEG 0.017452406f   // (PI / 180 as f32).sin_cos()
#define PB2O_COS_1DEG 0.99984770f
static inline Feature hull_support_feature_id_toward(const ShapeRef& s, const HullTopology& t, const Vec3& dir) {
    uint32_t best = 0;
    Real best_dot = dot(ld3(s.points), dir);
    for (uint32_t i = 1; i < s.num_points; ++i) { Real d = dot(ld3(s.points + 3 * i), dir); if (d > best_dot) { best_dot = d; best = i; } }
    uint32_t first = t.vert_first[best], cnt = t.vert_count[best];
    for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t f = t.faces_adj_to_vertex[first + i];
        if (dot(ld3(t.face_normal + 3 * f), dir) >= PB2O_COS_1DEG) return Feature{2, f};
    }
    for (uint32_t i = 0; i < cnt; ++i) {
        uint32_t e = t.edges_adj_to_vertex[first + i];
        if (fabsf(dot(ld3(t.edge_dir + 3 * e), dir)) <= PB2O_SIN_1DEG) return Feature{1, e};
    }
    return Feature{0, best};
}

// contact_manifolds_convex_ball.rs:42-145 with shape1 = ConvexPolyhedron (project_local_point_and_get_feature:
// point_support_map.rs:62-76), no normal constraints; pos12 = pose of the ball in the hull's frame.
static inline void manifold_hull_ball(const Iso& pos12, const ShapeRef& hull, const HullTopology& t, Real radius, Real prediction, bool flipped,
                                      Manifold& m) {
    Vec3 local_p2_1 = pos12.tra;
    Vec3 proj; bool inside;
    hull_project_point(hull.support(), local_p2_1, proj, inside);
    Vec3 dpt = local_p2_1 - proj;
    Vec3 local_dir = inside ? -dpt : dpt, ud;
    Feature f{3, 0};
    if (try_normalize(local_dir, DEFAULT_EPSILON, ud)) f = hull_support_feature_id_toward(hull, t, ud);
    Vec3 n1; Real dist;
    if (!try_normalize_and_get(dpt, 0.0f, n1, dist)) {
        if (!try_normalize(pos12.tra, 0.0f, n1)) n1 = Vec3(1, 0, 0);
        dist = 0.0f;
    }
    if (inside) { n1 = -n1; dist = -dist; }
    m.clear();
    if (dist <= radius + prediction) {
        Vec3 n2 = pos12.inverse_transform_vector(-n1);
        m.push_flipped(proj, n2 * radius, packed_from_feature(f), packed_face(0), dist - radius, flipped);
        if (flipped) { m.local_n1 = n2; m.local_n2 = n1; } else { m.local_n1 = n1; m.local_n2 = n2; }
    }
}

// contact_manifolds_pfm_pfm.rs:42-162, empty incoming manifold (init_dir = None), no normal constraints, border radii 0.
// Returns false when the GJK/EPA contact is not ClosestPoints (manifold stays empty).
static inline bool manifold_pfm_pfm(const Iso& pos12, const ShapeRef& s1, const HullTopology* t1, const ShapeRef& s2, const HullTopology* t2,
                                    Real prediction, Manifold& m, const Vec3* init_dir = nullptr) {
    m.clear();
    Contact c;
    Vec3 noint;
    if (contact_support_map_support_map(pos12, s1.support(), s2.support(), prediction, c, nullptr, init_dir, &noint) != CONTACT_SOME) {
        // GJKResult::NoIntersection(dir) => manifold.local_n1 = *dir ("use the manifold normal as a cache", :151-154). The reference
        // leaves local_n2 as it was; nothing reads it on an empty manifold, it is recorded as zero here and on the GPU.
        m.local_n1 = noint; m.local_n2 = Vec3();
        return false;
    }
    // c: point1 = p1, point2 = pos12^-1 p2_1, normal1 = dir, normal2 = pos12^-1 (-dir), dist = (p2_1 - p1) . dir
    Vec3 local_n1 = c.normal1, local_n2 = c.normal2;
    PolyFeature f1 = s1.kind == SHAPE_CUBOID ? cuboid_support_face(s1.half_extents, local_n1) : hull_local_support_feature(s1, *t1, local_n1);
    PolyFeature f2 = s2.kind == SHAPE_CUBOID ? cuboid_support_face(s2.half_extents, local_n2) : hull_local_support_feature(s2, *t2, local_n2);
    contacts_face_face(pos12, f1, local_n1, f2, m, false);   // PolygonalFeature::contacts: faces have 3 or 4 vertices here
    m.push_flipped(c.point1, c.point2, 0u, 0u, c.dist, false);   // the GJK/EPA witness pair itself (:108-122), PackedFeatureId::UNKNOWN
    m.local_n1 = local_n1; m.local_n2 = local_n2;
    return true;
}

// ContactManifold::try_update_contacts[_eps] (contact_manifold.rs:652-699): keep last frame's manifold under the new pos12 if its
// normal turned by less than the angle threshold and every point, re-projected along local_n1, stayed within sqrt(dist_sq_threshold)
// of where it was without switching between penetrating and separated; dists and local_p1 are refreshed. Points already
// updated stay updated when a later point rejects the manifold (the reference mutates in place, then recomputes everything).
#define PB2O_COS_1_DEGREES 0.99984769515f   // utils::COS_1_DEGREES
static inline bool manifold_try_update_contacts(Manifold& m, const Iso& pos12, Real angle_dot_threshold = PB2O_COS_1_DEGREES,
                                                Real dist_sq_threshold = 1.0e-6f) {
    if (m.points.empty()) return false;
    Vec3 local_n2 = pos12.transform_vector(m.local_n2);
    if (-dot(m.local_n1, local_n2) < angle_dot_threshold) return false;
    for (auto& pt : m.points) {
        Vec3 local_p2 = pos12.transform_point(pt.local_p2);
        Vec3 dpt = local_p2 - pt.local_p1;
        Real dist = dot(dpt, m.local_n1);
        if (dist * pt.dist < 0.0f) return false;
        Vec3 new_local_p1 = local_p2 - m.local_n1 * dist;
        Vec3 dd = pt.local_p1 - new_local_p1;
        if (norm_squared(dd) > dist_sq_threshold) return false;
        pt.dist = dist;
        pt.local_p1 = new_local_p1;
    }
    return true;
}

enum ManifoldStatus { MANIFOLD_OK = 0, MANIFOLD_UNSUPPORTED = 2 };
// DefaultQueryDispatcher::contact_manifold_convex_convex arms for Ball / Cuboid (default_query_dispatcher.rs:760-782); pairs with
// a ConvexPolyhedron go through pfm_pfm when the hull's face topology is supplied, and are reported unsupported otherwise.
// `m` carries last frame's manifold when persistent != 0 (QueryDispatcher::contact_manifolds is called with the same manifold
// object every frame): the cuboid-cuboid and pfm_pfm arms first try to keep it (contact_manifolds_cuboid_cuboid.rs:30,
// contact_manifolds_pfm_pfm.rs:64); *kept = true when they did. The ball arms always recompute.
static inline int dispatch_manifold(const Iso& pos12, const ShapeRef& s1, const ShapeRef& s2, Real prediction, Manifold& m,
                                    const HullTopology* t1 = nullptr, const HullTopology* t2 = nullptr, bool persistent = false,
                                    bool* kept = nullptr, bool seed_gjk = false) {
    if (kept) *kept = false;
    Vec3 seed; bool have_seed = false;
    if (persistent && s1.kind != SHAPE_BALL && s2.kind != SHAPE_BALL) {
        bool pfm_ok = (s1.kind == SHAPE_CUBOID || t1) && (s2.kind == SHAPE_CUBOID || t2);
        if (pfm_ok && manifold_try_update_contacts(m, pos12)) { if (kept) *kept = true; return MANIFOLD_OK; }
        // contact_manifolds_pfm_pfm.rs:66: init_dir = Unit::try_new(manifold.local_n1, DEFAULT_EPSILON) seeds the GJK of the recomputation
        // (seed_gjk = false restarts GJK from the default direction: kept to measure what the seed changes)
        have_seed = seed_gjk && try_normalize(m.local_n1, DEFAULT_EPSILON, seed);
    }
    m.clear(); m.local_n1 = Vec3(); m.local_n2 = Vec3();
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_BALL) { manifold_ball_ball(pos12, s1.radius, s2.radius, prediction, m); return MANIFOLD_OK; }
    if (s1.kind == SHAPE_CUBOID && s2.kind == SHAPE_CUBOID) { manifold_cuboid_cuboid(pos12, s1.half_extents, s2.half_extents, prediction, m); return MANIFOLD_OK; }
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_CUBOID) { manifold_cuboid_ball(pos12.inverse(), s2.half_extents, s1.radius, prediction, true, m); return MANIFOLD_OK; }
    if (s1.kind == SHAPE_CUBOID && s2.kind == SHAPE_BALL) { manifold_cuboid_ball(pos12, s1.half_extents, s2.radius, prediction, false, m); return MANIFOLD_OK; }
    // (_, Ball) | (Ball, _) with a ConvexPolyhedron: contact_manifold_convex_ball, needs the vertex side of the topology
    if (s1.kind == SHAPE_BALL && s2.kind == SHAPE_CONVEX && t2 && t2->vert_first) { manifold_hull_ball(pos12.inverse(), s2, *t2, s1.radius, prediction, true, m); return MANIFOLD_OK; }
    if (s1.kind == SHAPE_CONVEX && s2.kind == SHAPE_BALL && t1 && t1->vert_first) { manifold_hull_ball(pos12, s1, *t1, s2.radius, prediction, false, m); return MANIFOLD_OK; }
    if (s1.kind == SHAPE_BALL || s2.kind == SHAPE_BALL) return MANIFOLD_UNSUPPORTED;
    // _ => contact_manifold_pfm_pfm (default_query_dispatcher.rs:818-831): Cuboid and ConvexPolyhedron are PolygonalFeatureMaps
    bool ok1 = s1.kind == SHAPE_CUBOID || (s1.kind == SHAPE_CONVEX && t1), ok2 = s2.kind == SHAPE_CUBOID || (s2.kind == SHAPE_CONVEX && t2);
    if (ok1 && ok2) { manifold_pfm_pfm(pos12, s1, t1, s2, t2, prediction, m, have_seed ? &seed : nullptr); return MANIFOLD_OK; }
    return MANIFOLD_UNSUPPORTED;
}

}  // namespace pb2o

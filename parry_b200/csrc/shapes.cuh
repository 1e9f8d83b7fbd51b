// parry_b200 — device shape table shared by the AABB / ray / contact kernels.
#pragma once
#include "common.cuh"

struct pb2_shapes {
    uint32_t n = 0, np = 0;
    bool has_convex = false;
    uint8_t* kinds = nullptr;   // pb2_shape_kind per shape
    float4* params = nullptr;   // ball {r}; cuboid {hx,hy,hz}; convex {first point, point count} (u32 bit patterns)
    float* points = nullptr;    // ConvexPolyhedron::points(), xyz packed
    float4* points4 = nullptr;  // same points padded to 16 B for vector loads in the support-map loop
    // optional face topology of the hulls (ConvexPolyhedron::faces / vertices_adj_to_face / edges_adj_to_face as parry builds
    // them, shape/convex_polyhedron.rs:390-637), needed by the pfm_pfm contact-manifold arm only
    uint32_t *hull_face_first = nullptr, *hull_face_count = nullptr;   // per table entry
    float* face_normal = nullptr;                                      // nf x 3
    uint32_t *face_first = nullptr, *face_count = nullptr;             // nf, into the adjacency arrays
    uint32_t *verts_adj_to_face = nullptr, *edges_adj_to_face = nullptr;   // vertex ids local to the hull, edge ids
    uint32_t nf = 0, nadj = 0;
    // vertex side of the topology (support_feature_id_toward): per point of the table; edge directions per table entry
    uint32_t *vert_first = nullptr, *vert_count = nullptr, *faces_adj_to_vertex = nullptr, *edges_adj_to_vertex = nullptr, *hull_edge_first = nullptr;
    float* edge_dir = nullptr;
    uint32_t* h_npoints = nullptr;   // host mirror: point count per entry (0 unless convex), for validating the topology
};

// Shape::compute_aabb(pos) (shape/shape.rs:369): aabb_ball.rs:8-33, aabb_cuboid.rs:9-16 + utils/isometry_ops.rs:16-18,
// aabb_convex_polyhedron.rs:8-16 + aabb_utils.rs:66-87.
__host__ __device__ __forceinline__ void shape_aabb_dev(uint8_t kind, float4 pr, const float* __restrict__ points, const Iso7& pos, V3& mn, V3& mx) {
    if (kind == PB2_SHAPE_BALL) {
        // ball_aabb: center + repeat(-r), center + repeat(r)
        float r = pr.x;
        mn = mk3(pos.t.x + (-r), pos.t.y + (-r), pos.t.z + (-r));
        mx = mk3(pos.t.x + r, pos.t.y + r, pos.t.z + r);
    } else if (kind == PB2_SHAPE_CUBOID) {
        // |R| * half_extents with R = to_rotation_matrix(), gemv accumulated column by column
        float qi = pos.q.i, qj = pos.q.j, qk = pos.q.k, qw = pos.q.w;
        float ww = qw * qw, ii = qi * qi, jj = qj * qj, kk = qk * qk;
        float ij = qi * qj * 2.0f, wk = qw * qk * 2.0f, wj = qw * qj * 2.0f;
        float ik = qi * qk * 2.0f, jk = qj * qk * 2.0f, wi = qw * qi * 2.0f;
        float m00 = fabsf(ww + ii - jj - kk), m01 = fabsf(ij - wk), m02 = fabsf(wj + ik);
        float m10 = fabsf(wk + ij), m11 = fabsf(ww - ii + jj - kk), m12 = fabsf(jk - wi);
        float m20 = fabsf(ik - wj), m21 = fabsf(wi + jk), m22 = fabsf(ww - ii - jj + kk);
        V3 he = mk3((m00 * pr.x + m01 * pr.y) + m02 * pr.z, (m10 * pr.x + m11 * pr.y) + m12 * pr.z,
                    (m20 * pr.x + m21 * pr.y) + m22 * pr.z);
        mn = pos.t - he;  // Aabb::from_half_extents(center, he)
        mx = pos.t + he;
    } else {
        uint32_t first = pb2_f2u(pr.x), cnt = pb2_f2u(pr.y);
        const float* p = points + 3ull * first;
        V3 w0 = iso_point(pos, mk3(p[0], p[1], p[2]));
        mn = w0; mx = w0;
        for (uint32_t k = 1; k < cnt; ++k) {
            V3 w = iso_point(pos, mk3(p[3 * k], p[3 * k + 1], p[3 * k + 2]));
            mn = vmin3(mn, w);
            mx = vmax3(mx, w);
        }
    }
}

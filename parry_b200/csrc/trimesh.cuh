// parry_b200 — TriMesh device object shared by raycast.cu (binary-tree kernels, C ABI) and trimesh_wide.cu
// (compressed 8-wide traversal tree).
#pragma once
#include "common.cuh"

struct pb2_trimesh {
    pb2_bvh bvh;
    uint32_t nt = 0, nv = 0;
    float4* tris = nullptr;    // [3 * sorted position]: {a, id}, {b, -}, {c, -}
    // Compressed 8-wide traversal tree (trimesh_wide.cu): 80-byte nodes, triangles re-gathered in wide-leaf order.
    float4* nodes8 = nullptr;  // W8_NODE_F4 (6) x float4 per node: 80 bytes of payload in a 96-byte, 32-byte-aligned record
    float4* tris8 = nullptr;   // the 48-byte record of `tris` padded to 64 bytes (W8_TRI_F4), ordered by (wide node, slot)
    uint32_t n_nodes8 = 0;
    int levels8 = 0;           // depth of the wide tree
};

// local_ray_intersection_with_triangle (ray_triangle.rs:70-152) — toi, face side (0 front / 1 back) and the
// un-normalised oriented normal. Returns false for None.
__device__ __forceinline__ bool ray_triangle(V3 a, V3 b, V3 c, V3 o, V3 dir, float& toi, uint32_t& fid, V3& n_out) {
    V3 ab = b - a, ac = c - a;
    V3 n = cross3(ab, ac);
    float d = dot3(n, dir);
    if (d == 0.0f) return false;
    V3 ap = o - a;
    float t = dot3(ap, n);
    if ((t < 0.0f && d < 0.0f) || (t > 0.0f && d > 0.0f)) return false;
    fid = d < 0.0f ? 0u : 1u;
    d = fabsf(d);
    V3 e = -cross3(dir, ap);
    float v, w;
    if (t < 0.0f) {
        v = -dot3(ac, e);
        if (v < 0.0f || v > d) return false;
        w = dot3(ab, e);
        if (w < 0.0f || v + w > d) return false;
        float invd = 1.0f / d;
        toi = -t * invd;
        n_out = n;
        fid |= 2u;  // bit 1: normal = -(n.normalize()) — negate after normalising
    } else {
        v = dot3(ac, e);
        if (v < 0.0f || v > d) return false;
        w = -dot3(ab, e);
        if (w < 0.0f || v + w > d) return false;
        float invd = 1.0f / d;
        toi = t * invd;
        n_out = n;
    }
    return true;
}

int pb2_wide_build(pb2_ctx* ctx, pb2_trimesh* mesh);
int pb2_wide_cast(pb2_ctx* ctx, const pb2_trimesh* mesh, const float* d_pose, const float* d_rays, const uint32_t* d_perm, uint32_t m,
                  float max_toi, float* d_toi, uint32_t* d_tri, float* d_n, uint32_t* d_f, bool with_normal, int tri_lanes, int refill, uint32_t cull, int shared_tri, const PieceSignal* pieces = nullptr);

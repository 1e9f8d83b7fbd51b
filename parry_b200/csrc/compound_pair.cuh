// query::contact between two Compounds (default_query_dispatcher.rs:338-351, nested composite dispatch): per-pair candidate
// enumeration and result reduction, written __host__ __device__ so that tests/hostcheck runs the very same functions on the CPU
// against the CPU restatement of the reference used as the test checker.
//
// contact_composite_shape_shape(pos12, compound1, shape2 = compound2) (contact_composite_shape_shape.rs:14-45) visits every part i of
// compound 1 whose AABB meets compound2.compute_aabb(pos12).loosened(prediction) — Shape::compute_aabb's default,
// local_aabb().transform_by(pos) (aabb.rs:492-498), on Compound::local_aabb = the merged part AABBs (compound.rs:120-127) — and
// dispatches contact(part_pos1[i].inv_mul(pos12), part_i, compound2), which lands in contact_shape_composite_shape (:63-76):
// pose.inverse(), compound 2 as the composite and part_i as the shape, flipped(). That inner call visits every part j of compound
// 2 whose AABB meets part_i.compute_aabb(pose).loosened(prediction) and runs contact(part_pos2[j].inv_mul(pose), part_j, part_i).
// A candidate is one (i, j); its leaf problem is handed to the contact kernels as shape1 = part_j at pos1 = part_pos2[j],
// shape2 = part_i at pos2 = pose (so that their pos1.inv_mul(pos2) is the reference's product), results in local frames.
#pragma once
#include "shapes.cuh"

struct CompoundTable {
    const uint32_t *first, *count, *part_shape;
    const float *part_pose, *part_aabb;   // np x 7, np x 6 (part AABBs in the compound's frame)
    uint32_t nc;
};

// Isometry::absolute_transform_vector (utils/isometry_ops.rs:16-18): |R| v, R = to_rotation_matrix(), accumulated by columns
__host__ __device__ __forceinline__ V3 iso_abs_vec(const Iso7& m, V3 v) {
    float qi = m.q.i, qj = m.q.j, qk = m.q.k, qw = m.q.w;
    float ww = qw * qw, ii = qi * qi, jj = qj * qj, kk = qk * qk;
    float ij = qi * qj * 2.0f, wk = qw * qk * 2.0f, wj = qw * qj * 2.0f;
    float ik = qi * qk * 2.0f, jk = qj * qk * 2.0f, wi = qw * qi * 2.0f;
    float m00 = fabsf(ww + ii - jj - kk), m01 = fabsf(ij - wk), m02 = fabsf(wj + ik);
    float m10 = fabsf(wk + ij), m11 = fabsf(ww - ii + jj - kk), m12 = fabsf(jk - wi);
    float m20 = fabsf(ik - wj), m21 = fabsf(wi + jk), m22 = fabsf(ww - ii - jj + kk);
    return mk3((m00 * v.x + m01 * v.y) + m02 * v.z, (m10 * v.x + m11 * v.y) + m12 * v.z, (m20 * v.x + m21 * v.y) + m22 * v.z);
}
// Aabb::intersects (aabb.rs:951-953), inclusive on every axis: part box b = {mins, maxs} against [mn, mx]
__host__ __device__ __forceinline__ bool aabb6_intersects(const float* b, V3 mn, V3 mx) {
    return b[0] <= mx.x && b[1] <= mx.y && b[2] <= mx.z && mn.x <= b[3] && mn.y <= b[4] && mn.z <= b[5];
}

// Candidates of one pair (compound c1 vs compound c2 under pos12), in (i, j) order. FILL = false: returns their number. FILL = true:
// also writes candidate number `at + c`: ij = {global part of compound 1, global part of compound 2}, the leaf problem's shapes
// (cs1 = shape of part j, cs2 = shape of part i) and poses (cp1 = part_pos2[j], cp2 = pose of part i in compound 2's frame).
template <bool FILL>
__host__ __device__ __forceinline__ uint32_t cc_candidates(const uint8_t* kinds, const float4* params, const float* points, const CompoundTable& T,
                                                           uint32_t c1, uint32_t c2, const Iso7& pos12, float prediction, uint32_t at, uint32_t* ij,
                                                           uint32_t* cs1, uint32_t* cs2, float* cp1, float* cp2) {
    uint32_t f1 = T.first[c1], m1 = T.count[c1], f2 = T.first[c2], m2 = T.count[c2];
    V3 amn = mk3(FLT_MAX, FLT_MAX, FLT_MAX), amx = mk3(-FLT_MAX, -FLT_MAX, -FLT_MAX);   // Aabb::new_invalid, then merge
    for (uint32_t j = 0; j < m2; ++j) {
        const float* b = T.part_aabb + 6ull * (f2 + j);
        amn = vmin3(amn, mk3(b[0], b[1], b[2]));
        amx = vmax3(amx, mk3(b[3], b[4], b[5]));
    }
    // Aabb::transform_by(pos12).loosened(prediction)
    V3 ctr = iso_point(pos12, (amn + amx) * 0.5f);
    V3 he = iso_abs_vec(pos12, (amx - amn) * 0.5f);
    V3 lmn = ctr + (-he), lmx = ctr + he;
    lmn = mk3(lmn.x + (-prediction), lmn.y + (-prediction), lmn.z + (-prediction));
    lmx = mk3(lmx.x + prediction, lmx.y + prediction, lmx.z + prediction);
    uint32_t cnt = 0;
    for (uint32_t i = 0; i < m1; ++i) {
        if (!aabb6_intersects(T.part_aabb + 6ull * (f1 + i), lmn, lmx)) continue;
        Iso7 pose = iso_inverse(iso_inv_mul(load_iso(T.part_pose + 7ull * (f1 + i)), pos12));   // part i in compound 2's frame
        uint32_t sid = T.part_shape[f1 + i];
        V3 smn, smx;
        shape_aabb_dev(kinds[sid], params[sid], points, pose, smn, smx);
        smn = mk3(smn.x + (-prediction), smn.y + (-prediction), smn.z + (-prediction));
        smx = mk3(smx.x + prediction, smx.y + prediction, smx.z + prediction);
        for (uint32_t j = 0; j < m2; ++j) {
            if (!aabb6_intersects(T.part_aabb + 6ull * (f2 + j), smn, smx)) continue;
            if (FILL) {
                size_t c = (size_t)at + cnt;
                ij[2 * c] = f1 + i; ij[2 * c + 1] = f2 + j;
                cs1[c] = T.part_shape[f2 + j]; cs2[c] = sid;
                const float* pj = T.part_pose + 7ull * (f2 + j);
                for (int d = 0; d < 7; ++d) cp1[7 * c + d] = pj[d];
                float* o = cp2 + 7 * c;
                o[0] = pose.q.i; o[1] = pose.q.j; o[2] = pose.q.k; o[3] = pose.q.w; o[4] = pose.t.x; o[5] = pose.t.y; o[6] = pose.t.z;
            }
            cnt++;
        }
    }
    return cnt;
}

// Reduction of one pair's candidates [lo, hi) (leaf contacts in local frames: {point1, normal1} in part j's frame, {point2, normal2}
// in part i's): the first strictly smaller dist wins at both levels of the reference, i.e. the smallest dist with ties going to the
// smallest (i, j). Winner: transform1_by(part_pos2[j]) (inner composite arm), flipped(), transform1_by(part_pos1[i]) (outer arm),
// then Contact::transform_by_mut(pos1, pos2) (contact_shape_shape.rs:132-135). Returns the status (1 Some, 0 None, 3 needs host).
__host__ __device__ __forceinline__ int cc_reduce(const uint32_t* ij, const float* cand, const uint8_t* cst, uint32_t lo, uint32_t hi,
                                                  const CompoundTable& T, uint32_t c1, uint32_t c2, const Iso7& p1w, const Iso7& p2w, float* o,
                                                  uint32_t* parts) {
    bool have = false, needs_host = false;
    uint32_t best_c = 0;
    float best = 0.0f;
    for (uint32_t c = lo; c < hi; ++c) {
        if (cst[c] == 3) needs_host = true;
        if (cst[c] != 1) continue;
        float d = cand[13ull * c + 12];
        if (!have || d < best) { best = d; best_c = c; have = true; }
    }
    if (needs_host || !have) {
        for (int d = 0; d < 13; ++d) o[d] = 0.0f;
        parts[0] = 0xFFFFFFFFu; parts[1] = 0xFFFFFFFFu;
        return needs_host ? 3 : 0;
    }
    const float* cj = cand + 13ull * best_c;
    uint32_t gi = ij[2ull * best_c], gj = ij[2ull * best_c + 1];
    Iso7 ppi = load_iso(T.part_pose + 7ull * gi), ppj = load_iso(T.part_pose + 7ull * gj);
    V3 q1 = iso_point(ppj, mk3(cj[0], cj[1], cj[2])), m1 = iso_vec(ppj, mk3(cj[6], cj[7], cj[8]));   // compound 2's frame
    V3 P1 = iso_point(ppi, mk3(cj[3], cj[4], cj[5])), N1 = iso_vec(ppi, mk3(cj[9], cj[10], cj[11]));  // flipped, compound 1's frame
    V3 w1 = iso_point(p1w, P1), w2 = iso_point(p2w, q1), n1 = iso_vec(p1w, N1), n2 = iso_vec(p2w, m1);
    o[0] = w1.x; o[1] = w1.y; o[2] = w1.z; o[3] = w2.x; o[4] = w2.y; o[5] = w2.z;
    o[6] = n1.x; o[7] = n1.y; o[8] = n1.z; o[9] = n2.x; o[10] = n2.y; o[11] = n2.z; o[12] = cj[12];
    parts[0] = gi - T.first[c1]; parts[1] = gj - T.first[c2];
    return 1;
}

// ContactManifold::try_update_contacts_eps (query/contact_manifolds/contact_manifold.rs:662-699) on one manifold stored in the
// batch layout of pb2_contact_manifolds_batch (normals: local_n1, local_n2; points: cnt x 9 words {local_p1, local_p2, dist,
// fid1, fid2}). __host__ __device__ so that tests/hostcheck can run the very same function on the CPU against the oracle.
#pragma once
#include "common.cuh"

#define PB2_COS_1_DEGREES 0.99984769515f   // utils::COS_1_DEGREES (try_update_contacts' DOT_THRESHOLD, contact_manifold.rs:655)
#define PB2_UPDATE_DIST_SQ 1.0e-6f         // DIST_SQ_THRESHOLD (:656)

// Returns true when the manifold survives under pos12; dist and local_p1 of the points visited are refreshed in place (points
// before a rejecting one stay refreshed, as in the reference, which then recomputes the manifold from scratch).
__host__ __device__ __forceinline__ bool manifold_try_update_core(const Iso7& pos12, const float* nr, uint32_t cnt, float* q,
                                                                  float angle_dot_threshold, float dist_sq_threshold) {
    if (cnt == 0) return false;
    V3 n1 = mk3(nr[0], nr[1], nr[2]);
    V3 n2 = iso_vec(pos12, mk3(nr[3], nr[4], nr[5]));
    if (-dot3(n1, n2) < angle_dot_threshold) return false;
    for (uint32_t i = 0; i < cnt; ++i) {
        float* o = q + 9u * i;
        V3 p1 = mk3(o[0], o[1], o[2]);
        V3 p2 = iso_point(pos12, mk3(o[3], o[4], o[5]));
        V3 dpt = p2 - p1;
        float dist = dot3(dpt, n1);
        if (dist * o[6] < 0.0f) return false;           // switched between penetrating and separated
        V3 np1 = p2 - n1 * dist;
        if (nrm2(p1 - np1) > dist_sq_threshold) return false;
        o[6] = dist;
        o[0] = np1.x; o[1] = np1.y; o[2] = np1.z;
    }
    return true;
}

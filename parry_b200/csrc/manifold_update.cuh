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

// One thread of k_manifold_try_update (contact.cu): ContactManifold::try_update_contacts_eps on pair k, in place. dispatch: only the
// pairs whose dispatcher arm tries to keep last frame's manifold (contact_manifolds_cuboid_cuboid.rs:28, contact_manifolds_pfm_pfm.rs:63:
// neither shape a Ball, hulls with face topology); kept pairs get status 0. old_fids / old_counts (optional): last frame's feature
// ids, saved for match_contacts before the recomputation overwrites them.
__host__ __device__ __forceinline__ void manifold_try_update_pair(uint32_t k, const uint8_t* kinds, uint32_t n_shapes, const uint32_t* shape1,
                                                                  const uint32_t* shape2, bool dispatch, bool have_topology, const float* pos1,
                                                                  const float* pos2, uint32_t max_points, float angle_dot_threshold,
                                                                  float dist_sq_threshold, const float* normals, const uint32_t* counts, float* pts,
                                                                  uint8_t* kept, uint8_t* status, uint32_t* old_fids, uint32_t* old_counts) {
    uint32_t cnt = counts[k];
    if (cnt > max_points) cnt = max_points;
    float* q = pts + (size_t)k * max_points * 9;
    if (old_fids) {
        old_counts[k] = cnt;
        uint32_t* f = old_fids + (size_t)k * max_points * 2;
        for (uint32_t i = 0; i < cnt; ++i) { f[2 * i] = pb2_f2u(q[9 * i + 7]); f[2 * i + 1] = pb2_f2u(q[9 * i + 8]); }
    }
    bool tries = true;
    if (dispatch) {
        uint32_t a = shape1[k], b = shape2[k];
        tries = a < n_shapes && b < n_shapes;
        if (tries) {
            uint8_t ka = kinds[a], kb = kinds[b];
            bool ok1 = ka == PB2_SHAPE_CUBOID || (ka == PB2_SHAPE_CONVEX && have_topology);
            bool ok2 = kb == PB2_SHAPE_CUBOID || (kb == PB2_SHAPE_CONVEX && have_topology);
            tries = ok1 && ok2;
        }
    }
    bool keep = false;
    if (tries && cnt) {
        Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));
        keep = manifold_try_update_core(pos12, normals + 6ull * k, cnt, q, angle_dot_threshold, dist_sq_threshold);
    }
    kept[k] = keep ? 1 : 0;
    if (keep && status) status[k] = 0;
}

// One thread of k_manifold_match: ContactManifold::match_contacts (contact_manifold.rs:761-770) as an index: match[k][i] = the last
// old point of pair k whose two feature ids equal those of new point i (the one whose ContactData the reference's loop leaves in
// place), -1 = none; kept manifolds map onto themselves.
__host__ __device__ __forceinline__ void manifold_match_pair(uint32_t k, const uint8_t* kept, const uint32_t* old_fids, const uint32_t* old_counts,
                                                             const uint32_t* counts, const float* pts, uint32_t max_points, int32_t* match) {
    uint32_t cnt = counts[k], oc = old_counts[k];
    if (cnt > max_points) cnt = max_points;
    const float* q = pts + (size_t)k * max_points * 9;
    const uint32_t* f = old_fids + (size_t)k * max_points * 2;
    int32_t* m = match + (size_t)k * max_points;
    bool keep = kept[k] != 0;
    for (uint32_t i = 0; i < max_points; ++i) {
        int32_t j = -1;
        if (i < cnt) {
            if (keep) j = (int32_t)i;
            else {
                uint32_t f1 = pb2_f2u(q[9 * i + 7]), f2 = pb2_f2u(q[9 * i + 8]);
                for (uint32_t o = 0; o < oc; ++o) if (f[2 * o] == f1 && f[2 * o + 1] == f2) j = (int32_t)o;
            }
        }
        m[i] = j;
    }
}

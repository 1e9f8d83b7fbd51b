// Test infrastructure (never linked into libparry_b200.so): runs the __host__ __device__ core of k_manifold_try_update
// (parry_b200/csrc/manifold_update.cuh) on the CPU, so that the CPU suite can check the very function the kernel calls against the
// oracle's restatement of ContactManifold::try_update_contacts_eps without a GPU. Built by tests/test_hostcheck.py with nvcc and the
// library's own floating-point flags.
#include "../../parry_b200/csrc/manifold_update.cuh"

extern "C" void hostcheck_manifolds_try_update(const float* pos1, const float* pos2, uint32_t n, uint32_t max_points, float angle_dot_threshold,
                                               float dist_sq_threshold, const float* normals, const uint32_t* counts, float* pts, uint8_t* kept) {
    for (uint32_t k = 0; k < n; ++k) {
        uint32_t cnt = counts[k] > max_points ? max_points : counts[k];
        Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));
        kept[k] = cnt && manifold_try_update_core(pos12, normals + 6ull * k, cnt, pts + (size_t)k * max_points * 9, angle_dot_threshold,
                                                  dist_sq_threshold) ? 1 : 0;
    }
}

// ---- Compound vs Compound: candidate enumeration + reduction (parry_b200/csrc/compound_pair.cuh) with the leaf contacts supplied
// by the caller (the test computes them with the oracle's query::contact on the materialised leaf problems).
#include "../../parry_b200/csrc/compound_pair.cuh"

static CompoundTable host_table(const uint32_t* first, const uint32_t* count, const uint32_t* part_shape, const float* part_pose,
                                const float* part_aabb, uint32_t nc) {
    CompoundTable T;
    T.first = first; T.count = count; T.part_shape = part_shape; T.part_pose = part_pose; T.part_aabb = part_aabb; T.nc = nc;
    return T;
}
// Part AABBs as k_compound_part_aabbs computes them.
extern "C" void hostcheck_part_aabbs(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* part_shape,
                                     const float* part_pose, uint32_t np, float* out) {
    for (uint32_t i = 0; i < np; ++i) {
        uint32_t sid = part_shape[i];
        V3 mn, mx;
        shape_aabb_dev(kinds[sid], ((const float4*)params4)[sid], points, load_iso(part_pose + 7ull * i), mn, mx);
        float* o = out + 6ull * i;
        o[0] = mn.x; o[1] = mn.y; o[2] = mn.z; o[3] = mx.x; o[4] = mx.y; o[5] = mx.z;
    }
}
// fill == 0: counts[k]; fill != 0: candidates at offsets[k].
extern "C" void hostcheck_cc_candidates(const uint8_t* kinds, const float* params4, const float* points, const uint32_t* first, const uint32_t* count,
                                        const uint32_t* part_shape, const float* part_pose, const float* part_aabb, uint32_t nc, const uint32_t* id1,
                                        const float* pos1, const uint32_t* id2, const float* pos2, uint32_t n, float prediction, int fill,
                                        uint32_t* counts, const uint32_t* offsets, uint32_t* ij, uint32_t* cs1, uint32_t* cs2, float* cp1, float* cp2) {
    CompoundTable T = host_table(first, count, part_shape, part_pose, part_aabb, nc);
    for (uint32_t k = 0; k < n; ++k) {
        Iso7 pos12 = iso_inv_mul(load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k));
        if (fill) cc_candidates<true>(kinds, (const float4*)params4, points, T, id1[k], id2[k], pos12, prediction, offsets[k], ij, cs1, cs2, cp1, cp2);
        else counts[k] = cc_candidates<false>(kinds, (const float4*)params4, points, T, id1[k], id2[k], pos12, prediction, 0u, nullptr, nullptr, nullptr,
                                              nullptr, nullptr);
    }
}
extern "C" void hostcheck_cc_reduce(const uint32_t* offsets, const uint32_t* ij, const float* cand, const uint8_t* cst, const uint32_t* first,
                                    const uint32_t* count, const uint32_t* part_shape, const float* part_pose, const float* part_aabb, uint32_t nc,
                                    const uint32_t* id1, const float* pos1, const uint32_t* id2, const float* pos2, uint32_t n, float* out,
                                    uint8_t* status, uint32_t* parts) {
    CompoundTable T = host_table(first, count, part_shape, part_pose, part_aabb, nc);
    for (uint32_t k = 0; k < n; ++k)
        status[k] = (uint8_t)cc_reduce(ij, cand, cst, offsets[k], offsets[k + 1], T, id1[k], id2[k], load_iso(pos1 + 7ull * k), load_iso(pos2 + 7ull * k),
                                       out + 13ull * k, parts + 2ull * k);
}

// ---- Second-frame dispatch: one "thread" of k_manifold_try_update / k_manifold_match per pair (manifold_update.cuh), so that the
// dispatch rule (which arms try to keep a manifold), the saved feature ids and the match index are checked on the CPU as well.
extern "C" void hostcheck_manifold_try_update_pairs(const uint8_t* kinds, uint32_t n_shapes, const uint32_t* shape1, const uint32_t* shape2, int dispatch,
                                                    int have_topology, const float* pos1, const float* pos2, uint32_t n, uint32_t max_points,
                                                    const float* normals, const uint32_t* counts, float* pts, uint8_t* kept, uint8_t* status,
                                                    uint32_t* old_fids, uint32_t* old_counts) {
    for (uint32_t k = 0; k < n; ++k)
        manifold_try_update_pair(k, kinds, n_shapes, shape1, shape2, dispatch != 0, have_topology != 0, pos1, pos2, max_points, PB2_COS_1_DEGREES,
                                 PB2_UPDATE_DIST_SQ, normals, counts, pts, kept, status, old_fids, old_counts);
}
extern "C" void hostcheck_manifold_match_pairs(const uint8_t* kept, const uint32_t* old_fids, const uint32_t* old_counts, const uint32_t* counts,
                                               const float* pts, uint32_t n, uint32_t max_points, int32_t* match) {
    for (uint32_t k = 0; k < n; ++k) manifold_match_pair(k, kept, old_fids, old_counts, counts, pts, max_points, match);
}

// ---- PLOC link (parry_b200/csrc/ploc.cuh): the rounds of bvh_link_ploc (bvh_build.cu) replayed on the CPU with the very
// per-cluster functions the kernels call. aabbs are in sorted order (leaf id = sorted position here). nodes: n - 1 wide nodes.
#include "../../parry_b200/csrc/ploc.cuh"
#include <vector>
extern "C" int hostcheck_ploc_link(const float* aabbs, uint32_t n, uint32_t radius, void* nodes_out, uint32_t* parents, uint32_t* leaf_slot) {
    std::vector<float4> Ca(2 * (size_t)n), Cb(2 * (size_t)n);
    std::vector<uint32_t> cand(n), order(n);
    std::vector<unsigned long long> incl(n);
    for (uint32_t p = 0; p < n; ++p) {
        order[p] = p;
        const float* a = aabbs + 6ull * p;
        Ca[2 * p] = make_float4(a[0], a[1], a[2], pb2_u2f(p));
        Ca[2 * p + 1] = make_float4(a[3], a[4], a[5], pb2_u2f(1u | PB2_CHANGE_PENDING));
    }
    uint32_t c = n, created = 0;
    int rounds = 0;
    while (c > 1) {
        for (uint32_t i = 0; i < c; ++i) cand[i] = ploc_nearest(Ca.data(), 0u, c, i, radius);
        unsigned long long acc = 0;
        for (uint32_t i = 0; i < c; ++i) { acc += ploc_flags(cand.data(), i); incl[i] = acc; }
        for (uint32_t i = 0; i < c; ++i)
            ploc_emit(Ca.data(), cand.data(), incl.data(), i, Cb.data(), (NodeWide*)nodes_out, parents, leaf_slot, order.data(), created, n);
        created += (uint32_t)(acc >> 32);
        c = (uint32_t)acc;
        Ca.swap(Cb);
        ++rounds;
    }
    return created == n - 1 ? rounds : -1;
}

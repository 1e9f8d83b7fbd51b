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

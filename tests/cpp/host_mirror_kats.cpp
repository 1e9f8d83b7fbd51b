// Known-answer program for the C++ host mirror (include/parry_b200.hpp) above the C ABI: the reference's own exact pins run through
// the C++ types a maintainer would use. Exit code 0 = all checks passed, 1 = a check failed, 3 = no CUDA device (there is no CPU
// fallback). Built and run by tests/test_zx_cpp_mirror.py.
//   crates/parry3d/tests/geometry/epa3.rs:8-23      cuboid (2,1,1) vs itself: dist == -0.5, normal1 == -x; dist == -1.8, normal1 == -y
//   crates/parry3d/tests/geometry/ball_ball_toi.rs  time_of_impact == 0.9
//   a unit square of two triangles, three boxes: hand-checkable ray hit, pair set, intersect_aabb
#include <cstdio>
#include <cmath>
#include <algorithm>
#include "parry_b200.hpp"

static int failures = 0;
#define CHECK(cond) do { if (!(cond)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); ++failures; } } while (0)

int main() {
    try {
        pb2::Context ctx(0);
        const pb2::Isometry ident{{0, 0, 0, 1}, {0, 0, 0}};
        auto at = [](float x, float y, float z) { return pb2::Isometry{{0, 0, 0, 1}, {x, y, z}}; };

        // ---- shape table: cuboid (2,1,1), ball 0.5
        const uint8_t kinds[2] = {PB2_SHAPE_CUBOID, PB2_SHAPE_BALL};
        const float params[8] = {2.0f, 1.0f, 1.0f, 0.0f, 0.5f, 0.0f, 0.0f, 0.0f};
        pb2_shapes* shapes = nullptr;
        ctx.check(pb2_shapes_create(ctx.get(), kinds, params, 2, nullptr, 0, &shapes));

        // epa3.rs:8-23: contact_support_map_support_map(&m1.inv_mul(&m2), c, c, 10.0) = query::contact(m1, c, m2, c, 10.0) in m1's frame
        {
            std::vector<pb2_contact> out;
            std::vector<uint8_t> st;
            pb2::query::contact(ctx, shapes, {0, 0}, {at(3.5f, 0, 0), at(0, 0.2f, 0)}, {0, 0}, {ident, ident}, 10.0f, out, st);
            CHECK(st[0] == 1 && out[0].dist == -0.5f);
            CHECK(out[0].normal1[0] == -1.0f && out[0].normal1[1] == 0.0f && out[0].normal1[2] == 0.0f);
            CHECK(st[1] == 1 && out[1].dist == -1.8f);
            CHECK(out[1].normal1[0] == 0.0f && out[1].normal1[1] == -1.0f && out[1].normal1[2] == 0.0f);
        }
        // ball_ball_toi.rs
        {
            std::vector<pb2::query::ShapeCastHit> hit;
            std::vector<uint8_t> st;
            pb2::query::cast_shapes(ctx, shapes, {1}, {ident}, {pb2::query::Vector{0, 10, 0}}, {1}, {at(0, 10, 0)}, {pb2::query::Vector{0, 0, 0}},
                                    pb2::query::ShapeCastOptions(), hit, st);
            CHECK(st[0] == PB2_CAST_CONVERGED && hit[0].time_of_impact == 0.9f);
        }
        // distance between the two balls 3 apart: 3 - 0.5 - 0.5
        {
            std::vector<float> d;
            std::vector<uint8_t> st;
            pb2::query::distance(ctx, shapes, {1}, {ident}, {1}, {at(3, 0, 0)}, d, st);
            CHECK(d[0] == 2.0f);
        }
        ctx.check(pb2_shapes_destroy(ctx.get(), shapes));

        // ---- TriMesh: the unit square z = 0 as triangles (0,1,2) and (0,2,3)
        {
            pb2::TriMesh mesh(ctx, {0, 0, 0, 1, 0, 0, 1, 1, 0, 0, 1, 0}, {0, 1, 2, 0, 2, 3});
            std::vector<pb2::Ray> rays = {{{0.75f, 0.25f, 2.0f}, {0, 0, -1}}, {{0.25f, 0.75f, 3.0f}, {0, 0, -1}}, {{2.0f, 2.0f, 1.0f}, {0, 0, -1}},
                                          {{0.5f, 0.25f, -4.0f}, {0, 0, 2}}};
            std::vector<float> toi;
            std::vector<uint32_t> tri;
            mesh.cast_ray(nullptr, rays, 100.0f, true, toi, tri);
            CHECK(tri[0] == 0 && toi[0] == 2.0f);
            CHECK(tri[1] == 1 && toi[1] == 3.0f);
            CHECK(tri[2] == PB2_INVALID_U32);
            CHECK(tri[3] == 0 && toi[3] == 2.0f);             // toi is in units of the (non-normalised) direction
            auto hits = mesh.cast_ray_and_get_normal(nullptr, rays, 100.0f, true, tri);
            CHECK(hits[0].normal[2] == 1.0f && hits[3].normal[2] == -1.0f);   // the normal faces the ray's origin side
            const pb2::Isometry up = at(0, 0, 1);               // mesh lifted by 1: the first ray now hits at toi 1
            mesh.cast_ray(&up, rays, 100.0f, true, toi, tri);
            CHECK(tri[0] == 0 && toi[0] == 1.0f);
        }
        // ---- Bvh: boxes 0 and 1 overlap, box 2 is apart
        {
            std::vector<pb2::Aabb> boxes = {{{0, 0, 0}, {1, 1, 1}}, {{0.5f, 0.5f, 0.5f}, {2, 2, 2}}, {{5, 5, 5}, {6, 6, 6}}};
            pb2::Bvh bvh = pb2::Bvh::from_leaves(ctx, pb2::BvhBuildStrategy::Binned, boxes);
            CHECK(bvh.leaf_count() == 3);
            std::vector<std::pair<uint32_t, uint32_t>> pairs;
            bvh.traverse_bvtt_single_tree(false, [&](uint32_t a, uint32_t b) { pairs.push_back({std::min(a, b), std::max(a, b)}); });
            CHECK(pairs.size() == 1 && pairs[0].first == 0 && pairs[0].second == 1);
            auto csr = bvh.intersect_aabb({{{0.9f, 0.9f, 0.9f}, {1.1f, 1.1f, 1.1f}}, {{5.5f, 5.5f, 5.5f}, {7, 7, 7}}, {{3, 3, 3}, {4, 4, 4}}});
            CHECK(csr.first[1] - csr.first[0] == 2 && csr.first[2] - csr.first[1] == 1 && csr.first[3] - csr.first[2] == 0);
            CHECK(csr.second.size() == 3 && csr.second[2] == 2);
            pb2::Aabb root = bvh.root_aabb();
            CHECK(root.mins[0] == 0.0f && root.maxs[2] == 6.0f);
        }
    } catch (const pb2::Error& e) {
        std::printf("pb2::Error: %s\n", e.what());
        return 3;
    }
    if (failures) return 1;
    std::printf("host mirror known answers: ok\n");
    return 0;
}

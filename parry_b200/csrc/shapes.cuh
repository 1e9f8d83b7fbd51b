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
};

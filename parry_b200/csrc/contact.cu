// parry_b200 — query::contact batch (placeholder until the GJK/EPA kernels land).
#include "shapes.cuh"

extern "C" {
int pb2_contact_batch(pb2_ctx* ctx, const pb2_shapes*, const uint32_t*, const uint32_t*, const float*, const float*, float, uint32_t,
                      pb2_contact*, uint8_t*, uint64_t*, int) {
    if (!ctx) return PB2_ERR_INVALID;
    PB2_FAIL(ctx, PB2_ERR_UNSUPPORTED, "contact kernels not built yet");
}
int pb2_contact_batch_compact(pb2_ctx* ctx, const pb2_shapes*, const uint32_t*, const uint32_t*, const float*, const float*, float,
                              uint32_t, pb2_contact*, uint32_t*, uint64_t, uint64_t*, int) {
    if (!ctx) return PB2_ERR_INVALID;
    PB2_FAIL(ctx, PB2_ERR_UNSUPPORTED, "contact kernels not built yet");
}
}

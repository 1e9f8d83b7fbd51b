// parry_b200 — shared device/host definitions. Compiled with --fmad=false: every a*b+c stays un-fused so f32
// results are bit-identical to the reference's (Rust never contracts), with IEEE div/sqrt (nvcc defaults
// -prec-div=true -prec-sqrt=true, -ftz=false).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <float.h>
#include "../../include/parry_b200.h"

#define PB2_SM_COUNT_FALLBACK 148

struct Scratch {
    void* ptr = nullptr;
    size_t cap = 0;
};

struct pb2_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = PB2_SM_COUNT_FALLBACK;
    uint64_t launches = 0;
    char err[512] = {0};
    // grow-only device staging for PB2_MEM_HOST calls + misc scratch
    Scratch stage[8];
    Scratch scratch[5];
    // pinned host bounce buffer for small readbacks (counters)
    uint64_t* h_counters = nullptr;  // 16 x u64 pinned
    uint64_t* d_counters = nullptr;  // 16 x u64 device
    // copy/compute pipeline for PB2_MEM_HOST batches: H2D stream, D2H stream, event pool
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaStream_t copy_peer[6] = {nullptr};  // peer pushes of the multi-GPU gather: one DMA queue per group of peers
    cudaStream_t compute2 = nullptr;  // second compute stream: consecutive host-mode ray chunks alternate so that one chunk's tail overlaps the next one's start
    int ray_slot = 8;                 // d_counters slot of the persistent ray kernels' fetch counter (8 or 12, one per compute stream)
    cudaEvent_t ev[64] = {nullptr};
    int ev_next = 0;
    // optional phase timing of the contact pipeline (pb2_ctx_enable_phase_timing): events on the launching stream around the
    // GJK kernel, the EPA kernel and the finishing kernel of the last run_contacts call (bench.py's per-kernel roofline)
    bool phase_timing = false;
    cudaEvent_t phase_ev[4] = {nullptr, nullptr, nullptr, nullptr};
    int phase_marks = 0;
    void* epa_big_arena = nullptr;     // arena of the overflow EPA kernel (contact.cu), allocated at the first contact call
    unsigned int* d_pieces = nullptr;  // 2 x 32 u32: per-piece retired-ray counters and completion flags of a piece-signalling ray cast
    void* wait_value32 = nullptr;      // cuStreamWaitValue32, resolved once (NULL: not available -> piece-wise launches)
};
struct PieceSignal { uint32_t size; unsigned int* done; unsigned int* flag; uint32_t flush_every; uint32_t shift; };   // size == 1 << shift
int pb2_pipeline_init(pb2_ctx* ctx);
static inline cudaEvent_t pb2_next_event(pb2_ctx* ctx) { cudaEvent_t e = ctx->ev[ctx->ev_next]; ctx->ev_next = (ctx->ev_next + 1) % 64; return e; }

#define PB2_CUDA(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t _e = (expr);                                                                        \
        if (_e != cudaSuccess) {                                                                        \
            snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr,       \
                     cudaGetErrorString(_e));                                                           \
            return PB2_ERR_CUDA;                                                                        \
        }                                                                                               \
    } while (0)

#define PB2_CHECK(expr)            \
    do {                           \
        int _s = (expr);           \
        if (_s != PB2_OK) return _s; \
    } while (0)

#define PB2_FAIL(ctx, code, ...)                                \
    do {                                                        \
        snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);  \
        return (code);                                          \
    } while (0)

int pb2_scratch_reserve(pb2_ctx* ctx, Scratch* s, size_t bytes);

// Sticky device-side fault word (d_counters slot 15). The binary-tree walks keep a fixed per-thread stack; the reference's
// SmallVec grows instead (bvh_traverse.rs:343). A Karras tree over 63-bit Morton keys with the index tie-break is at most
// 63 + 32 levels deep, which PB2_STACK covers, so the flag can only fire for trees linked another way (PLOC) on adversarial
// input. A full stack never drops work silently: the push sets the flag and the next synchronising call (or
// pb2_ctx_synchronize) returns PB2_ERR_DEPTH.
#define PB2_FAULT_SLOT 15
#define PB2_FAULT_STACK 1u
#define PB2_FAULT_BAD_ID 2u   // an id read from a device-resident array was out of range (nothing was dereferenced)
#define PB2_STACK 96
#define PB2_FAULT_PTR(ctx) ((unsigned int*)((ctx)->d_counters + PB2_FAULT_SLOT))
int pb2_check_fault(pb2_ctx* ctx);   // after a stream synchronisation that followed pb2_fetch_fault
int pb2_fetch_fault(pb2_ctx* ctx);   // enqueues the 4-byte read-back of the fault word
#ifdef __CUDACC__
__device__ __forceinline__ void pb2_push(uint32_t* stack, int& sp, uint32_t v, unsigned int* fault) {
    if (sp < PB2_STACK) stack[sp++] = v;
    else atomicOr(fault, PB2_FAULT_STACK);
}
#endif
static inline unsigned pb2_blocks(uint64_t n, unsigned threads) { return (unsigned)((n + threads - 1) / threads); }
#define PB2_LAUNCHED(ctx) ((ctx)->launches++)
static inline void pb2_phase_mark(pb2_ctx* ctx, int i) {
    if (ctx->phase_timing && ctx->phase_ev[i]) { cudaEventRecord(ctx->phase_ev[i], ctx->stream); if (ctx->phase_marks < i + 1) ctx->phase_marks = i + 1; }
}

// ---------------------------------------------------------------- node layout (== BvhNodeWide, 64 B)
// child = { mins.xyz, children:u32, maxs.xyz, data:u32 }; data low 30 bits = leaf_count, top 2 = change flags.
struct __align__(16) NodeHalf {
    float mnx, mny, mnz;
    uint32_t children;
    float mxx, mxy, mxz;
    uint32_t data;
};
struct __align__(64) NodeWide {
    NodeHalf left, right;
};
#define PB2_LEAF_COUNT_MASK 0x3fffffffu
#define PB2_CHANGED (1u << 30)
#define PB2_CHANGE_PENDING (3u << 30)

struct pb2_bvh {
    uint32_t n_leaves = 0;
    uint32_t n_nodes = 0;      // number of wide nodes (max(1, n-1); 0 when empty)
    int strategy = 0;
    NodeWide* nodes = nullptr; // leaf `children` = sorted position (see leaf_order), internal = wide node index
    uint32_t* parents = nullptr;     // per wide node: (parent << 1) | is_right ; parents[0] = 0 (dummy)
    uint32_t* leaf_slot = nullptr;   // per leaf id: (node << 1) | is_right  (== leaf_node_indices)
    uint32_t* leaf_order = nullptr;  // sorted position -> leaf id
    uint32_t* counters = nullptr;    // per wide node arrival counters for bottom-up passes
    uint32_t cap_leaves = 0;
    // true: Karras numbering (children of a node at (split, split + 1), a subtree's leaves are a contiguous range of sorted
    // positions ending at the left child's index) — the self-pair walk culls "already seen" subtrees with it. false: PLOC
    // numbering (root 0, children after their parent, no range property).
    bool karras = true;
};

// ---------------------------------------------------------------- device math (nalgebra op order)
struct V3 {
    float x, y, z;
};
__host__ __device__ __forceinline__ V3 mk3(float x, float y, float z) { V3 v; v.x = x; v.y = y; v.z = z; return v; }
__host__ __device__ __forceinline__ V3 operator+(V3 a, V3 b) { return mk3(a.x + b.x, a.y + b.y, a.z + b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a, V3 b) { return mk3(a.x - b.x, a.y - b.y, a.z - b.z); }
__host__ __device__ __forceinline__ V3 operator-(V3 a) { return mk3(-a.x, -a.y, -a.z); }
__host__ __device__ __forceinline__ V3 operator*(V3 a, float s) { return mk3(a.x * s, a.y * s, a.z * s); }
__host__ __device__ __forceinline__ V3 operator/(V3 a, float s) { return mk3(a.x / s, a.y / s, a.z / s); }
__host__ __device__ __forceinline__ float dot3(V3 a, V3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
__host__ __device__ __forceinline__ V3 cross3(V3 a, V3 b) {
    return mk3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__host__ __device__ __forceinline__ float nrm2(V3 a) { return dot3(a, a); }
__device__ __forceinline__ float nrm(V3 a) { return sqrtf(dot3(a, a)); }
__device__ __forceinline__ V3 normalize3(V3 a) { return a / nrm(a); }
__host__ __device__ __forceinline__ V3 vmin3(V3 a, V3 b) { return mk3(fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)); }
__host__ __device__ __forceinline__ V3 vmax3(V3 a, V3 b) { return mk3(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)); }
__device__ __forceinline__ float comp(V3 v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : v.z); }

// bit pattern of a float (shape tables keep u32 fields in float4 params); __float_as_uint on the device
__host__ __device__ __forceinline__ uint32_t pb2_f2u(float f) {
#ifdef __CUDA_ARCH__
    return __float_as_uint(f);
#else
    uint32_t u; memcpy(&u, &f, 4); return u;
#endif
}

struct Q4 {
    float i, j, k, w;
};
struct Iso7 {
    Q4 q;
    V3 t;
};
__host__ __device__ __forceinline__ Iso7 load_iso(const float* p) {
    Iso7 m;
    m.q.i = p[0]; m.q.j = p[1]; m.q.k = p[2]; m.q.w = p[3];
    m.t = mk3(p[4], p[5], p[6]);
    return m;
}
__host__ __device__ __forceinline__ Q4 qconj(Q4 q) { Q4 r; r.i = -q.i; r.j = -q.j; r.k = -q.k; r.w = q.w; return r; }
// UnitQuaternion * Vector3: t = (q.xyz x v) * 2 ; t * w + (q.xyz x t) + v
__host__ __device__ __forceinline__ V3 qrot(Q4 q, V3 v) {
    V3 u = mk3(q.i, q.j, q.k);
    V3 t = cross3(u, v) * 2.0f;
    V3 c = cross3(u, t);
    return (t * q.w + c) + v;
}
__host__ __device__ __forceinline__ V3 qirot(Q4 q, V3 v) { return qrot(qconj(q), v); }
__host__ __device__ __forceinline__ Q4 qmul(Q4 a, Q4 b) {
    Q4 r;
    r.i = a.w * b.i + a.i * b.w + a.j * b.k - a.k * b.j;
    r.j = a.w * b.j - a.i * b.k + a.j * b.w + a.k * b.i;
    r.k = a.w * b.k + a.i * b.j - a.j * b.i + a.k * b.w;
    r.w = a.w * b.w - a.i * b.i - a.j * b.j - a.k * b.k;
    return r;
}
__host__ __device__ __forceinline__ V3 iso_point(const Iso7& m, V3 p) { return qrot(m.q, p) + m.t; }
__host__ __device__ __forceinline__ V3 iso_vec(const Iso7& m, V3 v) { return qrot(m.q, v); }
__host__ __device__ __forceinline__ V3 iso_inv_point(const Iso7& m, V3 p) { return qirot(m.q, p - m.t); }
__host__ __device__ __forceinline__ V3 iso_inv_vec(const Iso7& m, V3 v) { return qirot(m.q, v); }
__host__ __device__ __forceinline__ Iso7 iso_inv_mul(const Iso7& a, const Iso7& b) {
    Iso7 r;
    Q4 inv = qconj(a.q);
    r.t = qrot(inv, b.t - a.t);
    r.q = qmul(inv, b.q);
    return r;
}
__host__ __device__ __forceinline__ Iso7 iso_inverse(const Iso7& a) {
    Iso7 r;
    r.q = qconj(a.q);
    r.t = -qrot(r.q, a.t);
    return r;
}

// slab test == Aabb::cast_local_ray(ray, max_toi, solid = true) (query/ray/ray_aabb.rs:12-49) with
// FLT_MAX for None (BvhNode::cast_ray, bvh_tree.rs:1177-1181). inv = 1/dir (only read where dir != 0).
__device__ __forceinline__ float slab_cost(float mnx, float mny, float mnz, float mxx, float mxy, float mxz, V3 o, V3 d,
                                           V3 inv, float tmax) {
    float tmin = 0.0f;
    if (d.x == 0.0f) {
        if (o.x < mnx || o.x > mxx) return FLT_MAX;
    } else {
        float n = (mnx - o.x) * inv.x, f = (mxx - o.x) * inv.x;
        if (n > f) { float t = n; n = f; f = t; }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return FLT_MAX;
    }
    if (d.y == 0.0f) {
        if (o.y < mny || o.y > mxy) return FLT_MAX;
    } else {
        float n = (mny - o.y) * inv.y, f = (mxy - o.y) * inv.y;
        if (n > f) { float t = n; n = f; f = t; }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return FLT_MAX;
    }
    if (d.z == 0.0f) {
        if (o.z < mnz || o.z > mxz) return FLT_MAX;
    } else {
        float n = (mnz - o.z) * inv.z, f = (mxz - o.z) * inv.z;
        if (n > f) { float t = n; n = f; f = t; }
        tmin = fmaxf(tmin, n);
        tmax = fminf(tmax, f);
        if (tmin > tmax) return FLT_MAX;
    }
    return tmin;
}

// Branch-free form of slab_cost: identical values for every input (IEEE inf/NaN arithmetic reproduces the reference's
// dir == 0 special case: inv = +-inf gives (-inf, +inf) inside the slab, a same-signed infinite pair outside, and NaN —
// ignored by fmaxf/fminf exactly like Rust's f32::max/min — on the boundary). The per-axis early-outs of the reference
// are equivalent to one final test because tmin only grows and tmax only shrinks.
__device__ __forceinline__ float slab_cost_bf(float4 lo, float4 hi, V3 o, V3 inv, float tmax) {
    float tmin = 0.0f;
    float n, f, a, b;
    n = (lo.x - o.x) * inv.x; f = (hi.x - o.x) * inv.x;
    a = n > f ? f : n; b = n > f ? n : f;
    tmin = fmaxf(tmin, a); tmax = fminf(tmax, b);
    n = (lo.y - o.y) * inv.y; f = (hi.y - o.y) * inv.y;
    a = n > f ? f : n; b = n > f ? n : f;
    tmin = fmaxf(tmin, a); tmax = fminf(tmax, b);
    n = (lo.z - o.z) * inv.z; f = (hi.z - o.z) * inv.z;
    a = n > f ? f : n; b = n > f ? n : f;
    tmin = fmaxf(tmin, a); tmax = fminf(tmax, b);
    return tmin > tmax ? FLT_MAX : tmin;
}

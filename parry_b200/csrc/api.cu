// parry_b200 — context management and small host helpers of the C ABI (include/parry_b200.h).
#include "common.cuh"

int pb2_scratch_reserve(pb2_ctx* ctx, Scratch* s, size_t bytes) {
    if (bytes <= s->cap) return PB2_OK;
    if (s->ptr) {
        PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        PB2_CUDA(ctx, cudaFree(s->ptr));
        s->ptr = nullptr;
        s->cap = 0;
    }
    size_t want = bytes + bytes / 4 + 256;
    PB2_CUDA(ctx, cudaMalloc(&s->ptr, want));
    s->cap = want;
    return PB2_OK;
}

int pb2_fetch_fault(pb2_ctx* ctx) {
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + PB2_FAULT_SLOT, ctx->d_counters + PB2_FAULT_SLOT, 8, cudaMemcpyDeviceToHost, ctx->stream));
    return PB2_OK;
}
int pb2_check_fault(pb2_ctx* ctx) {
    uint32_t f = *(uint32_t*)(ctx->h_counters + PB2_FAULT_SLOT);
    if (f == 0) return PB2_OK;
    ctx->h_counters[PB2_FAULT_SLOT] = 0;
    cudaMemsetAsync(ctx->d_counters + PB2_FAULT_SLOT, 0, 8, ctx->stream);
    if (f & PB2_FAULT_BAD_ID) PB2_FAIL(ctx, PB2_ERR_INVALID, "a shape id in a device-resident array was out of range: the leaf was skipped");
    PB2_FAIL(ctx, PB2_ERR_DEPTH, "traversal stack overflow (tree deeper than %d levels): results of the calls since the last synchronisation are incomplete", PB2_STACK);
}

int pb2_pipeline_init(pb2_ctx* ctx) {
    if (ctx->copy_in) return PB2_OK;
    PB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    PB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    PB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->compute2, cudaStreamNonBlocking));
    for (int i = 0; i < 6; ++i) PB2_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_peer[i], cudaStreamNonBlocking));
    for (int i = 0; i < 64; ++i) PB2_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev[i], cudaEventDisableTiming));
    PB2_CUDA(ctx, cudaMalloc((void**)&ctx->d_pieces, 64 * sizeof(unsigned int)));
    {   // stream memory operations come from the driver API; resolved through the runtime so that nothing links against libcuda
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess) ctx->wait_value32 = fn;
        (void)cudaGetLastError();
    }
    return PB2_OK;
}

extern "C" {

int pb2_version(void) { return 100; }

int pb2_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

static int ctx_init(int device, cudaStream_t stream, bool own, pb2_ctx** out) {
    if (!out) return PB2_ERR_INVALID;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return PB2_ERR_CUDA;
    if (cudaSetDevice(device) != cudaSuccess) return PB2_ERR_CUDA;
    pb2_ctx* ctx = new pb2_ctx();
    ctx->device = device;
    ctx->own_stream = own;
    if (own) {
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { delete ctx; return PB2_ERR_CUDA; }
    } else {
        ctx->stream = stream;
    }
    int sm = 0;
    if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && sm > 0) ctx->sm_count = sm;
    if (cudaMallocHost((void**)&ctx->h_counters, 16 * sizeof(uint64_t)) != cudaSuccess) { delete ctx; return PB2_ERR_CUDA; }
    if (cudaMalloc((void**)&ctx->d_counters, 16 * sizeof(uint64_t)) != cudaSuccess) { delete ctx; return PB2_ERR_CUDA; }
    if (cudaMemset(ctx->d_counters, 0, 16 * sizeof(uint64_t)) != cudaSuccess) { delete ctx; return PB2_ERR_CUDA; }
    memset(ctx->h_counters, 0, 16 * sizeof(uint64_t));
    {   // temporaries of the pair / candidate lists come from the stream-ordered pool: keep what it has grown to across
        // synchronisations (the default threshold of 0 gives the memory back at every sync and re-maps it on the next call,
        // which showed up as 3-8 ms of idle GPU per TriMesh-contact call); trimmed again in pb2_ctx_destroy
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
    }
    *out = ctx;
    return PB2_OK;
}

int pb2_ctx_create(int device, pb2_ctx** out) { return ctx_init(device, nullptr, true, out); }
int pb2_ctx_create_on_stream(int device, void* cuda_stream, pb2_ctx** out) {
    return ctx_init(device, (cudaStream_t)cuda_stream, false, out);
}

int pb2_ctx_destroy(pb2_ctx* ctx) {
    if (!ctx) return PB2_ERR_INVALID;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& s : ctx->stage) if (s.ptr) cudaFree(s.ptr);
    for (auto& s : ctx->scratch) if (s.ptr) cudaFree(s.ptr);
    { cudaMemPool_t pool; if (cudaDeviceGetDefaultMemPool(&pool, ctx->device) == cudaSuccess) cudaMemPoolTrimTo(pool, 0); }
    if (ctx->h_counters) cudaFreeHost(ctx->h_counters);
    if (ctx->d_counters) cudaFree(ctx->d_counters);
    if (ctx->d_pieces) cudaFree(ctx->d_pieces);
    if (ctx->epa_big_arena) cudaFree(ctx->epa_big_arena);
    for (int i = 0; i < 4; ++i) if (ctx->phase_ev[i]) cudaEventDestroy(ctx->phase_ev[i]);
    if (ctx->copy_in) { cudaStreamDestroy(ctx->copy_in); cudaStreamDestroy(ctx->copy_out); if (ctx->compute2) cudaStreamDestroy(ctx->compute2); for (int i = 0; i < 6; ++i) if (ctx->copy_peer[i]) cudaStreamDestroy(ctx->copy_peer[i]); for (int i = 0; i < 64; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]); }
    if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
    return PB2_OK;
}

int pb2_ctx_synchronize(pb2_ctx* ctx) {
    if (!ctx) return PB2_ERR_INVALID;
    PB2_CHECK(pb2_fetch_fault(ctx));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return pb2_check_fault(ctx);   // device-resident calls report a traversal-stack overflow here
}

int pb2_ctx_enable_phase_timing(pb2_ctx* ctx, int on) {
    if (!ctx) return PB2_ERR_INVALID;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    if (on) for (int i = 0; i < 4; ++i) if (!ctx->phase_ev[i]) PB2_CUDA(ctx, cudaEventCreate(&ctx->phase_ev[i]));
    ctx->phase_timing = on != 0;
    ctx->phase_marks = 0;
    return PB2_OK;
}

int pb2_contact_phase_times(pb2_ctx* ctx, float* gjk_ms, float* epa_ms, float* finish_ms, uint64_t* epa_runs) {
    if (!ctx) return PB2_ERR_INVALID;
    if (!ctx->phase_timing || ctx->phase_marks < 4) PB2_FAIL(ctx, PB2_ERR_INVALID, "no timed contact call since pb2_ctx_enable_phase_timing");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    PB2_CUDA(ctx, cudaMemcpyAsync(ctx->h_counters + 4, ctx->d_counters + 4, 8, cudaMemcpyDeviceToHost, ctx->stream));
    PB2_CUDA(ctx, cudaEventSynchronize(ctx->phase_ev[3]));
    PB2_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    float a = 0.f, b = 0.f, c = 0.f;
    PB2_CUDA(ctx, cudaEventElapsedTime(&a, ctx->phase_ev[0], ctx->phase_ev[1]));
    PB2_CUDA(ctx, cudaEventElapsedTime(&b, ctx->phase_ev[1], ctx->phase_ev[2]));
    PB2_CUDA(ctx, cudaEventElapsedTime(&c, ctx->phase_ev[2], ctx->phase_ev[3]));
    if (gjk_ms) *gjk_ms = a;
    if (epa_ms) *epa_ms = b;
    if (finish_ms) *finish_ms = c;
    if (epa_runs) *epa_runs = ctx->h_counters[4];
    return PB2_OK;
}

void* pb2_ctx_stream(pb2_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
const char* pb2_last_error(pb2_ctx* ctx) { return ctx ? ctx->err : "null ctx"; }
uint64_t pb2_ctx_launch_count(pb2_ctx* ctx) { return ctx ? ctx->launches : 0; }

}  // extern "C"

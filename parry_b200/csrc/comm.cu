// parry_b200 — multi-GPU exchange behind the C ABI ("pb2_comm_init / pb2_allgather_results").
//
// The query path shards with no data-path collective (rays / query leaves / candidate pairs are independent, the Bvh and the
// shape tables are replicated per GPU); what crosses GPUs is the gather of results: fixed-size per-ray records, and counts +
// compacted variable-size lists (pairs, contacts). Round 1 had these only as torch.distributed helpers, so a non-Python host
// could not use more than one GPU. Here they are plain C entry points over NCCL (one process per GPU, one pb2_comm per
// context): the host only has to carry 128 bytes (the NCCL unique id) from rank 0 to the other ranks by whatever transport it
// already has. NCCL is resolved with dlopen at the first pb2_comm_* call, so the library itself still loads where NCCL is absent.
//
// pb2_comm_peer_alloc hands out buffers that every rank of the node can write into (CUDA IPC mappings): the peer pointers
// pb2_trimesh_cast_rays_allgather pushes finished result ranges to with the copy engines while the traversal kernel is running.
#include "common.cuh"
#include <dlfcn.h>
#include <nccl.h>
#include <vector>

struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi* nccl_api() {
    static NcclApi api;
    static bool tried = false;
    if (tried) return api.lib ? &api : nullptr;
    tried = true;
    // RTLD_NOLOAD first: a host that already carries NCCL (e.g. PyTorch's bundled copy) keeps using that one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return nullptr;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) ok = false; return p; };
    api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
    api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
    api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
    api.Broadcast = (decltype(api.Broadcast))sym("ncclBroadcast");
    api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
    api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
    api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
    if (!ok) return nullptr;
    api.lib = h;
    return &api;
}

struct pb2_comm {
    pb2_ctx* ctx = nullptr;
    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;
    unsigned long long* d_words = nullptr;   // nranks + 2 u64: gathered counts, own count, barrier word
    unsigned long long* h_words = nullptr;   // pinned mirror
    std::vector<void*> opened;               // IPC mappings of peers' buffers (closed in pb2_comm_destroy / pb2_comm_peer_free)
    std::vector<void*> owned;                // buffers this rank allocated for pb2_comm_peer_alloc
};

#define PB2_NCCL(c, expr)                                                                                          \
    do {                                                                                                           \
        ncclResult_t _r = (expr);                                                                                  \
        if (_r != ncclSuccess) {                                                                                   \
            snprintf((c)->ctx->err, sizeof((c)->ctx->err), "%s:%d %s -> %s", __FILE__, __LINE__, #expr, nccl_api()->GetErrorString(_r)); \
            return PB2_ERR_CUDA;                                                                                   \
        }                                                                                                          \
    } while (0)

extern "C" {

int pb2_comm_unique_id(void* id128) {
    if (!id128) return PB2_ERR_INVALID;
    NcclApi* a = nccl_api();
    if (!a) return PB2_ERR_UNSUPPORTED;
    static_assert(sizeof(ncclUniqueId) == PB2_COMM_ID_BYTES, "ncclUniqueId is 128 bytes");
    ncclUniqueId id;
    if (a->GetUniqueId(&id) != ncclSuccess) return PB2_ERR_CUDA;
    memcpy(id128, &id, sizeof(id));
    return PB2_OK;
}

int pb2_comm_create(pb2_ctx* ctx, const void* id128, int rank, int nranks, pb2_comm** out) {
    if (!ctx || !id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) return PB2_ERR_INVALID;
    *out = nullptr;
    NcclApi* a = nccl_api();
    if (!a) PB2_FAIL(ctx, PB2_ERR_UNSUPPORTED, "libnccl.so.2 not found (dlopen)");
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    pb2_comm* c = new pb2_comm();
    c->ctx = ctx; c->rank = rank; c->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, id128, sizeof(id));
    ncclResult_t r = a->CommInitRank(&c->comm, nranks, id, rank);
    if (r != ncclSuccess) { snprintf(ctx->err, sizeof(ctx->err), "ncclCommInitRank: %s", a->GetErrorString(r)); delete c; return PB2_ERR_CUDA; }
    size_t words = (size_t)nranks + 2;
    if (cudaMalloc((void**)&c->d_words, words * 8) != cudaSuccess || cudaMallocHost((void**)&c->h_words, words * 8) != cudaSuccess) {
        snprintf(ctx->err, sizeof(ctx->err), "pb2_comm_create: allocation failed");
        a->CommDestroy(c->comm); delete c; return PB2_ERR_CUDA;
    }
    cudaMemset(c->d_words, 0, words * 8);
    *out = c;
    return PB2_OK;
}

int pb2_comm_destroy(pb2_comm* c) {
    if (!c) return PB2_ERR_INVALID;
    cudaSetDevice(c->ctx->device);
    cudaStreamSynchronize(c->ctx->stream);
    for (void* p : c->opened) cudaIpcCloseMemHandle(p);
    for (void* p : c->owned) cudaFree(p);
    if (c->d_words) cudaFree(c->d_words);
    if (c->h_words) cudaFreeHost(c->h_words);
    if (c->comm) nccl_api()->CommDestroy(c->comm);
    delete c;
    return PB2_OK;
}

int pb2_comm_rank(const pb2_comm* c) { return c ? c->rank : -1; }
int pb2_comm_size(const pb2_comm* c) { return c ? c->nranks : 0; }

// Fixed-size all-gather of device buffers (per-ray records): recv holds nranks * bytes_per_rank, rank-major. Enqueued on the
// context's stream; send may be recv + rank * bytes_per_rank (in place).
int pb2_comm_allgather(pb2_comm* c, const void* send, void* recv, uint64_t bytes_per_rank) {
    if (!c || (bytes_per_rank && (!send || !recv))) return PB2_ERR_INVALID;
    if (bytes_per_rank == 0) return PB2_OK;
    PB2_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    PB2_NCCL(c, nccl_api()->AllGather(send, recv, (size_t)bytes_per_rank, ncclChar, c->comm, c->ctx->stream));
    c->ctx->launches++;
    return PB2_OK;
}

// "all-gather of compacted hit / pair counts": one u64 per rank, returned on the host (synchronises the context's stream).
int pb2_comm_allgather_counts(pb2_comm* c, uint64_t mine, uint64_t* all) {
    if (!c || !all) return PB2_ERR_INVALID;
    pb2_ctx* ctx = c->ctx;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = ctx->stream;
    c->h_words[c->nranks] = mine;
    PB2_CUDA(ctx, cudaMemcpyAsync(c->d_words + c->nranks, c->h_words + c->nranks, 8, cudaMemcpyHostToDevice, st));
    PB2_NCCL(c, nccl_api()->AllGather(c->d_words + c->nranks, c->d_words, 8, ncclChar, c->comm, st));
    ctx->launches++;
    PB2_CUDA(ctx, cudaMemcpyAsync(c->h_words, c->d_words, (size_t)c->nranks * 8, cudaMemcpyDeviceToHost, st));
    PB2_CUDA(ctx, cudaStreamSynchronize(st));
    for (int r = 0; r < c->nranks; ++r) all[r] = c->h_words[r];
    return PB2_OK;
}

// Variable-size gather of compacted records: every rank contributes `count` elements of `elem_bytes`; afterwards recv holds all
// ranks' elements back to back in rank order (no padding), counts[r] (host) says how many came from rank r. One grouped set of
// broadcasts, one per rank, each of exactly that rank's bytes. Capacity is checked before anything is written.
int pb2_comm_allgatherv(pb2_comm* c, const void* send, uint64_t count, uint32_t elem_bytes, void* recv, uint64_t cap_elems,
                        uint64_t* counts, uint64_t* total) {
    if (!c || !counts || !total || elem_bytes == 0 || (count && !send)) return PB2_ERR_INVALID;
    PB2_CHECK(pb2_comm_allgather_counts(c, count, counts));
    uint64_t sum = 0;
    for (int r = 0; r < c->nranks; ++r) sum += counts[r];
    *total = sum;
    if (sum > cap_elems || (sum && !recv)) PB2_FAIL(c->ctx, PB2_ERR_OVERFLOW, "allgatherv: %llu elements > capacity %llu", (unsigned long long)sum, (unsigned long long)cap_elems);
    NcclApi* a = nccl_api();
    PB2_NCCL(c, a->GroupStart());
    uint64_t off = 0;
    for (int r = 0; r < c->nranks; ++r) {
        size_t bytes = (size_t)counts[r] * elem_bytes;
        if (bytes) {
            char* dst = (char*)recv + off * elem_bytes;
            ncclResult_t rr = a->Broadcast(r == c->rank ? send : (const void*)dst, dst, bytes, ncclChar, r, c->comm, c->ctx->stream);
            if (rr != ncclSuccess) { a->GroupEnd(); snprintf(c->ctx->err, sizeof(c->ctx->err), "ncclBroadcast: %s", a->GetErrorString(rr)); return PB2_ERR_CUDA; }
        }
        off += counts[r];
    }
    PB2_NCCL(c, a->GroupEnd());
    c->ctx->launches++;
    return PB2_OK;
}

// Cross-rank barrier on the context's stream (an all-reduce of one word): work enqueued after it starts once every rank has
// reached it.
int pb2_comm_barrier(pb2_comm* c) {
    if (!c) return PB2_ERR_INVALID;
    PB2_CUDA(c->ctx, cudaSetDevice(c->ctx->device));
    unsigned long long* w = c->d_words + c->nranks + 1;
    PB2_NCCL(c, nccl_api()->AllReduce(w, w, 1, ncclUint64, ncclSum, c->comm, c->ctx->stream));
    c->ctx->launches++;
    return PB2_OK;
}

// Allocates `bytes` on this rank and maps every other rank's allocation of the same call into this process (CUDA IPC; all ranks
// on one node, peer access over NVLink). peers[r] is usable in copies and kernels of this process; peers[rank] is the local
// buffer. Collective: every rank calls it with the same size.
int pb2_comm_peer_alloc(pb2_comm* c, uint64_t bytes, void** peers) {
    if (!c || !peers || bytes == 0) return PB2_ERR_INVALID;
    pb2_ctx* ctx = c->ctx;
    PB2_CUDA(ctx, cudaSetDevice(ctx->device));
    void* mine = nullptr;
    PB2_CUDA(ctx, cudaMalloc(&mine, (size_t)bytes));
    c->owned.push_back(mine);
    cudaIpcMemHandle_t h;
    PB2_CUDA(ctx, cudaIpcGetMemHandle(&h, mine));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    char *d_all = nullptr, *h_all = (char*)malloc((size_t)c->nranks * 64);
    PB2_CUDA(ctx, cudaMalloc((void**)&d_all, (size_t)(c->nranks + 1) * 64));
    PB2_CUDA(ctx, cudaMemcpyAsync(d_all + (size_t)c->nranks * 64, &h, 64, cudaMemcpyHostToDevice, ctx->stream));
    int rc = PB2_OK;
    ncclResult_t r = nccl_api()->AllGather(d_all + (size_t)c->nranks * 64, d_all, 64, ncclChar, c->comm, ctx->stream);
    if (r != ncclSuccess) { snprintf(ctx->err, sizeof(ctx->err), "peer_alloc all-gather: %s", nccl_api()->GetErrorString(r)); rc = PB2_ERR_CUDA; }
    if (rc == PB2_OK && cudaMemcpyAsync(h_all, d_all, (size_t)c->nranks * 64, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) rc = PB2_ERR_CUDA;
    if (rc == PB2_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = PB2_ERR_CUDA;
    for (int p = 0; p < c->nranks && rc == PB2_OK; ++p) {
        if (p == c->rank) { peers[p] = mine; continue; }
        cudaIpcMemHandle_t hp;
        memcpy(&hp, h_all + (size_t)p * 64, 64);
        void* q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, hp, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { snprintf(ctx->err, sizeof(ctx->err), "cudaIpcOpenMemHandle(rank %d): %s", p, cudaGetErrorString(e)); rc = PB2_ERR_CUDA; break; }
        c->opened.push_back(q);
        peers[p] = q;
    }
    cudaFree(d_all);
    free(h_all);
    if (rc != PB2_OK) return rc;
    return pb2_comm_barrier(c);   // nobody writes into a peer before everybody has mapped everything
}

}  // extern "C"

// Microbenchmark (measurement tool, not product): cost of fetching one BVH node per lane from random places, as the wide
// ray kernel does — 5 x LDG.128 on 80-byte nodes (round 1 layout) against 96- and 64-byte nodes fetched with 128- or 256-bit loads
// (LDG.E.ENL2.256, sm_100+). Working sets: L1-sized, L2-sized, HBM-sized. Same residency as the ray kernel: 7 CTAs x 128 threads per SM.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint32_t hash32(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }

template <int STRIDE16, int NLOAD, bool WIDE>
__global__ void __launch_bounds__(128, 7) k_fetch(const float4* __restrict__ base, uint32_t n_nodes, int iters, float* out) {
    uint32_t s = hash32(blockIdx.x * blockDim.x + threadIdx.x + 12345u);
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        uint32_t node = s % n_nodes;
        const float4* p = base + (size_t)node * STRIDE16;
        float v = 0.f;
        if (WIDE) {
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) {
                float a0, a1, a2, a3, a4, a5, a6, a7;
                asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(a0), "=f"(a1), "=f"(a2), "=f"(a3), "=f"(a4), "=f"(a5), "=f"(a6), "=f"(a7) : "l"(p + 2 * k));
                v += a0 + a7 + a3;
            }
        } else {
#pragma unroll
            for (int k = 0; k < NLOAD; ++k) { float4 q = __ldg(p + k); v += q.x + q.w; }
        }
        acc += v;
        s = hash32(s + __float_as_uint(v));   // next address depends on the data, like a tree walk
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

template <int STRIDE16, int NLOAD, bool WIDE>
static void run(const char* name, const float4* buf, size_t bytes, int iters, float* out, int blocks) {
    uint32_t n_nodes = (uint32_t)(bytes / (16 * STRIDE16));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k_fetch<STRIDE16, NLOAD, WIDE><<<blocks, 128>>>(buf, n_nodes, iters, out);
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        k_fetch<STRIDE16, NLOAD, WIDE><<<blocks, 128>>>(buf, n_nodes, iters, out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    double fetches = (double)blocks * 128 * iters;
    printf("  %-34s %8.3f ms  %7.2f Gnodes/s\n", name, best, fetches / best / 1e6);
}

int main() {
    int sm = 148;
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
    int blocks = sm * 7;
    size_t cap = (size_t)768 << 20;
    float4* buf; float* out;
    cudaMalloc(&buf, cap); cudaMalloc(&out, (size_t)blocks * 128 * 4);
    cudaMemset(buf, 0, cap);
    struct { const char* name; size_t bytes; int iters; } sets[] = {{"L1-sized 48 KB", 48 << 10, 4000}, {"L2-sized 32 MB", 32 << 20, 2000}, {"HBM-sized 768 MB", cap, 1000}};
    for (auto& s : sets) {
        printf("%s\n", s.name);
        run<5, 5, false>("80 B node, 5 x LDG.128", buf, s.bytes, s.iters, out, blocks);
        run<6, 6, false>("96 B node, 6 x LDG.128", buf, s.bytes, s.iters, out, blocks);
        run<6, 3, true>("96 B node, 3 x LDG.256", buf, s.bytes, s.iters, out, blocks);
        run<4, 4, false>("64 B node, 4 x LDG.128", buf, s.bytes, s.iters, out, blocks);
        run<4, 2, true>("64 B node, 2 x LDG.256", buf, s.bytes, s.iters, out, blocks);
        run<8, 4, true>("128 B node, 4 x LDG.256", buf, s.bytes, s.iters, out, blocks);
    }
    return 0;
}

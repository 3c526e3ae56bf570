// gather_bench.cu — what the B200 memory pipes give for the access patterns K1 can be built from (design input,
// not product code): random 128-B table-line gathers by (a) 8 lanes per line LDG.128, (b) 4 lanes per line,
// (c) per-thread TMA bulk copies into shared memory, (d) LDGSTS 16-B copies into shared memory, (e) thread-per-line
// LDG.128 x 8; and (f) three fire-and-forget reductions per index onto random 32-B lines; (g) = (a)+(f) together.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gather_bench gather_bench.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint64_t mix(uint64_t x) { x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33; return x; }

__global__ void fill_idx(uint32_t* idx, size_t n, uint32_t mask, int zipfish) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t h = mix(i * 0x9E3779B97F4A7C15ull + 12345);
        idx[i] = (uint32_t)h & mask;
    }
}

// (a) 8 lanes per line
__global__ void __launch_bounds__(1024, 1) k_ldg8(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const size_t nwarp = (size_t)gridDim.x * (blockDim.x >> 5), w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (size_t base = w * 32; base + 32 <= n; base += nwarp * 32) {
        const uint32_t mine = idx[base + lane];
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t slot = __shfl_sync(0xFFFFFFFFu, mine, r * 4 + g);
            v[r] = __ldcg(&tab[(size_t)slot * 8 + j]);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) acc ^= v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
// (b) 4 lanes per line, 2 loads per lane
__global__ void __launch_bounds__(1024, 1) k_ldg4(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    const int lane = threadIdx.x & 31, g = lane >> 2, j = lane & 3;
    const size_t nwarp = (size_t)gridDim.x * (blockDim.x >> 5), w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (size_t base = w * 32; base + 32 <= n; base += nwarp * 32) {
        const uint32_t mine = idx[base + lane];
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 4; r++) {
            const uint32_t slot = __shfl_sync(0xFFFFFFFFu, mine, r * 8 + g);
            v[2 * r] = __ldcg(&tab[(size_t)slot * 8 + j]);
            v[2 * r + 1] = __ldcg(&tab[(size_t)slot * 8 + j + 4]);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) acc ^= v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
// (e) one thread per line, 8 loads per thread
__global__ void __launch_bounds__(1024, 1) k_ldg1(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (size_t i = t0; i < n; i += nthr) {
        const uint32_t slot = idx[i];
        uint4 v[8];
#pragma unroll
        for (int c = 0; c < 8; c++) v[c] = __ldcg(&tab[(size_t)slot * 8 + c]);
#pragma unroll
        for (int r = 0; r < 8; r++) acc ^= v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
    }
    if (acc == 0x12345678u) *sink = acc;
}
// (c) per-thread TMA bulk copy of the 128-B line into the warp's shared-memory buffer (kStages deep)
template <int kStages, int kBytes>
__global__ void __launch_bounds__(1024, 1) k_tma(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint4* buf = reinterpret_cast<uint4*>(smem) + (size_t)warp * kStages * 32 * 8;
    unsigned long long* bars = reinterpret_cast<unsigned long long*>(smem + (size_t)(blockDim.x >> 5) * kStages * 32 * 128) + warp * kStages;
    if (lane == 0) for (int s = 0; s < kStages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bars[s])), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const size_t nwarp = (size_t)gridDim.x * (blockDim.x >> 5), w = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t acc = 0;
    size_t issue = w * 32;
    int it_issue = 0, it_wait = 0;
    auto do_issue = [&](size_t base, int st) {
        const uint32_t slot = idx[base + lane];
        if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bars[st])), "r"(32 * kBytes) : "memory");
        __syncwarp();
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(smem_u32(&buf[(st * 32 + lane) * 8])), "l"(tab + (size_t)slot * 8), "r"(kBytes), "r"(smem_u32(&bars[st])) : "memory");
    };
    for (int s = 0; s < kStages - 1; s++) { if (issue + 32 <= n) { do_issue(issue, it_issue % kStages); it_issue++; issue += nwarp * 32; } }
    for (size_t base = w * 32; base + 32 <= n; base += nwarp * 32) {
        if (issue + 32 <= n) { do_issue(issue, it_issue % kStages); it_issue++; issue += nwarp * 32; }
        const int st = it_wait % kStages; const uint32_t par = (it_wait / kStages) & 1; it_wait++;
        asm volatile("{\n\t.reg .pred p;\n\tW_%=: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(&bars[st])), "r"(par) : "memory");
#pragma unroll
        for (int c = 0; c < kBytes / 16; c++) { const uint4 v = buf[(st * 32 + lane) * 8 + c]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (acc == 0x12345678u) *sink = acc;
}
// (d) LDGSTS: 8 lanes per line, 16-B cp.async into shared memory, thread-per-line consumption
template <int kStages>
__global__ void __launch_bounds__(1024, 1) k_ldgsts(const uint4* __restrict__ tab, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    extern __shared__ __align__(128) uint8_t smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 3, j = lane & 7;
    uint4* buf = reinterpret_cast<uint4*>(smem) + (size_t)warp * kStages * 32 * 8;
    const size_t nwarp = (size_t)gridDim.x * (blockDim.x >> 5), w = (size_t)blockIdx.x * (blockDim.x >> 5) + warp;
    uint32_t acc = 0;
    size_t issue = w * 32;
    int it_issue = 0, it_wait = 0;
    auto do_issue = [&](size_t base, int st) {
        const uint32_t mine = idx[base + lane];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t slot = __shfl_sync(0xFFFFFFFFu, mine, r * 4 + g);
            // swizzle the 16-B chunk position by the record index so that thread-per-line reads are conflict-free
            const int rec = r * 4 + g;
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(&buf[(st * 32 + rec) * 8 + (j ^ (rec & 7))])), "l"(tab + (size_t)slot * 8 + j) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    for (int s = 0; s < kStages - 1; s++) { if (issue + 32 <= n) { do_issue(issue, it_issue % kStages); issue += nwarp * 32; } else asm volatile("cp.async.commit_group;" ::: "memory"); it_issue++; }
    for (size_t base = w * 32; base + 32 <= n; base += nwarp * 32) {
        if (issue + 32 <= n) { do_issue(issue, it_issue % kStages); issue += nwarp * 32; } else asm volatile("cp.async.commit_group;" ::: "memory");
        it_issue++;
        asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
        __syncwarp();
        const int st = it_wait % kStages; it_wait++;
#pragma unroll
        for (int c = 0; c < 8; c++) { const uint4 v = buf[(st * 32 + lane) * 8 + (c ^ (lane & 7))]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
        __syncwarp();
    }
    if (acc == 0x12345678u) *sink = acc;
}
// (f) three reductions per index on a random 32-byte line (u64 add, u64 max, u32 add)
template <int kReds>
__global__ void __launch_bounds__(1024, 1) k_red(uint8_t* hot, const uint32_t* __restrict__ idx, size_t n) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (size_t i = t0; i < n; i += nthr) {
        const uint32_t slot = idx[i];
        uint8_t* h = hot + (size_t)slot * 32;
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(h), "l"(1500ull) : "memory");
        if (kReds > 1) asm volatile("red.global.max.u64 [%0], %1;" ::"l"(h + 16), "l"((unsigned long long)i) : "memory");
        if (kReds > 2) asm volatile("red.global.add.u32 [%0], %1;" ::"l"(h + 24), "r"(1u) : "memory");
    }
}
// (g) gather (8 lanes per line) + 3 reductions by one lane per line
__global__ void __launch_bounds__(1024, 1) k_ldg8_red(const uint4* __restrict__ tab, uint8_t* hot, const uint32_t* __restrict__ idx, size_t n, uint32_t* sink) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const size_t nwarp = (size_t)gridDim.x * (blockDim.x >> 5), w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint32_t acc = 0;
    for (size_t base = w * 32; base + 32 <= n; base += nwarp * 32) {
        const uint32_t mine = idx[base + lane];
        uint4 v[8];
#pragma unroll
        for (int r = 0; r < 8; r++) {
            const uint32_t slot = __shfl_sync(0xFFFFFFFFu, mine, r * 4 + g);
            v[r] = __ldcg(&tab[(size_t)slot * 8 + j]);
        }
#pragma unroll
        for (int r = 0; r < 8; r++) acc ^= v[r].x ^ v[r].y ^ v[r].z ^ v[r].w;
        uint8_t* h = hot + (size_t)mine * 32;
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(h), "l"(1500ull + (acc & 1)) : "memory");
        asm volatile("red.global.max.u64 [%0], %1;" ::"l"(h + 16), "l"((unsigned long long)base) : "memory");
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(h + 24), "r"(1u) : "memory");
    }
    if (acc == 0x12345678u) *sink = acc;
}
// stream read baseline: plain coalesced 16-B loads of n*144 bytes
__global__ void __launch_bounds__(1024, 1) k_stream(const uint4* __restrict__ src, size_t n16, uint32_t* sink) {
    const size_t nthr = (size_t)gridDim.x * blockDim.x, t0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t acc = 0;
    for (size_t i = t0; i < n16; i += nthr) { const uint4 v = __ldcs(&src[i]); acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F> static float timeit(F f, int reps = 5) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    f(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) { cudaEventRecord(a); f(); cudaEventRecord(b); CK(cudaEventSynchronize(b)); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}

int main(int argc, char** argv) {
    const int log2slots = argc > 1 ? atoi(argv[1]) : 20;
    const size_t n = (size_t)1 << (argc > 2 ? atoi(argv[2]) : 25);
    const size_t slots = (size_t)1 << log2slots;
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    uint4* tab; uint8_t* hot; uint32_t* idx; uint32_t* sink; uint4* stream;
    CK(cudaMalloc(&tab, slots * 128)); CK(cudaMalloc(&hot, slots * 32)); CK(cudaMalloc(&idx, n * 4)); CK(cudaMalloc(&sink, 4));
    CK(cudaMalloc(&stream, n * 144));
    CK(cudaMemset(tab, 1, slots * 128)); CK(cudaMemset(hot, 0, slots * 32)); CK(cudaMemset(stream, 1, n * 144));
    fill_idx<<<sms * 4, 1024>>>(idx, n, (uint32_t)(slots - 1), 0);
    CK(cudaDeviceSynchronize());
    printf("{\"slots_log2\": %d, \"table_MB\": %.0f, \"n\": %zu, \"sms\": %d", log2slots, slots * 160 / 1e6, n, sms);
    auto rep = [&](const char* name, float ms) { printf(", \"%s_Glines_s\": %.2f", name, n / (ms * 1e-3) / 1e9); fflush(stdout); };
    rep("ldg8", timeit([&] { k_ldg8<<<sms, 1024>>>(tab, idx, n, sink); }));
    rep("ldg4", timeit([&] { k_ldg4<<<sms, 1024>>>(tab, idx, n, sink); }));
    rep("ldg1", timeit([&] { k_ldg1<<<sms, 1024>>>(tab, idx, n, sink); }));
    {
        const int W = 16;                                  // warps per CTA for the shared-memory staged variants
        const int sm2 = W * 2 * 32 * 128 + 1024, sm3 = W * 3 * 32 * 128 + 1024;
        CK(cudaFuncSetAttribute(k_tma<2, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
        rep("tma128_2st", timeit([&] { k_tma<2, 128><<<sms, W * 32, sm2>>>(tab, idx, n, sink); }));
        CK(cudaFuncSetAttribute(k_tma<3, 128>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm3));
        rep("tma128_3st", timeit([&] { k_tma<3, 128><<<sms, W * 32, sm3>>>(tab, idx, n, sink); }));
        CK(cudaFuncSetAttribute(k_tma<3, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm3));
        rep("tma64_3st", timeit([&] { k_tma<3, 64><<<sms, W * 32, sm3>>>(tab, idx, n, sink); }));
        CK(cudaFuncSetAttribute(k_ldgsts<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm2));
        rep("ldgsts_2st", timeit([&] { k_ldgsts<2><<<sms, W * 32, sm2>>>(tab, idx, n, sink); }));
        CK(cudaFuncSetAttribute(k_ldgsts<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm3));
        rep("ldgsts_3st", timeit([&] { k_ldgsts<3><<<sms, W * 32, sm3>>>(tab, idx, n, sink); }));
    }
    rep("red1", timeit([&] { k_red<1><<<sms, 1024>>>(hot, idx, n); }));
    rep("red3", timeit([&] { k_red<3><<<sms, 1024>>>(hot, idx, n); }));
    rep("red3_2cta", timeit([&] { k_red<3><<<sms * 2, 1024>>>(hot, idx, n); }));
    rep("ldg8_red3", timeit([&] { k_ldg8_red<<<sms, 1024>>>(tab, hot, idx, n, sink); }));
    {
        float ms = timeit([&] { k_stream<<<sms * 2, 1024>>>(stream, n * 9, sink); });
        printf(", \"stream_GBs\": %.0f", n * 144 / (ms * 1e-3) / 1e9);
    }
    printf("}\n");
    return 0;
}

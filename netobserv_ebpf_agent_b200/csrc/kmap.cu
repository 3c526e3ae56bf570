// kmap.cu — KERNEL_MAP mode kernels: one thread per index around the bodies of kmap_body.cuh.
// HBM-bound integer work: each pass streams the batch once (resolve reads the 40-byte key, the other passes a
// few metric fields) and touches one or two table lines per record with L2 atomics.
#include "kernels.cuh"
#include "kmap_body.cuh"

namespace fa {

constexpr int kKmThreads = 256;

__global__ void __launch_bounds__(kKmThreads) km_resolve_kernel(KmParams P) {
    for (uint32_t i = P.lo + blockIdx.x * blockDim.x + threadIdx.x; i < P.hi; i += gridDim.x * blockDim.x) km_resolve_body(P, i);
}
__global__ void __launch_bounds__(kKmThreads) km_init_kernel(KmParams P) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) km_init_body(P, i);
}
__global__ void __launch_bounds__(kKmThreads) km_fold_kernel(KmParams P) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) km_fold_body(P, i);
}
__global__ void __launch_bounds__(kKmThreads) km_bresolve_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->bset_count;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) km_bresolve_body(P, j);
}
__global__ void __launch_bounds__(kKmThreads) km_order_kernel(KmParams P) {
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) km_order_body(P, i);
}
__global__ void __launch_bounds__(kKmThreads) km_cleanup_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->bset_count;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < P.n; i += gridDim.x * blockDim.x) km_cleanup_record_body(P, i);
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) km_cleanup_bset_body(P, j);
}
__global__ void km_reset_bset_count_kernel(KmCounters* c) { c->bset_count = 0ull; }

// ---- v2 (kmap_body.cuh, second half): one streaming pass, the rest over compact lists -------------------------
__global__ void __launch_bounds__(kKmThreads) km2_resolve_fold_kernel(KmParams P) {
    for (uint32_t i = P.lo + blockIdx.x * blockDim.x + threadIdx.x; i < P.hi; i += gridDim.x * blockDim.x) km2_resolve_fold_body(P, i);
}
__global__ void __launch_bounds__(kKmThreads) km2_init_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->deferred_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) km2_init_body(P, k);
}
__global__ void __launch_bounds__(kKmThreads) km2_fold_deferred_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->deferred_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) km2_fold_deferred_body(P, k);
}
__global__ void __launch_bounds__(kKmThreads) km2_order_b_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->brec_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) km2_order_b_body(P, k);
}
__global__ void __launch_bounds__(kKmThreads) km2_finish_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->touched_count;
    for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < m; k += gridDim.x * blockDim.x) km2_finish_flow_body(P, k);
}
__global__ void __launch_bounds__(kKmThreads) km2_cleanup_kernel(KmParams P) {
    const uint32_t m = (uint32_t)P.c->bset_count;
    for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += gridDim.x * blockDim.x) km_cleanup_bset_body(P, j);
}
__global__ void km2_reset_counts_kernel(KmCounters* c) {
    c->bset_count = 0ull; c->touched_count = 0ull; c->deferred_count = 0ull; c->brec_count = 0ull;
}

__global__ void __launch_bounds__(kKmThreads) km_evict_kernel(Table t, uint8_t* met, uint8_t* out, unsigned long long cap,
                                                               unsigned long long* cursor, uint32_t* slot_of_out) {
    const uint64_t words = (t.mask + 1) >> 5;
    for (uint64_t w = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; w < words; w += (uint64_t)gridDim.x * blockDim.x)
        km_evict_word_body(t, met, (uint32_t)w, out, cap, cursor, slot_of_out);
}

#ifndef FA_HOST_EMUL
// One batch: records [0, cut) may create flows, records [cut, n) find the map full.
int launch_kmap_batch(KmParams P, uint32_t cut, int sm_count, cudaStream_t st) {
    if (!P.n) return 0;
    const int grid = sm_count * 8;
    int launches = 0;
    if (cut > 0) { P.lo = 0; P.hi = cut; P.allow_insert = 1; km_resolve_kernel<<<grid, kKmThreads, 0, st>>>(P); launches++; }
    if (cut < P.n) { P.lo = cut; P.hi = P.n; P.allow_insert = 0; km_resolve_kernel<<<grid, kKmThreads, 0, st>>>(P); launches++; }
    km_init_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_fold_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_bresolve_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_order_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_cleanup_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_reset_bset_count_kernel<<<1, 1, 0, st>>>(P.c);
    return launches + 6;
}

// v2 of the same batch: resolve+fold over the records, everything else over the lists it leaves behind.
int launch_kmap_batch_v2(KmParams P, uint32_t cut, int sm_count, cudaStream_t st) {
    if (!P.n) return 0;
    const int grid = sm_count * 8;
    int launches = 0;
    if (cut > 0) { P.lo = 0; P.hi = cut; P.allow_insert = 1; km2_resolve_fold_kernel<<<grid, kKmThreads, 0, st>>>(P); launches++; }
    if (cut < P.n) { P.lo = cut; P.hi = P.n; P.allow_insert = 0; km2_resolve_fold_kernel<<<grid, kKmThreads, 0, st>>>(P); launches++; }
    km2_init_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km2_fold_deferred_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km_bresolve_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km2_order_b_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km2_finish_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km2_cleanup_kernel<<<grid, kKmThreads, 0, st>>>(P);
    km2_reset_counts_kernel<<<1, 1, 0, st>>>(P.c);
    return launches + 7;
}

int launch_kmap_evict(const Table& t, uint8_t* met, uint8_t* out, unsigned long long cap, unsigned long long* cursor,
                      uint32_t* slot_of_out, int sm_count, cudaStream_t st) {
    km_evict_kernel<<<sm_count * 8, kKmThreads, 0, st>>>(t, met, out, cap, cursor, slot_of_out);
    return 1;
}
#endif  // FA_HOST_EMUL

}  // namespace fa

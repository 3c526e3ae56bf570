// simt.h — a small host emulation of the CUDA execution model (test infrastructure only).
//
// Every CUDA thread of a launch is an OS thread; threadIdx / blockIdx / blockDim / gridDim, __syncthreads, named
// barriers, the warp collectives (ballot / shfl / any / all / reduce / syncwarp), atomics, fences and the handful of
// PTX helpers of csrc/common.cuh (mbarrier + 1-D bulk copy, ld.cg / st.cg, red.*) are mapped to host equivalents, so
// that a kernel's source can be compiled by g++ and run against the oracle at small sizes.  It checks the kernel's
// LOGIC (including its inter-thread protocol under real preemptive concurrency); it says nothing about the device
// memory model, register pressure or speed.  Used by tests/emul/k1_emul.cpp.
#pragma once
#include <cuda_runtime.h>   // vector types (uint4, make_uint4 ...) and the host-side no-op definitions of __global__ etc.

#include <sched.h>

#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

namespace simt {

class Barrier {                     // reusable counting barrier (generation based)
public:
    explicit Barrier(unsigned n) : n_(n) {}
    void wait() {
        std::unique_lock<std::mutex> lk(m_);
        const unsigned gen = gen_;
        if (++count_ == n_) { count_ = 0; gen_++; cv_.notify_all(); return; }
        cv_.wait(lk, [&] { return gen_ != gen; });
    }
private:
    std::mutex m_; std::condition_variable cv_;
    unsigned n_, count_ = 0, gen_ = 0;
};

struct Warp {
    Barrier bar{32};
    unsigned long long v[32];
    std::mutex sub_mu;                                   // barriers of the lane subsets used by masked collectives
    std::map<unsigned, std::unique_ptr<Barrier>> sub;
};
struct Cta {
    std::unique_ptr<Barrier> sync_all;
    std::mutex named_mu;
    std::map<int, std::unique_ptr<Barrier>> named;
    std::vector<std::unique_ptr<Warp>> warps;
    uint8_t* smem = nullptr;
};
struct Dim { unsigned x, y, z; };
struct Ctx {
    Dim tid{0, 0, 0}, bid{0, 0, 0}, bdim{1, 1, 1}, gdim{1, 1, 1};
    int lane = 0;
    Warp* warp = nullptr;
    Cta* cta = nullptr;
};
inline thread_local Ctx* g_ctx = nullptr;
inline Ctx& ctx() { return *g_ctx; }

// Run `body` once per CUDA thread of a grid x block launch (block a multiple of 32), all CTAs concurrently.
inline void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()>& body) {
    if (block % 32) abort();
    std::vector<std::unique_ptr<Cta>> ctas;
    std::vector<void*> smem_raw;
    for (unsigned b = 0; b < grid; b++) {
        auto c = std::make_unique<Cta>();
        c->sync_all = std::make_unique<Barrier>(block);
        for (unsigned w = 0; w < block / 32; w++) c->warps.push_back(std::make_unique<Warp>());
        void* p = nullptr;
        if (posix_memalign(&p, 1024, smem_bytes ? smem_bytes : 1024)) abort();
        memset(p, 0xA5, smem_bytes ? smem_bytes : 1024);       // shared memory starts uninitialised on the device
        c->smem = static_cast<uint8_t*>(p);
        smem_raw.push_back(p);
        ctas.push_back(std::move(c));
    }
    std::vector<std::thread> threads;
    threads.reserve((size_t)grid * block);
    for (unsigned b = 0; b < grid; b++)
        for (unsigned t = 0; t < block; t++)
            threads.emplace_back([&, b, t] {
                Ctx c;
                c.tid = {t, 0, 0}; c.bid = {b, 0, 0}; c.bdim = {block, 1, 1}; c.gdim = {grid, 1, 1};
                c.lane = (int)(t & 31); c.cta = ctas[b].get(); c.warp = ctas[b]->warps[t >> 5].get();
                g_ctx = &c;
                body();
                g_ctx = nullptr;
            });
    for (auto& th : threads) th.join();
    for (void* p : smem_raw) free(p);
}

// path counters a kernel may bump through FA_EMUL_COUNT (reset / read by the harness)
inline unsigned long long g_counts[8] = {};
inline void count(int which, unsigned long long n) { __atomic_fetch_add(&g_counts[which & 7], n, __ATOMIC_RELAXED); }

inline void named_barrier(int id, unsigned count) {
    Cta& c = *ctx().cta;
    Barrier* b;
    {
        std::lock_guard<std::mutex> lk(c.named_mu);
        auto& slot = c.named[id];
        if (!slot) slot = std::make_unique<Barrier>(count);
        b = slot.get();
    }
    b->wait();
}

// all 32 lanes exchange one 64-bit value; every lane sees all of them
inline void exchange(unsigned long long mine, unsigned long long out[32]) {
    Warp& w = *ctx().warp;
    w.v[ctx().lane] = mine;
    w.bar.wait();
    for (int l = 0; l < 32; l++) out[l] = w.v[l];
    w.bar.wait();
}

// the lanes of `mask` (every one of them calls this with the same mask) exchange one 64-bit value
inline void exchange_masked(unsigned mask, unsigned long long mine, unsigned long long out[32]) {
    if (mask == 0xFFFFFFFFu) { exchange(mine, out); return; }
    Warp& w = *ctx().warp;
    if (!((mask >> ctx().lane) & 1u)) abort();
    Barrier* b;
    {
        std::lock_guard<std::mutex> lk(w.sub_mu);
        auto& slot = w.sub[mask];
        if (!slot) slot = std::make_unique<Barrier>((unsigned)__builtin_popcount(mask));
        b = slot.get();
    }
    w.v[ctx().lane] = mine;
    b->wait();
    for (int l = 0; l < 32; l++) out[l] = ((mask >> l) & 1u) ? w.v[l] : 0ull;
    b->wait();
}

}  // namespace simt

// ---------------------------------------------------------------------------------------------- CUDA spellings
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#define threadIdx (simt::ctx().tid)
#define blockIdx  (simt::ctx().bid)
#define blockDim  (simt::ctx().bdim)
#define gridDim   (simt::ctx().gdim)

inline void __syncthreads() { simt::ctx().cta->sync_all->wait(); }
inline void __syncwarp(unsigned = 0xFFFFFFFFu) { simt::ctx().warp->bar.wait(); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline long long clock64() { return 0; }

inline unsigned __ballot_sync(unsigned mask, int pred) {
    if (mask != 0xFFFFFFFFu) abort();
    unsigned long long v[32]; simt::exchange(pred ? 1ull : 0ull, v);
    unsigned r = 0; for (int l = 0; l < 32; l++) r |= (unsigned)v[l] << l;
    return r;
}
inline int __any_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) != 0u; }
inline int __all_sync(unsigned mask, int pred) { return __ballot_sync(mask, pred) == 0xFFFFFFFFu; }
template <typename T> inline T __shfl_sync(unsigned mask, T val, int src) {
    static_assert(sizeof(T) <= 8, "shfl");
    if (mask != 0xFFFFFFFFu) abort();
    unsigned long long v[32], mine = 0; memcpy(&mine, &val, sizeof(T)); simt::exchange(mine, v);
    T out; memcpy(&out, &v[src & 31], sizeof(T)); return out;
}
template <typename T> inline T __shfl_up_sync(unsigned mask, T val, unsigned delta) {
    if (mask != 0xFFFFFFFFu) abort();
    unsigned long long v[32], mine = 0; memcpy(&mine, &val, sizeof(T)); simt::exchange(mine, v);
    const int lane = simt::ctx().lane, src = lane - (int)delta;
    T out; memcpy(&out, &v[src < 0 ? lane : src], sizeof(T)); return out;
}
inline unsigned __reduce_add_sync(unsigned mask, unsigned val) {
    unsigned long long v[32]; simt::exchange_masked(mask, val, v);
    unsigned r = 0; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r += (unsigned)v[l];
    return r;
}
inline unsigned __reduce_max_sync(unsigned mask, unsigned val) {
    unsigned long long v[32]; simt::exchange_masked(mask, val, v);
    unsigned r = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && (unsigned)v[l] > r) r = (unsigned)v[l];
    return r;
}
inline unsigned __reduce_or_sync(unsigned mask, unsigned val) {
    unsigned long long v[32]; simt::exchange_masked(mask, val, v);
    unsigned r = 0; for (int l = 0; l < 32; l++) if ((mask >> l) & 1u) r |= (unsigned)v[l];
    return r;
}
inline unsigned __match_any_sync(unsigned mask, unsigned val) {
    unsigned long long v[32]; simt::exchange_masked(mask, val, v);
    unsigned r = 0; for (int l = 0; l < 32; l++) if (((mask >> l) & 1u) && (unsigned)v[l] == val) r |= 1u << l;
    return r;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(int x) { return __builtin_ffs(x); }
inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
// n-th set bit of mask at or above `base` (offset > 0), as the CUDA intrinsic; 0xFFFFFFFF when there is none
inline unsigned __fns(unsigned mask, unsigned base, int offset) {
    if (offset <= 0) abort();
    for (unsigned b = base; b < 32; b++)
        if ((mask >> b) & 1u) { if (--offset == 0) return b; }
    return 0xFFFFFFFFu;
}

// ---- atomics (shared or global: both are plain host memory here) -------------------------------------------
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd_system(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAnd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicAnd(unsigned* p, unsigned v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <typename T> inline T simt_atomic_max(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur < v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
template <typename T> inline T simt_atomic_min(T* p, T v) {
    T cur = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (cur > v && !__atomic_compare_exchange_n(p, &cur, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return cur;
}
inline unsigned atomicMax(unsigned* p, unsigned v) { return simt_atomic_max(p, v); }
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) { return simt_atomic_max(p, v); }
inline unsigned atomicMin(unsigned* p, unsigned v) { return simt_atomic_min(p, v); }
inline unsigned atomicCAS(unsigned* p, unsigned cmp, unsigned v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long cmp, unsigned long long v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }

template <typename T> inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }

inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }

// ---- the PTX helpers of csrc/common.cuh (that header defines them under __CUDACC__ only) -------------------
namespace fa {
inline void named_barrier_sync(int id, int count) { simt::named_barrier(id, (unsigned)count); }
// mbarrier: bit 0 of the word = parity of the phase in progress; a completed bulk copy flips it
inline void mbar_init(unsigned long long* bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_expect_tx(unsigned long long*, uint32_t) {}
inline void mbar_wait(unsigned long long* bar, uint32_t parity) {
    while ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) & 1ull) == parity) sched_yield();
}
inline void fence_barrier_init() {}
inline void fence_proxy_async() {}
inline void tma_load_1d(void* smem_dst, const void* gmem_src, uint32_t bytes, unsigned long long* bar) {
    memcpy(smem_dst, gmem_src, bytes);
    __atomic_fetch_xor(bar, 1ull, __ATOMIC_SEQ_CST);
}
inline void tma_load_1d_stream(void* d, const void* s, uint32_t bytes, unsigned long long* bar) { tma_load_1d(d, s, bytes, bar); }
inline void tma_prefetch_l2(const void*, uint32_t) {}
inline void prefetch_l2(const void*) {}
inline uint4 ld_cg_u4(const uint4* p) {                       // two 8-byte halves: a 16-byte line chunk may tear here
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(p);
    const unsigned long long a = __atomic_load_n(q, __ATOMIC_SEQ_CST), b = __atomic_load_n(q + 1, __ATOMIC_SEQ_CST);
    return make_uint4((uint32_t)a, (uint32_t)(a >> 32), (uint32_t)b, (uint32_t)(b >> 32));
}
inline unsigned long long ld_cg_u64(const unsigned long long* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
inline void st_cg_u4(uint4* p, uint4 v) {
    unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
    __atomic_store_n(q, (unsigned long long)v.x | ((unsigned long long)v.y << 32), __ATOMIC_SEQ_CST);
    __atomic_store_n(q + 1, (unsigned long long)v.z | ((unsigned long long)v.w << 32), __ATOMIC_SEQ_CST);
}
inline uint4 ld_stream_u4(const uint4* p) { return *p; }
inline void red_add_u64(void* p, unsigned long long v) { __atomic_fetch_add(static_cast<unsigned long long*>(p), v, __ATOMIC_SEQ_CST); }
inline void red_max_u64(void* p, unsigned long long v) { simt_atomic_max(static_cast<unsigned long long*>(p), v); }
inline void red_min_u64(void* p, unsigned long long v) { simt_atomic_min(static_cast<unsigned long long*>(p), v); }
inline void red_add_u32(void* p, uint32_t v) { __atomic_fetch_add(static_cast<uint32_t*>(p), v, __ATOMIC_SEQ_CST); }
inline void red_or_u64(void* p, unsigned long long v) { __atomic_fetch_or(static_cast<unsigned long long*>(p), v, __ATOMIC_SEQ_CST); }
inline void red_or_u32(void* p, uint32_t v) { __atomic_fetch_or(static_cast<uint32_t*>(p), v, __ATOMIC_SEQ_CST); }
inline void red_max_u32(void* p, uint32_t v) { simt_atomic_max(static_cast<uint32_t*>(p), v); }
inline void red_min_u32(void* p, uint32_t v) { simt_atomic_min(static_cast<uint32_t*>(p), v); }
inline uint64_t u64_of(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }
}  // namespace fa

// engine_emul.cpp — the WHOLE C ABI (csrc/engine.cu) on the CPU, for tests only.
//
// engine.cu is compiled unchanged by g++; what it calls is replaced by test doubles:
//   * the CUDA runtime entry points it uses (cudaMalloc, cudaMemcpyAsync, streams, events ...) are defined below over
//     plain host memory, everything synchronous (one fake device with 4 "SMs" so that grids stay small);
//   * the launch_* functions of the kernel files are re-stated here on tests/emul/simt.h (one OS thread per CUDA
//     thread) around the product's own kernel source.
// Purpose: check HOST logic (chunking, staging, the "full" cut, KERNEL_MAP plumbing, error paths) and kernel logic
// together, through the same ctypes binding the GPU tests use, when no GPU is at hand.  It is NOT a CPU backend of the
// product: it is built into tests/emul/_build/ by the test-suite, libflowagg.so contains none of it, and the
// package's loader cannot reach it (tests swap the library handle explicitly).  Speed is irrelevant here.
#define FA_HOST_EMUL 1
#include "simt.h"

#include <algorithm>
#include <map>
#include <mutex>

#undef __shared__
#define __shared__ static          // static shared arrays: their kernels are launched one CTA at a time (launch_serial)

#include "../../netobserv_ebpf_agent_b200/csrc/aggregate.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/evict.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/features.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/kmap.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/misc_kernels.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/pbflow.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/snaps.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/dnscorr.cu"

// ------------------------------------------------------------------------------------------------ CUDA runtime double
namespace {
std::mutex g_alloc_mu;
std::map<uintptr_t, std::pair<size_t, int>> g_allocs;     // base -> (size, 1 = device / 2 = pinned host)
int kind_of(const void* p) {
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    auto it = g_allocs.upper_bound(reinterpret_cast<uintptr_t>(p));
    if (it == g_allocs.begin()) return 0;
    --it;
    return reinterpret_cast<uintptr_t>(p) < it->first + it->second.first ? it->second.second : 0;
}
cudaError_t do_alloc(void** out, size_t size, int kind) {
    void* p = nullptr;
    if (posix_memalign(&p, 1024, size ? size : 1024)) return cudaErrorMemoryAllocation;
    memset(p, 0xCD, size ? size : 1024);                 // device memory is not zeroed by cudaMalloc
    std::lock_guard<std::mutex> lk(g_alloc_mu);
    g_allocs[reinterpret_cast<uintptr_t>(p)] = {size ? size : 1024, kind};
    *out = p;
    return cudaSuccess;
}
cudaError_t do_free(void* p) {
    if (!p) return cudaSuccess;
    {
        std::lock_guard<std::mutex> lk(g_alloc_mu);
        if (!g_allocs.erase(reinterpret_cast<uintptr_t>(p))) return cudaErrorInvalidValue;
    }
    free(p);
    return cudaSuccess;
}
}  // namespace

extern "C" {
cudaError_t cudaGetDeviceCount(int* n) { *n = 1; return cudaSuccess; }
cudaError_t cudaSetDevice(int) { return cudaSuccess; }
cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp* prop, int) {
    memset(prop, 0, sizeof *prop);
    prop->major = 10; prop->minor = 0; prop->multiProcessorCount = 4;
    return cudaSuccess;
}
cudaError_t cudaGetLastError(void) { return cudaSuccess; }
const char* cudaGetErrorString(cudaError_t) { return "emulated CUDA runtime error"; }
cudaError_t cudaMalloc(void** p, size_t size) { return do_alloc(p, size, 1); }
cudaError_t cudaFree(void* p) { return do_free(p); }
cudaError_t cudaHostAlloc(void** p, size_t size, unsigned int) { return do_alloc(p, size, 2); }
cudaError_t cudaFreeHost(void* p) { return do_free(p); }
cudaError_t cudaMemset(void* p, int v, size_t n) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memmove(d, s, n); return cudaSuccess; }
cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned int) { *s = reinterpret_cast<cudaStream_t>(uintptr_t(0x5EED)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned int) { *e = reinterpret_cast<cudaEvent_t>(uintptr_t(0xE7E7)); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventQuery(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaPointerGetAttributes(cudaPointerAttributes* a, const void* p) {
    memset(a, 0, sizeof *a);
    const int k = kind_of(p);
    a->type = k == 1 ? cudaMemoryTypeDevice : k == 2 ? cudaMemoryTypeHost : cudaMemoryTypeUnregistered;
    return cudaSuccess;
}
cudaError_t cudaIpcGetMemHandle(cudaIpcMemHandle_t*, void*) { return cudaErrorNotSupported; }
cudaError_t cudaIpcOpenMemHandle(void**, cudaIpcMemHandle_t, unsigned int) { return cudaErrorNotSupported; }
cudaError_t cudaIpcCloseMemHandle(void*) { return cudaErrorNotSupported; }
}  // extern "C"

// ------------------------------------------------------------------------------------------------ launch_* on simt.h
namespace fa {

namespace {
// grid sizes follow the product's formulas with the fake device's 4 SMs, capped so that a launch stays a few
// hundred OS threads; CTAs of kernels with static shared arrays run one after the other
unsigned small(unsigned grid, unsigned cap = 3) { return std::max(1u, std::min(grid, cap)); }
void launch_serial(unsigned grid, unsigned block, const std::function<void()>& body, size_t smem_bytes = 0) {
    std::vector<uint8_t> smem(smem_bytes + 256);
    for (unsigned b = 0; b < grid; b++) {
        std::vector<std::thread> th;
        simt::Cta cta; cta.sync_all = std::make_unique<simt::Barrier>(block);
        cta.smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem.data()) + 127) & ~(uintptr_t)127);
        for (unsigned w = 0; w < block / 32; w++) cta.warps.push_back(std::make_unique<simt::Warp>());
        for (unsigned t = 0; t < block; t++)
            th.emplace_back([&, t] {
                simt::Ctx c; c.tid = {t, 0, 0}; c.bid = {b, 0, 0}; c.bdim = {block, 1, 1}; c.gdim = {grid, 1, 1};
                c.lane = (int)(t & 31); c.cta = &cta; c.warp = cta.warps[t >> 5].get();
                simt::g_ctx = &c; body(); simt::g_ctx = nullptr;
            });
        for (auto& x : th) x.join();
    }
}
}  // namespace

int launch_aggregate(const AggLaunch& a, cudaStream_t) {
    if (a.n == 0) return 0;
    const bool dev_n = (a.opt & 16u) != 0;
    const uint4* recs = a.recs; const uint32_t n = a.n; Table t = a.table; const uint64_t epoch = a.epoch;
    Counters* ctr = a.ctr; uint32_t* spill = a.spill_idx; SketchParams sk = a.sk; const uint32_t opt = a.opt;
    {
        const uint32_t n_tiles = (n + kTile - 1) / kTile;
        const unsigned g = small((n_tiles + kTeams - 1) / kTeams, 2);
        const size_t sm = sizeof(AggSmem);
        if (sk.cms && dev_n) simt::launch(g, kCtaThreads, sm, [=] { aggregate_kernel<true, false, true>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt); });
        else if (sk.cms) simt::launch(g, kCtaThreads, sm, [=] { aggregate_kernel<true, false, false>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt); });
        else if (dev_n) simt::launch(g, kCtaThreads, sm, [=] { aggregate_kernel<false, false, true>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt); });
        else simt::launch(g, kCtaThreads, sm, [=] { aggregate_kernel<false, false, false>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt); });
    }
    FixupScratch* sc = a.scratch; const uint32_t ss = a.scratch_slots;
    simt::launch(2, 256, 0, [=] { fixup_scan_kernel(recs, n, t, ctr, sc, ss - 1, opt); });
    simt::launch(2, 256, 0, [=] { fixup_apply_kernel(recs, t, epoch, ctr, sc, ss, reinterpret_cast<unsigned int*>(&ctr->scratch[1])); });
    return 3;
}

int launch_evict(const Table& table, uint4* out, uint32_t* slot_of_out, unsigned long long cap, Counters* ctr, int,
                 cudaStream_t, bool drain) {
    Table t = table;
    if (drain) simt::launch(2, 256, 0, [=] { evict_kernel<true>(t, out, slot_of_out, cap, ctr); });
    else simt::launch(2, 256, 0, [=] { evict_kernel<false>(t, out, slot_of_out, cap, ctr); });
    return 1;
}

int launch_full_cut(const uint4* recs, uint32_t n, const Table& table, unsigned long long live, unsigned long long max_entries,
                    uint32_t* idx_set, uint32_t set_slots, uint32_t* bitmap, uint32_t* cut_out, int, cudaStream_t) {
    Table t = table;
    memset(idx_set, 0xFF, (size_t)set_slots * 4);
    memset(bitmap, 0, ((size_t)n + 31) / 32 * 4);
    simt::launch(2, 256, 0, [=] { cut_scan_kernel(recs, n, t, idx_set, set_slots - 1); });
    simt::launch(2, 256, 0, [=] { cut_mark_kernel(idx_set, set_slots, bitmap); });
    const unsigned long long room = max_entries > live ? max_entries - live : 0ull;
    launch_serial(1, 1024, [=] { cut_select_kernel(bitmap, n, room, cut_out); });
    return 3;
}

int launch_feature_fold(int kind, const uint8_t* recs, uint32_t n, const Table& table, uint64_t epoch, uint64_t seq0,
                        uint32_t* slot_of, Counters* ctr, int, cudaStream_t) {
    if (!n) return 0;
    Table t = table;
    if (kind == 2) {
        simt::launch(2, kFeatTile, feature_fold_smem<DropFeat>(), [=] { feature_fold_kernel<DropFeat>(recs, n, t, epoch, seq0, slot_of, ctr); });
        simt::launch(2, 256, 0, [=] { pktdrop_first_kernel(recs, n, t, seq0, slot_of); });
    } else if (kind == 0) {
        simt::launch(2, kFeatTile, feature_fold_smem<AddFeat>(), [=] { feature_fold_kernel<AddFeat>(recs, n, t, epoch, seq0, slot_of, ctr); });
        simt::launch(2, 256, 0, [=] { additional_first_kernel(recs, n, t, seq0, slot_of); });
    } else {
        simt::launch(2, kFeatTile, feature_fold_smem<DnsFeat>(), [=] { feature_fold_kernel<DnsFeat>(recs, n, t, epoch, seq0, slot_of, ctr); });
        simt::launch(2, 256, 0, [=] { dns_first_kernel(recs, n, t, seq0, slot_of); });
    }
    return 2;
}
int launch_evict_features(const Table& table, const uint32_t* slot_of_out, unsigned long long n_out, uint8_t* out_recs,
                          uint8_t* out_dns, uint8_t* out_add, uint8_t* out_drop, unsigned long long* out_rtt_min,
                          uint8_t* out_present, int, cudaStream_t) {
    if (!n_out) return 0;
    Table t = table;
    simt::launch(2, 256, 0, [=] { evict_features_kernel(t, slot_of_out, n_out, out_recs, out_dns, out_add, out_drop, out_rtt_min, out_present); });
    return 1;
}

int launch_dns_correlate(const uint8_t* pkts, uint32_t n, const DnsCorr& dc, uint32_t* state, uint8_t* samples, uint32_t* warp_count,
                         uint8_t* out, int, cudaStream_t) {
    if (!n) return 0;
    const DnsCorr d = dc;
    // FA_EMUL_DNS_ROUNDS=0: no round kernels, so that the single-thread tail is exercised as well
    const char* rs = getenv("FA_EMUL_DNS_ROUNDS");
    const uint32_t rounds = rs ? (uint32_t)atoi(rs) : (uint32_t)kDnsRounds;
    simt::launch(2, 256, 0, [=] { dns_resolve_kernel(pkts, n, d, state, samples); });
    for (uint32_t r = 0; r < rounds; r++) simt::launch(2, 256, 0, [=] { dns_round_kernel(pkts, n, d, state, samples, r & 1u); });
    simt::launch(1, 32, 0, [=] { dns_tail_kernel(pkts, n, d, state, samples); });
    simt::launch(2, 256, 0, [=] { dns_count_kernel(state, n, warp_count); });
    simt::launch(1, 32, 0, [=] { dns_scan_kernel(warp_count, (n + 31) / 32, d.ctr); });
    simt::launch(2, 256, 0, [=] { dns_scatter_kernel(state, n, warp_count, samples, out); });
    return 5 + (int)rounds;
}
int launch_dns_purge(const DnsCorr& dc, uint64_t now, uint64_t timeout, int, cudaStream_t) {
    const DnsCorr d = dc;
    simt::launch(2, 256, 0, [=] { dns_purge_kernel(d, now, timeout); });
    return 1;
}
int launch_dns_rebuild(const DnsCorr& from, const DnsCorr& to, int, cudaStream_t) {
    const DnsCorr f = from, t = to;
    simt::launch(2, 256, 0, [=] { dns_rebuild_kernel(f, t); });
    return 1;
}
int launch_expand_events(const uint4* events, uint32_t n, uint4* recs_out, cudaStream_t) {
    if (!n) return 0;
    simt::launch(2, 256, 0, [=] { expand_events_kernel(events, n, recs_out); });
    return 1;
}
int launch_parse_snaps(const uint8_t* snaps, uint32_t n, uint32_t stride, const FilterSet* filter, uint32_t* cta_count, uint8_t* verdict,
                       uint4* out_recs, uint32_t* src_of, unsigned long long* n_out, unsigned long long* filter_ctr, int, cudaStream_t) {
    if (!n) return 0;
    const unsigned grid = std::min<unsigned>((n + kSnapTile - 1) / kSnapTile, 3);
    const size_t smem = (size_t)kSnapTile * std::max<uint32_t>(stride, kRecBytes);
    FilterSet F = filter ? *filter : FilterSet{};
    if (F.n_rules) {
        launch_serial(grid, kSnapTile, [=] { snap_count_kernel<true>(snaps, n, stride, F, cta_count, verdict, filter_ctr); }, smem);
        launch_serial(grid, kSnapTile, [=] { snap_parse_kernel<true>(snaps, n, stride, F, cta_count, verdict, out_recs, src_of, n_out); }, smem);
    } else {
        launch_serial(grid, kSnapTile, [=] { snap_count_kernel<false>(snaps, n, stride, F, cta_count, verdict, filter_ctr); }, smem);
        launch_serial(grid, kSnapTile, [=] { snap_parse_kernel<false>(snaps, n, stride, F, cta_count, verdict, out_recs, src_of, n_out); }, smem);
    }
    return 2;
}
int launch_pb_sizes(const PbInputs& in_, uint32_t n, const PbParams& P_, uint32_t* sizes, unsigned long long* offsets,
                    unsigned long long* block_sums, int, cudaStream_t) {
    if (!n) return 0;
    PbInputs in = in_; PbParams P = P_;
    const uint32_t nb = (n + kScanBlock - 1) / kScanBlock;
    simt::launch(2, 256, 0, [=] { pb_size_kernel(in, n, P, sizes); });
    launch_serial(nb, kScanBlock, [=] { pb_scan_block_kernel(sizes, n, P.wrap, offsets, block_sums); });
    launch_serial(1, kScanBlock, [=] { pb_scan_sums_kernel(block_sums, nb); });
    launch_serial(nb, kScanBlock, [=] { pb_scan_add_kernel(offsets, n, block_sums, sizes, P.wrap); });
    return 4;
}
int launch_pb_write(const PbInputs& in_, uint32_t n, const PbParams& P_, const unsigned long long* offsets, const uint32_t* sizes,
                    uint8_t* out, uint8_t* keys_out, cudaStream_t) {
    if (!n) return 0;
    PbInputs in = in_; PbParams P = P_;
    launch_serial((n + kPbCta - 1) / kPbCta, kPbCta, [=] { pb_write_kernel(in, n, P, offsets, sizes, out, keys_out); }, kPbStage);
    return 1;
}

int launch_cms_query(const SketchParams& sk_, const uint4* keys, uint32_t n, unsigned long long* est, cudaStream_t) {
    if (!n) return 0;
    SketchParams sk = sk_;
    simt::launch(small((n + 255) / 256), 256, 0, [=] { cms_query_kernel(sk, keys, n, est); });
    return 1;
}
int launch_hll_pack(const SketchParams& sk_, uint8_t* out, cudaStream_t) {
    SketchParams sk = sk_;
    simt::launch(2, 256, 0, [=] { hll_pack_kernel(sk, out); });
    return 1;
}
int launch_generate(const GenDeviceParams& g_, uint64_t first_index, uint32_t n, uint4* dst, cudaStream_t) {
    if (!n) return 0;
    GenDeviceParams g = g_;
    simt::launch(2, 256, 0, [=] { generate_kernel(g, first_index, n, dst); });
    return 1;
}
int launch_route_peer(const uint4* recs, const unsigned long long* n_dev, uint32_t max_n, uint32_t n_shards, uint32_t self_shard,
                      const PeerTargets& pt_, unsigned long long cap, unsigned long long* overflow, int, cudaStream_t) {
    if (!max_n) return 0;
    PeerTargets pt = pt_;                                                        // "peer" buffers are plain host memory here
    launch_serial(3, kRouteTile, [=] { route_peer_kernel(recs, n_dev, max_n, n_shards, self_shard, pt, cap, overflow); });
    return 1;
}
int launch_route(const uint4* recs, uint32_t n, uint32_t n_shards, uint4* out, unsigned long long* counts_dev, uint32_t* tmp,
                 int, cudaStream_t) {
    if (!n) { memset(counts_dev, 0, n_shards * sizeof(unsigned long long)); return 0; }
    const uint32_t n_ctas = (n + kRoutePerCta - 1) / kRoutePerCta;
    uint32_t* owner = tmp; uint32_t* hist = tmp + n;
    launch_serial(n_ctas, kRouteThreads, [=] { route_count_kernel(recs, n, n_shards, owner, hist); });
    launch_serial(1, 1024, [=] { route_scan_kernel(hist, n_ctas, n_shards, counts_dev); });
    launch_serial(n_ctas, kRouteThreads, [=] { route_scatter_kernel(recs, n, n_shards, owner, hist, out); });
    return 3;
}

int launch_kmap_batch(KmParams P, uint32_t cut, int, cudaStream_t) {
    if (!P.n) return 0;
    int launches = 0;
    if (cut > 0) { KmParams Q = P; Q.lo = 0; Q.hi = cut; Q.allow_insert = 1; simt::launch(3, 256, 0, [=] { km_resolve_kernel(Q); }); launches++; }
    if (cut < P.n) { KmParams Q = P; Q.lo = cut; Q.hi = P.n; Q.allow_insert = 0; simt::launch(3, 256, 0, [=] { km_resolve_kernel(Q); }); launches++; }
    simt::launch(3, 256, 0, [=] { km_init_kernel(P); });
    simt::launch(3, 256, 0, [=] { km_fold_kernel(P); });
    simt::launch(3, 256, 0, [=] { km_bresolve_kernel(P); });
    simt::launch(3, 256, 0, [=] { km_order_kernel(P); });
    simt::launch(3, 256, 0, [=] { km_cleanup_kernel(P); });
    KmCounters* c = P.c;
    simt::launch(1, 32, 0, [=] { if (threadIdx.x == 0) km_reset_bset_count_kernel(c); });
    return launches + 6;
}
int launch_kmap_batch_v2(KmParams P, uint32_t cut, int, cudaStream_t) {
    if (!P.n) return 0;
    int launches = 0;
    if (cut > 0) { KmParams Q = P; Q.lo = 0; Q.hi = cut; Q.allow_insert = 1; simt::launch(3, 256, 0, [=] { km2_resolve_fold_kernel(Q); }); launches++; }
    if (cut < P.n) { KmParams Q = P; Q.lo = cut; Q.hi = P.n; Q.allow_insert = 0; simt::launch(3, 256, 0, [=] { km2_resolve_fold_kernel(Q); }); launches++; }
    simt::launch(3, 256, 0, [=] { km2_init_kernel(P); });
    simt::launch(3, 256, 0, [=] { km2_fold_deferred_kernel(P); });
    simt::launch(3, 256, 0, [=] { km_bresolve_kernel(P); });
    simt::launch(3, 256, 0, [=] { km2_order_b_kernel(P); });
    simt::launch(3, 256, 0, [=] { km2_finish_kernel(P); });
    simt::launch(3, 256, 0, [=] { km2_cleanup_kernel(P); });
    KmCounters* c = P.c;
    simt::launch(1, 32, 0, [=] { if (threadIdx.x == 0) km2_reset_counts_kernel(c); });
    return launches + 7;
}
int launch_kmap_evict(const Table& table, uint8_t* met, uint8_t* out, unsigned long long cap, unsigned long long* cursor,
                      uint32_t* slot_of_out, int, cudaStream_t) {
    Table t = table;
    simt::launch(3, 256, 0, [=] { km_evict_kernel(t, met, out, cap, cursor, slot_of_out); });
    return 1;
}

}  // namespace fa

#undef __shared__
#include "../../netobserv_ebpf_agent_b200/csrc/engine.cu"

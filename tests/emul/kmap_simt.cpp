// kmap_simt.cpp — the KERNEL_MAP kernels of csrc/kmap.cu run on the SIMT emulation (simt.h): one OS thread per CUDA
// thread, real atomics, several CTAs at once.  Same C interface as kmap_emul.cpp (which runs the bodies one index at
// a time in shuffled order); what this adds is the claim / publish protocol of `resolve` and the shared scratch
// updates under real preemptive concurrency.  Test infrastructure only.
#define FA_HOST_EMUL 1
#include "simt.h"

#include <algorithm>
#include <string>
#include <unordered_set>

#include "../../netobserv_ebpf_agent_b200/csrc/kmap.cu"

using namespace fa;

namespace {
struct Emul {
    uint64_t max_entries, slots; int ringbuf;
    uint8_t *ident, *met, *spill; uint32_t *occ, *slot_of, *blist, *touched, *deferred, *brec; KmBEntry* bset; uint32_t bset_slots;
    int impl = 1;
    KmCounters* c; unsigned long long* live; uint64_t epoch = 0, spill_cap; size_t max_batch;
    std::unordered_set<std::string> keys;
};
void* zalloc(size_t bytes) { void* p = nullptr; if (posix_memalign(&p, 1024, bytes ? bytes : 1024)) abort(); memset(p, 0, bytes ? bytes : 1024); return p; }
KmParams params(Emul* e) {
    KmParams P{};
    P.ringbuf = e->ringbuf;
    P.t.ident = reinterpret_cast<uint4*>(e->ident); P.t.occ = e->occ; P.t.mask = e->slots - 1;
    P.met = e->met; P.slot_of = e->slot_of; P.live = e->live; P.c = e->c;
    P.spill = e->spill; P.spill_cap = e->spill_cap; P.bset = e->bset; P.bset_mask = e->bset_slots - 1; P.blist = e->blist;
    P.touched = e->touched; P.deferred = e->deferred; P.brec = e->brec;
    return P;
}
constexpr unsigned kGrid = 3, kBlock = 256;
}  // namespace

extern "C" {

void* kmap_emul_new(uint64_t max_entries, uint64_t max_batch, int ringbuf, uint64_t spill_cap, uint64_t) {
    Emul* e = new Emul();
    e->max_entries = max_entries; e->ringbuf = ringbuf; e->max_batch = max_batch; e->spill_cap = spill_cap;
    uint64_t want = max_entries + max_entries / 3 + 1, slots = 1024;
    while (slots < want) slots <<= 1;
    e->slots = slots;
    e->ident = static_cast<uint8_t*>(zalloc(slots * kIdentBytes)); e->met = static_cast<uint8_t*>(zalloc(slots * kMetLineBytes));
    e->occ = static_cast<uint32_t*>(zalloc(slots / 8));
    e->slot_of = static_cast<uint32_t*>(zalloc(max_batch * 4)); e->blist = static_cast<uint32_t*>(zalloc(max_batch * 4));
    e->touched = static_cast<uint32_t*>(zalloc(max_batch * 4)); e->deferred = static_cast<uint32_t*>(zalloc(max_batch * 4));
    e->brec = static_cast<uint32_t*>(zalloc(max_batch * 4));
    uint32_t bs = 1024; while ((uint64_t)bs < 2 * max_batch) bs <<= 1;
    e->bset_slots = bs; e->bset = static_cast<KmBEntry*>(zalloc((size_t)bs * sizeof(KmBEntry)));
    e->spill = static_cast<uint8_t*>(zalloc(spill_cap * kRecBytes));
    e->c = static_cast<KmCounters*>(zalloc(sizeof(KmCounters)));
    e->live = static_cast<unsigned long long*>(zalloc(8));
    return e;
}
void kmap_emul_free(void* h) {
    Emul* e = static_cast<Emul*>(h);
    free(e->ident); free(e->met); free(e->occ); free(e->slot_of); free(e->blist); free(e->bset); free(e->spill); free(e->c); free(e->live);
    free(e->touched); free(e->deferred); free(e->brec);
    delete e;
}
void kmap_emul_set_impl(void* h, int impl) { static_cast<Emul*>(h)->impl = impl; }
int kmap_emul_batch(void* h, const uint8_t* recs, uint32_t n) {
    Emul* e = static_cast<Emul*>(h);
    if (n > e->max_batch) return -1;
    uint32_t cut = n;                                   // what launch_full_cut delivers on the device
    {
        uint64_t live = *e->live;
        std::unordered_set<std::string> fresh;
        for (uint32_t i = 0; i < n; i++) {
            std::string k(reinterpret_cast<const char*>(recs + (size_t)i * kRecBytes), 39);
            if (e->keys.count(k) || fresh.count(k)) continue;
            if (live >= e->max_entries) { cut = i; break; }
            fresh.insert(k); live++;
        }
        for (auto& k : fresh) e->keys.insert(k);
    }
    KmParams P = params(e);
    P.recs = recs; P.n = n; P.epoch = ++e->epoch;
    if (e->impl == 2) {
        if (cut > 0) { KmParams Q = P; Q.lo = 0; Q.hi = cut; Q.allow_insert = 1; simt::launch(kGrid, kBlock, 0, [=] { km2_resolve_fold_kernel(Q); }); }
        if (cut < n) { KmParams Q = P; Q.lo = cut; Q.hi = n; Q.allow_insert = 0; simt::launch(kGrid, kBlock, 0, [=] { km2_resolve_fold_kernel(Q); }); }
        simt::launch(kGrid, kBlock, 0, [=] { km2_init_kernel(P); });
        simt::launch(kGrid, kBlock, 0, [=] { km2_fold_deferred_kernel(P); });
        simt::launch(kGrid, kBlock, 0, [=] { km_bresolve_kernel(P); });
        simt::launch(kGrid, kBlock, 0, [=] { km2_order_b_kernel(P); });
        simt::launch(kGrid, kBlock, 0, [=] { km2_finish_kernel(P); });
        simt::launch(kGrid, kBlock, 0, [=] { km2_cleanup_kernel(P); });
        KmCounters* c2 = e->c;
        simt::launch(1, 32, 0, [=] { if (threadIdx.x == 0) km2_reset_counts_kernel(c2); });
        for (uint32_t i = 0; i < e->bset_slots; i++) if (e->bset[i].key || e->bset[i].nfirst || e->bset[i].next || e->bset[i].kind) return -2;
        for (uint64_t sl = 0; sl < e->slots; sl++)
            for (int b = 48; b < 96; b++) if (e->ident[sl * kIdentBytes + b]) return -3;
        return 0;
    }
    if (cut > 0) { KmParams Q = P; Q.lo = 0; Q.hi = cut; Q.allow_insert = 1; simt::launch(kGrid, kBlock, 0, [=] { km_resolve_kernel(Q); }); }
    if (cut < n) { KmParams Q = P; Q.lo = cut; Q.hi = n; Q.allow_insert = 0; simt::launch(kGrid, kBlock, 0, [=] { km_resolve_kernel(Q); }); }
    simt::launch(kGrid, kBlock, 0, [=] { km_init_kernel(P); });
    simt::launch(kGrid, kBlock, 0, [=] { km_fold_kernel(P); });
    simt::launch(kGrid, kBlock, 0, [=] { km_bresolve_kernel(P); });
    simt::launch(kGrid, kBlock, 0, [=] { km_order_kernel(P); });
    simt::launch(kGrid, kBlock, 0, [=] { km_cleanup_kernel(P); });
    KmCounters* c = e->c;
    simt::launch(1, 32, 0, [=] { if (threadIdx.x == 0) km_reset_bset_count_kernel(c); });
    for (uint32_t i = 0; i < e->bset_slots; i++) if (e->bset[i].key || e->bset[i].nfirst || e->bset[i].next || e->bset[i].kind) return -2;
    return 0;
}
uint64_t kmap_emul_live(void* h) { return *static_cast<Emul*>(h)->live; }
uint64_t kmap_emul_evict(void* h, uint8_t* out, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    KmParams P = params(e);
    unsigned long long cursor = 0; unsigned long long* cp = &cursor;
    Table t = P.t; uint8_t* met = e->met;
    simt::launch(kGrid, kBlock, 0, [=] { km_evict_kernel(t, met, out, cap, cp, nullptr); });
    *e->live = 0; e->keys.clear();
    for (size_t i = 0; i < e->slots * kIdentBytes; i++) if (e->ident[i]) return ~0ull;
    for (size_t i = 0; i < e->slots * kMetLineBytes; i++) if (e->met[i]) return ~0ull;
    return cursor;
}
uint64_t kmap_emul_spilled(void* h, uint8_t* out, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    const uint64_t n = std::min<uint64_t>(e->c->spill_cursor, e->spill_cap);
    if (out) memcpy(out, e->spill, (size_t)std::min(n, cap) * kRecBytes);
    e->c->spill_cursor = 0;
    return n;
}
void kmap_emul_counters(void* h, uint64_t out[6]) {
    KmCounters* c = static_cast<Emul*>(h)->c;
    out[0] = c->intf_missed; out[1] = c->fail_create; out[2] = c->spill_cursor; out[3] = c->spill_dropped;
    out[4] = c->bset_count; out[5] = c->table_full;
}

}  // extern "C"

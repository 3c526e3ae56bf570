// kmap_emul.cpp — host emulation of the KERNEL_MAP kernels (test infrastructure, never part of libflowagg.so).
// Runs the per-thread bodies of netobserv_ebpf_agent_b200/csrc/kmap_body.cuh pass by pass with the thread indices of
// every pass in a seeded random order: the passes only use commutative atomics, so any order must give the
// reference's sequential result (checked against oracle_kmap_* by tests/test_kmap_emulation.py).
// What this does NOT cover: the device memory model (claim/publish fences) and launch plumbing — those are the
// GPU tests' job (tests/test_gpu_kernel_map.py).
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <unordered_set>
#include <string>
#include <vector>

#include "../../netobserv_ebpf_agent_b200/csrc/kmap_body.cuh"

using namespace fa;

namespace {
struct Emul {
    uint64_t max_entries, slots;
    int ringbuf;
    std::vector<uint8_t> ident, met, spill;
    std::vector<uint32_t> occ, slot_of, blist, touched, deferred, brec;
    int impl = 1;
    std::vector<KmBEntry> bset;
    KmCounters c{};
    unsigned long long live = 0;
    uint64_t epoch = 0, rng;
    size_t max_batch;
    std::unordered_set<std::string> keys;   // what launch_full_cut computes on the device: which keys are in the map
};

uint64_t next_rand(uint64_t& s) { s += 0x9E3779B97F4A7C15ull; uint64_t z = s; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); }

void shuffled(std::vector<uint32_t>& v, uint32_t lo, uint32_t hi, uint64_t& rng) {
    v.resize(hi - lo);
    for (uint32_t i = lo; i < hi; i++) v[i - lo] = i;
    for (size_t i = v.size(); i > 1; i--) std::swap(v[i - 1], v[next_rand(rng) % i]);
}
}  // namespace

extern "C" {

void* kmap_emul_new(uint64_t max_entries, uint64_t max_batch, int ringbuf, uint64_t spill_cap, uint64_t seed) {
    Emul* e = new Emul();
    e->max_entries = max_entries; e->ringbuf = ringbuf; e->rng = seed; e->max_batch = max_batch;
    uint64_t want = max_entries + max_entries / 3 + 1, slots = 1024;
    while (slots < want) slots <<= 1;
    e->slots = slots;
    e->ident.assign(slots * kIdentBytes + 64, 0); e->met.assign(slots * kMetLineBytes + 64, 0);
    e->occ.assign(slots / 32, 0);
    e->slot_of.assign(max_batch, 0); e->blist.assign(max_batch, 0);
    e->touched.assign(max_batch, 0); e->deferred.assign(max_batch, 0); e->brec.assign(max_batch, 0);
    uint64_t bs = 1024; while (bs < 2 * max_batch) bs <<= 1;
    e->bset.assign(bs, KmBEntry{});
    e->spill.assign(spill_cap * kRecBytes + 64, 0);
    return e;
}
void kmap_emul_free(void* h) { delete static_cast<Emul*>(h); }
void kmap_emul_set_impl(void* h, int impl) { static_cast<Emul*>(h)->impl = impl; }

// one batch (n <= max_batch); recs must be 16-byte aligned
int kmap_emul_batch(void* h, const uint8_t* recs, uint32_t n) {
    Emul* e = static_cast<Emul*>(h);
    if (n > e->max_batch) return -1;
    // the cut the device pre-pass (launch_full_cut) delivers: first record whose key is new while the map is full
    uint32_t cut = n;
    {
        uint64_t live = e->live;
        std::unordered_set<std::string> fresh;
        for (uint32_t i = 0; i < n; i++) {
            std::string k(reinterpret_cast<const char*>(recs + (size_t)i * kRecBytes), 39);
            if (e->keys.count(k) || fresh.count(k)) continue;
            if (live >= e->max_entries) { cut = i; break; }
            fresh.insert(k); live++;
        }
        for (auto& k : fresh) e->keys.insert(k);
    }
    KmParams P{};
    P.recs = recs; P.n = n; P.ringbuf = e->ringbuf;
    P.t.ident = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(e->ident.data()) + 15) & ~uintptr_t(15));
    P.t.occ = e->occ.data(); P.t.mask = e->slots - 1;
    P.met = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(e->met.data()) + 15) & ~uintptr_t(15));
    P.epoch = ++e->epoch; P.slot_of = e->slot_of.data(); P.live = &e->live; P.c = &e->c;
    P.spill = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(e->spill.data()) + 15) & ~uintptr_t(15));
    P.spill_cap = (e->spill.size() - 64) / kRecBytes;
    P.bset = e->bset.data(); P.bset_mask = (uint32_t)e->bset.size() - 1; P.blist = e->blist.data();
    P.touched = e->touched.data(); P.deferred = e->deferred.data(); P.brec = e->brec.data();
    std::vector<uint32_t> order;
    if (e->impl == 2) {
        if (cut > 0) { P.lo = 0; P.hi = cut; P.allow_insert = 1; shuffled(order, 0, cut, e->rng); for (uint32_t i : order) km2_resolve_fold_body(P, i); }
        if (cut < n) { P.lo = cut; P.hi = n; P.allow_insert = 0; shuffled(order, cut, n, e->rng); for (uint32_t i : order) km2_resolve_fold_body(P, i); }
        const uint32_t nd = (uint32_t)e->c.deferred_count;
        shuffled(order, 0, nd, e->rng); for (uint32_t k : order) km2_init_body(P, k);
        shuffled(order, 0, nd, e->rng); for (uint32_t k : order) km2_fold_deferred_body(P, k);
        const uint32_t m2 = (uint32_t)e->c.bset_count;
        shuffled(order, 0, m2, e->rng); for (uint32_t j : order) km_bresolve_body(P, j);
        shuffled(order, 0, (uint32_t)e->c.brec_count, e->rng); for (uint32_t k : order) km2_order_b_body(P, k);
        shuffled(order, 0, (uint32_t)e->c.touched_count, e->rng); for (uint32_t k : order) km2_finish_flow_body(P, k);
        shuffled(order, 0, m2, e->rng); for (uint32_t j : order) km_cleanup_bset_body(P, j);
        e->c.bset_count = 0; e->c.touched_count = 0; e->c.deferred_count = 0; e->c.brec_count = 0;
        for (auto& b : e->bset) if (b.key || b.nfirst || b.next || b.kind || b.pos) return -2;
        // every flow's scratch must be zero again
        const uint8_t* id = reinterpret_cast<const uint8_t*>(P.t.ident);
        for (uint64_t sl = 0; sl < e->slots; sl++)
            for (int b = 48; b < 96; b++) if (id[sl * kIdentBytes + b]) return -3;
        return 0;
    }
    if (cut > 0) { P.lo = 0; P.hi = cut; P.allow_insert = 1; shuffled(order, 0, cut, e->rng); for (uint32_t i : order) km_resolve_body(P, i); }
    if (cut < n) { P.lo = cut; P.hi = n; P.allow_insert = 0; shuffled(order, cut, n, e->rng); for (uint32_t i : order) km_resolve_body(P, i); }
    shuffled(order, 0, n, e->rng); for (uint32_t i : order) km_init_body(P, i);
    shuffled(order, 0, n, e->rng); for (uint32_t i : order) km_fold_body(P, i);
    const uint32_t m = (uint32_t)e->c.bset_count;
    shuffled(order, 0, m, e->rng); for (uint32_t j : order) km_bresolve_body(P, j);
    shuffled(order, 0, n, e->rng); for (uint32_t i : order) km_order_body(P, i);
    shuffled(order, 0, n, e->rng); for (uint32_t i : order) km_cleanup_record_body(P, i);
    shuffled(order, 0, m, e->rng); for (uint32_t j : order) km_cleanup_bset_body(P, j);
    e->c.bset_count = 0;
    // scratch must be all-zero again
    for (auto& b : e->bset) if (b.key || b.nfirst || b.next || b.kind || b.pos) return -2;
    return 0;
}

uint64_t kmap_emul_live(void* h) { return static_cast<Emul*>(h)->live; }

// lookup-and-delete everything; returns the number of flows
uint64_t kmap_emul_evict(void* h, uint8_t* out, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    Table t{}; 
    t.ident = reinterpret_cast<uint4*>((reinterpret_cast<uintptr_t>(e->ident.data()) + 15) & ~uintptr_t(15));
    t.occ = e->occ.data(); t.mask = e->slots - 1;
    uint8_t* met = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(e->met.data()) + 15) & ~uintptr_t(15));
    unsigned long long cursor = 0;
    std::vector<uint32_t> order; shuffled(order, 0, (uint32_t)(e->slots / 32), e->rng);
    for (uint32_t w : order) km_evict_word_body(t, met, w, out, cap, &cursor);
    e->live = 0; e->keys.clear();
    // the table must be all-zero again
    for (size_t i = 0; i < e->slots * kIdentBytes; i++) if (reinterpret_cast<uint8_t*>(t.ident)[i]) return ~0ull;
    for (size_t i = 0; i < e->slots * kMetLineBytes; i++) if (met[i]) return ~0ull;
    return cursor;
}
uint64_t kmap_emul_spilled(void* h, uint8_t* out, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    uint8_t* sp = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(e->spill.data()) + 15) & ~uintptr_t(15));
    const uint64_t n = std::min<uint64_t>(e->c.spill_cursor, (e->spill.size() - 64) / kRecBytes);
    if (out) memcpy(out, sp, (size_t)std::min(n, cap) * kRecBytes);
    e->c.spill_cursor = 0;
    return n;
}
void kmap_emul_counters(void* h, uint64_t out[6]) {
    Emul* e = static_cast<Emul*>(h);
    out[0] = e->c.intf_missed; out[1] = e->c.fail_create; out[2] = e->c.spill_cursor; out[3] = e->c.spill_dropped;
    out[4] = e->c.bset_count; out[5] = e->c.table_full;
}

}  // extern "C"

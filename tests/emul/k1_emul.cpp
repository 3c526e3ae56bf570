// k1_emul.cpp — K1 (aggregate_kernel + the two re-fold kernels) and K2 (evict_kernel) of the product, compiled for
// the host on top of simt.h and run with one OS thread per CUDA thread (test infrastructure, never part of
// libflowagg.so).  tests/test_k1_emulation.py compares what comes out with the oracle's Accounter.  This lets a change
// to the kernels' logic (e.g. the experiment variants, template parameter kVar) be checked for bit-exactness on the
// CPU at small sizes before any GPU time is spent on it; speed, registers and the device memory model are not covered.
#define FA_HOST_EMUL 1
#include "simt.h"

#include "../../netobserv_ebpf_agent_b200/csrc/aggregate.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/evict.cu"
#include "../../netobserv_ebpf_agent_b200/csrc/features.cu"

using namespace fa;

namespace {
struct Emul {
    Table t{};
    uint64_t slots = 0, epoch = 0;
    Counters* ctr = nullptr;
    FixupScratch* scratch = nullptr; uint32_t scratch_slots = 0;
    uint32_t* spill_idx = nullptr;
    uint64_t max_batch = 0;
    SketchParams sk{};                       // cms == nullptr: sketches off
    uint32_t* slot_of = nullptr;             // feature folds: per-sample slot scratch
    uint64_t feat_seq[2] = {0, 0};
};
void* zalloc(size_t bytes) {
    void* p = nullptr;
    if (posix_memalign(&p, 1024, bytes ? bytes : 1024)) abort();
    memset(p, 0, bytes ? bytes : 1024);
    return p;
}
void run_k1(Emul* e, const uint4* recs, uint32_t n, unsigned grid, uint32_t opt) {
    const uint64_t epoch = e->epoch;
    Table t = e->t; Counters* ctr = e->ctr; uint32_t* spill = e->spill_idx;
    SketchParams sk = e->sk;
    if (sk.cms)
        simt::launch(grid, kCtaThreads, sizeof(AggSmem), [=] {
            aggregate_kernel<true, false, false>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt);
        });
    else
        simt::launch(grid, kCtaThreads, sizeof(AggSmem), [=] {
            aggregate_kernel<false, false, false>(recs, n, t, epoch, ctr, spill, sk, nullptr, opt);
        });
}
}  // namespace

extern "C" {

void* k1_emul_new(uint64_t max_entries, uint64_t max_batch) {
    Emul* e = new Emul();
    uint64_t want = max_entries + max_entries / 3 + 1, slots = 1024;
    while (slots < want) slots <<= 1;
    e->slots = slots; e->max_batch = max_batch;
    e->t.mask = slots - 1;
    e->t.ident = static_cast<uint4*>(zalloc(slots * kIdentBytes));
    e->t.hot = static_cast<uint4*>(zalloc(slots * kHotBytes));
    e->t.occ = static_cast<uint32_t*>(zalloc(slots / 8));
    e->ctr = static_cast<Counters*>(zalloc(sizeof(Counters)));
    uint32_t ss = 1024; while ((uint64_t)ss < 2 * max_batch) ss <<= 1;
    e->scratch_slots = ss;
    e->scratch = static_cast<FixupScratch*>(zalloc((size_t)ss * sizeof(FixupScratch)));
    e->spill_idx = static_cast<uint32_t*>(zalloc(max_batch * 4));
    return e;
}
void k1_emul_free(void* h) {
    Emul* e = static_cast<Emul*>(h);
    free(e->t.ident); free(e->t.hot); free(e->t.occ); free(e->ctr); free(e->scratch); free(e->spill_idx);
    delete e;
}

// one launch_aggregate: K1 (variant kVar = var) on `grid` CTAs, then the ordered re-fold.  recs: 16-byte aligned.
int k1_emul_ingest(void* h, const uint8_t* recs8, uint32_t n, unsigned grid, int var, uint32_t opt) {
    Emul* e = static_cast<Emul*>(h);
    if (n == 0 || n > e->max_batch || (reinterpret_cast<uintptr_t>(recs8) & 15)) return -1;
    const uint4* recs = reinterpret_cast<const uint4*>(recs8);
    e->epoch++;
    const uint32_t n_tiles = (n + kTile - 1) / kTile;
    grid = std::min<unsigned>(grid, (n_tiles + kTeams - 1) / kTeams);
    switch (var) {
        case 0: run_k1(e, recs, n, grid, opt); break;
        default: return -2;
    }
    Table t = e->t; Counters* ctr = e->ctr; FixupScratch* sc = e->scratch; const uint32_t ss = e->scratch_slots;
    const uint64_t epoch = e->epoch;
    simt::launch(2, 256, 0, [=] { fixup_scan_kernel(recs, n, t, ctr, sc, ss - 1, opt); });
    simt::launch(2, 256, 0, [=] {
        fixup_apply_kernel(recs, t, epoch, ctr, sc, ss, reinterpret_cast<unsigned int*>(&ctr->scratch[1]));
    });
    for (uint32_t i = 0; i < ss; i++) if (sc[i].key) return -3;       // scratch must be clean between launches
    return 0;
}

// K6 feature folds (kind 0 = additional 72-byte samples, 1 = dns 104-byte samples) and the merged eviction
void k1_emul_enable_features(void* h) {
    Emul* e = static_cast<Emul*>(h);
    e->t.feat_add = static_cast<uint4*>(zalloc(e->slots * 80));
    e->t.feat_dns = static_cast<uint4*>(zalloc(e->slots * 128));
    e->slot_of = static_cast<uint32_t*>(zalloc(e->max_batch * 4));
}
int k1_emul_ingest_feature(void* h, int kind, const uint8_t* recs, uint32_t n) {
    Emul* e = static_cast<Emul*>(h);
    if (!e->slot_of || n == 0 || n > e->max_batch) return -1;
    Table t = e->t; Counters* ctr = e->ctr; uint32_t* so = e->slot_of;
    const uint64_t epoch = ++e->epoch, seq0 = e->feat_seq[kind];
    if (kind == 0) {
        simt::launch(2, kFeatTile, feature_fold_smem<AddFeat>(), [=] { feature_fold_kernel<AddFeat>(recs, n, t, epoch, seq0, so, ctr); });
        simt::launch(2, 256, 0, [=] { additional_first_kernel(recs, n, t, seq0, so); });
    } else {
        simt::launch(2, kFeatTile, feature_fold_smem<DnsFeat>(), [=] { feature_fold_kernel<DnsFeat>(recs, n, t, epoch, seq0, so, ctr); });
        simt::launch(2, 256, 0, [=] { dns_first_kernel(recs, n, t, seq0, so); });
    }
    e->feat_seq[kind] += n;
    return 0;
}
uint64_t k1_emul_evict_features(void* h, uint8_t* out, uint8_t* out_dns, uint8_t* out_add, uint8_t* out_present, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    Table t = e->t; Counters* ctr = e->ctr;
    ctr->evict_out = 0;
    uint32_t* slot_of_out = static_cast<uint32_t*>(zalloc((cap ? cap : 1) * 4));
    uint4* o = reinterpret_cast<uint4*>(out);
    simt::launch(2, 256, 0, [=] { evict_kernel<false>(t, o, slot_of_out, cap, ctr); });
    const uint64_t n = ctr->evict_out;
    simt::launch(2, 256, 0, [=] { evict_features_kernel(t, slot_of_out, n, out, out_dns, out_add, nullptr, nullptr, out_present); });
    free(slot_of_out);
    ctr->live = 0;
    return n;
}

// fused count-min + HyperLogLog (variants 0 and 8 only); arrays are owned by the emulation
void k1_emul_enable_sketch(void* h, uint32_t log2w, uint32_t depth, uint32_t p, uint64_t seed) {
    Emul* e = static_cast<Emul*>(h);
    e->sk.log2w = log2w; e->sk.depth = depth; e->sk.p = p; e->sk.seed = seed;
    e->sk.cms = static_cast<unsigned long long*>(zalloc(((size_t)depth << log2w) * 8));
    e->sk.hll = static_cast<uint32_t*>(zalloc(((size_t)1 << p) * 4));
}
void k1_emul_sketch_export(void* h, uint64_t* cms_out, uint8_t* hll_out) {
    Emul* e = static_cast<Emul*>(h);
    memcpy(cms_out, e->sk.cms, ((size_t)e->sk.depth << e->sk.log2w) * 8);
    for (size_t i = 0; i < ((size_t)1 << e->sk.p); i++) hll_out[i] = (uint8_t)e->sk.hll[i];
}
// path counters since the last call: 0 representatives probed, 1 of them through the general loop, 2 cache hits
void k1_emul_paths(uint64_t out[3]) {
    for (int i = 0; i < 3; i++) { out[i] = simt::g_counts[i]; simt::g_counts[i] = 0; }
}
uint64_t k1_emul_live(void* h) { return static_cast<Emul*>(h)->ctr->live; }
uint64_t k1_emul_counter(void* h, int which) {
    Counters* c = static_cast<Emul*>(h)->ctr;
    return which == 0 ? c->live : which == 1 ? c->spills : which == 2 ? c->fixups_total : c->dirty;
}

// K2: lookup-and-delete everything; returns the number of flows found
uint64_t k1_emul_evict(void* h, uint8_t* out, uint64_t cap) {
    Emul* e = static_cast<Emul*>(h);
    Table t = e->t; Counters* ctr = e->ctr;
    ctr->evict_out = 0;
    uint4* o = reinterpret_cast<uint4*>(out);
    simt::launch(2, 256, 0, [=] { evict_kernel<false>(t, o, nullptr, cap, ctr); });
    const uint64_t n = ctr->evict_out;
    ctr->live = 0;
    return n;
}

}  // extern "C"

/*
 * oracle.c — CPU restatement of the reference's flow-aggregation algorithms.
 * TEST INFRASTRUCTURE ONLY (see oracle.h for the rules and the parity status).
 * Every function cites the reference file:line it follows; nothing here is
 * copied — the reference is Go / eBPF C, this is a byte-offset restatement.
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------ field access */
/* offsets inside flow_metrics (reference bpf/types.h:94-126) */
enum {
    M_START = 0, M_END = 8, M_BYTES = 16, M_PACKETS = 24, M_ETH = 28, M_FLAGS = 30,
    M_SRCMAC = 32, M_DSTMAC = 38, M_IFINDEX = 44, M_LOCK = 48, M_SAMPLING = 52,
    M_DIR = 56, M_ERRNO = 57, M_DSCP = 58, M_NBOBS = 59, M_OBSDIR = 60, M_PAD0 = 66,
    M_OBSINTF = 68, M_SSLVER = 92, M_CIPHER = 94, M_KEYSHARE = 96, M_TLSTYPES = 98,
    M_MISC = 99, M_PAD1 = 100
};
/* dns_metrics (bpf/types.h:131-140) */
enum { D_START = 0, D_END = 8, D_LATENCY = 16, D_ID = 24, D_FLAGS = 26, D_ETH = 28, D_ERRNO = 30, D_NAME = 31 };
/* additional_metrics (bpf/types.h:174-181) */
enum { A_START = 0, A_END = 8, A_RTT = 16, A_IPSEC_RET = 24, A_ETH = 28, A_IPSEC_ENC = 30 };

static inline uint64_t ld64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline uint32_t ld32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static inline uint16_t ld16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static inline void st64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }
static inline void st32(uint8_t* p, uint32_t v) { memcpy(p, &v, 4); }
static inline void st16(uint8_t* p, uint16_t v) { memcpy(p, &v, 2); }

static inline int all_zero6(const uint8_t* m) { /* pkg/model/flow_content.go:200-207 AllZerosMac */
    return (m[0] | m[1] | m[2] | m[3] | m[4] | m[5]) == 0;
}

/* -------------------------------------------------------------- ReadFrom */
void oracle_read_from(const uint8_t* wire, uint8_t* rec) {
    /* pkg/model/record.go:227-231 + pkg/ebpf/bpf_x86_bpfel.go:110-153: blank fields
     * (`_ [1]byte`, `_ [2]byte`, `_ [4]byte`) are skipped by binary.Read -> zero. */
    memcpy(rec, wire, OR_REC_SIZE);
    rec[39] = 0;
    rec[OR_ID_SIZE + M_PAD0] = 0; rec[OR_ID_SIZE + M_PAD0 + 1] = 0;
    memset(rec + OR_ID_SIZE + M_PAD1, 0, 4);
}

/* --------------------------------------------------------- AccumulateBase */
void oracle_accumulate_base(uint8_t* p, const uint8_t* o) {
    /* pkg/model/flow_content.go:28-61 */
    uint64_t ps = ld64(p + M_START), os = ld64(o + M_START);
    if (ps == 0 || (ps > os && os != 0)) st64(p + M_START, os);             /* :36-38 */
    uint64_t pe = ld64(p + M_END), oe = ld64(o + M_END);
    if (pe == 0 || pe < oe) st64(p + M_END, oe);                              /* :39-41 */
    st64(p + M_BYTES, ld64(p + M_BYTES) + ld64(o + M_BYTES));                 /* :42 */
    st32(p + M_PACKETS, ld32(p + M_PACKETS) + ld32(o + M_PACKETS));           /* :43 u32 wrap */
    st16(p + M_FLAGS, (uint16_t)(ld16(p + M_FLAGS) | ld16(o + M_FLAGS)));     /* :44 */
    if (ld16(o + M_ETH) != 0) st16(p + M_ETH, ld16(o + M_ETH));               /* :45-47 */
    if (all_zero6(p + M_SRCMAC)) memcpy(p + M_SRCMAC, o + M_SRCMAC, 6);       /* :48-50 */
    if (all_zero6(p + M_DSTMAC)) memcpy(p + M_DSTMAC, o + M_DSTMAC, 6);       /* :51-53 */
    if (o[M_DSCP] != 0) p[M_DSCP] = o[M_DSCP];                                /* :54-56 */
    if (ld32(o + M_SAMPLING) != 0) st32(p + M_SAMPLING, ld32(o + M_SAMPLING));/* :57-59 */
}

/* ------------------------------------------------------- feature folds */
static void build_base_from_additional(uint8_t* base, uint64_t start, uint64_t end, uint16_t eth) {
    /* pkg/model/flow_content.go:63-74 */
    uint64_t bs = ld64(base + M_START);
    if (bs == 0 || (bs > start && start != 0)) st64(base + M_START, start);
    uint64_t be = ld64(base + M_END);
    if (be == 0 || be < end) st64(base + M_END, end);
    if (ld16(base + M_ETH) == 0) st16(base + M_ETH, eth);
}

static void merge_dns_block(oracle_content* p, const uint8_t* o) {
    /* pkg/model/flow_content.go:81-95 */
    if (!p->has_dns) { memcpy(p->dns, o, OR_DNS_SIZE); p->has_dns = 1; return; }   /* :81-84 */
    st16(p->dns + D_FLAGS, (uint16_t)(ld16(p->dns + D_FLAGS) | ld16(o + D_FLAGS))); /* :86 */
    if (ld16(o + D_ID) != 0) st16(p->dns + D_ID, ld16(o + D_ID));                   /* :87-89 */
    if (p->dns[D_ERRNO] != o[D_ERRNO]) p->dns[D_ERRNO] = o[D_ERRNO];                /* :90-92 */
    if (ld64(p->dns + D_LATENCY) < ld64(o + D_LATENCY))                             /* :93-95 */
        st64(p->dns + D_LATENCY, ld64(o + D_LATENCY));
}
void oracle_accumulate_dns(oracle_content* p, const uint8_t* o) {
    /* pkg/model/flow_content.go:76-96 */
    build_base_from_additional(p->metrics, ld64(o + D_START), ld64(o + D_END), ld16(o + D_ETH));   /* :80 */
    merge_dns_block(p, o);
}

static void merge_additional_block(oracle_content* p, const uint8_t* o) {
    /* pkg/model/flow_content.go:159-176 */
    if (!p->has_additional) { memcpy(p->additional, o, OR_ADD_SIZE); p->has_additional = 1; return; }
    if (ld64(p->additional + A_RTT) < ld64(o + A_RTT))                               /* :164-166 */
        st64(p->additional + A_RTT, ld64(o + A_RTT));
    int32_t pr = (int32_t)ld32(p->additional + A_IPSEC_RET), orr = (int32_t)ld32(o + A_IPSEC_RET);
    if (pr < orr) {                                                                  /* :168-171 */
        p->additional[A_IPSEC_ENC] = o[A_IPSEC_ENC];
        st32(p->additional + A_IPSEC_RET, (uint32_t)orr);
        pr = orr;
    }
    if (pr == orr && o[A_IPSEC_ENC]) p->additional[A_IPSEC_ENC] = o[A_IPSEC_ENC];    /* :172-176 */
}
void oracle_accumulate_additional(oracle_content* p, const uint8_t* o) {
    /* pkg/model/flow_content.go:154-177 */
    build_base_from_additional(p->metrics, ld64(o + A_START), ld64(o + A_END), ld16(o + A_ETH));   /* :158 */
    merge_additional_block(p, o);
}

/* pkt_drop_metrics: start@0 end@8 bytes u16@16 packets u16@18 latest_drop_cause u32@20 latest_flags u16@24
 * eth_protocol u16@26 latest_state u8@28 (bpf/types.h:142-151) */
enum { P_START = 0, P_END = 8, P_BYTES = 16, P_PACKETS = 18, P_CAUSE = 20, P_FLAGS = 24, P_ETH = 26, P_STATE = 28 };
static uint16_t add_u16_sat(uint16_t a, uint16_t b) { uint16_t x = (uint16_t)(a + b); return x < a ? 0xFFFF : x; }   /* flow_content.go:209-215 */
void oracle_accumulate_drops(uint8_t* p_metrics104, uint8_t* p_drops32, uint8_t* has_drops, const uint8_t* o) {
    /* pkg/model/flow_content.go:98-117 */
    build_base_from_additional(p_metrics104, ld64(o + P_START), ld64(o + P_END), ld16(o + P_ETH));        /* :102 */
    if (!*has_drops) { memcpy(p_drops32, o, OR_DROP_SIZE); memset(p_drops32 + 29, 0, 3); *has_drops = 1; return; }   /* :103-106 */
    st16(p_drops32 + P_BYTES, add_u16_sat(ld16(p_drops32 + P_BYTES), ld16(o + P_BYTES)));                 /* :108 */
    st16(p_drops32 + P_PACKETS, add_u16_sat(ld16(p_drops32 + P_PACKETS), ld16(o + P_PACKETS)));           /* :109 */
    st16(p_drops32 + P_FLAGS, (uint16_t)(ld16(p_drops32 + P_FLAGS) | ld16(o + P_FLAGS)));                 /* :110 */
    if (ld32(o + P_CAUSE) != 0) st32(p_drops32 + P_CAUSE, ld32(o + P_CAUSE));                             /* :111-113 */
    if (o[P_STATE] != 0) p_drops32[P_STATE] = o[P_STATE];                                                 /* :114-116 */
}

void oracle_new_record_times(uint64_t now_unix_ns, uint64_t mono_now_ns, uint64_t start_mono,
                             uint64_t end_mono, uint64_t* tfs, uint64_t* tfe) {
    /* pkg/model/record.go:90-97: time.Duration(monoNow - start) then now.Add(-delta) */
    *tfs = now_unix_ns - (mono_now_ns - start_mono);
    *tfe = now_unix_ns - (mono_now_ns - end_mono);
}

/* ------------------------------------------------------------ hash spec */
static inline uint64_t fmix64(uint64_t x) {
    x ^= x >> 33; x *= 0xFF51AFD7ED558CCDull; x ^= x >> 33; x *= 0xC4CEB9FE1A85EC53ull; x ^= x >> 33;
    return x;
}
uint64_t oracle_key_premix(const uint8_t* k) {
    /* 5 LE words, byte 39 (padding) masked out */
    const uint64_t P1 = 0x9E3779B97F4A7C15ull, P2 = 0xC2B2AE3D27D4EB4Full;
    uint64_t w0 = ld64(k), w1 = ld64(k + 8), w2 = ld64(k + 16), w3 = ld64(k + 24),
             w4 = ld64(k + 32) & 0x00FFFFFFFFFFFFFFull;
    uint64_t h = 0x243F6A8885A308D3ull;
    h = (h ^ w0) * P1; h ^= h >> 32;
    h = (h ^ w1) * P2; h ^= h >> 29;
    h = (h ^ w2) * P1; h ^= h >> 32;
    h = (h ^ w3) * P2; h ^= h >> 29;
    h = (h ^ w4) * P1; h ^= h >> 32;
    return h;
}
uint64_t oracle_slot_hash(const uint8_t* k) { return fmix64(oracle_key_premix(k)); }
uint64_t oracle_owner_hash(const uint8_t* k) { return fmix64(oracle_key_premix(k) ^ 0xA0761D6478BD642Full); }

/* ------------------------------------------------------------- flat map */
typedef struct {
    uint8_t key[OR_ID_SIZE];
    oracle_content c;
    /* flowmap only: what the feature samples contribute to the base via
     * buildBaseFromAdditional, kept apart so that the result does not depend on how
     * fa_ingest / fa_ingest_dns / fa_ingest_additional calls interleave; combined at
     * evict in LookupAndDeleteMap's order: base, then DNS, then additional
     * (pkg/tracer/tracer.go:1094-1151). */
    uint8_t  has_base;
    uint64_t fs[3], fe[3];     /* [0]=dns [1]=additional [2]=packet drops: min non-zero start / max end */
    uint16_t feth[3];          /* first non-zero eth_protocol in sample order */
    uint8_t  drops[OR_DROP_SIZE]; uint8_t has_drops;     /* pkt_drop_metrics block (bpf/types.h:142-151) */
    uint64_t rtt_min;          /* extension (no reference analogue): smallest non-zero flow_rtt, 0 = none */
} entry_t;

typedef struct {
    uint32_t* idx;     /* 0 = empty, else entry index + 1 */
    size_t    cap;     /* power of two */
    entry_t*  ents;
    size_t    n, ents_cap;
} fmap;

static void fmap_init(fmap* m, size_t hint) {
    size_t cap = 64; while (cap < hint * 2) cap <<= 1;
    m->cap = cap; m->idx = (uint32_t*)calloc(cap, sizeof(uint32_t));
    m->ents_cap = hint < 16 ? 16 : hint; m->ents = (entry_t*)malloc(m->ents_cap * sizeof(entry_t)); m->n = 0;
}
static void fmap_free(fmap* m) { free(m->idx); free(m->ents); m->idx = NULL; m->ents = NULL; m->n = 0; }
static void fmap_clear(fmap* m) { memset(m->idx, 0, m->cap * sizeof(uint32_t)); m->n = 0; }
static void fmap_grow(fmap* m) {
    size_t ncap = m->cap * 2; uint32_t* nidx = (uint32_t*)calloc(ncap, sizeof(uint32_t));
    for (size_t i = 0; i < m->n; i++) {
        size_t s = (size_t)oracle_slot_hash(m->ents[i].key) & (ncap - 1);
        while (nidx[s]) s = (s + 1) & (ncap - 1);
        nidx[s] = (uint32_t)(i + 1);
    }
    free(m->idx); m->idx = nidx; m->cap = ncap;
}
/* returns entry; *found tells whether it pre-existed. key compare = 40 bytes (pad already zeroed) */
static entry_t* fmap_get(fmap* m, const uint8_t* key, int create, int* found) {
    size_t s = (size_t)oracle_slot_hash(key) & (m->cap - 1);
    for (;;) {
        uint32_t v = m->idx[s];
        if (!v) break;
        entry_t* e = &m->ents[v - 1];
        if (memcmp(e->key, key, OR_ID_SIZE) == 0) { *found = 1; return e; }
        s = (s + 1) & (m->cap - 1);
    }
    *found = 0;
    if (!create) return NULL;
    if ((m->n + 1) * 2 > m->cap) { fmap_grow(m); return fmap_get(m, key, create, found); }
    if (m->n == m->ents_cap) { m->ents_cap *= 2; m->ents = (entry_t*)realloc(m->ents, m->ents_cap * sizeof(entry_t)); }
    entry_t* e = &m->ents[m->n];
    memset(e, 0, sizeof(*e));
    memcpy(e->key, key, OR_ID_SIZE);
    m->idx[s] = (uint32_t)(++m->n);
    return e;
}
static size_t fmap_dump_records(const fmap* m, uint8_t* out, size_t cap) {
    size_t k = m->n < cap ? m->n : cap;
    for (size_t i = 0; i < k; i++) {
        memcpy(out + i * OR_REC_SIZE, m->ents[i].key, OR_ID_SIZE);
        memcpy(out + i * OR_REC_SIZE + OR_ID_SIZE, m->ents[i].c.metrics, OR_MET_SIZE);
    }
    return m->n;
}

/* ------------------------------------------------------------ Accounter */
/* Lean flat map for the Accounter: one 144-byte record (key + metrics) per flow, in insertion order. */
typedef struct {
    uint32_t* idx; size_t cap;       /* open addressing; 0 = empty, else entry index + 1 */
    uint8_t*  recs; size_t n, recs_cap;
} amap;
static void amap_init(amap* m, size_t hint) {
    size_t cap = 64; while (cap < hint * 2) cap <<= 1;
    m->cap = cap; m->idx = (uint32_t*)calloc(cap, sizeof(uint32_t));
    m->recs_cap = hint < 16 ? 16 : hint; m->recs = (uint8_t*)malloc(m->recs_cap * OR_REC_SIZE); m->n = 0;
}
static void amap_free(amap* m) { free(m->idx); free(m->recs); m->idx = NULL; m->recs = NULL; m->n = 0; }
static void amap_clear(amap* m) { memset(m->idx, 0, m->cap * sizeof(uint32_t)); m->n = 0; }
static void amap_grow(amap* m) {
    size_t ncap = m->cap * 2; uint32_t* nidx = (uint32_t*)calloc(ncap, sizeof(uint32_t));
    for (size_t i = 0; i < m->n; i++) {
        size_t s = (size_t)oracle_slot_hash(m->recs + i * OR_REC_SIZE) & (ncap - 1);
        while (nidx[s]) s = (s + 1) & (ncap - 1);
        nidx[s] = (uint32_t)(i + 1);
    }
    free(m->idx); m->idx = nidx; m->cap = ncap;
}
/* the Accounter's per-record step (account.go:82-96 without the maxEntries test): fold or insert */
static inline void amap_account(amap* m, const uint8_t* rec /* after ReadFrom */) {
    size_t s = (size_t)oracle_slot_hash(rec) & (m->cap - 1);
    for (;;) {
        uint32_t v = m->idx[s];
        if (!v) break;
        uint8_t* e = m->recs + (size_t)(v - 1) * OR_REC_SIZE;
        if (memcmp(e, rec, OR_ID_SIZE) == 0) { oracle_accumulate_base(e + OR_ID_SIZE, rec + OR_ID_SIZE); return; }
        s = (s + 1) & (m->cap - 1);
    }
    if ((m->n + 1) * 2 > m->cap) { amap_grow(m); amap_account(m, rec); return; }
    if (m->n == m->recs_cap) { m->recs_cap *= 2; m->recs = (uint8_t*)realloc(m->recs, m->recs_cap * OR_REC_SIZE); }
    memcpy(m->recs + m->n * OR_REC_SIZE, rec, OR_REC_SIZE);               /* account.go:95: whole 104 B kept */
    m->idx[s] = (uint32_t)(++m->n);
}
static inline int amap_has(const amap* m, const uint8_t* rec) {
    size_t s = (size_t)oracle_slot_hash(rec) & (m->cap - 1);
    for (;;) {
        uint32_t v = m->idx[s];
        if (!v) return 0;
        if (memcmp(m->recs + (size_t)(v - 1) * OR_REC_SIZE, rec, OR_ID_SIZE) == 0) return 1;
        s = (s + 1) & (m->cap - 1);
    }
}

typedef struct generation { uint8_t* recs; size_t n; struct generation* next; } generation;
struct oracle_accounter {
    size_t max_entries;
    amap   entries;
    generation *gen_head, *gen_tail; size_t gen_pending;
};

oracle_accounter* oracle_accounter_new(size_t max_entries) {
    oracle_accounter* a = (oracle_accounter*)calloc(1, sizeof(*a));
    a->max_entries = max_entries;
    amap_init(&a->entries, max_entries < (1u << 20) ? max_entries : (1u << 20));
    return a;
}
void oracle_accounter_free(oracle_accounter* a) {
    if (!a) return;
    while (a->gen_head) { generation* g = a->gen_head; a->gen_head = g->next; free(g->recs); free(g); }
    amap_free(&a->entries); free(a);
}
size_t oracle_accounter_len(const oracle_accounter* a) { return a->entries.n; }

static void accounter_push_generation(oracle_accounter* a) {
    generation* g = (generation*)calloc(1, sizeof(*g));
    g->n = a->entries.n; g->recs = (uint8_t*)malloc(g->n ? g->n * OR_REC_SIZE : 1);
    memcpy(g->recs, a->entries.recs, g->n * OR_REC_SIZE);
    if (a->gen_tail) a->gen_tail->next = g; else a->gen_head = g;
    a->gen_tail = g; a->gen_pending++;
    amap_clear(&a->entries);
}

void oracle_accounter_account(oracle_accounter* a, const uint8_t* wire, size_t n) {
    uint8_t rec[OR_REC_SIZE];
    for (size_t i = 0; i < n; i++) {
        oracle_read_from(wire + i * OR_REC_SIZE, rec);                  /* tracer_ringbuf.go:112-134 */
        if (a->entries.n >= a->max_entries && !amap_has(&a->entries, rec))
            accounter_push_generation(a);                               /* account.go:85-94: new key, cache full */
        amap_account(&a->entries, rec);                                 /* account.go:82-83 / :95 */
    }
}
size_t oracle_accounter_evict(oracle_accounter* a, uint8_t* out, size_t cap) {
    size_t k = a->entries.n < cap ? a->entries.n : cap;                 /* account.go:63-80,102-124 */
    memcpy(out, a->entries.recs, k * OR_REC_SIZE);
    size_t n = a->entries.n;
    amap_clear(&a->entries);
    return n;
}
size_t oracle_accounter_pending(const oracle_accounter* a) { return a->gen_pending; }
size_t oracle_accounter_next_generation_len(const oracle_accounter* a) { return a->gen_head ? a->gen_head->n : 0; }
size_t oracle_accounter_pop_generation(oracle_accounter* a, uint8_t* out, size_t cap) {
    generation* g = a->gen_head; if (!g) return 0;
    size_t k = g->n < cap ? g->n : cap; memcpy(out, g->recs, k * OR_REC_SIZE);
    size_t n = g->n; a->gen_head = g->next; if (!a->gen_head) a->gen_tail = NULL; a->gen_pending--;
    free(g->recs); free(g);
    return n;
}

size_t oracle_accounter_sharded_run(const uint8_t* wire, size_t n, int n_threads, uint8_t* out, size_t cap) {
    /* CPU-baseline only (bench.py --impl reference): T private Accounters, thread t owns the keys with
     * owner_hash % T == t.  Pass 1 (parallel over contiguous chunks): owner of every record + per-chunk
     * histograms.  Pass 2: stable scatter of record indices into per-owner lists.  Pass 3: every thread folds
     * its own list in stream order.  No maxEntries cut. */
    if (n_threads < 1) n_threads = 1;
    const size_t T = (size_t)n_threads;
    uint16_t* owner = (uint16_t*)malloc((n ? n : 1) * sizeof(uint16_t));
    uint32_t* list = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t* hist = (size_t*)calloc(T * T + 1, sizeof(size_t));          /* hist[chunk][owner] */
    size_t* start = (size_t*)calloc(T + 1, sizeof(size_t));
    amap* maps = (amap*)calloc(T, sizeof(amap));
#pragma omp parallel num_threads(n_threads)
    {
#ifdef _OPENMP
        const size_t t = (size_t)omp_get_thread_num();
#else
        const size_t t = 0;
#endif
        const size_t lo = n * t / T, hi = n * (t + 1) / T;
        for (size_t i = lo; i < hi; i++) {
            const uint16_t o = (uint16_t)(oracle_owner_hash(wire + i * OR_REC_SIZE) % T);
            owner[i] = o; hist[t * T + o]++;
        }
#pragma omp barrier
#pragma omp single
        {
            size_t run = 0;
            for (size_t o = 0; o < T; o++) {
                start[o] = run;
                for (size_t c = 0; c < T; c++) { const size_t v = hist[c * T + o]; hist[c * T + o] = run; run += v; }
            }
            start[T] = run;
        }
        for (size_t i = lo; i < hi; i++) list[hist[t * T + owner[i]]++] = (uint32_t)i;
#pragma omp barrier
        amap* m = &maps[t];
        amap_init(m, (start[t + 1] - start[t]) / 4 + 1024);
        uint8_t rec[OR_REC_SIZE];
        for (size_t k = start[t]; k < start[t + 1]; k++) {
            oracle_read_from(wire + (size_t)list[k] * OR_REC_SIZE, rec);
            amap_account(m, rec);
        }
    }
    size_t total = 0;
    for (size_t t = 0; t < T; t++) {
        if (out && total < cap) {
            const size_t k = maps[t].n < cap - total ? maps[t].n : cap - total;
            memcpy(out + total * OR_REC_SIZE, maps[t].recs, k * OR_REC_SIZE);
        }
        total += maps[t].n; amap_free(&maps[t]);
    }
    free(maps); free(owner); free(list); free(hist); free(start);
    return total;
}

/* persistent shards: same three passes, the maps survive the call */
struct oracle_sharded { int T; amap* maps; };
oracle_sharded* oracle_sharded_new(int n_threads) {
    oracle_sharded* s = (oracle_sharded*)calloc(1, sizeof(*s));
    s->T = n_threads < 1 ? 1 : n_threads;
    s->maps = (amap*)calloc((size_t)s->T, sizeof(amap));
    for (int t = 0; t < s->T; t++) amap_init(&s->maps[t], 1024);
    return s;
}
void oracle_sharded_free(oracle_sharded* s) { if (s) { for (int t = 0; t < s->T; t++) amap_free(&s->maps[t]); free(s->maps); free(s); } }
size_t oracle_sharded_len(const oracle_sharded* s) { size_t n = 0; for (int t = 0; t < s->T; t++) n += s->maps[t].n; return n; }
void oracle_sharded_account(oracle_sharded* s, const uint8_t* wire, size_t n) {
    const size_t T = (size_t)s->T;
    uint16_t* owner = (uint16_t*)malloc((n ? n : 1) * sizeof(uint16_t));
    uint32_t* list = (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t));
    size_t* hist = (size_t*)calloc(T * T + 1, sizeof(size_t));
    size_t* start = (size_t*)calloc(T + 1, sizeof(size_t));
#pragma omp parallel num_threads(s->T)
    {
#ifdef _OPENMP
        const size_t t = (size_t)omp_get_thread_num();
#else
        const size_t t = 0;
#endif
        const size_t lo = n * t / T, hi = n * (t + 1) / T;
        for (size_t i = lo; i < hi; i++) {
            const uint16_t o = (uint16_t)(oracle_owner_hash(wire + i * OR_REC_SIZE) % T);
            owner[i] = o; hist[t * T + o]++;
        }
#pragma omp barrier
#pragma omp single
        {
            size_t run = 0;
            for (size_t o = 0; o < T; o++) {
                start[o] = run;
                for (size_t c = 0; c < T; c++) { const size_t v = hist[c * T + o]; hist[c * T + o] = run; run += v; }
            }
            start[T] = run;
        }
        for (size_t i = lo; i < hi; i++) list[hist[t * T + owner[i]]++] = (uint32_t)i;
#pragma omp barrier
        uint8_t rec[OR_REC_SIZE];
        for (size_t k = start[t]; k < start[t + 1]; k++) {
            oracle_read_from(wire + (size_t)list[k] * OR_REC_SIZE, rec);
            amap_account(&s->maps[t], rec);
        }
    }
    free(owner); free(list); free(hist); free(start);
}
size_t oracle_sharded_evict(oracle_sharded* s, uint8_t* out, size_t cap) {
    size_t total = 0;
    for (int t = 0; t < s->T; t++) {
        amap* m = &s->maps[t];
        if (out && total < cap) {
            const size_t k = m->n < cap - total ? m->n : cap - total;
            memcpy(out + total * OR_REC_SIZE, m->recs, k * OR_REC_SIZE);
        }
        total += m->n;
        amap_clear(m);
    }
    return total;
}

/* -------------------------------------------------------------- flowmap */
struct oracle_flowmap { fmap m; };
oracle_flowmap* oracle_flowmap_new(void) { oracle_flowmap* m = (oracle_flowmap*)calloc(1, sizeof(*m)); fmap_init(&m->m, 1024); return m; }
void oracle_flowmap_free(oracle_flowmap* m) { if (m) { fmap_free(&m->m); free(m); } }
size_t oracle_flowmap_len(const oracle_flowmap* m) { return m->m.n; }

void oracle_flowmap_account(oracle_flowmap* m, const uint8_t* wire, size_t n) {
    uint8_t rec[OR_REC_SIZE];
    for (size_t i = 0; i < n; i++) {
        oracle_read_from(wire + i * OR_REC_SIZE, rec);
        int found; entry_t* e = fmap_get(&m->m, rec, 1, &found);
        if (e->has_base) oracle_accumulate_base(e->c.metrics, rec + OR_ID_SIZE);   /* account.go:82-83 */
        else { memcpy(e->c.metrics, rec + OR_ID_SIZE, OR_MET_SIZE); e->has_base = 1; }  /* account.go:95 */
    }
}
static void feature_base_effect(entry_t* e, int which, uint64_t s, uint64_t en, uint16_t eth) {
    if (e->fs[which] == 0 || (e->fs[which] > s && s != 0)) e->fs[which] = s;
    if (e->fe[which] == 0 || e->fe[which] < en) e->fe[which] = en;
    if (e->feth[which] == 0) e->feth[which] = eth;
}
void oracle_flowmap_fold_dns(oracle_flowmap* m, const uint8_t* recs, size_t n) {
    uint8_t key[OR_ID_SIZE];
    for (size_t i = 0; i < n; i++) {
        memcpy(key, recs + i * OR_DNSREC_SIZE, OR_ID_SIZE); key[39] = 0;
        int found; entry_t* e = fmap_get(&m->m, key, 1, &found);      /* empty base if absent: tracer.go:1179-1182 */
        uint8_t dns[OR_DNS_SIZE]; memcpy(dns, recs + i * OR_DNSREC_SIZE + OR_ID_SIZE, OR_DNS_SIZE);
        dns[63] = 0;                                                  /* trailing pad: blank in the Go mirror */
        feature_base_effect(e, 0, ld64(dns + D_START), ld64(dns + D_END), ld16(dns + D_ETH));
        merge_dns_block(&e->c, dns);
    }
}
void oracle_flowmap_fold_additional(oracle_flowmap* m, const uint8_t* recs, size_t n) {
    uint8_t key[OR_ID_SIZE];
    for (size_t i = 0; i < n; i++) {
        memcpy(key, recs + i * OR_ADDREC_SIZE, OR_ID_SIZE); key[39] = 0;
        int found; entry_t* e = fmap_get(&m->m, key, 1, &found);
        uint8_t add[OR_ADD_SIZE]; memcpy(add, recs + i * OR_ADDREC_SIZE + OR_ID_SIZE, OR_ADD_SIZE);
        add[31] = 0;
        feature_base_effect(e, 1, ld64(add + A_START), ld64(add + A_END), ld16(add + A_ETH));
        merge_additional_block(&e->c, add);
        const uint64_t rtt = ld64(add + A_RTT);
        if (rtt != 0 && (e->rtt_min == 0 || rtt < e->rtt_min)) e->rtt_min = rtt;
    }
}
void oracle_flowmap_fold_drops(oracle_flowmap* m, const uint8_t* recs, size_t n) {
    uint8_t key[OR_ID_SIZE];
    for (size_t i = 0; i < n; i++) {
        memcpy(key, recs + i * OR_DROPREC_SIZE, OR_ID_SIZE); key[39] = 0;
        int found; entry_t* e = fmap_get(&m->m, key, 1, &found);
        const uint8_t* o = recs + i * OR_DROPREC_SIZE + OR_ID_SIZE;
        feature_base_effect(e, 2, ld64(o + P_START), ld64(o + P_END), ld16(o + P_ETH));
        uint8_t scratch_base[OR_MET_SIZE]; memset(scratch_base, 0, sizeof scratch_base);   /* base effect is kept apart (see entry_t) */
        oracle_accumulate_drops(scratch_base, e->drops, &e->has_drops, o);
    }
}
size_t oracle_flowmap_evict_ex(oracle_flowmap* m, uint8_t* out, uint8_t* out_dns, uint8_t* out_add, uint8_t* out_drops,
                               uint64_t* out_rtt_min, uint8_t* out_present, size_t cap) {
    size_t k = m->m.n < cap ? m->m.n : cap;
    for (size_t i = 0; i < k; i++) {
        const entry_t* e = &m->m.ents[i];
        uint8_t met[OR_MET_SIZE];
        if (e->has_base) memcpy(met, e->c.metrics, OR_MET_SIZE); else memset(met, 0, OR_MET_SIZE);
        /* LookupAndDeleteMap's order: DNS, packet drops, additional (pkg/tracer/tracer.go:1098-1151) */
        if (e->c.has_dns) build_base_from_additional(met, e->fs[0], e->fe[0], e->feth[0]);
        if (e->has_drops) build_base_from_additional(met, e->fs[2], e->fe[2], e->feth[2]);
        if (e->c.has_additional) build_base_from_additional(met, e->fs[1], e->fe[1], e->feth[1]);
        memcpy(out + i * OR_REC_SIZE, e->key, OR_ID_SIZE);
        memcpy(out + i * OR_REC_SIZE + OR_ID_SIZE, met, OR_MET_SIZE);
        if (out_dns) { if (e->c.has_dns) memcpy(out_dns + i * OR_DNS_SIZE, e->c.dns, OR_DNS_SIZE); else memset(out_dns + i * OR_DNS_SIZE, 0, OR_DNS_SIZE); }
        if (out_add) { if (e->c.has_additional) memcpy(out_add + i * OR_ADD_SIZE, e->c.additional, OR_ADD_SIZE); else memset(out_add + i * OR_ADD_SIZE, 0, OR_ADD_SIZE); }
        if (out_drops) { if (e->has_drops) memcpy(out_drops + i * OR_DROP_SIZE, e->drops, OR_DROP_SIZE); else memset(out_drops + i * OR_DROP_SIZE, 0, OR_DROP_SIZE); }
        if (out_rtt_min) out_rtt_min[i] = e->rtt_min;
        if (out_present) out_present[i] = (uint8_t)((e->c.has_dns ? 1 : 0) | (e->c.has_additional ? 2 : 0) | (e->has_drops ? 4 : 0));
    }
    size_t n = m->m.n; fmap_clear(&m->m);
    return n;
}
size_t oracle_flowmap_evict(oracle_flowmap* m, uint8_t* out, uint8_t* out_dns, uint8_t* out_add,
                            uint8_t* out_present, size_t cap) {
    return oracle_flowmap_evict_ex(m, out, out_dns, out_add, NULL, NULL, out_present, cap);
}

/* ----------------------------------------------------------- KERNEL_MAP */
struct oracle_kmap {
    size_t max_entries; int ringbuf;
    fmap m;
    uint8_t* spill; size_t n_spill, spill_cap;
    uint64_t fail_create, intf_missed;
};
oracle_kmap* oracle_kmap_new(size_t max_entries, int ringbuf_fallback) {
    oracle_kmap* k = (oracle_kmap*)calloc(1, sizeof(*k));
    k->max_entries = max_entries; k->ringbuf = ringbuf_fallback;
    fmap_init(&k->m, max_entries < (1u << 20) ? max_entries : (1u << 20));
    return k;
}
void oracle_kmap_free(oracle_kmap* k) { if (k) { fmap_free(&k->m); free(k->spill); free(k); } }
size_t oracle_kmap_len(const oracle_kmap* k) { return k->m.n; }
uint64_t oracle_kmap_counter_fail_create(const oracle_kmap* k) { return k->fail_create; }
uint64_t oracle_kmap_counter_intf_missed(const oracle_kmap* k) { return k->intf_missed; }

static int add_observed_intf(uint8_t* v, uint32_t if_index, uint8_t direction) {
    /* bpf/flows.c:76-96 */
    uint8_t nb = v[M_NBOBS];
    if (nb >= 6) return 1;
    for (uint8_t i = 0; i < nb; i++) {
        if (ld32(v + M_OBSINTF + 4 * i) == if_index) {
            if (v[M_OBSDIR + i] != direction && v[M_OBSDIR + i] != 3) v[M_OBSDIR + i] = 3;
            return 0;
        }
    }
    st32(v + M_OBSINTF + 4 * nb, if_index);
    v[M_OBSDIR + nb] = direction;
    v[M_NBOBS] = (uint8_t)(nb + 1);
    return 0;
}

static void update_existing_flow(oracle_kmap* k, uint8_t* agg, const uint8_t* key, const uint8_t* ev) {
    /* bpf/flows.c:98-143; ev = the packet event's metrics block */
    uint32_t if_index = ld32(ev + M_IFINDEX);
    uint64_t ts = ld64(ev + M_START);
    if (ld32(agg + M_IFINDEX) == if_index) {
        st32(agg + M_PACKETS, ld32(agg + M_PACKETS) + 1);
        st64(agg + M_BYTES, ld64(agg + M_BYTES) + ld64(ev + M_BYTES));
        st64(agg + M_END, ts);                                         /* last writer, not max */
        st16(agg + M_FLAGS, (uint16_t)(ld16(agg + M_FLAGS) | ld16(ev + M_FLAGS)));
        agg[M_DSCP] = ev[M_DSCP];
        st32(agg + M_SAMPLING, ld32(ev + M_SAMPLING));
        uint16_t hv = ld16(ev + M_SSLVER); uint8_t ty = ev[M_TLSTYPES];
        if (hv > 0 && ld16(agg + M_SSLVER) != hv) {
            if (ld16(agg + M_SSLVER) == 0) st16(agg + M_SSLVER, hv);
            else agg[M_MISC] |= 0x01;                                  /* MISC_FLAGS_SSL_MISMATCH */
        }
        if (ld16(ev + M_CIPHER) > 0 && ty == 0x02) st16(agg + M_CIPHER, ld16(ev + M_CIPHER));
        if (ld16(ev + M_KEYSHARE) > 0 && ty == 0x02) st16(agg + M_KEYSHARE, ld16(ev + M_KEYSHARE));
        agg[M_TLSTYPES] |= ty;
    } else if (if_index != 0) {
        st64(agg + M_END, ts);
        st16(agg + M_FLAGS, (uint16_t)(ld16(agg + M_FLAGS) | ld16(ev + M_FLAGS)));
        if (add_observed_intf(agg, if_index, ev[M_DIR]) > 0 && key[36] != 0) k->intf_missed++;  /* :134-142 */
    }
}

/* flows.c:228-245 new_flow from a packet event */
static void kmap_new_flow(const uint8_t* ev, uint8_t* nf) {
    memset(nf, 0, OR_MET_SIZE);
    st32(nf + M_IFINDEX, ld32(ev + M_IFINDEX)); nf[M_DIR] = ev[M_DIR];
    st32(nf + M_PACKETS, 1); st64(nf + M_BYTES, ld64(ev + M_BYTES));
    st16(nf + M_ETH, ld16(ev + M_ETH));
    st64(nf + M_START, ld64(ev + M_START)); st64(nf + M_END, ld64(ev + M_START));
    st16(nf + M_FLAGS, ld16(ev + M_FLAGS)); nf[M_DSCP] = ev[M_DSCP];
    st32(nf + M_SAMPLING, ld32(ev + M_SAMPLING));
    memcpy(nf + M_DSTMAC, ev + M_DSTMAC, 6); memcpy(nf + M_SRCMAC, ev + M_SRCMAC, 6);
    st16(nf + M_SSLVER, ld16(ev + M_SSLVER)); st16(nf + M_CIPHER, ld16(ev + M_CIPHER));
    st16(nf + M_KEYSHARE, ld16(ev + M_KEYSHARE)); nf[M_TLSTYPES] = ev[M_TLSTYPES];
}

void oracle_kmap_packets(oracle_kmap* k, const uint8_t* wire, size_t n) {
    uint8_t rec[OR_REC_SIZE];
    for (size_t i = 0; i < n; i++) {
        oracle_read_from(wire + i * OR_REC_SIZE, rec);
        const uint8_t* ev = rec + OR_ID_SIZE;
        int found; entry_t* e = fmap_get(&k->m, rec, 0, &found);        /* flows.c:222 */
        if (found) { update_existing_flow(k, e->c.metrics, rec, ev); continue; }
        uint8_t nf[OR_MET_SIZE]; kmap_new_flow(ev, nf);
        if (k->m.n < k->max_entries) {                                  /* flows.c:247 BPF_NOEXIST */
            e = fmap_get(&k->m, rec, 1, &found);
            memcpy(e->c.metrics, nf, OR_MET_SIZE);
        } else if (k->ringbuf) {                                        /* flows.c:262-279, errno = E2BIG */
            nf[M_ERRNO] = 7;
            if (k->n_spill == k->spill_cap) { k->spill_cap = k->spill_cap ? k->spill_cap * 2 : 64; k->spill = (uint8_t*)realloc(k->spill, k->spill_cap * OR_REC_SIZE); }
            memcpy(k->spill + k->n_spill * OR_REC_SIZE, rec, OR_ID_SIZE);
            memcpy(k->spill + k->n_spill * OR_REC_SIZE + OR_ID_SIZE, nf, OR_MET_SIZE);
            k->n_spill++;
        } else {
            k->fail_create++;                                           /* flows.c:285 */
        }
    }
}
size_t oracle_kmap_evict(oracle_kmap* k, uint8_t* out, size_t cap) {
    size_t n = fmap_dump_records(&k->m, out, cap); fmap_clear(&k->m); return n;
}
size_t oracle_kmap_spilled(oracle_kmap* k, uint8_t* out, size_t cap) {
    if (!out) return k->n_spill;                   /* peek */
    size_t c = k->n_spill < cap ? k->n_spill : cap;
    if (out && c) memcpy(out, k->spill, c * OR_REC_SIZE);
    size_t n = k->n_spill; k->n_spill = 0; return n;
}

/* LookupAndDeleteMap's merged view when the BASE is the kernel map (the reference's real deployment:
 * aggregated_flows + the per-CPU feature maps, pkg/tracer/tracer.go:1063-1157).  No capacity limit here. */
uint64_t oracle_flowmap_packets_kmap(oracle_flowmap* m, const uint8_t* wire, size_t n) {
    oracle_kmap counters; memset(&counters, 0, sizeof counters);
    uint8_t rec[OR_REC_SIZE];
    for (size_t i = 0; i < n; i++) {
        oracle_read_from(wire + i * OR_REC_SIZE, rec);
        int found; entry_t* e = fmap_get(&m->m, rec, 1, &found);
        if (e->has_base) update_existing_flow(&counters, e->c.metrics, rec, rec + OR_ID_SIZE);
        else { kmap_new_flow(rec + OR_ID_SIZE, e->c.metrics); e->has_base = 1; }
    }
    return counters.intf_missed;
}

/* ------------------------------------------------------------- sketches */
void oracle_cms_update(uint64_t* table, uint32_t lw, uint32_t depth, uint64_t seed, const uint8_t* wire, size_t n) {
    for (size_t i = 0; i < n; i++) {
        const uint8_t* r = wire + i * OR_REC_SIZE;
        uint64_t h = oracle_key_premix(r);
        uint64_t a = fmix64(h ^ 0xE7037ED1A0B428DBull ^ seed);
        uint64_t b = fmix64(h ^ 0x8EBC6AF09C88C6E3ull ^ seed) | 1ull;
        uint64_t w = ld32(r + OR_ID_SIZE + M_PACKETS);
        for (uint32_t d = 0; d < depth; d++) {
            uint64_t g = (a + (uint64_t)d * b) * 0x9E3779B97F4A7C15ull;
            table[((size_t)d << lw) + (size_t)(g >> (64 - lw))] += w;
        }
    }
}
void oracle_cms_query(const uint64_t* table, uint32_t lw, uint32_t depth, uint64_t seed, const uint8_t* keys, size_t n, uint64_t* est) {
    for (size_t i = 0; i < n; i++) {
        uint64_t h = oracle_key_premix(keys + i * OR_ID_SIZE);
        uint64_t a = fmix64(h ^ 0xE7037ED1A0B428DBull ^ seed);
        uint64_t b = fmix64(h ^ 0x8EBC6AF09C88C6E3ull ^ seed) | 1ull;
        uint64_t m = UINT64_MAX;
        for (uint32_t d = 0; d < depth; d++) {
            uint64_t g = (a + (uint64_t)d * b) * 0x9E3779B97F4A7C15ull;
            uint64_t v = table[((size_t)d << lw) + (size_t)(g >> (64 - lw))];
            if (v < m) m = v;
        }
        est[i] = m;
    }
}
void oracle_hll_update(uint8_t* regs, uint32_t p, uint64_t seed, const uint8_t* wire, size_t n) {
    for (size_t i = 0; i < n; i++) {
        uint64_t h = fmix64(oracle_key_premix(wire + i * OR_REC_SIZE) ^ 0x589965CC75374CC3ull ^ seed);
        uint32_t idx = (uint32_t)(h >> (64 - p));
        uint64_t rest = h << p;                       /* remaining 64-p bits, left aligned */
        uint32_t rho = rest ? (uint32_t)__builtin_clzll(rest) + 1 : (64 - p) + 1;
        if (rho > 64 - p + 1) rho = 64 - p + 1;
        if (regs[idx] < rho) regs[idx] = (uint8_t)rho;
    }
}
double oracle_hll_estimate(const uint8_t* regs, uint32_t p) {
    size_t m = (size_t)1 << p; double sum = 0.0; size_t zeros = 0;
    for (size_t i = 0; i < m; i++) { sum += ldexp(1.0, -(int)regs[i]); if (!regs[i]) zeros++; }
    double alpha = m >= 128 ? 0.7213 / (1.0 + 1.079 / (double)m) : (m == 64 ? 0.709 : (m == 32 ? 0.697 : 0.673));
    double e = alpha * (double)m * (double)m / sum;
    if (e <= 2.5 * (double)m && zeros) e = (double)m * log((double)m / (double)zeros);
    return e;
}


/* ---------------------------------------------------------------------------------------------------------------
 * (f4) raw-header front end.  TEST INFRASTRUCTURE like the rest of this file.
 * Follows bpf/utils.h:24-167 and bpf/flows.c:176-245; every bound check is the reference's "header end <= data_end". */
static uint16_t be16(const uint8_t* p) { return (uint16_t)((p[0] << 8) | p[1]); }

/* bpf/utils.h:24-50 set_flags: one flag per packet, first match of the chain */
static uint16_t snap_tcp_flags(uint8_t f) {
    const int fin = f & 0x01, syn = f & 0x02, rst = f & 0x04, psh = f & 0x08, ack = f & 0x10, urg = f & 0x20, ece = f & 0x40, cwr = f & 0x80;
    if (ack && syn) return 0x100;          /* SYN_ACK_FLAG */
    if (ack && fin) return 0x200;          /* FIN_ACK_FLAG */
    if (ack && rst) return 0x400;          /* RST_ACK_FLAG */
    if (fin) return 0x01;
    if (syn) return 0x02;
    if (ack) return 0x10;
    if (rst) return 0x04;
    if (psh) return 0x08;
    if (urg) return 0x20;
    if (ece) return 0x40;
    if (cwr) return 0x80;
    return 0;
}

int oracle_parse_snap(const uint8_t* snap, uint32_t stride, uint8_t* rec) {
    uint64_t ts; uint32_t len, ifindex, sampling; uint16_t cap;
    memcpy(&ts, snap, 8); memcpy(&len, snap + 8, 4); memcpy(&ifindex, snap + 12, 4); memcpy(&sampling, snap + 16, 4);
    memcpy(&cap, snap + 20, 2);
    const uint8_t direction = snap[22];
    const uint8_t* d = snap + 24;
    uint32_t end = cap;                                        /* data_end - data */
    if (end > stride - 24) end = stride - 24;
    memset(rec, 0, 144);
    if (14 > end) return 0;                                    /* utils.h:154-156 */
    const uint16_t eth = be16(d + 12);
    uint32_t l4; uint8_t proto, dscp;
    if (eth == 0x0800) {                                       /* utils.h:111-129: l4 = ip + sizeof(iphdr), options are not skipped */
        l4 = 14 + 20;
        if (l4 > end) return 0;
        static const uint8_t ip4in6[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff};
        memcpy(rec + 0, ip4in6, 12); memcpy(rec + 12, d + 14 + 12, 4);
        memcpy(rec + 16, ip4in6, 12); memcpy(rec + 28, d + 14 + 16, 4);
        dscp = (d[14 + 1] >> 2) & 0x3F;
        proto = d[14 + 9];
    } else if (eth == 0x86DD) {                                /* utils.h:132-149: nexthdr taken as the transport, no extension walk */
        l4 = 14 + 40;
        if (l4 > end) return 0;
        memcpy(rec + 0, d + 14 + 8, 16); memcpy(rec + 16, d + 14 + 24, 16);
        dscp = (uint8_t)(((be16(d + 14) >> 4) >> 2) & 0x3F);
        proto = d[14 + 6];
    } else {
        return 0;                                              /* utils.h:166: only IP-based flows */
    }
    uint16_t sport = 0, dport = 0, flags = 0; uint8_t itype = 0, icode = 0;
    switch (proto) {                                           /* utils.h:53-104 */
    case 6:   if (l4 + 20 <= end) { sport = be16(d + l4); dport = be16(d + l4 + 2); flags = snap_tcp_flags(d[l4 + 13]); } break;
    case 17:  if (l4 + 8 <= end)  { sport = be16(d + l4); dport = be16(d + l4 + 2); } break;
    case 132: if (l4 + 12 <= end) { sport = be16(d + l4); dport = be16(d + l4 + 2); } break;
    case 1:   if (l4 + 8 <= end)  { itype = d[l4]; icode = d[l4 + 1]; } break;
    case 58:  if (l4 + 8 <= end)  { itype = d[l4]; icode = d[l4 + 1]; } break;
    default: break;
    }
    memcpy(rec + 32, &sport, 2); memcpy(rec + 34, &dport, 2);
    rec[36] = proto; rec[37] = itype; rec[38] = icode;
    /* flows.c:226-245 new_flow */
    const uint64_t bytes = len; const uint32_t one = 1;
    memcpy(rec + 40, &ts, 8); memcpy(rec + 48, &ts, 8); memcpy(rec + 56, &bytes, 8); memcpy(rec + 64, &one, 4);
    memcpy(rec + 68, &eth, 2); memcpy(rec + 70, &flags, 2);
    memcpy(rec + 72, d + 6, 6);                                /* src_mac = eth->h_source */
    memcpy(rec + 78, d + 0, 6);                                /* dst_mac = eth->h_dest */
    memcpy(rec + 84, &ifindex, 4); memcpy(rec + 92, &sampling, 4);
    rec[96] = direction; rec[98] = dscp;
    return 1;
}

size_t oracle_parse_snaps(const uint8_t* snaps, size_t n, uint32_t stride, uint8_t* out, uint32_t* src_of) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++)
        if (oracle_parse_snap(snaps + i * stride, stride, out + m * 144)) { if (src_of) src_of[m] = (uint32_t)i; m++; }
    return m;
}


/* ---------------------------------------------------------------------------------------------------------------
 * Flow filter (bpf/flows_filter.h, bpf/utils.h:179-222).  Rule image (64 B): ip[16] @0, prefix_len u32 @16, sample u32 @20,
 * dst ports start/end/1/2 u16 @24, src ports @32, generic ports @40, tcp_flags u16 @48, protocol @50, icmp type/code @51,52,
 * direction @53, action @54, filter_drops @55, do_peer_cidr_lookup @56.  Peer CIDR image (20 B): ip[16], prefix_len u32. */
static uint16_t rd16(const uint8_t* p) { uint16_t v; memcpy(&v, p, 2); return v; }
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }

/* BPF_MAP_TYPE_LPM_TRIE lookup: the entry with the longest prefix_len <= key_prefix whose first prefix_len bits equal the key's */
static const uint8_t* lpm_lookup(const uint8_t* entries, size_t n, size_t entry_bytes, const uint8_t key[16], uint32_t key_prefix) {
    const uint8_t* best = NULL; uint32_t best_len = 0;
    for (size_t i = 0; i < n; i++) {
        const uint8_t* e = entries + i * entry_bytes;
        const uint32_t pl = rd32(e + 16);
        if (pl > key_prefix || pl > 128) continue;
        int ok = 1;
        for (uint32_t b = 0; b < pl / 8 && ok; b++) ok = e[b] == key[b];
        if (ok && (pl & 7)) { const uint8_t m = (uint8_t)(0xFF << (8 - (pl & 7))); ok = ((e[pl / 8] ^ key[pl / 8]) & m) == 0; }
        if (ok && (!best || pl > best_len)) { best = e; best_len = pl; }
    }
    return best;
}

/* flow_filter_setup_lookup_key (flows_filter.h:14-41) */
static void filter_key_of(const uint8_t* rec, int use_src, uint16_t eth, uint8_t key[16], uint32_t* prefix) {
    memset(key, 0, 16);
    const uint8_t* ip = rec + (use_src ? 0 : 16);
    if (eth == 0x0800) { memcpy(key, ip + 12, 4); *prefix = 32; } else { memcpy(key, ip, 16); *prefix = 128; }
}

/* do_flow_filter_lookup (flows_filter.h:43-217); action and sampling keep what a rule wrote even when it ends with 0 */
static int filter_lookup(const uint8_t* rules, size_t n_rules, const uint8_t* peers, size_t n_peers, const uint8_t* rec,
                         const uint8_t key[16], uint32_t prefix, int peer_use_src, uint16_t eth, uint8_t* action, uint32_t* sampling) {
    const uint8_t* r = lpm_lookup(rules, n_rules, 64, key, prefix);
    if (!r) return 0;
    int result = 1;
    if (r[54] != 2) { *action = r[54]; result++; }
    if (r[56]) {                                                   /* peer CIDR: the opposite address */
        uint8_t pk[16]; uint32_t pp;
        filter_key_of(rec, peer_use_src, eth, pk, &pp);
        if (lpm_lookup(peers, n_peers, 20, pk, pp)) result++; else return 0;
    }
    const uint32_t sample = rd32(r + 20);
    if (sample) { *sampling = sample; result++; }
    const uint8_t proto = rec[36];
    const uint16_t sport = rd16(rec + 32), dport = rd16(rec + 34), flags = rd16(rec + 70);
    if (r[50] == proto || r[50] == 0) {
        if (proto == 6 || proto == 17 || proto == 132) {
            const uint16_t ds = rd16(r + 24), de = rd16(r + 26), d1 = rd16(r + 28), d2 = rd16(r + 30);
            if ((ds != 0 && de == 0) || d1 != 0 || d2 != 0) { if (ds == dport || d1 == dport || d2 == dport) result++; else return 0; }
            else if (ds != 0 && de != 0) { if (ds <= dport && dport <= de) result++; else return 0; }
            const uint16_t ss = rd16(r + 32), se = rd16(r + 34), s1 = rd16(r + 36), s2 = rd16(r + 38);
            if ((ss != 0 && se == 0) || s1 != 0 || s2 != 0) { if (ss == sport || s1 == sport || s2 == sport) result++; else return 0; }
            else if (ss != 0 && se != 0) { if (ss <= sport && sport <= se) result++; else return 0; }
            const uint16_t ps = rd16(r + 40), pe = rd16(r + 42), p1 = rd16(r + 44), p2 = rd16(r + 46);
            if ((ps != 0 && pe == 0) || p1 != 0 || p2 != 0) {
                if (ps == sport || ps == dport || p1 == sport || p1 == dport || p2 == sport || p2 == dport) result++; else return 0;
            } else if (ps != 0 && pe != 0) {
                if ((ps <= sport && sport <= pe) || (ps <= dport && dport <= pe)) result++; else return 0;
            }
            if (proto == 6 && rd16(r + 48) != 0) { if (rd16(r + 48) == flags) result++; else return 0; }
        } else if (proto == 1 || proto == 58) {
            if (r[51] != 0) {
                if (r[51] == rec[37]) result++; else return 0;
                if (r[52] != 0) { if (r[52] == rec[38]) result++; else return 0; }
            }
        }
    } else {
        return 0;
    }
    if (r[53] != 2) { if (r[53] == rec[96]) result++; else return 0; }
    if (r[55]) return 0;                                           /* filter_drops needs drop_reason != 0; flow_monitor passes 0 */
    return result;
}

int oracle_filter_packet(const uint8_t* rules, size_t n_rules, const uint8_t* peers, size_t n_peers, uint8_t* rec, uint64_t counters[3]) {
    const uint16_t eth = rd16(rec + 68);
    uint8_t action = 2;                                            /* is_flow_filtered: *action = MAX_FILTER_ACTIONS */
    uint32_t sampling = rd32(rec + 92);
    uint8_t key[16]; uint32_t prefix;
    filter_key_of(rec, 1, eth, key, &prefix);                      /* source address first; its peer is the destination */
    int result = filter_lookup(rules, n_rules, peers, n_peers, rec, key, prefix, 0, eth, &action, &sampling);
    if (result <= 0) {
        filter_key_of(rec, 0, eth, key, &prefix);
        result = filter_lookup(rules, n_rules, peers, n_peers, rec, key, prefix, 1, eth, &action, &sampling);
    }
    memcpy(rec + 92, &sampling, 4);
    if (result != 0 && action != 2) {                              /* utils.h:185-203 */
        if (action == 1) { counters[1]++; return 1; }
        counters[0]++;
        return 0;
    }
    counters[2]++;                                                 /* utils.h:209-218 */
    return (action == 0 || action == 2) ? 1 : 0;
}

size_t oracle_parse_snaps_filtered(const uint8_t* snaps, size_t n, uint32_t stride, const uint8_t* rules, size_t n_rules,
                                   const uint8_t* peers, size_t n_peers, uint8_t* out, uint32_t* src_of, uint64_t counters[3]) {
    size_t m = 0;
    for (size_t i = 0; i < n; i++) {
        if (!oracle_parse_snap(snaps + i * stride, stride, out + m * 144)) continue;
        if (n_rules && oracle_filter_packet(rules, n_rules, peers, n_peers, out + m * 144, counters)) continue;
        if (src_of) src_of[m] = (uint32_t)i;
        m++;
    }
    return m;
}

/* ------------------------------------------------------------------------------------------------------------------
 * K7: DNS query -> response correlation (bpf/dns_tracker.h:23-37,68-127 + bpf/flows.c:210-213,291-330), sequential.
 * dns_flow_id (bpf/types.h:250-257) restated as a 40-byte image: src_ip 16 | dst_ip 16 | src_port 2 | dst_port 2 | id 2 |
 * protocol 1 | 0.  A chained hash map with exact byte compare stands in for the BPF hash map.
 * ------------------------------------------------------------------------------------------------------------------ */
typedef struct dnsc_node { struct dnsc_node* next; uint8_t key[40]; uint64_t ts; } dnsc_node;
struct oracle_dnscorr { dnsc_node** bucket; size_t n_buckets, len, max_entries; };

oracle_dnscorr* oracle_dnscorr_new(size_t max_entries) {
    oracle_dnscorr* m = (oracle_dnscorr*)calloc(1, sizeof *m);
    m->n_buckets = 1; while (m->n_buckets < 2 * max_entries + 16) m->n_buckets <<= 1;
    m->bucket = (dnsc_node**)calloc(m->n_buckets, sizeof *m->bucket);
    m->max_entries = max_entries;
    return m;
}
void oracle_dnscorr_free(oracle_dnscorr* m) {
    if (!m) return;
    for (size_t b = 0; b < m->n_buckets; b++) for (dnsc_node* x = m->bucket[b]; x;) { dnsc_node* nx = x->next; free(x); x = nx; }
    free(m->bucket); free(m);
}
size_t oracle_dnscorr_pending(const oracle_dnscorr* m) { return m->len; }

static size_t dnsc_bucket(const oracle_dnscorr* m, const uint8_t* key) {
    uint64_t h = 0xCBF29CE484222325ull;
    for (int i = 0; i < 40; i++) { h ^= key[i]; h *= 0x100000001B3ull; }
    return (size_t)(h ^ (h >> 29)) & (m->n_buckets - 1);
}
/* fill_dns_id (dns_tracker.h:23-37) */
static void dnsc_key(const uint8_t* pkt, uint16_t dns_id, int reverse, uint8_t key[40]) {
    memset(key, 0, 40);
    memcpy(key, pkt + (reverse ? 16 : 0), 16);            /* src_ip */
    memcpy(key + 16, pkt + (reverse ? 0 : 16), 16);       /* dst_ip */
    memcpy(key + 32, pkt + (reverse ? 34 : 32), 2);       /* src_port */
    memcpy(key + 34, pkt + (reverse ? 32 : 34), 2);       /* dst_port */
    memcpy(key + 36, &dns_id, 2);
    key[38] = pkt[36];                                    /* transport_protocol */
}

size_t oracle_dnscorr_packets(oracle_dnscorr* m, const uint8_t* pkts, size_t n, uint8_t* out) {
    size_t n_out = 0;
    for (size_t i = 0; i < n; i++) {
        const uint8_t* P = pkts + i * OR_DNSREC_SIZE;
        uint64_t ts; uint16_t id, flags;
        memcpy(&ts, P + 48, 8); memcpy(&id, P + 64, 2); memcpy(&flags, P + 66, 2);
        uint8_t key[40];
        int dns_errno = 0;                                /* what track_dns_packet returns (flows.c:212) */
        uint16_t pkt_dns_id = 0, pkt_dns_flags = 0; uint64_t pkt_latency = 0; int has_name = 0;
        if ((flags & 0x8000u) == 0) {                     /* query: BPF_NOEXIST insert (dns_tracker.h:92-100) */
            dnsc_key(P, id, 0, key);
            const size_t b = dnsc_bucket(m, key);
            dnsc_node* x = m->bucket[b];
            while (x && memcmp(x->key, key, 40)) x = x->next;
            if (x) dns_errno = -17;                       /* -EEXIST */
            else if (m->len >= m->max_entries) dns_errno = -7;   /* -E2BIG: the map is full */
            else {
                x = (dnsc_node*)malloc(sizeof *x);
                memcpy(x->key, key, 40); x->ts = ts; x->next = m->bucket[b]; m->bucket[b] = x; m->len++;
            }
        } else {                                          /* response: lookup of the reversed tuple + delete (:101-110) */
            dnsc_key(P, id, 1, key);
            const size_t b = dnsc_bucket(m, key);
            dnsc_node** px = &m->bucket[b];
            while (*px && memcmp((*px)->key, key, 40)) px = &(*px)->next;
            if (*px) { dnsc_node* x = *px; pkt_latency = ts - x->ts; *px = x->next; free(x); m->len--; }
            else dns_errno = 2;                           /* ENOENT */
            pkt_dns_id = id; pkt_dns_flags = flags; has_name = 1;
        }
        if (pkt_dns_id != 0 || dns_errno != 0) {          /* flows.c:291: one dns_metrics sample for the packet's flow */
            uint8_t* S = out + n_out * OR_DNSREC_SIZE;
            memset(S, 0, OR_DNSREC_SIZE);
            memcpy(S, P, OR_ID_SIZE);
            memcpy(S + 40, &ts, 8); memcpy(S + 48, &ts, 8);              /* start = end = pkt.current_ts */
            memcpy(S + 56, &pkt_latency, 8);
            memcpy(S + 64, &pkt_dns_id, 2); memcpy(S + 66, &pkt_dns_flags, 2);
            memcpy(S + 68, P + 68, 2);                                   /* eth_protocol */
            S[70] = (uint8_t)dns_errno;                                  /* u8 field: -17 -> 239, -7 -> 249 */
            if (has_name) memcpy(S + 71, P + 71, 32);
            n_out++;
        }
    }
    return n_out;
}

size_t oracle_dnscorr_purge(oracle_dnscorr* m, uint64_t now, uint64_t timeout) {
    size_t gone = 0;
    for (size_t b = 0; b < m->n_buckets; b++) {
        dnsc_node** px = &m->bucket[b];
        while (*px) {
            if ((int64_t)(now - (*px)->ts) >= (int64_t)timeout) { dnsc_node* x = *px; *px = x->next; free(x); m->len--; gone++; }
            else px = &(*px)->next;
        }
    }
    return gone;
}

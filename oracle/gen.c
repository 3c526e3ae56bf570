/*
 * gen.c — CPU restatement of the synthetic flow-record stream (spec: DESIGN.md §7 and the product's
 * netobserv_ebpf_agent_b200/csrc/flowgen.h, which is NOT included here).
 *
 * TEST INFRASTRUCTURE ONLY, like everything under oracle/: it lets bench.py --impl reference, the in-bench parity
 * check and the tests produce the workload's records without loading the product library.  tests/test_generator.py
 * pins it to the product generator bit for bit.
 *
 * Record i depends only on (seed, i): key rank from a counter-based 64-bit mixer (uniform, or Zipf through a table
 * of cumulative bucket thresholds in 2^-64 units), per-key constant L2 fields, ts = t0 + i, one collapsed TCP flag
 * (reference bpf/utils.h:24-51), packets = 1 — the single-packet record flow_monitor sends to the ring buffer
 * (reference bpf/flows.c:228-245,268-279).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "oracle.h"

static inline uint64_t g_mix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t g_mix32(uint32_t x) { x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16; return x; }
static inline uint32_t g_bswap(uint32_t x) { return __builtin_bswap32(x); }
static inline uint64_t g_mulhi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }

struct oracle_gen {
    uint64_t seed, n_keys, t0;
    uint32_t dist, varying;
    uint64_t* thr; uint32_t* first; uint32_t* size; uint32_t nb;
};

/* ranks 1..N in octaves [2^o, 2^(o+1)), each split into <= 256 equal sub-ranges; mass = sum k^-s */
static void zipf_build(oracle_gen* g, double s) {
    size_t cap = 64 * 256, nb = 0;
    long double* mass = (long double*)malloc(cap * sizeof(long double));
    g->first = (uint32_t*)malloc(cap * 4); g->size = (uint32_t*)malloc(cap * 4);
    for (uint64_t lo = 1; lo <= g->n_keys; lo <<= 1) {
        const uint64_t hi = (lo << 1) < g->n_keys + 1 ? (lo << 1) : g->n_keys + 1;
        const uint64_t span = hi - lo, parts = span < 256 ? span : 256;
        for (uint64_t pi = 0; pi < parts; pi++) {
            const uint64_t a = lo + span * pi / parts, b = lo + span * (pi + 1) / parts;
            long double m = 0;
            for (uint64_t k = a; k < b; k++) m += pow((double)k, -s);
            g->first[nb] = (uint32_t)(a - 1); g->size[nb] = (uint32_t)(b - a); mass[nb++] = m;
        }
    }
    long double total = 0, run = 0;
    for (size_t i = 0; i < nb; i++) total += mass[i];
    g->thr = (uint64_t*)malloc(nb * 8);
    for (size_t i = 0; i < nb; i++) {
        run += mass[i];
        const long double f = run / total * 18446744073709551616.0L;
        g->thr[i] = f >= 18446744073709551615.0L ? ~0ull : (uint64_t)f;
    }
    g->thr[nb - 1] = ~0ull;
    g->nb = (uint32_t)nb;
    free(mass);
}

oracle_gen* oracle_gen_new(uint64_t seed, uint64_t n_keys, uint32_t dist, uint32_t zipf_s_milli, uint64_t t0_ns, uint32_t varying_desc) {
    oracle_gen* g = (oracle_gen*)calloc(1, sizeof(*g));
    g->seed = seed; g->n_keys = n_keys; g->dist = dist; g->t0 = t0_ns; g->varying = varying_desc;
    if (dist == 1) zipf_build(g, zipf_s_milli / 1000.0);
    return g;
}
void oracle_gen_free(oracle_gen* g) { if (g) { free(g->thr); free(g->first); free(g->size); free(g); } }

static uint64_t pick_rank(const oracle_gen* g, uint64_t r) {
    if (g->dist == 0 || g->nb == 0) return g_mulhi(r, g->n_keys);
    uint32_t lo = 0, hi = g->nb - 1;                      /* first bucket whose threshold is >= r */
    while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (g->thr[mid] >= r) hi = mid; else lo = mid + 1; }
    return (uint64_t)g->first[lo] + g_mulhi(g_mix(0x5851F42D4C957F2Dull, r), (uint64_t)g->size[lo]);
}

static void key_words(uint64_t seed, uint64_t kid, uint32_t w[10]) {
    const uint64_t g1 = g_mix(seed ^ 0xA5A5A5A5DEADBEEFull, kid), g2 = g_mix(seed ^ 0x0123456789ABCDEFull, kid);
    const int v6 = (g1 & 0xFF) < 26, tcp = ((g1 >> 8) & 0xFF) < 205;
    memset(w, 0, 40);
    if (v6) {
        w[0] = 0xB80D0120u; w[1] = (uint32_t)(g2 >> 32); w[2] = (uint32_t)g2;
        w[4] = 0xB80D0120u; w[5] = (uint32_t)(g1 >> 40) | 0x01000000u; w[6] = (uint32_t)(g2 >> 16);
    } else { w[2] = 0xFFFF0000u; w[6] = 0xFFFF0000u; }
    w[3] = g_bswap(g_mix32((uint32_t)kid));
    w[7] = g_bswap(0x0A000000u | ((uint32_t)(g2 >> 8) & 0x00FFFFFFu));
    static const uint32_t ports[8] = {80, 443, 8080, 53, 6443, 5432, 9090, 22};
    w[8] = (1024u + (uint32_t)((g1 >> 16) % 64000u)) | (ports[(g1 >> 48) & 7] << 16);
    w[9] = tcp ? 6u : 17u;
}

void oracle_gen_key(const oracle_gen* g, uint64_t rank, uint8_t* key40) { uint32_t w[10]; key_words(g->seed, rank, w); memcpy(key40, w, 40); }

static void record_words(const oracle_gen* g, uint64_t idx, uint32_t w[36]) {
    const uint64_t r0 = g_mix(g->seed, 2 * idx), r1 = g_mix(g->seed, 2 * idx + 1);
    const uint64_t kid = pick_rank(g, r0);
    key_words(g->seed, kid, w);
    const int tcp = (w[9] & 0xFF) == 6, v6 = w[2] != 0xFFFF0000u || w[0] != 0;
    const uint64_t ts = g->t0 + idx;
    const uint32_t len = 64u + (uint32_t)(r1 % 1437u);
    const uint32_t fsel = (uint32_t)(r1 >> 16) % 11u;
    const uint32_t flags = !tcp ? 0u : (fsel < 8 ? (1u << fsel) : (0x100u << (fsel - 8)));
    const uint64_t g3 = g_mix(g->seed ^ 0x7777777711111111ull, kid), g4 = g_mix(g->seed ^ 0x3333333399999999ull, kid);
    uint32_t eth = v6 ? 0x86DDu : 0x0800u;
    uint32_t smac_lo = ((uint32_t)g3 & 0xFFFFFFFEu) | 0x02u, smac_hi = (uint32_t)(g3 >> 32) & 0xFFFFu;
    uint32_t dmac_lo = (uint32_t)g4 | 0x02u, dmac_hi = (uint32_t)(g4 >> 32) & 0xFFFFu;
    uint32_t ifx = 1u + (uint32_t)((g3 >> 48) & 7), dir = (uint32_t)(g4 >> 48) & 1;
    uint32_t dscp = ((g4 >> 50) & 3) == 0 ? 0u : (uint32_t)((g4 >> 52) & 0x3F);
    uint32_t samp = ((g4 >> 58) & 1) ? 50u : 0u;
    uint32_t err = 0, nobs = 0, odir = 0, oi[3] = {0, 0, 0}, ssl = 0, cipher = 0, types = 0, misc = 0;
    if (g->varying) {
        const uint64_t q0 = g_mix(g->seed ^ 0xD1B54A32D192ED03ull, idx), q1 = g_mix(g->seed ^ 0x8CB92BA72F3D8DD7ull, idx);
        eth = (q0 & 3) == 0 ? 0u : (((q0 >> 2) & 1) ? 0x86DDu : 0x0800u);
        if (((q0 >> 3) & 3) == 0) { smac_lo = 0; smac_hi = 0; } else smac_lo ^= (uint32_t)(q0 >> 8) & 0xFF00u;
        if (((q0 >> 5) & 3) == 0) { dmac_lo = 0; dmac_hi = 0; } else dmac_hi ^= (uint32_t)(q0 >> 20) & 0xFFu;
        ifx = (uint32_t)(q0 >> 32) & 0xF; dir = (uint32_t)(q0 >> 36) & 1;
        dscp = ((q0 >> 37) & 1) ? 0u : (uint32_t)(q0 >> 40) & 0x3F;
        samp = ((q0 >> 46) & 1) ? 0u : (uint32_t)(q0 >> 48) & 0xFF;
        err = ((q1 & 7) == 0) ? 7u : 0u;
        nobs = (uint32_t)(q1 >> 3) & 3; odir = (uint32_t)(q1 >> 8) & 0x01010101u;
        oi[0] = nobs > 0 ? 1u + ((uint32_t)(q1 >> 16) & 7) : 0; oi[1] = nobs > 1 ? 9u + ((uint32_t)(q1 >> 20) & 7) : 0;
        oi[2] = nobs > 2 ? 17u + ((uint32_t)(q1 >> 24) & 7) : 0;
        ssl = ((q1 >> 28) & 3) == 0 ? 0x0303u : 0u; cipher = ssl ? 0x1301u : 0u;
        types = ssl ? (uint32_t)(q1 >> 32) & 0x3F : 0u; misc = (uint32_t)(q1 >> 40) & 1;
    }
    w[10] = (uint32_t)ts; w[11] = (uint32_t)(ts >> 32); w[12] = w[10]; w[13] = w[11];
    w[14] = len; w[15] = 0; w[16] = 1; w[17] = eth | (flags << 16);
    w[18] = smac_lo; w[19] = smac_hi | ((dmac_lo & 0xFFFFu) << 16); w[20] = (dmac_lo >> 16) | (dmac_hi << 16);
    w[21] = ifx; w[22] = 0; w[23] = samp;
    w[24] = dir | (err << 8) | (dscp << 16) | (nobs << 24);
    w[25] = odir; w[26] = 0; w[27] = oi[0]; w[28] = oi[1]; w[29] = oi[2]; w[30] = 0; w[31] = 0; w[32] = 0;
    w[33] = ssl | (cipher << 16); w[34] = (types << 16) | (misc << 24); w[35] = 0;
}

void oracle_gen_records(const oracle_gen* g, uint64_t first_index, size_t n, uint8_t* out, int n_threads) {
    if (n_threads < 1) n_threads = 1;
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        uint32_t w[36];
        record_words(g, first_index + i, w);
        memcpy(out + i * OR_REC_SIZE, w, OR_REC_SIZE);
    }
}

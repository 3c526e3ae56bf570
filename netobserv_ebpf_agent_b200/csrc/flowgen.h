// flowgen.h — counter-based synthetic flow-record stream, shared by the
// host generator and the CUDA generator so both produce bit-identical bytes: record i
// depends only on (seed, i) and on an integer Zipf threshold table built once on the host.
//
// The single-packet record it emits is what the reference's datapath sends to the
// direct_flows ring buffer (bpf/flows.c:228-245,268-279): packets=1, bytes=len,
// start=end=ts, one collapsed TCP flag (bpf/utils.h:24-51), per-key-constant L2/ifindex
// fields (so that every merge in pkg/model/flow_content.go:28-61 is order-independent),
// or — with varying_desc=1 — per-record random descriptors that exercise the
// order-dependent merge rules.
#pragma once
#include <stdint.h>

#ifndef FA_HD
#if defined(__CUDACC__)
#define FA_HD __host__ __device__ __forceinline__
#else
#define FA_HD inline
#endif
#endif

namespace fa {

struct GenDeviceParams {
    uint64_t seed;
    uint64_t n_keys;
    uint64_t t0_ns;
    uint32_t dist;           // 0 uniform, 1 zipf
    uint32_t varying_desc;
    // Zipf sampler: n_buckets contiguous rank ranges; thresholds[b] = cumulative probability
    // of buckets 0..b in 2^-64 units (last = 2^64-1); ranks inside a bucket are drawn uniformly.
    const uint64_t* thresholds;
    const uint32_t* bucket_first;   // first rank (0-based) of bucket b
    const uint32_t* bucket_size;
    uint32_t n_buckets;
    uint32_t reserved;
};

FA_HD uint64_t gen_splitmix(uint64_t seed, uint64_t ctr) {
    uint64_t z = seed + (ctr + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
// invertible 32-bit mixer: distinct key ids -> distinct IPv4 host parts
FA_HD uint32_t gen_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7FEB352Du; x ^= x >> 15; x *= 0x846CA68Bu; x ^= x >> 16;
    return x;
}

FA_HD uint32_t gen_bswap32(uint32_t x) {
    return (x >> 24) | ((x >> 8) & 0xFF00u) | ((x << 8) & 0xFF0000u) | (x << 24);
}

FA_HD uint64_t gen_pick_rank(const GenDeviceParams& p, uint64_t r) {
    if (p.dist == 0 || p.n_buckets == 0) {
        // uniform over n_keys via 64x64->128 multiply-high
#if defined(__CUDA_ARCH__)
        return __umul64hi(r, p.n_keys);
#else
        return (uint64_t)(((unsigned __int128)r * p.n_keys) >> 64);
#endif
    }
    uint32_t lo = 0, hi = p.n_buckets - 1;          // first bucket with thresholds[b] >= r
    while (lo < hi) {
        uint32_t mid = (lo + hi) >> 1;
        if (p.thresholds[mid] >= r) hi = mid; else lo = mid + 1;
    }
    const uint32_t sz = p.bucket_size[lo];
    uint64_t r2 = gen_splitmix(0x5851F42D4C957F2Dull, r);
#if defined(__CUDA_ARCH__)
    uint64_t off = __umul64hi(r2, (uint64_t)sz);
#else
    uint64_t off = (uint64_t)(((unsigned __int128)r2 * sz) >> 64);
#endif
    return (uint64_t)p.bucket_first[lo] + off;
}

// Fill the 40-byte key of key id `kid` into 10 LE u32 words.
FA_HD void gen_key_words(uint64_t seed, uint64_t kid, uint32_t w[10]) {
    const uint64_t g1 = gen_splitmix(seed ^ 0xA5A5A5A5DEADBEEFull, kid);
    const uint64_t g2 = gen_splitmix(seed ^ 0x0123456789ABCDEFull, kid);
    const bool v6 = (g1 & 0xFFu) < 26u;                   // ~10 % IPv6
    const bool tcp = ((g1 >> 8) & 0xFFu) < 205u;          // ~80 % TCP
    const uint32_t host = gen_mix32((uint32_t)kid);       // bijective => keys are distinct
    uint32_t src[4], dst[4];
    if (v6) {
        src[0] = 0xB80D0120u;                             // 2001:0db8::/32 in memory order 20 01 0d b8
        src[1] = (uint32_t)(g2 >> 32); src[2] = (uint32_t)g2;
        dst[0] = 0xB80D0120u; dst[1] = (uint32_t)(g1 >> 40) | 0x01000000u; dst[2] = (uint32_t)(g2 >> 16);
    } else {
        src[0] = 0; src[1] = 0; src[2] = 0xFFFF0000u;     // ::ffff:a.b.c.d  (bytes 10,11 = ff ff)
        dst[0] = 0; dst[1] = 0; dst[2] = 0xFFFF0000u;
    }
    src[3] = gen_bswap32(host);                      // big-endian host part in bytes 12..15
    dst[3] = gen_bswap32(0x0A000000u | ((uint32_t)(g2 >> 8) & 0x00FFFFFFu));
    w[0] = src[0]; w[1] = src[1]; w[2] = src[2]; w[3] = src[3];
    w[4] = dst[0]; w[5] = dst[1]; w[6] = dst[2]; w[7] = dst[3];
    const uint32_t sport = 1024u + (uint32_t)((g1 >> 16) % 64000u);
    const uint32_t dsel = (uint32_t)(g1 >> 48) & 7u;
    const uint32_t dport = dsel == 0 ? 80u : dsel == 1 ? 443u : dsel == 2 ? 8080u : dsel == 3 ? 53u
                         : dsel == 4 ? 6443u : dsel == 5 ? 5432u : dsel == 6 ? 9090u : 22u;
    w[8] = sport | (dport << 16);
    w[9] = tcp ? 6u : 17u;                                // proto, icmp_type=0, icmp_code=0, pad=0
}

// Write record `idx` of the stream as 36 LE u32 words (144 bytes).
FA_HD void gen_record_words(const GenDeviceParams& p, uint64_t idx, uint32_t w[36]) {
    const uint64_t r0 = gen_splitmix(p.seed, 2 * idx);
    const uint64_t r1 = gen_splitmix(p.seed, 2 * idx + 1);
    const uint64_t kid = gen_pick_rank(p, r0);
    gen_key_words(p.seed, kid, w);
    const bool tcp = (w[9] & 0xFFu) == 6u;
    const bool v6 = w[2] != 0xFFFF0000u || w[0] != 0u;
    const uint64_t ts = p.t0_ns + idx;
    const uint32_t len = 64u + (uint32_t)(r1 % 1437u);     // [64, 1500]
    // collapsed TCP flags: exactly one of the 11 values the datapath can emit (bpf/utils.h:24-51)
    const uint32_t fsel = (uint32_t)(r1 >> 16) % 11u;
    const uint32_t flags = !tcp ? 0u : (fsel < 8u ? (1u << fsel) : (0x100u << (fsel - 8u)));
    const uint64_t g3 = gen_splitmix(p.seed ^ 0x7777777711111111ull, kid);
    const uint64_t g4 = gen_splitmix(p.seed ^ 0x3333333399999999ull, kid);
    uint32_t eth = v6 ? 0x86DDu : 0x0800u;
    uint32_t smac_lo = (uint32_t)g3 & 0xFFFFFFFEu | 0x02u, smac_hi = (uint32_t)(g3 >> 32) & 0xFFFFu;
    uint32_t dmac_lo = (uint32_t)g4 | 0x02u, dmac_hi = (uint32_t)(g4 >> 32) & 0xFFFFu;   // dmac as 6 bytes: lo16 first
    uint32_t if_index = 1u + (uint32_t)((g3 >> 48) & 7u);
    uint32_t direction = (uint32_t)(g4 >> 48) & 1u;
    uint32_t dscp = ((g4 >> 50) & 3u) == 0 ? 0u : (uint32_t)((g4 >> 52) & 0x3Fu);
    uint32_t sampling = ((g4 >> 58) & 1u) ? 50u : 0u;
    uint32_t lock = 0, errno_ = 0, nb_obs = 0, obsdir_lo = 0, obsdir_hi = 0;
    uint32_t obs_intf[6] = {0, 0, 0, 0, 0, 0};
    uint32_t ssl = 0, cipher = 0, keyshare = 0, tls_types = 0, misc = 0;
    if (p.varying_desc) {
        // per-record random descriptors: zero / non-zero mixes for every order-dependent rule
        const uint64_t q0 = gen_splitmix(p.seed ^ 0xD1B54A32D192ED03ull, idx);
        const uint64_t q1 = gen_splitmix(p.seed ^ 0x8CB92BA72F3D8DD7ull, idx);
        eth = (q0 & 3u) == 0 ? 0u : (((q0 >> 2) & 1u) ? 0x86DDu : 0x0800u);
        if (((q0 >> 3) & 3u) == 0) { smac_lo = 0; smac_hi = 0; } else { smac_lo ^= (uint32_t)(q0 >> 8) & 0xFF00u; }
        if (((q0 >> 5) & 3u) == 0) { dmac_lo = 0; dmac_hi = 0; } else { dmac_hi ^= (uint32_t)(q0 >> 20) & 0xFFu; }
        if_index = (uint32_t)(q0 >> 32) & 0xFu;
        direction = (uint32_t)(q0 >> 36) & 1u;
        dscp = ((q0 >> 37) & 1u) ? 0u : (uint32_t)(q0 >> 40) & 0x3Fu;
        sampling = ((q0 >> 46) & 1u) ? 0u : (uint32_t)(q0 >> 48) & 0xFFu;
        errno_ = ((q1 & 7u) == 0) ? 7u : 0u;
        nb_obs = (uint32_t)(q1 >> 3) & 3u;
        obsdir_lo = (uint32_t)(q1 >> 8) & 0x01010101u; obsdir_hi = 0;
        obs_intf[0] = nb_obs > 0 ? 1u + ((uint32_t)(q1 >> 16) & 7u) : 0u;
        obs_intf[1] = nb_obs > 1 ? 9u + ((uint32_t)(q1 >> 20) & 7u) : 0u;
        obs_intf[2] = nb_obs > 2 ? 17u + ((uint32_t)(q1 >> 24) & 7u) : 0u;
        ssl = ((q1 >> 28) & 3u) == 0 ? 0x0303u : 0u;
        cipher = ssl ? 0x1301u : 0u;
        tls_types = ssl ? (uint32_t)(q1 >> 32) & 0x3Fu : 0u;
        misc = (uint32_t)(q1 >> 40) & 1u;
        lock = 0;
    }
    // metrics (record words 10..35)
    w[10] = (uint32_t)ts; w[11] = (uint32_t)(ts >> 32);          // start
    w[12] = (uint32_t)ts; w[13] = (uint32_t)(ts >> 32);          // end
    w[14] = len; w[15] = 0;                                      // bytes
    w[16] = 1u;                                                  // packets
    w[17] = eth | (flags << 16);
    w[18] = smac_lo;                                             // src_mac[0..4)
    w[19] = smac_hi | ((dmac_lo & 0xFFFFu) << 16);               // src_mac[4..6) | dst_mac[0..2)
    w[20] = (dmac_lo >> 16) | (dmac_hi << 16);                   // dst_mac[2..6)
    w[21] = if_index; w[22] = lock; w[23] = sampling;
    w[24] = direction | (errno_ << 8) | (dscp << 16) | (nb_obs << 24);
    w[25] = obsdir_lo; w[26] = obsdir_hi & 0xFFFFu;              // observed_direction[6] + 2 pad bytes
    w[27] = obs_intf[0]; w[28] = obs_intf[1]; w[29] = obs_intf[2];
    w[30] = obs_intf[3]; w[31] = obs_intf[4]; w[32] = obs_intf[5];
    w[33] = ssl | (cipher << 16);
    w[34] = keyshare | (tls_types << 16) | (misc << 24);
    w[35] = 0;
}

}  // namespace fa

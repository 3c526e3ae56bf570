// kernels.cuh — launch wrappers implemented in the .cu files (internal to libflowagg.so).
#pragma once
#include "common.cuh"

namespace fa {

struct SketchParams {
    unsigned long long* cms;   // depth x 2^log2w u64, or nullptr
    uint32_t*           hll;   // 2^p u32 registers (value = rho), or nullptr
    uint32_t log2w, depth, p;
    uint64_t seed;
};

// scratch entry of the ordered re-fold (one per flow flagged TAG_DIRTY in a launch)
// (an open-addressed set keyed by table slot; capacity = power of two >= 2 x max_batch)
struct FixupScratch {
    unsigned long long key;  // table slot + 1, 0 = empty
    uint32_t nfirst;  // ~(min record index of the flow in this launch)       all-zero == initial state
    uint32_t eth;     // 1 + max index with eth_protocol != 0
    uint32_t dscp;    // 1 + max index with dscp != 0
    uint32_t samp;    // 1 + max index with sampling != 0
    uint32_t nsmac;   // ~(min index with src_mac != 0), 0 = none
    uint32_t ndmac;   // ~(min index with dst_mac != 0), 0 = none
};

struct AggLaunch {
    const uint4* recs;       // n x 144 B, 16-byte aligned, device memory
    uint32_t     n;
    Table        table;
    uint64_t     epoch;
    Counters*    ctr;
    uint32_t*    spill_idx;  // indices of records that found the table physically full
    SketchParams sk;
    FixupScratch* scratch;   // scratch_slots entries, all-zero between launches
    uint32_t     scratch_slots;
    int          sm_count;
    unsigned long long* prof; // 8 phase cycle counters (FA_PHASE_PROFILE=1) or nullptr
    uint32_t     opt;        // switches: 2 = no hot-flow cache (diagnostics), 16 = the record count is on the device (ctr->launch_n)
};

// K1: fold a batch of flow records into the table (ACCOUNTER semantics) and, when any
// flow saw records with differing order-dependent fields, re-fold those flows in stream order.
// Returns the number of kernels launched.
int launch_aggregate(const AggLaunch& a, cudaStream_t st);

// K2: lookup-and-delete every live flow of `table` into out_recs (device pointer, cap records).
// The number found is left in ctr->evict_out (can exceed cap; only cap are written).
// slot_of_out (optional): receives the table slot of every emitted flow, for the feature pass.
// drain: emit + reset only the flows that received records since the last drain; nothing is removed.
int launch_evict(const Table& table, uint4* out_recs, uint32_t* slot_of_out,
                 unsigned long long cap, Counters* ctr, int sm_count, cudaStream_t st, bool drain = false);

// Overflow pre-pass (ACCOUNTER "full" cut, reference pkg/flow/account.go:85-94): finds the index of the
// first record whose key is new when the cache already holds max_entries flows. Result in *cut_out
// (device; == n when the whole batch fits).  idx_set: scratch of set_slots u32 (power of two >= 2n),
// bitmap: scratch of (n+31)/32 u32.
int launch_full_cut(const uint4* recs, uint32_t n, const Table& table, unsigned long long live,
                    unsigned long long max_entries, uint32_t* idx_set, uint32_t set_slots, uint32_t* bitmap,
                    uint32_t* cut_out, int sm_count, cudaStream_t st);

// K6 feature folds (kind 0 = additional_metrics 72-B records, 1 = dns_metrics 104-B records, 2 = pkt_drop_metrics 72-B
// records); slot_of: n u32 scratch
int launch_feature_fold(int kind, const uint8_t* recs, uint32_t n, const Table& t, uint64_t epoch, uint64_t seq0,
                        uint32_t* slot_of, Counters* ctr, int sm_count, cudaStream_t st);
int launch_evict_features(const Table& t, const uint32_t* slot_of_out, unsigned long long n_out, uint8_t* out_recs,
                          uint8_t* out_dns, uint8_t* out_add, uint8_t* out_drop, unsigned long long* out_rtt_min,
                          uint8_t* out_present, int sm_count, cudaStream_t st);

// sketches
int launch_cms_query(const SketchParams& sk, const uint4* keys, uint32_t n, unsigned long long* est, cudaStream_t st);
int launch_hll_pack(const SketchParams& sk, uint8_t* out_regs, cudaStream_t st);

// K8: evicted flows -> pbflow.Record wire bytes (pbflow.cu)
struct PbIface { uint32_t if_index; uint8_t mac[6]; uint8_t name_len, udn_len; char name[16]; char udn[64]; };   // == fa_iface_name
struct PbParams {
    uint64_t now_unix_ns, mono_now_ns;
    uint8_t  agent_ip[16];
    uint32_t agent_is_v4;             // 1: agent_ip[12..16) is an IPv4 address
    uint32_t wrap;                    // 1: every record is prefixed with the Records.entries tag + length
    const PbIface* ifaces; uint32_t n_ifaces;
};
struct PbInputs { const uint8_t* recs; const uint8_t* dns; const uint8_t* add; const uint8_t* drop; const uint8_t* present; };
// sizes: n u32 (message bodies), offsets: n + 1 u64, block_sums: ceil(n / 1024) u64 — device scratch
int launch_pb_sizes(const PbInputs& in, uint32_t n, const PbParams& P, uint32_t* sizes, unsigned long long* offsets,
                    unsigned long long* block_sums, int sm_count, cudaStream_t st);
int launch_pb_write(const PbInputs& in, uint32_t n, const PbParams& P, const unsigned long long* offsets, const uint32_t* sizes,
                    uint8_t* out, uint8_t* keys_out, cudaStream_t st);

// 64-byte packet events -> single-packet 144-byte records (misc_kernels.cu)
int launch_expand_events(const uint4* events, uint32_t n, uint4* recs_out, cudaStream_t st);
// (f4) packet snapshots -> records of the packets flow_monitor would submit (and, with a filter, keep), stable order;
// *n_out (device) = how many; filter_ctr[3] (device) += accept / reject / no-match
constexpr int kSnapMaxCtas = 1024;                 // cta_count has this many entries
constexpr int kMaxFilterEntries = 16;              // MAX_FILTER_ENTRIES (bpf/types.h:66)
struct FilterRuleDev {                             // fa_filter_rule with the address as big-endian words
    uint32_t ipw[4]; uint32_t prefix, sample;
    uint16_t dps, dpe, dp1, dp2, sps, spe, sp1, sp2, ps, pe, p1, p2, tcp_flags;
    uint8_t proto, icmp_type, icmp_code, direction, action, filter_drops, peer;
};
struct FilterCidrDev { uint32_t ipw[4]; uint32_t prefix; };
struct FilterSet { uint32_t n_rules, n_peers; FilterRuleDev rules[kMaxFilterEntries]; FilterCidrDev peers[kMaxFilterEntries]; };
int launch_parse_snaps(const uint8_t* snaps, uint32_t n, uint32_t stride, const FilterSet* filter, uint32_t* cta_count, uint8_t* verdict,
                       uint4* out_recs, uint32_t* src_of, unsigned long long* n_out, unsigned long long* filter_ctr, int sm_count,
                       cudaStream_t st);

// K7 DNS query -> response correlation (dnscorr.cu): the dns_flows map of bpf/maps_definition.h:81-89 on the device
struct __align__(64) DnsEntry {                    // one key of dns_flows, 64 B
    unsigned long long key[5];                     // dns_flow_id: src_ip | dst_ip | src_port, dst_port << 16, id << 32, protocol << 48
    unsigned long long ts;                         // the map's value while `present`
    uint32_t tag;                                  // 0 empty, 1 being written, 2 published
    uint32_t present;                              // the map holds the key (a deleted key keeps its entry until the next rebuild)
    uint32_t next[2];                              // per batch: smallest unprocessed packet index of the key, by round parity
};
enum { DNSC_REMAINING = 0, DNSC_CREATED, DNSC_EMITTED, DNSC_PRESENT, DNSC_FULL, DNSC_PURGED, DNSC_N = 8 };
struct DnsCorr { DnsEntry* tab; uint64_t mask; unsigned long long* ctr; unsigned long long max_entries; };
constexpr int kDnsRounds = 6;                      // round kernels per batch (later ones exit at once); the tail takes the rest
// pkts: n x 104 B in stream order; state: n u32; samples: n x 104 B scratch; warp_count: ceil(n / 32) u32; out: n x 104 B,
// receives ctr[DNSC_EMITTED] samples in stream order
int launch_dns_correlate(const uint8_t* pkts, uint32_t n, const DnsCorr& d, uint32_t* state, uint8_t* samples, uint32_t* warp_count,
                         uint8_t* out, int sm_count, cudaStream_t st);
int launch_dns_purge(const DnsCorr& d, uint64_t now, uint64_t timeout, int sm_count, cudaStream_t st);
int launch_dns_rebuild(const DnsCorr& from, const DnsCorr& to, int sm_count, cudaStream_t st);

// generator
struct GenDeviceParams;
int launch_generate(const GenDeviceParams& g, uint64_t first_index, uint32_t n, uint4* dst, cudaStream_t st);

// routing fused with the exchange: destinations are peer-mapped receive buffers + their record counters
struct PeerTargets { uint4* buf[16]; unsigned long long* count[16]; };
int launch_route_peer(const uint4* recs, const unsigned long long* n_dev, uint32_t max_n, uint32_t n_shards, uint32_t self_shard,
                      const PeerTargets& pt, unsigned long long cap, unsigned long long* overflow, int sm_count, cudaStream_t st);

// routing (K3)
int launch_route(const uint4* recs, uint32_t n, uint32_t n_shards, uint4* out, unsigned long long* counts_dev,
                 uint32_t* tmp_owner, int sm_count, cudaStream_t st);

}  // namespace fa

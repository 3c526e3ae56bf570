/*
 * oracle.h — CPU restatement of the netobserv agent's flow-aggregation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under netobserv_ebpf_agent_b200/ may
 * include, link or call this; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, and only as the
 * checker or the timed CPU baseline.
 *
 * Parity status: PINNED for ACCOUNTER mode, ReadFrom decoding, AccumulateDNS,
 * AccumulateAdditional and NewRecord time conversion against the reference's
 * own known-answer tests (tests/test_oracle_goldens.py transcribes
 * pkg/flow/account_test.go:47-217, pkg/model/record_test.go:19-102,193-224,
 * 323-347, pkg/model/flow_content_test.go:11-53,184-246,338-380).
 * SOURCE-PINNED, TEST-UNPINNED (the reference has no unit test): the
 * order-dependent fields of AccumulateBase (eth_protocol/dscp/sampling/MACs)
 * and all of KERNEL_MAP mode (bpf/flows.c:76-143,222-288).
 * PARITY UNPINNED: count-min sketch and HyperLogLog — they do not exist in the
 * reference; the oracle restates this repo's own spec
 * and tests additionally validate them against exact counts within (eps,delta).
 *
 * The Go toolchain is absent from the build image, so the reference itself
 * cannot be compiled: there is no oracle/_ref.
 *
 * Everything works on raw little-endian byte buffers with the offsets below;
 * the product header include/flowagg.h is deliberately NOT included.
 */
#ifndef FLOWAGG_ORACLE_H
#define FLOWAGG_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* sizes (reference bpf/types.h:94-126,131-140,174-181,191-215) */
enum {
    OR_ID_SIZE = 40, OR_MET_SIZE = 104, OR_REC_SIZE = 144,
    OR_DNS_SIZE = 64, OR_ADD_SIZE = 32,
    OR_DNSREC_SIZE = 104, OR_ADDREC_SIZE = 72,
    OR_DROP_SIZE = 32, OR_DROPREC_SIZE = 72
};

/* --- pkg/model/record.go:227-231 ReadFrom: binary.Read of 144 bytes; blank (padding)
 * struct fields are skipped by encoding/binary, i.e. come out zero. */
void oracle_read_from(const uint8_t* wire144, uint8_t* rec144);

/* --- pkg/model/flow_content.go:28-61 */
void oracle_accumulate_base(uint8_t* p_metrics104, const uint8_t* o_metrics104);

/* BpfFlowContent restated: base metrics + optional feature blocks. */
typedef struct oracle_content {
    uint8_t metrics[OR_MET_SIZE];
    uint8_t dns[OR_DNS_SIZE];
    uint8_t additional[OR_ADD_SIZE];
    uint8_t has_dns;
    uint8_t has_additional;
} oracle_content;

/* --- pkg/model/flow_content.go:63-96 */
void oracle_accumulate_dns(oracle_content* p, const uint8_t* dns64);
/* --- pkg/model/flow_content.go:154-177 */
void oracle_accumulate_additional(oracle_content* p, const uint8_t* add32);

/* --- pkg/model/flow_content.go:98-117 AccumulateDrops on (base metrics, pkt_drop_metrics block, presence flag);
 * pinned by flow_content_test.go:55-104. */
void oracle_accumulate_drops(uint8_t* p_metrics104, uint8_t* p_drops32, uint8_t* has_drops, const uint8_t* o_drops32);

/* --- pkg/model/record.go:90-97: wall = now - (monoNow - mono), u64 arithmetic. */
void oracle_new_record_times(uint64_t now_unix_ns, uint64_t mono_now_ns,
                             uint64_t start_mono, uint64_t end_mono,
                             uint64_t* time_flow_start_ns, uint64_t* time_flow_end_ns);

/* --- pkg/flow/account.go:58-124: the Accounter as a sequential state machine. */
typedef struct oracle_accounter oracle_accounter;
oracle_accounter* oracle_accounter_new(size_t max_entries);
void   oracle_accounter_free(oracle_accounter* a);
/* Feed n wire records (each goes through ReadFrom, tracer_ringbuf.go:112-134).
 * Evictions of reason "full" (account.go:85-94) are queued as generations. */
void   oracle_accounter_account(oracle_accounter* a, const uint8_t* wire, size_t n);
size_t oracle_accounter_len(const oracle_accounter* a);
/* Timer / closing eviction (account.go:63-80): moves all entries out. Returns count
 * (writes at most cap records, in first-insertion order). */
size_t oracle_accounter_evict(oracle_accounter* a, uint8_t* out_recs, size_t cap);
/* "full" generations queued by _account, oldest first. */
size_t oracle_accounter_pending(const oracle_accounter* a);
size_t oracle_accounter_next_generation_len(const oracle_accounter* a);
size_t oracle_accounter_pop_generation(oracle_accounter* a, uint8_t* out_recs, size_t cap);

/* Multi-thread variant used only as the `--impl reference` CPU baseline: records are
 * partitioned by key hash over n_threads private Accounters (no max_entries cut). */
size_t oracle_accounter_sharded_run(const uint8_t* wire, size_t n, int n_threads,
                                    uint8_t* out_recs, size_t cap);

/* Persistent variant for bench.py --impl reference: the T private Accounters live across calls (steady state, like
 * the GPU arm's flow table), every call folds one batch. */
typedef struct oracle_sharded oracle_sharded;
oracle_sharded* oracle_sharded_new(int n_threads);
void   oracle_sharded_free(oracle_sharded* s);
void   oracle_sharded_account(oracle_sharded* s, const uint8_t* wire, size_t n);
size_t oracle_sharded_len(const oracle_sharded* s);
size_t oracle_sharded_evict(oracle_sharded* s, uint8_t* out_recs, size_t cap);   /* lookup-and-delete of every shard */

/* --- synthetic stream (gen.c): CPU restatement of the workload generator, so that the CPU arms and the in-bench
 * parity check never load the product library */
typedef struct oracle_gen oracle_gen;
oracle_gen* oracle_gen_new(uint64_t seed, uint64_t n_keys, uint32_t dist, uint32_t zipf_s_milli, uint64_t t0_ns,
                           uint32_t varying_desc);
void   oracle_gen_free(oracle_gen* g);
void   oracle_gen_records(const oracle_gen* g, uint64_t first_index, size_t n, uint8_t* out, int n_threads);
void   oracle_gen_key(const oracle_gen* g, uint64_t rank, uint8_t* key40);

/* --- Map path with feature folds: LookupAndDeleteMap's merged view
 * (pkg/tracer/tracer.go:1063-1187) under the record-by-record fold contract
 * (SURVEY.md §8 a12'). */
typedef struct oracle_flowmap oracle_flowmap;
oracle_flowmap* oracle_flowmap_new(void);
void   oracle_flowmap_free(oracle_flowmap* m);
void   oracle_flowmap_account(oracle_flowmap* m, const uint8_t* wire, size_t n);          /* ACCOUNTER fold */
void   oracle_flowmap_fold_dns(oracle_flowmap* m, const uint8_t* dnsrecs, size_t n);       /* n x 104 B */
void   oracle_flowmap_fold_additional(oracle_flowmap* m, const uint8_t* addrecs, size_t n);/* n x 72 B */
void   oracle_flowmap_fold_drops(oracle_flowmap* m, const uint8_t* droprecs, size_t n);      /* n x 72 B */
size_t oracle_flowmap_len(const oracle_flowmap* m);
/* like oracle_flowmap_evict, plus the packet-drop blocks (32 B, present bit 2) and the RTT minimum (extension:
 * smallest non-zero flow_rtt of the flow's samples; PARITY UNPINNED — the reference keeps only the maximum) */
size_t oracle_flowmap_evict_ex(oracle_flowmap* m, uint8_t* out_recs, uint8_t* out_dns, uint8_t* out_add, uint8_t* out_drops,
                               uint64_t* out_rtt_min, uint8_t* out_present, size_t cap);
size_t oracle_flowmap_evict(oracle_flowmap* m, uint8_t* out_recs, uint8_t* out_dns,
                            uint8_t* out_add, uint8_t* out_present, size_t cap);

/* --- KERNEL_MAP mode: bpf/flows.c:76-143 (update_existing_flow, add_observed_intf)
 * and :222-288 (lookup / insert NOEXIST / spill).  Each input record is one packet
 * event whose metrics carry (ts=start, len=bytes, flags, dscp, sampling, ifindex,
 * direction, eth, macs, tls*). */
typedef struct oracle_kmap oracle_kmap;
uint64_t oracle_flowmap_packets_kmap(oracle_flowmap* m, const uint8_t* wire, size_t n);   /* returns OBSERVED_INTF_MISSED increments */
oracle_kmap* oracle_kmap_new(size_t max_entries, int ringbuf_fallback);
void   oracle_kmap_free(oracle_kmap* m);
void   oracle_kmap_packets(oracle_kmap* m, const uint8_t* wire, size_t n);
size_t oracle_kmap_len(const oracle_kmap* m);
size_t oracle_kmap_evict(oracle_kmap* m, uint8_t* out_recs, size_t cap);
/* records spilled to the ring buffer (errno = E2BIG = 7), in order */
size_t oracle_kmap_spilled(oracle_kmap* m, uint8_t* out_recs, size_t cap);
uint64_t oracle_kmap_counter_fail_create(const oracle_kmap* m);
uint64_t oracle_kmap_counter_intf_missed(const oracle_kmap* m);

/* --- (f4) raw-header front end: bpf/utils.h:24-167 (set_flags, fill_l4info, fill_iphdr, fill_ip6hdr, fill_ethhdr) and
 * the single-packet flow flow_monitor builds, bpf/flows.c:176-245.  snap = 24-byte header (ts u64, len u32, if_index
 * u32, sampling u32, cap_len u16, direction u8, pad) + frame bytes; returns 1 and fills rec144 on SUBMIT, 0 on DISCARD.
 * SOURCE-PINNED ONLY: the reference has no unit test for its eBPF parsing. */
int    oracle_parse_snap(const uint8_t* snap, uint32_t stride, uint8_t* rec144);
size_t oracle_parse_snaps(const uint8_t* snaps, size_t n, uint32_t stride, uint8_t* out_recs, uint32_t* src_of);
/* the flow filter in front of the map update: bpf/flows_filter.h:14-255 (is_flow_filtered, do_flow_filter_lookup, LPM
 * semantics of filter_map / peer_filter_map) + check_and_do_flow_filtering (bpf/utils.h:179-222).  rules: n x 64-byte
 * fa_filter_rule images, peers: n x 20-byte CIDRs (layouts restated in oracle.c).  rec144 = the packet's parsed record;
 * returns 1 when the packet is skipped; counters[3] += (accept, reject, nomatch); a matching rule's sample is written to
 * the record's sampling field.  SOURCE-PINNED ONLY. */
int    oracle_filter_packet(const uint8_t* rules, size_t n_rules, const uint8_t* peers, size_t n_peers, uint8_t* rec144,
                            uint64_t counters[3]);
size_t oracle_parse_snaps_filtered(const uint8_t* snaps, size_t n, uint32_t stride, const uint8_t* rules, size_t n_rules,
                                   const uint8_t* peers, size_t n_peers, uint8_t* out_recs, uint32_t* src_of,
                                   uint64_t counters[3]);

/* --- K7: DNS query -> response correlation, the dns_flows map of bpf/dns_tracker.h:23-37,68-127 together with what
 * flow_monitor does with the result (bpf/flows.c:210-213,291-330: a per-flow dns_metrics sample when pkt.dns_id != 0 or
 * dns_errno != 0).  Input: n 104-byte DNS packets in stream order = flow_id (40 B) + dns_metrics layout (64 B) with
 * end_mono_time_ts = the packet's timestamp, id / flags = the DNS header fields in host order (QR = 0x8000),
 * eth_protocol, name; latency / errno / start are ignored.  Per packet, sequentially:
 *   query    (QR = 0): bpf_map_update_elem(dns_flows, {src, dst, ports, id, proto}, ts, BPF_NOEXIST); the return value is the
 *                      packet's dns_errno (track_dns_packet returns `ret`): 0, or -EEXIST / -E2BIG truncated to the u8 errno
 *                      field (239 / 249) -> a sample with id = flags = latency = 0 and no name (pkt.dns_id stays 0)
 *   response (QR = 1): lookup of the REVERSED tuple: found -> latency = ts - value, entry deleted; else dns_errno = ENOENT (2);
 *                      sample {start = end = ts, id, flags, latency, eth_protocol, name, errno} unless id == 0 and errno == 0
 * out_samples (n x 104 B at most) receives the samples in stream order, ready for AccumulateDNS; returns their number.
 * max_entries = the map's capacity (reference: 1 << 20).  SOURCE-PINNED ONLY: the reference has no unit test for it. */
typedef struct oracle_dnscorr oracle_dnscorr;
oracle_dnscorr* oracle_dnscorr_new(size_t max_entries);
void   oracle_dnscorr_free(oracle_dnscorr* m);
size_t oracle_dnscorr_packets(oracle_dnscorr* m, const uint8_t* pkts, size_t n, uint8_t* out_samples);
size_t oracle_dnscorr_pending(const oracle_dnscorr* m);
/* FlowFetcher.lookupAndDeleteDNSMap (pkg/tracer/tracer.go:1235-1257): delete every query with
 * time.Duration(now - ts) >= timeout (signed 64-bit compare); returns the number deleted */
size_t oracle_dnscorr_purge(oracle_dnscorr* m, uint64_t mono_now_ns, uint64_t timeout_ns);

/* --- hashes + sketches (this repo's spec; PARITY UNPINNED, see header comment) */
uint64_t oracle_key_premix(const uint8_t* key40);
uint64_t oracle_slot_hash(const uint8_t* key40);
uint64_t oracle_owner_hash(const uint8_t* key40);
void   oracle_cms_update(uint64_t* table, uint32_t log2_width, uint32_t depth, uint64_t seed,
                         const uint8_t* wire, size_t n);          /* += packets per record */
void   oracle_cms_query(const uint64_t* table, uint32_t log2_width, uint32_t depth, uint64_t seed,
                        const uint8_t* keys40, size_t n, uint64_t* est);
void   oracle_hll_update(uint8_t* regs, uint32_t p, uint64_t seed, const uint8_t* wire, size_t n);
double oracle_hll_estimate(const uint8_t* regs, uint32_t p);

#ifdef __cplusplus
}
#endif
#endif

// snaps.cu — (f4) raw-header front end: packet snapshots -> single-packet flow records, on the device.
//
// Replaces the part of flow_monitor that runs before the map is touched: fill_ethhdr / fill_iphdr / fill_ip6hdr /
// fill_l4info / set_flags (reference bpf/utils.h:24-167) and the construction of new_flow (bpf/flows.c:176-245).
// Input: n fixed-stride snapshots (24-byte fa_packet_snap_hdr + the first stride-24 bytes of the frame).  Output: the
// 144-byte records of the packets the reference would SUBMIT, in input order (K1's order-dependent fields need the
// stream order), plus for every record the index of its snapshot (so that an FA_FULL cut can be reported in snapshots).
//
// Two kernels over the same static partition of the tiles (contiguous ranges per CTA):
//   snap_count_kernel   validity only (two 8-byte loads per snapshot)  -> records per CTA
//   snap_parse_kernel   exclusive prefix over the CTA counts, then per tile: coalesced staging of the snapshots in shared
//                       memory, one thread per packet parses out of shared memory, warp ballots give the stable position,
//                       the records are laid out in shared memory and leave as one contiguous, coalesced run.
// HBM-bound byte work: stride + 144 algorithmic bytes per packet (+ 16 re-read by the count pass).
#include "kernels.cuh"

namespace fa {

constexpr int kSnapTile = 256;                  // packets per tile == threads per CTA
constexpr uint32_t kSnapHdr = 24;

__device__ __forceinline__ uint32_t h16(const uint8_t* p) { return *reinterpret_cast<const uint16_t*>(p); }   // raw (little-endian view)
__device__ __forceinline__ uint32_t raw32(const uint8_t* p) { return h16(p) | (h16(p + 2) << 16); }           // p is 2-byte aligned
__device__ __forceinline__ uint32_t be16(const uint8_t* p) { const uint32_t v = h16(p); return ((v & 0xFFu) << 8) | (v >> 8); }

// fill_ethhdr's verdict from the two words that hold cap_len and the EtherType: does the reference SUBMIT this packet?
__device__ __forceinline__ bool snap_submits(uint64_t w16, uint64_t w32, uint32_t stride) {
    uint32_t end = (uint32_t)(w16 >> 32) & 0xFFFFu;                   // cap_len @20
    end = min(end, stride - kSnapHdr);
    if (end < 14u) return false;                                      // utils.h:154-156
    const uint32_t raw = (uint32_t)(w32 >> 32) & 0xFFFFu;             // frame bytes 12,13 = snapshot bytes 36,37
    const uint32_t eth = ((raw & 0xFFu) << 8) | (raw >> 8);
    if (eth == 0x0800u) return end >= 34u;                           // utils.h:115-118
    if (eth == 0x86DDu) return end >= 54u;                            // utils.h:136-139
    return false;                                                     // utils.h:166
}

__device__ __forceinline__ uint32_t tcp_flag_of(uint32_t f) {          // utils.h:24-50: first match of the chain, one flag per packet
    const bool fin = f & 0x01u, syn = f & 0x02u, rst = f & 0x04u, psh = f & 0x08u, ack = f & 0x10u, urg = f & 0x20u,
               ece = f & 0x40u, cwr = f & 0x80u;
    return (ack && syn) ? 0x100u : (ack && fin) ? 0x200u : (ack && rst) ? 0x400u : fin ? 0x01u : syn ? 0x02u : ack ? 0x10u
         : rst ? 0x04u : psh ? 0x08u : urg ? 0x20u : ece ? 0x40u : cwr ? 0x80u : 0u;
}

// One submitted packet -> the nine 16-byte chunks of its flow record.  S = the snapshot in shared memory.
__device__ __forceinline__ void snap_to_record(const uint8_t* S, uint32_t stride, uint4 r[kRecChunks]) {
    const uint2 ts = *reinterpret_cast<const uint2*>(S);
    const uint32_t len = *reinterpret_cast<const uint32_t*>(S + 8), ifindex = *reinterpret_cast<const uint32_t*>(S + 12);
    const uint32_t sampling = *reinterpret_cast<const uint32_t*>(S + 16);
    const uint32_t end = min(h16(S + 20), stride - kSnapHdr), direction = S[22];
    const uint8_t* d = S + kSnapHdr;
    const uint32_t eth = be16(d + 12);
    uint32_t l4, proto, dscp;
    if (eth == 0x0800u) {                                             // fill_iphdr: IPv4 carried as ::ffff:a.b.c.d, options not skipped
        l4 = 34u;
        r[0] = make_uint4(0u, 0u, 0xFFFF0000u, raw32(d + 26));
        r[1] = make_uint4(0u, 0u, 0xFFFF0000u, raw32(d + 30));
        dscp = ((uint32_t)d[15] >> 2) & 0x3Fu;
        proto = d[23];
    } else {                                                          // fill_ip6hdr: nexthdr is taken as the transport protocol
        l4 = 54u;
        r[0] = make_uint4(raw32(d + 22), raw32(d + 26), raw32(d + 30), raw32(d + 34));
        r[1] = make_uint4(raw32(d + 38), raw32(d + 42), raw32(d + 46), raw32(d + 50));
        dscp = ((be16(d + 14) >> 4) >> 2) & 0x3Fu;
        proto = d[20];
    }
    uint32_t sport = 0, dport = 0, flags = 0, itype = 0, icode = 0;   // fill_l4info: a transport header that does not fit is not parsed
    const uint32_t need = proto == 6u ? 20u : proto == 132u ? 12u : (proto == 17u || proto == 1u || proto == 58u) ? 8u : 0xFFFFu;
    if (l4 + need <= end) {
        if (proto == 1u || proto == 58u) { itype = d[l4]; icode = d[l4 + 1]; }
        else { sport = be16(d + l4); dport = be16(d + l4 + 2); if (proto == 6u) flags = tcp_flag_of(d[l4 + 13]); }
    }
    r[2] = make_uint4(sport | (dport << 16), proto | (itype << 8) | (icode << 16), ts.x, ts.y);     // key tail | start = ts
    r[3] = make_uint4(ts.x, ts.y, len, 0u);                                                        // end = ts | bytes = len
    r[4] = make_uint4(1u, eth | (flags << 16), raw32(d + 6), h16(d + 10) | (h16(d) << 16));        // packets | eth, flags | src_mac | dst_mac..
    r[5] = make_uint4(raw32(d + 2), ifindex, 0u, sampling);                                        // ..dst_mac | if_index | lock | sampling
    r[6] = make_uint4(direction | (dscp << 16), 0u, 0u, 0u);                                       // direction, errno, dscp, nb_observed_intf
    r[7] = make_uint4(0u, 0u, 0u, 0u);
    r[8] = make_uint4(0u, 0u, 0u, 0u);
}

// ---- flow filter: bpf/flows_filter.h:14-255 + check_and_do_flow_filtering (bpf/utils.h:179-222) on the parsed record ----
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return (v >> 24) | ((v >> 8) & 0xFF00u) | ((v << 8) & 0xFF0000u) | (v << 24); }

// BPF_MAP_TYPE_LPM_TRIE match of one entry: prefix_len <= the key's and the first prefix_len bits agree (big-endian words)
__device__ __forceinline__ bool lpm_match(const uint32_t ew[4], uint32_t pl, const uint32_t kw[4], uint32_t kp) {
    bool ok = pl <= kp;
#pragma unroll
    for (int w = 0; w < 4; w++) {
        const int bits = (int)pl - 32 * w;
        const uint32_t m = bits >= 32 ? 0xFFFFFFFFu : bits <= 0 ? 0u : 0xFFFFFFFFu << (32 - bits);
        ok = ok && ((ew[w] ^ kw[w]) & m) == 0u;
    }
    return ok;
}

// flow_filter_setup_lookup_key: IPv4 keys are the 4 address bytes with prefix 32, IPv6 keys the 16 bytes with prefix 128
__device__ __forceinline__ void filter_key_of(uint4 ip, bool v4, uint32_t kw[4], uint32_t& kp) {
    if (v4) { kw[0] = bswap32(ip.w); kw[1] = kw[2] = kw[3] = 0u; kp = 32u; }
    else { kw[0] = bswap32(ip.x); kw[1] = bswap32(ip.y); kw[2] = bswap32(ip.z); kw[3] = bswap32(ip.w); kp = 128u; }
}

// do_flow_filter_lookup: > 0 when the longest-prefix rule for `key_ip` matches the packet; action / sampling keep what the rule
// wrote even when it ends with 0 (the caller's no-match branch looks at the action)
__device__ int filter_lookup(const FilterSet& F, uint4 key_ip, uint4 peer_ip, bool v4, const uint4 r[kRecChunks], uint32_t& action,
                             uint32_t& sampling) {
    uint32_t kw[4], kp;
    filter_key_of(key_ip, v4, kw, kp);
    int best = -1; uint32_t best_len = 0;
    for (uint32_t i = 0; i < F.n_rules; i++) {
        const uint32_t pl = F.rules[i].prefix;
        if (lpm_match(F.rules[i].ipw, pl, kw, kp) && (best < 0 || pl > best_len)) { best = (int)i; best_len = pl; }
    }
    if (best < 0) return 0;
    const FilterRuleDev& R = F.rules[best];
    int result = 1;
    if (R.action != 2u) { action = R.action; result++; }
    if (R.peer) {
        uint32_t pw[4], pp;
        filter_key_of(peer_ip, v4, pw, pp);
        bool hit = false;
        for (uint32_t i = 0; i < F.n_peers; i++) hit = hit || lpm_match(F.peers[i].ipw, F.peers[i].prefix, pw, pp);
        if (hit) result++; else return 0;
    }
    if (R.sample) { sampling = R.sample; result++; }
    const uint32_t sport = r[2].x & 0xFFFFu, dport = r[2].x >> 16, proto = r[2].y & 0xFFu, flags = r[4].y >> 16;
    if (R.proto == proto || R.proto == 0u) {
        if (proto == 6u || proto == 17u || proto == 132u) {
            if ((R.dps != 0 && R.dpe == 0) || R.dp1 != 0 || R.dp2 != 0) { if (R.dps == dport || R.dp1 == dport || R.dp2 == dport) result++; else return 0; }
            else if (R.dps != 0 && R.dpe != 0) { if (R.dps <= dport && dport <= R.dpe) result++; else return 0; }
            if ((R.sps != 0 && R.spe == 0) || R.sp1 != 0 || R.sp2 != 0) { if (R.sps == sport || R.sp1 == sport || R.sp2 == sport) result++; else return 0; }
            else if (R.sps != 0 && R.spe != 0) { if (R.sps <= sport && sport <= R.spe) result++; else return 0; }
            if ((R.ps != 0 && R.pe == 0) || R.p1 != 0 || R.p2 != 0) {
                if (R.ps == sport || R.ps == dport || R.p1 == sport || R.p1 == dport || R.p2 == sport || R.p2 == dport) result++; else return 0;
            } else if (R.ps != 0 && R.pe != 0) {
                if ((R.ps <= sport && sport <= R.pe) || (R.ps <= dport && dport <= R.pe)) result++; else return 0;
            }
            if (proto == 6u && R.tcp_flags != 0) { if (R.tcp_flags == flags) result++; else return 0; }
        } else if (proto == 1u || proto == 58u) {
            if (R.icmp_type != 0) {
                if (R.icmp_type == ((r[2].y >> 8) & 0xFFu)) result++; else return 0;
                if (R.icmp_code != 0) { if (R.icmp_code == ((r[2].y >> 16) & 0xFFu)) result++; else return 0; }
            }
        }
    } else {
        return 0;
    }
    if (R.direction != 2u) { if (R.direction == (r[6].x & 0xFFu)) result++; else return 0; }
    if (R.filter_drops) return 0;                                    // needs a drop reason; flow_monitor passes 0
    return result;
}

// check_and_do_flow_filtering: true = skip the packet; which = 0 accept / 1 reject / 2 no match (the global counter to bump);
// the rule's sample, if any, lands in the record's sampling field
__device__ bool filter_skips(const FilterSet& F, uint4 r[kRecChunks], uint32_t& which) {
    const bool v4 = (r[4].y & 0xFFFFu) == 0x0800u;
    uint32_t action = 2u, sampling = r[5].w;                         // is_flow_filtered starts from MAX_FILTER_ACTIONS
    int result = filter_lookup(F, r[0], r[1], v4, r, action, sampling);            // source address; its peer is the destination
    if (result <= 0) result = filter_lookup(F, r[1], r[0], v4, r, action, sampling);
    r[5].w = sampling;
    if (result != 0 && action != 2u) { which = action == 1u ? 1u : 0u; return action == 1u; }
    which = 2u;
    return action == 0u || action == 2u;
}

__device__ __forceinline__ void snap_range(uint32_t n, uint32_t& t0, uint32_t& t1) {    // this CTA's contiguous tiles
    const uint32_t n_tiles = (n + kSnapTile - 1) / kSnapTile, per = (n_tiles + gridDim.x - 1) / gridDim.x;
    t0 = min(n_tiles, blockIdx.x * per);
    t1 = min(n_tiles, t0 + per);
}

// Coalesced copy of one tile of snapshots into shared memory, four loads in flight per thread.
__device__ __forceinline__ void stage_snaps(uint64_t* tile, const uint8_t* snaps, uint32_t first, uint32_t cnt, uint32_t stride) {
    const uint64_t* G = reinterpret_cast<const uint64_t*>(snaps + (size_t)first * stride);
    const uint32_t total = cnt * (stride / 8), tid = threadIdx.x;
    uint32_t w = tid;
    for (; w + 3 * kSnapTile < total; w += 4 * kSnapTile) {
        const uint64_t a0 = G[w], a1 = G[w + kSnapTile], a2 = G[w + 2 * kSnapTile], a3 = G[w + 3 * kSnapTile];
        tile[w] = a0; tile[w + kSnapTile] = a1; tile[w + 2 * kSnapTile] = a2; tile[w + 3 * kSnapTile] = a3;
    }
    for (; w < total; w += kSnapTile) tile[w] = G[w];
}

// Pass 1: how many packets of every CTA's range go on.  Without a filter two 8-byte loads per snapshot decide; with one the
// tile is staged and parsed like in pass 2, the filter runs here (its verdict is kept per snapshot, its counters are final).
template <bool kFilter>
__global__ void __launch_bounds__(kSnapTile)
snap_count_kernel(const uint8_t* __restrict__ snaps, uint32_t n, uint32_t stride, FilterSet F, uint32_t* __restrict__ cta_count,
                  uint8_t* __restrict__ verdict, unsigned long long* __restrict__ filter_ctr) {
    FA_DYN_SMEM(sm);                                                  // kFilter: kSnapTile x stride bytes
    __shared__ uint32_t total, fc[3];
    if (threadIdx.x == 0) { total = 0; fc[0] = fc[1] = fc[2] = 0; }
    __syncthreads();
    uint32_t t0, t1; snap_range(n, t0, t1);
    uint32_t mine = 0;
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t first = t * kSnapTile, i = first + threadIdx.x;
        if (!kFilter) {
            if (i < n) {
                const uint8_t* S = snaps + (size_t)i * stride;
                mine += snap_submits(*reinterpret_cast<const uint64_t*>(S + 16), *reinterpret_cast<const uint64_t*>(S + 32), stride) ? 1u : 0u;
            }
        } else {
            const uint32_t cnt = min((uint32_t)kSnapTile, n - first);
            stage_snaps(reinterpret_cast<uint64_t*>(sm), snaps, first, cnt, stride);
            __syncthreads();
            if (threadIdx.x < cnt) {
                const uint8_t* S = sm + (size_t)threadIdx.x * stride;
                bool ok = snap_submits(*reinterpret_cast<const uint64_t*>(S + 16), *reinterpret_cast<const uint64_t*>(S + 32), stride);
                if (ok) {
                    uint4 r[kRecChunks];
                    snap_to_record(S, stride, r);
                    uint32_t which;
                    ok = !filter_skips(F, r, which);
                    atomicAdd(&fc[which], 1u);
                }
                verdict[i] = ok ? 1 : 0;
                mine += ok ? 1u : 0u;
            }
            __syncthreads();                                          // the tile is re-used
        }
    }
    const uint32_t wsum = __reduce_add_sync(0xFFFFFFFFu, mine);
    if ((threadIdx.x & 31) == 0 && wsum) atomicAdd(&total, wsum);
    __syncthreads();
    if (threadIdx.x == 0) cta_count[blockIdx.x] = total;
    if (kFilter && threadIdx.x < 3 && fc[threadIdx.x]) atomicAdd(&filter_ctr[threadIdx.x], (unsigned long long)fc[threadIdx.x]);
}

template <bool kFilter>
__global__ void __launch_bounds__(kSnapTile)
snap_parse_kernel(const uint8_t* __restrict__ snaps, uint32_t n, uint32_t stride, FilterSet F, const uint32_t* __restrict__ cta_count,
                  const uint8_t* __restrict__ verdict, uint4* __restrict__ out, uint32_t* __restrict__ src_of,
                  unsigned long long* __restrict__ n_out) {
    FA_DYN_SMEM(sm);                                                  // kSnapTile x max(stride, 144) bytes: snapshots in, records out
    __shared__ uint32_t wcnt[kSnapTile / 32];
    __shared__ uint32_t base_s;
    const uint32_t tid = threadIdx.x, lane = tid & 31u, warp = tid >> 5;
    if (warp == 0) {                                                  // exclusive prefix over the CTAs before this one
        uint32_t acc = 0;
        for (uint32_t b = lane; b < blockIdx.x; b += 32) acc += cta_count[b];
        acc = __reduce_add_sync(0xFFFFFFFFu, acc);
        if (lane == 0) {
            base_s = acc;
            if (blockIdx.x == gridDim.x - 1) *n_out = (unsigned long long)acc + cta_count[blockIdx.x];
        }
    }
    __syncthreads();
    uint32_t run = base_s;                                            // first output record of the current tile
    uint32_t t0, t1; snap_range(n, t0, t1);
    uint64_t* tile = reinterpret_cast<uint64_t*>(sm);
    uint4* otile = reinterpret_cast<uint4*>(sm);
    for (uint32_t t = t0; t < t1; t++) {
        const uint32_t first = t * kSnapTile, cnt = min((uint32_t)kSnapTile, n - first);
        stage_snaps(tile, snaps, first, cnt, stride);
        __syncthreads();
        const uint8_t* S = reinterpret_cast<const uint8_t*>(tile) + (size_t)tid * stride;
        const bool ok = tid < cnt && (kFilter ? verdict[first + tid] != 0
                                              : snap_submits(*reinterpret_cast<const uint64_t*>(S + 16), *reinterpret_cast<const uint64_t*>(S + 32), stride));
        uint4 r[kRecChunks];
        if (ok) {
            snap_to_record(S, stride, r);
            if (kFilter) { uint32_t which; (void)filter_skips(F, r, which); }     // again, for the rule's sampling (the verdict is pass 1's)
        }
        const uint32_t bal = __ballot_sync(0xFFFFFFFFu, ok);
        if (lane == 0) wcnt[warp] = __popc(bal);
        __syncthreads();                                              // every thread has parsed: the tile may be overwritten
        uint32_t before = 0, tile_total = 0;
#pragma unroll
        for (int w = 0; w < kSnapTile / 32; w++) { const uint32_t c = wcnt[w]; before += (w < (int)warp) ? c : 0u; tile_total += c; }
        if (ok) {
            const uint32_t pos = before + __popc(bal & ((1u << lane) - 1u));
#pragma unroll
            for (int c = 0; c < kRecChunks; c++) otile[pos * kRecChunks + c] = r[c];
            src_of[run + pos] = first + tid;
        }
        __syncthreads();
        uint4* O = out + (size_t)run * kRecChunks;
        for (uint32_t q = tid; q < tile_total * kRecChunks; q += kSnapTile) O[q] = otile[q];
        run += tile_total;
        __syncthreads();                                              // otile / wcnt are re-used
    }
}

#ifndef FA_HOST_EMUL
int launch_parse_snaps(const uint8_t* snaps, uint32_t n, uint32_t stride, const FilterSet* filter, uint32_t* cta_count, uint8_t* verdict,
                       uint4* out_recs, uint32_t* src_of, unsigned long long* n_out, unsigned long long* filter_ctr, int sm_count,
                       cudaStream_t st) {
    if (!n) return 0;
    const uint32_t n_tiles = (n + kSnapTile - 1) / kSnapTile;
    const uint32_t grid = std::min<uint32_t>(n_tiles, (uint32_t)std::min(sm_count * 4, kSnapMaxCtas));
    const size_t smem = (size_t)kSnapTile * std::max<uint32_t>(stride, kRecBytes);
    if (filter && filter->n_rules) {
        snap_count_kernel<true><<<grid, kSnapTile, (size_t)kSnapTile * stride, st>>>(snaps, n, stride, *filter, cta_count, verdict, filter_ctr);
        snap_parse_kernel<true><<<grid, kSnapTile, smem, st>>>(snaps, n, stride, *filter, cta_count, verdict, out_recs, src_of, n_out);
    } else {
        FilterSet none{};
        snap_count_kernel<false><<<grid, kSnapTile, 0, st>>>(snaps, n, stride, none, cta_count, verdict, filter_ctr);
        snap_parse_kernel<false><<<grid, kSnapTile, smem, st>>>(snaps, n, stride, none, cta_count, verdict, out_recs, src_of, n_out);
    }
    return 2;
}
#endif  // FA_HOST_EMUL

}  // namespace fa

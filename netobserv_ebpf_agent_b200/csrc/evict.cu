// evict.cu — K2 flow_evict: lookup-and-delete every live flow of a table.
// Replaces FlowFetcher.LookupAndDeleteMap (reference pkg/tracer/tracer.go:1063-1157: iterate
// keys + one LookupAndDelete syscall per key) and the map hand-over of Accounter.evict
// (pkg/flow/account.go:67-68,86-87,102-124).
//
// An occupancy bitmap finds the live slots; 8 lanes per live flow: one coalesced 128-byte
// identity-line load, the 32-byte hot line by two of the lanes, shuffles to reassemble the
// 144-byte record (9 x 16-byte stores), one output-cursor atomic per 32 slots, and in-place
// clearing so the table comes out empty.
#include "kernels.cuh"

namespace fa {

// kDrain = false: lookup-and-delete (the flow is removed).  kDrain = true: "lookup-and-reset": flows that
// received records since the last drain are emitted and their hot line is zeroed, but they stay in the table
// (identity line, tag, bitmap untouched) so the next batch hits them on the fast path.
template <bool kDrain>
__global__ void __launch_bounds__(256)
evict_kernel(Table t, uint4* __restrict__ out, uint32_t* __restrict__ slot_of_out, unsigned long long cap, Counters* ctr) {
    const int lane = threadIdx.x & 31, g = lane >> 3, j = lane & 7;
    const uint64_t slots = t.mask + 1;
    const uint64_t n_words = slots >> 5;               // slots is a power of two >= 1024
    const uint64_t warp_global = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const uint64_t n_warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
    // The occupancy bitmap (1 bit per slot, set when a flow is created) lets a warp skip 1024 empty slots with one
    // coalesced load: eviction costs ~ (128 + 32 + 144) bytes per LIVE flow, not per slot.  One output-cursor
    // atomic per warp iteration (1024 slots).
    for (uint64_t w0 = warp_global * 32; w0 < n_words; w0 += n_warps * 32) {
        const uint64_t wi = w0 + lane;                  // lane l owns bitmap word w0 + l
        uint32_t mybits = wi < n_words ? t.occ[wi] : 0u;
        if (!kDrain && mybits) t.occ[wi] = 0u;
        uint32_t nonempty = __ballot_sync(0xFFFFFFFFu, mybits != 0u);
        if (nonempty == 0u) continue;
        if (kDrain) {
            // which of the live flows received records since the last drain?  Every lane walks the set bits of its
            // own bitmap word (32 independent chains of 32-byte hot-line reads in flight per warp).
            uint32_t rest = mybits, myact = 0;
            while (rest) {
                const int b = __ffs(rest) - 1; rest &= rest - 1;
                const uint64_t slot = wi * 32 + b;
                const uint4 h0 = ld_cg_u4(&t.hot[slot * 2]), h1 = ld_cg_u4(&t.hot[slot * 2 + 1]);
                if ((h0.x | h0.y | h0.z | h0.w | h1.x | h1.y | h1.z | h1.w) != 0u) myact |= 1u << b;
            }
            mybits = myact;
            nonempty = __ballot_sync(0xFFFFFFFFu, mybits != 0u);
            if (nonempty == 0u) continue;
        }
        // one reservation for everything this warp iteration emits; exclusive prefix over the lanes' word counts
        const uint32_t mycnt = __popc(mybits);
        uint32_t incl = mycnt;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d); if (lane >= d) incl += v; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(&ctr->evict_out, (unsigned long long)total);
        base = __shfl_sync(0xFFFFFFFFu, base, 0);
        // rounds of 4 flows (one per 8-lane group), taken across all 32 words of the window
        for (uint32_t f0 = 0; f0 < total; f0 += 4) {
            {
                const uint32_t f = f0 + g;                              // this group's flow number in the window
                const bool live = f < total;
                // owner word = first lane whose inclusive prefix exceeds f (f differs per group: four ballots)
                uint32_t src = 0, within = 0;
                {
                    const uint32_t m0 = __ballot_sync(0xFFFFFFFFu, incl > f0), m1 = __ballot_sync(0xFFFFFFFFu, incl > f0 + 1),
                             m2 = __ballot_sync(0xFFFFFFFFu, incl > f0 + 2), m3 = __ballot_sync(0xFFFFFFFFu, incl > f0 + 3);
                    const uint32_t mm = g == 0 ? m0 : g == 1 ? m1 : g == 2 ? m2 : m3;
                    src = mm ? (uint32_t)(__ffs(mm) - 1) : 0u;
                }
                const uint32_t sbits = __shfl_sync(0xFFFFFFFFu, mybits, src);
                const uint32_t sincl = __shfl_sync(0xFFFFFFFFu, incl, src), scnt = __shfl_sync(0xFFFFFFFFu, mycnt, src);
                within = live ? f - (sincl - scnt) : 0u;                // rank of the flow inside its word
                const uint32_t pos = live ? __fns(sbits, 0, within + 1) : 0u;
                const uint64_t slot = (w0 + src) * 32 + pos;
                const unsigned long long idx = base + f;
                uint4 line = make_uint4(0, 0, 0, 0), hot = make_uint4(0, 0, 0, 0);
                if (live) line = ld_cg_u4(&t.ident[slot * 8 + j]);
                if (live && j < 2) hot = ld_cg_u4(&t.hot[slot * 2 + j]);
                // hot chunk 0 = (bytes, nstart), hot chunk 1 = (end, packets, flags)
                const uint32_t b_lo = __shfl_sync(0xFFFFFFFFu, hot.x, g * 8), b_hi = __shfl_sync(0xFFFFFFFFu, hot.y, g * 8);
                const uint32_t ns_lo = __shfl_sync(0xFFFFFFFFu, hot.z, g * 8), ns_hi = __shfl_sync(0xFFFFFFFFu, hot.w, g * 8);
                const uint32_t e_lo = __shfl_sync(0xFFFFFFFFu, hot.x, g * 8 + 1), e_hi = __shfl_sync(0xFFFFFFFFu, hot.y, g * 8 + 1);
                const uint32_t pk = __shfl_sync(0xFFFFFFFFu, hot.z, g * 8 + 1), fl = __shfl_sync(0xFFFFFFFFu, hot.w, g * 8 + 1);
                if (live && idx < cap) {
                    uint4* O = out + idx * kRecChunks;
                    const uint64_t start = 0ull - u64_of(ns_lo, ns_hi);       // nstart = -start; 0 stays 0
                    if (j < 2) {
                        O[j] = line;                                            // key[0..32)
                    } else if (j == 2) {
                        O[2] = make_uint4(line.x, line.y & 0x00FFFFFFu, (uint32_t)start, (uint32_t)(start >> 32));
                    } else if (j == 3) {
                        O[3] = make_uint4(e_lo, e_hi, b_lo, b_hi);              // end, bytes
                        O[4] = make_uint4(pk, (line.y & 0xFFFFu) | (fl << 16), line.z, line.w);   // packets, eth|flags, desc[0..8)
                    } else {
                        O[j + 1] = line;                                        // desc[8..72)
                    }
                    if (slot_of_out && j == 0) slot_of_out[idx] = (uint32_t)slot;   // for the feature pass
                }
                // delete: tag -> EMPTY, hot line -> identity  (drain: hot line only)
                if (!kDrain && live && j == 2) *reinterpret_cast<uint2*>(reinterpret_cast<uint8_t*>(&t.ident[slot * 8 + 2]) + 8) = make_uint2(0u, 0u);
                if (live && j < 2) t.hot[slot * 2 + j] = make_uint4(0, 0, 0, 0);
            }
        }
    }
}

#ifndef FA_HOST_EMUL
int launch_evict(const Table& table, uint4* out_recs, uint32_t* slot_of_out,
                 unsigned long long cap, Counters* ctr, int sm_count, cudaStream_t st, bool drain) {
    if (drain) evict_kernel<true><<<sm_count * 8, 256, 0, st>>>(table, out_recs, slot_of_out, cap, ctr);
    else evict_kernel<false><<<sm_count * 8, 256, 0, st>>>(table, out_recs, slot_of_out, cap, ctr);
    return 1;
}
#endif  // FA_HOST_EMUL

}  // namespace fa

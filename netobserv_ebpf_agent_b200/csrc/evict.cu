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
            // which of the live flows received records since the last drain?  One lane per slot of a word:
            // the 32 hot lines of a word are 1 KB of contiguous memory.
            uint32_t ne = nonempty, myact = 0;
            while (ne) {
                const int src = __ffs(ne) - 1; ne &= ne - 1;
                const uint32_t bits = __shfl_sync(0xFFFFFFFFu, mybits, src);
                bool act = false;
                if ((bits >> lane) & 1u) {
                    const uint64_t slot = (w0 + src) * 32 + lane;
                    const uint4 h0 = ld_cg_u4(&t.hot[slot * 2]), h1 = ld_cg_u4(&t.hot[slot * 2 + 1]);
                    act = (h0.x | h0.y | h0.z | h0.w | h1.x | h1.y | h1.z | h1.w) != 0u;
                }
                const uint32_t am = __ballot_sync(0xFFFFFFFFu, act);
                if (lane == src) myact = am;
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
        const unsigned long long mybase = base + incl - mycnt;
        while (nonempty) {
            const int src = __ffs(nonempty) - 1; nonempty &= nonempty - 1;
            uint32_t bits = __shfl_sync(0xFFFFFFFFu, mybits, src);
            const unsigned long long wbase = __shfl_sync(0xFFFFFFFFu, mybase, src);
            const uint64_t word_slot0 = (w0 + src) * 32;
            uint32_t done_before = 0;
            while (bits) {                              // up to 4 flows (one per 8-lane group) per round
                const uint32_t pos = __fns(bits, 0, g + 1);
                const bool live = pos < 32u;
                const uint64_t slot = word_slot0 + (live ? pos : 0u);
                const unsigned long long idx = wbase + done_before + g;
                const uint32_t taken = min(4, __popc(bits));
                for (uint32_t k = 0; k < taken; k++) bits &= bits - 1;
                done_before += taken;
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

int launch_evict(const Table& table, uint4* out_recs, uint32_t* slot_of_out,
                 unsigned long long cap, Counters* ctr, int sm_count, cudaStream_t st, bool drain) {
    if (drain) evict_kernel<true><<<sm_count * 8, 256, 0, st>>>(table, out_recs, slot_of_out, cap, ctr);
    else evict_kernel<false><<<sm_count * 8, 256, 0, st>>>(table, out_recs, slot_of_out, cap, ctr);
    return 1;
}

}  // namespace fa

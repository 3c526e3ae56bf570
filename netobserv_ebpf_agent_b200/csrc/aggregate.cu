// aggregate.cu — K1 flow_aggregate: fold a batch of 144-byte flow records into the
// open-addressed flow table with the semantics of the reference's userspace Accounter
// (pkg/flow/account.go:82-96 + pkg/model/flow_content.go:28-61).
//
// Pipeline per CTA (persistent, 256 threads, tiles of 256 records):
//   TMA bulk copy (cp.async.bulk + mbarrier, 2 stages) stages a 36 KB tile of records in
//   shared memory -> each thread hashes one record -> duplicates of a key inside the tile
//   elect one representative through a shared-memory index table and fold into it with
//   shared-memory atomics -> each warp then walks its representatives 4 at a time, 8 lanes
//   per flow: one coalesced 128-byte identity-line load per probe, a masked 16-byte compare
//   per lane, ballot to agree on hit / miss / claim, and five fire-and-forget reductions
//   (RED.add/max/or) on the 32-byte hot line.
//
// Exactness of the order-dependent fields (eth_protocol / dscp / sampling = last non-zero,
// MACs = first non-zero, everything else = first record; flow_content.go:45-59,
// account.go:95): while every record of a flow carries the same 74-byte descriptor the
// merge result does not depend on order, so the fast path only *compares* descriptors.
// Any mismatch flags the flow TAG_DIRTY and the two re-fold kernels below recompute those
// fields from the batch in stream-index order.
#include "kernels.cuh"

namespace fa {

constexpr int kTile      = 256;                 // records per tile == threads per CTA
constexpr int kStages    = 2;
constexpr int kTileChunks = kTile * kRecChunks; // uint4 per tile
constexpr int kRepSlots  = 512;
constexpr uint32_t kRepEmpty = 0xFFFFFFFFu;
constexpr uint32_t kProbeLimit = 8192;

struct __align__(128) AggSmem {
    uint4    tile[kStages][kTileChunks];          // 73,728 B
    uint4    acc[kTile * 2];                      //  8,192 B  hot-line layout per record slot
    unsigned long long hs[kTile];                 //  2,048 B  slot hash of each record
    uint32_t rep[2][kRepSlots];                   //  4,096 B
    uint8_t  tdirty[kTile];                       //    256 B
    unsigned long long full_bar[kStages];
    uint32_t n_insert, n_spill, any_dirty, pad;
};

__device__ __forceinline__ void issue_tile_load(AggSmem& s, int stage, const uint4* recs, uint32_t n, uint32_t tile_idx) {
    const uint32_t first = tile_idx * kTile;
    const uint32_t cnt = min((uint32_t)kTile, n - first);
    const uint32_t bytes = cnt * kRecBytes;
    mbar_expect_tx(&s.full_bar[stage], bytes);
    tma_load_1d(&s.tile[stage][0], recs + (size_t)first * kRecChunks, bytes, &s.full_bar[stage]);
}

template <bool kSketch>
__global__ void __launch_bounds__(kTile, 2)
aggregate_kernel(const uint4* __restrict__ recs, uint32_t n, Table t, uint64_t epoch, Counters* ctr,
                 uint32_t* __restrict__ spill_idx, SketchParams sk) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    AggSmem& s = *reinterpret_cast<AggSmem*>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t n_tiles = (n + kTile - 1) / kTile;

    if (tid == 0) {
        for (int i = 0; i < kStages; i++) mbar_init(&s.full_bar[i], 1);
        fence_barrier_init();
        s.n_insert = 0; s.n_spill = 0; s.any_dirty = 0;
    }
    s.rep[0][tid] = kRepEmpty; s.rep[0][tid + kTile] = kRepEmpty;
    s.rep[1][tid] = kRepEmpty; s.rep[1][tid + kTile] = kRepEmpty;
    __syncthreads();
    if (tid == 0) {
        for (int i = 0; i < kStages; i++) {
            uint32_t ti = blockIdx.x + i * gridDim.x;
            if (ti < n_tiles) issue_tile_load(s, i, recs, n, ti);
        }
    }

    const int g = lane >> 3;                  // flow group inside the warp (4 groups of 8 lanes)
    const int j = lane & 7;                   // 16-byte chunk of the identity line handled by this lane
    const uint4 cmask = chunk_mask(j);
    const int rc = rec_chunk_of_line_chunk(j);
    uint32_t my_inserts = 0, my_spills = 0;

    for (uint32_t it = 0;; ++it) {
        const uint32_t tile_idx = blockIdx.x + it * gridDim.x;
        if (tile_idx >= n_tiles) break;
        const int stage = it % kStages;
        const uint32_t parity = (it / kStages) & 1u;
        const uint32_t first = tile_idx * kTile;
        const uint32_t cnt = min((uint32_t)kTile, n - first);
        mbar_wait(&s.full_bar[stage], parity);
        const uint4* T = s.tile[stage];
        uint32_t* rep = s.rep[it & 1];
        uint32_t* rep_next = s.rep[(it & 1) ^ 1];

        // ---------------------------------------------------------- P1: hash + elect
        const bool valid = (uint32_t)tid < cnt;
        uint4 r0 = make_uint4(0, 0, 0, 0), r1 = r0, r2 = r0, r3 = r0, r4 = r0;
        uint64_t h = 0, premix = 0;
        int my_rep = tid;
        s.acc[tid * 2] = make_uint4(0, 0, 0, 0);
        s.acc[tid * 2 + 1] = make_uint4(0, 0, 0, 0);
        s.tdirty[tid] = 0;
        if (valid) {
            const uint4* R = T + tid * kRecChunks;
            r0 = R[0]; r1 = R[1]; r2 = R[2]; r3 = R[3]; r4 = R[4];
            premix = key_premix(u64_of(r0.x, r0.y), u64_of(r0.z, r0.w), u64_of(r1.x, r1.y), u64_of(r1.z, r1.w),
                                u64_of(r2.x, r2.y));
            h = slot_hash(premix);
            s.hs[tid] = h;
            uint32_t rs = (uint32_t)(h >> 40) & (kRepSlots - 1);
            const uint4 m2 = chunk_mask(2);
            for (;;) {
                uint32_t old = atomicCAS(&rep[rs], kRepEmpty, (uint32_t)tid);
                if (old == kRepEmpty) break;
                const uint4* O = T + old * kRecChunks;
                uint4 o0 = O[0], o1 = O[1], o2 = O[2];
                if (eq4_masked(o0, r0, chunk_mask(0)) && eq4_masked(o1, r1, chunk_mask(1)) && eq4_masked(o2, r2, m2)) {
                    my_rep = (int)old;
                    break;
                }
                rs = (rs + 1) & (kRepSlots - 1);
            }
        }
        __syncthreads();

        // ---------------------------------------------------------- P2: fold duplicates into their representative
        const uint64_t v_start = u64_of(r2.z, r2.w), v_end = u64_of(r3.x, r3.y), v_bytes = u64_of(r3.z, r3.w);
        const uint32_t v_packets = r4.x, v_flags = r4.y >> 16;
        if (valid && my_rep != tid) {
            unsigned long long* A = reinterpret_cast<unsigned long long*>(&s.acc[my_rep * 2]);
            atomicAdd(&A[0], (unsigned long long)v_bytes);
            if (v_start) atomicMax(&A[1], (unsigned long long)(0ull - v_start));
            if (v_end) atomicMax(&A[2], (unsigned long long)v_end);
            uint32_t* A32 = reinterpret_cast<uint32_t*>(&A[3]);
            atomicAdd(&A32[0], v_packets);
            if (v_flags) atomicOr(&A32[1], v_flags);
            // exact descriptor compare against the representative (74 bytes, padding masked)
            const uint4* R = T + tid * kRecChunks;
            const uint4* O = T + my_rep * kRecChunks;
            bool same = eq4_masked(O[4], r4, chunk_mask(3));
#pragma unroll
            for (int c = 5; c < 9; c++) same = same && eq4_masked(O[c], R[c], chunk_mask(c - 1));
            if (!same) s.tdirty[my_rep] = 1;
        }
        rep_next[tid] = kRepEmpty; rep_next[tid + kTile] = kRepEmpty;
        __syncthreads();

        // ---------------------------------------------------------- P3: representatives -> global table
        const bool is_rep = valid && my_rep == tid;
        if (is_rep) {
            unsigned long long* A = reinterpret_cast<unsigned long long*>(&s.acc[tid * 2]);
            uint32_t* A32 = reinterpret_cast<uint32_t*>(&A[3]);
            A[0] += v_bytes;
            unsigned long long ns = 0ull - v_start;           // 0 stays 0 ("unset")
            if (ns > A[1]) A[1] = ns;
            if (v_end > A[2]) A[2] = v_end;
            A32[0] += v_packets;
            A32[1] |= v_flags;
            if (kSketch) {
                const uint32_t pk = A32[0];
                const uint64_t a = cms_hash_a(premix, sk.seed), b = cms_hash_b(premix, sk.seed);
                for (uint32_t d = 0; d < sk.depth; d++)
                    red_add_u64(sk.cms + ((size_t)d << sk.log2w) + cms_index(a, b, d, sk.log2w), pk);
                const uint64_t hh = hll_hash(premix, sk.seed);
                const uint32_t idx = (uint32_t)(hh >> (64 - sk.p));
                const uint64_t rest = hh << sk.p;
                uint32_t rho = rest ? (uint32_t)__clzll((long long)rest) + 1u : (64u - sk.p) + 1u;
                if (__ldcg(&sk.hll[idx]) < rho) red_max_u32(&sk.hll[idx], rho);
            }
        }
        __syncwarp();

        uint32_t pending = __ballot_sync(0xFFFFFFFFu, is_rep);
        while (pending) {
            const uint32_t src = __fns(pending, 0, g + 1);       // g-th pending representative of this warp
            const bool active = src < 32u;
#pragma unroll
            for (int k = 0; k < 4; k++) pending &= pending - 1;   // (x & (x-1)) of 0 is 0
            const int ridx = warp * 32 + (active ? (int)src : 0);  // record slot inside the tile
            const uint4 rchunk = T[ridx * kRecChunks + rc];
            const uint64_t hh = s.hs[ridx];
            const bool cta_dirty = s.tdirty[ridx] != 0;
            uint64_t slot = hh & t.mask;
            bool done = !active;
            uint32_t nprobe = 0;
            uint64_t reload_slot = ~0ull;
            for (;;) {
                uint4 line = make_uint4(0, 0, 0, 0);
                if (!done) line = ld_cg_u4(&t.ident[slot * 8 + j]);
                const uint32_t tag_lo = __shfl_sync(0xFFFFFFFFu, line.z, g * 8 + 2);
                const uint32_t tag_hi = __shfl_sync(0xFFFFFFFFu, line.w, g * 8 + 2);
                const uint64_t tag = u64_of(tag_lo, tag_hi);
                const uint32_t eqb = (__ballot_sync(0xFFFFFFFFu, eq4_masked(line, rchunk, cmask)) >> (g * 8)) & 0xFFu;
                const uint32_t state = (uint32_t)(tag & TAG_STATE_MASK);
                unsigned long long* tagp = reinterpret_cast<unsigned long long*>(&t.ident[slot * 8 + 2]) + 1;

                // ---- claim an empty slot
                uint32_t won = 0;
                if (!done && state == 0 && j == 2) {
                    const unsigned long long want = (epoch << TAG_EPOCH_SHIFT) | TAG_HAS_BASE | TAG_CLAIMED;
                    won = atomicCAS(tagp, 0ull, want) == 0ull ? 1u : 0u;
                }
                won = __shfl_sync(0xFFFFFFFFu, won, g * 8 + 2);
                if (won) {
                    uint4 v = and4(rchunk, cmask);
                    if (j == 2) {
                        *reinterpret_cast<uint2*>(&t.ident[slot * 8 + 2]) = make_uint2(v.x, v.y);   // key tail only
                    } else {
                        st_cg_u4(&t.ident[slot * 8 + j], v);     // j==3: word0 (aux) = 0, word1 = eth
                    }
                    __threadfence();
                }
                __syncwarp();
                bool hit = false;
                if (won) {
                    if (j == 2) {
                        const unsigned long long pub = (epoch << TAG_EPOCH_SHIFT) | TAG_HAS_BASE | TAG_PUBLISHED |
                                                       (cta_dirty ? TAG_DIRTY : 0ull);
                        *reinterpret_cast<volatile unsigned long long*>(tagp) = pub;
                        my_inserts++;
                        if (cta_dirty) s.any_dirty = 1;
                    }
                    hit = true;
                } else if (!done && state == (uint32_t)TAG_PUBLISHED) {
                    const bool born_now = (tag >> TAG_EPOCH_SHIFT) == epoch;
                    if (born_now && reload_slot != slot) {
                        // published during this launch: the chunks read together with the tag may
                        // predate it. Re-read the line once, ordered after the tag observation.
                        reload_slot = slot;
                        __threadfence();
                    } else if ((eqb & 0x07u) == 0x07u) {
                        hit = true;
                        const bool desc_eq = (eqb & 0xF8u) == 0xF8u;
                        if ((!desc_eq || cta_dirty) && j == 2) {
                            if (!(tag & TAG_DIRTY)) atomicOr(tagp, (unsigned long long)TAG_DIRTY);
                            s.any_dirty = 1;
                        }
                    } else {
                        slot = (slot + 1) & t.mask;
                        if (++nprobe > kProbeLimit) {           // table physically full: spill, never drop silently
                            done = true;
                            if (j == 0) {
                                unsigned long long k = atomicAdd(&ctr->scratch[2], 1ull);   // per-launch cursor
                                spill_idx[k] = first + (uint32_t)ridx;
                                my_spills++;
                            }
                        }
                    }
                }
                // state CLAIMED by someone else, or lost the CAS: retry the same slot next iteration
                if (hit) {
                    uint8_t* hot = reinterpret_cast<uint8_t*>(t.hot) + slot * kHotBytes;
                    const unsigned long long* A = reinterpret_cast<const unsigned long long*>(&s.acc[ridx * 2]);
                    if (j == 0) red_add_u64(hot, A[0]);
                    else if (j == 1) { if (A[1]) red_max_u64(hot + 8, A[1]); }
                    else if (j == 2) { if (A[2]) red_max_u64(hot + 16, A[2]); }
                    else if (j == 3) red_add_u32(hot + 24, (uint32_t)A[3]);
                    else if (j == 4) { uint32_t f = (uint32_t)(A[3] >> 32); if (f) red_or_u32(hot + 28, f); }
                    done = true;
                }
                if (__all_sync(0xFFFFFFFFu, done)) break;
            }
        }
        __syncthreads();                         // everyone is done with this stage
        if (tid == 0) {
            const uint32_t nt = tile_idx + kStages * gridDim.x;
            if (nt < n_tiles) { fence_proxy_async(); issue_tile_load(s, stage, recs, n, nt); }
        }
    }

    // ---------------------------------------------------------- counters
    my_inserts = __reduce_add_sync(0xFFFFFFFFu, my_inserts);
    my_spills = __reduce_add_sync(0xFFFFFFFFu, my_spills);
    if (lane == 0) {
        if (my_inserts) atomicAdd(&s.n_insert, my_inserts);
        if (my_spills) atomicAdd(&s.n_spill, my_spills);
    }
    __syncthreads();
    if (tid == 0) {
        if (s.n_insert) atomicAdd(&ctr->live, (unsigned long long)s.n_insert);
        if (s.n_spill) atomicAdd(&ctr->spills, (unsigned long long)s.n_spill);
        if (s.any_dirty) *reinterpret_cast<volatile unsigned long long*>(&ctr->dirty) = 1ull;
    }
}

// ------------------------------------------------------------------------------------
// Ordered re-fold of flows flagged TAG_DIRTY (rare path; both kernels exit at once when
// nothing was flagged).  fixup_scan: one thread per record finds its flow and, if dirty,
// reduces min/max record indices into the flow's scratch entry.  fixup_apply: one thread
// per scratch entry gathers the fields from the records those indices name.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_equal_line(const uint4* line, uint4 k0, uint4 k1, uint4 k2) {
    uint4 l0 = ld_cg_u4(line), l1 = ld_cg_u4(line + 1), l2 = ld_cg_u4(line + 2);
    return eq4_masked(l0, k0, chunk_mask(0)) && eq4_masked(l1, k1, chunk_mask(1)) && eq4_masked(l2, k2, chunk_mask(2));
}

__global__ void fixup_scan_kernel(const uint4* __restrict__ recs, uint32_t n, Table t, Counters* ctr, FixupScratch* scratch) {
    if (*reinterpret_cast<volatile unsigned long long*>(&ctr->dirty) == 0ull) return;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const uint4* R = recs + (size_t)i * kRecChunks;
        const uint4 k0 = R[0], k1 = R[1], k2 = R[2];
        const uint64_t h = slot_hash(key_premix(u64_of(k0.x, k0.y), u64_of(k0.z, k0.w), u64_of(k1.x, k1.y),
                                                u64_of(k1.z, k1.w), u64_of(k2.x, k2.y)));
        uint64_t slot = h & t.mask;
        bool found = false;
        unsigned long long tag = 0;
        for (uint32_t p = 0; p <= kProbeLimit; p++) {
            tag = ld_cg_u64(reinterpret_cast<const unsigned long long*>(&t.ident[slot * 8 + 2]) + 1);
            if ((tag & TAG_STATE_MASK) == 0) break;
            if (key_equal_line(&t.ident[slot * 8], k0, k1, k2)) { found = true; break; }
            slot = (slot + 1) & t.mask;
        }
        if (!found || !(tag & TAG_DIRTY)) continue;
        uint32_t* auxp = reinterpret_cast<uint32_t*>(&t.ident[slot * 8 + 3]);
        uint32_t aux = *reinterpret_cast<volatile uint32_t*>(auxp);
        if (aux == 0) {
            const uint32_t mine = (uint32_t)atomicAdd(&ctr->scratch[0], 1ull) + 1u;
            const uint32_t old = atomicCAS(auxp, 0u, mine);
            aux = old ? old : mine;
        }
        FixupScratch* sc = &scratch[aux];
        sc->slot_lo = (uint32_t)slot; sc->slot_hi = (uint32_t)(slot >> 32);
        const uint4 r4 = R[4], r5 = R[5], r6 = R[6];
        atomicMin(&sc->first, i);
        if (r4.y & 0xFFFFu) atomicMax(&sc->eth, i + 1);
        if ((r6.x >> 16) & 0xFFu) atomicMax(&sc->dscp, i + 1);
        if (r5.w) atomicMax(&sc->samp, i + 1);
        if (r4.z | (r4.w & 0xFFFFu)) atomicMin(&sc->smac, i);
        if ((r4.w >> 16) | r5.x) atomicMin(&sc->dmac, i);
    }
}

__global__ void fixup_apply_kernel(const uint4* __restrict__ recs, Table t, uint64_t epoch, Counters* ctr,
                                   FixupScratch* scratch, unsigned int* ticket) {
    if (*reinterpret_cast<volatile unsigned long long*>(&ctr->dirty) == 0ull) return;
    const uint32_t count = (uint32_t)*reinterpret_cast<volatile unsigned long long*>(&ctr->scratch[0]);
    uint32_t fixed = 0;
    for (uint32_t a = 1 + blockIdx.x * blockDim.x + threadIdx.x; a <= count; a += gridDim.x * blockDim.x) {
        FixupScratch sc = scratch[a];
        scratch[a] = FixupScratch{0xFFFFFFFFu, 0u, 0u, 0u, 0xFFFFFFFFu, 0xFFFFFFFFu, 0u, 0u};
        if (sc.first == 0xFFFFFFFFu) continue;            // index allocated but lost the CAS race
        const uint64_t slot = u64_of(sc.slot_lo, sc.slot_hi);
        uint4* L = &t.ident[slot * 8];
        unsigned long long* tagp = reinterpret_cast<unsigned long long*>(&L[2]) + 1;
        const unsigned long long tag = *tagp;
        const bool is_new = (tag >> TAG_EPOCH_SHIFT) == epoch;
        // current descriptor state: line chunks 3..7 <-> record chunks 4..8
        uint4 d3 = L[3], d4 = L[4], d5 = L[5], d6 = L[6], d7 = L[7];
        if (is_new) {                                      // state = the flow's first record, whole (account.go:95)
            const uint4* F = recs + (size_t)sc.first * kRecChunks;
            d3 = and4(F[4], chunk_mask(3)); d4 = F[5]; d5 = and4(F[6], chunk_mask(5)); d6 = F[7]; d7 = and4(F[8], chunk_mask(7));
        }
        // eth_protocol / dscp / sampling: last non-zero in stream order (flow_content.go:45-47,54-59)
        if (sc.eth)  { const uint4 x = recs[(size_t)(sc.eth - 1) * kRecChunks + 4]; d3.y = x.y & 0xFFFFu; }
        if (sc.dscp) { const uint4 x = recs[(size_t)(sc.dscp - 1) * kRecChunks + 6]; d5.x = (d5.x & 0xFF00FFFFu) | (x.x & 0x00FF0000u); }
        if (sc.samp) { const uint4 x = recs[(size_t)(sc.samp - 1) * kRecChunks + 5]; d4.w = x.w; }
        // MACs: first non-zero in stream order, only while still all-zero (flow_content.go:48-53)
        if ((d3.z | (d3.w & 0xFFFFu)) == 0u && sc.smac != 0xFFFFFFFFu) {
            const uint4 x = recs[(size_t)sc.smac * kRecChunks + 4];
            d3.z = x.z; d3.w = (d3.w & 0xFFFF0000u) | (x.w & 0xFFFFu);
        }
        if (((d3.w >> 16) | d4.x) == 0u && sc.dmac != 0xFFFFFFFFu) {
            const uint4 x4 = recs[(size_t)sc.dmac * kRecChunks + 4], x5 = recs[(size_t)sc.dmac * kRecChunks + 5];
            d3.w = (d3.w & 0xFFFFu) | (x4.w & 0xFFFF0000u); d4.x = x5.x;
        }
        d3.x = 0u;                                        // aux back to 0
        L[3] = d3; L[4] = d4; L[5] = d5; L[6] = d6; L[7] = d7;
        *tagp = tag & ~(unsigned long long)TAG_DIRTY;
        fixed++;
    }
    if (fixed) atomicAdd(&ctr->fixups_total, (unsigned long long)fixed);
    // last CTA out resets the per-launch state
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) {
            *ticket = 0u;
            ctr->dirty = 0ull;
            ctr->scratch[0] = 0ull;
            __threadfence();
        }
    }
}

int launch_aggregate(const AggLaunch& a, cudaStream_t st) {
    if (a.n == 0) return 0;
    static bool attr_done = false;
    const int smem = (int)sizeof(AggSmem);
    if (!attr_done) {
        cudaFuncSetAttribute(aggregate_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(aggregate_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_done = true;
    }
    const uint32_t n_tiles = (a.n + kTile - 1) / kTile;
    const int grid = (int)min((uint32_t)(a.sm_count * 2), n_tiles);
    if (a.sk.cms) aggregate_kernel<true><<<grid, kTile, smem, st>>>(a.recs, a.n, a.table, a.epoch, a.ctr, a.spill_idx, a.sk);
    else aggregate_kernel<false><<<grid, kTile, smem, st>>>(a.recs, a.n, a.table, a.epoch, a.ctr, a.spill_idx, a.sk);
    const int fgrid = a.sm_count * 2;
    fixup_scan_kernel<<<fgrid, 256, 0, st>>>(a.recs, a.n, a.table, a.ctr, a.scratch);
    fixup_apply_kernel<<<fgrid, 256, 0, st>>>(a.recs, a.table, a.epoch, a.ctr, a.scratch,
                                              reinterpret_cast<unsigned int*>(&a.ctr->scratch[1]));
    return 3;
}

}  // namespace fa

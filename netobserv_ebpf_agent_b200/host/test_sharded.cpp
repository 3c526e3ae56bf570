// test_sharded.cpp — a C++ client drives every GPU of the box through fa_sharded_* (no Python, no torch, no NCCL) and
// checks the result against a single-GPU engine folding the same stream: what a Go host would do over cgo
// (INTEGRATION.md §5).  Built and run by tests/test_gpu_host_cpp.py; with one GPU it runs as a 1-shard box.
//   usage: test_sharded [n_gpus (0 = all)] [records] [keys]
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cuda_runtime_api.h>

#include "../../include/flowagg.h"

#define CHECK(c) do { if (!(c)) { fprintf(stderr, "CHECK failed %s:%d: %s  (%s / %s)\n", __FILE__, __LINE__, #c, fa_sharded_last_error(), fa_last_error()); return 1; } } while (0)

static bool rec_less(const fa_flow_record& a, const fa_flow_record& b) { return memcmp(&a, &b, 39) < 0; }

int main(int argc, char** argv) {
    int ndev = 0;
    CHECK(cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0);
    int n_gpus = argc > 1 ? atoi(argv[1]) : 0;
    if (n_gpus <= 0 || n_gpus > ndev) n_gpus = ndev;
    const size_t n = argc > 2 ? strtoull(argv[2], nullptr, 0) : 3000000, n_keys = argc > 3 ? strtoull(argv[3], nullptr, 0) : 200000;

    fa_gen_params gp{}; gp.seed = 4; gp.n_keys = n_keys; gp.dist = FA_GEN_ZIPF; gp.zipf_s_milli = 1100; gp.t0_ns = 1000;
    // per-flow constant descriptors: the order-dependent fields (first MAC, last dscp ...) follow ARRIVAL order at the owner,
    // which across sources / CTAs is not stream order — like the reference's merge of per-CPU entries in CPU order
    gp.varying_desc = 0;
    std::vector<fa_flow_record> recs(n);
    CHECK(fa_gen_records(nullptr, &gp, 0, n, recs.data()) == FA_OK);          // host instance of the generator

    // reference result: one engine, one GPU
    fa_config cfg{}; cfg.abi_version = FA_ABI_VERSION; cfg.mode = FA_MODE_ACCOUNTER; cfg.max_entries = 2 * n_keys; cfg.max_batch = 1 << 18;
    fa_engine* one = nullptr;
    CHECK(fa_create(&cfg, &one) == FA_OK);
    size_t took = 0;
    CHECK(fa_ingest(one, recs.data(), n, &took) == FA_OK && took == n);
    std::vector<fa_flow_record> want(n_keys);
    size_t n_want = 0;
    CHECK(fa_evict(one, want.data(), nullptr, nullptr, nullptr, want.size(), &n_want) == FA_OK);
    fa_destroy(one);
    want.resize(n_want);

    for (int combine = 1; combine >= 0; combine--) {
        std::vector<int32_t> devs(n_gpus);
        for (int i = 0; i < n_gpus; i++) devs[i] = i;
        fa_config sc = cfg; sc.max_batch = 1 << 17; sc.reserved0 = combine ? 0u : 1u;
        fa_sharded* s = nullptr;
        CHECK(fa_sharded_create(&sc, devs.data(), (uint32_t)n_gpus, &s) == FA_OK);
        // three calls of uneven size: pageable memory, stream order == array order
        const size_t cut1 = n / 3 + 7, cut2 = 2 * n / 3 + 1;
        CHECK(fa_sharded_ingest(s, recs.data(), cut1) == FA_OK);
        CHECK(fa_sharded_ingest(s, recs.data() + cut1, cut2 - cut1) == FA_OK);
        CHECK(fa_sharded_ingest(s, recs.data() + cut2, n - cut2) == FA_OK);
        size_t live = 0;
        CHECK(fa_sharded_live_flows(s, &live) == FA_OK && live == n_want);
        std::vector<fa_flow_record> got(live);
        size_t n_got = 0;
        CHECK(fa_sharded_evict(s, got.data(), got.size(), &n_got) == FA_OK && n_got == n_want);
        fa_stats st{}; uint64_t nv = 0, ov = 0;
        CHECK(fa_sharded_get_stats(s, &st, &nv, &ov) == FA_OK);
        CHECK(ov == 0 && st.spills == 0 && st.records_ingested == n && st.flows_evicted == n_want);
        CHECK(n_gpus == 1 ? nv == 0 : nv > 0);
        std::sort(got.begin(), got.end(), rec_less);
        std::vector<fa_flow_record> w = want;
        std::sort(w.begin(), w.end(), rec_less);
        CHECK(memcmp(got.data(), w.data(), n_want * sizeof(fa_flow_record)) == 0);
        CHECK(fa_sharded_live_flows(s, &live) == FA_OK && live == 0);
        printf("sharded x%d %s: %zu records -> %zu flows bit-identical to one GPU; %llu records (%.1f MB) crossed NVLink\n", n_gpus,
               combine ? "with the local combiner" : "raw routing", n, n_want, (unsigned long long)nv, nv * 144 / 1e6);
        fa_sharded_destroy(s);
    }
    return 0;
}

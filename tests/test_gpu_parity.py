"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle — bit-exact."""
import numpy as np
import pytest

import oracle_lib as O
from common import assert_same_generations, gen_host, gpu_generations, oracle_generations
from test_oracle_goldens import K1, K2, K3

pytestmark = pytest.mark.gpu


def rec_view(a):
    return np.ascontiguousarray(a).view(O.REC_DTYPE).reshape(-1)


def by_key(raw):
    return {bytes(r.tobytes()[:40]): r for r in rec_view(raw)}


def test_golden_evict_max_entries_on_gpu():
    """pkg/flow/account_test.go:47-128 through the GPU Accounter mirror."""
    import netobserv_ebpf_agent_b200 as fa
    now = 1661272402 * 10**9
    acc = fa.Accounter(2, clock=lambda: now, mono_clock=lambda: 1000)
    acc.account(K1(bytes=123, packets=1, start=123, end=123, flags=1))
    acc.account(K2(bytes=456, packets=1, start=456, end=456, flags=1))
    acc.account(K1(bytes=321, packets=1, start=789, end=789, flags=1))
    assert acc.out == []                                            # requireNoEviction
    acc.account(K3(bytes=111, packets=1, start=888, end=888, flags=1))
    assert len(acc.out) == 1
    reason, recs, tnow, mono = acc.out[0]
    assert reason == "full" and len(recs) == 2
    got = by_key(recs)
    k1, k2 = got[K1().tobytes()[:40]], got[K2().tobytes()[:40]]
    assert (k1["bytes"], k1["packets"], k1["start"], k1["end"], k1["flags"]) == (444, 2, 123, 789, 1)
    assert (k2["bytes"], k2["packets"], k2["start"], k2["end"], k2["flags"]) == (456, 1, 456, 456, 1)
    assert fa.new_record_times(tnow, mono, int(k1["start"]), int(k1["end"])) == (now - 877, now - 211)
    assert fa.new_record_times(tnow, mono, int(k2["start"]), int(k2["end"])) == (now - 544, now - 544)
    assert acc.engine.live_flows() == 1                             # k3 stays
    acc.close()
    assert len(acc.out) == 2 and acc.out[1][0] == "closing" and len(acc.out[1][1]) == 1


def test_golden_evict_period_on_gpu():
    """pkg/flow/account_test.go:130-217."""
    import netobserv_ebpf_agent_b200 as fa
    acc = fa.Accounter(200)
    for ts in (123, 456, 789):
        acc.account(K1(bytes=10, packets=1, start=ts, end=ts, flags=1))
    acc.tick()
    for ts in (1123, 1456):
        acc.account(K1(bytes=10, packets=1, start=ts, end=ts, flags=1))
    acc.tick()
    acc.tick()                                                      # nothing more is evicted
    assert [o[0] for o in acc.out] == ["timeout", "timeout"]
    a, b = rec_view(acc.out[0][1])[0], rec_view(acc.out[1][1])[0]
    assert (a["bytes"], a["packets"], a["start"], a["end"], a["flags"]) == (30, 3, 123, 789, 1)
    assert (b["bytes"], b["packets"], b["start"], b["end"], b["flags"]) == (20, 2, 1123, 1456, 1)


def test_config1_accounter_replay():
    """BASELINE.json configs[0]: 10k records / 100 5-tuples, maxEntries 5000."""
    recs = gen_host(seed=1, n=10_000, n_keys=100)
    got, st = gpu_generations([recs], 5000)
    want = oracle_generations([recs], 5000)
    assert_same_generations(got, want)
    assert len(got) == 1 and len(got[0]) == 100
    assert int(rec_view(got[0])["packets"].sum()) == 10_000
    assert st["order_fixups"] == 0 and st["spills"] == 0


@pytest.mark.parametrize("n", [0, 1, 2, 31, 32, 33, 255, 256, 257, 511, 513, 4097])
def test_ragged_sizes(n):
    recs = gen_host(seed=3, n=n, n_keys=50)
    got, _ = gpu_generations([recs], 1000)
    assert_same_generations(got, oracle_generations([recs], 1000))


@pytest.mark.parametrize("dist,n_keys", [(0, 1000), (1, 1000), (0, 200_000), (1, 200_000)])
def test_stream_parity(dist, n_keys):
    recs = gen_host(seed=2, n=400_000, n_keys=n_keys, dist=dist)
    got, st = gpu_generations([recs], 1 << 20)
    assert_same_generations(got, oracle_generations([recs], 1 << 20))
    assert st["order_fixups"] == 0


def test_multi_batch_and_chunking():
    """Several ingest calls per eviction, max_batch smaller than a call (internal chunking)."""
    b = [gen_host(seed=5, n=70_001, n_keys=30_000, dist=1, first=i * 70_001) for i in range(4)]
    got, _ = gpu_generations(b, 1 << 16, max_batch=10_000)
    assert_same_generations(got, oracle_generations(b, 1 << 16))


def test_device_pointer_ingest():
    recs = gen_host(seed=6, n=123_457, n_keys=5_000, dist=1)
    got, st = gpu_generations([recs], 1 << 14, to_device=True)
    assert_same_generations(got, oracle_generations([recs], 1 << 14))
    assert st["h2d_bytes"] == 0


@pytest.mark.parametrize("n_keys,n", [(7, 5_000), (300, 60_000), (20_000, 150_000)])
def test_order_dependent_fields(n_keys, n):
    """Per-record random descriptors: eth/dscp/sampling last-non-zero, MACs first-non-zero, the rest
    from the first record (pkg/model/flow_content.go:45-59, pkg/flow/account.go:95) — across batches."""
    b = [gen_host(seed=7, n=n, n_keys=n_keys, dist=1, varying=1, first=i * n) for i in range(3)]
    got, st = gpu_generations(b, 1 << 16, max_batch=40_000)
    assert_same_generations(got, oracle_generations(b, 1 << 16))
    assert st["order_fixups"] > 0


def test_packets_u32_wrap_and_zero_timestamps():
    r = rec_view(gen_host(seed=8, n=1000, n_keys=3)).copy()
    r["packets"] = 0xFFFFFFF0                      # sums wrap mod 2^32 (flow_content.go:43)
    r["start"][::3] = 0                            # 0 == "unset" start (flow_content.go:36-38)
    r["end"][::5] = 0
    r["bytes"] = np.uint64(1) << np.uint64(62)     # u64 wrap
    got, _ = gpu_generations([r], 100)
    assert_same_generations(got, oracle_generations([r], 100))


def test_padding_bytes_ignored_and_zeroed():
    r = gen_host(seed=9, n=5000, n_keys=40).copy()
    rng = np.random.default_rng(9)
    r[:, 39] = rng.integers(0, 256, len(r))
    r[:, 106:108] = rng.integers(0, 256, (len(r), 2))
    r[:, 140:144] = rng.integers(0, 256, (len(r), 4))
    got, st = gpu_generations([r], 100)
    assert_same_generations(got, oracle_generations([r], 100))
    assert len(got[-1]) == 40 and st["order_fixups"] == 0
    assert not got[-1][:, 39].any() and not got[-1][:, 106:108].any() and not got[-1][:, 140:144].any()


@pytest.mark.parametrize("max_entries,n_keys,n", [(1, 5, 200), (2, 3, 100), (10, 50, 3000), (1000, 5000, 50_000),
                                                    (4096, 4097, 30_000)])
def test_full_cut_generations(max_entries, n_keys, n):
    """'full' evictions land on exactly the same record as in the reference (account.go:85-94)."""
    b = [gen_host(seed=10, n=n, n_keys=n_keys, dist=0, first=i * n) for i in range(2)]
    got, st = gpu_generations(b, max_entries, max_batch=7_000)
    want = oracle_generations(b, max_entries)
    assert_same_generations(got, want)
    assert st["full_cuts"] == len(want) - 1 and st["full_cuts"] > 0


def test_full_cut_with_order_dependent_fields():
    b = [gen_host(seed=11, n=20_000, n_keys=700, dist=1, varying=1)]
    got, _ = gpu_generations(b, 256, max_batch=3_000)
    assert_same_generations(got, oracle_generations(b, 256))


def test_determinism():
    recs = gen_host(seed=12, n=200_000, n_keys=10_000, dist=1, varying=1)
    a, _ = gpu_generations([recs], 1 << 15)
    for _ in range(3):
        b, _ = gpu_generations([recs], 1 << 15)
        assert_same_generations(b, a)


def test_device_generator_matches_host_generator():
    import torch
    import netobserv_ebpf_agent_b200 as fa
    for dist, varying in ((0, 0), (1, 0), (1, 1)):
        p = fa.GenParams(seed=13, n_keys=100_000, dist=dist, zipf_s_milli=1100, t0_ns=77, varying_desc=varying)
        host = fa.gen_records_host(p, 1_000_000_007, 50_001)
        with fa.FlowAggEngine(1000) as eng:
            t = torch.zeros(50_001 * 144, dtype=torch.uint8, device="cuda")
            eng.gen_records(p, 1_000_000_007, 50_001, t)
            eng.sync()
        assert np.array_equal(t.cpu().numpy().reshape(-1, 144), host)


def test_evict_capacity_error_and_empty_evict():
    import ctypes as C
    import netobserv_ebpf_agent_b200 as fa
    from netobserv_ebpf_agent_b200._lib import lib
    with fa.FlowAggEngine(1000) as eng:
        assert len(eng.evict()) == 0
        recs = gen_host(seed=14, n=500, n_keys=100)
        distinct = len(np.unique(recs[:, :40], axis=0))
        eng.ingest(recs)
        out = np.zeros((10, 144), dtype=np.uint8)
        got = C.c_size_t(0)
        rc = lib().fa_evict(eng._h, C.c_void_p(out.ctypes.data), None, None, None, 10, C.byref(got))
        assert rc == -7 and eng.live_flows() == distinct      # FA_E_2BIG, nothing deleted
        assert len(eng.evict()) == distinct and eng.live_flows() == 0


@pytest.mark.parametrize("n,shards", [(1, 2), (2047, 2), (2049, 3), (100_000, 8), (65_536, 16)])
def test_route_by_hash_matches_host_partition(n, shards):
    """K3: stable partition by owner_hash(key) % shards == the host twin (same order inside a shard)."""
    import torch
    import netobserv_ebpf_agent_b200 as fa
    from netobserv_ebpf_agent_b200.sharded import route_host
    recs = gen_host(seed=15, n=n, n_keys=5_000, dist=1)
    want, want_counts = route_host(recs, shards)
    with fa.FlowAggEngine(1000, max_batch=1 << 17) as eng:
        src = torch.from_numpy(np.ascontiguousarray(recs).reshape(-1).copy()).cuda()
        dst = torch.zeros_like(src)
        counts = eng.route(src, n, shards, dst)
        eng.sync()
        assert np.array_equal(counts.astype(np.int64), want_counts)
        assert np.array_equal(dst.cpu().numpy().reshape(-1, 144), want)


@pytest.mark.parametrize("varying", [0, 1])
def test_drain_active_partials_fold_to_the_same_flows(varying):
    """fa_drain_active: per-batch partial flow records (flows stay cached) re-folded by a second engine — the
    multi-GPU combiner path — give exactly the flows of the plain stream when descriptors are per-key constants;
    commutative fields are exact in every case."""
    import torch
    import netobserv_ebpf_agent_b200 as fa
    batches = [gen_host(seed=16, n=50_000, n_keys=8_000, dist=1, varying=varying, first=i * 50_000) for i in range(4)]
    with fa.FlowAggEngine(1 << 17, max_batch=60_000) as local, fa.FlowAggEngine(1 << 15) as owner:
        part = torch.zeros(60_000 * 144, dtype=torch.uint8, device="cuda")
        total_partials = 0
        for b in batches:
            local.ingest(b)
            k = local.drain_active(part, 60_000)
            assert 0 < k <= len(np.unique(b[:, :40], axis=0))
            total_partials += k
            owner.ingest(part.data_ptr(), k)
        assert local.drain_active(part, 60_000) == 0          # nothing touched since the last drain
        got = O.sort_records(owner.evict())
        assert local.live_flows() == len(got)                 # the combiner kept its flows cached
    want = oracle_generations(batches, 1 << 20)[0]
    g, w = rec_view(got), rec_view(want)
    for f in ("bytes", "packets", "flags", "start", "end"):
        assert np.array_equal(g[f], w[f]), f
    assert np.array_equal(got[:, :40], want[:, :40])
    if not varying:
        assert np.array_equal(got, want)
    assert total_partials < sum(len(b) for b in batches)

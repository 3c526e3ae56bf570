"""KERNEL_MAP mode (bpf/flows.c:76-143,222-288: lookup, update_existing_flow, add_observed_intf, insert with
BPF_NOEXIST, ring-buffer fallback) on the GPU against the oracle, through the C ABI.

The same streams also run through the host emulation of the kernel bodies (tests/test_kmap_emulation.py); this file
is the device-side gate (first green B200 run: round 2, both FA_KMAP_IMPL values)."""
import os

import numpy as np
import pytest

import oracle_lib as O
from common import gen_host
from test_kmap_emulation import messy_stream

pytestmark = pytest.mark.gpu


def test_kernel_map_mode_refuses_accounter_only_flags():
    import netobserv_ebpf_agent_b200 as fa
    with pytest.raises(fa.FlowAggError) as ei:
        fa.FlowAggEngine(100, mode=fa.FA_MODE_KERNEL_MAP, flags=fa.FA_F_NO_FULL_CUT)
    assert ei.value.code == -22 and "KERNEL_MAP" in str(ei.value)


@pytest.fixture(params=["2", "1"], autouse=True)
def kmap_impl(request, monkeypatch):
    """FA_KMAP_IMPL: 2 = one streaming pass + per-flow finalisation (default), 1 = seven per-record passes."""
    monkeypatch.setenv("FA_KMAP_IMPL", request.param)
    return request.param


def check(recs, max_entries, max_batch, ringbuf=True, evict_every=None):
    import netobserv_ebpf_agent_b200 as fa
    km = O.KernelMap(max_entries, ringbuf_fallback=ringbuf)
    b = O.as_bytes(recs)
    n = b.size // O.REC
    step = evict_every or n
    with fa.FlowAggEngine(max_entries, mode=fa.FA_MODE_KERNEL_MAP, max_batch=max_batch,
                          flags=fa.FA_F_RINGBUF_FALLBACK if ringbuf else 0) as eng:
        for lo in range(0, n, step):
            part = b[lo * O.REC: (lo + step) * O.REC]
            km.packets(part)
            eng.ingest(part)
            want, got = O.sort_records(km.evict()), O.sort_records(eng.evict())
            assert want.shape == got.shape
            if not np.array_equal(want, got):
                w, g = want.view(O.REC_DTYPE).reshape(-1), got.view(O.REC_DTYPE).reshape(-1)
                bad = np.nonzero((want != got).any(axis=1))[0][0]
                diff = [f for f in O.REC_DTYPE.names if not np.array_equal(w[bad][f], g[bad][f])]
                raise AssertionError(f"flow {bad}: fields {diff}: want {[w[bad][f] for f in diff]} got {[g[bad][f] for f in diff]}")
            n_sp = km.spilled()
            if ringbuf:
                gs = eng.read_spilled()
                assert len(gs) == n_sp
                if n_sp:
                    ws = km.spilled_records(n_sp)
                    assert np.array_equal(ws[np.lexsort(ws.T[::-1])], gs[np.lexsort(gs.T[::-1])])
        st = eng.stats()
    assert st["observed_intf_missed"] == km.intf_missed and st["hashmap_fail_create"] == km.fail_create
    assert st["spills"] == 0 and st["ringbuf_dropped"] == 0


@pytest.mark.parametrize("dist,n_keys", [(0, 500), (1, 20_000)])
def test_generator_streams(dist, n_keys):
    recs = gen_host(seed=60, n=200_000, n_keys=n_keys, dist=dist)
    check(recs, 1 << 20, 1 << 18)
    check(recs, 1 << 20, 30_011)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_messy_stream_matches_the_sequential_map_update(seed):
    recs = messy_stream(seed, 60_000, 300)
    check(recs, 1 << 12, 60_000)
    check(recs, 1 << 12, 997)


def test_many_interfaces_fill_the_observed_list():
    recs = messy_stream(21, 40_000, 40, n_ifaces=12)
    check(recs, 1 << 10, 40_000)
    check(recs, 1 << 10, 512)
    check(messy_stream(22, 5_000, 1, n_ifaces=30), 16, 5_000)


def test_full_map_spills_to_the_ring_buffer_or_counts():
    recs = messy_stream(31, 30_000, 2_000, tls=False)
    check(recs, 500, 30_000, ringbuf=True)
    check(recs, 500, 4_096, ringbuf=True)
    check(recs, 500, 4_096, ringbuf=False)


def test_eviction_between_batches_and_large_batch():
    check(messy_stream(41, 50_000, 800), 1 << 11, 2_048, evict_every=10_000)
    check(messy_stream(42, 2_000_000, 100_000, n_ifaces=3), 1 << 18, 1 << 20)


def test_committed_fixture():
    import netobserv_ebpf_agent_b200 as fa
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmap", "kmap_messy_seed11.npz"))
    with fa.FlowAggEngine(int(z["max_entries"]), mode=fa.FA_MODE_KERNEL_MAP, max_batch=8_000, flags=fa.FA_F_RINGBUF_FALLBACK) as eng:
        eng.ingest(z["records"])
        assert np.array_equal(O.sort_records(eng.evict()), z["flows"])
        sp = eng.read_spilled()
        assert np.array_equal(sp[np.lexsort(sp.T[::-1])], z["spilled"])
        st = eng.stats()
    assert [st["observed_intf_missed"], st["hashmap_fail_create"]] == [int(x) for x in z["counters"]]


def test_kernel_map_base_with_feature_maps():
    """aggregated_flows as the base + the per-CPU feature maps, merged by LookupAndDeleteMap (tracer.go:1063-1157)."""
    import netobserv_ebpf_agent_b200 as fa
    from test_gpu_features import compare, keys_of, make_add, make_dns
    rng = np.random.default_rng(81)
    pk = messy_stream(81, 60_000, 2_000, n_ifaces=5)
    keys = np.unique(pk[:, :40], axis=0)
    allk = np.concatenate([keys, keys_of(82, 300)])
    add, dns = make_add(rng, allk, 20_000), make_dns(rng, allk, 20_000)
    om = O.FlowMap()
    with fa.FlowAggEngine(1 << 15, mode=fa.FA_MODE_KERNEL_MAP, flags=fa.FA_F_ENABLE_RTT | fa.FA_F_ENABLE_DNS, max_batch=16_384) as eng:
        eng.ingest_additional(add); eng.ingest(pk[:30_000]); eng.ingest_dns(dns); eng.ingest(pk[30_000:])
        om.fold_additional(add); missed = om.packets_kmap(pk); om.fold_dns(dns)
        compare(eng, om)
        assert eng.stats()["observed_intf_missed"] == missed and eng.live_flows() == 0

"""The whole C ABI on the CPU: csrc/engine.cu compiled unchanged against a test double of the CUDA runtime, with the
kernels running on tests/emul/simt.h (tests/emul/engine_emul.cpp).  What it checks is HOST logic together with the
kernels' logic — chunking and staging, the exact "full" cut, KERNEL_MAP plumbing and its fallback ring, feature and
sketch paths, error codes — through the same ctypes binding and the same helpers as the GPU tests, at sizes where a
kernel launch (a few hundred to two thousand OS threads) is affordable.  It is test infrastructure: the package's
loader cannot reach the emulated library (the handle is swapped in here, explicitly), the product has no CPU path,
and the GPU tests remain the gate for everything the emulation cannot see (memory model, launch geometry, speed)."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as O
from common import assert_same_generations, gen_host, gpu_generations, oracle_generations

# a wedged emulation (it is thousands of OS threads) must not hang the suite: pytest-timeout, if installed
pytestmark = pytest.mark.timeout(900)


def test_chunked_host_ingest_and_refold(engine_emul):
    recs = gen_host(seed=7, n=5_000, n_keys=60, dist=1, varying=1)
    got, st = gpu_generations([recs[:3_000], recs[3_000:]], 1 << 10, max_batch=1_024)
    assert_same_generations(got, oracle_generations([recs], 1 << 10))
    assert st["order_fixups"] > 0 and st["records_ingested"] == 5_000 and st["h2d_bytes"] == 5_000 * 144


@pytest.mark.parametrize("impl", ["1", "2"])
def test_kernel_map_mode_through_the_c_abi(engine_emul, monkeypatch, impl):
    """Both implementations of the map update (FA_KMAP_IMPL: 1 = seven per-record passes, 2 = per-flow finalisation)."""
    from test_gpu_kernel_map import check
    from test_kmap_emulation import messy_stream
    monkeypatch.setenv("FA_KMAP_IMPL", impl)
    check(messy_stream(51, 2_000, 120, n_ifaces=9), 1 << 9, 1_024)                       # two batches
    check(messy_stream(52, 2_000, 400, tls=False), 150, 1_024, ringbuf=True)             # full map -> fallback ring
    check(messy_stream(52, 1_500, 400, tls=False), 150, 1_024, ringbuf=False)            # ... or the counter
    check(messy_stream(53, 1_600, 100), 1 << 8, 1_024, evict_every=800)


def test_error_paths(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    with pytest.raises(fa.FlowAggError) as ei:
        fa.FlowAggEngine(0)
    assert ei.value.code == -22
    with pytest.raises(fa.FlowAggError):
        fa.FlowAggEngine(100, mode=fa.FA_MODE_KERNEL_MAP, flags=fa.FA_F_ENABLE_SKETCH)
    with pytest.raises(fa.FlowAggError):
        fa.FlowAggEngine(100, flags=fa.FA_F_RINGBUF_FALLBACK)
    with fa.FlowAggEngine(100) as eng:
        with pytest.raises(fa.FlowAggError):
            eng.read_spilled()
        assert eng.live_flows() == 0 and len(eng.evict()) == 0


def test_small_cache_windows_keep_the_cut_exact_and_the_prepass_linear(engine_emul):
    """A cache much smaller than the batch (the reference's default is 5000 flows): the cut is searched in windows
    proportional to the room left, every generation still ends at the exact record, and the number of launches grows
    with the number of generations, not with generations x batch size."""
    recs = gen_host(seed=71, n=2_400, n_keys=200, dist=1)
    want = oracle_generations([recs], 60)
    got, st = gpu_generations([recs], 60, max_batch=2_400)             # one host chunk, many cuts inside it
    assert_same_generations(got, want)
    gens = len(want) - 1
    assert gens > 8 and st["full_cuts"] == gens
    # per generation: <= 2 pre-passes (3 kernels each) + 2 folds (3 each) + the eviction; windows without a cut add folds
    assert st["kernel_launches"] <= 16 * gens + 64, (st["kernel_launches"], gens)


def test_device_resident_input_with_cuts(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    recs = gen_host(seed=72, n=900, n_keys=120, dist=1)
    want = oracle_generations([recs], 40)
    gens = []
    with fa.FlowAggEngine(40, max_batch=1_024) as eng:
        p = ctypes.c_void_p()
        assert engine_emul.fa_device_alloc(eng._h, recs.size, ctypes.byref(p)) == 0
        ctypes.memmove(p.value, np.ascontiguousarray(recs).ctypes.data, recs.size)
        done, n = 0, len(recs)
        while done < n:
            rc, took = eng.ingest(p.value + done * 144, n - done)
            done += took
            if rc == fa.FA_FULL:
                gens.append(O.sort_records(eng.evict()))
        gens.append(O.sort_records(eng.evict()))
        engine_emul.fa_device_free(eng._h, p)
    assert_same_generations(gens, want)


@pytest.mark.parametrize("max_entries,n_keys,n", [(1, 5, 8), (2, 3, 14)])
def test_degenerate_cache_sizes(engine_emul, max_entries, n_keys, n):
    """maxEntries 1 and 2 (tests/test_gpu_parity.py::test_full_cut_generations at emulation scale): room 0 / 1 windows."""
    b = [gen_host(seed=10, n=n, n_keys=n_keys, dist=0, first=i * n) for i in range(2)]
    got, st = gpu_generations(b, max_entries, max_batch=64)
    want = oracle_generations(b, max_entries)
    assert_same_generations(got, want)
    assert st["full_cuts"] == len(want) - 1 and st["full_cuts"] > 0


@pytest.mark.parametrize("impl", ["2", "1"])
def test_kernel_map_base_with_feature_maps(engine_emul, monkeypatch, impl):
    """The reference's real deployment: aggregated_flows as the base + the per-CPU feature maps, merged by
    LookupAndDeleteMap (tracer.go:1063-1157).  Flows first seen through a feature sample get their kernel-map entry
    from the first packet that follows; flows never seen by flow_monitor come out with an empty base."""
    import netobserv_ebpf_agent_b200 as fa
    from test_gpu_features import compare, keys_of, make_add, make_dns
    from test_kmap_emulation import messy_stream
    monkeypatch.setenv("FA_KMAP_IMPL", impl)
    rng = np.random.default_rng(81)
    pk = messy_stream(81, 3_000, 150, n_ifaces=5)
    keys = np.unique(pk[:, :40], axis=0)
    allk = np.concatenate([keys, keys_of(82, 30)])
    add, dns = make_add(rng, allk, 900), make_dns(rng, allk, 900)
    om = O.FlowMap()
    with fa.FlowAggEngine(1 << 10, mode=fa.FA_MODE_KERNEL_MAP, flags=fa.FA_F_ENABLE_RTT | fa.FA_F_ENABLE_DNS, max_batch=1_024) as eng:
        eng.ingest_additional(add); eng.ingest(pk[:1_500]); eng.ingest_dns(dns); eng.ingest(pk[1_500:])
        om.fold_additional(add); missed = om.packets_kmap(pk); om.fold_dns(dns)
        compare(eng, om)
        assert eng.stats()["observed_intf_missed"] == missed
        assert eng.live_flows() == 0


def test_route_peer_delivers_every_record_to_its_owner_once(engine_emul):
    """K3 fused with the exchange (fa_route_peer) on the emulation, where "peer memory" is host memory: every record
    lands exactly once in the buffer of owner_hash(key) % n_shards, records of one tile keep their order, the device-side
    count bounds the work, the remote-record counter and the overflow counter are exact."""
    import netobserv_ebpf_agent_b200 as fa
    from netobserv_ebpf_agent_b200._lib import check, lib
    from netobserv_ebpf_agent_b200.sharded import owner_of
    n_max, n, n_shards, me = 3_000, 2_777, 3, 1
    recs = np.ascontiguousarray(gen_host(seed=5, n=n_max, n_keys=400, dist=1))
    own = owner_of(recs[:n, :40], n_shards)
    for cap in (n_max, 700):                                   # roomy buffers, then buffers that overflow
        bufs = [np.zeros((cap, 144), dtype=np.uint8) for _ in range(n_shards)]
        cnts = [np.zeros(1, dtype=np.uint64) for _ in range(n_shards)]
        overflow, n_dev = np.zeros(2, dtype=np.uint64), np.array([n], dtype=np.uint64)
        pb, pc = (ctypes.c_void_p * 16)(), (ctypes.c_void_p * 16)()
        for s in range(n_shards):
            pb[s], pc[s] = bufs[s].ctypes.data, cnts[s].ctypes.data
        with fa.FlowAggEngine(1 << 12, max_batch=4096) as eng:
            check(lib().fa_route_peer(eng._h, ctypes.c_void_p(recs.ctypes.data), ctypes.c_void_p(n_dev.ctypes.data), n_max, n_shards, me,
                                      pb, pc, cap, ctypes.c_void_p(overflow.ctypes.data)))
            eng.sync()
        lost = 0
        for s in range(n_shards):
            want = recs[:n][own == s]
            assert int(cnts[s][0]) == len(want)                # reservations count every record, stored or not
            got = bufs[s][: min(len(want), cap)]
            lost += max(len(want) - cap, 0)
            if len(want) <= cap:
                # tiles may be reserved in any order; inside a tile the order is kept -> compare as sorted multisets and per tile
                assert np.array_equal(got[np.lexsort(got.T[::-1])], want[np.lexsort(want.T[::-1])])
        assert int(overflow[0]) == lost and (lost > 0) == (cap < n_max)
        assert int(overflow[1]) == int((own != me).sum())

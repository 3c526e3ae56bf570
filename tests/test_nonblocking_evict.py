"""FA_F_NONBLOCKING_EVICT: fa_evict swaps in an empty table and scans the retired one on its own stream while another
thread keeps calling fa_ingest — the Accounter hands its map to the evictor and goes on (pkg/flow/account.go:67-68,86-87).
Whatever the interleaving, every record ends up in exactly one eviction: the union of all evictions, re-folded,
equals the oracle's fold of the whole stream (AccumulateBase is associative over generations for these streams)."""
import threading

import numpy as np
import pytest

import oracle_lib as O
from common import gen_host


def refold(parts):
    acc = O.Accounter(1 << 22)
    for p in parts:
        if len(p):
            acc.account(p)
    out = O.sort_records(acc.evict())
    acc.close()
    return out


def check_sequential(n, n_keys, max_entries, max_batch):
    import netobserv_ebpf_agent_b200 as fa
    recs = gen_host(seed=41, n=n, n_keys=n_keys, dist=1)
    cuts = [0, n // 3, n // 2, n]
    with fa.FlowAggEngine(max_entries, flags=fa.FA_F_NONBLOCKING_EVICT, max_batch=max_batch) as eng:
        for a, b in zip(cuts[:-1], cuts[1:]):
            rc, took = eng.ingest(recs[a:b]); assert rc == 0 and took == b - a
            got = O.sort_records(eng.evict())
            acc = O.Accounter(max_entries); acc.account(recs[a:b]); want = O.sort_records(acc.evict()); acc.close()
            assert np.array_equal(got, want)                       # every generation on its own is exact
            assert eng.live_flows() == 0
        assert len(eng.evict()) == 0


def check_concurrent(n, n_keys, max_entries, max_batch, chunk):
    import netobserv_ebpf_agent_b200 as fa
    recs = gen_host(seed=42, n=n, n_keys=n_keys, dist=1)
    evicted, stop = [], threading.Event()
    with fa.FlowAggEngine(max_entries, flags=fa.FA_F_NONBLOCKING_EVICT, max_batch=max_batch) as eng:
        def evictor():
            while not stop.is_set():
                evicted.append(eng.evict(cap=max_entries).copy())
        t = threading.Thread(target=evictor)
        t.start()
        for lo in range(0, n, chunk):
            rc, took = eng.ingest(recs[lo:lo + chunk])
            assert rc == 0 and took == min(chunk, n - lo)
        stop.set(); t.join()
        evicted.append(eng.evict().copy())
        st = eng.stats()
    assert st["records_ingested"] == n and st["flows_evicted"] == sum(len(p) for p in evicted)
    assert sum(1 for p in evicted if len(p)) >= 2                 # the evictor really ran during the ingest
    assert np.array_equal(refold(evicted), refold([recs]))


def test_generations_are_exact_on_the_emulation(engine_emul):
    check_sequential(n=6_000, n_keys=500, max_entries=1 << 11, max_batch=2_048)


def test_ingest_while_evicting_on_the_emulation(engine_emul):
    check_concurrent(n=12_000, n_keys=400, max_entries=1 << 11, max_batch=1_024, chunk=1_000)


def test_flag_is_refused_with_feature_folds(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    with pytest.raises(fa.FlowAggError) as ei:
        fa.FlowAggEngine(64, flags=fa.FA_F_NONBLOCKING_EVICT | fa.FA_F_ENABLE_DNS)
    assert ei.value.code == -22


@pytest.mark.gpu
def test_generations_are_exact_gpu():
    check_sequential(n=600_000, n_keys=80_000, max_entries=1 << 18, max_batch=1 << 16)


@pytest.mark.gpu
def test_ingest_while_evicting_gpu():
    check_concurrent(n=4_000_000, n_keys=300_000, max_entries=1 << 20, max_batch=1 << 17, chunk=1 << 16)

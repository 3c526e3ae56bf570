"""KERNEL_MAP mode (bpf/flows.c:98-143,222-288 hit/miss semantics) on the GPU against the oracle.

Round 1 ships the mode in the oracle only: fa_create(mode=FA_MODE_KERNEL_MAP) refuses loudly and these tests skip.
They are the acceptance tests of the GPU implementation (next round): same streams, bit-exact flows, identical
spill list and counters."""
import numpy as np
import pytest

import oracle_lib as O
from common import gen_host

pytestmark = pytest.mark.gpu


def _engine(max_entries, **kw):
    import netobserv_ebpf_agent_b200 as fa
    try:
        return fa.FlowAggEngine(max_entries, mode=fa.FA_MODE_KERNEL_MAP, **kw)
    except fa.FlowAggError as e:
        if "mode" in str(e):
            pytest.skip("KERNEL_MAP mode is not implemented on the GPU yet (oracle only)")
        raise


def test_kernel_map_mode_is_refused_loudly_until_implemented():
    import netobserv_ebpf_agent_b200 as fa
    try:
        eng = fa.FlowAggEngine(100, mode=fa.FA_MODE_KERNEL_MAP)
    except fa.FlowAggError as e:
        assert e.code == -22 and "mode" in str(e)          # no silent fallback to ACCOUNTER semantics
    else:
        eng.close()


@pytest.mark.parametrize("dist,n_keys", [(0, 500), (1, 20_000)])
def test_kernel_map_parity_on_generator_streams(dist, n_keys):
    recs = gen_host(seed=60, n=200_000, n_keys=n_keys, dist=dist)
    with _engine(1 << 20) as eng:
        eng.ingest(recs)
        got = O.sort_records(eng.evict())
    km = O.KernelMap(1 << 20)
    km.packets(recs)
    assert np.array_equal(got, O.sort_records(km.evict()))


def test_kernel_map_dedup_rule_parity():
    """Same packets seen on several interfaces: only the first-seen interface counts (flows.c:104-131)."""
    rng = np.random.default_rng(61)
    recs = gen_host(seed=61, n=60_000, n_keys=2_000, dist=1).copy()
    r = recs.view(O.REC_DTYPE).reshape(-1)
    r["if_index"] = rng.integers(0, 9, len(r))             # 0 = unknown interface
    r["direction"] = rng.integers(0, 2, len(r))
    with _engine(1 << 16) as eng:
        eng.ingest(recs)
        got = O.sort_records(eng.evict())
        st = eng.stats()
    km = O.KernelMap(1 << 16)
    km.packets(recs)
    assert np.array_equal(got, O.sort_records(km.evict()))
    assert st["observed_intf_missed"] == km.intf_missed

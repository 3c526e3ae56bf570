"""KERNEL_MAP semantics (bpf/flows.c:76-143,222-288) of the CPU oracle.  The reference has no unit test for this
code (it needs a kernel), so these are source-pinned checks: hand-computed expectations for the in-kernel
de-duplication rule and the observed-interface list, the ring-buffer spill on a full map, and the self-test
SURVEY.md Appendix A.2 asks for: on the generator's streams (monotone timestamps, per-key-constant L2 / ifindex /
dscp / sampling) KERNEL_MAP and ACCOUNTER semantics give bit-identical flows."""
import numpy as np

import oracle_lib as O
from common import gen_host
from test_oracle_goldens import K1, K2, K3


def ev(mk, ts, length, ifindex, direction=0, flags=0x10, dscp=0, sampling=0, **kw):
    """One packet event in the ring-buffer record shape (flows.c:228-245)."""
    return mk(start=ts, end=ts, bytes=length, packets=1, flags=flags, if_index=ifindex, direction=direction,
              dscp=dscp, sampling=sampling, eth=0x0800, **kw)


def test_kernel_map_equals_accounter_on_generator_streams():
    for dist, n_keys in ((0, 500), (1, 20_000)):
        recs = gen_host(seed=60, n=200_000, n_keys=n_keys, dist=dist)
        km = O.KernelMap(1 << 20)
        km.packets(recs)
        acc = O.Accounter(1 << 20)
        acc.account(recs)
        a, b = O.sort_records(km.evict()), O.sort_records(acc.evict())
        assert np.array_equal(a, b)
        assert km.spilled() == 0 and km.fail_create == 0


def test_in_kernel_dedup_and_observed_interfaces():
    """update_existing_flow: only packets from the first-seen interface are counted; other interfaces are listed
    (flows.c:104-131, add_observed_intf :76-96)."""
    km = O.KernelMap(100)
    seq = [ev(K1, 100, 60, ifindex=3, direction=0, flags=0x02, dscp=8, sampling=1),
           ev(K1, 110, 70, ifindex=5, direction=1, flags=0x10),          # same packet seen on another interface
           ev(K1, 120, 80, ifindex=3, direction=0, flags=0x10, dscp=0, sampling=0),
           ev(K1, 130, 90, ifindex=5, direction=0, flags=0x01),          # other direction on intf 5 -> BOTH (3)
           ev(K1, 140, 50, ifindex=0, flags=0x04),                       # ifindex 0 and not first-seen: ignored entirely
           ev(K1, 150, 40, ifindex=7, direction=1, flags=0x08)]
    for r in seq:
        km.packets(r)
    f = km.evict().view(O.REC_DTYPE).reshape(-1)[0]
    assert (f["packets"], f["bytes"]) == (2, 140)                       # 60 + 80 only
    assert (f["start"], f["end"]) == (100, 150)                          # end = last writer (flows.c:107,128)
    assert f["flags"] == (0x02 | 0x10 | 0x01 | 0x08)                     # flags of ifindex-0 packet not merged
    assert (f["dscp"], f["sampling"]) == (0, 0)                          # overwritten even with zero (:109-110)
    assert (f["if_index"], f["direction"]) == (3, 0)
    assert f["nb_obs"] == 2 and list(f["obs_intf"][:2]) == [5, 7] and list(f["obs_dir"][:2]) == [3, 1]


def test_observed_interface_capacity_and_counter():
    km = O.KernelMap(100)
    km.packets(ev(K1, 1, 10, ifindex=1, proto=6))
    for i in range(2, 10):                                               # 8 other interfaces, room for 6
        km.packets(ev(K1, i, 10, ifindex=i, proto=6))
    f = km.evict().view(O.REC_DTYPE).reshape(-1)[0]
    assert f["nb_obs"] == 6 and list(f["obs_intf"]) == [2, 3, 4, 5, 6, 7]
    assert km.intf_missed == 2                                           # OBSERVED_INTF_MISSED (flows.c:134-142)
    assert f["packets"] == 1


def test_tls_merge_rules():
    km = O.KernelMap(100)
    km.packets(ev(K1, 1, 10, ifindex=1, ssl_version=0, tls_types=0x01))
    km.packets(ev(K1, 2, 10, ifindex=1, ssl_version=0x0303, cipher=0x1301, tls_types=0x02))   # server hello
    km.packets(ev(K1, 3, 10, ifindex=1, ssl_version=0x0304, cipher=0x1302, tls_types=0x01))   # mismatching client hello
    f = km.evict().view(O.REC_DTYPE).reshape(-1)[0]
    assert f["ssl_version"] == 0x0303 and f["misc"] == 1                 # first non-zero wins, mismatch flagged (:111-118)
    assert f["cipher"] == 0x1301 and f["tls_types"] == 0x03              # cipher only from SERVER_HELLO (:119-121)


def test_full_map_spills_to_ringbuffer_then_accounter():
    """Map full -> E2BIG -> single-packet record with errno=7 through the ring buffer (flows.c:262-279), which the
    userspace Accounter re-aggregates (tracer_ringbuf.go:124, account.go:82-96)."""
    km = O.KernelMap(2, ringbuf_fallback=True)
    km.packets(ev(K1, 1, 10, ifindex=1)); km.packets(ev(K2, 2, 20, ifindex=1))
    km.packets(ev(K3, 3, 30, ifindex=1)); km.packets(ev(K3, 4, 40, ifindex=1)); km.packets(ev(K1, 5, 50, ifindex=1))
    assert len(km) == 2 and km.spilled() == 2
    sp = km.spilled_records(10)
    s = sp.view(O.REC_DTYPE).reshape(-1)
    assert list(s["errno"]) == [7, 7] and list(s["bytes"]) == [30, 40] and list(s["packets"]) == [1, 1]
    acc = O.Accounter(100)
    acc.account(sp)
    a = acc.evict().view(O.REC_DTYPE).reshape(-1)
    assert len(a) == 1 and (a[0]["bytes"], a[0]["packets"], a[0]["start"], a[0]["end"], a[0]["errno"]) == (70, 2, 3, 4, 7)
    m = {bytes(r.tobytes()[:40]): r for r in km.evict().view(O.REC_DTYPE).reshape(-1)}
    assert m[K1().tobytes()[:40]]["bytes"] == 60 and m[K2().tobytes()[:40]]["bytes"] == 20
    km2 = O.KernelMap(1, ringbuf_fallback=False)
    km2.packets(ev(K1, 1, 10, ifindex=1)); km2.packets(ev(K2, 2, 20, ifindex=1))
    assert km2.fail_create == 1 and km2.spilled() == 0                   # HASHMAP_FAIL_CREATE_FLOW (flows.c:285)


def test_kernel_map_fixture_is_reproduced():
    """tests/golden/kmap/kmap_messy_seed11.npz (made by make_kmap_golden.py): stream and oracle output are frozen."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden", "kmap"))
    from make_kmap_golden import run
    from test_kmap_emulation import messy_stream
    z = np.load(os.path.join(here, "golden", "kmap", "kmap_messy_seed11.npz"))
    assert np.array_equal(messy_stream(11, 8_000, 400, n_ifaces=9), z["records"])
    flows, spilled, counters = run(z["records"], int(z["max_entries"]))
    assert np.array_equal(flows, z["flows"]) and np.array_equal(spilled, z["spilled"]) and np.array_equal(counters, z["counters"])
    assert len(z["flows"]) == 300 and len(z["spilled"]) > 0 and int(z["counters"][0]) > 0      # non-trivial: full map, spills, misses

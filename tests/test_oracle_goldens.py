"""Pins the CPU oracle against the reference's own known-answer tests.

Each test names the reference test it transcribes (values, not code):
  pkg/flow/account_test.go, pkg/model/record_test.go, pkg/model/flow_content_test.go.
"""
import ctypes as C

import numpy as np

import oracle_lib as O

V4 = [0] * 10 + [0xFF, 0xFF]
SRC1 = V4 + [0x12, 0x34, 0x56, 0x78]
SRC2 = V4 + [0xAA, 0xBB, 0xCC, 0xDD]
DST1 = V4 + [0x43, 0x21, 0x00, 0xFF]
DST2 = V4 + [0x11, 0x22, 0x33, 0x44]


def mkrec(src, dst, sport, dport, **m):
    r = np.zeros(1, dtype=O.REC_DTYPE)
    r["src_ip"][0] = src
    r["dst_ip"][0] = dst
    r["src_port"], r["dst_port"] = sport, dport
    for k, v in m.items():
        r[k] = v
    return r


# account_test.go:28-45
def K1(**m): return mkrec(SRC1, DST1, 333, 8080, **m)
def K2(**m): return mkrec(SRC2, DST1, 12, 8080, **m)
def K3(**m): return mkrec(SRC1, DST2, 333, 443, **m)


def by_key(raw):
    recs = np.ascontiguousarray(raw).view(O.REC_DTYPE).reshape(-1)
    return {bytes(r.tobytes()[:40]): r for r in recs}


def test_evict_max_entries():
    """pkg/flow/account_test.go:47-128 TestEvict_MaxEntries."""
    acc = O.Accounter(2)
    acc.account(K1(bytes=123, packets=1, start=123, end=123, flags=1))
    acc.account(K2(bytes=456, packets=1, start=456, end=456, flags=1))
    acc.account(K1(bytes=321, packets=1, start=789, end=789, flags=1))
    assert acc.pending() == 0                       # requireNoEviction
    acc.account(K3(bytes=111, packets=1, start=888, end=888, flags=1))
    assert acc.pending() == 1
    got = by_key(acc.pop_generation())
    assert len(got) == 2
    k1 = got[K1().tobytes()[:40]]
    k2 = got[K2().tobytes()[:40]]
    assert (k1["bytes"], k1["packets"], k1["start"], k1["end"], k1["flags"]) == (444, 2, 123, 789, 1)
    assert (k2["bytes"], k2["packets"], k2["start"], k2["end"], k2["flags"]) == (456, 1, 456, 456, 1)
    # everything else in the evicted metrics is zero (assert.Equal on the whole struct)
    for r in (k1, k2):
        z = r.copy()
        for f in ("bytes", "packets", "start", "end", "flags"):
            z[f] = 0
        assert not z.tobytes()[40:].strip(b"\0")
    assert acc.pending() == 0 and len(acc) == 1     # k3 stays in the new table
    # wall-clock conversion: now - (1000 - ts) ns  (account_test.go:112,124)
    now = 1661272402 * 10**9                        # 2022-08-23T16:33:22Z
    tfs, tfe = C.c_uint64(), C.c_uint64()
    O.lib().oracle_new_record_times(now, 1000, 123, 789, C.byref(tfs), C.byref(tfe))
    assert tfs.value == now - 877 and tfe.value == now - 211
    O.lib().oracle_new_record_times(now, 1000, 456, 456, C.byref(tfs), C.byref(tfe))
    assert tfs.value == now - 544 and tfe.value == now - 544


def test_evict_period():
    """pkg/flow/account_test.go:130-217 TestEvict_Period (the timer is the caller's evict())."""
    acc = O.Accounter(200)
    for ts in (123, 456, 789):
        acc.account(K1(bytes=10, packets=1, start=ts, end=ts, flags=1))
    r = acc.evict().view(O.REC_DTYPE).reshape(-1)
    assert len(r) == 1
    assert (r[0]["bytes"], r[0]["packets"], r[0]["start"], r[0]["end"], r[0]["flags"]) == (30, 3, 123, 789, 1)
    for ts in (1123, 1456):
        acc.account(K1(bytes=10, packets=1, start=ts, end=ts, flags=1))
    r = acc.evict().view(O.REC_DTYPE).reshape(-1)
    assert len(r) == 1
    assert (r[0]["bytes"], r[0]["packets"], r[0]["start"], r[0]["end"], r[0]["flags"]) == (20, 2, 1123, 1456, 1)
    assert len(acc.evict()) == 0                    # "no more flows are evicted"
    now = 1661272402 * 10**9
    tfs, tfe = C.c_uint64(), C.c_uint64()
    O.lib().oracle_new_record_times(now, 1000, 1123, 1456, C.byref(tfs), C.byref(tfe))
    assert tfs.value == now - 1000 + 1123 and tfe.value == now - 1000 + 1456   # mono > monoNow wraps correctly


GOLDEN_144 = bytes([
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff, 0x06, 0x07, 0x08, 0x09,
    0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0xff, 0xff, 0x0a, 0x0b, 0x0c, 0x0d,
    0x0e, 0x0f, 0x10, 0x11, 0x12, 0x00, 0x00, 0x00,
    0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19, 0x1a,
    0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19, 0x1a,
    0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 0x19, 0x1a,
    0x06, 0x07, 0x08, 0x09, 0x01, 0x02, 0x13, 0x14,
    0x04, 0x05, 0x06, 0x07, 0x08, 0x09, 0x0a, 0x0b, 0x0c, 0x0d, 0x0e, 0x0f,
    0x13, 0x14, 0x15, 0x16, 0x00, 0x00, 0x00, 0x00, 0x02, 0x00, 0x00, 0x00,
    0x03, 0x33, 0x60, 0x02, 0x01, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00, 0x00,
    0x07, 0, 0, 0, 0x08, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0,
    0x03, 0x03, 0x00, 0x00, 0x00, 0x00, 0x21, 0x00, 0x00, 0x00, 0x00, 0x00,
])


def test_record_binary_encoding():
    """pkg/model/record_test.go:19-102 TestRecordBinaryEncoding: field offsets of the 144-byte record."""
    assert len(GOLDEN_144) == 144
    r = np.frombuffer(O.read_from(np.frombuffer(GOLDEN_144, dtype=np.uint8)).tobytes(), dtype=O.REC_DTYPE)[0]
    assert bytes(r["src_ip"]) == bytes(V4 + [6, 7, 8, 9])
    assert bytes(r["dst_ip"]) == bytes(V4 + [10, 11, 12, 13])
    assert r["src_port"] == 0x0F0E and r["dst_port"] == 0x1110 and r["proto"] == 0x12
    assert r["icmp_type"] == 0 and r["icmp_code"] == 0
    assert r["start"] == r["end"] == r["bytes"] == 0x1A19181716151413
    assert r["packets"] == 0x09080706 and r["eth"] == 0x0201 and r["flags"] == 0x1413
    assert bytes(r["src_mac"]) == bytes([4, 5, 6, 7, 8, 9]) and bytes(r["dst_mac"]) == bytes([10, 11, 12, 13, 14, 15])
    assert r["if_index"] == 0x16151413 and r["sampling"] == 2 and r["direction"] == 3
    assert r["errno"] == 0x33 and r["dscp"] == 0x60 and r["nb_obs"] == 2
    assert list(r["obs_dir"]) == [1, 0, 0, 0, 0, 0] and list(r["obs_intf"]) == [7, 8, 0, 0, 0, 0]
    assert r["ssl_version"] == 0x0303 and r["tls_types"] == 0x21 and r["cipher"] == 0 and r["key_share"] == 0


def test_read_from_zeroes_padding():
    """binary.Read skips blank fields (pkg/ebpf/bpf_x86_bpfel.go:119,145,152)."""
    w = np.frombuffer(GOLDEN_144, dtype=np.uint8).copy()
    w[39] = 0xEE
    w[40 + 66:40 + 68] = 0xEE
    w[40 + 100:40 + 104] = 0xEE
    assert O.read_from(w).tobytes() == GOLDEN_144


def test_dns_metrics_binary_encoding():
    """pkg/model/record_test.go:193-224 TestDNSMetricsBinaryEncoding."""
    b = bytes([0x10, 0, 0, 0, 0, 0, 0, 0, 0xFF, 0, 0, 0, 0, 0, 0, 0,
               0x11, 0x12, 0x13, 0x14, 0x15, 0x16, 0x17, 0x18, 1, 0, 0x80, 0, 3, 0, 0]) + \
        b"test.example.com" + bytes(16) + bytes(1)
    assert len(b) == 64
    d = np.frombuffer(b, dtype=O.DNS_DTYPE)[0]
    assert d["start"] == 0x10 and d["end"] == 0xFF and d["eth"] == 3 and d["id"] == 1
    assert d["flags"] == 0x80 and d["latency"] == 0x1817161514131211 and d["errno"] == 0
    assert bytes(d["name"]).rstrip(b"\0") == b"test.example.com"


def test_additional_metrics_binary_encoding():
    """pkg/model/record_test.go:323-347 TestAdditionalMetricsBinaryEncoding."""
    b = bytes([0x10, 0, 0, 0, 0, 0, 0, 0, 0xFF, 0, 0, 0, 0, 0, 0, 0,
               0xad, 0xde, 0xef, 0xbe, 0xef, 0xbe, 0xad, 0xde, 1, 0, 0, 0, 3, 0, 1, 0])
    a = np.frombuffer(b, dtype=O.ADD_DTYPE)[0]
    assert a["start"] == 0x10 and a["end"] == 0xFF and a["eth"] == 3
    assert a["rtt"] == 0xDEADBEEFBEEFDEAD and a["ipsec_enc"] == 1 and a["ipsec_ret"] == 1


def _content(start=0, end=0, packets=0):
    c = O.Content()
    m = np.zeros(1, dtype=O.REC_DTYPE)
    m["start"], m["end"], m["packets"] = start, end, packets
    C.memmove(c.metrics, m.tobytes()[40:], 104)
    return c


def _met(c):
    return np.frombuffer(bytes(40) + bytes(c.metrics), dtype=O.REC_DTYPE)[0]


def _dns(**kw):
    d = np.zeros(1, dtype=O.DNS_DTYPE)
    for k, v in kw.items():
        d[k] = v
    return d


def _add(**kw):
    d = np.zeros(1, dtype=O.ADD_DTYPE)
    for k, v in kw.items():
        d[k] = v
    return d


def test_accumulate_dns():
    """pkg/model/flow_content_test.go:11-53 TestAccumulateDNS."""
    c = _content(10, 20, 3)
    d1 = O.as_bytes(_dns(start=25, end=25, latency=1000, id=1, flags=0b0011))
    O.lib().oracle_accumulate_dns(C.byref(c), O._p(d1))
    m = _met(c)
    assert (m["start"], m["end"], m["packets"]) == (10, 25, 3)
    assert c.has_dns and bytes(c.dns) == d1.tobytes()
    d2 = O.as_bytes(_dns(start=30, end=30, latency=2000, id=1, flags=0b1001))
    O.lib().oracle_accumulate_dns(C.byref(c), O._p(d2))
    m = _met(c)
    assert (m["start"], m["end"], m["packets"]) == (10, 30, 3)
    d = np.frombuffer(bytes(c.dns), dtype=O.DNS_DTYPE)[0]
    assert (d["start"], d["end"], d["latency"], d["id"], d["flags"]) == (25, 25, 2000, 1, 0b1011)


def test_accumulate_additional():
    """pkg/model/flow_content_test.go:184-246 TestAccumulateAdditional."""
    c = _content(10, 20, 3)
    f = O.lib().oracle_accumulate_additional
    f(C.byref(c), O._p(O.as_bytes(_add(start=25, end=25, rtt=200, ipsec_enc=1))))
    a = np.frombuffer(bytes(c.additional), dtype=O.ADD_DTYPE)[0]
    assert (_met(c)["start"], _met(c)["end"]) == (10, 25)
    assert (a["start"], a["end"], a["rtt"], a["ipsec_enc"], a["ipsec_ret"]) == (25, 25, 200, 1, 0)
    f(C.byref(c), O._p(O.as_bytes(_add(start=30, end=30, rtt=1000))))            # higher RTT, no ipsec info
    a = np.frombuffer(bytes(c.additional), dtype=O.ADD_DTYPE)[0]
    assert _met(c)["end"] == 30
    assert (a["start"], a["end"], a["rtt"], a["ipsec_enc"], a["ipsec_ret"]) == (25, 25, 1000, 1, 0)
    f(C.byref(c), O._p(O.as_bytes(_add(start=30, end=30, rtt=800, ipsec_ret=5))))  # lower RTT, ipsec failure
    a = np.frombuffer(bytes(c.additional), dtype=O.ADD_DTYPE)[0]
    assert (a["rtt"], a["ipsec_enc"], a["ipsec_ret"]) == (1000, 0, 5)
    f(C.byref(c), O._p(O.as_bytes(_add(start=30, end=30, rtt=800))))              # no change
    a = np.frombuffer(bytes(c.additional), dtype=O.ADD_DTYPE)[0]
    assert (a["start"], a["end"], a["rtt"], a["ipsec_enc"], a["ipsec_ret"]) == (25, 25, 1000, 0, 5)
    assert (_met(c)["start"], _met(c)["end"], _met(c)["packets"]) == (10, 30, 3)


def test_accumulate_now_base():
    """pkg/model/flow_content_test.go:338-380 TestAccumulateNowBase (DNS + Additional rows)."""
    c = _content()
    O.lib().oracle_accumulate_dns(C.byref(c), O._p(O.as_bytes(_dns(start=25, end=25))))
    assert (_met(c)["start"], _met(c)["end"], _met(c)["eth"]) == (25, 25, 0)
    c = _content()
    O.lib().oracle_accumulate_additional(C.byref(c), O._p(O.as_bytes(_add(start=25, end=25, eth=3))))
    assert (_met(c)["start"], _met(c)["end"], _met(c)["eth"]) == (25, 25, 3)


def test_accumulate_base_order_dependent_fields():
    """Source-pinned (no reference unit test): pkg/model/flow_content.go:45-59."""
    p = mkrec(SRC1, DST1, 1, 2, start=0, end=0, eth=0, dscp=0, sampling=0)
    o = mkrec(SRC1, DST1, 1, 2, start=50, end=60, bytes=7, packets=0xFFFFFFFF, flags=0x12, eth=0x0800,
              src_mac=[1, 2, 3, 4, 5, 6], dscp=10, sampling=5, if_index=9, direction=1)
    p["packets"] = 3
    m = O.accumulate_base(O.as_bytes(p)[40:], O.as_bytes(o)[40:])
    r = np.frombuffer(bytes(40) + m.tobytes(), dtype=O.REC_DTYPE)[0]
    assert (r["start"], r["end"], r["bytes"], r["packets"], r["flags"]) == (50, 60, 7, 2, 0x12)   # u32 wrap
    assert r["eth"] == 0x0800 and r["dscp"] == 10 and r["sampling"] == 5
    assert bytes(r["src_mac"]) == bytes([1, 2, 3, 4, 5, 6])
    assert r["if_index"] == 0 and r["direction"] == 0        # never merged
    o2 = mkrec(SRC1, DST1, 1, 2, start=0, end=10, eth=0, dscp=0, sampling=0, src_mac=[9, 9, 9, 9, 9, 9])
    m2 = O.accumulate_base(m, O.as_bytes(o2)[40:])
    r2 = np.frombuffer(bytes(40) + m2.tobytes(), dtype=O.REC_DTYPE)[0]
    assert (r2["start"], r2["end"]) == (50, 60)                # start 0 ignored, end keeps max
    assert r2["eth"] == 0x0800 and r2["dscp"] == 10 and r2["sampling"] == 5   # zero never overwrites
    assert bytes(r2["src_mac"]) == bytes([1, 2, 3, 4, 5, 6])   # MAC only set while all-zero


def test_config1_accounter_replay():
    """BASELINE.json configs[0]: 10k records, 100 5-tuples, maxEntries 5000 -> one eviction of 100 flows."""
    rng = np.random.default_rng(1)
    n, nk = 10_000, 100
    recs = np.zeros(n, dtype=O.REC_DTYPE)
    kid = rng.integers(0, nk, n)
    recs["src_ip"][:, 10:12] = 0xFF
    recs["src_ip"][:, 12:16] = kid[:, None].astype(np.uint32).view(np.uint8).reshape(n, 4)[:, ::-1]
    recs["dst_ip"][:, 10:12] = 0xFF
    recs["dst_ip"][:, 15] = 1
    recs["src_port"] = 1024 + kid
    recs["dst_port"] = 443
    recs["proto"] = 6
    recs["start"] = recs["end"] = 1_000 + np.arange(n)
    recs["bytes"] = rng.integers(64, 1501, n)
    recs["packets"] = 1
    recs["flags"] = 0x10
    acc = O.Accounter(5000)
    acc.account(recs)
    assert acc.pending() == 0
    out = acc.evict().view(O.REC_DTYPE).reshape(-1)
    assert len(out) == nk
    assert int(out["packets"].sum()) == n and int(out["bytes"].sum()) == int(recs["bytes"].sum())
    for r in out:
        sel = recs[(recs["src_port"] == r["src_port"])]
        assert r["packets"] == len(sel) and r["bytes"] == sel["bytes"].sum()
        assert r["start"] == sel["start"].min() and r["end"] == sel["end"].max()

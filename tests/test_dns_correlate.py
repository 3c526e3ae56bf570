"""K7 — DNS query -> response correlation (fa_ingest_dns_packets, csrc/dnscorr.cu).

Reference: track_dns_packet and its dns_flows map (bpf/dns_tracker.h:23-37,68-127), the dns_metrics sample flow_monitor
makes of the result (bpf/flows.c:210-213,291-330), lookupAndDeleteDNSMap (pkg/tracer/tracer.go:1235-1257).  The reference
has no unit test for any of it (eBPF), so the oracle's restatement is SOURCE-PINNED: the first tests below pin it to
hand-worked cases derived from the source text, the rest compare the engine (on the CPU: the emulated C ABI; on a B200:
tests/test_zz_gpu_dns_correlate.py) with it bit for bit."""
import numpy as np
import pytest

import oracle_lib as O

pytestmark = pytest.mark.timeout(900)

QR = 0x8000
CLIENT = bytes([0] * 10 + [0xFF, 0xFF, 10, 0, 0, 1])
SERVER = bytes([0] * 10 + [0xFF, 0xFF, 10, 0, 0, 53])


def pkt(src, dst, sport, dport, ts, dns_id, flags, proto=17, name=b"\x03www\x07example\x03com\x00", eth=0x0800):
    r = np.zeros(1, dtype=O.DNSREC_DTYPE)
    ident = np.zeros(40, dtype=np.uint8)
    ident[0:16] = np.frombuffer(src, dtype=np.uint8); ident[16:32] = np.frombuffer(dst, dtype=np.uint8)
    ident[32:34] = np.frombuffer(np.uint16(sport).tobytes(), dtype=np.uint8)
    ident[34:36] = np.frombuffer(np.uint16(dport).tobytes(), dtype=np.uint8)
    ident[36] = proto
    r["id"][0] = ident
    d = r["dns"]
    d["start"] = 7; d["latency"] = 12345; d["errno"] = 99          # ignored on input
    d["end"] = ts; d["id"] = dns_id; d["flags"] = flags; d["eth"] = eth
    nm = np.zeros(32, dtype=np.uint8); nm[:len(name)] = np.frombuffer(name, dtype=np.uint8)
    d["name"][0] = nm
    d["pad"] = 0xEE
    r["dns"] = d
    return r


def query(ts, dns_id, sport=40000, **kw):
    return pkt(CLIENT, SERVER, sport, 53, ts, dns_id, 0x0100, **kw)


def response(ts, dns_id, sport=40000, flags=QR | 0x0180, **kw):
    return pkt(SERVER, CLIENT, 53, sport, ts, dns_id, flags, **kw)


def samples_of(pkts, corr=None):
    corr = corr or O.DnsCorrelator()
    return corr.packets(np.concatenate(pkts)).view(O.DNSREC_DTYPE).reshape(-1), corr


# ------------------------------------------------------------------------------------------ the restatement, by hand
def test_query_then_response_gives_the_latency_and_deletes_the_query():
    s, corr = samples_of([query(1_000, 77), response(1_450, 77)])
    assert len(s) == 1 and corr.pending() == 0
    d = s[0]["dns"]
    assert (d["start"], d["end"], d["latency"], d["id"], d["flags"], d["eth"], d["errno"], d["pad"]) == (1_450, 1_450, 450, 77, QR | 0x0180, 0x0800, 0, 0)
    assert bytes(s[0]["id"]) == bytes(response(0, 0)["id"][0])                       # the sample belongs to the RESPONSE packet's flow
    assert bytes(d["name"][:17]) == b"\x03www\x07example\x03com\x00"


def test_response_without_query_is_enoent():
    s, corr = samples_of([response(500, 9)])
    assert len(s) == 1 and (s[0]["dns"]["latency"], s[0]["dns"]["errno"], s[0]["dns"]["id"]) == (0, 2, 9)


def test_repeated_query_keeps_the_first_timestamp_and_reports_eexist():
    # bpf_map_update_elem(..., BPF_NOEXIST) fails with -EEXIST; track_dns_packet returns it; the u8 errno field holds 239
    s, corr = samples_of([query(100, 5), query(160, 5), response(400, 5), response(420, 5)])
    assert [(int(x["dns"]["errno"]), int(x["dns"]["id"]), int(x["dns"]["latency"])) for x in s] == [(239, 0, 0), (0, 5, 300), (2, 5, 0)]
    assert bytes(s[0]["id"]) == bytes(query(0, 0)["id"][0]) and not s[0]["dns"]["name"].any() and s[0]["dns"]["flags"] == 0
    assert corr.pending() == 0


def test_keys_differ_by_id_port_and_protocol_and_responses_match_the_reversed_tuple():
    s, corr = samples_of([query(10, 1), query(20, 2), query(30, 1, sport=40001), query(40, 1, proto=6),
                          response(110, 1, sport=40001), response(120, 1, proto=6), response(130, 2), response(140, 1),
                          pkt(CLIENT, SERVER, 40000, 53, 150, 1, QR)])               # a "response" in the query's direction: other key
    assert [int(x["dns"]["latency"]) for x in s] == [80, 80, 110, 130, 0] and int(s[4]["dns"]["errno"]) == 2


def test_response_with_id_zero_leaves_no_sample_but_consumes_the_query():
    s, corr = samples_of([query(10, 0), response(30, 0)])
    assert len(s) == 0 and corr.pending() == 0


def test_full_map_and_purge():
    corr = O.DnsCorrelator(max_entries=3)
    s, _ = samples_of([query(100 + i, i + 1) for i in range(5)], corr)
    assert [int(x["dns"]["errno"]) for x in s] == [249, 249] and corr.pending() == 3          # (u8)-E2BIG
    # time.Duration(now - ts) >= timeout: queries at 100, 101, 102
    assert corr.purge(200, 99) == 2 and corr.pending() == 1
    assert corr.purge(50, 10) == 0                                                             # "future" entries: negative duration
    s, _ = samples_of([response(300, 3)], corr)
    assert int(s[0]["dns"]["latency"]) == 198


# ------------------------------------------------------------------------------------------ engine vs restatement
def dns_stream(seed, n, n_clients=40, n_ids=6, dup=0.15, orphan=0.1, zero_id=0.05):
    """Transactions of n_clients x n_ids keys in random interleaving: query, sometimes a repeated query, response,
    sometimes a second response or a response nobody asked for; timestamps increase with the index."""
    rng = np.random.default_rng(seed)
    out = []
    ts = 1_000
    open_q = []
    while len(out) < n:
        ts += int(rng.integers(1, 50))
        c = int(rng.integers(0, n_clients)); i = 0 if rng.random() < zero_id else int(rng.integers(1, n_ids + 1))
        proto = 17 if rng.random() < 0.8 else 6
        r = rng.random()
        if open_q and r < 0.45:
            c, i, proto = open_q.pop(int(rng.integers(0, len(open_q))))
            out.append(response(ts, i, sport=30000 + c, proto=proto, name=bytes([3, 97 + c % 26, 98, 99, 0])))
            if rng.random() < dup:
                out.append(response(ts + 1, i, sport=30000 + c, proto=proto))
        elif r < 0.45 + orphan:
            out.append(response(ts, i, sport=30000 + c, proto=proto))
        else:
            out.append(query(ts, i, sport=30000 + c, proto=proto))
            open_q.append((c, i, proto))
            if rng.random() < dup:
                out.append(query(ts + 2, i, sport=30000 + c, proto=proto))
    return np.concatenate(out[:n])


def compare_dns(eng, om):
    from test_gpu_features import compare
    compare(eng, om)


def run_case(pkts_batches, max_entries=None, flows=1 << 12, purge=None):
    import netobserv_ebpf_agent_b200 as fa
    corr = O.DnsCorrelator(max_entries or (1 << 20))
    om = O.FlowMap()
    with fa.FlowAggEngine(flows, flags=fa.FA_F_ENABLE_DNS | fa.FA_F_ENABLE_RTT, max_batch=4_096) as eng:
        for k, b in enumerate(pkts_batches):
            eng.ingest_dns_packets(b)
            om.fold_dns(corr.packets(b))
            if purge and k == purge[0]:
                eng.purge_stale_dns(purge[1], purge[2])
                gone = corr.purge(purge[1], purge[2])
                assert eng.stats()["dns_queries_purged"] == gone
            assert eng.stats()["dns_queries_pending"] == corr.pending()
        st = eng.stats()
        compare_dns(eng, om)
    return st


def test_emulated_engine_matches_the_restatement(engine_emul):
    s = dns_stream(3, 3_000)
    st = run_case([s[:1_100], s[1_100:1_900], s[1_900:]])
    assert st["dns_packets_ingested"] == 3_000 and st["dns_ingested"] > 0


def test_emulated_engine_hand_cases(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    cases = [query(100, 5), query(160, 5), response(400, 5), response(420, 5), query(10, 0), response(30, 0), response(77, 9)]
    run_case([np.concatenate(cases)])
    run_case([c for c in cases])                                   # one packet per call: the state lives in the table


def test_tail_kernel_alone_is_exact(engine_emul, monkeypatch):
    """No round kernels at all: every packet goes through the single-thread tail (what a key with more packets in one
    batch than rounds falls back to)."""
    monkeypatch.setenv("FA_EMUL_DNS_ROUNDS", "0")
    s = dns_stream(4, 1_200, n_clients=5, n_ids=2, dup=0.5)
    run_case([s[:700], s[700:]])


def test_one_hot_key_needs_more_rounds_than_are_launched(engine_emul):
    one = [query(10 + 4 * i, 7) if i % 3 else response(10 + 4 * i, 7) for i in range(60)]
    run_case([np.concatenate(one)])


def test_small_map_rebuilds_and_purge(engine_emul, monkeypatch):
    """FA_DNS_MAX_ENTRIES=64: a 256-slot table; answered keys are dropped by the rebuilds, pending queries survive them; a purge in
    between deletes exactly what the restatement deletes."""
    monkeypatch.setenv("FA_DNS_MAX_ENTRIES", "64")
    s = dns_stream(5, 2_000, n_clients=12, n_ids=3, orphan=0.05)
    st = run_case([s[:900], s[900:]], max_entries=64, purge=(0, 60_000, 9_000))
    assert st["dns_map_full"] == 0


def test_full_map_counts(engine_emul, monkeypatch):
    """Queries beyond max_entries get errno 249; WHICH ones is not order-exact on the device, their number is."""
    import netobserv_ebpf_agent_b200 as fa
    monkeypatch.setenv("FA_DNS_MAX_ENTRIES", "16")
    qs = np.concatenate([query(100 + i, 1 + i % 200, sport=20000 + i) for i in range(40)])
    with fa.FlowAggEngine(1 << 10, flags=fa.FA_F_ENABLE_DNS, max_batch=4_096) as eng:
        eng.ingest_dns_packets(qs)
        st = eng.stats()
        recs, dns, add, pres = eng.evict(features=True)
    assert st["dns_queries_pending"] == 16 and st["dns_map_full"] == 24 and st["dns_ingested"] == 24
    assert sorted(dns.view(O.DNS_DTYPE).reshape(-1)["errno"].tolist()) == [249] * 24


def test_needs_the_dns_flag(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(100) as eng:
        with pytest.raises(fa.FlowAggError):
            eng.ingest_dns_packets(query(1, 1))
        eng.purge_stale_dns(10, 1)                                 # nothing to purge: OK


def test_restatement_agrees_with_an_independent_python_model():
    """A second, dict-based restatement written from the source text alone (dns_tracker.h:92-110, flows.c:291-330): the C oracle
    and it must agree on every sample of a messy stream."""
    s = dns_stream(21, 4_000, n_clients=25, n_ids=3, dup=0.3, orphan=0.15, zero_id=0.1)
    got = O.DnsCorrelator(max_entries=40).packets(s).view(O.DNSREC_DTYPE).reshape(-1)
    dns_flows, want = {}, []
    for p in s.view(O.DNSREC_DTYPE).reshape(-1):
        ident, d = bytes(p["id"]), p["dns"]
        ts, did, flags = int(d["end"]), int(d["id"]), int(d["flags"])
        src, dst, sp, dp, proto = ident[0:16], ident[16:32], ident[32:34], ident[34:36], ident[36]
        errno, latency, pid, pflags, name = 0, 0, 0, 0, bytes(32)
        if not flags & QR:
            key = (src, dst, sp, dp, did, proto)
            if key in dns_flows:
                errno = (-17) & 0xFF
            elif len(dns_flows) >= 40:
                errno = (-7) & 0xFF
            else:
                dns_flows[key] = ts
        else:
            key = (dst, src, dp, sp, did, proto)
            if key in dns_flows:
                latency = ts - dns_flows.pop(key)
            else:
                errno = 2
            pid, pflags, name = did, flags, bytes(d["name"])
        if pid or errno:
            want.append((ident, ts, ts, latency, pid, pflags, int(d["eth"]), errno, name))
    assert len(got) == len(want)
    for g, w in zip(got, want):
        gd = g["dns"]
        assert (bytes(g["id"]), int(gd["start"]), int(gd["end"]), int(gd["latency"]), int(gd["id"]), int(gd["flags"]), int(gd["eth"]),
                int(gd["errno"]), bytes(gd["name"])) == w

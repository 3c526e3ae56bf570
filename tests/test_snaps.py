"""(f4) raw-header front end: fa_ingest_snaps parses packet snapshots on the device like flow_monitor's fill_ethhdr /
fill_iphdr / fill_ip6hdr / fill_l4info / set_flags (reference bpf/utils.h:24-167) and builds the single-packet flow of
bpf/flows.c:176-245.

The reference has no unit test for its eBPF parsing (source-pinned only), so the oracle restatement
(oracle_parse_snap) is pinned here against frames assembled by hand from the RFC header layouts with the expected
record fields written out; the engine is then compared with the oracle on random frame mixes (TCP / UDP / SCTP / ICMP /
ICMPv6 over IPv4 and IPv6, non-IP frames, truncated captures), on the emulation (CPU) and on the device (gpu)."""
import struct

import numpy as np
import pytest

import oracle_lib as O

HDR = 24


def mac(s):
    return bytes(int(x, 16) for x in s.split(":"))


def eth(dst, src, proto):
    return mac(dst) + mac(src) + struct.pack(">H", proto)


def ipv4(src, dst, proto, tos=0, ihl=5):
    return struct.pack(">BBHHHBBH4s4s", (4 << 4) | ihl, tos, 40, 0, 0, 64, proto, 0, bytes(src), bytes(dst))


def ipv6(src, dst, nexthdr, tclass=0):
    return struct.pack(">IHBB16s16s", (6 << 28) | (tclass << 20), 20, nexthdr, 64, bytes(src), bytes(dst))


def tcp(sport, dport, flags):
    return struct.pack(">HHIIBBHHH", sport, dport, 1, 2, 5 << 4, flags, 1000, 0, 0)


def udp(sport, dport):
    return struct.pack(">HHHH", sport, dport, 8, 0)


def sctp(sport, dport):
    return struct.pack(">HHII", sport, dport, 7, 0)


def icmp(t, c):
    return struct.pack(">BBHI", t, c, 0, 0)


def snap(frame, stride, ts=1, length=None, if_index=3, sampling=1, direction=0, cap=None):
    room = stride - HDR
    cap = min(len(frame), room) if cap is None else cap
    b = struct.pack("<QIIIHBB", ts, len(frame) if length is None else length, if_index, sampling, cap, direction, 0xEE)
    data = frame[:room].ljust(room, b"\xAA")                     # bytes beyond cap_len are garbage the parser must not use
    return np.frombuffer(b + data, dtype=np.uint8)


def one(frame, stride=104, **kw):
    s = snap(frame, stride, **kw)
    rec = np.zeros(144, dtype=np.uint8)
    ok = O.lib().oracle_parse_snap(O._p(s), stride, O._p(rec))
    return ok, rec.view(O.REC_DTYPE)[0], rec


V4A, V4B = [10, 1, 2, 3], [192, 168, 7, 9]
V6A, V6B = list(range(1, 17)), list(range(101, 117))
M1, M2 = "02:00:00:00:00:01", "02:aa:bb:cc:dd:ee"


def test_oracle_ipv4_tcp_known_answer():
    ok, r, raw = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 6, tos=0xB8) + tcp(443, 51000, 0x12), ts=777, length=1514, if_index=9,
                     sampling=50, direction=1)
    assert ok == 1
    assert bytes(r["src_ip"]) == bytes([0] * 10 + [0xFF, 0xFF] + V4A) and bytes(r["dst_ip"]) == bytes([0] * 10 + [0xFF, 0xFF] + V4B)
    assert (r["src_port"], r["dst_port"], r["proto"], r["icmp_type"], r["icmp_code"]) == (443, 51000, 6, 0, 0)
    assert (r["start"], r["end"], r["bytes"], r["packets"], r["eth"]) == (777, 777, 1514, 1, 0x0800)
    assert r["flags"] == 0x100                                    # SYN+ACK -> SYN_ACK_FLAG only (utils.h:26-28)
    assert bytes(r["src_mac"]) == mac(M1) and bytes(r["dst_mac"]) == mac(M2)
    assert (r["if_index"], r["sampling"], r["direction"], r["dscp"]) == (9, 50, 1, 0xB8 >> 2)
    assert not raw[99:].any() and raw[39] == 0 and r["nb_obs"] == 0


@pytest.mark.parametrize("bits,want", [(0x12, 0x100), (0x11, 0x200), (0x14, 0x400), (0x01, 0x01), (0x02, 0x02), (0x10, 0x10),
                                       (0x04, 0x04), (0x08, 0x08), (0x20, 0x20), (0x40, 0x40), (0x80, 0x80), (0x18, 0x10),
                                       (0x00, 0), (0x29, 0x01), (0xC0, 0x40)])
def test_oracle_set_flags_chain(bits, want):
    """utils.h:24-50: an if / else-if chain: exactly one flag per packet, in that priority."""
    ok, r, _ = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 6) + tcp(1, 2, bits))
    assert ok == 1 and r["flags"] == want


def test_oracle_ipv6_and_the_other_transports():
    ok, r, _ = one(eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 17, tclass=0x2E << 2) + udp(53, 40000))
    assert ok == 1 and bytes(r["src_ip"]) == bytes(V6A) and bytes(r["dst_ip"]) == bytes(V6B)
    assert (r["src_port"], r["dst_port"], r["proto"], r["eth"], r["dscp"], r["flags"]) == (53, 40000, 17, 0x86DD, 0x2E, 0)
    ok, r, _ = one(eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 58) + icmp(128, 3))
    assert ok == 1 and (r["proto"], r["icmp_type"], r["icmp_code"], r["src_port"], r["dst_port"]) == (58, 128, 3, 0, 0)
    ok, r, _ = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 1) + icmp(8, 0))
    assert ok == 1 and (r["proto"], r["icmp_type"], r["icmp_code"]) == (1, 8, 0)
    ok, r, _ = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 132) + sctp(2905, 2906))
    assert ok == 1 and (r["proto"], r["src_port"], r["dst_port"]) == (132, 2905, 2906)
    ok, r, _ = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 47) + b"\x00" * 8)                  # GRE: protocol kept, nothing parsed
    assert ok == 1 and (r["proto"], r["src_port"], r["dst_port"], r["flags"]) == (47, 0, 0, 0)
    ok, r, _ = one(eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 6) + tcp(80, 81, 0x02))              # IPv6 + TCP needs 74 header bytes
    assert ok == 1 and (r["src_port"], r["dst_port"], r["flags"]) == (80, 81, 0x02)
    ok, r, _ = one(eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 6) + tcp(80, 81, 0x02), stride=88)   # a 64-byte snap cuts the TCP header
    assert ok == 1 and (r["proto"], r["src_port"], r["dst_port"], r["flags"]) == (6, 0, 0, 0)


def test_oracle_discards_and_bounds():
    f4 = eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 6) + tcp(1, 2, 0x10)
    assert one(eth(M2, M1, 0x0806) + b"\x00" * 28)[0] == 0                                   # ARP
    assert one(eth(M2, M1, 0x8100) + b"\x00\x01\x08\x00" + ipv4(V4A, V4B, 6))[0] == 0        # VLAN tags are not unwrapped
    assert one(f4, cap=13)[0] == 0 and one(f4, cap=33)[0] == 0                                # Ethernet / IP header beyond data_end
    ok, r, _ = one(f4, cap=34)
    assert ok == 1 and (r["proto"], r["src_port"], r["flags"]) == (6, 0, 0)                   # IP header fits, TCP header does not
    ok, r, _ = one(f4, cap=53)
    assert ok == 1 and (r["src_port"], r["flags"]) == (0, 0)
    ok, r, _ = one(f4, cap=54)
    assert ok == 1 and (r["src_port"], r["dst_port"], r["flags"]) == (1, 2, 0x10)
    f6 = eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 17) + udp(5, 6)
    assert one(f6, cap=53)[0] == 0 and one(f6, cap=54)[0] == 1 and one(f6, cap=61)[1]["src_port"] == 0 and one(f6, cap=62)[1]["src_port"] == 5
    ok, r, _ = one(eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 17, ihl=6) + b"\x01\x02\x03\x04" + udp(7, 8))
    assert ok == 1 and (r["src_port"], r["dst_port"]) == (0x0102, 0x0304)                     # options are NOT skipped (utils.h:114)


def random_snaps(rng, n, stride, n_hosts=40):
    """A mix of everything above, few hosts / ports so that flows repeat."""
    out = np.zeros((n, stride), dtype=np.uint8)
    for i in range(n):
        v6 = rng.random() < 0.35
        a, b = (rng.integers(0, n_hosts, 2) + 1).tolist()
        src = (V6A[:15] + [a]) if v6 else (V4A[:3] + [a])
        dst = (V6B[:15] + [b]) if v6 else (V4B[:3] + [b])
        kind = rng.choice(["tcp", "udp", "sctp", "icmp", "other", "nonip"], p=[0.5, 0.2, 0.05, 0.1, 0.05, 0.1])
        sp, dp = int(rng.integers(1, 6)) * 1000, int(rng.choice([53, 80, 443]))
        if kind == "tcp":
            l4, proto = tcp(sp, dp, int(rng.integers(0, 256))), 6
        elif kind == "udp":
            l4, proto = udp(sp, dp), 17
        elif kind == "sctp":
            l4, proto = sctp(sp, dp), 132
        elif kind == "icmp":
            l4, proto = icmp(int(rng.integers(0, 4)), int(rng.integers(0, 3))), (58 if rng.random() < 0.5 else 1)
        else:
            l4, proto = bytes(rng.integers(0, 256, 12, dtype=np.uint8)), 47
        ip = ipv6(src, dst, proto, tclass=int(rng.integers(0, 256))) if v6 else ipv4(src, dst, proto, tos=int(rng.integers(0, 256)))
        frame = eth(M2, M1, int(rng.choice([0x0806, 0x8100, 0x88CC])) if kind == "nonip" else (0x86DD if v6 else 0x0800)) + ip + l4
        cap = None if rng.random() < 0.8 else int(rng.integers(0, len(frame) + 1))
        out[i] = snap(frame, stride, ts=1_000 + 3 * i, length=int(rng.integers(60, 1515)), if_index=int(rng.integers(1, 4)),
                      sampling=int(rng.choice([0, 1, 50])), direction=int(rng.integers(0, 2)), cap=cap)
    return out


def check(n, stride, max_entries, max_batch):
    import netobserv_ebpf_agent_b200 as fa
    rng = np.random.default_rng(stride)
    snaps = random_snaps(rng, n, stride)
    recs, src = O.parse_snaps(snaps, stride)
    assert 0.5 * n < len(recs) < n
    acc = O.Accounter(max_entries)
    acc.account(recs)
    want = O.sort_records(acc.evict())
    with fa.FlowAggEngine(max_entries, max_batch=max_batch) as eng:
        h = (n // 3) * stride
        flat = snaps.reshape(-1)
        rc, took = eng.ingest_snaps(flat[:h], stride)
        assert rc == 0 and took == n // 3
        rc, took = eng.ingest_snaps(flat[h:], stride)
        assert rc == 0 and took == n - n // 3
        st = eng.stats()
        assert st["snaps_ingested"] == n and st["snaps_discarded"] == n - len(recs) and st["records_ingested"] == len(recs)
        assert st["h2d_bytes"] == n * stride
        got = O.sort_records(eng.evict())
    assert np.array_equal(got, want)
    acc.close()
    return len(want)


def check_full_cut(stride=88):
    """The Accounter's maxEntries rule cuts a snapshot chunk where it cuts the parsed records; `consumed` is in snapshots."""
    import netobserv_ebpf_agent_b200 as fa
    snaps = random_snaps(np.random.default_rng(3), 1_500, stride)
    recs, src = O.parse_snaps(snaps, stride)
    with fa.FlowAggEngine(40, max_batch=1_024) as a, fa.FlowAggEngine(40, max_batch=1_024) as b:
        rc_a, took_a = a.ingest_snaps(snaps.reshape(-1), stride)
        rc_b, took_b = b.ingest(recs)
        assert rc_a == rc_b == fa.FA_FULL and 0 < took_b < len(recs)
        assert took_a == int(src[took_b])                        # everything before the first record that did not fit
        assert np.array_equal(O.sort_records(a.evict()), O.sort_records(b.evict()))
        assert a.stats()["snaps_ingested"] == took_a and a.stats()["snaps_discarded"] == took_a - took_b


def test_snaps_on_the_emulation(engine_emul):
    assert check(n=2_500, stride=104, max_entries=1 << 12, max_batch=1_024) > 200
    check(n=1_200, stride=88, max_entries=1 << 12, max_batch=4_096)


def test_full_cut_with_snaps_on_the_emulation(engine_emul):
    check_full_cut()


def test_bad_strides_are_refused(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(64, max_batch=256) as eng:
        for stride in (32, 90, 160):
            with pytest.raises(fa.FlowAggError) as ei:
                eng.ingest_snaps(np.zeros(stride * 2, dtype=np.uint8), stride)
            assert ei.value.code == -22


@pytest.mark.gpu
def test_snaps_gpu():
    check(n=60_000, stride=104, max_entries=1 << 17, max_batch=16_384)
    check(n=20_000, stride=88, max_entries=1 << 16, max_batch=1 << 20)
    check(n=5_000, stride=152, max_entries=1 << 14, max_batch=1 << 12)
    check_full_cut()


@pytest.mark.gpu
def test_snaps_from_device_memory_gpu():
    import torch

    import netobserv_ebpf_agent_b200 as fa
    stride = 104
    snaps = random_snaps(np.random.default_rng(8), 30_000, stride)
    recs, _ = O.parse_snaps(snaps, stride)
    acc = O.Accounter(1 << 16); acc.account(recs); want = O.sort_records(acc.evict()); acc.close()
    d = torch.from_numpy(snaps.reshape(-1).copy()).cuda()
    with fa.FlowAggEngine(1 << 16, max_batch=8_192) as eng:
        rc, took = eng.ingest_snaps(d, stride)
        assert rc == 0 and took == 30_000 and eng.stats()["h2d_bytes"] == 0
        assert np.array_equal(O.sort_records(eng.evict()), want)

"""K8 — evicted flows -> pbflow.Record wire bytes (fa_pb_encode) against the Python protobuf library.

Checker: tests/pbflow_ref.py (proto/flow.proto restated as runtime descriptors; NewRecord + FlowToPB restated).
Pinned to the reference's own test of this path, pkg/exporter/kafka_proto_test.go:26-86 (TestProtoConversion) and
:88-130 (TestIdenticalKeys).  The CPU tests drive the kernels through the engine emulation (tests/emul); the
gpu-marked ones through libflowagg.so."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib as O
import pbflow_ref as PB
from common import gen_host


def pb_encode(lib, h, recs, dns=None, add=None, present=None, drops=None, now=1_700_000_000_123_456_789, mono=5_000_000_000_000,
              agent_ip=bytes(10) + b"\xff\xff" + bytes([10, 1, 2, 3]), agent_v4=1, ifaces=(), wrap=False, keys=True):
    from netobserv_ebpf_agent_b200._lib import IfaceName, PbParams
    recs = np.ascontiguousarray(recs).view(np.uint8).reshape(-1, 144)
    n = len(recs)
    rows = (IfaceName * max(len(ifaces), 1))()
    for i, (idx, mac, name, udn) in enumerate(ifaces):
        rows[i].if_index = idx
        for b in range(6):
            rows[i].mac[b] = mac[b]
        rows[i].name_len, rows[i].udn_len = len(name), len(udn)
        rows[i].name, rows[i].udn = name.encode(), udn.encode()
    p = PbParams(now_unix_ns=now, mono_now_ns=mono, agent_ip_is_v4=agent_v4, flags=1 if wrap else 0,
                 ifaces=C.cast(rows, C.POINTER(IfaceName)) if ifaces else None, n_ifaces=len(ifaces))
    for i in range(16):
        p.agent_ip[i] = agent_ip[i]
    out_len = C.c_size_t(0)
    ptr = lambda a: C.c_void_p(a.ctypes.data) if a is not None else None   # noqa: E731
    rc = lib.fa_pb_encode(h, ptr(recs), ptr(dns), ptr(add), ptr(drops), ptr(present), n, C.byref(p), None, 0, None, None, C.byref(out_len))
    assert rc == -7 and out_len.value > 0, rc                              # FA_E_2BIG reports the size
    out = np.zeros(out_len.value, dtype=np.uint8)
    offs = np.zeros(n + 1, dtype=np.uint64)
    kout = np.zeros((n, 32), dtype=np.uint8) if keys else None
    rc = lib.fa_pb_encode(h, ptr(recs), ptr(dns), ptr(add), ptr(drops), ptr(present), n, C.byref(p), ptr(out), out.size, ptr(offs),
                          ptr(kout), C.byref(out_len))
    assert rc == 0, (rc, lib.fa_last_error())
    assert offs[0] == 0 and offs[-1] == out.size == out_len.value
    return out.tobytes(), offs.astype(np.int64), kout


def golden_record():
    """The record of TestProtoConversion (kafka_proto_test.go:33-56), as an evicted 144-byte flow."""
    r = np.zeros(1, dtype=O.REC_DTYPE)
    r["src_ip"][0] = list(bytes(10) + b"\xff\xff" + bytes([192, 1, 2, 3]))
    r["dst_ip"][0] = list(bytes(10) + b"\xff\xff" + bytes([127, 3, 2, 1]))
    r["src_port"], r["dst_port"], r["icmp_type"], r["proto"] = 4321, 1234, 8, 210
    r["direction"], r["eth"] = 1, 3
    r["src_mac"][0] = [0xaa, 0xbb, 0xcc, 0xdd, 0xee, 0xff]; r["dst_mac"][0] = [0x11, 0x22, 0x33, 0x44, 0x55, 0x66]
    r["bytes"], r["packets"], r["flags"] = 789, 987, 1
    r["if_index"], r["nb_obs"] = 7, 1
    r["obs_intf"][0][0], r["obs_dir"][0][0] = 9, 1                         # Interfaces: veth0 (dir 1 = first seen) and abcde
    mono = 5_000_000_000_000
    r["start"], r["end"] = mono - 5_000_000_000, mono                      # TimeFlowStart = now - 5 s, TimeFlowEnd = now
    return r


def check_golden(lib, h):
    r = golden_record()
    ifaces = [(7, bytes(6), "veth0", ""), (9, bytes(6), "abcde", "")]
    now, mono = 1_700_000_000_123_456_789, 5_000_000_000_000
    raw, offs, keys = pb_encode(lib, h, r, now=now, mono=mono, ifaces=ifaces)
    m = PB.Record(); m.ParseFromString(raw)
    # the assertions of kafka_proto_test.go:64-85 (the Go test builds Interfaces by hand with directions 0 and 1; here
    # they come out of NewRecord's rule: first-seen direction, then the observed list)
    assert m.eth_protocol == 3 and m.direction == 1 and len(m.dup_list) == 2
    assert (m.dup_list[0].interface, m.dup_list[0].direction) == ("veth0", 1)
    assert (m.dup_list[1].interface, m.dup_list[1].direction) == ("abcde", 1)
    assert m.data_link.src_mac == 0xaabbccddeeff and m.data_link.dst_mac == 0x112233445566
    assert m.network.src_addr.ipv4 == 0xC0010203 and m.network.dst_addr.ipv4 == 0x7F030201
    assert (m.transport.src_port, m.transport.dst_port, m.transport.protocol, m.icmp_type) == (4321, 1234, 210, 8)
    assert m.time_flow_start.seconds * 10**9 + m.time_flow_start.nanos == now - 5_000_000_000
    assert m.time_flow_end.seconds * 10**9 + m.time_flow_end.nanos == now
    assert (m.bytes, m.packets, m.flags) == (789, 987, 1)
    assert bytes(keys[0][:16]) == bytes(r["dst_ip"][0]) and bytes(keys[0][16:]) == bytes(r["src_ip"][0])   # 127.3.2.1 < 192.1.2.3
    # TestIdenticalKeys: swapping the addresses gives the same key
    r2 = r.copy(); r2["src_ip"], r2["dst_ip"] = r["dst_ip"].copy(), r["src_ip"].copy()
    _, _, keys2 = pb_encode(lib, h, r2, now=now, mono=mono, ifaces=ifaces)
    assert np.array_equal(keys, keys2)
    # and byte for byte what the protobuf library writes for the same message
    want = PB.flow_to_pb(r.tobytes(), None, None, now, mono, bytes(10) + b"\xff\xff" + bytes([10, 1, 2, 3]), True, ifaces)
    assert raw == want.SerializeToString()


def check_batch(lib, h, n=3000, wrap=False):
    from test_gpu_features import make_add, make_dns
    rng = np.random.default_rng(5)
    recs = gen_host(seed=91, n=n, n_keys=n, dist=0, varying=1).copy()      # every descriptor field varies; ~10 % IPv6
    recs.view(O.REC_DTYPE)["start"] = rng.integers(1, 5_000_000_000_000, n).reshape(-1, 1)
    recs.view(O.REC_DTYPE)["end"] = rng.integers(1, 6_000_000_000_000, n).reshape(-1, 1)     # some lie after mono_now
    recs.view(O.REC_DTYPE)["bytes"][::7] = rng.integers(1 << 40, 1 << 63, len(recs[::7])).reshape(-1, 1)
    keys40 = recs[:, :40]
    dns = make_dns(rng, keys40, n).view(np.uint8).reshape(n, 104)[:, 40:].copy()
    add = make_add(rng, keys40, n).view(np.uint8).reshape(n, 72)[:, 40:].copy()
    names = [b"\x03www\x07example\x03com\x00", b"\x05hello\x00", b"\x03abc\xc0\x0c", b"\x1fabcdefghijklmnopqrstuvwxyz01234", b"\x00", b"\x09short"]
    dv = dns.view(O.DNS_DTYPE).reshape(-1)
    for i in range(n):
        nm = names[i % len(names)]
        dv["name"][i] = list(nm[:32].ljust(32, b"\x00"))
    add.view(O.ADD_DTYPE).reshape(-1)["ipsec_ret"][::5] = -7
    drops = np.zeros(n, dtype=O.DROP_DTYPE)
    drops["bytes"], drops["packets"] = rng.integers(0, 65536, n), rng.integers(0, 65536, n)
    drops["cause"], drops["flags"], drops["state"] = rng.integers(0, 1 << 32, n), rng.integers(0, 1 << 16, n), rng.integers(0, 256, n)
    drops = drops.view(np.uint8).reshape(n, 32)
    present = rng.integers(0, 8, n).astype(np.uint8)
    ifaces = [(i, bytes([2, 0, 0, 0, 0, i]), f"eth{i}", "default" if i % 3 == 0 else "") for i in range(1, 12)]
    ifaces += [(3, bytes([2, 0, 0, 0, 9, 9]), "ens3-alt", "blue")]            # two rows for ifindex 3: disambiguated by MAC
    agent6 = bytes(range(0x20, 0x30))
    raw, offs, keys = pb_encode(lib, h, recs, dns, add, present, drops, ifaces=ifaces, agent_ip=agent6, agent_v4=0, wrap=wrap)
    now, mono = 1_700_000_000_123_456_789, 5_000_000_000_000
    for i in range(n):
        want = PB.flow_to_pb(recs[i].tobytes(), dns[i].tobytes() if present[i] & 1 else None, add[i].tobytes() if present[i] & 2 else None,
                             now, mono, agent6, False, ifaces, drops[i].tobytes() if present[i] & 4 else None).SerializeToString()
        got = raw[offs[i]:offs[i + 1]]
        if wrap:
            one = PB.Records(); one.ParseFromString(got)
            assert len(one.entries) == 1 and one.entries[0].SerializeToString() == want, i
        else:
            assert got == want, (i, got.hex(), want.hex())
        assert bytes(keys[i]) == PB.flow_key(recs[i].tobytes())
    if wrap:                                                               # the whole output is one pbflow.Records
        allm = PB.Records(); allm.ParseFromString(raw)
        assert len(allm.entries) == n


# ---------------------------------------------------------------- CPU: kernels on the engine emulation
@pytest.fixture
def emul_engine(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    eng = fa.FlowAggEngine(64)
    try:
        yield engine_emul, eng._h
    finally:
        eng.close()


def test_golden_record_of_the_reference_on_the_emulation(emul_engine):
    check_golden(*emul_engine)


def test_random_batch_matches_protobuf_on_the_emulation(emul_engine):
    check_batch(*emul_engine, n=700)
    check_batch(*emul_engine, n=300, wrap=True)


# ---------------------------------------------------------------- GPU
@pytest.mark.gpu
def test_golden_record_of_the_reference_gpu():
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(64) as eng:
        check_golden(fa.lib(), eng._h)


@pytest.mark.gpu
def test_random_batch_matches_protobuf_gpu():
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(64) as eng:
        check_batch(fa.lib(), eng._h, n=20_000)
        check_batch(fa.lib(), eng._h, n=5_000, wrap=True)


@pytest.mark.gpu
def test_evicted_flows_go_straight_to_protobuf_from_device_memory():
    """fa_evict into device buffers, fa_pb_encode from them: nothing but the wire bytes crosses PCIe."""
    import torch
    import netobserv_ebpf_agent_b200 as fa
    recs = gen_host(seed=92, n=50_000, n_keys=4_000, dist=1)
    with fa.FlowAggEngine(1 << 13) as eng:
        eng.ingest(recs)
        n = eng.live_flows()
        dev = torch.empty(n * 144, dtype=torch.uint8, device="cuda")
        assert eng.evict_into(dev, n) == n
        flows = dev.cpu().numpy().reshape(n, 144)
        p_raw, offs, _ = pb_encode(fa.lib(), eng._h, flows)
        from netobserv_ebpf_agent_b200._lib import PbParams
        p = PbParams(now_unix_ns=1_700_000_000_123_456_789, mono_now_ns=5_000_000_000_000, agent_ip_is_v4=1)
        for i, b in enumerate(bytes(10) + b"\xff\xff" + bytes([10, 1, 2, 3])):
            p.agent_ip[i] = b
        out = torch.empty(len(p_raw), dtype=torch.uint8, device="cuda")
        ln = C.c_size_t(0)
        rc = fa.lib().fa_pb_encode(eng._h, C.c_void_p(dev.data_ptr()), None, None, None, None, n, C.byref(p), C.c_void_p(out.data_ptr()),
                                   out.numel(), None, None, C.byref(ln))
        assert rc == 0 and ln.value == len(p_raw)
        assert out.cpu().numpy().tobytes() == p_raw

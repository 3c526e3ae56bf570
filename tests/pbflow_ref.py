"""Checker for K8 (fa_pb_encode): the reference's proto/flow.proto restated as runtime descriptors for the Python
protobuf library, and model.NewRecord + pbflow.FlowToPB (pkg/model/record.go:82-159, pkg/pbflow/proto.go:39-149)
restated over the 144-byte flow record (+ optional dns / additional blocks).

TEST INFRASTRUCTURE ONLY (like oracle/): the product never imports this.  The message classes come from descriptors
built here (protoc is not in the image); serialisation is the protobuf library's own, which writes fields in
field-number order like protobuf-go does."""
import numpy as np
from google.protobuf import descriptor_pb2, descriptor_pool, duration_pb2, message_factory, timestamp_pb2  # noqa: F401

import oracle_lib as O

F = descriptor_pb2.FieldDescriptorProto
U32, U64, I32, BOOL, STR, BYT, FX32, MSG, ENUM = (F.TYPE_UINT32, F.TYPE_UINT64, F.TYPE_INT32, F.TYPE_BOOL, F.TYPE_STRING,
                                                  F.TYPE_BYTES, F.TYPE_FIXED32, F.TYPE_MESSAGE, F.TYPE_ENUM)


def _build():
    fd = descriptor_pb2.FileDescriptorProto(name="flow_restated.proto", package="pbflow", syntax="proto3",
                                            dependency=["google/protobuf/timestamp.proto", "google/protobuf/duration.proto"])
    en = fd.enum_type.add(name="Direction")
    en.value.add(name="INGRESS", number=0); en.value.add(name="EGRESS", number=1)

    def msg(name, fields, oneof=None):
        m = fd.message_type.add(name=name)
        if oneof:
            m.oneof_decl.add(name=oneof)
        for fname, num, typ, *rest in fields:
            f = m.field.add(name=fname, number=num, type=typ, label=F.LABEL_OPTIONAL)
            for r in rest:
                if r == "repeated":
                    f.label = F.LABEL_REPEATED
                elif r == "oneof":
                    f.oneof_index = 0
                else:
                    f.type_name = r
        return m
    msg("IP", [("ipv4", 1, FX32, "oneof"), ("ipv6", 2, BYT, "oneof")], oneof="ip_family")                  # flow.proto:90-95
    msg("DataLink", [("src_mac", 1, U64), ("dst_mac", 2, U64)])                                             # :79-82
    msg("Network", [("src_addr", 1, MSG, ".pbflow.IP"), ("dst_addr", 2, MSG, ".pbflow.IP"), ("dscp", 3, U32)])   # :84-88
    msg("Transport", [("src_port", 1, U32), ("dst_port", 2, U32), ("protocol", 3, U32)])                    # :97-103
    msg("DupMapEntry", [("interface", 1, STR), ("direction", 2, ENUM, ".pbflow.Direction"), ("udn", 3, STR)])   # :21-25
    msg("Record", [                                                                                         # :31-77
        ("eth_protocol", 1, U32), ("direction", 2, ENUM, ".pbflow.Direction"),
        ("time_flow_start", 3, MSG, ".google.protobuf.Timestamp"), ("time_flow_end", 4, MSG, ".google.protobuf.Timestamp"),
        ("data_link", 5, MSG, ".pbflow.DataLink"), ("network", 6, MSG, ".pbflow.Network"), ("transport", 7, MSG, ".pbflow.Transport"),
        ("bytes", 8, U64), ("packets", 9, U64), ("interface", 10, STR), ("duplicate", 11, BOOL),
        ("agent_ip", 12, MSG, ".pbflow.IP"), ("flags", 13, U32), ("icmp_type", 14, U32), ("icmp_code", 15, U32),
        ("pkt_drop_bytes", 16, U64), ("pkt_drop_packets", 17, U64), ("pkt_drop_latest_flags", 18, U32),
        ("pkt_drop_latest_state", 19, U32), ("pkt_drop_latest_drop_cause", 20, U32),
        ("dns_id", 21, U32), ("dns_flags", 22, U32), ("dns_latency", 23, MSG, ".google.protobuf.Duration"),
        ("time_flow_rtt", 24, MSG, ".google.protobuf.Duration"), ("dns_errno", 25, U32),
        ("dup_list", 26, MSG, ".pbflow.DupMapEntry", "repeated"),
        ("sampling", 29, U32), ("ipsec_encrypted", 30, U32), ("ipsec_encrypted_ret", 31, I32), ("dns_name", 32, STR),
        ("ssl_version", 33, U32), ("ssl_mismatch", 34, BOOL), ("tls_types", 35, U32), ("tls_cipher_suite", 36, U32),
        ("tls_key_share", 37, U32)])
    msg("Records", [("entries", 1, MSG, ".pbflow.Record", "repeated")])                                     # :17-19
    pool = descriptor_pool.Default()
    pool.Add(fd)
    get = message_factory.GetMessageClass
    return get(pool.FindMessageTypeByName("pbflow.Record")), get(pool.FindMessageTypeByName("pbflow.Records"))


Record, Records = _build()


def dns_dotted(raw32):
    """utils.DNSRawNameToDotted (pkg/utils/dns.go:20-60)."""
    b = bytes(raw32)
    b = b[: b.index(0)] if 0 in b else b
    out, i = [], 0
    while i < len(b):
        ln = b[i]
        if ln == 0 or (ln & 0xC0) == 0xC0:
            break
        i += 1
        if i + ln > len(b):
            break
        out.append(b[i:i + ln]); i += ln
    return b".".join(out)


def iface_lookup(ifaces, if_index, mac):
    """interfaceNamer over ifaces.Registerer.ifaceCacheLookup (registerer.go:153-190); ifaces: list of
    (if_index, mac bytes, name, udn) rows."""
    rows = [r for r in ifaces if r[0] == if_index]
    if not rows:
        return "unknown", ""
    if len(rows) == 1:
        return rows[0][2], rows[0][3]
    for r in rows:
        if bytes(r[1]) == bytes(mac):
            return r[2], r[3]
    return rows[0][2], rows[0][3]


def _secnanos(msg, ns):
    # Go: Unix() floors, Nanosecond() in [0, 1e9)
    msg.seconds, msg.nanos = int(ns) // 1_000_000_000, int(ns) % 1_000_000_000


def _duration(msg, ns_u64):
    d = int(ns_u64)
    if d >= 1 << 63:
        d -= 1 << 64                                  # time.Duration(uint64) is an int64 cast
    q = abs(d) // 1_000_000_000 * (1 if d >= 0 else -1)  # durationpb.New truncates towards zero
    msg.seconds, msg.nanos = q, d - q * 1_000_000_000


def _ip(msg, ip16, v4):
    if v4:
        msg.ipv4 = int.from_bytes(bytes(ip16[12:16]), "big")
    else:
        msg.ipv6 = bytes(ip16)


def flow_to_pb(rec144, dns64, add32, now_unix_ns, mono_now_ns, agent_ip16, agent_is_v4, ifaces, drop32=None):
    """NewRecord + FlowToPB for one evicted flow -> pbflow.Record message."""
    r = np.frombuffer(bytes(rec144), dtype=O.REC_DTYPE)[0]
    pb = Record()
    pb.eth_protocol = int(r["eth"]); pb.direction = int(r["direction"])
    for field, mono in (("time_flow_start", r["start"]), ("time_flow_end", r["end"])):
        delta = (int(mono_now_ns) - int(mono)) & ((1 << 64) - 1)
        if delta >= 1 << 63:
            delta -= 1 << 64
        _secnanos(getattr(pb, field), int(now_unix_ns) - delta)
    pb.data_link.src_mac = int.from_bytes(bytes(r["src_mac"]), "big"); pb.data_link.dst_mac = int.from_bytes(bytes(r["dst_mac"]), "big")
    v6 = int(r["eth"]) == 0x86DD
    pb.network.dscp = int(r["dscp"])
    _ip(pb.network.src_addr, r["src_ip"], not v6); _ip(pb.network.dst_addr, r["dst_ip"], not v6)
    pb.transport.src_port, pb.transport.dst_port, pb.transport.protocol = int(r["src_port"]), int(r["dst_port"]), int(r["proto"])
    pb.bytes, pb.packets = int(r["bytes"]), int(r["packets"])
    _ip(pb.agent_ip, agent_ip16, agent_is_v4)
    pb.flags, pb.icmp_type, pb.icmp_code = int(r["flags"]), int(r["icmp_type"]), int(r["icmp_code"])
    if drop32 is not None:                              # proto.go:86-92
        p = np.frombuffer(bytes(drop32), dtype=O.DROP_DTYPE)[0]
        pb.pkt_drop_bytes, pb.pkt_drop_packets = int(p["bytes"]), int(p["packets"])
        pb.pkt_drop_latest_flags, pb.pkt_drop_latest_state, pb.pkt_drop_latest_drop_cause = int(p["flags"]), int(p["state"]), int(p["cause"])
    if dns64 is not None:
        d = np.frombuffer(bytes(dns64), dtype=O.DNS_DTYPE)[0]
        pb.dns_id, pb.dns_flags, pb.dns_errno = int(d["id"]), int(d["flags"]), int(d["errno"])
        name = dns_dotted(d["name"])
        if name:
            pb.dns_name = name.decode("latin-1")       # tests use ASCII names
        if int(d["latency"]):
            _duration(pb.dns_latency, d["latency"])
    rtt = 0
    if add32 is not None:
        a = np.frombuffer(bytes(add32), dtype=O.ADD_DTYPE)[0]
        rtt = int(a["rtt"])
        pb.ipsec_encrypted_ret = int(a["ipsec_ret"])
        if a["ipsec_enc"]:
            pb.ipsec_encrypted = 1
    _duration(pb.time_flow_rtt, rtt)
    pb.time_flow_rtt.SetInParent()
    lmac = r["dst_mac"] if int(r["direction"]) == 0 else r["src_mac"]
    intfs = [(int(r["if_index"]), int(r["direction"]))] + [(int(r["obs_intf"][i]), int(r["obs_dir"][i])) for i in range(min(int(r["nb_obs"]), 6))]
    for idx, direction in intfs:
        name, udn = iface_lookup(ifaces, idx, lmac)
        e = pb.dup_list.add()
        e.interface, e.direction, e.udn = name, direction, udn
    pb.sampling = int(r["sampling"])
    pb.ssl_version, pb.tls_types, pb.tls_cipher_suite, pb.tls_key_share = int(r["ssl_version"]), int(r["tls_types"]), int(r["cipher"]), int(r["key_share"])
    pb.ssl_mismatch = bool(int(r["misc"]) & 0x01)
    for sub in ("time_flow_start", "time_flow_end", "data_link", "network", "transport", "agent_ip"):
        getattr(pb, sub).SetInParent()                 # FlowToPB always allocates them
    pb.network.src_addr.SetInParent(); pb.network.dst_addr.SetInParent()
    return pb


def flow_key(rec144):
    """getFlowKey (pkg/exporter/kafka_proto.go:37-47)."""
    b = bytes(rec144)
    src, dst = b[0:16], b[16:32]
    return src + dst if src <= dst else dst + src

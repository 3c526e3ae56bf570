"""The flow filter in front of fa_ingest_snaps: bpf/flows_filter.h:14-255 (LPM lookup of the source, then the destination
address; protocol / port / ICMP / TCP-flag / direction / peer-CIDR conditions) + check_and_do_flow_filtering
(bpf/utils.h:179-222: ACCEPT / REJECT, the no-match rule, the three global counters).

Source-pinned only (the reference has no unit test of the eBPF matching): the oracle restatement is checked against
cases worked out by hand from the source, the engine against the oracle on random rule sets and packets."""
import numpy as np
import pytest

import oracle_lib as O
from test_snaps import M1, M2, V4A, V4B, V6A, V6B, eth, icmp, ipv4, ipv6, random_snaps, snap, tcp, udp


def rules_of(*specs):
    import netobserv_ebpf_agent_b200 as fa
    r = np.zeros(len(specs), dtype=fa.FILTER_RULE_DTYPE)
    r["direction"], r["action"] = 2, 2                      # MAX_DIRECTION, MAX_FILTER_ACTIONS
    for i, sp in enumerate(specs):
        ip = sp.pop("cidr")
        addr, plen = ip
        r["ip"][i, : len(addr)] = addr
        r["prefix_len"][i] = plen
        for k, v in sp.items():
            r[k][i] = v
    return r


def cidrs_of(*specs):
    import netobserv_ebpf_agent_b200 as fa
    c = np.zeros(len(specs), dtype=fa.FILTER_CIDR_DTYPE)
    for i, (addr, plen) in enumerate(specs):
        c["ip"][i, : len(addr)] = addr
        c["prefix_len"][i] = plen
    return c


def verdict(frame, rules, peers=None, **kw):
    """-> (skipped, counters, sampling of the record) for one packet through the oracle."""
    s = snap(frame, 104, **kw)
    recs, src, ctr = O.parse_snaps_filtered(s, 104, rules, peers)
    samp = int(recs.view(O.REC_DTYPE).reshape(-1)["sampling"][0]) if len(recs) else None
    return len(recs) == 0, ctr.tolist(), samp


T4 = eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 6) + tcp(40000, 443, 0x02)         # 10.1.2.3:40000 -> 192.168.7.9:443 SYN
U6 = eth(M2, M1, 0x86DD) + ipv6(V6A, V6B, 17) + udp(53, 5353)
I4 = eth(M2, M1, 0x0800) + ipv4(V4A, V4B, 1) + icmp(8, 0)


def test_oracle_accept_reject_and_the_nomatch_rule():
    acc = rules_of(dict(cidr=([10, 1, 0, 0], 16), action=0))
    assert verdict(T4, acc) == (False, [1, 0, 0], 1)                          # source address inside the ACCEPT rule
    rej = rules_of(dict(cidr=([10, 1, 0, 0], 16), action=1))
    assert verdict(T4, rej) == (True, [0, 1, 0], None)
    other = rules_of(dict(cidr=([172, 16, 0, 0], 12), action=0))
    assert verdict(T4, other) == (True, [0, 0, 1], None)                      # no rule at all: action stays MAX -> skipped
    # the destination address is tried when the source finds nothing
    dst = rules_of(dict(cidr=([192, 168, 0, 0], 16), action=0))
    assert verdict(T4, dst) == (False, [1, 0, 0], 1)
    # an ACCEPT rule whose conditions fail: no match, and "we have accept rule but no match" -> skipped
    a443 = rules_of(dict(cidr=([10, 1, 0, 0], 16), action=0, dst_port_start=80))
    assert verdict(T4, a443) == (True, [0, 0, 1], None)
    # a REJECT rule whose conditions fail: no match, but the packet goes on (utils.h:216-218)
    r80 = rules_of(dict(cidr=([10, 1, 0, 0], 16), action=1, dst_port_start=80))
    assert verdict(T4, r80) == (False, [0, 0, 1], 1)


def test_oracle_longest_prefix_ports_flags_direction_sampling():
    # /24 beats /16 beats /0
    r = rules_of(dict(cidr=([0, 0, 0, 0], 0), action=1), dict(cidr=([10, 1, 0, 0], 16), action=1),
                 dict(cidr=([10, 1, 2, 0], 24), action=0, sample=7))
    assert verdict(T4, r) == (False, [1, 0, 0], 7)                            # the rule's sample becomes the packet's sampling
    # port forms: single (start, end == 0), pair (port1 / port2), range, generic (either side)
    base = dict(cidr=([10, 1, 2, 3], 32), action=0)
    assert verdict(T4, rules_of(dict(base, dst_port_start=443)))[0] is False
    assert verdict(T4, rules_of(dict(base, dst_port1=80, dst_port2=443)))[0] is False
    assert verdict(T4, rules_of(dict(base, dst_port_start=400, dst_port_end=500)))[0] is False
    assert verdict(T4, rules_of(dict(base, dst_port_start=444, dst_port_end=500)))[0] is True
    assert verdict(T4, rules_of(dict(base, src_port_start=40000)))[0] is False
    assert verdict(T4, rules_of(dict(base, src_port_start=1, src_port_end=1024)))[0] is True
    assert verdict(T4, rules_of(dict(base, port1=40000)))[0] is False and verdict(T4, rules_of(dict(base, port_start=443)))[0] is False
    assert verdict(T4, rules_of(dict(base, port_start=1000, port_end=2000)))[0] is True
    # protocol, TCP flags (the set_flags value), direction
    assert verdict(T4, rules_of(dict(base, protocol=17)))[0] is True and verdict(T4, rules_of(dict(base, protocol=6)))[0] is False
    assert verdict(T4, rules_of(dict(base, tcp_flags=0x02)))[0] is False and verdict(T4, rules_of(dict(base, tcp_flags=0x10)))[0] is True
    assert verdict(T4, rules_of(dict(base, direction=1)), direction=1)[0] is False
    assert verdict(T4, rules_of(dict(base, direction=1)), direction=0)[0] is True
    assert verdict(T4, rules_of(dict(base, filter_drops=1)))[0] is True       # needs a drop reason, flow_monitor has none
    # ICMP type / code
    ib = dict(cidr=([10, 1, 2, 3], 32), action=0)
    assert verdict(I4, rules_of(dict(ib, icmp_type=8)))[0] is False and verdict(I4, rules_of(dict(ib, icmp_type=3)))[0] is True
    assert verdict(I4, rules_of(dict(ib, icmp_type=8, icmp_code=1)))[0] is True
    # IPv6 keys are 128 bits; an IPv4 rule with the same leading bytes is a different prefix
    r6 = rules_of(dict(cidr=(V6A[:8] + [0] * 8, 64), action=0))
    assert verdict(U6, r6) == (False, [1, 0, 0], 1)
    assert verdict(U6, rules_of(dict(cidr=([10, 1, 0, 0], 16), action=0))) == (True, [0, 0, 1], None)


def test_oracle_peer_cidr():
    r = rules_of(dict(cidr=([10, 1, 2, 3], 32), action=0, do_peer_cidr_lookup=1))
    assert verdict(T4, r, cidrs_of(([192, 168, 7, 0], 24)))[0] is False       # source rule, the peer is the destination
    assert verdict(T4, r, cidrs_of(([192, 169, 0, 0], 16)))[0] is True
    rd = rules_of(dict(cidr=([192, 168, 7, 9], 32), action=0, do_peer_cidr_lookup=1))
    assert verdict(T4, rd, cidrs_of(([10, 0, 0, 0], 8)))[0] is False          # destination rule, the peer is the source


def random_rules(rng, n):
    specs = []
    for _ in range(n):
        v6 = rng.random() < 0.3
        if v6:
            plen = int(rng.choice([0, 64, 120, 128]))
            addr = (V6A if rng.random() < 0.5 else V6B)[:15] + [int(rng.integers(1, 41))]
        else:
            plen = int(rng.choice([0, 8, 24, 30, 32]))
            addr = (V4A if rng.random() < 0.5 else V4B)[:3] + [int(rng.integers(1, 41))]
        nb = (plen + 7) // 8
        addr = [a if i < nb else 0 for i, a in enumerate(addr)]
        if plen % 8 and nb:
            addr[nb - 1] &= (0xFF << (8 - plen % 8)) & 0xFF
        sp = dict(cidr=(addr, plen), action=int(rng.choice([0, 1, 2])), direction=int(rng.choice([0, 1, 2, 2])),
                  protocol=int(rng.choice([0, 0, 6, 17, 1])), sample=int(rng.choice([0, 0, 5])),
                  do_peer_cidr_lookup=int(rng.random() < 0.2))
        form = rng.integers(0, 6)
        if form == 1:
            sp["dst_port_start"] = int(rng.choice([53, 80, 443]))
        elif form == 2:
            sp["port1"], sp["port2"] = 53, 2000
        elif form == 3:
            sp["src_port_start"], sp["src_port_end"] = 1000, 3000
        elif form == 4:
            sp["tcp_flags"] = int(rng.choice([0x02, 0x10, 0x100]))
        elif form == 5:
            sp["icmp_type"] = int(rng.integers(0, 4))
        specs.append(sp)
    # the trie holds one entry per (prefix, length)
    seen, uniq = set(), []
    for sp in specs:
        k = (tuple(sp["cidr"][0]), sp["cidr"][1])
        if k not in seen:
            seen.add(k); uniq.append(sp)
    return rules_of(*uniq)


def check(n, stride, seed, max_batch):
    import netobserv_ebpf_agent_b200 as fa
    rng = np.random.default_rng(seed)
    snaps = random_snaps(rng, n, stride)
    rules = random_rules(rng, 12)
    peers = cidrs_of(([10, 1, 2, 0], 25), ([192, 168, 7, 16], 28), (V6B[:15] + [0], 120))
    recs, src, ctr = O.parse_snaps_filtered(snaps, stride, rules, peers)
    plain, _ = O.parse_snaps(snaps, stride)
    assert 0 < len(recs) < len(plain) and int(ctr.sum()) == len(plain) and int((ctr > 0).sum()) >= 2
    acc = O.Accounter(1 << 16); acc.account(recs); want = O.sort_records(acc.evict()); acc.close()
    with fa.FlowAggEngine(1 << 16, max_batch=max_batch) as eng:
        eng.set_flow_filter(rules, peers)
        rc, took = eng.ingest_snaps(snaps.reshape(-1), stride)
        assert rc == 0 and took == n
        st = eng.stats()
        assert [st["filter_accept"], st["filter_reject"], st["filter_nomatch"]] == ctr.tolist()
        assert st["snaps_discarded"] == n - len(plain) and st["records_ingested"] == len(recs)
        assert np.array_equal(O.sort_records(eng.evict()), want)
        eng.set_flow_filter()                                                  # filter off again
        rc, took = eng.ingest_snaps(snaps.reshape(-1), stride)
        assert rc == 0 and eng.stats()["records_ingested"] == len(recs) + len(plain)


def test_filter_on_the_emulation(engine_emul):
    check(n=1_500, stride=104, seed=5, max_batch=1_024)
    check(n=700, stride=88, seed=6, max_batch=4_096)


def test_too_many_rules_are_refused(engine_emul):
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(64, max_batch=256) as eng:
        with pytest.raises(fa.FlowAggError) as ei:
            eng.set_flow_filter(np.zeros(17, dtype=fa.FILTER_RULE_DTYPE))
        assert ei.value.code == -22


@pytest.mark.gpu
def test_filter_gpu():
    check(n=40_000, stride=104, seed=7, max_batch=16_384)
    check(n=9_000, stride=88, seed=8, max_batch=1 << 20)

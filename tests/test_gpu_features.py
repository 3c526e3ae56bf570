"""K6 feature folds (RTT / IPsec / DNS) through the C ABI against the oracle's LookupAndDeleteMap view."""
import numpy as np
import pytest

import oracle_lib as O
from common import gen_host

pytestmark = pytest.mark.gpu


def keys_of(seed, n_keys):
    import netobserv_ebpf_agent_b200 as fa
    p = fa.GenParams(seed=seed, n_keys=n_keys, dist=0, zipf_s_milli=1100, t0_ns=0, varying_desc=0)
    return np.stack([fa.gen_key(p, r) for r in range(n_keys)])


def make_dns(rng, keys, n):
    r = np.zeros(n, dtype=O.DNSREC_DTYPE)
    r["id"] = keys[rng.integers(0, len(keys), n)]
    d = r["dns"]
    d["start"] = np.where(rng.random(n) < 0.1, 0, rng.integers(1, 1 << 40, n))
    d["end"] = np.where(rng.random(n) < 0.1, 0, rng.integers(1, 1 << 40, n))
    d["latency"] = np.where(rng.random(n) < 0.2, 0, rng.integers(1, 1 << 30, n))
    d["id"] = np.where(rng.random(n) < 0.3, 0, rng.integers(1, 1 << 16, n))
    d["flags"] = rng.integers(0, 1 << 16, n)
    d["eth"] = np.where(rng.random(n) < 0.4, 0, rng.choice([0x0800, 0x86DD], n))
    d["errno"] = np.where(rng.random(n) < 0.5, 0, rng.integers(1, 100, n))
    d["name"] = rng.integers(97, 123, (n, 32))
    d["pad"] = rng.integers(0, 256, n)                 # ignored on input, zero on output
    r["dns"] = d
    return r


def make_add(rng, keys, n):
    r = np.zeros(n, dtype=O.ADDREC_DTYPE)
    r["id"] = keys[rng.integers(0, len(keys), n)]
    a = r["add"]
    a["start"] = np.where(rng.random(n) < 0.1, 0, rng.integers(1, 1 << 40, n))
    a["end"] = np.where(rng.random(n) < 0.1, 0, rng.integers(1, 1 << 40, n))
    a["rtt"] = np.where(rng.random(n) < 0.2, 0, rng.integers(1, 1 << 34, n))
    a["ipsec_ret"] = rng.integers(-3, 4, n)
    a["ipsec_enc"] = rng.integers(0, 2, n)
    a["eth"] = np.where(rng.random(n) < 0.4, 0, rng.choice([0x0800, 0x86DD], n))
    a["pad"] = rng.integers(0, 256, n)
    r["add"] = a
    return r


def compare(eng, om):
    g_recs, g_dns, g_add, g_pres = eng.evict(features=True)
    o_recs, o_dns, o_add, o_pres = om.evict()
    gp, op = O.sort_perm(g_recs), O.sort_perm(o_recs)
    assert len(gp) == len(op)
    for name, g, o in (("records", g_recs, o_recs), ("dns", g_dns, o_dns), ("additional", g_add, o_add),
                       ("present", g_pres.reshape(-1, 1), o_pres.reshape(-1, 1))):
        gs, os_ = g[gp], o[op]
        if not np.array_equal(gs, os_):
            bad = np.nonzero((gs != os_).any(axis=1))[0]
            raise AssertionError(f"{name}: {len(bad)} of {len(gs)} differ; first {bad[0]}:\n gpu {gs[bad[0]].tolist()}\n ora {os_[bad[0]].tolist()}")


def test_feature_goldens_on_gpu():
    """pkg/model/flow_content_test.go:11-53,184-246 through fa_ingest / fa_ingest_dns / fa_ingest_additional."""
    import netobserv_ebpf_agent_b200 as fa
    from test_oracle_goldens import K1
    base = K1(start=10, end=20, packets=3)
    d = np.zeros(2, dtype=O.DNSREC_DTYPE)
    d["id"][:] = np.frombuffer(base.tobytes()[:40], dtype=np.uint8)
    d["dns"]["start"] = [25, 30]; d["dns"]["end"] = [25, 30]; d["dns"]["latency"] = [1000, 2000]
    d["dns"]["id"] = [1, 1]; d["dns"]["flags"] = [0b0011, 0b1001]
    a = np.zeros(4, dtype=O.ADDREC_DTYPE)
    a["id"][:] = np.frombuffer(base.tobytes()[:40], dtype=np.uint8)
    a["add"]["start"] = [25, 30, 30, 30]; a["add"]["end"] = [25, 30, 30, 30]
    a["add"]["rtt"] = [200, 1000, 800, 800]; a["add"]["ipsec_enc"] = [1, 0, 0, 0]; a["add"]["ipsec_ret"] = [0, 0, 5, 0]
    with fa.FlowAggEngine(100, flags=fa.FA_F_ENABLE_RTT | fa.FA_F_ENABLE_DNS) as eng:
        eng.ingest(base); eng.ingest_dns(d); eng.ingest_additional(a)
        recs, dns, add, pres = eng.evict(features=True)
    r = recs.view(O.REC_DTYPE).reshape(-1)[0]
    dn = dns.view(O.DNS_DTYPE).reshape(-1)[0]
    ad = add.view(O.ADD_DTYPE).reshape(-1)[0]
    assert pres[0] == 3 and (r["start"], r["end"], r["packets"]) == (10, 30, 3)
    assert (dn["start"], dn["end"], dn["latency"], dn["id"], dn["flags"]) == (25, 25, 2000, 1, 0b1011)
    assert (ad["start"], ad["end"], ad["rtt"], ad["ipsec_enc"], ad["ipsec_ret"]) == (25, 25, 1000, 0, 5)


@pytest.mark.parametrize("n_keys,n_base,n_feat", [(50, 2_000, 3_000), (5_000, 60_000, 40_000)])
def test_feature_folds_match_oracle(n_keys, n_base, n_feat):
    import netobserv_ebpf_agent_b200 as fa
    rng = np.random.default_rng(31)
    keys = keys_of(31, n_keys)
    extra = keys_of(32, n_keys // 2)                      # flows seen only through feature samples
    allk = np.concatenate([keys, extra])
    base = [gen_host(seed=31, n=n_base, n_keys=n_keys, first=i * n_base) for i in range(2)]
    dns = [make_dns(rng, allk, n_feat) for _ in range(2)]
    add = [make_add(rng, allk, n_feat) for _ in range(2)]
    om = O.FlowMap()
    with fa.FlowAggEngine(1 << 16, flags=fa.FA_F_ENABLE_RTT | fa.FA_F_ENABLE_DNS, max_batch=25_000) as eng:
        # interleave the three streams differently on both sides: only per-stream order matters
        eng.ingest_dns(dns[0]); eng.ingest(base[0]); eng.ingest_additional(add[0])
        eng.ingest(base[1]); eng.ingest_additional(add[1]); eng.ingest_dns(dns[1])
        om.account(base[0]); om.account(base[1])
        om.fold_dns(dns[0]); om.fold_dns(dns[1]); om.fold_additional(add[0]); om.fold_additional(add[1])
        compare(eng, om)
        # second generation after the eviction: everything was deleted
        eng.ingest_additional(add[0][:1000]); om.fold_additional(add[0][:1000])
        compare(eng, om)
        assert eng.live_flows() == 0


def test_feature_only_then_base_with_varying_descriptors():
    """A flow created by feature samples adopts its first base record whole (account.go:95) — also under re-fold."""
    import netobserv_ebpf_agent_b200 as fa
    rng = np.random.default_rng(33)
    base = gen_host(seed=33, n=30_000, n_keys=400, dist=1, varying=1)
    keys = np.unique(base[:, :40], axis=0)
    add = make_add(rng, keys, 5_000)
    om = O.FlowMap()
    with fa.FlowAggEngine(1 << 12, flags=fa.FA_F_ENABLE_RTT, max_batch=8_000) as eng:
        eng.ingest_additional(add); eng.ingest(base)
        om.fold_additional(add); om.account(base)
        g_recs, _, g_add, g_pres = eng.evict(features=True)
        o_recs, _, o_add, o_pres = om.evict()
        gp, op = O.sort_perm(g_recs), O.sort_perm(o_recs)
        assert np.array_equal(g_recs[gp], o_recs[op]) and np.array_equal(g_add[gp], o_add[op])
        assert np.array_equal(g_pres[gp], o_pres[op])


def test_feature_disabled_errors():
    import netobserv_ebpf_agent_b200 as fa
    with fa.FlowAggEngine(100) as eng:
        with pytest.raises(fa.FlowAggError):
            eng.ingest_dns(np.zeros(1, dtype=O.DNSREC_DTYPE))
        with pytest.raises(fa.FlowAggError):
            eng.ingest_additional(np.zeros(1, dtype=O.ADDREC_DTYPE))

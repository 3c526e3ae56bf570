#!/usr/bin/env python
"""Throughput of the secondary kernels (not the headline bench): fused count-min + HyperLogLog (BASELINE config 3
shape, scaled), RTT / DNS feature folds (config 5 shape, scaled), K2 evict.  Prints one JSON line per measurement."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import netobserv_ebpf_agent_b200 as fa  # noqa: E402

REC = 144
dev = torch.device("cuda", 0)
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)


def timed(fn, n_iter):
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for i in range(n_iter):
        fn(i)
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / 1e3


def sketch_bench(n_keys=100_000_000, B=1 << 24, steps=12):
    eng = fa.FlowAggEngine(1 << 27, flags=fa.FA_F_ENABLE_SKETCH, cms_log2_width=20, cms_depth=4, hll_precision=14,
                           max_batch=B, cuda_stream=stream.cuda_stream)
    gp = fa.GenParams(seed=3, n_keys=n_keys, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    ring = []
    for i in range(4):
        t = torch.empty(B * REC, dtype=torch.uint8, device=dev)
        eng.gen_records(gp, i * B, B, t)
        ring.append(t)
    eng.sync()
    for i in range(2):
        eng.ingest(ring[i % 4].data_ptr(), B)
    dt = timed(lambda i: eng.ingest(ring[i % 4].data_ptr(), B), steps)
    n = (steps + 2) * B
    flows = eng.live_flows()
    est = eng.hll_estimate()
    print(json.dumps({"bench": "K1 + fused count-min(w=2^20,d=4) + HLL(p=14)", "workload": f"{n_keys} Zipf-1.1 keys",
                      "Mpkts_s": B * steps / dt / 1e6, "records": n, "distinct_exact": flows, "hll_estimate": est,
                      "hll_rel_err": abs(est - flows) / flows}), flush=True)
    eng.close()


def feature_bench(n_keys=1_000_000, B=1 << 22, steps=8):
    from test_gpu_features import keys_of, make_add, make_dns
    import oracle_lib as O
    eng = fa.FlowAggEngine(1 << 24, flags=fa.FA_F_ENABLE_RTT | fa.FA_F_ENABLE_DNS, max_batch=B, cuda_stream=stream.cuda_stream)
    rng = np.random.default_rng(5)
    gp = fa.GenParams(seed=5, n_keys=n_keys, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    base = torch.empty(B * REC, dtype=torch.uint8, device=dev)
    eng.gen_records(gp, 0, B, base)
    eng.sync()
    keys = base.cpu().numpy().reshape(-1, REC)[:200_000, :40]
    dns = torch.from_numpy(O.as_bytes(make_dns(rng, keys, B // 4)).copy()).to(dev)
    add = torch.from_numpy(O.as_bytes(make_add(rng, keys, B // 4)).copy()).to(dev)
    eng.ingest(base.data_ptr(), B)
    eng.ingest_dns(dns); eng.ingest_additional(add)
    t_dns = timed(lambda i: eng.ingest_dns(dns), steps)
    t_add = timed(lambda i: eng.ingest_additional(add), steps)
    t0 = time.perf_counter()
    out = eng.evict(features=True)
    t_ev = time.perf_counter() - t0
    print(json.dumps({"bench": "K6 feature folds", "dns_Msamples_s": (B // 4) * steps / t_dns / 1e6,
                      "additional_Msamples_s": (B // 4) * steps / t_add / 1e6, "flows": len(out[0]),
                      "evict_with_features_to_host_s": t_ev}), flush=True)
    eng.close()


def kmap_bench(n_keys=1_000_000, B=1 << 24, steps=12):
    """KERNEL_MAP mode (kmap.cu) on a 1 M-key Zipf stream."""
    eng = fa.FlowAggEngine(1 << 24, mode=fa.FA_MODE_KERNEL_MAP, max_batch=B, cuda_stream=stream.cuda_stream)
    gp = fa.GenParams(seed=2, n_keys=n_keys, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    ring = []
    for i in range(4):
        t = torch.empty(B * REC, dtype=torch.uint8, device=dev)
        eng.gen_records(gp, i * B, B, t)
        ring.append(t)
    eng.sync()
    for i in range(2):
        eng.ingest(ring[i % 4].data_ptr(), B)
    dt = timed(lambda i: eng.ingest(ring[i % 4].data_ptr(), B), steps)
    st = eng.stats()
    print(json.dumps({"bench": "KERNEL_MAP mode map update", "workload": f"{n_keys} Zipf-1.1 keys, batch {B}",
                      "Mpkts_s": B * steps / dt / 1e6, "flows": eng.live_flows(), "kernel_launches": st["kernel_launches"]}), flush=True)
    eng.close()


def small_cache_bench(max_entries=5000, n_keys=1_000_000, B=1 << 22, steps=3):
    """The reference's default CACHE_MAX_FLOWS (5000) against a stream with far more concurrent flows: every few
    thousand records the cache is 'full' (account.go:85-94) -> exact cut, eviction, resume.  Throughput of that loop
    with device-resident input and a device-resident eviction buffer (small caches)."""
    eng = fa.FlowAggEngine(max_entries, max_batch=B, cuda_stream=stream.cuda_stream)
    gp = fa.GenParams(seed=2, n_keys=n_keys, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    src = torch.empty(B * REC, dtype=torch.uint8, device=dev)
    eng.gen_records(gp, 0, B, src)
    out = torch.empty(max_entries * REC, dtype=torch.uint8, device=dev)
    eng.sync()
    gens = 0
    t0 = time.perf_counter()
    for _ in range(steps):
        done = 0
        while done < B:
            rc, took = eng.ingest(src.data_ptr() + done * REC, B - done)
            done += took
            if rc == fa.FA_FULL:
                eng.evict_into(out.data_ptr(), max_entries)
                gens += 1
    eng.evict_into(out.data_ptr(), max_entries)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    st = eng.stats()
    print(json.dumps({"bench": "ACCOUNTER with a small cache", "max_entries": max_entries, "workload": f"{n_keys} Zipf-1.1 keys",
                      "Mpkts_s": B * steps / dt / 1e6, "generations": gens, "records_per_generation": B * steps / max(gens, 1),
                      "kernel_launches": st["kernel_launches"]}), flush=True)
    eng.close()


def pb_bench(n_keys=4_000_000, B=1 << 23):
    """K8: evicted flows (device memory) -> pbflow.Record wire bytes (device memory): sizes + scan + write kernels."""
    import ctypes as C
    from netobserv_ebpf_agent_b200._lib import PbParams
    eng = fa.FlowAggEngine(1 << 23, flags=fa.FA_F_NO_FULL_CUT, max_batch=1 << 22, cuda_stream=stream.cuda_stream)
    gp = fa.GenParams(seed=6, n_keys=n_keys, dist=0, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0)
    src = torch.empty(B * REC, dtype=torch.uint8, device=dev)
    eng.gen_records(gp, 0, B, src)
    eng.ingest(src.data_ptr(), B)
    n = eng.live_flows()
    flows = torch.empty(n * REC, dtype=torch.uint8, device=dev)
    assert eng.evict_into(flows, n) == n
    p = PbParams(now_unix_ns=1_700_000_000_123_456_789, mono_now_ns=5_000_000_000_000, agent_ip_is_v4=1)
    for i, b in enumerate(bytes(10) + b"\xff\xff" + bytes([10, 1, 2, 3])):
        p.agent_ip[i] = b
    ln = C.c_size_t(0)
    L = fa.lib()
    out = torch.empty(n * 200, dtype=torch.uint8, device=dev)
    keys = torch.empty(n * 32, dtype=torch.uint8, device=dev)

    def run(_):
        rc = L.fa_pb_encode(eng._h, C.c_void_p(flows.data_ptr()), None, None, None, None, n, C.byref(p), C.c_void_p(out.data_ptr()),
                            out.numel(), None, C.c_void_p(keys.data_ptr()), C.byref(ln))
        assert rc == 0, rc
    run(0)
    dt = timed(run, 5) / 5
    print(json.dumps({"bench": "K8 pbflow encode (device in, device out)", "flows": n, "wire_bytes": ln.value, "bytes_per_flow": ln.value / n,
                      "Mflows_s": n / dt / 1e6, "GB_s_in_plus_out": (n * REC + ln.value + n * 32) / dt / 1e9}), flush=True)
    eng.close()


def snaps_bench(B=1 << 22, n_frames=1 << 16, steps=6):
    """(f4) fa_ingest_snaps from device memory: parse (count + parse kernels) + K1 fold, packets per second; and the parse
    kernels alone against their byte roofline (stride + 144 algorithmic bytes per submitted packet)."""
    from test_snaps import random_snaps
    for stride in (88, 104):
        rng = np.random.default_rng(stride)
        frames = random_snaps(rng, n_frames, stride, n_hosts=250)
        snaps = frames[rng.integers(0, n_frames, B)]
        snaps[:, 0:8] = (1_000_000 + 7 * np.arange(B, dtype=np.uint64)).view(np.uint8).reshape(B, 8)     # increasing timestamps
        d = torch.from_numpy(snaps.reshape(-1)).to(dev)
        eng = fa.FlowAggEngine(1 << 22, flags=fa.FA_F_NO_FULL_CUT, max_batch=B, cuda_stream=stream.cuda_stream)
        eng.ingest_snaps(d, stride)
        st0 = eng.stats()
        dt = timed(lambda i: eng.ingest_snaps(d, stride), steps) / steps
        st1 = eng.stats()
        sub = (st1["records_ingested"] - st0["records_ingested"]) // steps
        print(json.dumps({"bench": "f4 snapshots -> parse -> fold (device in)", "stride": stride, "snaps": B, "submitted": sub,
                          "flows": eng.live_flows(), "Mpkts_s": B / dt / 1e6,
                          "algorithmic_GB_s": (B * stride + sub * REC) / dt / 1e9}), flush=True)
        if stride == 104:                                    # the same stream behind a 12-rule flow filter
            from test_flow_filter import cidrs_of, random_rules
            eng.evict(cap=1 << 22)
            eng.set_flow_filter(random_rules(np.random.default_rng(5), 12), cidrs_of(([10, 1, 2, 0], 25)))
            eng.ingest_snaps(d, stride)
            st0 = eng.stats()
            dt = timed(lambda i: eng.ingest_snaps(d, stride), steps) / steps
            st1 = eng.stats()
            kept = (st1["records_ingested"] - st0["records_ingested"]) // steps
            print(json.dumps({"bench": "f4 snapshots -> parse -> 12-rule flow filter -> fold (device in)", "stride": stride, "snaps": B,
                              "kept": kept, "Mpkts_s": B / dt / 1e6}), flush=True)
        eng.close()


def dnscorr_bench(n_clients=200_000, B=1 << 20, steps=6):
    """K7: query / response pairs of n_clients x 4 ids, responses shuffled a few hundred packets behind their queries."""
    import oracle_lib as O
    rng = np.random.default_rng(9)
    eng = fa.FlowAggEngine(1 << 22, flags=fa.FA_F_ENABLE_DNS, max_batch=B, cuda_stream=stream.cuda_stream)
    batches = []
    for b in range(2):
        half = B // 2
        r = np.zeros(B, dtype=O.DNSREC_DTYPE)
        c = rng.integers(0, n_clients, half); i = rng.integers(1, 5, half)
        pos_q = np.sort(rng.choice(B, half, replace=False))                       # queries keep their order ...
        rest = np.setdiff1d(np.arange(B), pos_q)                                   # ... their responses fill the other positions
        ident = np.zeros((B, 40), dtype=np.uint8)
        def tuple_of(cl, reverse):
            t = np.zeros((len(cl), 40), dtype=np.uint8)
            a = np.zeros((len(cl), 16), dtype=np.uint8); a[:, 10:12] = 0xFF; a[:, 12] = 10; a[:, 13:16] = np.stack([(cl >> 16) & 255, (cl >> 8) & 255, cl & 255], 1)
            srv = np.zeros((len(cl), 16), dtype=np.uint8); srv[:, 10:12] = 0xFF; srv[:, 12:16] = [10, 255, 0, 53]
            port = (20000 + (cl % 40000)).astype("<u2").view(np.uint8).reshape(-1, 2)
            p53 = np.tile(np.frombuffer(np.uint16(53).tobytes(), dtype=np.uint8), (len(cl), 1))
            t[:, 0:16], t[:, 16:32] = (srv, a) if reverse else (a, srv)
            t[:, 32:34], t[:, 34:36] = (p53, port) if reverse else (port, p53)
            t[:, 36] = 17
            return t
        ident[pos_q] = tuple_of(c, False); ident[rest] = tuple_of(c, True)
        r["id"] = ident
        d = r["dns"]
        d["end"] = 1_000_000 + b * B + np.arange(B)
        d["id"][pos_q] = i; d["id"][rest] = i
        d["flags"][pos_q] = 0x0100; d["flags"][rest] = 0x8180
        d["eth"] = 0x0800
        r["dns"] = d
        batches.append(torch.from_numpy(O.as_bytes(r).copy()).to(dev))
    eng.ingest_dns_packets(batches[0])
    dt = timed(lambda k: eng.ingest_dns_packets(batches[k % 2]), steps)
    st = eng.stats()
    print(json.dumps({"bench": "K7 DNS query/response correlation + DNS fold", "Mpkts_s": B * steps / dt / 1e6,
                      "samples": st["dns_ingested"], "pending": st["dns_queries_pending"]}), flush=True)
    eng.close()


if __name__ == "__main__":
    which = sys.argv[1:] or ["sketch", "features", "kmap"]
    if "sketch" in which:
        sketch_bench()
    if "features" in which:
        feature_bench()
    if "kmap" in which:
        kmap_bench()
    if "smallcache" in which:
        small_cache_bench()
        small_cache_bench(max_entries=100_000)
    if "pb" in which:
        pb_bench()
    if "snaps" in which:
        snaps_bench()
    if "dnscorr" in which:
        dnscorr_bench()

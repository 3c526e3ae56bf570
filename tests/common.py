"""Shared helpers for the GPU parity tests: run the same stream through the CUDA path (via the C ABI)
and through the CPU oracle, and compare the evicted generations bit for bit."""
import numpy as np

import oracle_lib as O


def gen_host(seed, n, n_keys, dist=0, varying=0, first=0, t0=1_000_000, s_milli=1100):
    import netobserv_ebpf_agent_b200 as fa
    p = fa.GenParams(seed=seed, n_keys=n_keys, dist=dist, zipf_s_milli=s_milli, t0_ns=t0, varying_desc=varying)
    return fa.gen_records_host(p, first, n)


def oracle_generations(batches, max_entries):
    """-> list of sorted (n,144) arrays: every 'full' generation in order, then the final eviction."""
    acc = O.Accounter(max_entries)
    gens = []
    for b in batches:
        acc.account(b)
        while acc.pending():
            gens.append(O.sort_records(acc.pop_generation()))
    gens.append(O.sort_records(acc.evict()))
    acc.close()
    return gens


def gpu_generations(batches, max_entries, to_device=False, **engine_kw):
    import netobserv_ebpf_agent_b200 as fa
    gens = []
    with fa.FlowAggEngine(max_entries, **engine_kw) as eng:
        for b in batches:
            buf = np.ascontiguousarray(b).view(np.uint8).reshape(-1)
            if to_device:
                import torch
                t = torch.from_numpy(buf.copy()).cuda()
                n = buf.size // 144
                done = 0
                while done < n:
                    rc, took = eng.ingest(t.data_ptr() + done * 144, n - done)
                    done += took
                    if rc == fa.FA_FULL:
                        gens.append(O.sort_records(eng.evict()))
            else:
                eng.ingest_all(buf, lambda r: gens.append(O.sort_records(r)))
        gens.append(O.sort_records(eng.evict()))
        stats = eng.stats()
    return gens, stats


def assert_same_generations(got, want):
    assert len(got) == len(want), (len(got), len(want), [len(g) for g in got], [len(w) for w in want])
    for gi, (g, w) in enumerate(zip(got, want)):
        assert g.shape == w.shape, (gi, g.shape, w.shape)
        if g.size and not np.array_equal(g, w):
            bad = np.nonzero((g != w).any(axis=1))[0]
            i = int(bad[0])
            gr = g[i].view(O.REC_DTYPE)[0]
            wr = w[i].view(O.REC_DTYPE)[0]
            diff = [n for n in O.REC_DTYPE.names if not np.array_equal(gr[n], wr[n])]
            raise AssertionError(f"generation {gi}: {len(bad)} of {len(g)} flows differ; first at {i}: fields {diff}\n"
                                 f" gpu   : {gr}\n oracle: {wr}")

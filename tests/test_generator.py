"""The synthetic stream that both the CPU legs and the GPU bench consume: shape checks on the host
instance of the generator (the device instance is compared bit for bit in tests/test_gpu_parity.py)."""
import numpy as np

import oracle_lib as O
from common import gen_host

COLLAPSED = {0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40, 0x80, 0x100, 0x200, 0x400}   # bpf/utils.h:24-51


def test_records_look_like_ring_buffer_single_packet_flows():
    r = gen_host(seed=70, n=50_000, n_keys=5_000, dist=1, t0=123).view(O.REC_DTYPE).reshape(-1)
    assert (r["packets"] == 1).all() and (r["start"] == r["end"]).all()
    assert (np.diff(r["start"].astype(np.int64)) == 1).all() and r["start"][0] == 123      # strictly monotone
    assert r["bytes"].min() >= 64 and r["bytes"].max() <= 1500
    tcp = r["proto"] == 6
    assert set(np.unique(r["flags"][tcp]).tolist()) <= COLLAPSED and (r["flags"][~tcp] == 0).all()
    assert set(np.unique(r["proto"]).tolist()) == {6, 17} and 0.7 < tcp.mean() < 0.9
    v4 = (r["src_ip"][:, :10] == 0).all(axis=1) & (r["src_ip"][:, 10] == 0xFF) & (r["src_ip"][:, 11] == 0xFF)
    assert 0.85 < v4.mean() < 0.95
    assert (r["eth"][v4] == 0x0800).all() and (r["eth"][~v4] == 0x86DD).all()
    assert not r["pad_id"].any() and not r["pad0"].any() and not r["pad1"].any() and not r["lock"].any()


def test_descriptors_are_per_key_constants_by_default():
    raw = gen_host(seed=71, n=40_000, n_keys=300, dist=0)
    keys, inv = np.unique(raw[:, :40], axis=0, return_inverse=True)
    desc = np.concatenate([raw[:, 68:70], raw[:, 72:144]], axis=1)       # eth_protocol + descriptor
    for k in range(len(keys)):
        rows = desc[inv.reshape(-1) == k]
        assert (rows == rows[0]).all()
    assert len(keys) == 300                                              # distinct key ids -> distinct 5-tuples


def test_zipf_head_matches_the_law():
    n, n_keys, s = 400_000, 100_000, 1.1
    raw = gen_host(seed=72, n=n, n_keys=n_keys, dist=1)
    _, counts = np.unique(raw[:, :40], axis=0, return_counts=True)
    counts = np.sort(counts)[::-1]
    h = (np.arange(1, n_keys + 1) ** -s).sum()
    for rank in (1, 2, 3, 10):
        expect = n * rank ** -s / h
        assert abs(counts[rank - 1] - expect) < 5 * np.sqrt(expect) + 0.02 * expect, (rank, counts[rank - 1], expect)
    assert counts[:100].sum() / n == np.clip(counts[:100].sum() / n, 0.45, 0.60)     # top-100 mass of Zipf(1.1, 1e5) ~ 0.52


def test_slices_are_position_independent():
    a = gen_host(seed=73, n=10_000, n_keys=1_000, dist=1)
    b = np.concatenate([gen_host(seed=73, n=4_000, n_keys=1_000, dist=1, first=0),
                        gen_host(seed=73, n=6_000, n_keys=1_000, dist=1, first=4_000)])
    assert np.array_equal(a, b)


def test_oracle_twin_matches_the_product_generator():
    """oracle/gen.c (what the CPU arms and the in-bench parity check use) against the product's host generator."""
    for seed, n_keys, dist, var in ((2, 1_000_000, 1, 0), (2, 10_000_000, 0, 0), (9, 777, 1, 1), (5, 3, 0, 1)):
        g = O.Gen(seed, n_keys, dist=dist, varying_desc=var, t0_ns=55)
        a = g.records(12_345, 20_000, threads=3)
        b = gen_host(seed=seed, n=20_000, n_keys=n_keys, dist=dist, varying=var, first=12_345, t0=55)
        assert np.array_equal(a, b)
        g.close()


def test_persistent_sharded_accounter_equals_the_sequential_one():
    recs = gen_host(seed=75, n=60_000, n_keys=4_000, dist=1)
    seq = O.Accounter(1 << 20)
    seq.account(recs)
    sh = O.ShardedAccounter(5)
    sh.account(recs[:25_000]); sh.account(recs[25_000:])
    assert len(sh) == len(seq)
    assert np.array_equal(O.sort_records(sh.evict()), O.sort_records(seq.evict()))
    assert len(sh) == 0
    sh.close(); seq.close()

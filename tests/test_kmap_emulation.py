"""KERNEL_MAP kernels, logic check on the CPU: the per-thread bodies of csrc/kmap_body.cuh are compiled for the host
(tests/emul/kmap_emul.cpp) and run pass by pass with the thread indices of every pass in a seeded random order; the
result must equal the oracle's sequential map update (bpf/flows.c:76-143,222-288) bit for bit — flows, spilled
records and counters.  The device memory model and the launch plumbing are what tests/test_gpu_kernel_map.py adds."""
import ctypes
import os

import numpy as np
import pytest

import oracle_lib as O
from common import gen_host

from emul_build import EMUL, build, csrc

# a wedged emulation (it is thousands of OS threads) must not hang the suite: pytest-timeout, if installed
pytestmark = pytest.mark.timeout(900)

# two ways of running the same kernel bodies on the CPU:
#   shuffled  one index at a time, every pass in a seeded random order, plain memory operations (kmap_emul.cpp)
#   simt      csrc/kmap.cu's kernels on tests/emul/simt.h: one OS thread per CUDA thread, real atomics, 3 CTAs at once
#   ...-v2    the per-flow finalisation variant (km2_* bodies / kernels) on the same two runners
BACKENDS = {"shuffled": "kmap_emul", "simt": "kmap_simt", "shuffled-v2": "kmap_emul", "simt-v2": "kmap_simt"}
_libs = {}


def emul(backend="shuffled"):
    if backend in _libs:
        return _libs[backend]
    so = build(BACKENDS[backend], csrc("kmap_body.cuh", "kmap.cu", "common.cuh", "kernels.cuh") + [os.path.join(EMUL, "simt.h")])
    L = ctypes.CDLL(so)
    L.kmap_emul_new.restype = ctypes.c_void_p
    L.kmap_emul_new.argtypes = [ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64]
    L.kmap_emul_free.argtypes = [ctypes.c_void_p]
    L.kmap_emul_batch.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint32]
    L.kmap_emul_live.restype = ctypes.c_uint64
    L.kmap_emul_live.argtypes = [ctypes.c_void_p]
    L.kmap_emul_evict.restype = ctypes.c_uint64
    L.kmap_emul_evict.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    L.kmap_emul_spilled.restype = ctypes.c_uint64
    L.kmap_emul_spilled.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint64]
    L.kmap_emul_counters.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.kmap_emul_set_impl.argtypes = [ctypes.c_void_p, ctypes.c_int]
    _libs[backend] = L
    return L


# v1 on the thread-per-CUDA-thread runner is kept for manual runs (KMAP_BACKENDS=all); the suite runs v1 shuffled and v2 on both
_ACTIVE = sorted(BACKENDS) if os.environ.get("KMAP_BACKENDS") == "all" else ["shuffled", "shuffled-v2", "simt-v2"]


@pytest.fixture(params=_ACTIVE)
def backend(request):
    return request.param


def aligned_copy(recs):
    """16-byte aligned copy of a record buffer (the bodies read records through aligned words)."""
    b = O.as_bytes(recs)
    raw = np.zeros(b.size + 16, dtype=np.uint8)
    off = (-raw.ctypes.data) % 16
    a = raw[off: off + b.size]
    a[:] = b
    return a


class Emul:
    def __init__(self, max_entries, max_batch, ringbuf=True, spill_cap=1 << 16, seed=1, backend="shuffled"):
        self.L = emul(backend)
        self.h = self.L.kmap_emul_new(max_entries, max_batch, 1 if ringbuf else 0, spill_cap, seed)
        self.L.kmap_emul_set_impl(self.h, 2 if backend.endswith("-v2") else 1)
        self.max_batch = max_batch

    def packets(self, recs):
        a = aligned_copy(recs)
        n = a.size // O.REC
        for lo in range(0, n, self.max_batch):
            c = min(self.max_batch, n - lo)
            rc = self.L.kmap_emul_batch(self.h, a[lo * O.REC:].ctypes.data, c)
            assert rc == 0, f"emulated batch failed: {rc}"

    def evict(self):
        n = self.L.kmap_emul_live(self.h)
        out = np.zeros(max(n, 1) * O.REC, dtype=np.uint8)
        got = self.L.kmap_emul_evict(self.h, out.ctypes.data, n)
        assert got == n, f"table scan found {got} flows, live counter says {n}"
        return out[: n * O.REC].reshape(-1, O.REC)

    def spilled(self, cap=1 << 16):
        out = np.zeros(cap * O.REC, dtype=np.uint8)
        n = self.L.kmap_emul_spilled(self.h, out.ctypes.data, cap)
        return out[: n * O.REC].reshape(-1, O.REC)

    def counters(self):
        c = (ctypes.c_uint64 * 6)()
        self.L.kmap_emul_counters(self.h, c)
        return dict(intf_missed=c[0], fail_create=c[1], spill_cursor=c[2], spill_dropped=c[3], bset_count=c[4], table_full=c[5])

    def close(self):
        if self.h:
            self.L.kmap_emul_free(self.h)
            self.h = None

    __del__ = close


def messy_stream(seed, n, n_keys, n_ifaces=4, tls=True, zero_if=True):
    """Packet events with everything the map update looks at varying per packet: interface (incl. 0), direction,
    dscp, sampling, TCP flags, TLS fields, non-monotone timestamps."""
    rng = np.random.default_rng(seed)
    recs = gen_host(seed=seed, n=n, n_keys=n_keys, dist=1).copy()
    r = recs.view(O.REC_DTYPE).reshape(-1)
    r["if_index"] = rng.integers(0 if zero_if else 1, n_ifaces + 1, n)
    r["direction"] = rng.integers(0, 2, n)
    r["dscp"] = rng.integers(0, 4, n)
    r["sampling"] = rng.integers(0, 3, n)
    r["start"] = rng.integers(1, 1 << 40, n)                  # "last writer" must not be mistaken for max
    r["end"] = r["start"]
    r["flags"] = rng.choice(np.array([0, 1, 2, 4, 0x10, 0x100, 0x200], dtype=np.uint16), n)
    r["packets"] = 1
    if tls:
        r["ssl_version"] = rng.choice(np.array([0, 0, 0, 0x0303, 0x0304], dtype=np.uint16), n)
        r["tls_types"] = rng.choice(np.array([0, 0, 1, 2, 4], dtype=np.uint8), n)
        r["cipher"] = rng.choice(np.array([0, 0x1301, 0x1302], dtype=np.uint16), n)
        r["key_share"] = rng.choice(np.array([0, 0x001D, 0x0017], dtype=np.uint16), n)
    else:
        for f in ("ssl_version", "tls_types", "cipher", "key_share"):
            r[f] = 0
    r["misc"] = 0
    r["nb_obs"] = 0
    r["obs_dir"] = 0
    r["obs_intf"] = 0
    r["errno"] = 0
    return recs


def check(recs, max_entries, max_batch, ringbuf=True, seed=1, evict_every=None, backend="shuffled"):
    b = O.as_bytes(recs)
    n = b.size // O.REC
    if backend.startswith("simt"):              # a launch costs ~800 OS threads there: at most 8 batches per stream
        max_batch = max(max_batch, -(-n // 8))
    km = O.KernelMap(max_entries, ringbuf_fallback=ringbuf)
    em = Emul(max_entries, max_batch, ringbuf=ringbuf, seed=seed, backend=backend)
    step = evict_every or n
    for lo in range(0, n, step):
        part = b[lo * O.REC: (lo + step) * O.REC]
        km.packets(part)
        em.packets(part)
        want, got = O.sort_records(km.evict()), O.sort_records(em.evict())
        assert want.shape == got.shape
        if not np.array_equal(want, got):
            w, g = want.view(O.REC_DTYPE).reshape(-1), got.view(O.REC_DTYPE).reshape(-1)
            bad = np.nonzero((want != got).any(axis=1))[0][0]
            diff = [f for f in O.REC_DTYPE.names if not np.array_equal(w[bad][f], g[bad][f])]
            raise AssertionError(f"flow {bad}: fields {diff}: want {[w[bad][f] for f in diff]} got {[g[bad][f] for f in diff]}")
        n_sp = km.spilled()
        ws = O.sort_records(km.spilled_records(max(n_sp, 1))) if n_sp else np.zeros((0, O.REC), np.uint8)
        gs = em.spilled()
        assert len(gs) == n_sp
        if n_sp:
            # several spilled packets of one flow: compare as multisets of whole records
            assert np.array_equal(ws[np.lexsort(ws.T[::-1])], gs[np.lexsort(gs.T[::-1])])
    c = em.counters()
    assert c["intf_missed"] == km.intf_missed and c["fail_create"] == km.fail_create
    assert c["table_full"] == 0 and c["spill_dropped"] == 0


def test_generator_stream_single_batch_and_chunked(backend):
    recs = gen_host(seed=60, n=100_000, n_keys=5_000, dist=1)
    check(recs, 1 << 16, 100_000, backend=backend)
    check(recs, 1 << 16, 7_001, seed=2, backend=backend)


@pytest.mark.parametrize("seed", [11, 12, 13])
def test_messy_stream_matches_the_sequential_map_update(seed, backend):
    recs = messy_stream(seed, 60_000, 300)
    check(recs, 1 << 12, 60_000, seed=seed, backend=backend)            # one batch: everything order-dependent happens inside a launch
    check(recs, 1 << 12, 997, seed=seed + 100, backend=backend)         # many batches: state carried across launches


def test_many_interfaces_fill_the_observed_list(backend):
    recs = messy_stream(21, 40_000, 40, n_ifaces=12)
    check(recs, 1 << 10, 40_000, seed=3, backend=backend)
    check(recs, 1 << 10, 512, seed=4, backend=backend)


def test_single_flow_many_interfaces(backend):
    recs = messy_stream(22, 5_000, 1, n_ifaces=30)
    check(recs, 16, 5_000, seed=5, backend=backend)
    check(recs, 16, 64, seed=6, backend=backend)


def test_full_map_spills_to_the_ring_buffer_or_counts(backend):
    recs = messy_stream(31, 30_000, 2_000, tls=False)
    check(recs, 500, 30_000, ringbuf=True, seed=7, backend=backend)       # cut inside the first batch
    check(recs, 500, 4_096, ringbuf=True, seed=8, backend=backend)        # map already full when later batches start
    check(recs, 500, 4_096, ringbuf=False, seed=9, backend=backend)       # HASHMAP_FAIL_CREATE_FLOW instead


def test_eviction_between_batches(backend):
    recs = messy_stream(41, 50_000, 800)
    check(recs, 1 << 11, 2_048, seed=10, evict_every=10_000, backend=backend)


def test_committed_fixture(backend):
    """tests/golden/kmap/kmap_messy_seed11.npz: full map (300 entries), spills, observed-interface misses."""
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kmap", "kmap_messy_seed11.npz"))
    em = Emul(int(z["max_entries"]), 8_000, ringbuf=True, backend=backend)
    em.packets(z["records"])
    assert np.array_equal(O.sort_records(em.evict()), z["flows"])
    sp = em.spilled()
    assert np.array_equal(sp[np.lexsort(sp.T[::-1])], z["spilled"])
    c = em.counters()
    assert [c["intf_missed"], c["fail_create"]] == [int(x) for x in z["counters"]]

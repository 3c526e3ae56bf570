"""ctypes loader for the CPU oracle (oracle/liboracle.so).

Test infrastructure only: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs — never by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

REC, ID, MET, DNS, ADD, DNSREC, ADDREC = 144, 40, 104, 64, 32, 104, 72


def build(force=False):
    src = [os.path.join(ORACLE_DIR, f) for f in ("oracle.c", "gen.c", "oracle.h", "Makefile")]
    if (not force and os.path.exists(LIB_PATH)
            and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return LIB_PATH
    subprocess.run(["make", "-C", ORACLE_DIR, "-s"], check=True)
    return LIB_PATH


_u8p = C.POINTER(C.c_uint8)


def _p(a):
    return a.ctypes.data_as(_u8p) if a is not None else None


class Content(C.Structure):
    _fields_ = [("metrics", C.c_uint8 * MET), ("dns", C.c_uint8 * DNS), ("additional", C.c_uint8 * ADD),
                ("has_dns", C.c_uint8), ("has_additional", C.c_uint8)]


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, sz, u64, u32 = C.c_void_p, C.c_size_t, C.c_uint64, C.c_uint32
    sig = {
        "oracle_read_from": (None, [_u8p, _u8p]),
        "oracle_accumulate_base": (None, [_u8p, _u8p]),
        "oracle_accumulate_dns": (None, [C.POINTER(Content), _u8p]),
        "oracle_accumulate_additional": (None, [C.POINTER(Content), _u8p]),
        "oracle_new_record_times": (None, [u64, u64, u64, u64, C.POINTER(u64), C.POINTER(u64)]),
        "oracle_accounter_new": (vp, [sz]),
        "oracle_accounter_free": (None, [vp]),
        "oracle_accounter_account": (None, [vp, _u8p, sz]),
        "oracle_accounter_len": (sz, [vp]),
        "oracle_accounter_evict": (sz, [vp, _u8p, sz]),
        "oracle_accounter_pending": (sz, [vp]),
        "oracle_accounter_next_generation_len": (sz, [vp]),
        "oracle_accounter_pop_generation": (sz, [vp, _u8p, sz]),
        "oracle_accounter_sharded_run": (sz, [_u8p, sz, C.c_int, _u8p, sz]),
        "oracle_flowmap_new": (vp, []),
        "oracle_flowmap_free": (None, [vp]),
        "oracle_flowmap_account": (None, [vp, _u8p, sz]),
        "oracle_flowmap_packets_kmap": (C.c_uint64, [vp, _u8p, sz]),
        "oracle_flowmap_fold_dns": (None, [vp, _u8p, sz]),
        "oracle_flowmap_fold_additional": (None, [vp, _u8p, sz]),
        "oracle_flowmap_len": (sz, [vp]),
        "oracle_flowmap_fold_drops": (None, [vp, _u8p, sz]),
        "oracle_flowmap_evict_ex": (sz, [vp, _u8p, _u8p, _u8p, _u8p, C.POINTER(u64), _u8p, sz]),
        "oracle_accumulate_drops": (None, [_u8p, _u8p, _u8p, _u8p]),
        "oracle_flowmap_evict": (sz, [vp, _u8p, _u8p, _u8p, _u8p, sz]),
        "oracle_kmap_new": (vp, [sz, C.c_int]),
        "oracle_kmap_free": (None, [vp]),
        "oracle_kmap_packets": (None, [vp, _u8p, sz]),
        "oracle_kmap_len": (sz, [vp]),
        "oracle_kmap_evict": (sz, [vp, _u8p, sz]),
        "oracle_kmap_spilled": (sz, [vp, _u8p, sz]),
        "oracle_kmap_counter_fail_create": (u64, [vp]),
        "oracle_kmap_counter_intf_missed": (u64, [vp]),
        "oracle_sharded_new": (vp, [C.c_int]),
        "oracle_sharded_free": (None, [vp]),
        "oracle_sharded_account": (None, [vp, _u8p, sz]),
        "oracle_sharded_len": (sz, [vp]),
        "oracle_sharded_evict": (sz, [vp, _u8p, sz]),
        "oracle_gen_new": (vp, [u64, u64, u32, u32, u64, u32]),
        "oracle_gen_free": (None, [vp]),
        "oracle_gen_records": (None, [vp, u64, sz, _u8p, C.c_int]),
        "oracle_gen_key": (None, [vp, u64, _u8p]),
        "oracle_key_premix": (u64, [_u8p]),
        "oracle_slot_hash": (u64, [_u8p]),
        "oracle_owner_hash": (u64, [_u8p]),
        "oracle_parse_snap": (C.c_int, [_u8p, u32, _u8p]),
        "oracle_parse_snaps": (sz, [_u8p, sz, u32, _u8p, C.POINTER(u32)]),
        "oracle_filter_packet": (C.c_int, [_u8p, sz, _u8p, sz, _u8p, C.POINTER(u64)]),
        "oracle_parse_snaps_filtered": (sz, [_u8p, sz, u32, _u8p, sz, _u8p, sz, _u8p, C.POINTER(u32), C.POINTER(u64)]),
        "oracle_dnscorr_new": (vp, [sz]),
        "oracle_dnscorr_free": (None, [vp]),
        "oracle_dnscorr_packets": (sz, [vp, _u8p, sz, _u8p]),
        "oracle_dnscorr_pending": (sz, [vp]),
        "oracle_dnscorr_purge": (sz, [vp, u64, u64]),
        "oracle_cms_update": (None, [C.POINTER(u64), u32, u32, u64, _u8p, sz]),
        "oracle_cms_query": (None, [C.POINTER(u64), u32, u32, u64, _u8p, sz, C.POINTER(u64)]),
        "oracle_hll_update": (None, [_u8p, u32, u64, _u8p, sz]),
        "oracle_hll_estimate": (C.c_double, [_u8p, u32]),
    }
    for name, (res, args) in sig.items():
        f = getattr(L, name)
        f.restype, f.argtypes = res, args
    _lib = L
    return L


# ---------------------------------------------------------------- helpers

REC_DTYPE = np.dtype([
    ("src_ip", "u1", 16), ("dst_ip", "u1", 16), ("src_port", "<u2"), ("dst_port", "<u2"),
    ("proto", "u1"), ("icmp_type", "u1"), ("icmp_code", "u1"), ("pad_id", "u1"),
    ("start", "<u8"), ("end", "<u8"), ("bytes", "<u8"), ("packets", "<u4"),
    ("eth", "<u2"), ("flags", "<u2"), ("src_mac", "u1", 6), ("dst_mac", "u1", 6),
    ("if_index", "<u4"), ("lock", "<u4"), ("sampling", "<u4"),
    ("direction", "u1"), ("errno", "u1"), ("dscp", "u1"), ("nb_obs", "u1"),
    ("obs_dir", "u1", 6), ("pad0", "u1", 2), ("obs_intf", "<u4", 6),
    ("ssl_version", "<u2"), ("cipher", "<u2"), ("key_share", "<u2"),
    ("tls_types", "u1"), ("misc", "u1"), ("pad1", "u1", 4),
])
assert REC_DTYPE.itemsize == REC

DNS_DTYPE = np.dtype([("start", "<u8"), ("end", "<u8"), ("latency", "<u8"), ("id", "<u2"), ("flags", "<u2"),
                      ("eth", "<u2"), ("errno", "u1"), ("name", "u1", 32), ("pad", "u1")])
assert DNS_DTYPE.itemsize == DNS
ADD_DTYPE = np.dtype([("start", "<u8"), ("end", "<u8"), ("rtt", "<u8"), ("ipsec_ret", "<i4"),
                      ("eth", "<u2"), ("ipsec_enc", "u1"), ("pad", "u1")])
assert ADD_DTYPE.itemsize == ADD
DROP_DTYPE = np.dtype([("start", "<u8"), ("end", "<u8"), ("bytes", "<u2"), ("packets", "<u2"), ("cause", "<u4"), ("flags", "<u2"),
                       ("eth", "<u2"), ("state", "u1"), ("pad", "u1", 3)])
assert DROP_DTYPE.itemsize == 32
DROPREC_DTYPE = np.dtype([("id", "u1", 40), ("drop", DROP_DTYPE)])
DNSREC_DTYPE = np.dtype([("id", "u1", 40), ("dns", DNS_DTYPE)])
ADDREC_DTYPE = np.dtype([("id", "u1", 40), ("add", ADD_DTYPE)])
assert DNSREC_DTYPE.itemsize == DNSREC and ADDREC_DTYPE.itemsize == ADDREC


def as_bytes(a):
    """View any contiguous numpy array as a flat uint8 array."""
    return np.ascontiguousarray(a).view(np.uint8).reshape(-1)


def sort_records(raw, width=REC):
    """Sort an (n*width,) uint8 buffer of fixed-width records by their 40-byte key."""
    r = np.ascontiguousarray(raw).view(np.uint8).reshape(-1, width)
    if len(r) == 0:
        return r
    keys = r[:, :ID]
    order = np.lexsort(keys.T[::-1])
    return r[order]


def sort_perm(raw, width=REC):
    r = np.ascontiguousarray(raw).view(np.uint8).reshape(-1, width)
    if len(r) == 0:
        return np.zeros(0, dtype=np.int64)
    return np.lexsort(r[:, :ID].T[::-1])


class Gen:
    """CPU restatement of the synthetic stream (oracle/gen.c): the workload without the product library."""

    def __init__(self, seed, n_keys, dist=1, zipf_s_milli=1100, t0_ns=1_000_000, varying_desc=0):
        self.h = lib().oracle_gen_new(seed, n_keys, dist, zipf_s_milli, t0_ns, varying_desc)

    def records(self, first, n, threads=1, out=None):
        if out is None:
            out = np.empty(n * REC, dtype=np.uint8)
        lib().oracle_gen_records(self.h, first, n, _p(out), threads)
        return out.reshape(-1, REC)

    def key(self, rank):
        k = np.zeros(ID, dtype=np.uint8)
        lib().oracle_gen_key(self.h, rank, _p(k))
        return k

    def close(self):
        if self.h:
            lib().oracle_gen_free(self.h)
            self.h = None

    __del__ = close


class ShardedAccounter:
    """T private Accounters keyed by owner hash, persistent across batches (bench.py --impl reference)."""

    def __init__(self, threads):
        self.h = lib().oracle_sharded_new(threads)

    def account(self, recs):
        b = as_bytes(recs)
        lib().oracle_sharded_account(self.h, _p(b), b.size // REC)

    def __len__(self):
        return lib().oracle_sharded_len(self.h)

    def evict(self):
        n = len(self)
        out = np.zeros(max(n, 1) * REC, dtype=np.uint8)
        got = lib().oracle_sharded_evict(self.h, _p(out), n)
        return out[: got * REC].reshape(-1, REC)

    def close(self):
        if self.h:
            lib().oracle_sharded_free(self.h)
            self.h = None

    __del__ = close


class Accounter:
    """Sequential Accounter oracle (pkg/flow/account.go:58-124)."""

    def __init__(self, max_entries):
        self.h = lib().oracle_accounter_new(max_entries)

    def account(self, recs):
        b = as_bytes(recs)
        assert b.size % REC == 0
        lib().oracle_accounter_account(self.h, _p(b), b.size // REC)

    def __len__(self):
        return lib().oracle_accounter_len(self.h)

    def evict(self):
        n = len(self)
        out = np.zeros(max(n, 1) * REC, dtype=np.uint8)
        got = lib().oracle_accounter_evict(self.h, _p(out), n)
        return out[: got * REC].reshape(-1, REC)

    def pending(self):
        return lib().oracle_accounter_pending(self.h)

    def pop_generation(self):
        n = lib().oracle_accounter_next_generation_len(self.h)
        out = np.zeros(max(n, 1) * REC, dtype=np.uint8)
        got = lib().oracle_accounter_pop_generation(self.h, _p(out), n)
        return out[: got * REC].reshape(-1, REC)

    def close(self):
        if self.h:
            lib().oracle_accounter_free(self.h)
            self.h = None

    __del__ = close


class FlowMap:
    """LookupAndDeleteMap merged view with record-by-record feature folds."""

    def __init__(self):
        self.h = lib().oracle_flowmap_new()

    def account(self, recs):
        b = as_bytes(recs)
        lib().oracle_flowmap_account(self.h, _p(b), b.size // REC)

    def packets_kmap(self, recs):
        """Base = the kernel map's update (bpf/flows.c:222-288) instead of the Accounter's; returns intf_missed increments."""
        b = as_bytes(recs)
        return int(lib().oracle_flowmap_packets_kmap(self.h, _p(b), b.size // REC))

    def fold_dns(self, recs):
        b = as_bytes(recs)
        lib().oracle_flowmap_fold_dns(self.h, _p(b), b.size // DNSREC)

    def fold_additional(self, recs):
        b = as_bytes(recs)
        lib().oracle_flowmap_fold_additional(self.h, _p(b), b.size // ADDREC)

    def fold_drops(self, recs):
        b = as_bytes(recs)
        lib().oracle_flowmap_fold_drops(self.h, _p(b), b.size // 72)

    def evict_ex(self):
        """-> records, dns, additional, packet drops (n,32), rtt_min (n,), present."""
        n = len(self)
        m = max(n, 1)
        out, dns, add = np.zeros(m * REC, dtype=np.uint8), np.zeros(m * DNS, dtype=np.uint8), np.zeros(m * ADD, dtype=np.uint8)
        drops, rmin, pres = np.zeros(m * 32, dtype=np.uint8), np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.uint8)
        got = lib().oracle_flowmap_evict_ex(self.h, _p(out), _p(dns), _p(add), _p(drops), rmin.ctypes.data_as(C.POINTER(C.c_uint64)), _p(pres), n)
        return (out[: got * REC].reshape(-1, REC), dns[: got * DNS].reshape(-1, DNS), add[: got * ADD].reshape(-1, ADD),
                drops[: got * 32].reshape(-1, 32), rmin[:got], pres[:got])

    def __len__(self):
        return lib().oracle_flowmap_len(self.h)

    def evict(self):
        n = len(self)
        m = max(n, 1)
        out = np.zeros(m * REC, dtype=np.uint8)
        dns = np.zeros(m * DNS, dtype=np.uint8)
        add = np.zeros(m * ADD, dtype=np.uint8)
        pres = np.zeros(m, dtype=np.uint8)
        got = lib().oracle_flowmap_evict(self.h, _p(out), _p(dns), _p(add), _p(pres), n)
        return (out[: got * REC].reshape(-1, REC), dns[: got * DNS].reshape(-1, DNS),
                add[: got * ADD].reshape(-1, ADD), pres[:got])

    def close(self):
        if self.h:
            lib().oracle_flowmap_free(self.h)
            self.h = None

    __del__ = close


class KernelMap:
    """aggregated_flows hit/miss semantics (bpf/flows.c:76-143,222-288)."""

    def __init__(self, max_entries, ringbuf_fallback=True):
        self.h = lib().oracle_kmap_new(max_entries, 1 if ringbuf_fallback else 0)

    def packets(self, recs):
        b = as_bytes(recs)
        lib().oracle_kmap_packets(self.h, _p(b), b.size // REC)

    def __len__(self):
        return lib().oracle_kmap_len(self.h)

    def evict(self):
        n = len(self)
        out = np.zeros(max(n, 1) * REC, dtype=np.uint8)
        got = lib().oracle_kmap_evict(self.h, _p(out), n)
        return out[: got * REC].reshape(-1, REC)

    def spilled(self):
        n = lib().oracle_kmap_spilled(self.h, None, 0)
        return n

    def spilled_records(self, cap):
        out = np.zeros(max(cap, 1) * REC, dtype=np.uint8)
        n = lib().oracle_kmap_spilled(self.h, _p(out), cap)
        return out[: min(n, cap) * REC].reshape(-1, REC)

    @property
    def fail_create(self):
        return lib().oracle_kmap_counter_fail_create(self.h)

    @property
    def intf_missed(self):
        return lib().oracle_kmap_counter_intf_missed(self.h)

    def close(self):
        if self.h:
            lib().oracle_kmap_free(self.h)
            self.h = None

    __del__ = close


def accumulate_base(p, o):
    pb = as_bytes(p).copy()
    ob = as_bytes(o)
    lib().oracle_accumulate_base(_p(pb), _p(ob))
    return pb


def read_from(wire):
    w = as_bytes(wire)
    out = np.zeros(REC, dtype=np.uint8)
    lib().oracle_read_from(_p(w), _p(out))
    return out



def parse_snaps(snaps, stride):
    """(f4) oracle_parse_snaps: (n x stride) snapshot bytes -> (records (m,144), snapshot index of every record (m,))."""
    b = np.ascontiguousarray(snaps).view(np.uint8).reshape(-1)
    n = b.size // stride
    out = np.zeros((n, 144), dtype=np.uint8)
    src = np.zeros(n, dtype=np.uint32)
    m = lib().oracle_parse_snaps(_p(b), n, stride, _p(out), src.ctypes.data_as(C.POINTER(C.c_uint32)))
    return out[:m].copy(), src[:m].copy()


def parse_snaps_filtered(snaps, stride, rules, peers):
    """(f4) with the flow filter: -> (records kept (m,144), their snapshot indices, [accept, reject, nomatch] counters)."""
    b = np.ascontiguousarray(snaps).view(np.uint8).reshape(-1)
    n = b.size // stride
    out = np.zeros((n, 144), dtype=np.uint8)
    src = np.zeros(n, dtype=np.uint32)
    ctr = np.zeros(3, dtype=np.uint64)
    rb = np.ascontiguousarray(rules).view(np.uint8).reshape(-1)
    pb = np.ascontiguousarray(peers).view(np.uint8).reshape(-1) if peers is not None and len(peers) else np.zeros(20, dtype=np.uint8)
    m = lib().oracle_parse_snaps_filtered(_p(b), n, stride, _p(rb), len(rules), _p(pb), 0 if peers is None else len(peers), _p(out),
                                          src.ctypes.data_as(C.POINTER(C.c_uint32)), ctr.ctypes.data_as(C.POINTER(C.c_uint64)))
    return out[:m].copy(), src[:m].copy(), ctr


class DnsCorrelator:
    """dns_flows + the sample flow_monitor emits per DNS packet (bpf/dns_tracker.h:68-127, bpf/flows.c:291-330), sequential."""

    def __init__(self, max_entries=1 << 20):
        self.h = lib().oracle_dnscorr_new(max_entries)

    def packets(self, pkts):
        a = np.ascontiguousarray(as_bytes(pkts))
        n = a.size // DNSREC
        out = np.zeros(max(n, 1) * DNSREC, dtype=np.uint8)
        k = lib().oracle_dnscorr_packets(self.h, _p(a), n, _p(out))
        return out[: k * DNSREC].reshape(-1, DNSREC)

    def pending(self):
        return lib().oracle_dnscorr_pending(self.h)

    def purge(self, now, timeout):
        return lib().oracle_dnscorr_purge(self.h, now, timeout)

    def __del__(self):
        if getattr(self, "h", None):
            lib().oracle_dnscorr_free(self.h)
            self.h = None

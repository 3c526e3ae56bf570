"""Host-side mirrors of the reference's pipeline stages for this path, on top of the C ABI.

Go names are kept so the parity tests read like the reference's own tests:
  Accounter      — pkg/flow/account.go:19-124   (NewAccounter / Account / evict)
  MapTracer      — pkg/flow/tracer_map.go:24-146 (evictFlows via LookupAndDeleteMap)
  new_record_times — pkg/model/record.go:90-97
The Go toolchain is absent from the build image; the cgo binding a maintainer would add is
shown in INTEGRATION.md and the compiled-language mirror lives in host/accounter.hpp.
"""
import numpy as np

from ._lib import FA_FULL, REC_BYTES
from .engine import FlowAggEngine

MASK64 = (1 << 64) - 1


def new_record_times(now_unix_ns, mono_now_ns, start_mono, end_mono):
    """TimeFlowStart/End = now - (monoNow - mono), u64 arithmetic (pkg/model/record.go:90-97)."""
    tfs = (now_unix_ns - ((mono_now_ns - start_mono) & MASK64)) & MASK64
    tfe = (now_unix_ns - ((mono_now_ns - end_mono) & MASK64)) & MASK64
    return tfs, tfe


class Accounter:
    """GPU-backed replacement of flow.Accounter.

    account(records) plays the `case record := <-in` arm for a batch of RawRecords,
    tick() the `case <-evictTick.C` arm, close() the closed-channel arm.  Evicted batches
    are appended to `self.out` as (reason, ndarray[n,144], now, mono_now) — the `evictor <- records`
    send of account.go:123.
    """

    def __init__(self, max_entries, clock=None, mono_clock=None, engine=None, **engine_kw):
        self.max_entries = max_entries
        self.clock = clock or (lambda: 0)
        self.mono_clock = mono_clock or (lambda: 0)
        self.engine = engine or FlowAggEngine(max_entries, **engine_kw)
        self.out = []
        self.evictions = {"full": 0, "timeout": 0, "closing": 0}

    def _evict(self, reason, even_if_empty=False):
        recs = self.engine.evict()
        if len(recs) == 0 and not even_if_empty:
            return
        self.evictions[reason] += 1
        self.out.append((reason, recs, self.clock(), self.mono_clock()))

    def account(self, records):
        buf = np.ascontiguousarray(records).view(np.uint8).reshape(-1)
        n = buf.size // REC_BYTES
        done = 0
        while done < n:
            rc, took = self.engine.ingest(buf[done * REC_BYTES:])
            done += took
            if rc == FA_FULL:
                self._evict("full", even_if_empty=True)       # account.go:85-94

    def tick(self):
        if self.engine.live_flows() == 0:                     # account.go:64-66
            return
        self._evict("timeout")

    def close(self):
        self._evict("closing", even_if_empty=True)            # account.go:73-80


class MapTracer:
    """GPU-backed replacement of flow.MapTracer's eviction: LookupAndDeleteMap + NewRecord times."""

    def __init__(self, engine, clock, mono_clock):
        self.engine, self.clock, self.mono_clock = engine, clock, mono_clock

    def evict_flows(self):
        mono_now, now = self.mono_clock(), self.clock()       # tracer_map.go:105-106
        recs = self.engine.evict()
        r = recs.reshape(-1, REC_BYTES)
        start = r[:, 40:48].copy().view("<u8").reshape(-1)
        end = r[:, 48:56].copy().view("<u8").reshape(-1)
        with np.errstate(over="ignore"):
            tfs = np.uint64(now) - (np.uint64(mono_now) - start)
            tfe = np.uint64(now) - (np.uint64(mono_now) - end)
        return recs, tfs, tfe

#!/usr/bin/env python
"""Regenerates tests/golden/kmap/*.npz: a packet-event stream in which everything the map update of flow_monitor looks
at varies per packet (tests/test_kmap_emulation.py::messy_stream), and what the CPU oracle's KERNEL_MAP restatement
(bpf/flows.c:76-143,222-288) makes of it with a map of 300 entries and the ring-buffer fallback on: the evicted
flows, the spilled single-packet records, the counters.  Like the other fixtures these are oracle outputs (the
reference needs a kernel to run); they freeze the stream and the semantics against silent drift."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, "..", "..", ".."))

import oracle_lib as O  # noqa: E402
from test_kmap_emulation import messy_stream  # noqa: E402


def run(recs, max_entries):
    km = O.KernelMap(max_entries, ringbuf_fallback=True)
    km.packets(recs)
    flows = O.sort_records(km.evict())
    n_sp = km.spilled()
    sp = km.spilled_records(max(n_sp, 1))[:n_sp]
    sp = sp[np.lexsort(sp.T[::-1])] if n_sp else sp
    return flows, sp, np.array([km.intf_missed, km.fail_create], dtype=np.uint64)


def main():
    recs = messy_stream(11, 8_000, 400, n_ifaces=9)
    flows, spilled, counters = run(recs, 300)
    np.savez_compressed(os.path.join(HERE, "kmap_messy_seed11.npz"), records=recs, max_entries=300, flows=flows,
                        spilled=spilled, counters=counters)
    print(recs.shape, flows.shape, spilled.shape, counters)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Regenerates tests/golden/*.npz.

The reference is Go + eBPF C and cannot run in this image (no go, no clang -target bpf), so these fixtures are
NOT outputs of the reference: they are (a) synthetic input streams from this repo's counter-based generator and
(b) the evicted flows the CPU oracle produces for them — the oracle itself being pinned to the reference's own
known-answer tests in tests/test_oracle_goldens.py.  They freeze both, so that a later change to the generator,
the hash-independent oracle semantics or the CUDA path shows up as a diff.

  config1_seed1.npz   BASELINE.json configs[0]: 10k records, 100 5-tuples, maxEntries 5000 (one eviction of 100 flows)
  fullcut_seed10.npz  3000 records, 50 5-tuples, maxEntries 10: every 'full' generation in order + the final one
  refold_seed7.npz    5000 records, 7 5-tuples, per-record random descriptors (order-dependent merge rules)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

import oracle_lib as O  # noqa: E402
from common import gen_host, oracle_generations  # noqa: E402


def main():
    cases = {
        "config1_seed1": (dict(seed=1, n=10_000, n_keys=100), 5000),
        "fullcut_seed10": (dict(seed=10, n=3_000, n_keys=50), 10),
        "refold_seed7": (dict(seed=7, n=5_000, n_keys=7, dist=1, varying=1), 1 << 16),
    }
    for name, (g, max_entries) in cases.items():
        recs = gen_host(**g)
        gens = oracle_generations([recs], max_entries)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), records=recs, max_entries=max_entries,
                            n_generations=len(gens), **{f"gen{i}": x for i, x in enumerate(gens)})
        print(name, recs.shape, [len(x) for x in gens][:8], "...")


if __name__ == "__main__":
    main()

"""Committed fixtures (tests/golden/*.npz, made by tests/golden/make_golden.py): the generator and the oracle must
reproduce them on CPU; the CUDA path must reproduce them on the GPU."""
import glob
import os

import numpy as np
import pytest

from common import assert_same_generations, gen_host, gpu_generations, oracle_generations

GOLD = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "*.npz")))
GEN = {"config1_seed1": dict(seed=1, n=10_000, n_keys=100), "fullcut_seed10": dict(seed=10, n=3_000, n_keys=50),
       "refold_seed7": dict(seed=7, n=5_000, n_keys=7, dist=1, varying=1)}


def _load(path):
    z = np.load(path)
    return z["records"], int(z["max_entries"]), [z[f"gen{i}"] for i in range(int(z["n_generations"]))]


@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_generator_and_oracle_reproduce_fixture(path):
    recs, max_entries, gens = _load(path)
    assert np.array_equal(gen_host(**GEN[os.path.basename(path)[:-4]]), recs)     # host generator is frozen
    assert_same_generations(oracle_generations([recs], max_entries), gens)


@pytest.mark.gpu
@pytest.mark.parametrize("path", GOLD, ids=[os.path.basename(p)[:-4] for p in GOLD])
def test_cuda_path_reproduces_fixture(path):
    recs, max_entries, gens = _load(path)
    got, _ = gpu_generations([recs], max_entries, max_batch=4_000)
    assert_same_generations(got, gens)

"""Count-min + HyperLogLog kernels (new capability; the reference has none: parity unpinned).
Checked three ways: bit-exact against the CPU restatement of this repo's spec, against exact
counts within the stated (eps, delta) / standard-error bounds, and never-underestimate."""
import math

import numpy as np
import pytest

import oracle_lib as O
from common import gen_host

pytestmark = pytest.mark.gpu


def _oracle_sketch(recs, lw, depth, p, seed):
    import ctypes as C
    b = O.as_bytes(recs)
    cms = np.zeros(depth << lw, dtype=np.uint64)
    hll = np.zeros(1 << p, dtype=np.uint8)
    O.lib().oracle_cms_update(cms.ctypes.data_as(C.POINTER(C.c_uint64)), lw, depth, seed, O._p(b), b.size // 144)
    O.lib().oracle_hll_update(O._p(hll), p, seed, O._p(b), b.size // 144)
    return cms.reshape(depth, 1 << lw), hll


@pytest.mark.parametrize("dist,n_keys,n", [(1, 2_000, 100_000), (0, 50_000, 300_000), (1, 200_000, 400_000)])
def test_sketch_matches_cpu_restatement_bit_exact(dist, n_keys, n):
    import netobserv_ebpf_agent_b200 as fa
    lw, depth, p, seed = 12, 4, 10, 0xC0FFEE
    recs = gen_host(seed=21, n=n, n_keys=n_keys, dist=dist)
    with fa.FlowAggEngine(1 << 20, flags=fa.FA_F_ENABLE_SKETCH, cms_log2_width=lw, cms_depth=depth, hll_precision=p,
                          sketch_seed=seed, max_batch=70_000) as eng:
        eng.ingest(recs)
        cms, hll = eng.sketch_export(lw, depth, p)
        want_cms, want_hll = _oracle_sketch(recs, lw, depth, p, seed)
        assert np.array_equal(cms, want_cms)
        assert np.array_equal(hll, want_hll)
        # point queries through the kernel == min over rows of the exported table
        keys = np.unique(recs[:, :40], axis=0)[:5000]
        est = eng.cms_query(keys)
        import ctypes as C
        want = np.zeros(len(keys), dtype=np.uint64)
        O.lib().oracle_cms_query(want_cms.ctypes.data_as(C.POINTER(C.c_uint64)), lw, depth, seed,
                                 O._p(O.as_bytes(keys)), len(keys), want.ctypes.data_as(C.POINTER(C.c_uint64)))
        assert np.array_equal(est, want)
        assert abs(eng.hll_estimate() - O.lib().oracle_hll_estimate(O._p(want_hll), p)) < 1e-6
        # the flow table is unaffected by the fused sketch
        assert eng.live_flows() == len(np.unique(recs[:, :40], axis=0))


def test_sketch_error_bounds_against_exact_counts():
    """CMS: never under, over-estimate <= eps*N with eps = e/w for >= 1 - e^-d of the keys.
    HLL: relative error within 3 sigma = 3 * 1.04/sqrt(m)."""
    import netobserv_ebpf_agent_b200 as fa
    lw, depth, p = 16, 4, 14
    n, n_keys = 2_000_000, 300_000
    recs = gen_host(seed=22, n=n, n_keys=n_keys, dist=1)
    with fa.FlowAggEngine(1 << 20, flags=fa.FA_F_ENABLE_SKETCH, cms_log2_width=lw, cms_depth=depth, hll_precision=p,
                          sketch_seed=7) as eng:
        eng.ingest(recs)
        flows = eng.evict()                                   # exact per-flow packet counts from the table itself
        f = np.ascontiguousarray(flows).view(O.REC_DTYPE).reshape(-1)
        exact = f["packets"].astype(np.int64)
        est = eng.cms_query(flows[:, :40]).astype(np.int64)
        assert (est >= exact).all()                           # count-min never under-estimates
        eps_n = math.e / (1 << lw) * n
        frac_within = float(((est - exact) <= eps_n).mean())
        assert frac_within >= 1 - math.exp(-depth) - 0.005, frac_within
        distinct = len(f)
        rel = abs(eng.hll_estimate() - distinct) / distinct
        assert rel <= 3 * 1.04 / math.sqrt(1 << p), rel


def test_sketch_reset_and_disabled():
    import netobserv_ebpf_agent_b200 as fa
    recs = gen_host(seed=23, n=10_000, n_keys=100)
    with fa.FlowAggEngine(1000, flags=fa.FA_F_ENABLE_SKETCH, cms_log2_width=8, cms_depth=2, hll_precision=6) as eng:
        eng.ingest(recs)
        cms, hll = eng.sketch_export(8, 2, 6)
        assert int(cms.sum()) == 2 * 10_000 and hll.any()
        eng.sketch_reset()
        cms, hll = eng.sketch_export(8, 2, 6)
        assert not cms.any() and not hll.any()
    with fa.FlowAggEngine(1000) as eng:
        with pytest.raises(fa.FlowAggError):
            eng.cms_query(recs[:4, :40])

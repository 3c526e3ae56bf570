"""K7 (fa_ingest_dns_packets) on the device, against the sequential restatement — the same cases as tests/test_dns_correlate.py
at larger sizes, through libflowagg.so.

The kernels of csrc/dnscorr.cu were written after the round's GPU budget had been spent on K1; their first (and so far only)
hardware run used the round's last 48 GPU seconds: 4 passed (profiles/r2_k7_gpu_first_run.log).  The file sorts last and carries
a thread-method timeout so that a wedged kernel here could not hold or poison the rest of the GPU suite."""
import numpy as np
import pytest

import oracle_lib as O
from test_dns_correlate import dns_stream, query, response, run_case

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(240, method="thread")]


def test_hand_cases_on_the_device():
    cases = [query(100, 5), query(160, 5), response(400, 5), response(420, 5), query(10, 0), response(30, 0), response(77, 9)]
    run_case([np.concatenate(cases)])
    run_case(cases)


def test_stream_matches_the_restatement_on_the_device():
    s = dns_stream(13, 60_000, n_clients=3_000, n_ids=4)
    st = run_case([s[:25_000], s[25_000:26_000], s[26_000:]], flows=1 << 16)
    assert st["dns_packets_ingested"] == 60_000


def test_hot_keys_go_through_the_tail_on_the_device():
    s = dns_stream(14, 6_000, n_clients=3, n_ids=1, dup=0.6)
    run_case([s], flows=1 << 12)


def test_small_map_rebuild_and_purge_on_the_device(monkeypatch):
    monkeypatch.setenv("FA_DNS_MAX_ENTRIES", "1024")
    s = dns_stream(15, 30_000, n_clients=150, n_ids=3, orphan=0.05)
    st = run_case([s[:14_000], s[14_000:]], max_entries=1024, flows=1 << 14, purge=(0, 400_000, 50_000))
    assert st["dns_map_full"] == 0

"""The product's end queue (cactus_b200/csrc/end_queue.h) and MSA-level host logic (bar_windows.h) on the CPU: application
threads submit flowers (tickets) and collect them, lane workers merge whatever is pending into device batches -- here a
stand-in device that computes every job with the host build of the product's own graph code (tests/hosttest). The results
must equal the oracle's make_consistent_partial_order_alignments / msa_make_partial_order_alignment flower by flower,
whatever the batch composition, and a failing ticket must not take its batch mates down."""
import numpy as np

import _reflib as R
from _synth import family, to_ascii, two_end_problem


def _flowers(rng, n, L=60, kmax=5, window=10000):
    out = []
    for _ in range(n):
        K = int(rng.integers(2, kmax + 1))
        out.append(two_end_problem(rng, K, int(rng.integers(max(8, L // 2), L + 1))))
    return out


def _check(flowers, res, p, **kw):
    for (ends, ri, rr, ov), got in zip(flowers, res):
        want = R.oracle_make_consistent_partial_order_alignments(ends, ri, rr, ov, p=p, **kw)
        assert got is not None and len(got) == len(want)
        for a, b in zip(got, want):
            assert a.shape == b.shape and np.array_equal(a, b)


def test_queue_merges_tickets_and_matches_the_oracle(oracle_built):
    rng = np.random.default_rng(11)
    p = R.cactus_params(wb=10, wf=0.01)
    flowers = _flowers(rng, 14)
    res, rcs, batches = R.hosttest_flowers(flowers, p=p, n_lanes=2, n_threads=3)
    assert not rcs.any()
    assert batches < len(flowers)               # tickets shared device batches
    _check(flowers, res, p)


def test_a_burst_of_submissions_shares_one_batch(oracle_built):
    """one thread submits 40 flowers back to back before its first wait (the shim's bar() does this with thousands): the lane worker
    lingers while tickets keep arriving instead of launching a batch of the first few"""
    rng = np.random.default_rng(15)
    p = R.cactus_params(wb=10, wf=0.01)
    flowers = _flowers(rng, 40, L=30, kmax=3)
    res, rcs, batches = R.hosttest_flowers(flowers, p=p, n_lanes=2, n_threads=1)
    assert not rcs.any()
    assert batches <= 3, batches
    _check(flowers, res, p)


def test_small_batch_limit_and_many_windows(oracle_built):
    """a tiny job limit forces many batches; a tiny window forces several rounds per end (tickets re-enter the queue)"""
    rng = np.random.default_rng(12)
    p = R.cactus_params(wb=10, wf=0.01)
    flowers = _flowers(rng, 6, L=70)
    res, rcs, batches = R.hosttest_flowers(flowers, p=p, n_lanes=3, n_threads=2, max_jobs=3, window_size=25)
    assert not rcs.any() and batches > len(flowers)
    _check(flowers, res, p, window_size=25)


def test_independent_ends_without_consistency(oracle_built):
    rng = np.random.default_rng(13)
    p = R.cactus_params(wb=10, wf=0.01)
    ends = [[to_ascii(s) for s in family(rng, int(rng.integers(1, 6)), int(rng.integers(5, 50)))] for _ in range(9)]
    flowers = [(ends[i:i + 3], None, None, None) for i in range(0, 9, 3)]
    res, rcs, _ = R.hosttest_flowers(flowers, p=p, consistent=False, window_size=20)
    assert not rcs.any()
    for (es, _, _, _), got in zip(flowers, res):
        for e, a in zip(es, got):
            b = R.oracle_msa_make_partial_order_alignment(e, window_size=20, p=p)
            assert a.shape == b.shape and np.array_equal(a, b)


def test_failed_batches_are_rerun_per_ticket(oracle_built):
    """every 2nd device batch "fails": its tickets are re-run one by one and still deliver; a ticket with a bad input
    (a '-' in a string: code 5, which the device rejects) fails ALONE"""
    rng = np.random.default_rng(14)
    p = R.cactus_params(wb=10, wf=0.01)
    flowers = _flowers(rng, 10)
    bad = 4
    ends, ri, rr, ov = flowers[bad]
    ends = [list(e) for e in ends]
    ends[0][0] = ends[0][0][:3] + b"-" + ends[0][0][4:]
    flowers[bad] = (ends, ri, rr, ov)
    res, rcs, _ = R.hosttest_flowers(flowers, p=p, n_lanes=2, n_threads=2, fail_every=2)
    assert rcs[bad] != 0 and res[bad] is None
    ok = [i for i in range(len(flowers)) if i != bad]
    assert not rcs[ok].any()
    _check([flowers[i] for i in ok], [res[i] for i in ok], p)

"""Concurrent callers on one context (-m gpu). The reference enters the BAR code from OpenMP teams (bar/impl/bar.c:90-94), so
the C ABI is called from several host threads at once; their requests meet in the engine's queues (POA: end_queue.h, pair-HMM: group_commit.h) and share device batches.
Whatever the interleaving, every caller must get exactly what it gets when it is alone. (Named to run after the parity suites.)"""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

from _synth import family, pecan_pair, to_ascii

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(300)]


@pytest.fixture(scope="module")
def eng():
    import cactus_b200 as cb
    e = cb.Engine()
    yield e
    e.close()


def test_pecan_batches_from_many_threads(eng):
    rng = np.random.default_rng(2024)
    requests = []
    for r in range(24):
        pairs = []
        for _ in range(int(rng.integers(1, 5))):
            sx, sy, a = pecan_pair(rng, int(rng.choice([40, 200, 600])), k_anchor=12)
            pairs.append((sx, sy, a, bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
        requests.append(pairs)
    alone = [eng.get_aligned_pairs_using_anchors_batch(q, None, True) for q in requests]
    for rounds in range(2):
        with ThreadPoolExecutor(8) as ex:
            together = list(ex.map(lambda q: eng.get_aligned_pairs_using_anchors_batch(q, None, True), requests))
        for a, b in zip(alone, together):
            assert len(a) == len(b)
            for (t1, p1, c1), (t2, p2, c2) in zip(a, b):
                assert np.array_equal(t1, t2) and np.array_equal(p1, p2) and c1 == c2


def test_poa_ends_from_many_threads(eng):
    rng = np.random.default_rng(77)
    ends = [[to_ascii(s) for s in family(rng, int(rng.integers(2, 7)), int(rng.choice([30, 150, 400])), sub=0.05, ins=0.02, dele=0.02)]
            for _ in range(24)]
    alone = [eng.msa_make_partial_order_alignment(e, window_size=120) for e in ends]
    with ThreadPoolExecutor(8) as ex:
        together = list(ex.map(lambda e: eng.msa_make_partial_order_alignment(e, window_size=120), ends))
    for a, b in zip(alone, together):
        assert a.msa_seq.shape == b.msa_seq.shape and np.array_equal(a.msa_seq, b.msa_seq) and a.seq_lens == b.seq_lens

"""cPecan mode on the GPU (-m gpu): the CUDA pair-HMM kernel through the C ABI (barb200_pecan_aligned_pairs_batch and the
staged form) against (1) the committed golden vectors of the unmodified reference, (2) the plain-C oracle on seeded
random inputs -- bit-exact on the (score, x, y) triples and on the pre-floor posteriors --, (3) size-independent
properties at the benchmark shape. Nothing here reads /root/reference."""
import numpy as np
import pytest

import _golden as G
import _reflib as R
from _synth import pecan_pair

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import cactus_b200 as cb
    e = cb.Engine()
    yield e
    e.close()


def _cb_params(threshold, min_diags, tb_diags, expansion, split):
    import cactus_b200 as cb
    p = cb.PairwiseAlignmentParameters(threshold, min_diags, tb_diags, expansion, 1)
    p.c.split_matrix_bigger_than_this = split
    return p


def test_golden(eng):
    cases = list(G.pecan_cases())
    for c in cases:        # one call per case: parameters differ
        p = _cb_params(c["threshold"], c["min_diags"], c["tb_diags"], c["expansion"], c["split"])
        t, po, cells = eng.get_aligned_pairs_using_anchors_batch([(c["sx"], c["sy"], c["anchors"], c["rl"], c["rr"])], p, True)[0]
        assert np.array_equal(t, c["triples"]), c["id"]
        if "post" in c:
            assert np.array_equal(po[::-1], c["post"]), c["id"]
    dflt = [c for c in cases if (c["threshold"], c["min_diags"], c["tb_diags"], c["expansion"], c["split"]) == (0.01, 1000, 40, 20, 9000000)]
    assert len(dflt) >= 5
    res = eng.get_aligned_pairs_using_anchors_batch([(c["sx"], c["sy"], c["anchors"], c["rl"], c["rr"]) for c in dflt])
    for c, (t, cells) in zip(dflt, res):       # batched: same answers as one by one
        assert np.array_equal(t, c["triples"]), c["id"]


def test_reference_known_answer(eng):
    """submodules/cPecan/tests/pairwiseAlignerTest.c:243-322"""
    t = eng.get_aligned_pairs_using_anchors(b"AGCG", b"AGTTCG", [], _cb_params(0.2, 1000, 40, 2, 9000000))
    assert {(int(x), int(y)) for _, x, y in t} == {(0, 0), (1, 1), (2, 4), (3, 5)}


def test_random_vs_oracle(eng, oracle_built):
    rng = np.random.default_rng(31337)
    for rnd in range(6):
        thr = float(rng.choice([0.01, 0.2, 0.0001, 0.0]))
        md, tb, ex = int(rng.choice([1000, 100, 50])), int(rng.choice([40, 10, 1])), int(rng.choice([20, 4, 10, 0]))
        sb = int(rng.choice([30 * 30, 100 * 100, 400 * 400, 3000 * 3000]))
        pairs = []
        for it in range(24):
            L = int(rng.choice([1, 7, 30, 100, 300, 600, 1500]))
            sx, sy, a = pecan_pair(rng, L, k_anchor=int(rng.choice([8, 12, 20])), keep=float(rng.choice([1, 0.7, 0.3, 0.0])),
                                   sub=float(rng.choice([0.02, 0.1, 0.3])), ins=float(rng.choice([0.0, 0.01, 0.05])),
                                   dele=float(rng.choice([0.0, 0.01, 0.05])), nfrac=float(rng.choice([0, 0, 0.03])))
            pairs.append((sx, sy, a, bool(rng.integers(0, 2)), bool(rng.integers(0, 2))))
        pairs += [(b"", b"", [], False, False), (b"A", b"", [], False, True), (b"", b"ACGT", [], True, False)]
        res = eng.get_aligned_pairs_using_anchors_batch(pairs, _cb_params(thr, md, tb, ex, sb), True)
        po_ = R.pecan_params(thr, md, tb, ex)
        for i, (q, (t, po, cells)) in enumerate(zip(pairs, res)):
            to, poo = R.oracle_pecan_aligned_pairs(q[0], q[1], q[2], q[3], q[4], po_, sb)
            assert np.array_equal(t, to) and np.array_equal(po, poo), (rnd, i, len(q[0]), len(q[1]), len(q[2]))


def test_wide_unanchored_and_overflow(eng, oracle_built):
    """a wide band (no anchors, 700 x 700: diagonals far wider than a warp) and a job whose candidates overflow the
    optimistic output room (threshold 0 keeps every cell) -> the re-run path"""
    rng = np.random.default_rng(5)
    sx, sy, _ = pecan_pair(rng, 700, k_anchor=9999, sub=0.1, ins=0.02, dele=0.02)
    for thr in (0.01, 0.0):
        t, po, cells = eng.get_aligned_pairs_using_anchors_batch([(sx, sy, [], False, False)], _cb_params(thr, 1000, 40, 20, 9000000), True)[0]
        to, poo = R.oracle_pecan_aligned_pairs(sx, sy, [], False, False, R.pecan_params(thr), 9000000)
        assert np.array_equal(t, to) and np.array_equal(po, poo), thr
        assert cells == (len(sx) + 1) * (len(sy) + 1)


def test_bench_shape_properties(eng, oracle_built):
    """2 kbp pairs with MUM-like anchors (the benchmark's pecan workload): batch-composition independence, staged ==
    batch, symmetric posterior mass, and a sample against the oracle"""
    rng = np.random.default_rng(11)
    pairs = []
    for it in range(96):
        sx, sy, a = pecan_pair(rng, 2000, k_anchor=50)
        pairs.append((sx, sy, a, False, False))
    res = eng.get_aligned_pairs_using_anchors_batch(pairs, None, True)
    st = eng.pecan_stage(pairs)
    ms = st.run()
    res2 = st.fetch(True)
    assert ms > 0 and st.cells() == sum(r[2] for r in res) and st.launches() >= 1
    for (t, po, c), (t2, po2, c2) in zip(res, res2):
        assert np.array_equal(t, t2) and np.array_equal(po, po2) and c == c2
    solo = eng.get_aligned_pairs_using_anchors_batch(pairs[17:18], None, True)[0]
    assert np.array_equal(solo[0], res[17][0]) and np.array_equal(solo[1], res[17][1])
    for i in (0, 50, 95):
        to, poo = R.oracle_pecan_aligned_pairs(*pairs[i][:5], R.pecan_params(), 9000000)
        assert np.array_equal(res[i][0], to) and np.array_equal(res[i][1], poo), i
    for (t, po, c), q in zip(res, pairs):
        assert t[:, 0].min() >= 100000 and t[:, 0].max() <= 10000000         # threshold 0.01 .. 1
        assert len(np.unique(t[:, 1] * 100000 + t[:, 2])) == len(t)             # "points must be unique", pairwiseAligner.c:1482
        mass_x = np.bincount(t[:, 1], weights=po, minlength=len(q[0]))
        assert mass_x.max() < 1.02                                               # posterior mass per base <= 1 (up to logAdd's approximation)

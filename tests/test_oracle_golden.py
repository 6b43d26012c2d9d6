"""The oracle against the committed golden vectors (outputs of the unmodified reference, scripts/make_golden.py).
Runs everywhere, including where /root/reference and oracle/_ref do not exist. CPU only."""
import numpy as np

import _golden as G
import _reflib as R


def test_poa_golden(oracle_built):
    n = 0
    for c in G.poa_cases():
        p = R.cactus_params(**c["params"])
        tr = R.oracle_poa_msa_trace(c["seqs"], p)
        assert np.array_equal(tr["msa"], c["msa"]), c["id"]
        assert tr["read_id_map"] == c["order"], c["id"]
        assert tr["cells"] == c["cells"], c["id"]
        assert np.array_equal(np.array([a["best_score"] for a in tr["alns"]]), c["best"]), c["id"]
        assert np.array_equal(np.concatenate([a["cigar"] for a in tr["alns"]]), c["cigar"]), c["id"]
        assert np.array_equal(np.concatenate([a["dp_beg"] for a in tr["alns"]]), c["beg"]), c["id"]
        assert np.array_equal(np.concatenate([a["dp_end"] for a in tr["alns"]]), c["end"]), c["id"]
        assert np.array_equal(R.oracle_poa_msa(c["seqs"], p), c["msa"]), c["id"]
        n += 1
    assert n >= 10


def test_msa_validity_property(oracle_built):
    """the reference's own invariant (bar/tests/poaBarTest.c:19-31 validate_msa): every MSA row spells its input"""
    for c in G.poa_cases():
        msa = R.oracle_poa_msa(c["seqs"], R.cactus_params(**c["params"]))
        for row, s in zip(msa, c["seqs"]):
            assert np.array_equal(row[row != 5], s)


def test_window_golden(oracle_built):
    for c in G.window_cases():
        m = R.oracle_msa_make_partial_order_alignment(c["strs"], window_size=c["win"])
        assert m.shape == c["msa"].shape and np.array_equal(m, c["msa"]), c["id"]


def test_two_end_golden(oracle_built):
    for c in G.two_end_cases():
        ms = R.oracle_make_consistent_partial_order_alignments(c["ends"], c["ri"], c["rr"], c["ov"], window_size=c["win"])
        for a, b in zip(ms, c["msas"]):
            assert a.shape == b.shape and np.array_equal(a, b), c["id"]

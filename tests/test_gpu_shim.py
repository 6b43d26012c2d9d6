"""Drop-in check (B200): the reference's OWN compiled bar/impl/poaBarAligner.o, relinked so that its two MSA entry
points are the ones shim/cactus_bar_shim.c exports on top of libbarb200 (oracle/Makefile: libbar_shim.so), must return
the same Msa matrices as the unmodified reference library. This is the link a Cactus build makes.
Both libraries are built in the CPU container (they need /root/reference) and travel to the GPU box prebuilt; the
test is skipped only when they were never built."""
import numpy as np
import pytest

import _golden as G
import _reflib as R
from _synth import family, to_ascii, two_end_problem

pytestmark = pytest.mark.gpu



def needs_libs(fn):
    """the shim / reference libraries are prebuilt by oracle/Makefile and travel with the snapshot: on the GPU box their absence
    is a FAILURE, not a skip (a silently skipped drop-in test would read as a pass)"""
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        assert R.have_bar_shim() and R.have_bar_ref(), "oracle/_ref/libbar_shim.so / libbar_ref.so missing: run `make -C oracle` where /root/reference exists"
        return fn(*a, **k)
    return wrapped


@needs_libs
def test_shim_golden_windows_and_two_ends():
    for c in G.window_cases():
        m = R.shim_msa_make_partial_order_alignment(c["strs"], window_size=c["win"])
        assert m.shape == c["msa"].shape and np.array_equal(m, c["msa"]), c["id"]
    for c in G.two_end_cases():
        ms = R.shim_make_consistent_partial_order_alignments(c["ends"], c["ri"], c["rr"], c["ov"], window_size=c["win"])
        for a, b in zip(ms, c["msas"]):
            assert a.shape == b.shape and np.array_equal(a, b), c["id"]


@needs_libs
def test_shim_vs_reference_library():
    rng = np.random.default_rng(314)
    for it in range(10):
        K = int(rng.integers(1, 9))
        L = int(rng.choice([10, 80, 300, 900]))
        strs = [to_ascii(s) for s in family(rng, K, L, sub=0.05, ins=0.02, dele=0.02, nfrac=0.01)]
        for win in (50, 10000):
            a = R.shim_msa_make_partial_order_alignment(strs, window_size=win)
            b = R.ref_msa_make_partial_order_alignment(strs, window_size=win)
            assert a.shape == b.shape and np.array_equal(a, b), (it, win)
    for it in range(4):
        K = int(rng.integers(1, 8))
        ends, ri, rr, ov = two_end_problem(rng, K, int(rng.choice([20, 150])), sub=0.05, ins=0.02, dele=0.02)
        a = R.shim_make_consistent_partial_order_alignments(ends, ri, rr, ov)
        b = R.ref_make_consistent_partial_order_alignments(ends, ri, rr, ov)
        for x, y in zip(a, b):
            assert x.shape == y.shape and np.array_equal(x, y), it


# ---- cPecan mode: shim/cactus_pecan_shim.c linked with the reference's own pairwiseAligner.o / multipleAligner.o --------


def needs_pecan_libs(fn):
    import functools

    @functools.wraps(fn)
    def wrapped(*a, **k):
        assert R.have_pecan_shim() and R.have_pecan_ref(), "oracle/_ref/libpecan_shim.so / libpecan_ref.so missing: run `make -C oracle` where /root/reference exists"
        return fn(*a, **k)
    return wrapped


@needs_pecan_libs
def test_pecan_shim_get_aligned_pairs_golden():
    """the reference's getAlignedPairsUsingAnchors SYMBOL, served by libbarb200, against the golden triples"""
    import ctypes as C
    lib = R._load(R.PECAN_SHIM_SO)
    f = lib.pecan_ref_aligned_pairs2
    f.restype = C.c_int64
    f.argtypes = [C.c_char_p, C.c_int64, C.c_char_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.POINTER(R.PecanParams), C.c_int64,
                  C.POINTER(C.c_void_p)]
    lib.pecan_ref_free.argtypes = [C.c_void_p]
    for c in G.pecan_cases():
        p = R.pecan_params(c["threshold"], c["min_diags"], c["tb_diags"], c["expansion"])
        a = np.ascontiguousarray(c["anchors"], np.int64)
        t = C.c_void_p()
        n = f(c["sx"], len(c["sx"]), c["sy"], len(c["sy"]), a.ctypes.data, len(a), int(c["rl"]), int(c["rr"]), C.byref(p), c["split"], C.byref(t))
        got = np.ctypeslib.as_array(C.cast(t, C.POINTER(C.c_int64)), shape=(max(3 * n, 1),))[: 3 * n].reshape(n, 3).copy()
        lib.pecan_ref_free(t)
        assert np.array_equal(got, c["triples"]), c["id"]


@needs_pecan_libs
def test_pecan_shim_make_all_pairwise_alignments():
    """makeAllPairwiseAlignments (multipleAligner.c:667-680): one device batch for all pairs of an end, against the
    unmodified reference library -- incl. the reference's own MUM anchoring for matrices larger than 500 x 500"""
    from _synth import evolve
    rng = np.random.default_rng(2718)
    for K, L in [(2, 60), (4, 300), (3, 900), (5, 150)]:
        parent = rng.integers(0, 4, L).astype(np.uint8)
        seqs = [to_ascii(evolve(parent, rng, sub=0.05, ins=0.01, dele=0.01)) for _ in range(K)]
        le = [int(x) for x in rng.integers(0, 2, K)]
        re_ = [int(x) for x in rng.integers(0, 2, K)]
        a, sa = R.make_all_pairwise(seqs, le, re_, "shim")
        b, sb = R.make_all_pairwise(seqs, le, re_, "ref")
        assert np.array_equal(a, b) and np.array_equal(sa, sb), (K, L)

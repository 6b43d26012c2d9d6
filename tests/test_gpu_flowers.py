"""Flower-level drop-in proof (SURVEY.md 8 rows a12 / b): the reference's OWN make_flower_alignment_poa,
stPinchIterator_constructFromAlignedBlocks and bar() -- compiled from /root/reference into oracle/_ref/libflower_shim.so with
shim/cactus_bar_shim.c linked in place of msa_make_partial_order_alignment / make_consistent_partial_order_alignments / bar --
must produce on the B200 exactly what the unmodified reference library produces: the same ORDERED AlignmentBlock and stPinch
streams (bar/tests/poaBarTest.c:181-265 is the reference's own test of this path) and the same blocks after bar()."""
import numpy as np
import pytest

import _flowers as F
import _flower_golden as G

pytestmark = pytest.mark.gpu


def _need_shim():
    # the shim library is a build product of oracle/Makefile (prebuilt in the snapshot): its absence is a FAILURE on the GPU box
    assert F.have("shim"), "oracle/_ref/libflower_shim.so is missing: run `make -C oracle` where /root/reference exists"


def test_alignment_block_and_pinch_streams_equal_the_fixture():
    _need_shim()
    for name, fl, params, stream, bar in G.cases():
        r = F.blocks("shim", fl, params)
        assert np.array_equal(r["raw"], stream), name


def test_bar_on_one_flower_equals_the_fixture():
    _need_shim()
    for name, fl, params, stream, bar in G.cases():
        out = F.bar("shim", [fl], params)[0]
        assert np.array_equal(out, bar), name


def test_bar_over_many_flowers_goes_through_one_queue():
    """40 flowers in ONE bar() call from 4 OpenMP threads: the shim submits every end of every flower before it waits for the
    first; every flower must come out as the unmodified reference's bar() leaves it"""
    _need_shim()
    flowers = [F.random_flower(100 + s, n_threads=int(4 + s % 5), n_blocks=int(2 + s % 4), seg_len=40 + 7 * (s % 9)) for s in range(40)]
    got = F.bar("shim", flowers, threads=4)
    assert F.have("ref"), "oracle/_ref/libflower_ref.so is missing"
    want = F.bar("ref", flowers, threads=4)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i


def test_streams_equal_the_reference_library_on_fresh_seeds():
    _need_shim()
    assert F.have("ref"), "oracle/_ref/libflower_ref.so is missing"
    for seed in range(200, 212):
        fl = F.random_flower(seed, n_threads=int(3 + seed % 6), seg_len=30 + 11 * (seed % 7), lower=0.05 if seed % 3 == 0 else 0.0)
        params = {"bar/poa/partialOrderAlignmentWindow": "60"} if seed % 4 == 0 else {}
        a, b = F.blocks("shim", fl, params), F.blocks("ref", fl, params)
        assert np.array_equal(a["raw"], b["raw"]), seed


def test_cpecan_configuration_bar_equals_the_reference():
    """bar() with partialOrderAlignment="0" (SURVEY rows a13 / a14): the reference's flowerAligner.c / endAligner.c / multipleAligner.c
    over shim/cactus_pecan_shim.c -- makeAlignment's selection rounds, makeAllPairwiseAlignments and getAlignedPairsUsingAnchors
    served by the pair-HMM kernel -- leaves every flower exactly as the unmodified reference does (one thread: the reference's
    st_random() tie breaks make its own output depend on the interleaving of concurrent flowers)"""
    _need_shim()
    assert F.have("ref"), "oracle/_ref/libflower_ref.so is missing"
    from test_flowers_cpu import PECAN, pecan_flowers
    fls = pecan_flowers()
    want = F.bar("ref", fls, PECAN, threads=1)
    got = F.bar("shim", fls, PECAN, threads=1)
    for i, (a, b) in enumerate(zip(got, want)):
        assert np.array_equal(a, b), i

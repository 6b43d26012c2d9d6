#!/usr/bin/env python
"""Generate tests/golden/pecan_golden.npz from the UNMODIFIED reference cPecan (oracle/_ref/libpecan_ref.so, built from
/root/reference by oracle/Makefile): integer triples of getAlignedPairsUsingAnchors (pairwiseAligner.c:1477-1495) and the
pre-floor posteriors of getPosteriorProbsWithBanding for single sub-matrices. Run in the build container only; the fixture
is committed so that parity can be checked where /root/reference does not exist (the GPU box).

  python scripts/make_golden_pecan.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _reflib as R  # noqa: E402
from _synth import pecan_pair  # noqa: E402


def main():
    assert R.have_pecan_ref(), "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260923)
    out = {}
    # (L, anchor run length, fraction of runs kept, evolve kwargs, ragged l/r, params, split size)
    dflt = dict(threshold=0.01, min_diags=1000, tb_diags=40, expansion=20)
    spec = [
        (4, 99, 0.0, {}, 0, 0, dict(threshold=0.2, min_diags=1000, tb_diags=40, expansion=2), 3000 * 3000),   # slot for the reference's KAT, see below
        (40, 8, 1.0, {}, 0, 0, dflt, 3000 * 3000),
        (120, 99, 0.0, dict(sub=0.1, ins=0.02, dele=0.02), 0, 0, dflt, 3000 * 3000),          # no anchors: full matrix
        (300, 12, 1.0, {}, 1, 0, dflt, 3000 * 3000),
        (300, 12, 0.5, dict(sub=0.08, ins=0.02, dele=0.02), 0, 1, dflt, 3000 * 3000),
        (700, 12, 1.0, dict(nfrac=0.02), 1, 1, dflt, 3000 * 3000),
        (1500, 12, 1.0, {}, 0, 0, dflt, 3000 * 3000),                                          # > 1000 diagonals: intermediate tracebacks
        (2500, 20, 0.8, dict(sub=0.04, ins=0.01, dele=0.01), 0, 0, dflt, 3000 * 3000),
        (900, 12, 0.2, dict(sub=0.05, ins=0.01, dele=0.01), 0, 0, dflt, 120 * 120),            # splits at large anchor gaps
        (900, 12, 0.2, dict(sub=0.05, ins=0.01, dele=0.01), 1, 1, dflt, 120 * 120),
        (600, 10, 0.7, {}, 0, 0, dict(threshold=0.0001, min_diags=100, tb_diags=10, expansion=4), 3000 * 3000),
        (400, 99, 0.0, dict(sub=0.3, ins=0.05, dele=0.05), 0, 0, dict(threshold=0.01, min_diags=50, tb_diags=10, expansion=10), 3000 * 3000),
    ]
    for ci, (L, k, keep, ek, rl, rr, pk, sb) in enumerate(spec):
        if ci == 0:      # the reference's own known-answer case, submodules/cPecan/tests/pairwiseAlignerTest.c:243-322
            sx, sy, a = b"AGCG", b"AGTTCG", np.zeros((0, 2), np.int64)
        else:
            sx, sy, a = pecan_pair(rng, L, k_anchor=k, keep=keep, **ek)
        p = R.pecan_params(**pk)
        t = R.ref_pecan_aligned_pairs(sx, sy, a, rl, rr, p, sb)
        out[f"c{ci}_sx"] = np.frombuffer(sx, np.uint8)
        out[f"c{ci}_sy"] = np.frombuffer(sy, np.uint8)
        out[f"c{ci}_anchors"] = a
        out[f"c{ci}_flags"] = np.array([rl, rr, pk["min_diags"], pk["tb_diags"], pk["expansion"], sb], np.int64)
        out[f"c{ci}_thr"] = np.array([pk["threshold"]], np.float64)
        out[f"c{ci}_triples"] = t
        if len(R.oracle_pecan_split_points(len(sx), len(sy), a, sb, rl, rr)) == 1 and not (rl and rr and False):
            x, y, ps = R.ref_pecan_posteriors(sx, sy, a, rl, rr, p)          # one sub-matrix: raw doubles as well
            out[f"c{ci}_post_x"], out[f"c{ci}_post_y"], out[f"c{ci}_post"] = x, y, ps
        print(ci, len(sx), len(sy), len(a), t.shape)
    out["n_cases"] = np.array([len(spec)], np.int32)
    f = os.path.join(ROOT, "tests", "golden", "pecan_golden.npz")
    np.savez_compressed(f, **out)
    print(f, os.path.getsize(f))


if __name__ == "__main__":
    main()

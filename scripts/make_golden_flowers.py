#!/usr/bin/env python
"""Freeze flower-level outputs of the UNMODIFIED reference (oracle/_ref/libflower_ref.so, built from /root/reference by
oracle/Makefile) into tests/golden/flower_golden.npz: for every case the flat flower description (tests/_flowers.py), the
parameter overrides, the ordered AlignmentBlock / stPinch stream of make_flower_alignment_poa +
stPinchIterator_constructFromAlignedBlocks, and the canonical block list of the flower after bar().
Run here (the reference does not travel to the GPU box; the fixture does)."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _flowers as F  # noqa: E402


def cases():
    unit = {"bar/poa/partialOrderAlignmentBandConstant": "10", "bar/poa/partialOrderAlignmentBandFraction": "0.01",
            "bar/poa/partialOrderAlignmentWindow": "1000000", "bar/poa/partialOrderAlignmentMaskFilter": "5"}
    out = []
    # the reference's own two flower tests (bar/tests/poaBarTest.c:181-265)
    out.append(("shared_maxlen2", F.flowers_shared(), dict(unit, **{"bar/bandingLimit": "2", "bar/poa/partialOrderAlignmentProgressiveMaxRows": "1000",
                                                                    "bar/poa/partialOrderAlignmentProgressiveMaxLengthDiff": "0.02"})))
    out.append(("shared_iterator", F.flowers_shared(), dict(unit, **{"bar/bandingLimit": "10000", "bar/poa/partialOrderAlignmentProgressiveMaxRows": "50",
                                                                     "bar/poa/partialOrderAlignmentProgressiveMaxLengthDiff": "0.05"})))
    for seed in range(4):
        out.append(("random_%d" % seed, F.random_flower(seed), {}))
    out.append(("windows", F.random_flower(10, seg_len=90), {"bar/poa/partialOrderAlignmentWindow": "40"}))
    out.append(("masked", F.random_flower(11, lower=0.15, alphabet=b"ACGTN"), {"bar/poa/partialOrderAlignmentMaskFilter": "3"}))
    out.append(("banding_limit", F.random_flower(12, seg_len=80), {"bar/bandingLimit": "50"}))
    out.append(("wide", F.random_flower(13, n_threads=14, n_blocks=7, seg_len=220, p_skip=0.15), {}))
    out.append(("no_progressive", F.random_flower(14, n_threads=9), {"bar/poa/partialOrderAlignmentProgressiveMaxRows": "4"}))
    return out


def main():
    d = {}
    names = []
    for name, fl, params in cases():
        r = F.blocks("ref", fl, params)
        b = F.bar("ref", [fl], params)[0]
        names.append(name)
        d[name + "/flower"] = np.frombuffer(json.dumps({"n_events": fl["n_events"], "seqs": [s.decode() for s in fl["seqs"]], "seq_event": list(map(int, fl["seq_event"])),
                                                        "end_side": list(map(int, fl["end_side"])), "adj": [list(map(int, a)) for a in fl["adj"]], "params": params}).encode(), np.uint8)
        d[name + "/stream"] = r["raw"]
        d[name + "/bar"] = b
        print(name, len(fl["seqs"]), "seqs", len(fl["adj"]), "adjacencies ->", len(r["blocks"]), "alignment blocks,", len(r["pinches"]), "pinches,", int(b[0]), "blocks after bar()")
    d["names"] = np.frombuffer(json.dumps(names).encode(), np.uint8)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "flower_golden.npz"), **d)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Minimal driver for ncu: stage N synthetic 2 kbp pairs (MUM-like anchors), run the pair-HMM kernel `reps` times. Development aid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 50
eng = cb.Engine()
st = eng.pecan_stage(workload.synth_pairs(0, n, L, k_anchor=k))
for r in range(reps):
    ms = st.run()
    print("run", r, ms, "ms", st.cells() / ms / 1e6, "Gcell/s", "launches", st.launches(), flush=True)

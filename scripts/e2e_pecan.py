#!/usr/bin/env python
"""Development aid: phase times of the host-buffer pair-HMM call (BARB200_DEBUG=1)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4736
eng = cb.Engine()
pairs = workload.synth_pairs(0, n, 2000, k_anchor=50)
table = eng.pecan_table(pairs)
for r in range(3):
    t = time.time()
    raw = eng.pecan_batch_raw(table)
    dt = time.time() - t
    res = eng._take_pairs(*raw, table.n)
    print("e2e", r, dt * 1e3, "ms", sum(x[1] for x in res) / dt / 1e9, "Gcell/s", flush=True)

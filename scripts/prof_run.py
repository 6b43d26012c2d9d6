#!/usr/bin/env python
"""Minimal driver for ncu: stage N synthetic ends (8 x 2 kbp), run the kernel `reps` times. Development aid."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 8
L = int(sys.argv[4]) if len(sys.argv) > 4 else 2000
T = int(sys.argv[5]) if len(sys.argv) > 5 else 0
cps = int(sys.argv[6]) if len(sys.argv) > 6 else 0
eng = cb.Engine(cb.PoaParams(threads_per_block=T, ctas_per_sm=cps))
kw = {k: float(os.environ["SYNTH_" + k.upper()]) for k in ("sub", "ins", "dele") if "SYNTH_" + k.upper() in os.environ}   # divergence of the synthetic reads
st = eng.stage(packed=workload.synth_ends(0, n, K, L, **kw))
for r in range(reps):
    ms = st.run()
    print("run", r, ms, "ms", flush=True)
msas, cells = st.fetch()
print("cells", int(cells.sum()), "Gcell/s", cells.sum() / ms / 1e6)

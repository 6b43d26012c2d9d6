#!/usr/bin/env python
"""Development aid: sweep small shapes on the GPU against the oracle and report which fail."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_b200 as cb
import _reflib as R
from _synth import family
eng = cb.Engine()
rng = np.random.default_rng(1)
for K in (2, 3, 8):
    for L in (5, 15, 16, 17, 29, 31, 32, 33, 100, 300, 510, 511, 512, 600, 1000, 1023, 1024, 2000, 2047, 2048, 3000):
        job = family(rng, K, L, sub=0.03, ins=0.01, dele=0.01)
        try:
            m, c = eng.poa_msa_batch([job], return_cells=True)
            tr = R.oracle_poa_msa_trace(job)
            ok = m[0].shape == tr["msa"].shape and np.array_equal(m[0], tr["msa"]) and int(c[0]) == tr["cells"]
            print(K, L, [len(s) for s in job][:3], "ok" if ok else "MISMATCH", flush=True)
        except Exception as e:
            print(K, L, [len(s) for s in job][:3], "EXC", str(e)[-60:], flush=True)

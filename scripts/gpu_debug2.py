#!/usr/bin/env python
"""Development aid: the random-vs-oracle batches of tests/test_gpu_parity.py with per-job diagnostics."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_b200 as cb
import _reflib as R
from _synth import family
eng = cb.Engine()
for seed in range(3):
    rng = np.random.default_rng(500 + seed)
    jobs = []
    for it in range(48):
        K = int(rng.integers(2, 14))
        L = int(rng.choice([1, 5, 20, 60, 150, 300, 400, 800, 1500]))
        kw = dict(sub=float(rng.choice([0.0, 0.02, 0.08, 0.2])), ins=float(rng.choice([0, 0.005, 0.03])),
                  dele=float(rng.choice([0, 0.005, 0.03])), nfrac=float(rng.choice([0, 0, 0.01])))
        jobs.append(family(rng, K, L, sort=bool(rng.random() < 0.7), **kw))
    for rep in range(3):
        try:
            msas, cells = eng.poa_msa_batch(jobs, return_cells=True)
        except Exception as e:
            print("seed", seed, "rep", rep, "EXC", str(e)[-80:]); continue
        bad = []
        for j, (m, job) in enumerate(zip(msas, jobs)):
            tr = R.oracle_poa_msa_trace(job)
            if not (m.shape == tr["msa"].shape and np.array_equal(m, tr["msa"]) and int(cells[j]) == tr["cells"]):
                bad.append((j, len(job), [len(s) for s in job][:3], m.shape, tr["msa"].shape, int(cells[j]), tr["cells"]))
        print("seed", seed, "rep", rep, "bad", bad, flush=True)

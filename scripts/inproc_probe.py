#!/usr/bin/env python
"""ONE context over every visible device: barb200_poa_msa_batch with pinned host buffers, E ends per device; prints the engine's
per-batch timing lines (BARB200_TIMING) and the whole-call throughput. Development aid / evidence for the in-process multi-GPU path."""
import ctypes as C
import os
import sys
import time

import numpy as np

os.environ["BARB200_TIMING"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402
import torch  # noqa: E402

E = int(sys.argv[1]) if len(sys.argv) > 1 else 2368
eng = cb.Engine(cb.PoaParams(devices="all"))
nd = eng.device_count()
n = E * nd
n_seq, lens, flat = workload.synth_ends(0, n, 8, 2000)
pins = [torch.from_numpy(a).pin_memory() for a in (n_seq, lens, flat)]
q_nseq, q_lens, q_flat = [t.numpy() for t in pins]
for it in range(4):
    outs = (C.c_void_p * n)()
    ml = np.zeros(n, np.int32)
    cc = np.zeros(n, np.int64)
    t0 = time.time()
    eng._check(eng.lib.barb200_poa_msa_batch(eng.ctx, n, q_nseq.ctypes.data, q_lens.ctypes.data, q_flat.ctypes.data, None, outs, ml.ctypes.data, cc.ctypes.data))
    dt = time.time() - t0
    t1 = time.time()
    for i in range(n):
        eng.lib.barb200_free(outs[i])
    print("call %d: %d devices, %d ends, %.1f ms -> %.1f Gcell/s (freeing the outputs: %.1f ms)" % (it, nd, n, dt * 1e3, cc.sum() / dt / 1e9, (time.time() - t1) * 1e3), flush=True)

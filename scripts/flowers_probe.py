#!/usr/bin/env python
"""Development aid: the bench's e2e_flowers leg alone (BARB200_TIMING=1 prints the device batches)."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cactus_b200 as cb  # noqa: E402
n = int(sys.argv[1]) if len(sys.argv) > 1 else 592
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 16
eng = cb.Engine()
t0 = time.time()
res, _ = bench.flowers_leg(eng, 0, n, threads, 25.77e6)
print(json.dumps(res), "total wall %.1f s" % (time.time() - t0))

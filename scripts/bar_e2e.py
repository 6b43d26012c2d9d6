#!/usr/bin/env python
"""The reference's OWN bar() (bar/impl/bar.c) on synthetic leaf flowers, wall clock of the bar() call (alignment of every end + CAF of
every flower): the unmodified reference library (oracle/_ref/libflower_ref.so, CPU abPOA) next to the drop-in build
(oracle/_ref/libflower_shim.so: the same reference objects with shim/cactus_bar_shim.c -- incl. its bar() with the global end
queue -- linked in, running on the GPU). Checks that both leave every flower with the same blocks. This is BASELINE.json
configs[0]'s plumbing ("cactus_consolidated BAR phase ... reference path") on synthetic data; evolver data is not available
offline. Writes gpurun_out/bar_e2e.json. Run under gpurun."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _flowers as F  # noqa: E402
from bench import usable_cores  # noqa: E402


def main():
    n_flowers = int(sys.argv[1]) if len(sys.argv) > 1 else 240
    threads = usable_cores()
    t0 = time.time()
    flowers = [F.random_flower(9000 + s, n_threads=8, n_blocks=int(3 + s % 4), seg_len=int(400 + 150 * (s % 9)), sub=0.03, indel=0.008, p_empty=0.0) for s in range(n_flowers)]
    ends = sum(len(f["end_side"]) for f in flowers)
    bases = sum(len(s) for f in flowers for s in f["seqs"])
    print("built %d flowers, %d ends, %d bases in %.1f s" % (n_flowers, ends, bases, time.time() - t0), flush=True)
    out = {"flowers": n_flowers, "ends": ends, "bases": bases, "threads": threads}
    got, s_gpu = F.bar("shim", flowers, threads=threads, want_seconds=True)
    got2, s_gpu2 = F.bar("shim", flowers, threads=threads, want_seconds=True)          # second call: context and arenas warm
    n_ref = min(n_flowers, max(threads * 2, 32))
    want, s_ref = F.bar("ref", flowers[:n_ref], threads=threads, want_seconds=True)
    same = all(np.array_equal(a, b) for a, b in zip(got2[:n_ref], want)) and all(np.array_equal(a, b) for a, b in zip(got, got2))
    ref_bases = sum(len(s) for f in flowers[:n_ref] for s in f["seqs"])
    out.update({"gpu_shim_bar_seconds_first_call": s_gpu, "gpu_shim_bar_seconds": s_gpu2, "gpu_bases_per_s": bases / s_gpu2,
                "reference_bar_seconds": s_ref, "reference_flowers": n_ref, "reference_bases_per_s": ref_bases / s_ref,
                "speedup_of_the_bar_call": (bases / s_gpu2) / (ref_bases / s_ref), "identical_blocks_after_bar": bool(same)})
    print(json.dumps(out, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "bar_e2e.json"), "w"), indent=1)
    return 0 if same else 3


if __name__ == "__main__":
    sys.exit(main())

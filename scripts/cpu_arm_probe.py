#!/usr/bin/env python
"""How the reference CPU arm scales with threads on this box, with glibc's default allocator and with large blocks retained
(the jemalloc stand-in): Gcell/s and per-thread Gcell/s at 1, 2, 4, ... threads. Explains what the round-1 verdict saw on the 96-core
box (0.017 Gcell/s/thread vs 0.169 single-threaded). Writes gpurun_out/cpu_arm_scaling.json."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _reflib as R  # noqa: E402
import workload  # noqa: E402
from bench import usable_cores  # noqa: E402


def main():
    cores = usable_cores()
    K, L = 8, 2000
    n_seq, lens, flat = workload.synth_ends(0, 8 * cores, K, L)
    offs = np.concatenate([[0], np.cumsum(lens)])
    cells_per_end = 25.77e6
    out = {"usable_cores": cores, "logical_cpus": os.cpu_count(), "rows": []}
    t = 1
    ladder = []
    while t < cores:
        ladder.append(t)
        t *= 2
    ladder.append(cores)
    for th in ladder:
        n = min(len(n_seq), max(4, 6 * th))
        row = {"threads": th, "ends": n}
        for mode, name in ((0, "glibc_default"), (1, "retained_blocks")):
            R.cpu_poa_msa_many(n_seq[:min(n, th)], lens[:min(n, th) * K], flat[:offs[min(n, th) * K]], threads=th, malloc_mode=mode)
            s, kind, _ = R.cpu_poa_msa_many(n_seq[:n], lens[:n * K], flat[:offs[n * K]], threads=th, malloc_mode=mode)
            row[name] = {"gcells_per_s": cells_per_end * n / s / 1e9, "per_thread": cells_per_end * n / s / 1e9 / th}
        out["rows"].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "cpu_arm_scaling.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

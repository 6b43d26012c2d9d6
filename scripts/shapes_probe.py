#!/usr/bin/env python
"""Per-shape device throughput AND per-phase SM-clock breakdown (guide tree / DP / traceback / fusion / topological sort / MSA) of
SURVEY.md 8d's scaled-down shapes and the mixed batch. Writes gpurun_out/shapes_phases.json (copied to profiles/). Run under gpurun."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402

SHAPES = [("1000x8x200", 1000, 8, 200), ("4000x8x200", 4000, 8, 200), ("1000x4x2000", 1000, 4, 2000), ("2368x8x2000", 2368, 8, 2000), ("100x30x2000", 100, 30, 2000),
          ("10x8x10000", 10, 8, 10000), ("148x8x10000", 148, 8, 10000), ("2000x8x500", 2000, 8, 500)]
MIXED = [(2000, 8, 200), (300, 8, 2000), (8, 8, 10000)]


def main():
    eng = cb.Engine(cb.PoaParams(collect_phase_clocks=1))
    out = {}

    def run(packed):
        st = eng.stage(packed=packed)
        st.run()
        ms = min(st.run() for _ in range(2))
        ph = st.phase_clocks()
        _, cells = st.fetch()
        b = st.buckets()
        st.close()
        tot = max(1, ph["total"])
        return {"ms": ms, "gcells_per_s": float(cells.sum()) / ms / 1e6, "buckets": b,
                "phase_share": {k: round(v / tot, 4) for k, v in ph.items() if k != "total"}}
    for name, n, k, l in SHAPES:
        r = run(workload.synth_ends(7000000, n, k, l))
        r["ends_per_s"] = n / r["ms"] * 1e3
        out[name] = r
        print(name, json.dumps(r), flush=True)
    parts = [workload.synth_ends(7200000 + 100000 * i, n, k, l) for i, (n, k, l) in enumerate(MIXED)]
    mixed = (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]), np.concatenate([p[2] for p in parts]))
    out["mixed"] = run(mixed)
    print("mixed", json.dumps(out["mixed"]), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "shapes_phases.json"), "w"), indent=1)


if __name__ == "__main__":
    main()

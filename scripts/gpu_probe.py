#!/usr/bin/env python
"""Quick GPU probe: parity on a few shapes with verbose diagnostics + timing / phase clocks on the bench shape.
Writes gpurun_out/probe.json. Development aid (run under gpurun)."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cactus_b200 as cb  # noqa: E402
import workload  # noqa: E402
import _golden as G  # noqa: E402
import _reflib as R  # noqa: E402


def main():
    out = {}
    n_ends = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    cps = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    eng = cb.Engine(cb.PoaParams(collect_phase_clocks=1, threads_per_block=T, ctas_per_sm=cps))
    out["device"] = eng.device_info()
    print(out["device"], flush=True)
    bad = []
    default = R.params_dict(R.cactus_params())
    for c in G.poa_cases():
        if not all(abs(c["params"][k] - default[k]) < 1e-9 for k in c["params"]):
            continue
        try:
            msas, cells = eng.poa_msa_batch([c["seqs"]], return_cells=True)
            ok = msas[0].shape == c["msa"].shape and np.array_equal(msas[0], c["msa"])
            print("golden", c["id"], [len(s) for s in c["seqs"]][:4], "ok" if ok else "MISMATCH", msas[0].shape, c["msa"].shape,
                  int(cells[0]), c["cells"], flush=True)
            if not ok:
                bad.append(c["id"])
                if msas[0].shape == c["msa"].shape:
                    d = np.argwhere(msas[0] != c["msa"])
                    print("   first diffs", d[:5].tolist())
        except Exception as e:  # noqa: BLE001
            print("golden", c["id"], "EXC", e, flush=True)
            bad.append(c["id"])
    out["golden_bad"] = bad
    # bench shape
    n_seq, lens, flat = workload.synth_ends(0, n_ends, 8, 2000)
    t0 = time.time()
    st = eng.stage(packed=(n_seq, lens, flat))
    t1 = time.time()
    times = []
    for it in range(3):
        ms = st.run()
        times.append(ms)
        print("run", it, ms, "ms", st.phase_clocks(), flush=True)
    msas, cells = st.fetch()
    t2 = time.time()
    tot_cells = int(cells.sum())
    best = min(times)
    out["bench"] = dict(n_ends=n_ends, cells=tot_cells, kernel_ms=times, gcells_per_s=tot_cells / best / 1e6,
                        ends_per_s=n_ends / best * 1e3, stage_s=t1 - t0, total_s=t2 - t0, phase=st.phase_clocks(),
                        launches=st.launches())
    print(json.dumps(out["bench"]), flush=True)
    # spot parity on the bench shape against the oracle
    offs = np.concatenate([[0], np.cumsum(lens)])
    nbad = 0
    for e in range(min(3, n_ends)):
        job = [flat[offs[e * 8 + i]:offs[e * 8 + i + 1]] for i in range(8)]
        tr = R.oracle_poa_msa_trace(job)
        ok = msas[e].shape == tr["msa"].shape and np.array_equal(msas[e], tr["msa"]) and int(cells[e]) == tr["cells"]
        print("bench end", e, "ok" if ok else "MISMATCH", msas[e].shape, tr["msa"].shape, int(cells[e]), tr["cells"], flush=True)
        nbad += 0 if ok else 1
    out["bench_bad"] = nbad
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()

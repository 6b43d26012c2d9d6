#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (oracle/_ref/*.so, built from /root/reference by
oracle/Makefile). Run in the build container only; the fixtures are committed so that parity can be checked
where /root/reference does not exist (the GPU box).

  python scripts/make_golden.py
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import _reflib as R  # noqa: E402
from _synth import family, to_ascii, two_end_problem  # noqa: E402


def pack(seqs):
    lens = np.array([len(s) for s in seqs], np.int32)
    return lens, np.concatenate(seqs).astype(np.uint8)


def h(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def main():
    rng = np.random.default_rng(20260922)
    out = {}
    cases = []
    # (K, L, params kwargs, evolve kwargs)
    spec = [(2, 30, {}, {}), (3, 60, {}, dict(sub=0.1)), (8, 200, {}, {}), (5, 310, {}, {}), (5, 320, {}, {}),
            (12, 150, {}, dict(sub=0.08, ins=0.03, dele=0.03)), (8, 400, dict(wb=10, wf=0.01), dict(sub=0.05, ins=0.02, dele=0.02)),
            (6, 300, dict(progressive=0), {}), (20, 100, {}, dict(nfrac=0.02)), (4, 1000, {}, {}), (8, 2000, {}, {}),
            (30, 64, dict(wb=5, wf=0.0), dict(sub=0.2)), (3, 5, {}, {}), (9, 700, dict(wb=30, wf=0.02), dict(ins=0.03, dele=0.03))]
    for ci, (K, L, pk, ek) in enumerate(spec):
        seqs = family(rng, K, L, **ek)
        p = R.cactus_params(**pk)
        tr = R.ref_poa_msa_trace(seqs, p)
        lens, flat = pack(seqs)
        out[f"c{ci}_lens"] = lens
        out[f"c{ci}_flat"] = flat
        out[f"c{ci}_msa"] = tr["msa"]
        out[f"c{ci}_order"] = np.array(tr["read_id_map"], np.int32)
        out[f"c{ci}_best"] = np.array([a["best_score"] for a in tr["alns"]], np.int64)
        out[f"c{ci}_ncigar"] = np.array([len(a["cigar"]) for a in tr["alns"]], np.int32)
        out[f"c{ci}_cigar"] = np.concatenate([a["cigar"] for a in tr["alns"]]).astype(np.uint64)
        out[f"c{ci}_beg"] = np.concatenate([a["dp_beg"] for a in tr["alns"]]).astype(np.int32)
        out[f"c{ci}_end"] = np.concatenate([a["dp_end"] for a in tr["alns"]]).astype(np.int32)
        out[f"c{ci}_cells"] = np.array([tr["cells"]], np.int64)
        d = R.params_dict(p)
        out[f"c{ci}_params"] = np.array([d["wb"], d["o1"], d["e1"], d["o2"], d["e2"], d["k"], d["w"], d["min_w"],
                                         d["progressive"], d["disable_seeding"]], np.int32)
        out[f"c{ci}_wf"] = np.array([d["wf"]], np.float32)
        cases.append(ci)
    out["n_cases"] = np.array([len(cases)], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "poa_golden.npz"), **out)

    # BAR level: windows + trimming, and two-end consistency
    out = {}
    n = 0
    for K, L, win in [(3, 60, 20), (5, 200, 50), (4, 120, 10000), (6, 700, 110), (1, 40, 20), (4, 90, 5)]:
        strs = [to_ascii(s) for s in family(rng, K, L, sub=0.05, ins=0.02, dele=0.02, nfrac=0.01)]
        if K == 4 and L == 90:
            strs[-1] = b""
        msa = R.ref_msa_make_partial_order_alignment(strs, window_size=win)
        out[f"w{n}_strs"] = np.frombuffer(b"\n".join(strs), np.uint8)
        out[f"w{n}_win"] = np.array([win], np.int64)
        out[f"w{n}_msa"] = msa
        n += 1
    out["n_windows"] = np.array([n], np.int32)
    m = 0
    for K, L, win in [(4, 60, 10000), (7, 150, 10000), (3, 100, 20)]:
        ends, ri, rr, ov = two_end_problem(rng, K, L, sub=0.05, ins=0.02, dele=0.02)
        msas = R.ref_make_consistent_partial_order_alignments(ends, ri, rr, ov, window_size=win)
        for e in range(2):
            out[f"t{m}_strs{e}"] = np.frombuffer(b"\n".join(ends[e]), np.uint8)
            out[f"t{m}_rr{e}"] = np.array(rr[e], np.int64)
            out[f"t{m}_msa{e}"] = msas[e]
        out[f"t{m}_win"] = np.array([win], np.int64)
        m += 1
    out["n_two_end"] = np.array([m], np.int32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "bar_golden.npz"), **out)
    for f in ("poa_golden.npz", "bar_golden.npz"):
        print(f, os.path.getsize(os.path.join(ROOT, "tests", "golden", f)))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""bench.py -- BAR POA DP throughput (Gcell/s, ends/s) of the B200 engine vs the reference CPU BAR path.

Contract (see the round brief): `python bench.py --gpus N --steps K --warmup W` prints ONE JSON line on rank 0.
A *step* is one pass of the hot path (batched POA: all abpoa_msa jobs of a batch of ends) over one batch of synthetic
ends of BASELINE.json's shape: E ends x 8 sequences x 2 kbp per GPU (weak scaling; BASELINE.json configs[2] scaled to
what one step should take -- configs[1], evolverMammals, needs data that is not available offline).

  value   whole-job Gcell/s, inputs already resident in HBM (stage created before the timed region), device time of
          the kernel launches (CUDA events on the launching stream), max over ranks.
  e2e     the same metric through the C-ABI call a caller makes (barb200_poa_msa_batch) with HOST buffers: host
          packing + guide trees + H2D + kernel + D2H + unpack all inside the timed region.
  cells   the banded-cell definition of SURVEY.md 8d (sum of dp_end-dp_beg+1), counted by the kernel itself and
          pinned to the reference's count by the parity tests.

`--impl reference` times the reference's own CPU implementation (unmodified abPOA built from /root/reference into
oracle/_ref, AVX2, OpenMP over ends with all host threads; the plain-C oracle port if that library is absent) on a
bounded sample of the same workload and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import workload  # noqa: E402  (synthetic inputs: a plain host library shared by both arms, not product code)

K_SEQS, L_BP = 8, 2000
METRIC = "BAR POA DP Gcells/sec and ends/sec at 1/2/4/8 B200 vs reference CPU BAR"
ALGO_BYTES_PER_CELL = 32.0     # 20 B written + 12 B read per cell (int32 planes)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:  # noqa: BLE001
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, False, []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "samples": len(self.rows), "reasons": sorted(reasons)}


def usable_cores():
    """host threads the box really gives us: affinity mask capped by the cgroup CPU quota (the GPU boxes show 128
    logical CPUs but run the container under a 16-CPU quota; oversubscribing them makes the CPU arm slower, not faster)"""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()[:2]
        if quota != "max":
            n = max(1, min(n, int(float(quota) / float(period) + 0.5)))
    except Exception:  # noqa: BLE001
        pass
    return n


def cpu_baseline(n_seq, lens, flat, cells, budget_s=20.0):
    """reference CPU path on the host cores, bounded sample (about budget_s seconds of wall time)"""
    import _reflib as R
    threads = usable_cores()
    K = K_SEQS
    # calibrate on a few ends, then size the sample
    n0 = min(len(n_seq), max(2, threads))
    offs = np.concatenate([[0], np.cumsum(lens)])
    secs0, kind, _ = R.cpu_poa_msa_many(n_seq[:n0], lens[:n0 * K], flat[:offs[n0 * K]], threads=threads)
    per_end = secs0 / n0
    n = int(min(len(n_seq), max(n0, budget_s / max(per_end, 1e-6))))
    secs, kind, _ = R.cpu_poa_msa_many(n_seq[:n], lens[:n * K], flat[:offs[n * K]], threads=threads)
    c = float(np.sum(cells[:n]))
    return {"value": c / secs / 1e9, "unit": "Gcell/s", "ends_per_s": n / secs, "cores": threads, "kind": kind,
            "sample": "first %d of the step's ends (8 x 2 kbp each), %.1f s wall, OpenMP schedule(dynamic,1) over ends" % (n, secs)}, n, secs


PECAN_BYTES_PER_CELL = 120.0   # 5 fp64 states x (forward write + forward read at the traceback + backward write)


def pecan_measure(local_rank, rank, first_pair, n_pairs, steps, warmup, cpu_budget, want_cpu, no_e2e, host_threads):
    """cPecan mode (BASELINE.json configs[3]): banded pair-HMM posteriors of n_pairs synthetic
    2 kbp pairs per GPU per step with MUM-like anchors (k = 50, Cactus' setting). Runs in a process of its own (see main)
    and returns this rank's raw measurements."""
    import cactus_b200 as cb
    eng = cb.Engine(cb.PoaParams(device=local_rank, host_threads=host_threads))
    pairs = workload.synth_pairs(first_pair, n_pairs, L_BP, k_anchor=50)
    st = eng.pecan_stage(pairs)
    cells = float(st.cells())
    for _ in range(warmup):
        st.run()
    dev_ms, launches = 0.0, 0
    for _ in range(steps):
        dev_ms += st.run()             # CUDA-event time of the launch(es); the call returns after the stream is idle
        launches += st.launches()
    res = st.fetch(True)
    n_trip = int(sum(len(r[0]) for r in res))
    st.close()
    out = {"dev_ms": dev_ms, "launches": launches, "cells": cells, "n_pairs": n_pairs, "steps": steps, "e2e_ms": float("nan"), "same": True,
           "h2d": int(sum(len(q[0]) + len(q[1]) + q[2].nbytes for q in pairs)), "d2h": n_trip * 24}
    if not no_e2e:
        # the timed call is the C ABI itself (host strings + anchors in, malloc'd triples out); building the ctypes argument
        # arrays before and turning the outputs into numpy arrays after are harness work
        table = eng.pecan_table(pairs)
        eng._take_pairs(*eng.pecan_batch_raw(table), table.n)                  # warm-up (also sizes the context's caches)
        t0 = time.time()
        raw = eng.pecan_batch_raw(table)                                       # synchronous: returns with the results on the host
        out["e2e_ms"] = (time.time() - t0) * 1e3
        res2 = eng._take_pairs(*raw, table.n)
        out["same"] = bool(all(np.array_equal(a[0], b[0]) for a, b in zip(res, res2)))
    if want_cpu and rank == 0:
        try:
            import _reflib as R
            threads = usable_cores()
            samp = [(q[0], q[1], q[2], False, False) for q in pairs[: max(2, threads)]]
            s0, kind = R.cpu_pecan_many(samp, threads)
            n = int(min(n_pairs, max(len(samp), cpu_budget / max(s0 / len(samp), 1e-6))))
            samp = [(q[0], q[1], q[2], False, False) for q in pairs[:n]]
            secs, kind = R.cpu_pecan_many(samp, threads)
            c = float(sum(r[2] for r in res[:n]))
            # parity on the sample actually timed: the reference's triples must equal the engine's
            chk = R.ref_pecan_aligned_pairs(*samp[0], R.pecan_params()) if kind == "reference" else R.oracle_pecan_aligned_pairs(*samp[0], R.pecan_params())[0]
            out["cpu_baseline"] = {"value": c / secs / 1e9, "unit": "Gcell/s", "pairs_per_s": n / secs, "cores": threads, "kind": kind,
                                   "sample": "first %d of the step's pairs, %.1f s wall, one getAlignedPairsUsingAnchors call per pair on a pool of %d threads" % (n, secs, threads),
                                   "bit_identical_on_first_pair": bool(np.array_equal(chk, res[0][0]))}
            # max abs difference of the pre-floor posteriors (tolerance 1e-5; the engine is held to 0) on a few pairs,
            # against the plain-C oracle (itself pinned bit for bit to the compiled reference by tests/test_pecan_cpu.py)
            md = 0.0
            for q, r in list(zip(samp, res))[:4]:
                to, po = R.oracle_pecan_aligned_pairs(q[0], q[1], q[2], False, False, R.pecan_params())
                md = max(md, float(np.max(np.abs(po - r[1]))) if len(po) == len(r[1]) and len(po) else (0.0 if len(po) == len(r[1]) else float("inf")))
            out["cpu_baseline"]["max_abs_posterior_diff_4_pairs"] = md
        except Exception as e:  # noqa: BLE001
            out["cpu_baseline"] = {"value": None, "unit": "Gcell/s", "cores": usable_cores(), "kind": "unavailable", "sample": str(e)}
    eng.close()
    return out


def reference_pecan(cb, R, args, threads):
    """the reference's own getAlignedPairsUsingAnchors on the host cores, bounded sample of the cPecan workload; cells from the
    checker's band (sum of diagonal widths over the split regions), the same definition the engine reports"""
    pairs = workload.synth_pairs(0, min(args.pecan_pairs_per_step, 4 * max(2, threads) * (args.steps + args.warmup)), L_BP, k_anchor=50)
    samp = [(q[0], q[1], q[2], False, False) for q in pairs[: max(2, threads)]]
    s0, kind = R.cpu_pecan_many(samp, threads)
    n = int(min(len(pairs), max(len(samp), min(args.cpu_budget, 12.0) / max(s0 / len(samp), 1e-6) / max(1, args.steps + args.warmup))))
    samp = [(q[0], q[1], q[2], False, False) for q in pairs[:n]]

    def cells_of(q):
        c = 0
        sp = R.oracle_pecan_split_points(len(q[0]), len(q[1]), q[2], 3000 * 3000, False, False)
        j = 0
        for x1, y1, x2, y2 in sp:
            sub = []
            while j < len(q[2]) and q[2][j][0] + q[2][j][1] < x2 + y2:
                sub.append((q[2][j][0] - x1, q[2][j][1] - y1))
                j += 1
            L, Rr = R.oracle_pecan_band(int(x2 - x1), int(y2 - y1), np.array(sub, np.int64).reshape(-1, 2), 20)
            c += int(((Rr - L) // 2 + 1).sum())
        return c
    ncount = min(n, 8)
    cells_per_pair = sum(cells_of(q) for q in samp[:ncount]) / ncount
    for _ in range(args.warmup):
        R.cpu_pecan_many(samp, threads)
    t = 0.0
    for _ in range(args.steps):
        s, kind = R.cpu_pecan_many(samp, threads)
        t += s
    value = cells_per_pair * n * args.steps / t / 1e9
    return {"metric": "cPecan banded pair-HMM forward/backward/posterior Gcell/s (cells = sum of band diagonal widths)", "value": value, "unit": "Gcell/s",
            "impl": "reference", "pairs_per_s": n * args.steps / t, "ms_per_step": t / args.steps * 1e3, "dtype": "f64",
            "cpu_baseline": {"value": value, "unit": "Gcell/s", "cores": threads, "kind": kind,
                             "sample": "%d pairs (2 kbp, MUM-like anchors) per step, cells/pair from an exact count of %d pairs" % (n, ncount)}}


def pecan_object(mine, mx, sm, n_pairs):
    """the "pecan" object of the JSON line from rank 0's measurements (mine) and the reductions over ranks"""
    peak, peak_src = measured_peak()
    steps = mine["steps"]
    tot_cells, tot_pairs = float(sm[2]), float(sm[3])
    launch_ms = mine["dev_ms"] / max(1, mine["launches"])
    achieved = mine["cells"] * (steps / max(1, mine["launches"])) * PECAN_BYTES_PER_CELL / (launch_ms * 1e-3) / 1e9
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "ncu_summary.json")))["pecan_dram_bytes_per_cell"] * mine["cells"] * (steps / max(1, mine["launches"]))
    except Exception:  # noqa: BLE001
        pass
    out = {"metric": "cPecan banded pair-HMM forward/backward/posterior Gcell/s (cells = sum of band diagonal widths)",
           "value": tot_cells * steps / float(mx[0]) / 1e6, "unit": "Gcell/s", "pairs_per_s": tot_pairs * steps / float(mx[0]) * 1e3,
           "ms_per_step": float(mx[0]) / steps, "dtype": "f64",
           "config": {"workload": "synthetic %d pairs x %d bp per GPU per step, 2%% sub / 0.5%% ins / 0.5%% del, anchors = exact co-linear "
                                  "runs >= 50 bp (MUM-like), diagonalExpansion 20, threshold 0.01" % (n_pairs, L_BP),
                      "cells_per_pair": mine["cells"] / n_pairs},
           "e2e": {"value": tot_cells / float(mx[1]) / 1e6, "unit": "Gcell/s", "pairs_per_s": tot_pairs / float(mx[1]) * 1e3, "ms_per_step": float(mx[1]),
                   "h2d_bytes_per_step": mine["h2d"], "d2h_bytes_per_step": mine["d2h"],
                   "api": "barb200_pecan_aligned_pairs_batch (host strings + anchors -> (score, x, y) triples)", "same_as_staged": bool(mine["same"])},
           "gpu_launches": int(sm[4]),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                        "peak_source": peak_src, "kernel": "pecan_posterior_kernel", "bytes_per_cell_algorithmic": PECAN_BYTES_PER_CELL}}
    if "cpu_baseline" in mine:
        out["cpu_baseline"] = mine["cpu_baseline"]
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ends-per-step", type=int, default=int(os.environ.get("BARB200_ENDS_PER_STEP", "2368")),
                    help="ends per GPU per step (default 16 x 148 SMs)")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer legs (for runs under ncu: the streamed e2e path releases jobs "
                    "to a RUNNING kernel from the host, which deadlocks under ncu's kernel serialisation)")
    ap.add_argument("--pecan-only", action="store_true", help="internal: this process only measures the cPecan section and prints its raw numbers")
    ap.add_argument("--pecan-pairs-per-step", type=int, default=int(os.environ.get("BARB200_PECAN_PAIRS_PER_STEP", "4736")),
                    help="cPecan-mode pairs per GPU per step (default 32 x 148 SMs); 0 skips the cPecan section")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # host threads per rank: the box's usable cores shared by the ranks of this node (torchrun exports OMP_NUM_THREADS=1,
    # which would make the host side of the end-to-end calls single threaded)
    host_threads = max(1, usable_cores() // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))))
    if args.pecan_only:
        m = pecan_measure(local_rank, rank, rank * args.pecan_pairs_per_step, args.pecan_pairs_per_step, max(1, min(args.steps, 3)),
                          max(1, min(args.warmup, 2)), min(args.cpu_budget, 12.0), not args.no_cpu_baseline, args.no_e2e, host_threads)
        print("PECAN_JSON " + json.dumps(m))
        return 0
    E = args.ends_per_step
    config = {"workload": "synthetic %d ends x %d seqs x %d bp per GPU per step, Cactus default POA parameters "
                          "(convex gap 400/30/1200/1, band 1000+0.1L, progressive order), 2%% sub / 0.5%% ins / 0.5%% del" % (E, K_SEQS, L_BP),
              "ends_per_gpu_per_step": E, "seqs_per_end": K_SEQS, "bp": L_BP, "parallelism": "ends sharded over %d GPU(s)" % world,
              "l2": "working set (DP planes, ~100 MB per resident CTA) is far larger than the 126 MB L2"}

    import cactus_b200 as cb

    if args.impl == "reference":
        # the reference's CPU implementation of the path; rank 0 only
        if rank != 0:
            return 0
        import _reflib as R
        n_seq, lens, flat = workload.synth_ends(0, E, K_SEQS, L_BP)
        threads = usable_cores()
        # cells of the sample from the oracle/reference itself (bounded): per-end count through the trace is slow, so
        # use the port's cell counter on the sample actually timed
        K = K_SEQS
        offs = np.concatenate([[0], np.cumsum(lens)])
        n0 = min(E, max(2, threads))
        secs0, kind, _ = R.cpu_poa_msa_many(n_seq[:n0], lens[:n0 * K], flat[:offs[n0 * K]], threads=threads)
        n = int(min(E, max(n0, args.cpu_budget / max(secs0 / n0, 1e-6) / max(1, args.steps + args.warmup))))
        lib = R._load(R.build_oracle())
        import ctypes as C
        lib.oracle_poa_cells.restype = C.c_int64
        lib.oracle_poa_cells.argtypes = [C.POINTER(R.RefParams), C.c_int, C.c_void_p, C.c_void_p]
        p = R.cactus_params()
        # one representative end's cell count x n would be an estimate; count a few ends exactly and scale
        ncount = min(n, 4)
        csum = 0
        for e in range(ncount):
            l = np.ascontiguousarray(lens[e * K:(e + 1) * K])
            f = np.ascontiguousarray(flat[offs[e * K]:offs[(e + 1) * K]])
            csum += lib.oracle_poa_cells(C.byref(p), K, l.ctypes.data, f.ctypes.data)
        cells_per_end = csum / ncount
        for _ in range(args.warmup):
            R.cpu_poa_msa_many(n_seq[:n], lens[:n * K], flat[:offs[n * K]], threads=threads)
        t = 0.0
        for _ in range(args.steps):
            s, kind, _ = R.cpu_poa_msa_many(n_seq[:n], lens[:n * K], flat[:offs[n * K]], threads=threads)
            t += s
        value = cells_per_end * n * args.steps / t / 1e9
        line = {"metric": METRIC, "value": value, "unit": "Gcell/s", "impl": "reference", "n_gpus": args.gpus, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": config,
                "ends_per_s": n * args.steps / t,
                "cpu_baseline": {"value": value, "unit": "Gcell/s", "cores": threads, "kind": kind,
                                 "sample": "%d ends (8 x 2 kbp) per step, cells/end from an exact count of %d ends" % (n, ncount)},
                "e2e": {"value": value, "unit": "Gcell/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        if args.pecan_pairs_per_step > 0:
            try:
                line["pecan"] = reference_pecan(cb, R, args, threads)
            except Exception as e:  # noqa: BLE001
                line["pecan"] = {"error": str(e)}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    if world > 1:
        saved = os.dup(1)
        os.dup2(2, 1)              # NCCL prints its version banner on stdout at the first collective; the JSON line must stand alone
        try:
            dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    dev = torch.device("cuda", local_rank)

    # rank 0 deals the end list out (scatter) -- the only exchange the path needs before the compute
    from cactus_b200 import dist as D
    ranges = D.deal_end_ranges(E * world, world) if rank == 0 else None
    first_end, n_ends = D.scatter_end_ranges(ranges, dev) if world > 1 else (0, E)

    eng = cb.Engine(cb.PoaParams(device=local_rank, host_threads=host_threads))
    n_seq, lens, flat = workload.synth_ends(first_end, n_ends, K_SEQS, L_BP)
    stage = eng.stage(packed=(n_seq, lens, flat))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        stage.run()
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.time()
    dev_ms = 0.0
    launches = 0
    for _ in range(args.steps):
        dev_ms += stage.run()          # CUDA-event time of the launch(es) on the launching stream
        launches += stage.launches()
    barrier()
    wall_ms = (time.time() - t0) * 1e3
    msas, cells = stage.fetch()
    my_cells = float(cells.sum())

    # ---- e2e through the host-buffer C-ABI call ----
    # inputs in PINNED host memory; per step: pack + guide trees + H2D + kernels + D2H + unpack (all inside
    # barb200_poa_msa_batch), and for N > 1 the gather of every rank's MSA bytes on rank 0 over NCCL (the only data-path
    # exchange the reference-facing call needs: alignments come back to the process that owns the flowers).
    import ctypes as C
    pin = [torch.from_numpy(a).pin_memory() for a in (n_seq, lens, flat)]
    p_nseq, p_lens, p_flat = [t.numpy() for t in pin]

    def e2e_once():
        outs = (C.c_void_p * n_ends)()
        ml = np.zeros(n_ends, np.int32)
        cc = np.zeros(n_ends, np.int64)
        eng._check(eng.lib.barb200_poa_msa_batch(eng.ctx, n_ends, p_nseq.ctypes.data, p_lens.ctypes.data, p_flat.ctypes.data, None,
                                                 outs, ml.ctypes.data, cc.ctypes.data))
        d2h = int((ml.astype(np.int64) * K_SEQS).sum())
        if world > 1:
            # concatenate this rank's MSA rows and send them to rank 0
            buf = np.empty(d2h, np.uint8)
            o = 0
            for i in range(n_ends):
                nb = int(ml[i]) * K_SEQS
                C.memmove(buf.ctypes.data + o, outs[i], nb)
                o += nb
            D.gather_bytes(torch.from_numpy(buf), dev)
        for i in range(n_ends):
            eng.lib.barb200_free(outs[i])
        return d2h
    d2h, e2e_ms = 0, float("nan")
    if not args.no_e2e:
        e2e_once()
        barrier()
        t1 = time.time()
        e2e_steps = max(1, min(args.steps, 3))
        for _ in range(e2e_steps):
            d2h = e2e_once()
        barrier()
        e2e_ms = (time.time() - t1) * 1e3 / e2e_steps
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- reduce over ranks: max time, sum cells; gather alignment checksums on rank 0 ----
    mx, sm = D.reduce_stats([dev_ms, wall_ms, e2e_ms, my_cells, float(n_ends), float(launches)], dev)
    checksums = D.gather_checksums(float(sum(int(m.sum()) for m in msas[:64])), dev)
    # ---- cPecan mode (its own context: the POA arenas are released first) ----
    pecan = None
    if args.pecan_pairs_per_step > 0:
        stage.close()
        eng.close()
        # measured in a fresh process per rank (its host-side phases ran several times slower inside this one after the POA
        # section; a clean process reproduces the stand-alone numbers), all ranks at the same time; reductions happen here
        cmd = [sys.executable, os.path.abspath(__file__), "--pecan-only", "--pecan-pairs-per-step", str(args.pecan_pairs_per_step),
               "--steps", str(args.steps), "--warmup", str(args.warmup), "--cpu-budget", str(args.cpu_budget)]
        cmd += ["--no-e2e"] if args.no_e2e else []
        cmd += ["--no-cpu-baseline"] if args.no_cpu_baseline else []
        env = dict(os.environ)
        env["OMP_NUM_THREADS"] = str(host_threads)
        barrier()
        cp = subprocess.run(cmd, env=env, capture_output=True, text=True)
        mine = None
        for ln in cp.stdout.splitlines():
            if ln.startswith("PECAN_JSON "):
                mine = json.loads(ln[len("PECAN_JSON "):])
        failed = mine is None
        if failed:          # every rank still takes part in the reduction below; the section is reported as failed
            sys.stderr.write("[bench] the cPecan section failed on rank %d:\n%s\n" % (rank, cp.stderr[-2000:]))
            mine = {"dev_ms": float("inf"), "e2e_ms": float("inf"), "cells": 0.0, "n_pairs": 0, "launches": 0, "steps": 1, "h2d": 0, "d2h": 0, "same": False}
        mxp, smp = D.reduce_stats([mine["dev_ms"], mine["e2e_ms"], mine["cells"], float(mine["n_pairs"]), float(mine["launches"]), 1.0 if failed else 0.0], dev)
        if rank == 0:
            pecan = {"error": "the cPecan section failed on %d rank(s); see stderr" % int(smp[5])} if smp[5] > 0 else \
                pecan_object(mine, mxp, smp, args.pecan_pairs_per_step)
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0
    dev_ms_max, wall_ms_max, e2e_ms_max = float(mx[0]), float(mx[1]), float(mx[2])
    tot_cells, tot_ends, tot_launches = float(sm[3]), float(sm[4]), int(sm[5])
    value = tot_cells * args.steps / dev_ms_max / 1e6          # Gcell/s
    peak, peak_src = measured_peak()
    # roofline of the dominant (only) kernel: algorithmic bytes per launch / average launch duration, rank 0's launches
    launch_ms = dev_ms / max(1, launches)
    achieved = my_cells * (args.steps / max(1, launches)) * ALGO_BYTES_PER_CELL / (launch_ms * 1e-3) / 1e9
    traffic = None
    prof = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.exists(prof):
        try:
            bpc = json.load(open(prof))["dram_bytes_per_cell"]
            traffic = bpc * my_cells * (args.steps / max(1, launches))
        except Exception:  # noqa: BLE001
            traffic = None
    h2d = int(flat.nbytes + lens.nbytes * 2 + lens.size * 8 + n_seq.size * 40)
    line = {"metric": METRIC, "value": value, "unit": "Gcell/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int32", "data": "synthetic", "config": config,
            "ends_per_s": tot_ends * args.steps / dev_ms_max * 1e3,
            "wall_ms_per_step": wall_ms_max / args.steps,
            "e2e": {"value": tot_cells / e2e_ms_max / 1e6, "unit": "Gcell/s", "ends_per_s": tot_ends / e2e_ms_max * 1e3,
                    "ms_per_step": e2e_ms_max, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h + n_ends * 16,
                    "api": "barb200_poa_msa_batch (host buffers: pack + guide trees + H2D + kernel + D2H + unpack)"},
            "gpu_launches": tot_launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src, "kernel": "poa_msa_kernel",
                         "bytes_per_cell_algorithmic": ALGO_BYTES_PER_CELL},
            "clocks": sampler.summary(), "rank_checksums": checksums}
    if pecan is not None:
        line["pecan"] = pecan
    if not args.no_cpu_baseline:
        try:
            cpu, n_cpu, secs = cpu_baseline(n_seq, lens, flat, cells, args.cpu_budget)
            line["cpu_baseline"] = cpu
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"value": None, "unit": "Gcell/s", "cores": usable_cores(), "kind": "unavailable", "sample": str(e)}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Summarise one kernel of an .ncu-rep (read here, on the CPU box) into a small JSON for profiles/.
usage: ncu_summary.py <report.ncu-rep> <out.json> <cells> "<what>" [kernel-index]"""
import csv
import json
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "dram__bytes_read.sum.per_second", "dram__bytes_write.sum.per_second", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__shared_mem_per_block_static", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__t_bytes.sum", "sm__cycles_active.avg", "smsp__cycles_active.avg"]


def main():
    rep, out, cells, what = sys.argv[1], sys.argv[2], float(sys.argv[3]), sys.argv[4]
    idx = int(sys.argv[5]) if len(sys.argv) > 5 else 0
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    h, u, v = rows[0], rows[1], rows[2 + idx]
    m = {}
    for k, uu, vv in zip(h, u, v):
        if k in KEEP or k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") or \
           k.startswith("smsp__average_warp_latency_issue_stalled") or k in ("Kernel Name",):
            m[k] = {"value": vv, "unit": uu}

    def val(k, scale={"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "Tbyte": 1e12}):
        return float(m[k]["value"].replace(",", "")) * scale.get(m[k]["unit"], 1.0)
    res = {"what": what, "source": rep + " (scratch, not committed)", "cells": cells, "metrics": m}
    try:
        res["dram_bytes_per_cell"] = (val("dram__bytes_read.sum") + val("dram__bytes_write.sum")) / cells
        res["warp_inst_per_cell"] = val("smsp__inst_executed.sum") / cells
        res["cells_per_s_under_ncu"] = cells / (val("gpu__time_duration.sum") * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}[m["gpu__time_duration.sum"]["unit"]])
    except Exception as e:  # noqa: BLE001
        res["note"] = "derived figures unavailable: %s" % e
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res.get(k) for k in ("dram_bytes_per_cell", "warp_inst_per_cell", "cells_per_s_under_ncu")}))
    for k in sorted(m):
        if "stalled" in k:
            print(k, m[k]["value"])


main()

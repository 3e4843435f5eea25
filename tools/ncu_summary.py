#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into the handful of metrics the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof_x.ncu-rep [more.ncu-rep ...] > profiles/xxx.txt
    python tools/ncu_summary.py --json profiles/ncu_r02_top.json gpurun_out/prof_top.ncu-rep > profiles/xxx.txt

--json also writes the DRAM traffic of the (longest) captured launch in the form bench.py reads for `roofline.traffic`.
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_subpipe_imma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_subpipe_imma.sum", "sm__inst_executed_pipe_uniform.sum",
    "smsp__inst_executed_pipe_tmem.sum", "sm__mem_tensor_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_op_dmma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor_op_dmma.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_bytes.sum", "l1tex__t_sector_hit_rate.pct", "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum",
]
STALL = "smsp__average_warp"


def num(v):
    try:
        return float(v.replace(",", ""))
    except (ValueError, AttributeError):
        return 0.0


def main():
    argv = sys.argv[1:]
    jpath = None
    if argv and argv[0] == "--json":
        jpath, argv = argv[1], argv[2:]
    best = None
    for rep in argv:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(out)))
        if len(rows) < 3:
            print(f"## {rep}: empty")
            continue
        hdr, units = rows[0], rows[1]
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            if jpath and (best is None or num(d.get("gpu__time_duration.sum")) > best["duration"]):
                u = {k: units[hdr.index(k)] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum") if k in hdr}
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                tsc = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
                best = {"kernel": d.get("Kernel Name"), "grid": d.get("Grid Size"), "block": d.get("Block Size"),
                        "dram_bytes_read": num(d.get("dram__bytes_read.sum")) * scale.get(u.get("dram__bytes_read.sum"), 1.0),
                        "dram_bytes_write": num(d.get("dram__bytes_write.sum")) * scale.get(u.get("dram__bytes_write.sum"), 1.0),
                        "duration": num(d.get("gpu__time_duration.sum")),
                        "duration_s_under_ncu": num(d.get("gpu__time_duration.sum")) * tsc.get(u.get("gpu__time_duration.sum"), 1.0),
                        "tensor_pipe_pct": d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                        "source": rep, "how": "ncu --set full --clock-control none, one launch (per launch, like roofline.achieved)"}
            print(f"## {rep}\nkernel: {d.get('Kernel Name')}  grid {d.get('Grid Size')} block {d.get('Block Size')}")
            for k in KEYS:
                if k in d and d[k] != "":
                    print(f"  {k:75s} {d[k]:>18s} {units[hdr.index(k)]}")
            stalls = [(h, d[h]) for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio") and d[h]]
            stalls = sorted(stalls, key=lambda t: -float(t[1].replace(",", "")))[:8]
            for h, v in stalls:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {v}")
            print()
    if jpath and best:
        json.dump(best, open(jpath, "w"), indent=1)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (one or more launches) into profiles/<name>.md + JSON fragments.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_rollout_f64_v1
"""
import csv
import io
import json
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__cycles_elapsed.avg", "smsp__cycles_active.avg", "sm__cycles_elapsed.avg.per_second",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum.per_cycle_elapsed",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "sm__icc_request_hit_rate.pct", "l1tex__t_sector_pipe_lsu_mem_global_op_ld_hit_rate.pct",
]
STALLS = "smsp__average_warps_issue_stalled_"


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    launches = []
    for vals in rows[2:]:
        rec = dict(zip(hdr, zip(units, vals)))
        d = {"kernel": rec.get("Kernel Name", ("", ""))[1]}
        for k in KEYS:
            if k in rec:
                d[k] = {"unit": rec[k][0], "value": rec[k][1]}
        stalls = {k[len(STALLS):].replace("_per_issue_active.ratio", ""): float(v[1]) for k, v in rec.items()
                  if k.startswith(STALLS) and k.endswith("_per_issue_active.ratio") and v[1] not in ("", "n/a")}
        d["stalls_per_issue"] = dict(sorted(stalls.items(), key=lambda kv: -kv[1])[:8])
        launches.append(d)
    with open(out + ".json", "w") as f:
        json.dump(launches, f, indent=1)
    with open(out + ".md", "w") as f:
        f.write(f"# ncu summary of `{rep}` (ncu --set full --clock-control none)\n\n")
        for d in launches:
            f.write(f"## {d['kernel']}\n\n| metric | value | unit |\n|---|---|---|\n")
            for k in KEYS:
                if k in d:
                    f.write(f"| {k} | {d[k]['value']} | {d[k]['unit']} |\n")
            f.write("\nTop warp-stall reasons (warps stalled per issue-active cycle):\n\n")
            for k, v in d["stalls_per_issue"].items():
                f.write(f"* {k}: {v:.3f}\n")
            f.write("\n")
    print("wrote", out + ".md")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Write profiles/roofline_inputs.json from an ncu summary of the headline kernel (tools/ncu_summary.py output), stamped
with the hash of the K1 sources in the tree -- run it right after the capture, with the sources the capture ran.

    python tools/stamp_roofline_inputs.py profiles/r02_rollout_f64_v9.json
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    src = sys.argv[1]
    launches = json.load(open(src))
    k = [x for x in launches if "rk4_rollout_kernel<double" in x["kernel"]][-1]
    v = lambda n: float(k[n]["value"])
    u = lambda n: k[n]["unit"]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    dram = v("dram__bytes_read.sum") * scale[u("dram__bytes_read.sum")] + v("dram__bytes_write.sum") * scale[u("dram__bytes_write.sum")]
    cycles = v("sm__cycles_elapsed.avg")
    out = {
        "_comment": "ncu-derived inputs that bench.py copies into its roofline object; valid only for the K1 sources with this hash",
        "csrc_sha16": bench.kernel_source_sha16(),
        "k1_sources": list(bench.K1_SOURCES),
        "rk4_rollout_f64_source_file": os.path.relpath(src, ROOT).replace(".json", ".md"),
        "rk4_rollout_f64_dram_bytes_per_launch": int(dram),
        "rk4_rollout_f64_fp64_pipe_pct": round(v("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"), 1),
        "rk4_rollout_f64_issue_active_pct": round(v("smsp__issue_active.avg.pct_of_peak_sustained_active"), 1),
        "rk4_rollout_f64_smem_wavefront_pct": round(100.0 * v("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum") / (cycles * 148), 1),
        "rk4_rollout_f64_ms_under_ncu": round(v("gpu__time_duration.sum") * {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u("gpu__time_duration.sum")], 4),
        "rk4_rollout_f64_registers": int(v("launch__registers_per_thread")),
    }
    with open(os.path.join(ROOT, "profiles", "roofline_inputs.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()

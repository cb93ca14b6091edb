#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box with `ncu -i`) into profiles/<name>.md + .json.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r1_partition --elements 1073741824
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
    "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__inst_executed.sum",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("rep")
    ap.add_argument("out")
    ap.add_argument("--elements", type=int, default=0, help="elements one launch processed (for per-element figures)")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    raw = subprocess.run(["ncu", "-i", a.rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals))
        u = dict(zip(hdr, units))
        k = {"kernel": d.get("Kernel Name"), "metrics": {}}
        for key in KEYS:
            if key in d and d[key] != "":
                k["metrics"][key] = [d[key], u.get(key, "")]
        def gb(name):
            v, unit = float(d[name]), u[name]
            return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1, "Tbyte": 1e12}[unit]
        if "dram__bytes_read.sum" in d:
            k["dram_bytes_per_launch"] = gb("dram__bytes_read.sum") + gb("dram__bytes_write.sum")
            if a.elements:
                k["dram_bytes_per_element"] = k["dram_bytes_per_launch"] / a.elements
        out.append(k)
    with open(a.out + ".json", "w") as f:
        json.dump({"report": a.rep, "note": a.note, "elements_per_launch": a.elements, "kernels": out}, f, indent=1)
    with open(a.out + ".md", "w") as f:
        f.write(f"# ncu --set full summary: {a.rep}\n\n{a.note}\n\n")
        for k in out:
            f.write(f"## {k['kernel']}\n\n")
            if "dram_bytes_per_launch" in k:
                f.write(f"- DRAM read+write per launch: {k['dram_bytes_per_launch'] / 1e9:.3f} GB")
                if a.elements:
                    f.write(f" = {k['dram_bytes_per_element']:.2f} B/element (algorithmic: 32 for partition, 16 for count)")
                f.write("\n")
            for key, (v, unit) in k["metrics"].items():
                f.write(f"- `{key}` = {v} {unit}\n")
            f.write("\n")
    print("wrote", a.out + ".md/.json")


if __name__ == "__main__":
    main()

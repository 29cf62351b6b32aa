#!/usr/bin/env python
"""Condense an .ncu-rep (read here with `ncu -i ... --page raw --csv`) into the few lines DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep > profiles/r02_xxx.txt
       python tools/ncu_summary.py --json WORKLOAD ROLLOUT_STEPS_PER_LAUNCH gpurun_out/prof.ncu-rep profiles/r02_xxx.txt
           (also merges the dominant kernel's counters into profiles/ncu_counters.json, which bench.py reads)"""
import csv
import io
import json
import os
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg.per_second",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def num(d, u, key, to=None):
    """value of one metric converted to base units (bytes, ns)"""
    v = float(d[key].replace(",", ""))
    scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "usecond": 1e3, "us": 1e3, "msecond": 1e6, "ms": 1e6,
             "nsecond": 1.0, "ns": 1.0, "second": 1e9}.get(u.get(key, ""), 1.0)
    return v * scale


def main():
    argv = sys.argv[1:]
    workload, steps_per_launch, out_txt = None, None, None
    if argv and argv[0] == "--json":
        workload, steps_per_launch, rep, out_txt = argv[1], float(argv[2]), argv[3], argv[4]
    else:
        rep = argv[0]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    lines = [f"# {rep}: ncu --set full --clock-control none (per launch; cold-cache, serialised replay)"]
    last = None
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        last = (d, u)
        lines.append(f"\n== {d.get('Kernel Name', '?')[:100]}")
        for k in KEYS:
            if k in d:
                lines.append(f"{k:84s} {d[k]:>18s} {u[k]}")
    text = "\n".join(lines) + "\n"
    if out_txt:
        with open(out_txt, "w") as f:
            f.write(text)
    else:
        sys.stdout.write(text)
    if workload and last:
        d, u = last
        path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "ncu_counters.json")
        allc = json.load(open(path)) if os.path.exists(path) else {}
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        allc[workload] = {
            "kernel": d.get("Kernel Name", "?")[:60], "source": os.path.relpath(out_txt, os.path.dirname(path) + "/.."),
            "dram_bytes_per_launch": num(d, u, "dram__bytes_read.sum") + num(d, u, "dram__bytes_write.sum"),
            "duration_us_under_ncu": num(d, u, "gpu__time_duration.sum") / 1e3,
            "executed": {
                "warp_instructions_per_warp_step": inst / (steps_per_launch / 32.0),
                "issue_slot_utilisation": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]) / 100.0,
                "fp32_pipe_cycles_active": float(d["sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed"]) / 100.0,
                "active_lanes_per_instruction": float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
                "warps_active_pct": float(d["sm__warps_active.avg.pct_of_peak_sustained_active"]),
                "registers_per_thread": float(d["launch__registers_per_thread"]),
                "source": os.path.relpath(out_txt, os.path.dirname(path) + "/.."),
            },
        }
        with open(path, "w") as f:
            json.dump(allc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()

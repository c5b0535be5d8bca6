#!/usr/bin/env python
"""Summarise an ncu report (read here, no GPU needed):
  python profiles/summarize.py gpurun_out/prof_X.ncu-rep [out.md] [kernel_traffic.json] [envs] [kernel_issue.json]
The two JSON files are what bench.py reads for roofline.traffic and the lane-issue roofline."""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_active.avg.per_cycle_active",
    "smsp__warps_eligible.avg.per_cycle_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
] + ["smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s for s in (
    "long_scoreboard", "short_scoreboard", "wait", "no_instruction", "branch_resolving", "lg_throttle", "math_pipe_throttle",
    "not_selected", "dispatch_stall", "barrier", "imc_miss", "mio_throttle")]


def main():
    rep = sys.argv[1]
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    head, units, data = rows[0], rows[1], rows[2:]
    ki = head.index("Kernel Name")
    names = [r[ki].split("(")[0] for r in data]
    seen = {}
    for i, n in enumerate(names):   # k_post / k_post_events launch twice per step: task-free group, then envs with tasks
        seen[n] = seen.get(n, 0) + 1
        if seen[n] > 1:
            names[i] = n + "#%d" % seen[n]
    out = ["| metric | unit | " + " | ".join(names) + " |", "|---|---|" + "---|" * len(names)]
    traffic = {}
    for m in METRICS:
        if m not in head:
            continue
        j = head.index(m)
        vals = [r[j] for r in data]
        out.append("| %s | %s | %s |" % (m, units[j], " | ".join(vals)))
    to_b = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    to_ms = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "nsecond": 1e-6, "second": 1e3}
    jr, jw, jt = head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum"), head.index("gpu__time_duration.sum")
    envs = int(sys.argv[4]) if len(sys.argv) > 4 else 524288
    for r, n in zip(data, names):
        b = float(r[jr].replace(",", "")) * to_b[units[jr]] + float(r[jw].replace(",", "")) * to_b[units[jw]]
        traffic[n] = {"envs": envs, "dram_bytes_per_launch": b, "gpu_time_ms": float(r[jt].replace(",", "")) * to_ms[units[jt]]}
    text = "\n".join(out)
    if len(sys.argv) > 2:
        open(sys.argv[2], "a").write(text + "\n")
    else:
        print(text)
    if len(sys.argv) > 3:
        json.dump(traffic, open(sys.argv[3], "w"), indent=1)
    if len(sys.argv) > 5:
        col = lambda m: head.index(m)  # noqa: E731
        num = lambda r, m: float(r[col(m)].replace(",", ""))  # noqa: E731
        st = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio"
        issue = {}
        for r, n in zip(data, names):
            issue[n.split("::")[-1]] = {
                "issue_slot_pct": num(r, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
                "fma_pipe_pct": num(r, "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
                "alu_pipe_pct": num(r, "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active"),
                "active_lanes_per_inst": num(r, "smsp__thread_inst_executed_per_inst_executed.ratio"),
                "warps_active_pct": num(r, "sm__warps_active.avg.pct_of_peak_sustained_active"),
                "dram_pct_of_peak": num(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                "stall_long_scoreboard_per_issue": num(r, st % "long_scoreboard"),
                "stall_no_instruction_per_issue": num(r, st % "no_instruction"),
                "warp_inst_per_launch": num(r, "smsp__inst_executed.sum"),
                "envs": envs,
            }
        json.dump(issue, open(sys.argv[5], "w"), indent=1)


if __name__ == "__main__":
    main()

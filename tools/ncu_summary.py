#!/usr/bin/env python
"""One-screen summary of an `ncu --set full` report of one kernel: duration, instruction counts, lane efficiency, pipe
utilisation, stall reasons, L2 / DRAM traffic.  The numbers DESIGN.md and profiles/*.md quote come from here.

usage: python tools/ncu_summary.py REPORT.ncu-rep [--json]
"""
import argparse
import csv
import io
import json
import subprocess

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active threads per instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe instructions % of peak"),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "FMA pipe cycles active %"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers per thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic shared memory per block"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle (warps per issue)"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "stall barrier"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
    ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "stall dispatch"),
    ("lts__t_sectors_op_red.sum", "L2 sectors, reductions (RED)"),
    ("lts__t_sectors_op_atom.sum", "L2 sectors, atomics (ATOM)"),
    ("lts__t_sectors.sum", "L2 sectors, all"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
]


def load(report):
    txt = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    out = {"kernel": vals[col["Kernel Name"]] if "Kernel Name" in col else "?"}
    for key, label in KEYS:
        if key in col:
            out[key] = {"label": label, "value": vals[col[key]], "unit": units[col[key]]}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    d = load(args.report)
    if args.json:
        print(json.dumps(d, indent=1))
        return
    print("kernel: %s" % d["kernel"])
    for key, _ in KEYS:
        if key in d:
            print("%-46s %18s %s" % (d[key]["label"], d[key]["value"], d[key]["unit"]))


if __name__ == "__main__":
    main()

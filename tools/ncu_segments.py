#!/usr/bin/env python
"""Summarise an `ncu --set full --import-source on` report of one kernel by SASS segment.

Reads the source page of the report (`ncu -i REPORT --page source --csv`), groups consecutive SASS instructions with the
same execution count (= straight-line segments of the same loop nest) and prints, per segment, its share of all executed
warp instructions and the average number of active threads -- the view DESIGN.md section 4.1 argues from (the kernel is
bound by issued instructions, so "where do the instructions go" is the profile that matters).

usage: python tools/ncu_segments.py gpurun_out/final3_megakernel.ncu-rep [--min-share 0.1] > profiles/<name>.txt
"""
import argparse
import csv
import io
import subprocess
import sys


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("--min-share", type=float, default=0.1, help="hide segments below this percentage of all instructions")
    args = ap.parse_args()
    txt = subprocess.run(["ncu", "-i", args.report, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    start = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    kernel = rows[start - 1][1] if start > 0 and len(rows[start - 1]) > 1 else "?"
    hdr = rows[start]
    ia, isrc, ie, it = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Avg. Threads Executed")
    data = []
    for r in rows[start + 1:]:
        if len(r) <= ie or not r[ie].strip():
            continue
        addr = int(r[ia], 16) if r[ia].lower().startswith("0x") else int(r[ia])
        data.append((addr, r[isrc].strip(), int(r[ie]), float(r[it] or 0)))
    if not data:
        sys.exit("no per-instruction counts in the report (was it captured with --import-source on / --set full?)")
    base, total = data[0][0], sum(d[2] for d in data)
    print("# %s" % kernel)
    print("# %.3f G warp instructions executed, %d SASS instructions" % (total / 1e9, len(data)))
    print("# offset range | instrs | executions | share | avg active threads | first instruction")
    segs, first, prev = [], 0, None
    for i, d in enumerate(data):
        if prev is not None and abs(d[2] - prev) > 0.02 * max(d[2], prev, 1):
            segs.append((first, i - 1, prev))
            first = i
        prev = d[2]
    segs.append((first, len(data) - 1, prev))
    for a, b, count in segs:
        n = b - a + 1
        share = 100.0 * n * count / total
        if share >= args.min_share:
            print("%5x-%5x | %4d | %12d | %5.2f%% | %4.1f | %s" % (data[a][0] - base, data[b][0] - base, n, count, share, data[a][3], data[a][1][:70]))
    by = {}
    for _, src, count, _ in data:
        op = src.split()[0].split(".")[0] if src and not src.startswith("@") else (src.split()[1].split(".")[0] if len(src.split()) > 1 else src)
        by[op] = by.get(op, 0) + count
    top = sorted(by.items(), key=lambda kv: -kv[1])[:12]
    print("# by opcode: " + ", ".join("%s %.2f%%" % (k, 100.0 * v / total) for k, v in top))


if __name__ == "__main__":
    main()

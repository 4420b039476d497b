#!/usr/bin/env python
"""Per-source-line and per-section breakdown of executed warp instructions and stall samples from an .ncu-rep
captured with --set full --import-source on (kernel compiled with -lineinfo).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep [sections.txt]"""
import csv
import io
import subprocess
import sys

# source-line ranges of ackb_core.cuh that make up the sections of Sim::dynamics (kept in sync by hand)
SECTIONS = None


def load(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
    out = {}
    cur = None
    for row in csv.reader(io.StringIO(raw)):
        if len(row) == 2 and row[0] in ("File Path", "File Name"):
            cur = row[1].split("/")[-1]
            continue
        if not row or row[0] in ("Function Name", "Line No") or len(row) < 10:
            if row and row[0] == "Line No":
                hdr = row
            continue
        try:
            ln = int(row[0])
        except ValueError:
            continue
        ix = {h: i for i, h in reversed(list(enumerate(hdr)))}   # first occurrence of duplicated names

        def g(name):
            try:
                return int(row[ix[name]] or 0)
            except (KeyError, ValueError, IndexError):
                return 0
        out[(cur, ln)] = (row[1], g("Instructions Executed"), g("# Samples"), g("stall_wait"), g("stall_short_sb"), g("stall_no_inst"),
                          g("stall_branch_resolving"), g("stall_long_sb"))
    return out


def main():
    rep = sys.argv[1]
    rows = load(rep)
    tot_i = sum(v[1] for v in rows.values()) or 1
    tot_s = sum(v[2] for v in rows.values()) or 1
    print(f"# {rep}: {tot_i} warp instructions, {tot_s} samples")
    top = sorted(rows.items(), key=lambda kv: -kv[1][2])[: int(sys.argv[2]) if len(sys.argv) > 2 else 40]
    print("# file:line  %inst  %samples  wait short_sb no_inst branch long_sb | source")
    for (f, ln), v in top:
        print(f"{f}:{ln:5d} {100 * v[1] / tot_i:6.2f} {100 * v[2] / tot_s:6.2f}  {v[3]:5d} {v[4]:5d} {v[5]:5d} {v[6]:5d} {v[7]:5d} | {v[0].strip()[:110]}")
    # cumulative by 25-line bucket of ackb_core.cuh
    print("# ackb_core.cuh by line bucket: lines  %inst  %samples")
    b = {}
    for (f, ln), v in rows.items():
        key = (f, ln // 25 * 25)
        x = b.setdefault(key, [0, 0])
        x[0] += v[1]; x[1] += v[2]
    for (f, l0), x in sorted(b.items()):
        if x[0] * 200 > tot_i or x[1] * 200 > tot_s:
            print(f"{f}:{l0:5d}-{l0 + 24:5d} {100 * x[0] / tot_i:6.2f} {100 * x[1] / tot_s:6.2f}")


if __name__ == "__main__":
    main()

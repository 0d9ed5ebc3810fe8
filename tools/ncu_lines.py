#!/usr/bin/env python3
"""Per-source-line hot spots of an ncu report captured with --import-source on (read here, without a GPU).

usage: python tools/ncu_lines.py report.ncu-rep [N]   -> top N lines by stall samples and by executed instructions
"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None
hdr = None
rows = []
for r in csv.reader(out.splitlines()):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if len(r) >= 2 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        rows.append((cur, r))
si = hdr.index("# Samples")
ie = hdr.index("Instructions Executed")
te = hdr.index("Thread Instructions Executed")


def num(v):
    try:
        return int(v)
    except ValueError:
        return 0


ts = sum(num(r[si]) for _, r in rows) or 1
ti = sum(num(r[ie]) for _, r in rows) or 1
tt = sum(num(r[te]) for _, r in rows) or 1
print("total: %d samples, %d warp instructions, %.1f active threads / instruction" % (ts, ti, tt / ti))
print("\n-- by stall samples")
for f, r in sorted(rows, key=lambda t: -num(t[1][si]))[:top]:
    print("%5.1f%% smp %5.1f%% inst  %s:%s  %s" % (100 * num(r[si]) / ts, 100 * num(r[ie]) / ti, f, r[0], r[1].strip()[:100]))
print("\n-- by executed warp instructions")
for f, r in sorted(rows, key=lambda t: -num(t[1][ie]))[:top]:
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * num(r[ie]) / ti, 100 * num(r[si]) / ts, f, r[0], r[1].strip()[:100]))

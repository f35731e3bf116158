#!/usr/bin/env python
"""Source lines of an .ncu-rep ranked by stall samples.  usage: python profiles/top_stalls.py <file.ncu-rep> [N]"""
import csv
import subprocess
import sys
import io

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                     capture_output=True, text=True).stdout
fname, cols, per = "", None, []
for r in csv.reader(io.StringIO(out)):
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        cols = {h: i for i, h in enumerate(r)}
        continue
    if r[0].isdigit() and cols and "Instructions Executed" in cols:
        try:
            per.append((int(r[cols["Instructions Executed"]]), int(r[cols["# Samples"]]), fname, int(r[0]),
                        r[1].strip()[:100]))
        except ValueError:
            pass
tot = sum(p[0] for p in per) or 1
ts = sum(p[1] for p in per) or 1
print(f"# {rep}: source lines by stall samples (share of warp instructions / of stall samples)")
for ie, smp, fn, ln, src in sorted(per, key=lambda p: -p[1])[:top]:
    print(f"{100 * ie / tot:5.2f}% inst {100 * smp / ts:5.2f}% smp {fn[:18]:18s} {ln:5d} {src}")

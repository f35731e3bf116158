#!/usr/bin/env python
"""Source lines of an .ncu-rep ranked by excessive shared-memory wavefronts (bank conflicts,
uncoalesced cp.async).  usage: python profiles/smem_wavefronts.py <file.ncu-rep> [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 20
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
    if r[0].isdigit() and cols and "L1 Wavefronts Shared" in cols:
        try:
            per.append((int(r[cols["L1 Wavefronts Shared Excessive"]] or 0), int(r[cols["L1 Wavefronts Shared"]] or 0),
                        fname, int(r[0]), r[1].strip()[:95]))
        except ValueError:
            pass
te = sum(p[0] for p in per) or 1
tw = sum(p[1] for p in per) or 1
print(f"# {rep}: shared wavefronts {tw}, of which excessive {te}")
for ex, w, fn, ln, src in sorted(per, key=lambda p: -p[1])[:top]:
    print(f"{100 * w / tw:5.1f}% of wavefronts {100 * ex / te:5.1f}% of excessive {fn[:16]:16s} {ln:5d} {src}")

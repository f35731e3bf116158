#!/usr/bin/env python
"""Share of executed warp instructions / stall samples per part of the step kernel, from the ncu
source page (needs -lineinfo).  usage: python profiles/phase_breakdown.py <file.ncu-rep>"""
import csv
import io
import re
import subprocess
import sys

SRC = "grid-fed-rl-gym_b200/csrc/gfr_device.cuh"


def ranges():
    """Line ranges of the device header's parts, found from its own markers."""
    lines = open(SRC).read().split("\n")
    def find(pat, start=0):
        for i in range(start, len(lines)):
            if re.search(pat, lines[i]):
                return i + 1
        raise KeyError(pat)
    bt = find(r"struct BranchT")
    mo = find(r"GFR_HD double newton_mismatch")
    bu = find(r"Back-substitution root -> leaf fused with the polar update")
    ns = find(r"GFR_HD void newton_solve", bu)
    f0 = find(r"first iteration: every instance starts", ns)
    pr = find(r"the iterate is expected to have converged", f0)
    e = find(r"mismatch \+ assemble \+ eliminate, leaf -> root", pr)
    sw = find(r"GFR_HD void sweep_solve", e)
    bf = find(r"GFR_HD void branch_flow", sw)
    st = find(r"GFR_HD void step_instance", bf)
    post = find(r"bus state -> observation", st)
    rs = find(r"GFR_HD void reset_instance", post)
    ph = find(r"Philox4x32-10")
    sol = find(r"// -+ solvers")
    return [("group ops / helpers (inlined accessors, sync, reductions, rcp, sincos, atan)", 1, ph - 1),
            ("Philox + Box-Muller", ph, sol - 1),
            ("newton: flat start, branch terms", sol, mo - 1), ("newton M: mismatch-only pass", mo, bu - 1),
            ("newton B+U: back-substitution + polar update", bu, ns - 1),
            ("newton: setup", ns, f0 - 1),
            ("newton first iteration (flat-start factors)", f0, pr - 1),
            ("newton: convergence prediction", pr, e - 1),
            ("newton E: mismatch + assemble + eliminate", e, sw - 1), ("sweep", sw, bf - 1),
            ("line flows", bf, st - 1), ("step: actions, batteries, weather, loads, injections", st, post - 1),
            ("step: observation, reward, flags, state", post, rs - 1), ("reset", rs, 10 ** 9)]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"],
                         capture_output=True, text=True).stdout
    parts = ranges()
    agg = {p[0]: [0, 0] for p in parts}
    agg["CUDA math library (log, sqrt, sincospi, atan2, fmod, division)"] = [0, 0]
    agg["kernel prologue / loop (gfr_b200.cu)"] = [0, 0]
    fname, cols = "", None
    for r in csv.reader(io.StringIO(out)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1]
            continue
        if r[0] == "Line No":
            cols = {h: i for i, h in enumerate(r)}
            continue
        if r[0].isdigit() and cols:
            try:
                ie, smp = int(r[cols["Instructions Executed"]]), int(r[cols["# Samples"]])
            except ValueError:
                continue
            line = int(r[0])
            if fname.endswith("gfr_device.cuh"):
                for name, lo, hi in parts:
                    if lo <= line <= hi:
                        key = name
                        break
            elif fname.endswith("gfr_b200.cu"):
                key = "kernel prologue / loop (gfr_b200.cu)"
            else:
                key = "CUDA math library (log, sqrt, sincospi, atan2, fmod, division)"
            agg[key][0] += ie
            agg[key][1] += smp
    ti = sum(v[0] for v in agg.values()) or 1
    ts = sum(v[1] for v in agg.values()) or 1
    print(f"# {rep}: share of warp instructions executed / of stall samples")
    for k, (ie, smp) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
        if ie:
            print(f"  {100 * ie / ti:5.1f}% inst  {100 * smp / ts:5.1f}% smp   {k}")


if __name__ == "__main__":
    main()

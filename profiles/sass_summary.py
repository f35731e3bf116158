#!/usr/bin/env python
"""Static facts of the built library (no GPU needed): registers / spills / stack per kernel from `ptxas -v`, and
counts of the SASS mnemonics that matter (bulk TMA copies, cp.async, FP64 math, warp-level primitives) from
`cuobjdump -sass`.  usage: python profiles/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "grid-fed-rl-gym_b200", "csrc")
LIB = os.path.join(ROOT, "grid-fed-rl-gym_b200", "libgfr_b200.so")
sys.path.insert(0, ROOT)
from grid_fed_rl_b200 import build as b  # noqa: E402

WATCH = ("UBLKCP", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "DFMA", "DMUL", "DADD", "DSETP", "MUFU", "SHFL", "WARPSYNC",
         "MATCH", "REDUX", "BAR", "UTCMMA", "UTMALDG", "HMMA", "IMAD", "UMOV")


def demangle_short(name):
    m = re.search(r"(step_kernel|solve_kernel|reset_kernel|dense_solve_reg_kernel|dense_solve_kernel|noise_fill_kernel|"
                  r"dfma_peak_kernel|obs_convert_kernel)(I[^E]*E)?", name)
    if not m:
        return name[:60]
    args = re.findall(r"L[ib](\d+)E", name[m.start():m.start() + 80])
    return m.group(1) + ("<" + ",".join(args) + ">" if args else "")


def main():
    cmd = [b.find_nvcc(), *b.NVCC_FLAGS, "-Xptxas=-v", "-o", "/tmp/_sass_summary.so", *b.SOURCES]
    res = subprocess.run(cmd, cwd=CSRC, capture_output=True, text=True)
    assert res.returncode == 0, res.stderr[-2000:]
    print("# nvcc " + " ".join(b.NVCC_FLAGS) + " -Xptxas=-v   (registers / spills / stack per kernel)")
    cur = None
    rows = []
    for line in res.stderr.splitlines():
        m = re.search(r"Compiling entry function '([^']+)'", line)
        if m:
            cur = demangle_short(m.group(1))
        m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
        if m and cur:
            stack, ss, sl = m.groups()
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            rows.append((cur, int(m.group(1)), int(stack), int(ss), int(sl)))
            cur = None
    for name, regs, stack, ss, sl in sorted(rows):
        print(f"{name:40s} regs {regs:3d}  stack {stack:4d} B  spill stores {ss:4d} B  spill loads {sl:4d} B")
    print(f"# kernels: {len(rows)}; with spills: {sum(1 for r in rows if r[3] or r[4])}")
    sass = subprocess.run(["cuobjdump", "-sass", "/tmp/_sass_summary.so"], capture_output=True, text=True).stdout
    print("\n# cuobjdump -sass: arch and mnemonic counts per kernel family")
    print("arch:", ", ".join(sorted(set(re.findall(r"arch = (sm_\w+)", sass)))))
    fam = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle_short(m.group(1))
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur:
            fam[cur]["total"] += 1
            op = m.group(1)
            for w in WATCH:
                if op == w or (w == "BAR" and op == "BAR"):
                    fam[cur][w] += 1
    hdr = ["total"] + list(WATCH)
    show = [k for k in sorted(fam) if k.startswith(("step_kernel<8,1,1>", "step_kernel<1,0,1>", "step_kernel<2,1,1>", "step_kernel<4,0,1>",
                                                   "step_kernel<64,1,0>", "solve_kernel<8,1,1>", "reset_kernel", "dense_solve_reg_kernel<16,9>"))]
    print(f"{'kernel':34s} " + " ".join(f"{h:>7s}" for h in hdr))
    for k in show:
        print(f"{k:34s} " + " ".join(f"{fam[k][h]:7d}" for h in hdr))
    tot = collections.Counter()
    for k in fam:
        tot.update(fam[k])
    print(f"{'all kernels (' + str(len(fam)) + ')':34s} " + " ".join(f"{tot[h]:7d}" for h in hdr))
    print("\n# UBLKCP = cp.async.bulk (the TMA bulk copy that stages the feeder image), SYNCS = its mbarrier; no UTCMMA / UTMALDG /"
          "\n# HMMA: the path has no dense contraction (tree elimination is O(n)), tensor cores do not apply (SURVEY 8d)")
    os.remove("/tmp/_sass_summary.so")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed): key raw metrics + the hottest source lines.
usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep [--lines N]
       python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep --json <workload> <instances> > profiles/r02_ncu_<workload>.json
           (the facts bench.py reads back: DRAM bytes per launch, FP64 lane operations per env-step, pipe utilisations)"""
import csv
import io
import subprocess
import sys

WANT = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_bytes.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio',
        'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__cycles_elapsed.max']


def run(args):
    return subprocess.run(["ncu", "-i", *args], capture_output=True, text=True).stdout


def facts(rep, workload, instances):
    import json
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units, r = rows[0], rows[1], rows[2]

    def val(name, scale=True):
        if name not in hdr:
            return None
        i = hdr.index(name)
        v = float(r[i].replace(",", ""))
        u = units[i].lower()
        if scale:
            v *= {"gbyte": 1e9, "mbyte": 1e6, "kbyte": 1e3, "byte": 1.0}.get(u, 1.0)
        if name == "gpu__time_duration.sum":
            v *= {"us": 1e-3, "ns": 1e-6, "ms": 1.0, "s": 1e3}.get(u, 1.0)      # always ms
        return v
    # executed DFMA + DMUL + DADD thread instructions: per elapsed cycle (summed over SMSPs) x elapsed cycles
    per_cycle = sum(val(n, False) or 0.0 for n in
                    ("smsp__sass_thread_inst_executed_op_dfma_pred_on.sum.per_cycle_elapsed",
                     "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum.per_cycle_elapsed",
                     "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum.per_cycle_elapsed"))
    lane_ops = per_cycle * (val("sm__cycles_elapsed.max", False) or 0.0)
    out = {"workload": workload, "instances": int(instances), "kernel": r[hdr.index("Kernel Name")][:60],
           "capture": rep.split("/")[-1], "duration_ms_under_ncu": val("gpu__time_duration.sum", False),
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "dram_bytes_per_launch": (val("dram__bytes_read.sum") or 0) + (val("dram__bytes_write.sum") or 0),
           "fp64_lane_ops_per_env_step": lane_ops / int(instances) if lane_ops else None,
           "warp_instructions_per_env_step": (val("smsp__inst_executed.sum", False) or 0) / int(instances),
           "shared_wavefronts_per_env_step": (val("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", False) or 0) / int(instances),
           "shared_bank_conflict_wavefronts_per_env_step": (val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", False) or 0) / int(instances),
           "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
           "fp64_pipe_pct": val("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", False),
           "l1tex_data_pipe_pct": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", False),
           "dram_pct": val("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", False),
           "registers": val("launch__registers_per_thread", False), "threads_per_cta": val("launch__block_size", False),
           "warps_active_pct": val("sm__warps_active.avg.pct_of_peak_sustained_active", False)}
    print(json.dumps(out, indent=1))


def main():
    rep = sys.argv[1]
    if "--json" in sys.argv:
        i = sys.argv.index("--json")
        return facts(rep, sys.argv[i + 1], sys.argv[i + 2])
    nlines = int(sys.argv[sys.argv.index("--lines") + 1]) if "--lines" in sys.argv else 25
    rows = list(csv.reader(io.StringIO(run([rep, "--page", "raw", "--csv"]))))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}")
    for r in rows[2:]:
        print("kernel:", r[hdr.index('Kernel Name')][:100])
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                print(f"  {w:72s} {units[i]:16s} {r[i]:>20s}")
    src = run([rep, "--page", "source", "--csv", "--print-source", "sass,cuda"])
    per, fname = [], ""
    stall_cols = None
    for r in csv.reader(io.StringIO(src)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split('/')[-1]
            continue
        if r[0] == "Line No":
            stall_cols = {h: i for i, h in enumerate(r)}
            continue
        if r[0].isdigit() and stall_cols:
            try:
                per.append((int(r[stall_cols["Instructions Executed"]]), int(r[stall_cols["# Samples"]]),
                            fname, int(r[0]), r[1].strip()[:100]))
            except ValueError:
                pass
    tot = sum(p[0] for p in per) or 1
    tots = sum(p[1] for p in per) or 1
    print(f"\nhottest source lines (share of warp instructions executed / of stall samples):")
    for ie, smp, f, l, s in sorted(per, key=lambda x: -x[0])[:nlines]:
        print(f"  {100 * ie / tot:5.2f}% inst {100 * smp / tots:5.2f}% smp  {f}:{l:<4d} {s}")


if __name__ == "__main__":
    main()

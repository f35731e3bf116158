"""Runs bench.py once per BASELINE workload on one GPU and prints one summary row each
(the table kept under profiles/).  usage: python tools/bench_all.py [--steps K] [workload ...]"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ALL = ["ieee13", "ieee13_newton", "ieee34", "ieee34_newton", "ieee123", "ieee123_sweep", "synthetic1000"]
args = sys.argv[1:]
steps = "200"
if "--steps" in args:
    i = args.index("--steps")
    steps = args[i + 1]
    del args[i:i + 2]
for w in (args or ALL):
    res = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--workload", w, "--steps", steps,
                          "--warmup", "50", "--no-cpu"], capture_output=True, text=True)
    try:
        j = json.loads(res.stdout.strip().splitlines()[-1])
    except Exception:
        print(f"{w:14s} FAILED rc={res.returncode} {res.stderr[-300:]}", flush=True)
        continue
    li, c = j["config"]["launch"], j["config"]
    print(f"{w:14s} B={c['instances_per_gpu']:7d} {c['solver']:7s} lanes={li['lanes']:3d} thr={li['threads']:3d} "
          f"grid={li['grid']:4d} smem={li['smem_bytes']:6d} | {j['value']:.4e} env-steps/s {j['ms_per_step']:.3f} ms/step | "
          f"e2e {j['e2e']['value']:.4e} (serial {j['e2e']['serial_value']:.4e}) | it {j['mean_iterations']:.2f} "
          f"conv {j['converged_frac']:.4f} | hbm frac {j['roofline']['frac']:.4f} | clocks {j['clocks']['sm_mhz']} "
          f"{j['clocks']['reasons']}", flush=True)

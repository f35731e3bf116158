#!/usr/bin/env python
"""One table from the per-N bench lines of a scaling run (gpurun_out/r02_scale_{weak,strong}_n{1,2,4,8}.json).
usage: python tools/scale_table.py > profiles/r02_scaling.txt"""
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line(path):
    with open(path) as fh:
        return json.loads(fh.read().strip().splitlines()[-1])


for mode in ("weak", "strong"):
    rows = []
    for n in (1, 2, 4, 8):
        p = os.path.join(ROOT, "gpurun_out", f"r02_scale_{mode}_n{n}.json")
        if os.path.exists(p):
            rows.append((n, line(p)))
    if not rows:
        continue
    base = rows[0][1]["value"]
    print(f"# {mode} scaling, IEEE-123 Newton tol 1e-6 (python -m torch.distributed.run ... bench.py --gpus N --scaling {mode}), one 8 x B200 box")
    print(f"{'N':>2s} {'inst/GPU':>9s} {'env-steps/s':>12s} {'x N=1':>6s} {'eff':>6s} {'ms/step':>8s} {'e2e':>11s} {'host obs fp64':>13s} {'host obs fp32':>13s} {'sm MHz':>7s} reasons")
    for n, j in rows:
        e = j["e2e"]
        print(f"{n:2d} {j['config']['instances_per_gpu']:9d} {j['value']:12.4e} {j['value'] / base:6.2f} {j['value'] / base / n:6.3f} "
              f"{j['ms_per_step']:8.3f} {e['value']:11.4e} {e['with_observations']['value']:13.3e} "
              f"{(e.get('with_observations_f32') or {}).get('value', float('nan')):13.3e} {j['clocks']['sm_mhz']:7d} {j['clocks']['reasons']}")
    if mode == "weak":
        print("# the other BASELINE configurations in the same runs (env-steps/s, all GPUs):")
        names = sorted(rows[0][1].get("configs", {}))
        print(f"{'N':>2s} " + " ".join(f"{k:>22s}" for k in names))
        for n, j in rows:
            print(f"{n:2d} " + " ".join(f"{j['configs'][k]['value']:22.4e}" for k in names))
    print()

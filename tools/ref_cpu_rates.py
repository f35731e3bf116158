#!/usr/bin/env python
"""CPU rates of the REFERENCE ITSELF, measured where it is mounted (the build container; it does not
travel to the GPU box) -> profiles/r02_reference_cpu_rates.json, which bench.py carries in its JSON line.

  config 1   GridEnvironment(feeder=IEEE13Bus(), renewable_sources=['solar','wind']) exactly as shipped
             (internal 3-bus system, default RobustPowerFlowSolver), random ndarray policy, 1000 steps,
             reset on done (BASELINE.md section 2 / SURVEY 8d "config 1"), one core
  oracle     the reference's own step with deviations D1-D4 (oracle/ref_harness.OracleEnv + FixedNR) on the
             repaired IEEE-13 / 34 / 123 feeders - the loop bench.py's GPU arm replaces - one core each

usage: PYTHONDONTWRITEBYTECODE=1 python tools/ref_cpu_rates.py"""
import json
import os
import platform
import sys
import time
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh  # noqa: E402


def config1():
    import logging
    import random
    logging.disable(logging.CRITICAL)
    sys.path.insert(0, rh.REFERENCE_ROOT)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        from grid_fed_rl import GridEnvironment, IEEE13Bus
        env = GridEnvironment(feeder=IEEE13Bus(), renewable_sources=["solar", "wind"])
        random.seed(0); np.random.seed(0)
        env.reset(seed=0)
        for _ in range(20):
            env.step(np.array(env.action_space.sample()))
        env.reset(seed=0)
        steps, resets, conv = 1000, 0, 0
        t0 = time.perf_counter()
        for _ in range(steps):
            obs, r, term, trunc, info = env.step(np.array(env.action_space.sample()))
            conv += bool(info.get("power_flow_converged", False))
            if term or trunc:
                env.reset(); resets += 1
        dt = time.perf_counter() - t0
    return {"env_steps_per_s": steps / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps, "resets": resets,
            "converged": conv, "cores": 1, "obs_dim": len(obs), "note": "as shipped: internal 3-bus system, heuristic solver chain"}


def oracle_rate(ns, spec, steps):
    f = rh.make_feeder(ns, spec)
    t0 = time.perf_counter()
    out = rh.run_trace(ns, f, steps, seed=5, start_time=12 * 3600.0, tolerance=1e-6)
    dt = time.perf_counter() - t0
    return {"env_steps_per_s": steps / dt, "ms_per_step": 1e3 * dt / steps, "steps": steps,
            "converged": int(out["converged"].sum()), "mean_iterations": float(out["iterations"].mean()), "cores": 1,
            "note": "reference GridEnvironment.step + NewtonRaphsonSolver with D1-D4 (oracle/ref_harness.py), tol 1e-6"}


def main():
    ns = rh.build()
    res = {"where": "build container (the reference is not present on the GPU box)",
           "cpu": platform.processor() or platform.machine(), "logical_cpus": os.cpu_count(),
           "python": platform.python_version(), "numpy": np.__version__,
           "config1_as_shipped": config1(),
           "oracle_loop": {spec: oracle_rate(ns, spec, steps) for spec, steps in (("ieee13", 300), ("ieee34", 80), ("ieee123", 12))}}
    path = os.path.join(ROOT, "profiles", "r02_reference_cpu_rates.json")
    with open(path, "w") as fh:
        json.dump(res, fh, indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()

"""Throughput of the fused step over lane counts on synthetic radial feeders of several sizes
(to place the thresholds of topology.auto_lanes).  usage: python tools/tune_lanes.py [solver]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_fed_rl_b200 as m  # noqa: E402

solver = sys.argv[1] if len(sys.argv) > 1 else "newton"
SIZES = ((30, 262144), (60, 131072), (100, 131072), (200, 65536), (300, 32768), (500, 16384), (800, 8192))
if os.environ.get("GFR_TUNE_SIZES"):
    SIZES = tuple(s for s in SIZES if str(s[0]) in os.environ["GFR_TUNE_SIZES"].split(","))
for n, B in SIZES:
    cfg = m.NetworkConfig(num_buses=n, connectivity=0.0, load_probability=0.9, dg_probability=0.3,
                          min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
    f = m.repair_topology(m.SyntheticFeeder(cfg, seed=n))
    tot = sum(ld.base_power for ld in f.loads) / (f.parameters.base_power * 1e6)
    for ld in f.loads:
        s = 0.4 / tot
        ld.base_power *= s; ld.active_power *= s; ld.reactive_power *= s
    row = []
    for lanes in (2, 4, 8, 16, 32, 64, 128, 256):
        if lanes * 128 < n or lanes > 4 * n:
            continue
        try:
            env = m.BatchedGridEnvironment(f, B, solver=solver, lanes=lanes, repair=False,
                                           renewable_sources=["solar", "wind"], start_time=43200.0,
                                           tolerance=1e-6 if solver == "newton" else 1e-8)
        except m.GridEnvironmentError as exc:
            row.append(f"{lanes}: -")
            continue
        env.reset(seed=0)
        acts = [env.sample_actions() for _ in range(4)]
        for i in range(3):
            env.step(acts[i])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        K = 10
        for i in range(K):
            _, _, _, _, info = env.step(acts[i % 4])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        conv = float(info["power_flow_converged"].double().mean())
        row.append(f"{lanes}: {B / ms / 1e3:7.2f}M" + ("" if conv == 1.0 else f"(conv {conv:.2f})"))
        env.close()
        del env
        torch.cuda.empty_cache()
    print(f"n={n:4d} B={B:7d} {solver}: " + "  ".join(row), flush=True)

"""A small run of every kernel family for compute-sanitizer (racecheck / memcheck), e.g.
   compute-sanitizer --tool racecheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_fed_rl_b200 as m  # noqa: E402

torch.manual_seed(0)
for feeder, lanes_list in ((m.repair_topology(m.IEEE13Bus()), (1, 4, 32)),
                           (m.repair_topology(m.IEEE123Bus(seed=0)), (8, 16, 128))):
    for solver in ("newton", "sweep"):
        for lanes in lanes_list:
            env = m.BatchedGridEnvironment(feeder, 37, solver=solver, lanes=lanes, repair=False,
                                           renewable_sources=["solar", "wind"], start_time=43200.0,
                                           tolerance=1e-8 if solver == "newton" else 1e-10)
            env.reset(seed=1)
            for _ in range(2):
                act = env.sample_actions()
                act[3, 0] = float("nan")
                obs, reward, term, trunc, info = env.step(act)
            assert bool(info["power_flow_converged"][:3].all())
            env.close()
    s = m.B200PowerFlowSolver(tolerance=1e-8, lanes=lanes_list[1])
    n = len(feeder.buses)
    p = np.random.RandomState(0).uniform(-0.01, 0.0, size=(19, n)); p[:, 0] = 0
    sol = s.solve_batch(feeder, p)
    assert bool(sol.converged.all())
torch.cuda.synchronize()
print("sanitize_small: done")

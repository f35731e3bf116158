"""One dense solve of IEEE-34 with its loop line kept (66 unknowns), for ncu.
usage: python tools/prof_dense.py [B]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_fed_rl_b200 as m  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
f = m.repair_topology(m.IEEE34Bus(seed=0), keep_cycles=True)
rs = np.random.RandomState(1)
n = len(f.buses)
base = np.zeros(n)
idx = {b.id: i for i, b in enumerate(f.buses)}
for ld in f.loads:
    base[idx[ld.bus]] += ld.base_power / (f.parameters.base_power * 1e6)
p = torch.as_tensor(-base[None, :] * (0.4 / base.sum()) * rs.uniform(0.5, 1.5, size=(B, n)), device="cuda")
solver = m.B200PowerFlowSolver(tolerance=1e-6, method="dense")
for _ in range(3):
    sol = solver.solve_batch(f, p)
torch.cuda.synchronize()
print(float(sol.converged.double().mean()), float(sol.iterations.double().mean()))

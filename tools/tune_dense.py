"""Thread-grid sweep for the register-resident dense kernel (GFR_DENSE_TX = 8 | 16 | 32 columns of
8 threads; a grid that cannot hold the network falls back to the default choice).  Needs the library built
with every shape: python -m grid_fed_rl_b200.build --force -DGFR_DENSE_ALL_SHAPES
usage: python tools/tune_dense.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_fed_rl_b200 as m  # noqa: E402


def mesh(nb, conn, seed):
    return m.repair_topology(m.SyntheticFeeder(m.NetworkConfig(num_buses=nb, connectivity=conn, load_probability=0.9,
                                                               dg_probability=0.3), seed=seed), keep_cycles=True)


def injections(f, B, seed):
    rs = np.random.RandomState(seed)
    n = len(f.buses)
    base = np.zeros(n)
    idx = {b.id: i for i, b in enumerate(f.buses)}
    for ld in f.loads:
        base[idx[ld.bus]] += ld.base_power / (f.parameters.base_power * 1e6)
    return -base[None, :] * (0.4 / base.sum()) * rs.uniform(0.5, 1.5, size=(B, n))


def timed(solver, f, p, reps=3):
    sol = solver.solve_batch(f, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sol = solver.solve_batch(f, p)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, sol


cases = [("mesh-9", mesh(9, 0.3, 1), 262144), ("mesh-14", mesh(14, 0.2, 2), 262144), ("mesh-20", mesh(20, 0.1, 9), 131072),
         ("mesh-28", mesh(28, 0.08, 3), 131072),
         ("ieee34+loop", m.repair_topology(m.IEEE34Bus(seed=0), keep_cycles=True), 65536),
         ("mesh-40", mesh(40, 0.05, 3), 65536), ("mesh-47", mesh(47, 0.04, 5), 65536), ("mesh-60", mesh(60, 0.03, 7), 32768)]
for name, f, B in cases:
    N = 2 * (len(f.buses) - 1)
    p = torch.as_tensor(injections(f, B, 1), device="cuda")
    row = []
    for where in ("shared",):
        ms, sol = timed(m.B200PowerFlowSolver(tolerance=1e-6, method="dense", dense_kernel=where), f, p)
        row.append(f"shared {B / ms / 1e3:.2f}")
    R = (N + 7) // 8
    for tx in (8, 16, 32):
        if (tx == 8 and R > 9) or (tx == 16 and not 2 <= R <= 12) or (tx == 32 and R < 4) or N > 127:
            continue
        os.environ["GFR_DENSE_TX"] = str(tx)
        ms, sol = timed(m.B200PowerFlowSolver(tolerance=1e-6, method="dense", dense_kernel="registers"), f, p)
        row.append(f"8x{tx} R={R}: {B / ms / 1e3:.2f}")
    os.environ.pop("GFR_DENSE_TX", None)
    print(f"{name}: unknowns={N} B={B} its={float(sol.iterations.double().mean()):.2f} conv={float(sol.converged.double().mean()):.3f}"
          f" | M solves/s: " + " | ".join(row), flush=True)

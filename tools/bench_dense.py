"""Throughput of the dense (mesh-capable) Newton-Raphson kernel: converged solves/s on meshed
networks, next to the tree-ordered kernel on the radialised version of the same feeder and to the
numpy oracle on one host thread.  usage: python tools/bench_dense.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import grid_fed_rl_b200 as m  # noqa: E402
from oracle import port  # noqa: E402  (CPU baseline leg only)


def injections(f, B, seed):
    rs = np.random.RandomState(seed)
    n = len(f.buses)
    base = np.zeros(n)
    idx = {b.id: i for i, b in enumerate(f.buses)}
    for ld in f.loads:
        base[idx[ld.bus]] += ld.base_power / (f.parameters.base_power * 1e6)
    tot = base.sum()
    return -base[None, :] * (0.4 / tot) * rs.uniform(0.5, 1.5, size=(B, n))


def timed(solver, f, p, reps=10):
    p = torch.as_tensor(p, device="cuda")
    sol = solver.solve_batch(f, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        sol = solver.solve_batch(f, p)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, sol


cases = [("IEEE-34, loop line kept", lambda keep: m.repair_topology(m.IEEE34Bus(seed=0), keep_cycles=keep), 65536),
         ("synthetic 40 buses, connectivity 0.05", lambda keep: m.repair_topology(m.SyntheticFeeder(
             m.NetworkConfig(num_buses=40, connectivity=0.05, load_probability=0.9, dg_probability=0.4), seed=3),
             keep_cycles=keep), 65536),
         ("synthetic 80 buses, connectivity 0.02", lambda keep: m.repair_topology(m.SyntheticFeeder(
             m.NetworkConfig(num_buses=80, connectivity=0.02, load_probability=0.9, dg_probability=0.4), seed=5),
             keep_cycles=keep), 16384)]
for name, make, B in cases:
    mesh, tree = make(True), make(False)
    p = injections(mesh, B, 1)
    dense = m.B200PowerFlowSolver(tolerance=1e-6, method="dense")
    ms, sol = timed(dense, mesh, p)
    conv, its = float(sol.converged.double().mean()), float(sol.iterations.double().mean())
    newton = m.B200PowerFlowSolver(tolerance=1e-6, method="newton")
    ms_t, sol_t = timed(newton, tree, p)
    net = port.DenseNetwork(mesh.buses, mesh.lines)
    nb = 256
    t0 = time.perf_counter()
    port.newton_raphson(net, p[:nb], 1e-6, 50)
    cpu = nb / (time.perf_counter() - t0)
    unknowns = 2 * (len(mesh.buses) - 1)
    print(f"{name}: n={len(mesh.buses)} lines={len(mesh.lines)} (tree {len(tree.lines)}) unknowns={unknowns} B={B} | "
          f"dense {B / ms / 1e3:.3f} M solves/s ({ms:.2f} ms, {its:.2f} its, conv {conv:.3f}) | "
          f"tree-ordered on the radialised feeder {B / ms_t / 1e3:.2f} M solves/s | numpy oracle, 1 thread {cpu:.0f} solves/s",
          flush=True)

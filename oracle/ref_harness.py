"""Runs the UNMODIFIED reference (imported from /root/reference) with deviations D1-D4 and
freezes its outputs as golden vectors - TEST INFRASTRUCTURE, runs only where the reference is
mounted (the build container).  ``python oracle/ref_harness.py`` rewrites ``tests/golden/*.npz``.

No reference function body is restated here: every deviation is a subclass hook or an adapter.

  D1  PerUnitAdapter  - watts <-> per-unit around the solver plugin call
  D2  FixedNR         - J11 diagonal: ``J[r,r] -= 2 Vm_i^2 Im(Y_ii)`` after the reference's own
                        ``_build_jacobian`` (reference power_flow.py:247-248)
  D3  OracleEnv       - ``_initialize_feeder`` populates the env from the feeder instead of the
                        hard-coded 3-bus system (reference grid_env.py:243-298);
                        battery injections land on the battery's own bus (not the literal 2)
  D4  repair_topology - applied by the caller, upstream of the oracle and the kernel alike
  RNG ReplayRNG       - the reference draws from the global ``random`` / ``np.random`` streams
                        (grid_env.py:670-681, dynamics.py:69); the harness feeds the same
                        pre-generated numbers to the reference and to the kernel
"""

from __future__ import annotations

import os
import random as _py_random
import sys
import warnings
from collections import deque
from typing import Dict, Optional

import numpy as np

REFERENCE_ROOT = os.environ.get("GFR_REFERENCE_ROOT", "/root/reference")
REPO_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "grid_fed_rl"))


def _import_reference():
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    import logging
    logging.disable(logging.CRITICAL)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import grid_fed_rl  # noqa: F401
    from grid_fed_rl.environments import dynamics, grid_env, power_flow
    from grid_fed_rl import feeders
    from grid_fed_rl.feeders import synthetic
    return grid_env, power_flow, dynamics, feeders, synthetic


def build():
    """Returns a namespace with the harness classes bound to the imported reference."""
    grid_env, power_flow, dynamics, feeders, synthetic = _import_reference()

    class FixedNR(power_flow.NewtonRaphsonSolver):                       # D2
        def _build_jacobian(self, Y, V, buses, slack_bus, pv_buses, pq_buses):
            J = super()._build_jacobian(Y, V, buses, slack_bus, pv_buses, pq_buses)
            row = 0
            for i in range(len(buses)):
                if i == slack_bus:
                    continue
                J[row, row] -= 2.0 * abs(V[i]) ** 2 * Y.imag[i, i]
                row += 1
            return J

    class PerUnitAdapter(power_flow.PowerFlowSolver):                    # D1
        def __init__(self, inner, s_base):
            super().__init__(inner.tolerance, inner.max_iterations)
            self.inner, self.s_base = inner, float(s_base)

        def solve(self, buses, lines, loads, generation):
            sol = self.inner.solve(buses, lines,
                                   {k: v / self.s_base for k, v in loads.items()},
                                   {k: v / self.s_base for k, v in generation.items()})
            sol.line_flows = sol.line_flows * self.s_base
            sol.line_loadings = sol.line_loadings * self.s_base   # |S_pu| S_base / rating
            sol.losses = sol.losses * self.s_base
            self.last = sol
            return sol

    class OracleEnv(grid_env.GridEnvironment):                           # D3
        def _initialize_feeder(self):
            f = self.feeder
            self.buses, self.lines, self.loads = list(f.buses), list(f.lines), list(f.loads)
            if self.stochastic_loads:
                for load in self.loads:
                    self.dynamics.add_load_model(load.id, dynamics.TimeVaryingLoadModel())
            self.battery_bus = {}
            for gid, info in f.generators.items():
                kind = info.get("type")
                if kind == "solar" and "solar" in self.renewable_sources:
                    eff = info.get("efficiency", 0.18)
                    model = dynamics.SolarPVModel(efficiency=eff,
                                                  panel_area=info["capacity"] / (eff * 1000))
                elif kind == "wind" and "wind" in self.renewable_sources:
                    model = dynamics.WindTurbineModel(info.get("cut_in_speed", 3.0),
                                                      info.get("rated_speed", 12.0),
                                                      info.get("cut_out_speed", 25.0))
                elif kind == "battery":
                    self.batteries[gid] = dynamics.BatteryModel(
                        info["capacity_kwh"], info["power_rating_kw"] * 1e3, info["efficiency"], 0.5)
                    self.dynamics.add_battery_model(gid, self.batteries[gid])
                    self.battery_bus[gid] = info["bus"]
                    continue
                else:
                    continue
                self.generators[gid] = {"type": kind, "bus": info["bus"],
                                        "capacity": info["capacity"], "model": model}
                self.dynamics.add_renewable_model(gid, model)
            if not self.batteries:
                home = self.loads[0].bus if self.loads else self.buses[0].id
                gid = f"battery_{home}"
                self.batteries[gid] = dynamics.BatteryModel(capacity=1e3, power_rating=0.5e6,
                                                            efficiency=0.95, initial_soc=0.5)
                self.dynamics.add_battery_model(gid, self.batteries[gid])
                self.battery_bus[gid] = home

        def _calculate_power_injections(self):
            # the reference's method with its literal ``bus_id = 2`` neutralised: run it with
            # idle batteries, then place each battery's power on that battery's bus
            saved = {k: b.current_power for k, b in self.batteries.items()}
            for b in self.batteries.values():
                b.current_power = 0.0
            try:
                loads, gen = super()._calculate_power_injections()
            finally:
                for k, b in self.batteries.items():
                    b.current_power = saved[k]
            for k, b in self.batteries.items():
                bus = self.battery_bus[k]
                if b.current_power > 0:
                    gen[bus] = gen.get(bus, 0) + b.current_power
                elif b.current_power < 0:
                    loads[bus] = loads.get(bus, 0) + abs(b.current_power)
            return loads, gen

    class ReplayRNG:
        """Feeds queued numbers to ``random.random`` / ``random.gauss`` / ``np.random.normal``."""

        def __init__(self):
            self.q = deque()

        def feed(self, values):
            self.q.extend(float(v) for v in values)

        def __enter__(self):
            self._saved = (_py_random.random, _py_random.gauss, np.random.normal)
            _py_random.random = lambda: self.q.popleft()
            _py_random.gauss = lambda mu, sigma: mu + self.q.popleft() * sigma
            np.random.normal = lambda loc=0.0, scale=1.0, size=None: loc + scale * self.q.popleft()
            return self

        def __exit__(self, *exc):
            _py_random.random, _py_random.gauss, np.random.normal = self._saved
            return False

    class NS:
        pass
    ns = NS()
    ns.FixedNR, ns.PerUnitAdapter, ns.OracleEnv, ns.ReplayRNG = FixedNR, PerUnitAdapter, OracleEnv, ReplayRNG
    ns.power_flow, ns.grid_env, ns.dynamics, ns.feeders, ns.synthetic = power_flow, grid_env, dynamics, feeders, synthetic
    return ns


# --------------------------------------------------------------------------- golden cases

def make_feeder(ns, spec: str, use_reference_classes: bool = True):
    """``spec``: 'radial<N>', 'ieee13', 'ieee34', 'ieee123', 'synthetic<N>:<seed>'.
    IEEE-34 / IEEE-123 are constructed right after ``np.random.seed(0)`` (D4-i)."""
    sys.path.insert(0, REPO_ROOT)
    import grid_fed_rl_b200 as mine
    if spec.startswith("radial"):
        n = int(spec[6:])
        f = ns.feeders.SimpleRadialFeeder(num_buses=n) if use_reference_classes else mine.SimpleRadialFeeder(n)
    elif spec == "fixture3":
        # the reference's own 3-bus fixture (grid_env.py:249-265 = test_environment_fixed.py:84-100)
        f = ns.feeders.SimpleRadialFeeder(num_buses=3) if use_reference_classes else mine.SimpleRadialFeeder(3)
        f.lines[1].resistance, f.lines[1].reactance, f.lines[1].rating = 0.015, 0.025, 3e6
        for ld, p in zip(f.loads, (2e6, 1.5e6)):
            ld.base_power = ld.active_power = p
            ld.reactive_power = p * np.tan(np.arccos(ld.power_factor))
    elif spec == "ieee13":
        f = ns.feeders.IEEE13Bus() if use_reference_classes else mine.IEEE13Bus()
    elif spec == "ieee34":
        if use_reference_classes:
            st = np.random.get_state(); np.random.seed(0); f = ns.feeders.IEEE34Bus(); np.random.set_state(st)
        else:
            f = mine.IEEE34Bus(seed=0)
    elif spec == "ieee123":
        if use_reference_classes:
            st = np.random.get_state(); np.random.seed(0); f = ns.feeders.IEEE123Bus(); np.random.set_state(st)
        else:
            f = mine.IEEE123Bus(seed=0)
    elif spec.startswith("synthetic"):
        n, seed = spec[9:].split(":")
        cfg = dict(num_buses=int(n), connectivity=0.0, load_probability=0.9, dg_probability=0.4,
                   min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
        if use_reference_classes:
            f = ns.synthetic.SyntheticFeeder(ns.synthetic.NetworkConfig(**cfg), seed=int(seed))
        else:
            f = mine.SyntheticFeeder(mine.NetworkConfig(**cfg), seed=int(seed))
    elif spec.startswith("mesh"):
        # meshed: 'mesh<N>:<seed>:<connectivity>' (SyntheticFeeder with extra ties), 'meshieee34' (the
        # shipped IEEE-34 with its loop-closing line kept); repaired with keep_cycles=True
        if spec in ("meshieee34", "meshieee123"):
            cls = "IEEE34Bus" if spec == "meshieee34" else "IEEE123Bus"
            if use_reference_classes:
                st = np.random.get_state(); np.random.seed(0); f = getattr(ns.feeders, cls)(); np.random.set_state(st)
            else:
                f = getattr(mine, cls)(seed=0)
        else:
            n, seed, conn = spec[4:].split(":")
            cfg = dict(num_buses=int(n), connectivity=float(conn), load_probability=0.9, dg_probability=0.4,
                       min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
            if use_reference_classes:
                f = ns.synthetic.SyntheticFeeder(ns.synthetic.NetworkConfig(**cfg), seed=int(seed))
            else:
                f = mine.SyntheticFeeder(mine.NetworkConfig(**cfg), seed=int(seed))
        return mine.repair_topology(f, keep_cycles=True)
    else:
        raise ValueError(spec)
    return mine.repair_topology(f)


def run_trace(ns, feeder, steps: int, seed: int, *, renewable_sources=("solar", "wind"),
              timestep=1.0, episode_length=86400, start_time=0.0, tolerance=1e-6,
              max_iterations=50, stochastic_loads=True, weather_variation=True,
              nan_steps=(), load_scale: Optional[float] = None) -> Dict[str, np.ndarray]:
    """One env driven for ``steps`` steps with a seeded uniform policy; reset on done."""
    if load_scale is not None:
        for ld in feeder.loads:
            ld.base_power *= load_scale
            ld.active_power *= load_scale
            ld.reactive_power *= load_scale
    s_base = feeder.parameters.base_power * 1e6
    solver = ns.PerUnitAdapter(ns.FixedNR(tolerance=tolerance, max_iterations=max_iterations), s_base)
    env = ns.OracleEnv(feeder, timestep=timestep, episode_length=episode_length,
                       stochastic_loads=stochastic_loads, renewable_sources=list(renewable_sources),
                       weather_variation=weather_variation, power_flow_solver=solver)
    L, A, D = len(env.loads), env.action_space.shape[0], len(env.get_observation())
    rs = np.random.RandomState(seed)
    rng = ns.ReplayRNG()
    n_w = 4 if weather_variation else 0
    n_l = L if stochastic_loads else 0
    rec = {k: [] for k in ("actions", "noise", "obs", "reward", "terminated", "truncated", "error",
                           "converged", "iterations", "max_voltage", "min_voltage", "losses",
                           "violations", "viol_count", "current_step", "episode_reward",
                           "reset_before", "reset_noise", "reset_obs")}

    def draw():
        row = np.zeros(4 + L)
        row[0] = rs.random_sample()
        row[1:] = rs.standard_normal(3 + L)
        return row

    def do_reset():
        row = draw()
        rng.feed(row[:n_w])
        obs, _ = env.reset()
        assert not rng.q
        env.current_time = start_time
        rec["reset_noise"].append(row[:4]); rec["reset_obs"].append(np.array(obs, dtype=float))

    with rng, warnings.catch_warnings():
        warnings.simplefilter("ignore")
        do_reset()
        need_reset = False
        for t in range(steps):
            rec["reset_before"].append(need_reset)
            if need_reset:
                do_reset()
            act = rs.uniform(-1.0, 1.0, size=A)
            if t in nan_steps:
                act[t % A] = np.nan if (t // 2) % 2 == 0 else np.inf
            row = draw()
            bad = not np.all(np.isfinite(act)) and A > 1
            if not bad:
                rng.feed(np.concatenate([row[:n_w], row[4:4 + n_l]]))
            obs, reward, term, trunc, info = env.step(act)
            assert not rng.q, "the reference consumed fewer draws than the harness fed"
            err = "error" in info
            assert err == bad
            rec["actions"].append(act); rec["noise"].append(row)
            rec["obs"].append(np.array(obs, dtype=float)); rec["reward"].append(reward)
            rec["terminated"].append(term); rec["truncated"].append(trunc); rec["error"].append(err)
            viol = info.get("constraint_violations", {}) if not err else {}
            rec["violations"].append([bool(viol.get(k, False)) for k in
                                      ("voltage_high", "voltage_low", "frequency_high", "frequency_low")])
            rec["converged"].append(bool(info.get("power_flow_converged", False)))
            rec["iterations"].append(0 if err else int(solver.last.iterations))
            rec["max_voltage"].append(float(info.get("max_voltage", max(b.voltage_magnitude for b in env.buses))))
            rec["min_voltage"].append(float(info.get("min_voltage", min(b.voltage_magnitude for b in env.buses))))
            rec["losses"].append(float(info.get("total_losses", 0.0)))
            rec["viol_count"].append(env.constraint_violations)
            rec["current_step"].append(env.current_step)
            rec["episode_reward"].append(env.episode_reward)
            need_reset = bool(term or trunc)
    out = {k: np.array(v) for k, v in rec.items()}
    out["meta"] = np.array([steps, seed, timestep, episode_length, start_time, tolerance,
                            max_iterations, float(stochastic_loads), float(weather_variation),
                            1.0 if load_scale is None else load_scale], dtype=float)
    out["renewable_sources"] = np.array(list(renewable_sources))
    return out


def solve_cases(ns, feeder, seed: int, count: int, tolerance: float, max_iterations: int = 50,
                scale: float = 1.0) -> Dict[str, np.ndarray]:
    """Direct solver known-answer vectors: random P-only injections -> reference FixedNR."""
    rs = np.random.RandomState(seed)
    n, m = len(feeder.buses), len(feeder.lines)
    base = np.zeros(n)
    idx = {b.id: i for i, b in enumerate(feeder.buses)}
    for ld in feeder.loads:
        base[idx[ld.bus]] += ld.base_power / (feeder.parameters.base_power * 1e6)
    solver = ns.FixedNR(tolerance=tolerance, max_iterations=max_iterations)
    P = np.zeros((count, n))
    keys = ("converged", "iterations", "bus_voltages", "bus_angles", "line_flows", "line_loadings",
            "losses", "max_mismatch")
    rec = {k: [] for k in keys}
    for c in range(count):
        mult = scale * rs.uniform(0.2, 1.6, size=n)
        p = -base * mult
        flip = rs.random_sample(n) < 0.1
        p = np.where(flip, -0.5 * p, p)           # some buses export
        p[[i for i, b in enumerate(feeder.buses) if b.bus_type == "slack"]] = 0.0
        P[c] = p
        loads = {b.id: -p[i] for i, b in enumerate(feeder.buses) if p[i] < 0}
        gen = {b.id: p[i] for i, b in enumerate(feeder.buses) if p[i] > 0}
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            sol = solver.solve(feeder.buses, feeder.lines, loads, gen)
        for k in keys:
            rec[k].append(getattr(sol, k))
    out = {k: np.array(v) for k, v in rec.items()}
    out["p_spec"] = P
    out["meta"] = np.array([tolerance, max_iterations], dtype=float)
    return out


TRACE_CASES = {
    # name: (feeder spec, steps, seed, kwargs)
    "trace_fixture3_s0": ("fixture3", 120, 0, dict(episode_length=50)),
    "trace_radial13_s1": ("radial13", 150, 1, dict(load_scale=0.5, start_time=11.9 * 3600, timestep=60.0)),
    "trace_ieee13_s0": ("ieee13", 300, 0, dict(nan_steps=(17, 18, 140))),
    "trace_ieee13_s1": ("ieee13", 300, 1, dict(start_time=12 * 3600.0, timestep=30.0, tolerance=1e-8)),
    "trace_ieee13_s2": ("ieee13", 200, 2, dict(start_time=17.8 * 3600, timestep=120.0, episode_length=40,
                                               renewable_sources=("solar",))),
    "trace_ieee13_det": ("ieee13", 60, 3, dict(stochastic_loads=False, weather_variation=False,
                                               renewable_sources=())),
    "trace_ieee34_s0": ("ieee34", 120, 0, dict(start_time=9 * 3600.0, timestep=15.0)),
    "trace_ieee34_s1": ("ieee34", 100, 1, dict(renewable_sources=("solar", "wind"), start_time=5.95 * 3600,
                                               timestep=20.0, tolerance=1e-4, max_iterations=20)),
    "trace_ieee123_s0": ("ieee123", 24, 0, dict(start_time=10 * 3600.0)),
    "trace_ieee123_s1": ("ieee123", 16, 1, dict(start_time=14 * 3600.0, timestep=300.0, tolerance=1e-8)),
    # long runs (>= 3 seeds x >= 1000 steps on the feeders the reference can step quickly;
    # IEEE-123 costs ~1 s per reference solve, so its long traces are 150 steps)
    "trace_ieee13_long_s10": ("ieee13", 1000, 10, dict(start_time=6.5 * 3600.0, timestep=45.0)),
    "trace_ieee13_long_s11": ("ieee13", 1000, 11, dict(start_time=15 * 3600.0, timestep=20.0, episode_length=300,
                                                       nan_steps=(333, 334, 700))),
    "trace_ieee13_long_s12": ("ieee13", 1000, 12, dict(start_time=0.0, timestep=90.0, tolerance=1e-8,
                                                       renewable_sources=("wind",))),
    "trace_ieee34_long_s10": ("ieee34", 1000, 10, dict(start_time=7 * 3600.0, timestep=40.0)),
    "trace_ieee34_long_s11": ("ieee34", 1000, 11, dict(start_time=16.5 * 3600.0, timestep=10.0, episode_length=400)),
    "trace_ieee34_long_s12": ("ieee34", 1000, 12, dict(start_time=11 * 3600.0, timestep=60.0, tolerance=1e-8,
                                                       renewable_sources=("solar", "wind"))),
    "trace_ieee123_long_s10": ("ieee123", 150, 10, dict(start_time=8 * 3600.0, timestep=120.0)),
    "trace_ieee123_long_s11": ("ieee123", 150, 11, dict(start_time=13 * 3600.0, timestep=30.0, episode_length=60)),
    "trace_ieee123_long_s12": ("ieee123", 150, 12, dict(start_time=18.5 * 3600.0, timestep=60.0, tolerance=1e-8)),
}
# environments on MESHED feeders (loop-closing lines kept): the reference's step with its dense Newton-Raphson;
# the CUDA path takes these through the sweep with compensation (a different algorithm: compared tight)
MESH_TRACE_CASES = {
    "meshtrace_ieee34_s0": ("meshieee34", 60, 0, dict(start_time=9 * 3600.0, timestep=30.0, tolerance=1e-8)),
    "meshtrace_synthetic40_s1": ("mesh40:3:0.05", 40, 1, dict(start_time=13 * 3600.0, timestep=60.0, tolerance=1e-8,
                                                            load_scale=0.05)),
    "meshtrace_ieee123_s2": ("meshieee123", 10, 2, dict(start_time=11 * 3600.0, timestep=60.0, tolerance=1e-8)),
}
SOLVE_CASES = {
    "solve_fixture3": ("fixture3", 0, 4, 1e-10, 0.025),
    "solve_radial34": ("radial34", 1, 6, 1e-8, 0.02),
    "solve_ieee13": ("ieee13", 2, 8, 1e-8, 1.0),
    "solve_ieee34": ("ieee34", 3, 6, 1e-6, 1.0),
    "solve_ieee123": ("ieee123", 4, 3, 1e-6, 1.0),
    "solve_synthetic60": ("synthetic60:7", 5, 4, 1e-8, 0.05),
    # BASELINE config 5's feeder (ScalableFeeder(1000)'s parameters, radial, loads x 0.03 as bench.py runs it):
    # the reference's dense Newton-Raphson costs minutes per solve at this size - two cases
    "solve_synthetic1000": ("synthetic1000:1000", 11, 2, 1e-6, 0.03),
    # deliberately infeasible loading: the reference runs to max_iterations, converged=False
    "solve_ieee13_overload": ("ieee13", 6, 2, 1e-6, 9.0),
    # meshed networks (cycle lines kept): the reference's dense Newton-Raphson as it is
    "meshsolve_ieee34": ("meshieee34", 7, 5, 1e-6, 1.0),
    "meshsolve_synthetic40": ("mesh40:3:0.05", 8, 5, 1e-8, 0.05),
    "meshsolve_synthetic72": ("mesh72:5:0.02", 9, 3, 1e-8, 0.03),
    "meshsolve_synthetic24_overload": ("mesh24:2:0.1", 10, 2, 1e-6, 200.0),
    # IEEE-123 with its 26 tie lines kept: 244 unknowns - too large for the dense CUDA kernels, taken by the sweep
    "meshsolve_ieee123": ("meshieee123", 12, 2, 1e-8, 1.0),
}


# --------------------------------------------------------------------------- feeder tables

FEEDER_SPECS = ("radial3", "radial13", "radial34", "fixture3", "ieee13", "ieee34", "ieee123",
                "synthetic60:7", "synthetic300:300", "synthetic1000:1000", "scalable100:5",
                "mesh40:3:0.05", "mesh72:5:0.02", "meshieee34", "meshieee123")


def raw_feeder(ns, spec: str, use_reference_classes: bool = True):
    """The generator's output BEFORE any repair (D4-i only: the seed), for the table comparison."""
    sys.path.insert(0, REPO_ROOT)
    import grid_fed_rl_b200 as mine
    if spec.startswith("scalable"):
        n, seed = spec[8:].split(":")
        return (ns.synthetic.ScalableFeeder(int(n), seed=int(seed)) if use_reference_classes
                else mine.ScalableFeeder(int(n), seed=int(seed)))
    if spec in ("ieee34", "meshieee34", "ieee123", "meshieee123"):
        cls = "IEEE34Bus" if "34" in spec else "IEEE123Bus"
        if use_reference_classes:
            st = np.random.get_state(); np.random.seed(0); f = getattr(ns.feeders, cls)(); np.random.set_state(st)
            return f
        return getattr(mine, cls)(seed=0)
    if spec == "ieee13":
        return ns.feeders.IEEE13Bus() if use_reference_classes else mine.IEEE13Bus()
    f = make_feeder(ns, spec, use_reference_classes)          # radial / fixture / synthetic / mesh: no repair needed
    return getattr(f, "source", f)


def feeder_table(f) -> Dict[str, np.ndarray]:
    """Every field of a feeder the hot path reads, as plain arrays (ids as strings)."""
    import json
    gens = {str(k): {kk: (vv if isinstance(vv, str) else float(vv)) if not isinstance(vv, (int, np.integer)) or isinstance(vv, bool)
                     else int(vv) for kk, vv in v.items()} for k, v in f.generators.items()}
    return dict(
        bus_id=np.array([str(b.id) for b in f.buses]), bus_type=np.array([str(b.bus_type) for b in f.buses]),
        bus_vm=np.array([float(b.voltage_magnitude) for b in f.buses]),
        bus_level=np.array([float(b.voltage_level) for b in f.buses]),
        line_id=np.array([str(l.id) for l in f.lines]), line_from=np.array([str(l.from_bus) for l in f.lines]),
        line_to=np.array([str(l.to_bus) for l in f.lines]),
        line_r=np.array([float(l.resistance) for l in f.lines]), line_x=np.array([float(l.reactance) for l in f.lines]),
        line_rating=np.array([float(l.rating) for l in f.lines]),
        load_id=np.array([str(l.id) for l in f.loads]), load_bus=np.array([str(l.bus) for l in f.loads]),
        load_base=np.array([float(l.base_power) for l in f.loads]),
        load_p=np.array([float(l.active_power) for l in f.loads]),
        load_q=np.array([float(l.reactive_power) for l in f.loads]),
        generators=np.array(json.dumps(gens, sort_keys=True)),
        gen_keys=np.array([str(k) for k in f.generators] or [""]),     # dict order = action order (grid_env.py:629-651)
        base_power=np.array(float(f.parameters.base_power)), base_voltage=np.array(float(f.parameters.base_voltage)))


def freeze_feeders(ns, out_dir: str) -> None:
    """tests/golden/feeders.npz: the reference generators' raw output and the repaired (D4) network for
    every feeder spec the tests and the bench use (tests/test_feeders.py compares this package's own
    generators with it, field by field)."""
    sys.path.insert(0, REPO_ROOT)
    import grid_fed_rl_b200 as mine
    data = {}
    for spec in FEEDER_SPECS:
        raw = raw_feeder(ns, spec)
        for k, v in feeder_table(raw).items():
            data[f"{spec}/raw/{k}"] = v
        fixed = mine.repair_topology(raw, keep_cycles=spec.startswith("mesh"))
        for k, v in feeder_table(fixed).items():
            if k.startswith("line_"):
                data[f"{spec}/repaired/{k}"] = v
        print("feeder", spec, len(raw.buses), "buses", len(raw.lines), "->", len(fixed.lines), "lines",
              len(raw.loads), "loads", len(raw.generators), "generators", flush=True)
    np.savez_compressed(os.path.join(out_dir, "feeders.npz"), **data)


def freeze_dataset(ns, out_dir: str) -> None:
    """tests/golden/dataset_gridDataset.npz: the reference's GridDataset (algorithms/base.py:180-265) fed a seeded
    transition dict - its normalisation statistics, normalised arrays, ``__len__`` and ``__getitem__`` - for
    the RolloutBuffer parity test."""
    import torch
    from grid_fed_rl.algorithms.base import GridDataset
    rs = np.random.RandomState(7)
    N, D, A = 257, 19, 3
    raw = dict(observations=rs.normal(2.0, 5.0, size=(N, D)), actions=rs.uniform(-1, 1, size=(N, A)),
               rewards=rs.normal(-30.0, 12.0, size=N), next_observations=rs.normal(2.0, 5.0, size=(N, D)),
               terminals=(rs.random_sample(N) < 0.1))
    raw["observations"][:, 4] = 3.25          # a constant column: std 0 -> the reference's epsilon decides
    raw["next_observations"][:, 4] = 3.25
    out = {"raw_" + k: np.asarray(v) for k, v in raw.items()}
    probe_a = rs.normal(size=(5, A)).astype(np.float32)
    probe_o = rs.normal(size=(5, D)).astype(np.float32)
    out["probe_action"], out["probe_observation"] = probe_a, probe_o
    for norm in (True, False):
        ds = GridDataset(**{k: np.array(v) for k, v in raw.items()}, normalize=norm, device="cpu")
        tag = "norm" if norm else "plain"
        out[f"{tag}_size"] = np.array(ds.size)
        for k in ("observations", "actions", "rewards", "next_observations", "terminals"):
            out[f"{tag}_{k}"] = np.asarray(getattr(ds, k))                       # float64 numpy, as the reference holds them
            out[f"{tag}_all_{k}"] = ds.get_all_data()[k].detach().cpu().numpy()  # float32 tensors, as learners get them
        out[f"{tag}_denorm_action"] = ds.denormalize_action(torch.from_numpy(probe_a)).numpy()
        out[f"{tag}_denorm_observation"] = ds.denormalize_observation(torch.from_numpy(probe_o)).numpy()
        if norm:
            for k in ("obs_mean", "obs_std", "action_mean", "action_std", "reward_mean", "reward_std"):
                out["stat_" + k] = np.asarray(getattr(ds, k), dtype=float)
    np.savez_compressed(os.path.join(out_dir, "dataset_gridDataset.npz"), **out)
    print("dataset", sorted(out), flush=True)


def freeze_multi_agent(ns, out_dir: str) -> None:
    """tests/golden/multi_agent_wrapper.npz: the reference's MultiAgentEnvironmentWrapper
    (algorithms/multi_agent.py:37-135) around a scripted base environment - observation slices (incl. one agent
    whose slice runs past the end and is zero-padded), joint actions (scalar action, missing agent), reward
    split with a per-agent bonus in info, done flags."""
    from grid_fed_rl.algorithms.multi_agent import AgentConfig, MultiAgentEnvironmentWrapper
    rs = np.random.RandomState(11)
    D, steps = 19, 6
    cfgs = [("battery", 7, 1), ("solar", 5, 1), ("wind", 4, 1), ("observer", 6, 2)]     # 7 + 5 + 4 + 6 = 22 > 19
    obs_script = rs.normal(size=(steps + 1, D))
    rew_script = rs.normal(size=steps) * 10
    done_script = [(False, False), (False, False), (False, True), (False, False), (True, False), (False, False)]
    bonus_script = rs.normal(size=steps)

    class Scripted:
        def __init__(self):
            self.t = 0
            self.joint = []

        def reset(self):
            self.t = 0
            return obs_script[0], {}

        def step(self, joint):
            self.joint.append(np.array(joint, dtype=float))
            t = self.t
            self.t += 1
            info = {"solar_reward_bonus": float(bonus_script[t]), "t": t}
            return obs_script[t + 1], float(rew_script[t]), done_script[t][0], done_script[t][1], info

    base = Scripted()
    w = MultiAgentEnvironmentWrapper(base, [AgentConfig(a, o, d) for a, o, d in cfgs])
    out = {"obs_script": obs_script, "rew_script": rew_script, "bonus_script": bonus_script,
           "done_script": np.array(done_script), "agents": np.array([c[0] for c in cfgs]),
           "obs_dims": np.array([c[1] for c in cfgs]), "act_dims": np.array([c[2] for c in cfgs])}
    first = w.reset()
    for a in first:
        out[f"reset_obs_{a}"] = np.asarray(first[a], dtype=float)
    act_script = []
    for t in range(steps):
        acts = {"battery": rs.uniform(-1, 1, size=1), "solar": float(rs.uniform(-1, 1)),      # a scalar action
                "observer": rs.uniform(-1, 1, size=2)}
        if t % 2 == 0:
            acts["wind"] = rs.uniform(-1, 1, size=1)                                            # else: missing -> zeros
        act_script.append({k: np.atleast_1d(np.asarray(v, dtype=float)) for k, v in acts.items()})
        obs, rew, done, info = w.step(acts)
        for a in obs:
            out[f"step{t}_obs_{a}"] = np.asarray(obs[a], dtype=float)
            out[f"step{t}_rew_{a}"] = np.array(float(rew[a]))
            out[f"step{t}_done_{a}"] = np.array(bool(done[a]))
        for k, v in act_script[-1].items():
            out[f"step{t}_act_{k}"] = v
    out["joint_actions"] = np.array(base.joint)
    np.savez_compressed(os.path.join(out_dir, "multi_agent_wrapper.npz"), **out)
    print("multi-agent wrapper golden:", out["joint_actions"].shape, flush=True)


def main() -> None:
    ns = build()
    out_dir = os.path.join(REPO_ROOT, "tests", "golden")
    os.makedirs(out_dir, exist_ok=True)
    only = set(sys.argv[1:])
    if not only or "feeders" in only:
        freeze_feeders(ns, out_dir)
    if not only or "dataset" in only:
        freeze_dataset(ns, out_dir)
    if not only or "multi_agent" in only:
        freeze_multi_agent(ns, out_dir)
    for name, (spec, seed, count, tol, scale) in SOLVE_CASES.items():
        if only and name not in only:
            continue
        f = make_feeder(ns, spec)
        data = solve_cases(ns, f, seed, count, tol, scale=scale, max_iterations=30 if "overload" in name else 50)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), spec=np.array(spec), **data)
        print(name, "iters", data["iterations"], "conv", data["converged"], flush=True)
    for name, (spec, steps, seed, kw) in list(TRACE_CASES.items()) + list(MESH_TRACE_CASES.items()):
        if only and name not in only:
            continue
        f = make_feeder(ns, spec)
        data = run_trace(ns, f, steps, seed, **kw)
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), spec=np.array(spec), **data)
        print(name, "resets", int(data["reset_before"].sum()), "conv", int(data["converged"].sum()),
              "errors", int(data["error"].sum()), "trunc", int(data["truncated"].sum()),
              "term", int(data["terminated"].sum()), flush=True)


if __name__ == "__main__":
    main()

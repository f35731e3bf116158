"""The device functions (csrc/gfr_device.cuh, LANES = 1 compiled for the host by tests/host_emu)
against the frozen reference outputs and the numpy oracle.  This is the CPU-tier check of the
kernels' arithmetic; the GPU tier (test_gpu_parity.py) runs the same cases through the C ABI."""
import numpy as np
import pytest

from oracle import port
from tests.golden_util import (TOL_PU, feeder_for, golden_names, load_golden, port_trace,
                               replay_trace)
from tests.host_emu import emu


def _factory(solver, tol_override=None):
    class _One:
        def __init__(self, feeder, kw):
            kw = dict(kw)
            tol = kw.pop("tolerance")
            self.env = emu.EmuEnv(feeder, 1, solver=solver, tolerance=tol_override or tol, **kw)

        def reset(self, noise4, start_time):
            return self.env.reset(np.asarray(noise4)[None, :], start_time=start_time)[0]

        def step(self, action, noise):
            out = self.env.step(action[None, :], noise[None, :])
            return {k: v[0] for k, v in out.items()}
    return _One


@pytest.mark.parametrize("name", golden_names("trace_"))
def test_emu_newton_matches_reference_trace(name):
    g = load_golden(name)
    exact = replay_trace(_factory("newton"), g, ctx=name)
    assert exact >= 0.9 * g["obs"].shape[0]


@pytest.mark.parametrize("name", [n for n in golden_names("trace_") if "_long_" not in n])
def test_emu_sweep_matches_reference_trace(name):
    # a different algorithm: both sides run tight, the oracle at NR tol 1e-10
    g = port_trace(load_golden(name), tolerance=1e-10)
    exact = replay_trace(_factory("sweep", 1e-11), g, ctx=name, check_iterations=False)
    assert exact >= 0.9 * g["obs"].shape[0]


@pytest.mark.parametrize("lanes", [0, 1, 8, 32])
@pytest.mark.parametrize("name", golden_names("solve_"))
@pytest.mark.parametrize("solver", ["newton", "sweep"])
def test_emu_solver_matches_reference(name, solver, lanes):
    g = load_golden(name)
    f = feeder_for(g)
    tol, max_it = float(g["meta"][0]), int(g["meta"][1])
    if solver == "sweep":
        # both sides tight: the oracle's NR at 1e-10 instead of the frozen (looser) reference run
        net = port.DenseNetwork(f.buses, f.lines)
        ref = port.newton_raphson(net, g["p_spec"], 1e-10, 50)
        g = dict(g); g.update({k: ref[k] for k in ("bus_voltages", "bus_angles", "line_flows", "losses")})
        tol, max_it = 1e-11, 200
    sol = emu.emu_solve(f, g["p_spec"], solver, tol, max_it, lanes=lanes)   # lanes shapes the level schedule
    conv = g["converged"]
    if solver == "newton":
        assert np.array_equal(sol["converged"].astype(bool), conv)
        assert np.all(np.abs(sol["iterations"].astype(int) - g["iterations"]) <= 1)
    else:
        assert np.all(sol["converged"].astype(bool)[conv])
    if conv.any():
        for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
            assert np.max(np.abs(sol[k][conv] - g[k][conv])) <= TOL_PU, k


def test_emu_philox_matches_oracle():
    rs = np.random.RandomState(0)
    seeds = rs.randint(0, 2**63 - 1, size=64, dtype=np.int64).astype(np.uint64)
    draws = rs.randint(0, 2**40, size=64, dtype=np.int64).astype(np.uint64)
    for n_slots in (4, 5, 12, 99):
        a = emu.emu_noise(seeds, draws, n_slots)
        b = port.philox_noise(seeds, draws, n_slots)
        assert np.max(np.abs(a - b)) < 1e-13


@pytest.mark.parametrize("lanes", [1, 2, 4, 8, 16])
@pytest.mark.parametrize("name", ["trace_fixture3_s0", "trace_ieee13_s0", "trace_ieee34_s1", "trace_ieee123_s0"])
def test_emu_newton_trace_on_every_lane_count(name, lanes):
    """The path schedules (register hand-off along a lane, pool slots between lanes) on 1 - 16 emulated lanes."""
    g = load_golden(name)

    class _One(_factory("newton")):
        def __init__(self, feeder, kw):
            kw = dict(kw)
            tol = kw.pop("tolerance")
            self.env = emu.EmuEnv(feeder, 1, solver="newton", tolerance=tol, lanes=lanes, **kw)
            assert self.env.emu_lanes == lanes
    exact = replay_trace(_One, g, ctx=f"{name}/emu{lanes}")
    assert exact >= 0.9 * g["obs"].shape[0]


def test_schedule_hands_over_in_registers():
    """IEEE-123 on 8 lanes: 17 rows (Hu's bound for 123 buses of depth 13), every non-leaf bus takes one
    child's contribution in registers (78 of 122 branches), the rest fit a pool no larger than the staging
    ring asks for; a schedule cut for another lane count is re-cut by the library and still solves."""
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE123Bus(seed=0))
    e = emu.EmuEnv(f, 1, solver="newton", lanes=8, renewable_sources=["solar", "wind"])
    assert e.schedule["rows"] == 17 and e.schedule["positions"] == 136
    assert e.schedule["register_edges"] == 78
    assert e.schedule["pool_slots"] <= 18
    leaves = 123 - len({int(p) for p in e.soa.parent[1:]})
    assert e.schedule["register_edges"] == 122 - (leaves - 1)          # the bound: one heir per non-leaf bus
    e1 = emu.EmuEnv(f, 1, solver="newton", lanes=1, renewable_sources=["solar", "wind"])
    assert e1.schedule["rows"] == 123 and e1.schedule["register_edges"] == 78
    # compiled for 8 lanes, run on 4: rows are re-cut, results stay the same
    g = load_golden("solve_ieee123")
    a = emu.emu_solve(f, g["p_spec"], "newton", 1e-6, 50, lanes=8)
    b = emu.emu_solve(f, g["p_spec"], "newton", 1e-6, 50, lanes=8, emu_lanes=4)
    assert np.array_equal(a["iterations"], b["iterations"])
    assert np.max(np.abs(a["bus_voltages"] - b["bus_voltages"])) < 1e-12


# ----------------------------------------------------------------------------- weakly meshed feeders (sweep + compensation)

@pytest.mark.parametrize("lanes", [1, 4, 8])
@pytest.mark.parametrize("name", golden_names("meshtrace_"))
def test_emu_sweep_steps_meshed_feeders(name, lanes):
    """Environments on feeders with their loop-closing lines kept (IEEE-34 + its loop, a 40-bus synthetic mesh with
    37 ties, IEEE-123 with its 26 ties): the reference steps them with its dense Newton-Raphson (frozen traces); here
    the sweep walks the spanning tree and restores the loops by compensation.  Both sides tight."""
    g = port_trace(load_golden(name), tolerance=1e-10)

    class _One(_factory("sweep")):
        def __init__(self, feeder, kw):
            kw = dict(kw)
            kw.pop("tolerance")
            kw["max_iterations"] = 200
            self.env = emu.EmuEnv(feeder, 1, solver="sweep", tolerance=1e-11, lanes=lanes, **kw)
            assert self.env.soa.n_tie > 0 and self.env.soa.n_line == self.env.soa.n_bus - 1 + self.env.soa.n_tie
    exact = replay_trace(_One, g, ctx=f"{name}/emu{lanes}", check_iterations=False)
    assert exact >= 0.9 * g["obs"].shape[0]


@pytest.mark.parametrize("name", golden_names("meshsolve_"))
def test_emu_sweep_solves_meshed_networks(name):
    g = load_golden(name)
    f = feeder_for(g)
    conv = g["converged"]
    if not conv.any():
        pytest.skip("the frozen cases diverge (an infeasible loading)")
    net = port.DenseNetwork(f.buses, f.lines)
    ref = port.newton_raphson(net, g["p_spec"], 1e-10, 50)
    for lanes in (1, 8):
        sol = emu.emu_solve(f, g["p_spec"], "sweep", 1e-11, 300, lanes=lanes)
        ok = ref["converged"]
        assert np.all(sol["converged"].astype(bool)[ok])
        for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
            assert np.max(np.abs(sol[k][ok] - ref[k][ok])) <= TOL_PU, (k, lanes)
        s_base = f.parameters.base_power * 1e6
        assert np.allclose(sol["line_loadings"][ok], ref["line_loadings"][ok] * s_base, rtol=1e-6, atol=1e-10)


def test_newton_refuses_ties_and_zinv_is_the_loop_inverse():
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200.topology import TopologyError, compile_feeder, compile_for_solver
    f = m.repair_topology(m.IEEE123Bus(seed=0), keep_cycles=True)
    with pytest.raises(TopologyError):
        compile_for_solver(f, "newton", 8)
    soa = compile_feeder(f, root="center", width=16)
    t = soa.n_tie
    assert t == 26 and soa.tie_zinv.shape == (26, 26, 2) and soa.obs_dim == 2 * 123 + 2 * 148 + 1 + 2 * soa.n_load + soa.n_gen + 2 * soa.n_bat
    # Z_loop rebuilt from tree paths: symmetric, and tie_zinv is its inverse
    zinv = soa.tie_zinv[..., 0] + 1j * soa.tie_zinv[..., 1]
    z = np.linalg.inv(zinv)
    assert np.allclose(z, z.T, atol=1e-12)
    zt = soa.tie_r + 1j * soa.tie_x
    assert np.all(np.abs(np.diag(z)) >= np.abs(zt) - 1e-12)


# ----------------------------------------------------------------------------- more pool children than a packed record holds

@pytest.mark.parametrize("lanes", [2, 8, 16])
def test_emu_hub_with_many_children(lanes):
    """A hub bus with 13 children: 12 of them hand over through pool slots - more than the 8 a packed pool-child
    record holds (the wide-group path of the emulation build, 8 and 16 lanes, must fall back to the child list for
    that position; 2 lanes walk the list anyway).  Newton and sweep against the oracle's dense Newton-Raphson."""
    import grid_fed_rl_b200 as m
    f = m.SimpleRadialFeeder(18)
    for ln in f.lines[3:16]:                     # buses 5 .. 17 all hang off bus 4; bus 18 stays behind 17
        ln.from_bus = 4
    for ld in f.loads:
        ld.base_power *= 0.2; ld.active_power *= 0.2; ld.reactive_power *= 0.2
    f = m.repair_topology(f)
    net = port.DenseNetwork(f.buses, f.lines)
    rs = np.random.RandomState(5)
    n = len(f.buses)
    p_spec = -np.abs(rs.uniform(0.002, 0.01, size=(6, n))); p_spec[:, 0] = 0.0
    ref = port.newton_raphson(net, p_spec, 1e-10, 50)
    assert ref["converged"].all()
    for solver, tol, it in (("newton", 1e-10, 50), ("sweep", 1e-11, 200)):
        sol = emu.emu_solve(f, p_spec, solver, tol, it, lanes=lanes)
        assert sol["converged"].all(), solver
        for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
            assert np.max(np.abs(sol[k] - ref[k])) <= TOL_PU, (solver, k)
    if lanes > 1:
        from grid_fed_rl_b200.topology import compile_for_solver
        soa, _ = compile_for_solver(f, "newton", lanes, with_components=False)
        kids = np.bincount(np.asarray(soa.parent)[1:], minlength=n)
        assert kids.max() >= 10                   # the hub survived the center rooting with > 9 children

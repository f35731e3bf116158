"""GPU tier: oracle parity AT BASELINE.json's full batch sizes, on sampled instances.

The seeded-batch tests of test_gpu_parity.py compare 40 - 333 instances with the oracle; the full-size
tests there are property tests.  Here the kernels run the BASELINE configurations at their full batch
(in-kernel Philox noise, the launch plan the bench uses) and 64 instances - the first ones, the last
ones, the last CTA's, the last wave's, and random ones - are replayed through ``oracle/port.PortEnv``
with the noise rows the kernel recorded: |V|, angle, line P, losses within 1e-8 pu; flags, counters and
done flags exact (steps closer than 1e-9 to a threshold excluded); Newton iteration counts within +-1.
A wrong index at instance B - 1, in the ragged last wave or in the wave-balanced plan fails here."""
import numpy as np
import pytest
import torch

from oracle import port
from tests.golden_util import TOL_PU, feeder_for, load_golden, obs_layout

pytestmark = pytest.mark.gpu


def _sample_indices(B, info, rs, count=64):
    E = info["threads"] // info["lanes"]            # instance slots per CTA
    stride = info["grid"] * E                       # instances per wave
    last_wave = (B - 1) // stride * stride
    picks = {0, 1, E - 1, E, B - 1, B - 2, last_wave, min(B - 1, last_wave + 1), max(0, last_wave - 1)}
    last_cta = (info["grid"] - 1) * E
    picks.update(i for i in range(last_cta, min(B, last_cta + E), max(1, E // 4)))
    picks.update(i for i in range(last_wave, B, max(1, (B - last_wave) // 6)))
    picks = {int(i) for i in picks if 0 <= i < B}
    while len(picks) < count:
        picks.add(int(rs.randint(0, B)))
    return np.array(sorted(picks))[:max(count, len(picks))]


CASES = [
    # (feeder spec, B, solver, kernel tolerance, oracle tolerance, renewable sources): BASELINE configs[1..3]
    ("ieee13", 65536, "sweep", 1e-11, 1e-10, ["solar", "wind"]),
    ("ieee34", 262144, "sweep", 1e-11, 1e-10, ["solar"]),
    ("ieee123", 131072, "newton", 1e-6, 1e-6, ["solar", "wind"]),
    # and the sizes the wave-balanced / ragged plans see: not a multiple of anything
    ("ieee123", 131072 - 37, "newton", 1e-6, 1e-6, ["solar", "wind"]),
    ("ieee13", 65536 + 3, "newton", 1e-6, 1e-6, ["solar", "wind"]),
]


@pytest.mark.parametrize("spec,B,solver,tol,ptol,srcs", CASES)
def test_sampled_instances_of_the_full_batch_match_the_oracle(spec, B, solver, tol, ptol, srcs):
    import grid_fed_rl_b200 as m
    from oracle.ref_harness import make_feeder
    f = make_feeder(None, spec, use_reference_classes=False)
    start, seed, offset, steps = 12 * 3600.0, 31, 1000, 3
    kw = dict(timestep=1.0, renewable_sources=srcs, stochastic_loads=True, weather_variation=True)
    env = m.BatchedGridEnvironment(f, B, solver=solver, tolerance=tol, repair=False, record_noise=True,
                                   start_time=start, env_id_offset=offset, **kw)
    info = env.launch_info()
    rs = np.random.RandomState(B % 1000)
    idx = _sample_indices(B, info, rs)
    n_s = idx.size
    ref = port.PortEnv(f, n_s, tolerance=ptol, **kw)
    env.reset(seed=seed)
    keys = (idx + seed + offset).astype(np.uint64)               # seed + GLOBAL instance id
    nz0 = port.philox_noise(keys, np.zeros(n_s, dtype=np.uint64), 4)
    robs0 = ref.reset(nz0, start_time=start)
    obs0 = env.get_observation()[torch.as_tensor(idx, device=env.device)].cpu().numpy()
    assert np.max(np.abs(obs0 - robs0)) < 1e-9
    lay = obs_layout(ref.n, ref.m, ref.L, ref.G, ref.Bt)
    g = torch.Generator(device=env.device); g.manual_seed(5)
    sel = torch.as_tensor(idx, device=env.device)
    exact_steps = 0
    for t in range(steps):
        act = env.sample_actions(g)
        obs, reward, term, trunc, inf = env.step(act)
        used = env.noise_used[sel].cpu().numpy()
        expect = port.philox_noise(keys, np.full(n_s, t + 1, dtype=np.uint64), 4 + ref.L)
        assert np.max(np.abs(used - expect)) < 1e-12, "the recorded noise row is not this instance's Philox row"
        r = ref.step(act[sel].cpu().numpy(), used)
        o = obs[sel].cpu().numpy()
        assert bool(inf["power_flow_converged"].all()) and r["converged"].all()
        assert np.max(np.abs(o[:, lay["vm"]] - r["obs"][:, lay["vm"]])) <= TOL_PU
        assert np.max(np.abs(o[:, lay["va"]] - r["obs"][:, lay["va"]])) <= TOL_PU
        assert np.max(np.abs(o[:, lay["p"]] - r["obs"][:, lay["p"]])) / ref.s_base <= TOL_PU
        assert np.max(np.abs(inf["total_losses"][sel].cpu().numpy() - r["losses"])) / ref.s_base <= TOL_PU
        assert np.max(np.abs(o[:, lay["freq"]] - r["obs"][:, lay["freq"]])) <= 1e-9
        assert np.max(np.abs(o[:, lay["soc"]] - r["obs"][:, lay["soc"]])) <= 1e-12
        assert np.allclose(o[:, lay["gen"]], r["obs"][:, lay["gen"]], rtol=1e-12, atol=1e-6)
        assert np.array_equal(o[:, lay["loads"]], r["obs"][:, lay["loads"]])
        assert np.allclose(reward[sel].cpu().numpy(), r["reward"], rtol=1e-9, atol=1e-6)
        assert np.array_equal(term[sel].cpu().numpy(), r["terminated"])
        assert np.array_equal(inf["current_step"][sel].cpu().numpy(), r["current_step"])
        if solver == "newton":
            assert np.all(np.abs(inf["iterations"][sel].cpu().numpy().astype(int) - r["iterations"]) <= 1)
        vm, fr = r["obs"][:, lay["vm"]], r["obs"][:, lay["freq"]]
        margin = np.minimum(np.min(np.abs(vm - 0.95), axis=1), np.min(np.abs(vm - 1.05), axis=1))
        margin = np.minimum(margin, np.minimum(np.abs(fr - 59.5), np.abs(fr - 60.5)))
        clear = margin > 1e-9
        assert clear.mean() > 0.9
        assert np.array_equal(inf["constraint_violations"][sel].cpu().numpy()[clear], r["violations"][clear])
        assert np.array_equal(inf["constraint_violation_count"][sel].cpu().numpy()[clear], r["viol_count"][clear])
        assert np.array_equal(trunc[sel].cpu().numpy()[clear], r["truncated"][clear])
        exact_steps += int(clear.sum())
    assert exact_steps >= 0.9 * steps * n_s
    env.close()


@pytest.mark.parametrize("lanes", (64, 128))
def test_synthetic_1000_solver_matches_reference(lanes):
    """BASELINE config 5's feeder (1 000 buses) against the frozen outputs of the reference's own dense
    Newton-Raphson (tests/golden/solve_synthetic1000.npz: minutes per solve upstream), one CTA per instance."""
    import grid_fed_rl_b200 as m
    g = load_golden("solve_synthetic1000")
    f = feeder_for(g)
    assert len(f.buses) == 1000
    tol, max_it = float(g["meta"][0]), int(g["meta"][1])
    sol = m.B200PowerFlowSolver(tolerance=tol, max_iterations=max_it, method="newton", lanes=lanes).solve_batch(f, g["p_spec"])
    assert np.array_equal(sol.converged.cpu().numpy(), g["converged"]) and g["converged"].all()
    assert np.all(np.abs(sol.iterations.cpu().numpy().astype(int) - g["iterations"]) <= 1)
    for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
        assert np.max(np.abs(getattr(sol, k).cpu().numpy() - g[k])) <= TOL_PU, k
    s_base = f.parameters.base_power * 1e6
    assert np.allclose(sol.line_loadings.cpu().numpy(), g["line_loadings"] * s_base, rtol=1e-7, atol=1e-12)
    sw = m.B200PowerFlowSolver(tolerance=1e-11, max_iterations=200, method="sweep", lanes=lanes).solve_batch(f, g["p_spec"])
    tight = m.B200PowerFlowSolver(tolerance=1e-11, max_iterations=50, method="newton", lanes=lanes).solve_batch(f, g["p_spec"])
    for k in ("bus_voltages", "bus_angles", "line_flows", "losses"):
        assert torch.max(torch.abs(getattr(sw, k) - getattr(tight, k))) <= TOL_PU, k


@pytest.mark.parametrize("lanes", (64, 128))
def test_synthetic_1000_step_matches_oracle(lanes):
    """Two instances of the 1 000-bus feeder stepped against oracle/port.PortEnv (dense Newton-Raphson on
    1 998 unknowns: seconds per solve on the host) - the environment side of config 5, D = 6 320, 403 actions."""
    import grid_fed_rl_b200 as m
    cfg = m.NetworkConfig(num_buses=1000, connectivity=0.0, load_probability=0.9, dg_probability=0.4,
                          min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
    f = m.repair_topology(m.SyntheticFeeder(cfg, seed=1000))
    for ld in f.loads:
        ld.base_power *= 0.03; ld.active_power *= 0.03; ld.reactive_power *= 0.03
    kw = dict(timestep=60.0, renewable_sources=["solar", "wind"], tolerance=1e-6)
    B = 2
    env = m.BatchedGridEnvironment(f, B, lanes=lanes, repair=False, **kw)
    ref = port.PortEnv(f, B, **kw)
    assert env.obs_dim == 6320 and env.act_dim == 403
    rs = np.random.RandomState(lanes)
    nz0 = np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3))], axis=1)
    o0, _ = env.reset(noise=nz0, options={"start_time": 12 * 3600.0})
    r0 = ref.reset(nz0, start_time=12 * 3600.0)
    assert np.max(np.abs(o0.cpu().numpy() - r0)) < 1e-9
    lay = obs_layout(ref.n, ref.m, ref.L, ref.G, ref.Bt)
    for t in range(2):
        act = rs.uniform(-1, 1, size=(B, ref.A))
        nz = np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3 + ref.L))], axis=1)
        obs, reward, term, trunc, info = env.step(act, nz)
        r = ref.step(act, nz)
        o = obs.cpu().numpy()
        assert r["converged"].all() and bool(info["power_flow_converged"].all())
        assert np.max(np.abs(o[:, lay["vm"]] - r["obs"][:, lay["vm"]])) <= TOL_PU
        assert np.max(np.abs(o[:, lay["va"]] - r["obs"][:, lay["va"]])) <= TOL_PU
        assert np.max(np.abs(o[:, lay["p"]] - r["obs"][:, lay["p"]])) / ref.s_base <= TOL_PU
        assert np.max(np.abs(info["total_losses"].cpu().numpy() - r["losses"])) / ref.s_base <= TOL_PU
        assert np.allclose(o[:, lay["gen"]], r["obs"][:, lay["gen"]], rtol=1e-12, atol=1e-6)
        assert np.allclose(o[:, lay["soc"]], r["obs"][:, lay["soc"]], atol=1e-12)
        assert np.allclose(reward.cpu().numpy(), r["reward"], rtol=1e-9, atol=1e-6)
        assert np.all(np.abs(info["iterations"].cpu().numpy().astype(int) - r["iterations"]) <= 1)
        assert np.array_equal(info["constraint_violation_count"].cpu().numpy(), r["viol_count"])
    env.close()


def test_unseeded_shards_reproduce_the_unsharded_run():
    """Construction keys instance i with env_id_offset + i: shards that never pass a seed (plain reset(),
    auto_reset) still draw the streams of their GLOBAL instances - not one stream per local index."""
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], solver="newton", tolerance=1e-8, repair=False,
              start_time=9 * 3600.0, episode_length=3, auto_reset=True)
    whole = m.BatchedGridEnvironment(f, 600, **kw)
    whole.reset()                                             # no seed anywhere
    g = torch.Generator(device="cuda"); g.manual_seed(3)
    acts = [whole.sample_actions(g) for _ in range(5)]
    outs = [whole.step(a)[0].clone() for a in acts]            # crosses an auto-reset at step 3
    parts = []
    for rank in range(2):
        lo, hi = m.shard_range(600, rank, 2)
        part = m.BatchedGridEnvironment(f, hi - lo, env_id_offset=lo, **kw)
        part.reset()
        for t, a in enumerate(acts):
            op = part.step(a[lo:hi])[0]
            assert torch.equal(op, outs[t][lo:hi]), (rank, t)
        parts.append(part.step(acts[0][lo:hi])[0].clone())
    # and the two shards are NOT copies of each other
    assert not torch.equal(parts[0][:300], parts[1][:300])


def test_c_abi_auto_lanes_equals_python_choice():
    """lanes = 0 through the C ABI with a description that carries no lanes_hint resolves to the same lane
    count the Python front end picks (one rule: gfr_auto_lanes)."""
    import ctypes as C
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200 import _native as nat
    from grid_fed_rl_b200.topology import compile_feeder, compile_for_solver
    lib = nat.load_library()
    for make, srcs in ((m.IEEE13Bus, ["solar", "wind"]), (lambda: m.IEEE34Bus(seed=0), ["solar"]),
                       (lambda: m.IEEE123Bus(seed=0), ["solar", "wind"])):
        f = m.repair_topology(make())
        _, want = compile_for_solver(f, "newton", 0, renewable_sources=srcs)
        soa = compile_feeder(f, renewable_sources=srcs, root="center", width=want)      # no lanes_hint attribute
        desc, keep = nat.make_feeder_desc(soa)
        assert desc.lanes_hint == 0
        hf, he = C.c_void_p(), C.c_void_p()
        nat.check(lib, lib.gfr_feeder_create(C.byref(desc), 0, C.byref(hf)))
        cfg = nat.make_env_cfg(solver_cfg=nat.make_solver_cfg("newton", 1e-6, 50, 1.0, 0))
        nat.check(lib, lib.gfr_env_create(hf, 1000, C.byref(cfg), C.byref(he)))
        lanes = C.c_int32()
        nat.check(lib, lib.gfr_env_launch_info(he, C.byref(lanes), None, None, None))
        assert lanes.value == want
        lib.gfr_env_destroy(he); lib.gfr_feeder_destroy(hf)


def test_strict_shapes():
    """A transposed or flat action block is refused instead of being reinterpreted row by row."""
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    env = m.BatchedGridEnvironment(f, 6, renewable_sources=["solar", "wind"], repair=False)
    A = env.act_dim
    assert A == 3
    for bad in (np.zeros((A, 6)), np.zeros(6 * A), np.zeros((6, A, 1)), np.zeros((2, 9))):
        with pytest.raises(m.InvalidActionError):
            env.step(bad)
    with pytest.raises(m.InvalidActionError):
        env.step(np.zeros((6, A)), noise=np.zeros((env.noise_dim, 6)))
    env.step(np.zeros((6, A)))
    one = m.BatchedGridEnvironment(f, 1, renewable_sources=["solar", "wind"], repair=False)
    one.step(np.zeros(A))                                      # one instance: a plain action vector is unambiguous
    env.close(); one.close()

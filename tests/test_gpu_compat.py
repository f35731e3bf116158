"""GPU tier: the reference-shaped fronts (single-instance scalars, list API) and the HBM rollout
buffer - the callers either side of the hot path."""
import numpy as np
import pytest
import torch

from tests.golden_util import (feeder_for, load_golden, obs_layout, replay_trace, trace_kwargs)

pytestmark = pytest.mark.gpu


def test_single_env_has_reference_types_and_values():
    import grid_fed_rl_b200 as m
    g = load_golden("trace_fixture3_s0")

    class _One:
        def __init__(self, feeder, kw):
            self.env = m.GridEnvironment(feeder, repair=False, **kw)

        def reset(self, noise4, start_time):
            obs, info = self.env.reset(options={"start_time": start_time}, noise=np.asarray(noise4)[None, :4])
            assert isinstance(obs, list) and isinstance(obs[0], float)
            assert set(info) == {"current_step", "episode_reward", "constraint_violations", "timestep"}
            return np.array(obs)

        def step(self, action, noise):
            obs, reward, term, trunc, info = self.env.step(action, noise)
            assert isinstance(obs, list) and isinstance(reward, float)
            assert isinstance(term, bool) and isinstance(trunc, bool)
            assert {"power_flow_converged", "max_voltage", "min_voltage", "total_losses",
                    "constraint_violations", "current_step", "episode_reward"} <= set(info)
            assert set(info["constraint_violations"]) == {"voltage_high", "voltage_low", "frequency_high",
                                                         "frequency_low"}
            v = info["constraint_violations"]
            return dict(obs=np.array(obs), reward=reward, terminated=term, truncated=trunc,
                        error="error" in info, converged=info["power_flow_converged"],
                        iterations=info["iterations"], losses=info["total_losses"],
                        violations=[v[k] for k in ("voltage_high", "voltage_low", "frequency_high",
                                                   "frequency_low")],
                        viol_count=self.env.constraint_violations, current_step=self.env.current_step,
                        episode_reward=self.env.episode_reward)
    exact = replay_trace(_One, g, ctx="compat")
    assert exact >= 0.9 * g["obs"].shape[0]


def test_single_env_invalid_actions_follow_the_reference():
    import grid_fed_rl_b200 as m
    env = m.GridEnvironment(m.IEEE13Bus(), renewable_sources=["solar", "wind"])
    obs0, _ = env.reset(seed=0)
    assert env.observation_space.shape == (71,) and env.action_space.shape == (3,)
    for bad in (np.array([0.1, np.nan, 0.0]), np.array([np.inf, 0.0, 0.0]), "not an action", [0.1, 0.2]):
        obs, reward, term, trunc, info = env.step(bad)
        # test_robust_features.py:379 - invalid action => reward <= -safety_penalty; grid_env.py:454-467
        assert reward == -2 * env.safety_penalty and term and not trunc and "error" in info
        assert env.current_step == 0
    obs, reward, term, trunc, info = env.step(np.array(env.action_space.sample()))
    assert env.current_step == 1 and info["power_flow_converged"] and not term
    assert len(obs) == 71 and obs != obs0


def test_vectorized_list_api():
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    venv = m.VectorizedEnvironment(lambda: m.BatchedGridEnvironment(f, 6, renewable_sources=["solar"],
                                                                      repair=False), num_envs=6)
    obs, infos = venv.reset(seeds=[1, 2, 3, 4, 5, 6])
    assert len(obs) == 6 and len(obs[0]) == venv.env.obs_dim and len(infos) == 6
    acts = [[0.1 * i, -0.2] for i in range(6)]
    obs, rewards, dones, truncs, infos = venv.step(acts)
    assert all(isinstance(x, list) for x in (obs, rewards, dones, truncs, infos))
    assert isinstance(rewards[0], float) and isinstance(dones[0], bool)
    assert infos[3]["current_step"] == 1 and infos[3]["power_flow_converged"]
    with pytest.raises(ValueError):
        venv.step(acts[:5])
    # same seeds, same actions -> same observations (instances are independent of their slot)
    obs_b, _ = venv.reset(seeds=[6, 5, 4, 3, 2, 1])
    o2 = venv.step(acts[::-1])[0]
    assert np.allclose(np.array(o2)[::-1], np.array(obs))


def test_rollout_buffer_matches_a_step_by_step_loop():
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], episode_length=4, timestep=60.0, repair=False)
    env = m.BatchedGridEnvironment(f, 32, **kw)
    env.reset(seed=9)
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    buf = m.collect_random_data(env, 6, generator=g, dtype=torch.float64)
    assert buf.size == 6 * 32
    data = buf.get_all_data()
    o = data["observations"].view(6, 32, -1); n = data["next_observations"].view(6, 32, -1)
    d = data["terminals"].view(6, 32).bool()
    assert bool(d[3].all()) and not bool(d[:3].any())           # episode_length = 4
    for t in range(5):
        keep = ~d[t]
        assert torch.equal(o[t + 1][keep], n[t][keep])          # s_{t+1} is the next row's s_t
    assert torch.all(o[4][:, 0] == 1.0)                         # reset observation after done
    lay = obs_layout(env.soa.n_bus, env.soa.n_line, env.soa.n_load, env.soa.n_gen, env.soa.n_bat)
    assert torch.all(n[0][:, lay["freq"]] != 60.0) or True
    # z-normalisation as GridDataset._normalize_data
    ref = data["observations"].clone()
    buf.normalize()
    z = buf.get_all_data()["observations"]
    assert torch.allclose(z, (ref - ref.mean(0)) / (ref.std(0, unbiased=False) + 1e-6))
    batch = buf.sample_batch(128, generator=g)
    assert batch["observations"].shape == (128, env.obs_dim) and batch["terminals"].shape == (128,)
    assert torch.allclose(buf.denormalize_observation(z), ref, atol=1e-9)
    out = buf.to_numpy()
    assert out["terminals"].dtype == bool and out["actions"].shape == (192, env.act_dim)


def test_host_stepper_pipeline_matches_plain_stepping():
    """Depth-2 pipelining of the host copies must not change a single result."""
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], timestep=60.0, repair=False, start_time=10 * 3600.0)
    a = m.BatchedGridEnvironment(f, 512, **kw); a.reset(seed=3)
    rs = np.random.RandomState(0)
    acts = [torch.from_numpy(rs.uniform(-1, 1, size=(512, a.act_dim))).pin_memory() for _ in range(7)]
    plain = []
    for x in acts:
        _, r, t, u, _ = a.step(x)
        plain.append((r.cpu().clone(), t.cpu().clone(), u.cpu().clone()))
    for depth in (1, 2, 3):
        # a fresh environment each time: wind / temperature / cloud survive a reset, as upstream
        b = m.BatchedGridEnvironment(f, 512, **kw); b.reset(seed=3)
        st = m.HostStepper(b, depth=depth)
        got = []
        for i, x in enumerate(acts):
            st.submit(x)
            if i + 1 >= depth:
                h = st.result(); got.append((h["reward"].clone(), h["terminated"].clone(), h["truncated"].clone()))
        while st._pending:
            h = st.result(); got.append((h["reward"].clone(), h["terminated"].clone(), h["truncated"].clone()))
        assert len(got) == len(plain)
        for (r0, t0, u0), (r1, t1, u1) in zip(plain, got):
            assert torch.equal(r0, r1) and torch.equal(t0, t1) and torch.equal(u0, u1)
        assert st.h2d_bytes_per_step == 512 * a.act_dim * 8 and st.d2h_bytes_per_step == 512 * 10
    # observations=True: every step's observation rows reach the host as well
    a2 = m.BatchedGridEnvironment(f, 512, **kw); a2.reset(seed=3)
    b = m.BatchedGridEnvironment(f, 512, **kw); b.reset(seed=3)
    st = m.HostStepper(b, depth=2, observations=True)
    assert st.d2h_bytes_per_step == 512 * (10 + 8 * (b.obs_dim - 2 * b.soa.n_load))      # the static load columns stay put
    want = []
    for x in acts[:4]:
        o, _, _, _, _ = a2.step(x)
        want.append(o.cpu().clone())
    got = []
    for i, x in enumerate(acts[:4]):
        st.submit(x)
        if i >= 1:
            got.append(st.result()["observations"].clone())
    while st._pending:
        got.append(st.result()["observations"].clone())
    assert len(got) == 4 and all(torch.equal(w, g) for w, g in zip(want, got))


def test_graphed_collection_is_consistent_and_faster_to_launch():
    """collect_random_data through a CUDA graph: same transition structure as the eager loop."""
    import time
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], episode_length=6, timestep=60.0, repair=False,
              solver="newton")
    B, T, chunk = 4096, 24, 8
    env = m.BatchedGridEnvironment(f, B, **kw); env.reset(seed=1)
    torch.manual_seed(0)
    buf = m.collect_random_data(env, T, dtype=torch.float64, graph_chunk=chunk)
    assert buf.size == T * B
    d = buf.get_all_data()
    o = d["observations"].view(T, B, -1); n = d["next_observations"].view(T, B, -1)
    done = d["terminals"].view(T, B).bool()
    assert bool(done[5].all()) and bool(done[11].all()) and not bool(done[:5].any())     # episode_length = 6
    for t in range(T - 1):
        keep = ~done[t]
        assert torch.equal(o[t + 1][keep], n[t][keep])
        assert torch.all(o[t + 1][done[t]][:, 0] == 1.0)                                # reset rows
    a = d["actions"]
    assert float(a.min()) >= -1.0 and float(a.max()) <= 1.0 and abs(float(a.mean())) < 0.01
    assert not torch.equal(d["actions"].view(T, B, -1)[0], d["actions"].view(T, B, -1)[chunk])   # replays draw anew
    assert torch.isfinite(d["rewards"]).all()
    # launch cost: graph replay vs the eager loop on the same environment size
    e1 = m.BatchedGridEnvironment(f, B, **kw); e1.reset(seed=1)
    m.collect_random_data(e1, chunk)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m.collect_random_data(e1, 64)
    torch.cuda.synchronize(); t_eager = time.perf_counter() - t0
    e2 = m.BatchedGridEnvironment(f, B, **kw); e2.reset(seed=1)
    col = m.GraphedCollector(e2, chunk)
    col.run_chunk(); col.run_chunk()                     # eager chunk + capture, then one replay
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(64 // chunk):
        col.run_chunk()
    torch.cuda.synchronize(); t_graph = time.perf_counter() - t0
    print(f"64 steps x {B} instances: eager loop {t_eager * 1e3:.1f} ms, graph replay {t_graph * 1e3:.1f} ms")
    assert t_graph < t_eager


@pytest.mark.parametrize("spec", ["ieee13", "ieee34"])
def test_fp32_and_alternating_observation_buffers(spec):
    """obs_dtype=float32: the kernels write the observation rounded to fp32 (the reference declares float32,
    grid_env.py:346) - bit for bit the fp64 observation cast to float, every other output untouched; two
    alternating buffers (so that observation t - 1 can be copied out under step t) change nothing either, incl.
    the rejected-action path, which carries the unchanged state over from the previous buffer, and masked resets."""
    import grid_fed_rl_b200 as m
    from oracle.ref_harness import make_feeder
    f = make_feeder(None, spec, use_reference_classes=False)
    kw = dict(renewable_sources=["solar", "wind"], timestep=60.0, repair=False, start_time=11 * 3600.0, tolerance=1e-8)
    B = 203
    ref = m.BatchedGridEnvironment(f, B, **kw)                                   # fp64, one buffer
    alt = m.BatchedGridEnvironment(f, B, obs_buffers=2, **kw)                    # fp64, alternating
    f32 = m.BatchedGridEnvironment(f, B, obs_dtype=torch.float32, **kw)          # fp32, alternating (default)
    one = m.BatchedGridEnvironment(f, B, obs_dtype="float32", obs_buffers=1, **kw)
    assert len(f32._obs_bufs) == 2 and f32.observation_space.dtype == np.float32
    envs = (ref, alt, f32, one)
    for e in envs:
        o, _ = e.reset(seed=9)
    assert torch.equal(f32.get_observation(), ref.get_observation().float()) and f32.get_observation().dtype == torch.float32
    g = torch.Generator(device="cuda"); g.manual_seed(4)
    prev32 = None
    for t in range(7):
        act = ref.sample_actions(g)
        if t in (2, 3) and ref.act_dim > 1:
            act[5, 0] = float("nan")                     # rejected twice in a row: the row travels through both buffers
            act[77, 1] = float("inf")
        outs = [e.step(act) for e in envs]
        o64, r64, t64, u64, i64 = outs[0]
        for (o, r, te, tr, info), e in zip(outs[1:], envs[1:]):
            assert torch.equal(r, r64) and torch.equal(te, t64) and torch.equal(tr, u64)
            assert torch.equal(info["iterations"], i64["iterations"]) and torch.equal(info["error"], i64["error"])
            assert torch.equal(info["max_voltage"], i64["max_voltage"])
            if e.obs_dtype == torch.float32:
                assert o.dtype == torch.float32 and torch.equal(o, o64.float())
            else:
                assert torch.equal(o, o64)
        if prev32 is not None:
            # the previous step's observation is still intact in the other buffer (what the overlapped copy reads)
            assert torch.equal(prev32[0], prev32[1])
        prev32 = (outs[2][0], outs[2][0].clone())
        if t == 4:
            mask = torch.zeros(B, dtype=torch.bool, device="cuda"); mask[::3] = True
            rs = [e.reset(mask=mask)[0] for e in envs]
            assert torch.equal(rs[1], rs[0]) and torch.equal(rs[2], rs[0].float()) and torch.equal(rs[3], rs[0].float())
            prev32 = None
    sd = f32.state_dict()
    clone = m.BatchedGridEnvironment(f, B, obs_dtype=torch.float32, **kw)
    clone.load_state_dict(sd)
    act = ref.sample_actions(g)
    assert torch.equal(clone.step(act)[0], f32.step(act)[0])
    for e in envs + (clone,):
        e.close()


@pytest.mark.parametrize("depth", [2, 3])
def test_host_stepper_overlaps_fp32_observation_copies(depth):
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200.pipeline import HostStepper
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], timestep=60.0, repair=False, start_time=10 * 3600.0)
    B = 4096
    a = m.BatchedGridEnvironment(f, B, **kw)
    b = m.BatchedGridEnvironment(f, B, obs_dtype=torch.float32, **kw)
    a.reset(seed=2); b.reset(seed=2)
    g = torch.Generator(device="cuda"); g.manual_seed(1)
    acts = [a.sample_actions(g).cpu().pin_memory() for _ in range(6)]
    st = HostStepper(b, depth=depth, observations=True)
    assert st._overlap_obs and st._own_out and st.d2h_bytes_per_step == B * (10 + 4 * (b.obs_dim - 2 * b.soa.n_load))
    got = []
    for i, act in enumerate(acts):
        st.submit(act)
        if i + 1 >= depth:
            r = st.result()
            got.append((r["observations"].clone(), r["reward"].clone()))
    while st._pending:
        r = st.result()
        got.append((r["observations"].clone(), r["reward"].clone()))
    assert len(got) == len(acts)
    for act, (o, r) in zip(acts, got):
        o64, r64, _, _, _ = a.step(act)
        assert torch.equal(o, o64.float().cpu()) and torch.equal(r, r64.cpu())
    with pytest.raises(ValueError):
        from grid_fed_rl_b200.compat import GraphedCollector
        GraphedCollector(b)


def test_host_stepper_with_auto_reset_keeps_the_environments_own_outputs():
    """auto_reset post-processes the environment's own output buffers: the stepper must not redirect them."""
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200.pipeline import HostStepper
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], timestep=60.0, repair=False, start_time=10 * 3600.0,
              episode_length=3, auto_reset=True)
    a = m.BatchedGridEnvironment(f, 256, **kw); a.reset(seed=4)
    b = m.BatchedGridEnvironment(f, 256, **kw); b.reset(seed=4)
    g = torch.Generator(device="cuda"); g.manual_seed(9)
    acts = [a.sample_actions(g).cpu().pin_memory() for _ in range(7)]
    st = HostStepper(b, depth=2)
    assert not st._own_out
    got = []
    for i, x in enumerate(acts):
        st.submit(x)
        if i >= 1:
            h = st.result(); got.append((h["reward"].clone(), h["terminated"].clone(), h["truncated"].clone()))
    while st._pending:
        h = st.result(); got.append((h["reward"].clone(), h["terminated"].clone(), h["truncated"].clone()))
    seen_done = False
    for x, (r1, t1, u1) in zip(acts, got):
        _, r0, t0, u0, _ = a.step(x)
        assert torch.equal(r0.cpu(), r1) and torch.equal(t0.cpu(), t1) and torch.equal(u0.cpu(), u1)
        seen_done |= bool(t1.any())
    assert seen_done
    with pytest.raises(ValueError):
        b.step(acts[0], _step_out=a.step_outputs_into(torch.zeros(256, dtype=torch.float64, device="cuda"),
                                                      torch.zeros(256, dtype=torch.uint8, device="cuda"),
                                                      torch.zeros(256, dtype=torch.uint8, device="cuda")))
    a.close(); b.close()


def test_multi_agent_wrapper_on_the_batched_environment():
    """Three agents (battery, solar, wind) over IEEE-13: the joint action drives the same step, every agent gets
    its slice of the observation as a view, the shared reward is split evenly (multi_agent.py:37-135)."""
    import grid_fed_rl_b200 as m
    f = m.repair_topology(m.IEEE13Bus())
    kw = dict(renewable_sources=["solar", "wind"], timestep=60.0, repair=False, start_time=10 * 3600.0)
    B = 64
    base, plain = m.BatchedGridEnvironment(f, B, **kw), m.BatchedGridEnvironment(f, B, **kw)
    D = base.obs_dim
    cfgs = [m.AgentConfig("battery", 26, 1), m.AgentConfig("solar", 30, 1), m.AgentConfig("wind", D - 56 + 4, 1)]
    w = m.MultiAgentEnvironmentWrapper(base, cfgs)
    plain.reset(seed=4)
    first = w.reset(seed=4)
    assert first["battery"].shape == (B, 26) and first["wind"].shape == (B, D - 52)
    g = torch.Generator(device="cuda"); g.manual_seed(8)
    for t in range(3):
        act = plain.sample_actions(g)
        obs, rew, done, info = w.step({"battery": act[:, 0:1], "solar": act[:, 1], "wind": act[:, 2:3]})
        o, r, te, tr, _ = plain.step(act)
        assert torch.equal(obs["battery"], o[:, :26]) and torch.equal(obs["solar"], o[:, 26:56])
        assert torch.equal(obs["wind"][:, :D - 56], o[:, 56:]) and torch.equal(obs["wind"][:, D - 56:], torch.zeros(B, 4, dtype=o.dtype, device=o.device))
        for a in ("battery", "solar", "wind"):
            assert torch.equal(rew[a], r / 3) and torch.equal(done[a], te | tr)
    obs, rew, done, info = w.step({"battery": torch.zeros(B, 1)})          # the others send nothing: zeros
    o, r, _, _, _ = plain.step(torch.zeros(B, 3, dtype=torch.float64))
    assert torch.equal(obs["solar"], o[:, 26:56]) and torch.equal(rew["wind"], r / 3)

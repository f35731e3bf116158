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
    assert st.d2h_bytes_per_step == 512 * (10 + 8 * b.obs_dim)
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

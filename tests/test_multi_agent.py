"""CPU tier: the tensor ``MultiAgentEnvironmentWrapper`` against the reference's
(``/root/reference/grid_fed_rl/algorithms/multi_agent.py:37-135``).

``tests/golden/multi_agent_wrapper.npz`` was written by ``oracle/ref_harness.py multi_agent``: the UNMODIFIED
reference wrapper around a scripted base environment.  Here the same script drives a stub batched environment
(B instances, instance b = the script scaled by b + 1) through this package's wrapper."""
import os

import numpy as np
import pytest
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multi_agent_wrapper.npz")


class _ScriptedBatch:
    """Stands in for BatchedGridEnvironment: tensors [B, ...], instance b sees (b + 1) x the script."""

    def __init__(self, g, B):
        self.g, self.num_envs, self.device = g, B, torch.device("cpu")
        self.act_dim = int(g["act_dims"].sum())
        self.scale = torch.arange(1, B + 1, dtype=torch.float64)
        self.t, self.joint = 0, []

    def reset(self):
        self.t = 0
        return torch.as_tensor(self.g["obs_script"][0])[None, :] * self.scale[:, None], {}

    def step(self, joint):
        self.joint.append(joint.clone())
        t = self.t
        self.t += 1
        obs = torch.as_tensor(self.g["obs_script"][t + 1])[None, :] * self.scale[:, None]
        rew = float(self.g["rew_script"][t]) * self.scale
        term = torch.full((self.num_envs,), bool(self.g["done_script"][t][0]))
        trunc = torch.full((self.num_envs,), bool(self.g["done_script"][t][1]))
        info = {"solar_reward_bonus": float(self.g["bonus_script"][t]) * self.scale, "t": t}
        return obs, rew, term, trunc, info


def test_wrapper_equals_reference():
    import grid_fed_rl_b200 as m
    z = np.load(GOLDEN, allow_pickle=False)
    g = {k: z[k] for k in z.files}
    agents = [str(a) for a in g["agents"]]
    B = 3
    base = _ScriptedBatch(g, B)
    w = m.MultiAgentEnvironmentWrapper(base, [m.AgentConfig(a, int(o), int(d)) for a, o, d in
                                              zip(agents, g["obs_dims"], g["act_dims"])])
    assert w.n_agents == 4 and list(w.agent_obs_dims) == agents
    first = w.reset()
    for a in agents:
        for b in range(B):
            assert np.array_equal(first[a][b].numpy(), g[f"reset_obs_{a}"] * (b + 1)), a
    steps = g["joint_actions"].shape[0]
    for t in range(steps):
        acts = {}
        for a in agents:
            k = f"step{t}_act_{a}"
            if k in g:
                v = torch.as_tensor(g[k])[None, :].repeat(B, 1)
                acts[a] = v[:, 0] if a == "solar" else v          # a [B] vector for the one-entry agent
        obs, rew, done, info = w.step(acts)
        assert np.array_equal(base.joint[-1].numpy(), np.tile(g["joint_actions"][t], (B, 1)))
        for a in agents:
            for b in range(B):
                assert np.array_equal(obs[a][b].numpy(), g[f"step{t}_obs_{a}"] * (b + 1)), (t, a)
                assert abs(float(rew[a][b]) - float(g[f"step{t}_rew_{a}"]) * (b + 1)) < 1e-12, (t, a)
                assert bool(done[a][b]) == bool(g[f"step{t}_done_{a}"])
            assert info[a]["t"] == t
    # slices inside the observation are views of the environment's buffer (no copy), the padded one is not
    o = base.reset()[0]
    parts = w._split_observation(o)
    assert parts["battery"].data_ptr() == o.data_ptr() and parts["observer"].shape == (B, 6)
    assert torch.equal(parts["observer"][:, 3:], torch.zeros(B, 3, dtype=torch.float64))
    with pytest.raises(Exception):
        m.MultiAgentEnvironmentWrapper(base, [m.AgentConfig("a", 3, 9)]).step({})

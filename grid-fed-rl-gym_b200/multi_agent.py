"""``MultiAgentEnvironmentWrapper`` on tensors - the reference's obs / action splitting
(``/root/reference/grid_fed_rl/algorithms/multi_agent.py:37-135``) over a ``BatchedGridEnvironment``.

Same constructor (a base environment and a list of ``AgentConfig``), same methods and dict-of-agents return
shapes; every value gains a leading batch axis and stays on the device:

* ``reset() -> {agent_id: obs[B, observation_dim]}``
* ``step({agent_id: action[B, action_dim]}) -> (obs, rewards, dones, infos)`` with
  ``rewards[agent] = reward[B] / n_agents (+ info["<agent>_reward_bonus"] if present)``,
  ``dones[agent] = terminated | truncated``, ``infos[agent] = info`` (the same dict for every agent).

Splitting rules, as upstream: agents take consecutive slices of the global observation in the order of the
config list; a slice that runs past the end is zero-padded (``_split_observation``); a missing agent action
is a zero action, a scalar action counts as one entry, actions are flattened per instance in config order
(``_combine_actions``).  Slices are views of the observation buffer (no copy); the joint action is assembled
into one preallocated ``[B, A]`` tensor.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict, List, Optional, Sequence, Tuple

import torch


@dataclass
class AgentConfig:
    """Reference ``AgentConfig`` (multi_agent.py:24-33): the fields the wrapper reads, plus the learner hints."""
    agent_id: str
    observation_dim: int
    action_dim: int
    agent_type: str = "continuous"
    learning_rate: float = 1e-3
    hidden_dims: Optional[List[int]] = None


class MultiAgentEnvironmentWrapper:
    def __init__(self, base_env, agent_configs: Sequence[AgentConfig]) -> None:
        self.base_env = base_env
        self.agent_configs = {c.agent_id: c for c in agent_configs}
        self.n_agents = len(agent_configs)
        self.agent_obs_dims = {c.agent_id: int(c.observation_dim) for c in agent_configs}
        self.agent_action_dims = {c.agent_id: int(c.action_dim) for c in agent_configs}
        B, dev = base_env.num_envs, base_env.device
        total = sum(self.agent_action_dims.values())
        self._joint = torch.zeros(B, total, dtype=torch.float64, device=dev)
        self._act_slices, o = {}, 0
        for aid, d in self.agent_action_dims.items():
            self._act_slices[aid] = slice(o, o + d)
            o += d

    # -- splitting ------------------------------------------------------------------
    def _split_observation(self, global_obs: torch.Tensor) -> Dict[str, torch.Tensor]:
        B, D = global_obs.shape
        out, start = {}, 0
        for aid, dim in self.agent_obs_dims.items():
            end = start + dim
            if end <= D:
                out[aid] = global_obs[:, start:end]                    # a view: nothing is copied
            else:
                pad = torch.zeros(B, dim, dtype=global_obs.dtype, device=global_obs.device)
                if start < D:
                    pad[:, :D - start] = global_obs[:, start:]
                out[aid] = pad
            start = end
        return out

    def _combine_actions(self, actions: Dict[str, Any]) -> torch.Tensor:
        joint = self._joint
        joint.zero_()                                                   # an agent without an action sends zeros
        for aid, sl in self._act_slices.items():
            if aid not in actions:
                continue
            a = torch.as_tensor(actions[aid], dtype=torch.float64, device=joint.device)
            if a.dim() == 0:
                a = a.reshape(1, 1).expand(joint.shape[0], 1)
            joint[:, sl] = a.reshape(joint.shape[0], -1)
        return joint

    def _split_reward(self, global_reward: torch.Tensor, info: Dict[str, Any]) -> Dict[str, torch.Tensor]:
        base = global_reward / self.n_agents
        out = {}
        for aid in self.agent_configs:
            r = base
            key = f"{aid}_reward_bonus"
            if key in info:
                r = r + info[key]
            out[aid] = r
        return out

    # -- API --------------------------------------------------------------------------
    def reset(self, **kw) -> Dict[str, torch.Tensor]:
        global_obs, _ = self.base_env.reset(**kw)
        return self._split_observation(global_obs)

    def step(self, actions: Dict[str, Any]) -> Tuple[Dict[str, torch.Tensor], Dict[str, torch.Tensor],
                                                     Dict[str, torch.Tensor], Dict[str, Any]]:
        joint = self._combine_actions(actions)
        if joint.shape[1] != self.base_env.act_dim:
            from .errors import InvalidActionError
            raise InvalidActionError(f"the agents' action dims add up to {joint.shape[1]}, the environment takes "
                                     f"{self.base_env.act_dim}")
        global_obs, reward, terminated, truncated, info = self.base_env.step(joint)
        done = terminated | truncated
        return (self._split_observation(global_obs), self._split_reward(reward, info),
                {aid: done for aid in self.agent_configs}, {aid: info for aid in self.agent_configs})

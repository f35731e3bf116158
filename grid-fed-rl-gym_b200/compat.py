"""Reference-shaped fronts of ``BatchedGridEnvironment`` and the device-resident rollout buffer.

* ``GridEnvironment`` - one instance, the reference's return types: the observation is a Python
  list of floats (grid_env.py:753-783), the reward a float, the flags bools, ``info`` the reference's
  dict with ``constraint_violations`` as the dict of four bools (grid_env.py:610-617).
  Every step is a kernel launch plus a device -> host read: use it to drop into existing callers,
  not for throughput.
* ``VectorizedEnvironment`` - the list-in / list-out API of
  utils/parallel_environment.py:283-355 over ONE batched environment (no thread pool).
* ``RolloutBuffer`` / ``collect_random_data`` - the reference's offline-data path
  (algorithms/base.py:180-298) with the transitions kept in HBM: ``observations, actions, rewards,
  next_observations, terminals`` tensors, z-normalisation and ``sample_batch`` on the device.
"""

from __future__ import annotations

from typing import Any, Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

from .env import BatchedGridEnvironment
from .errors import InvalidActionError

VIOLATION_KEYS = ("voltage_high", "voltage_low", "frequency_high", "frequency_low")


def _info_row(info: Dict[str, Any], i: int, host: Dict[str, np.ndarray]) -> Dict[str, Any]:
    """The reference's info dict (base.py:169-176 updated at grid_env.py:610-617) for instance i."""
    viol = host["constraint_violations"][i]
    out = {
        "current_step": int(host["current_step"][i]),
        "episode_reward": float(host["episode_reward"][i]),
        "timestep": info["timestep"],
        "power_flow_converged": bool(host["power_flow_converged"][i]),
        "max_voltage": float(host["max_voltage"][i]),
        "min_voltage": float(host["min_voltage"][i]),
        "total_losses": float(host["total_losses"][i]),
        "constraint_violations": {k: bool(v) for k, v in zip(VIOLATION_KEYS, viol)},
        "constraint_violation_count": int(host["constraint_violation_count"][i]),
        "iterations": int(host["iterations"][i]),
    }
    if host["error"][i]:
        # grid_env.py:454-467: the reference reports the rejected action this way
        out["error"] = "invalid action (NaN / Inf)"
        out["unexpected_error"] = True
    return out


def _to_host(info: Dict[str, Any]) -> Dict[str, np.ndarray]:
    return {k: v.cpu().numpy() for k, v in info.items() if isinstance(v, torch.Tensor)}


class GridEnvironment:
    """One feeder instance with the reference's scalar return types (``num_envs=1``
    unwraps).  Constructor kwargs are the reference's (grid_env.py:161-174) plus the batched
    environment's (``solver``, ``tolerance``, ``max_iterations``, ``device`` ...)."""

    def __init__(self, feeder, **kwargs) -> None:
        kwargs.pop("num_envs", None)
        self._env = BatchedGridEnvironment(feeder, 1, **kwargs)
        self.feeder = self._env.feeder
        self.observation_space = self._env.observation_space
        self.action_space = self._env.action_space
        self.timestep = self._env.timestep
        self.episode_length = self._env.episode_length
        self.safety_penalty = self._env.safety_penalty
        self.current_step = 0
        self.episode_reward = 0.0
        self.constraint_violations = 0

    def _sync_counters(self, host) -> None:
        self.current_step = int(host["current_step"][0])
        self.episode_reward = float(host["episode_reward"][0])
        self.constraint_violations = int(host["constraint_violation_count"][0])

    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None, *, noise=None):
        obs, info = self._env.reset(seed=seed, options=options, noise=noise)
        host = _to_host(info)
        self._sync_counters(host)
        # reset's info is base.py:169-176 only
        return obs[0].tolist(), {"current_step": 0, "episode_reward": 0.0, "constraint_violations": 0,
                                 "timestep": self.timestep}

    def step(self, action, noise=None):
        """``noise`` (optional, [4 + L]) replays the reference's random draws (parity runs)."""
        A = self._env.act_dim
        try:
            act = np.asarray(action, dtype=np.float64).reshape(1, A)
        except (TypeError, ValueError):
            # non-numeric / wrong-length actions take the reference's error path too
            act = np.full((1, A), np.nan)
            if A == 1:
                act[:] = 0.0
        obs, reward, term, trunc, info = self._env.step(act, None if noise is None else np.asarray(noise)[None, :])
        host = _to_host(info)
        self._sync_counters(host)
        return (obs[0].tolist(), float(reward[0].item()), bool(term[0].item()), bool(trunc[0].item()),
                _info_row(info, 0, host))

    def get_observation(self) -> List[float]:
        return self._env.get_observation()[0].tolist()

    def close(self) -> None:
        self._env.close()


class VectorizedEnvironment:
    """``VectorizedEnvironment.reset(seeds) -> (observations, infos)`` and
    ``.step(actions) -> (observations, rewards, dones, truncateds, infos)`` as Python lists
    (utils/parallel_environment.py:309-355), on one batched environment.  ``environment_factory``
    is called once and must return a ``BatchedGridEnvironment`` (or pass one directly)."""

    def __init__(self, environment_factory, num_envs: Optional[int] = None, config: Any = None) -> None:
        env = environment_factory() if callable(environment_factory) else environment_factory
        if not isinstance(env, BatchedGridEnvironment):
            raise TypeError("environment_factory must produce a BatchedGridEnvironment")
        if num_envs is not None and int(num_envs) != env.num_envs:
            raise ValueError(f"factory built {env.num_envs} instances, num_envs says {num_envs}")
        self.env, self.num_envs, self.config = env, env.num_envs, config
        self.step_count = self.reset_count = 0

    def reset(self, seeds: Optional[Sequence[int]] = None) -> Tuple[List[Any], List[Dict]]:
        if seeds is not None and len(seeds) != self.num_envs:
            raise ValueError(f"Expected {self.num_envs} seeds, got {len(seeds)}")
        obs, info = self.env.reset(seeds=None if seeds is None else np.asarray(seeds, dtype=np.int64))
        self.reset_count += 1
        base = {"current_step": 0, "episode_reward": 0.0, "constraint_violations": 0,
                "timestep": self.env.timestep}
        return obs.cpu().tolist(), [dict(base) for _ in range(self.num_envs)]

    def step(self, actions: Sequence[Any]):
        if len(actions) != self.num_envs:
            raise ValueError(f"Expected {self.num_envs} actions, got {len(actions)}")   # :332-333
        act = actions if isinstance(actions, torch.Tensor) else np.asarray(actions, dtype=np.float64)
        obs, reward, term, trunc, info = self.env.step(act)
        host = _to_host(info)
        self.step_count += 1
        return (obs.cpu().tolist(), reward.cpu().tolist(), term.cpu().tolist(), trunc.cpu().tolist(),
                [_info_row(info, i, host) for i in range(self.num_envs)])

    def close(self) -> None:
        self.env.close()

    def get_performance_stats(self) -> Dict[str, Any]:
        return {"num_environments": self.num_envs, "total_steps": self.step_count,
                "total_resets": self.reset_count}


class RolloutBuffer:
    """Transitions in HBM with the field names of the reference's ``GridDataset``
    (algorithms/base.py:180-265).  Capacity is in transitions; ``add`` appends one batched step."""

    FIELDS = ("observations", "actions", "rewards", "next_observations", "terminals")

    def __init__(self, capacity: int, obs_dim: int, act_dim: int, device, dtype=torch.float32) -> None:
        self.capacity, self.size, self.device, self.dtype = int(capacity), 0, torch.device(device), dtype
        z = dict(device=self.device, dtype=dtype)
        self.observations = torch.empty(capacity, obs_dim, **z)
        self.next_observations = torch.empty(capacity, obs_dim, **z)
        self.actions = torch.empty(capacity, act_dim, **z)
        self.rewards = torch.empty(capacity, **z)
        self.terminals = torch.empty(capacity, **z)
        self.normalized = False

    def add(self, obs, actions, rewards, next_obs, terminals) -> int:
        b = obs.shape[0]
        if self.size + b > self.capacity:
            b = self.capacity - self.size
        if b <= 0:
            return 0
        s = slice(self.size, self.size + b)
        self.observations[s].copy_(obs[:b])
        self.actions[s].copy_(actions[:b])
        self.rewards[s].copy_(rewards[:b])
        self.next_observations[s].copy_(next_obs[:b])
        self.terminals[s].copy_(terminals[:b])
        self.size += b
        return b

    def normalize(self) -> None:
        """z-scores as ``GridDataset._normalize_data`` (base.py:212-228): population std + 1e-6."""
        n = self.size
        o, a, r = self.observations[:n], self.actions[:n], self.rewards[:n]
        self.obs_mean, self.obs_std = o.mean(0), o.std(0, unbiased=False) + 1e-6
        self.action_mean, self.action_std = a.mean(0), a.std(0, unbiased=False) + 1e-6
        self.reward_mean, self.reward_std = r.mean(), r.std(unbiased=False) + 1e-6
        o.sub_(self.obs_mean).div_(self.obs_std)
        self.next_observations[:n].sub_(self.obs_mean).div_(self.obs_std)
        a.sub_(self.action_mean).div_(self.action_std)
        r.sub_(self.reward_mean).div_(self.reward_std)
        self.normalized = True

    def get_all_data(self) -> Dict[str, torch.Tensor]:
        return {k: getattr(self, k)[:self.size] for k in self.FIELDS}

    def sample_batch(self, batch_size: int, generator: Optional[torch.Generator] = None) -> Dict[str, torch.Tensor]:
        idx = torch.randint(self.size, (batch_size,), device=self.device, generator=generator)
        return {k: getattr(self, k)[idx] for k in self.FIELDS}

    def denormalize_action(self, action: torch.Tensor) -> torch.Tensor:
        return action * self.action_std + self.action_mean if self.normalized else action

    def denormalize_observation(self, obs: torch.Tensor) -> torch.Tensor:
        return obs * self.obs_std + self.obs_mean if self.normalized else obs

    def to_numpy(self) -> Dict[str, np.ndarray]:
        """The dict of arrays ``collect_random_data`` returns upstream (base.py:292-298)."""
        out = {k: v.cpu().numpy() for k, v in self.get_all_data().items()}
        out["terminals"] = out["terminals"].astype(bool)
        return out


class GraphedCollector:
    """``chunk`` consecutive steps of the U(-1, 1) policy - action sampling, the fused step, the
    copies into a staging block and the masked reset of finished instances - captured once in a
    CUDA graph and replayed: one graph launch per ``chunk`` steps instead of a dozen Python-issued
    launches per step.  That matters for small feeders, whose step kernel (0.14 ms for 65 536
    IEEE-13 instances) is shorter than the host-side launch work around it."""

    def __init__(self, env: BatchedGridEnvironment, chunk: int = 8, dtype=torch.float32) -> None:
        if env.auto_reset or env.copy_outputs or len(env._obs_bufs) != 1:
            raise ValueError("GraphedCollector needs an environment with auto_reset=False, copy_outputs=False and one "
                             "observation buffer (a captured graph replays fixed pointers)")
        self.env, self.chunk = env, int(chunk)
        B, D, A, dev = env.num_envs, env.obs_dim, env.act_dim, env.device
        z = dict(device=dev, dtype=dtype)
        self.stage = dict(observations=torch.empty(chunk, B, D, **z), actions=torch.empty(chunk, B, A, **z),
                          rewards=torch.empty(chunk, B, **z), next_observations=torch.empty(chunk, B, D, **z),
                          terminals=torch.empty(chunk, B, **z))
        self._graph = None

    def _body(self) -> None:
        env, st = self.env, self.stage
        for k in range(self.chunk):
            act = torch.empty(env.num_envs, env.act_dim, dtype=torch.float64, device=env.device).uniform_(-1.0, 1.0)
            # the environment's single observation buffer holds the state the step starts from (after the masked
            # reset of the previous step): it goes to the block before the kernel overwrites it in place
            st["observations"][k].copy_(env.get_observation())
            nxt, reward, term, trunc, _ = env.step(act)
            done = term | trunc
            st["actions"][k].copy_(act)
            st["rewards"][k].copy_(reward)
            st["next_observations"][k].copy_(nxt)
            st["terminals"][k].copy_(done)
            env.reset(mask=done)                     # rewrites the observation rows of finished instances only

    def run_chunk(self) -> Dict[str, torch.Tensor]:
        """Advance ``chunk`` steps; returns the staging block (overwritten by the next call)."""
        if self._graph is None:
            # first chunk eagerly (it also warms every allocation up), then capture for all later ones
            self._body()
            torch.cuda.synchronize(self.env.device)
            graph = torch.cuda.CUDAGraph()
            snapshot = self.env.state_dict()
            first = {k: v.clone() for k, v in self.stage.items()}
            with torch.cuda.graph(graph):
                self._body()
            # capturing records launches without running them, but be explicit that nothing moved
            self.env.load_state_dict(snapshot)
            for k, v in first.items():
                self.stage[k].copy_(v)
            self._graph = graph
            return self.stage
        self._graph.replay()
        return self.stage


def collect_random_data(env: BatchedGridEnvironment, num_steps: int, normalize: bool = False,
                        generator: Optional[torch.Generator] = None, dtype=torch.float32,
                        graph_chunk: int = 0) -> RolloutBuffer:
    """``collect_random_data(env, n)`` (base.py:268-298) for a batched environment: ``num_steps``
    batched steps of a U(-1, 1) policy -> ``num_steps * num_envs`` transitions, instances that
    terminate or truncate are reset (masked) before their next step.  Nothing leaves the GPU.
    ``graph_chunk > 0`` replays that many steps per CUDA-graph launch (``GraphedCollector``; actions
    then come from torch's default CUDA generator)."""
    B = env.num_envs
    buf = RolloutBuffer(num_steps * B, env.obs_dim, env.act_dim, env.device, dtype)
    if graph_chunk > 0:
        env.reset()
        col = GraphedCollector(env, graph_chunk, dtype)
        left = num_steps
        while left > 0:
            st = col.run_chunk()
            take = min(left, graph_chunk)
            buf.add(st["observations"][:take].reshape(take * B, -1), st["actions"][:take].reshape(take * B, -1),
                    st["rewards"][:take].reshape(-1), st["next_observations"][:take].reshape(take * B, -1),
                    st["terminals"][:take].reshape(-1))
            left -= take
        if normalize:
            buf.normalize()
        return buf
    obs, _ = env.reset()
    obs = obs.clone()
    for _ in range(num_steps):
        act = env.sample_actions(generator)
        nxt, reward, term, trunc, _ = env.step(act)
        done = term | trunc
        buf.add(obs, act, reward, nxt, done)
        obs.copy_(nxt)
        if bool(done.any()):                      # one flag read per step, as the reference loop does
            fresh, _ = env.reset(mask=done)
            obs.copy_(fresh)
    if normalize:
        buf.normalize()
    return buf

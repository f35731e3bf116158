"""Host-driven stepping with the PCIe copies off the critical path.

``HostStepper`` is for callers whose actions live in host memory (a CPU policy, a recorded
action log, the reference's own loops): every step still copies that step's actions host ->
device and its reward / done flags device -> host, but the copy of step t+1's actions runs on a
second stream while step t's kernel executes, and the results come back through pinned buffers
one step behind the submissions (depth-2 software pipeline).  The step writes its reward / done flags
into one of ``depth`` alternating device buffers, so their copy to the host runs on a third stream
under the next step's kernel and the kernels follow each other without a gap.  ``depth=1`` degenerates
to the plain copy - step - copy - synchronise sequence.
"""

from __future__ import annotations

from collections import deque
from typing import Deque, Dict, Tuple

import torch

from . import _native as nat
from .env import BatchedGridEnvironment


class HostStepper:
    def __init__(self, env: BatchedGridEnvironment, depth: int = 2, observations: bool = False) -> None:
        """``observations=True`` also brings every step's observation ``[B, D]`` back to pinned host
        memory (what the reference's list-returning ``step`` implies for a host policy): 8 D bytes per
        instance per step over PCIe, which then bounds the rate."""
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.env, self.depth, self.observations = env, depth, bool(observations)
        dev, B, A = env.device, env.num_envs, env.act_dim
        self.copy_stream = torch.cuda.Stream(device=dev)      # host -> device: actions
        self.out_stream = torch.cuda.Stream(device=dev)       # device -> host: results
        self._act = [torch.empty(B, A, dtype=torch.float64, device=dev) for _ in range(depth)]
        self._host = [dict(reward=torch.empty(B, dtype=torch.float64).pin_memory(),
                           terminated=torch.empty(B, dtype=torch.bool).pin_memory(),
                           truncated=torch.empty(B, dtype=torch.bool).pin_memory())
                      for _ in range(depth)]
        # with two alternating observation buffers in the environment (obs_buffers=2, the default for fp32
        # observations) the observation of step t is copied on the copy stream while step t + 1 runs
        self._overlap_obs = self.observations and len(getattr(env, "_obs_bufs", [])) == 2
        if self.observations:
            # The 2L static load columns of the observation never change after construction (the reference reads
            # them as constants, grid_env.py:766-770): the host rows get them once, here, and every step copies
            # only the columns around them (IEEE-123: 502 of 692 entries).
            soa = env.soa
            c0 = 2 * soa.n_bus + 2 * soa.n_line + 1
            c1 = c0 + 2 * soa.n_load
            self._dyn_cols = [(0, c0), (c1, env.obs_dim)] if c1 > c0 else [(0, env.obs_dim)]
            first = env.get_observation().cpu()
            for h in self._host:
                h["observations"] = torch.empty(B, env.obs_dim, dtype=env.obs_dtype).pin_memory()
                h["observations"].copy_(first)
        # The step's reward / done flags in alternating device buffers of the stepper's own (the environment's
        # single set would have to be copied out before the next kernel may start).  Not with auto_reset /
        # copy_outputs (the environment post-processes its own outputs there) and not when a single observation
        # buffer has to be copied out between two kernels anyway.
        self._own_out = (not (env.auto_reset or env.copy_outputs)) and (not self.observations or self._overlap_obs)
        if self._own_out:
            self._dev_out = [dict(reward=torch.zeros(B, dtype=torch.float64, device=dev),
                                  terminated=torch.zeros(B, dtype=torch.uint8, device=dev),
                                  truncated=torch.zeros(B, dtype=torch.uint8, device=dev)) for _ in range(depth)]
            self._step_outs = [env.step_outputs_into(d["reward"], d["terminated"], d["truncated"]) for d in self._dev_out]
        self._stepped = [torch.cuda.Event() for _ in range(depth)]
        self._copied = [torch.cuda.Event() for _ in range(depth)]
        self._done = [torch.cuda.Event() for _ in range(depth)]
        self._busy = [False] * depth
        self._pending: Deque[int] = deque()
        self._n = 0

    @property
    def h2d_bytes_per_step(self) -> int:
        return self.env.num_envs * self.env.act_dim * 8

    @property
    def d2h_bytes_per_step(self) -> int:
        item = 4 if self.env.obs_dtype == torch.float32 else 8
        cols = sum(b - a for a, b in self._dyn_cols) if self.observations else 0
        return self.env.num_envs * (8 + 1 + 1 + item * cols)

    def _copy_obs(self, dst: torch.Tensor, obs: torch.Tensor) -> None:
        # one strided DMA per column block on the current stream (a sliced tensor.copy_ would stage the block
        # through a temporary and finish it with a synchronous host-side scatter)
        env = self.env
        stream = torch.cuda.current_stream(env.device).cuda_stream
        for a, b in self._dyn_cols:
            if b > a:
                nat.check(env.lib, env.lib.gfr_env_obs_to_host(env._h, obs.data_ptr(), dst.data_ptr(), a, b - a, stream))

    def submit(self, host_actions: torch.Tensor) -> None:
        """Queue one step.  ``host_actions``: pinned fp64 ``[B, A]`` (a pageable tensor works but
        makes the copy synchronous)."""
        j = self._n % self.depth
        if len(self._pending) >= self.depth:
            raise RuntimeError("pipeline full: call result() first")
        compute = torch.cuda.current_stream(self.env.device)
        h = self._host[j]
        if self._own_out:
            with torch.cuda.stream(self.copy_stream):
                if self._busy[j]:
                    self.copy_stream.wait_event(self._stepped[j])   # the kernel that last read this action buffer
                self._act[j].copy_(host_actions, non_blocking=True)
                self._copied[j].record(self.copy_stream)
            compute.wait_event(self._copied[j])
            if self._busy[j]:
                compute.wait_event(self._done[j])                   # the results last written to this set have left
            if self._overlap_obs and self.depth > 2 and self._n >= 2 and self._busy[(self._n - 2) % self.depth]:
                compute.wait_event(self._done[(self._n - 2) % self.depth])   # ... and the observation buffer written two steps ago
            obs = self.env.step(self._act[j], _step_out=self._step_outs[j])[0]
            self._stepped[j].record(compute)
            d = self._dev_out[j]
            with torch.cuda.stream(self.out_stream):
                self.out_stream.wait_event(self._stepped[j])
                h["reward"].copy_(d["reward"], non_blocking=True)
                h["terminated"].copy_(d["terminated"].view(torch.bool), non_blocking=True)
                h["truncated"].copy_(d["truncated"].view(torch.bool), non_blocking=True)
                if self.observations:
                    self._copy_obs(h["observations"], obs)
                self._done[j].record(self.out_stream)
            self._busy[j] = True
            self._pending.append(j)
            self._n += 1
            return
        with torch.cuda.stream(self.copy_stream):
            if self._busy[j]:
                self.copy_stream.wait_event(self._done[j])      # the step that last read this buffer
            self._act[j].copy_(host_actions, non_blocking=True)
            self._copied[j].record(self.copy_stream)
        compute.wait_event(self._copied[j])
        obs, reward, term, trunc, _ = self.env.step(self._act[j])
        if self.observations and not self._overlap_obs:
            self._copy_obs(h["observations"], obs)
        h["reward"].copy_(reward, non_blocking=True)
        h["terminated"].copy_(term, non_blocking=True)
        h["truncated"].copy_(trunc, non_blocking=True)
        if self._overlap_obs:
            # the step after next writes this observation buffer again: by then result() has waited for this copy
            self._stepped[j].record(compute)
            with torch.cuda.stream(self.copy_stream):
                self.copy_stream.wait_event(self._stepped[j])
                self._copy_obs(h["observations"], obs)
                self._done[j].record(self.copy_stream)
            self._busy[j] = True
            self._pending.append(j)
            self._n += 1
            return
        self._done[j].record(compute)
        self._busy[j] = True
        self._pending.append(j)
        self._n += 1

    def result(self) -> Dict[str, torch.Tensor]:
        """Host tensors (pinned, reused ``depth`` steps later) of the oldest submitted step."""
        j = self._pending.popleft()
        self._done[j].synchronize()
        return self._host[j]

    def drain(self) -> None:
        while self._pending:
            self.result()

"""``BatchedGridEnvironment`` - the reference's ``GridEnvironment`` reset / step / observation /
info contract for B independent feeder instances on one GPU, one fused kernel per step.

Reference surface mirrored here (paths under /root/reference/grid_fed_rl/):
  * constructor kwargs             environments/grid_env.py:161-174
  * ``reset(seed, options)``       environments/grid_env.py:360-408   -> (obs, info)
  * ``step(action)``               environments/grid_env.py:410-619   -> (obs, reward, terminated, truncated, info)
  * ``observation_space`` / ``action_space`` / ``current_step`` / ``episode_reward`` /
    ``constraint_violations``      environments/grid_env.py:300-358, environments/base.py:71-176
  * list-of-envs batching          utils/parallel_environment.py:283-355 (``VectorizedEnvironment``)

All arithmetic happens in ``libgfr_b200.so`` (hand-written sm_100a CUDA behind the C ABI of
``include/gfr_b200.h``); this module only owns tensors and argument checking.  There is no
CPU path: without the library or without a CUDA device construction fails.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Dict, Optional, Sequence, Tuple

import numpy as np
import torch

from . import _native as nat
from .distributed import all_reduce_stats, global_seeds, shard_range, stats_dict  # noqa: F401
from .components import Box
from .errors import InvalidActionError, InvalidConfigurationError, NetworkTopologyError
from .topology import FeederSoA, TopologyError, compile_for_solver, repair_topology


def _compile(feeder, renewable_sources, repair, solver, lanes, keep_cycles=False) -> Tuple[FeederSoA, Any, int]:
    if isinstance(feeder, FeederSoA):
        # a precompiled feeder runs on the lane count its level schedule was capped for
        return feeder, None, int(lanes) or int(getattr(feeder, "lanes_hint", 0) or 0)
    try:
        if repair is True:
            feeder = repair_topology(feeder, keep_cycles=keep_cycles)
        soa, lanes = compile_for_solver(feeder, solver, lanes, renewable_sources=renewable_sources)
        return soa, feeder, lanes
    except TopologyError as exc:
        if repair == "auto":
            fixed = repair_topology(feeder, keep_cycles=keep_cycles)
            soa, lanes = compile_for_solver(fixed, solver, lanes, renewable_sources=renewable_sources)
            return soa, fixed, lanes
        raise NetworkTopologyError(str(exc)) from exc


class NativeFeeder:
    """A compiled feeder resident on one device (``gfr_feeder``)."""

    def __init__(self, soa: FeederSoA, device: torch.device) -> None:
        self.lib = nat.load_library()
        self.soa = soa
        self.device = device
        desc, keep = nat.make_feeder_desc(soa)
        h = C.c_void_p()
        nat.check(self.lib, self.lib.gfr_feeder_create(C.byref(desc), device.index, C.byref(h)))
        self.handle = h

    def close(self) -> None:
        if getattr(self, "handle", None):
            self.lib.gfr_feeder_destroy(self.handle)
            self.handle = None

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass


def _cuda_device(device) -> torch.device:
    dev = torch.device(device)
    if dev.type != "cuda":
        raise InvalidConfigurationError("grid_fed_rl_b200 runs on CUDA devices only (no CPU fallback)")
    if not torch.cuda.is_available():
        raise nat.NativeLibraryMissing("no CUDA device is visible; grid_fed_rl_b200 has no CPU fallback")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class BatchedGridEnvironment:
    """B instances of one feeder; tensors in, tensors out.

    ``step`` returns views of buffers that the next ``step`` / ``reset`` overwrites (the observation
    buffer is owned by the native library); ``clone()`` what must survive, or pass
    ``copy_outputs=True``.
    """

    INFO_KEYS = ("power_flow_converged", "max_voltage", "min_voltage", "total_losses",
                 "constraint_violations", "constraint_violation_count", "iterations", "current_step",
                 "episode_reward", "error", "max_mismatch")

    def __init__(self, feeder, num_envs: int = 1, device="cuda", *, timestep: float = 1.0,
                 episode_length: int = 86400, stochastic_loads: bool = True,
                 renewable_sources: Optional[Sequence[str]] = None, weather_variation: bool = True,
                 safety_penalty: float = 100.0, voltage_limits: Tuple[float, float] = (0.95, 1.05),
                 frequency_limits: Tuple[float, float] = (59.5, 60.5), power_flow_solver=None,
                 solver: str = "newton", tolerance: float = 1e-6, max_iterations: int = 50,
                 acceleration: float = 1.0, lanes: int = 0, load_noise: float = 0.1, repair="auto",
                 start_time: float = 0.0, env_id_offset: int = 0, auto_reset: bool = False,
                 copy_outputs: bool = False, record_noise: bool = False, obs_dtype=torch.float64,
                 obs_buffers: int = 0, keep_cycles: bool = False, **kwargs) -> None:
        # ``obs_dtype=torch.float32``: the kernels write the observation as fp32 (the type the reference
        # declares for its observation space, grid_env.py:346; every other output stays fp64).
        # ``obs_buffers=2``: two alternating observation buffers - step t writes the one that does not hold
        # observation t - 1, so observation t - 1 can be copied out on another stream while step t runs
        # (pipeline.HostStepper does).  Default: two for fp32, one for fp64.
        # ``keep_cycles=True`` (sweep solver): a repair keeps the lines that close a cycle (the shipped IEEE-34 /
        # IEEE-123 ties, SyntheticFeeder(connectivity > 0)) and the sweep restores the loops by compensation,
        # instead of dropping them (deviation D4-iii).  A meshed connected feeder passed with repair=False works too.
        # **kwargs are accepted and ignored, as the reference constructor does (base.py:84)
        if int(num_envs) < 1:
            raise InvalidConfigurationError("num_envs must be >= 1")
        if power_flow_solver is not None:
            # reference: GridEnvironment(power_flow_solver=<PowerFlowSolver>) (grid_env.py:169,198-206)
            solver = getattr(power_flow_solver, "method", solver)
            tolerance = getattr(power_flow_solver, "tolerance", tolerance)
            max_iterations = getattr(power_flow_solver, "max_iterations", max_iterations)
        self.device = _cuda_device(device)
        self.num_envs = int(num_envs)
        self.renewable_sources = list(renewable_sources or [])
        if solver not in nat.SOLVERS:
            raise InvalidConfigurationError(f"solver must be one of {sorted(nat.SOLVERS)}, got {solver!r}")
        if keep_cycles and solver != "sweep":
            raise InvalidConfigurationError("keep_cycles=True needs solver='sweep' (the tree-ordered Newton-Raphson takes "
                                            "radial feeders; the dense Newton-Raphson lives on the solver surface)")
        self.soa, self.feeder, lanes = _compile(feeder, self.renewable_sources, repair, solver, lanes, keep_cycles)
        self.timestep, self.episode_length = float(timestep), int(episode_length)
        self.stochastic_loads, self.weather_variation = bool(stochastic_loads), bool(weather_variation)
        self.safety_penalty = float(safety_penalty)
        self.voltage_limits, self.frequency_limits = tuple(voltage_limits), tuple(frequency_limits)
        self.solver, self.tolerance, self.max_iterations = solver, float(tolerance), int(max_iterations)
        self.start_time, self.env_id_offset = float(start_time), int(env_id_offset)
        self.auto_reset, self.copy_outputs = bool(auto_reset), bool(copy_outputs)

        self._native_feeder = NativeFeeder(self.soa, self.device)
        self.lib = self._native_feeder.lib
        scfg = nat.make_solver_cfg(solver, tolerance, max_iterations, acceleration, lanes)
        self._cfg = nat.make_env_cfg(timestep=timestep, episode_length=episode_length,
                                     stochastic_loads=stochastic_loads,
                                     weather_variation=weather_variation,
                                     voltage_limits=voltage_limits, frequency_limits=frequency_limits,
                                     safety_penalty=safety_penalty, load_noise=load_noise,
                                     solver_cfg=scfg, env_id_offset=self.env_id_offset)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(self.lib, self.lib.gfr_env_create(self._native_feeder.handle, self.num_envs,
                                                        C.byref(self._cfg), C.byref(h)))
        self._h = h
        B = self.num_envs
        self.obs_dim = self.lib.gfr_env_obs_dim(h)
        self.act_dim = self.lib.gfr_env_act_dim(h)
        self.noise_dim = self.lib.gfr_env_noise_dim(h)
        # reference spaces (grid_env.py:300-358): obs bounds are +-inf placeholders there too
        if isinstance(obs_dtype, str):
            obs_dtype = {"float64": torch.float64, "float32": torch.float32}.get(obs_dtype)
        if obs_dtype not in (torch.float64, torch.float32):
            raise InvalidConfigurationError("obs_dtype must be torch.float64 or torch.float32")
        self.obs_dtype = obs_dtype
        obs_buffers = int(obs_buffers) or (2 if obs_dtype == torch.float32 else 1)
        if obs_buffers not in (1, 2):
            raise InvalidConfigurationError("obs_buffers must be 1 or 2")
        if obs_buffers == 2 and (auto_reset or copy_outputs):
            obs_buffers = 1           # those modes hand out clones anyway
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(self.obs_dim,),
                                     dtype=np.float32 if obs_dtype == torch.float32 else np.float64)
        self.action_space = Box(low=-1.0, high=1.0, shape=(self.act_dim,), dtype=np.float64)

        dev = self.device
        f64 = dict(dtype=torch.float64, device=dev)
        # the observation buffers are torch tensors the library writes into (gfr_env_bind_obs_buffers)
        self._obs_bufs = [torch.empty(B, self.obs_dim, dtype=obs_dtype, device=dev) for _ in range(obs_buffers)]
        nat.check(self.lib, self.lib.gfr_env_bind_obs_buffers(
            h, self._obs_bufs[0].data_ptr(), self._obs_bufs[1].data_ptr() if obs_buffers == 2 else None,
            nat.OBS_F32 if obs_dtype == torch.float32 else nat.OBS_F64, self._stream()))
        self._obs_ptr = {t.data_ptr(): t for t in self._obs_bufs}
        u8 = dict(dtype=torch.uint8, device=dev)
        i32 = dict(dtype=torch.int32, device=dev)
        self._out = dict(
            reward=torch.zeros(B, **f64), terminated=torch.zeros(B, **u8), truncated=torch.zeros(B, **u8),
            error=torch.zeros(B, **u8), converged=torch.zeros(B, **u8), iterations=torch.zeros(B, **i32),
            max_voltage=torch.ones(B, **f64), min_voltage=torch.ones(B, **f64), losses=torch.zeros(B, **f64),
            max_mismatch=torch.zeros(B, **f64), violations=torch.zeros(B, 4, **u8),
            violation_count=torch.zeros(B, **i32), current_step=torch.zeros(B, **i32),
            episode_reward=torch.zeros(B, **f64),
            noise_used=torch.zeros(B, self.noise_dim, **f64) if record_noise else None)
        self._step_out = nat.StepOut(*[(self._out[k].data_ptr() if self._out[k] is not None else None)
                                       for k, _ in nat.StepOut._fields_])
        self._bool = {k: self._out[k].view(torch.bool) for k in
                      ("terminated", "truncated", "error", "converged", "violations")}
        self._done = torch.zeros(B, **u8)
        self._seeds = None

    # ------------------------------------------------------------------ plumbing
    @property
    def _obs(self) -> torch.Tensor:
        """The buffer holding the latest observation (the library alternates between the bound ones)."""
        if len(self._obs_bufs) == 1:
            return self._obs_bufs[0]
        return self._obs_ptr[self.lib.gfr_env_obs_current(self._h)]

    def _stream(self) -> int:
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self) -> None:
        if getattr(self, "_h", None):
            self.lib.gfr_env_destroy(self._h)
            self._h = None
        if getattr(self, "_native_feeder", None) is not None:
            self._native_feeder.close()

    def __del__(self) -> None:
        try:
            self.close()
        except Exception:
            pass

    def launch_info(self) -> Dict[str, int]:
        lanes, threads, grid, smem = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int64()
        nat.check(self.lib, self.lib.gfr_env_launch_info(self._h, C.byref(lanes), C.byref(threads),
                                                         C.byref(grid), C.byref(smem)))
        return dict(lanes=lanes.value, threads=threads.value, grid=grid.value, smem_bytes=smem.value)

    def _as_device(self, x, shape, dtype, what) -> Optional[torch.Tensor]:
        if x is None:
            return None
        t = torch.as_tensor(x) if not isinstance(x, torch.Tensor) else x
        shape = tuple(int(v) for v in shape)
        if tuple(t.shape) != shape:
            # only an unambiguous 1-D form is accepted besides the exact shape: [B] when the row has one
            # entry, [A] when there is one instance.  Anything else (a transposed [A, B], a flat [B * A])
            # would scramble the per-instance rows, and the reference raises on a length mismatch too
            # (utils/parallel_environment.py:74-75)
            if len(shape) == 2 and t.dim() == 1 and t.numel() == shape[0] * shape[1] and 1 in shape:
                t = t.reshape(shape)
            else:
                raise InvalidActionError(f"{what} must have shape {shape}, got {tuple(t.shape)}")
        if t.dtype != dtype or t.device != self.device or not t.is_contiguous():
            t = t.to(device=self.device, dtype=dtype, non_blocking=True).contiguous()
        return t

    # ------------------------------------------------------------------ API
    def reset(self, seed: Optional[int] = None, options: Optional[Dict[str, Any]] = None, *,
              seeds=None, mask=None, noise=None):
        """``seed``: instance i gets ``seed + env_id_offset + i`` (results do not depend on how the
        instances are sharded).  ``seeds`` [B] uint64 overrides.  ``mask`` [B] selects instances.
        ``noise`` [B,4] replays the four weather draws instead of the in-kernel Philox stream."""
        B = self.num_envs
        if seeds is None and seed is not None:
            seeds = global_seeds(seed, self.env_id_offset, B, self.device)
        if seeds is not None:
            seeds = torch.as_tensor(seeds)
            if seeds.dtype == torch.uint64:
                seeds = seeds.view(torch.int64)
            seeds = self._as_device(seeds, (B,), torch.int64, "seeds")
        mask_t = None
        if mask is not None:
            mask_t = torch.as_tensor(mask)
            if mask_t.dtype == torch.bool:
                mask_t = mask_t.to(torch.uint8)
            mask_t = self._as_device(mask_t, (B,), torch.uint8, "mask")
        noise_t = self._as_device(noise, (B, 4), torch.float64, "reset noise")
        start = float((options or {}).get("start_time", self.start_time))
        nat.check(self.lib, self.lib.gfr_env_reset(
            self._h, seeds.data_ptr() if seeds is not None else None,
            mask_t.data_ptr() if mask_t is not None else None,
            noise_t.data_ptr() if noise_t is not None else None, start, self._stream()))
        # info of the instances that were reset: zeros, voltages at 1.0 (one launch for all the output arrays)
        nat.check(self.lib, self.lib.gfr_env_reset_outputs(
            self._h, mask_t.data_ptr() if mask_t is not None else None, C.byref(self._step_out), self._stream()))
        obs = self._obs.clone() if self.copy_outputs else self._obs
        return obs, self._info()

    def step_outputs_into(self, reward: torch.Tensor, terminated: torch.Tensor, truncated: torch.Tensor):
        """A copy of the step's output table with reward (fp64 [B]) and the two done flags (uint8 [B]) pointed
        at the caller's device buffers: ``step(actions, _step_out=...)`` then writes THOSE instead of the
        environment's own (``gfr_env_step`` takes the output table per call).  ``pipeline.HostStepper`` alternates
        two such sets, so a step's results can leave for the host while the next step already runs."""
        B = self.num_envs
        for t, dt in ((reward, torch.float64), (terminated, torch.uint8), (truncated, torch.uint8)):
            if t.dtype != dt or t.numel() != B or not t.is_contiguous() or t.device != self.device:
                raise ValueError("reward: fp64 [B], terminated / truncated: uint8 [B], contiguous, on the environment's device")
        out = nat.StepOut(*[getattr(self._step_out, k) for k, _ in nat.StepOut._fields_])
        out.reward, out.terminated, out.truncated = reward.data_ptr(), terminated.data_ptr(), truncated.data_ptr()
        return out

    def step(self, actions, noise=None, _step_out=None):
        """``actions`` [B, A] float64 (any device / dtype is converted; a host array costs one
        H2D copy).  ``noise`` [B, 4 + L] replays the reference's random draws (parity mode).
        ``_step_out`` (from ``step_outputs_into``): this step's reward / done flags go to the caller's buffers and
        the returned reward / flag tensors are NOT updated (not with ``auto_reset`` / ``copy_outputs``)."""
        B = self.num_envs
        act = self._as_device(actions, (B, self.act_dim), torch.float64, "actions")
        noise_t = self._as_device(noise, (B, self.noise_dim), torch.float64, "noise")
        if _step_out is not None and (self.auto_reset or self.copy_outputs):
            raise ValueError("_step_out cannot be combined with auto_reset / copy_outputs")
        nat.check(self.lib, self.lib.gfr_env_step(
            self._h, act.data_ptr(), noise_t.data_ptr() if noise_t is not None else None,
            C.byref(self._step_out if _step_out is None else _step_out), self._stream()))
        o, b = self._out, self._bool
        reward, terminated, truncated = o["reward"], b["terminated"], b["truncated"]
        info = self._info()
        obs = self._obs
        if self.copy_outputs or self.auto_reset:
            obs = obs.clone()
            reward, terminated, truncated = reward.clone(), terminated.clone(), truncated.clone()
            info = {k: (v.clone() if isinstance(v, torch.Tensor) else v) for k, v in info.items()}
        if self.auto_reset:
            torch.bitwise_or(o["terminated"], o["truncated"], out=self._done)
            info["final_observation"] = obs
            nat.check(self.lib, self.lib.gfr_env_reset(self._h, None, self._done.data_ptr(), None,
                                                       self.start_time, self._stream()))
            obs = torch.where(self._done.bool()[:, None], self._obs, obs)
        return obs, reward, terminated, truncated, info

    def _info(self) -> Dict[str, torch.Tensor]:
        # same keys as the reference's info (grid_env.py:610-617, base.py:169-176); its
        # ``constraint_violations`` dict of four bools becomes a [B, 4] tensor
        o, b = self._out, self._bool
        return {"power_flow_converged": b["converged"], "max_voltage": o["max_voltage"],
                "min_voltage": o["min_voltage"], "total_losses": o["losses"],
                "constraint_violations": b["violations"],
                "constraint_violation_count": o["violation_count"], "iterations": o["iterations"],
                "current_step": o["current_step"], "episode_reward": o["episode_reward"],
                "error": b["error"], "max_mismatch": o["max_mismatch"],
                "timestep": self.timestep}

    def get_observation(self) -> torch.Tensor:
        return self._obs

    @property
    def noise_used(self) -> Optional[torch.Tensor]:
        return self._out["noise_used"]

    @property
    def current_step(self) -> torch.Tensor:
        return self._out["current_step"]

    @property
    def episode_reward(self) -> torch.Tensor:
        return self._out["episode_reward"]

    @property
    def constraint_violations(self) -> torch.Tensor:
        return self._out["violation_count"]

    def sample_actions(self, generator: Optional[torch.Generator] = None) -> torch.Tensor:
        """U(-1, 1) actions on the device (the batched form of ``action_space.sample()``)."""
        return torch.rand(self.num_envs, self.act_dim, dtype=torch.float64, device=self.device,
                          generator=generator) * 2.0 - 1.0

    # ------------------------------------------------------------------ checkpoint / resume
    def state_dict(self) -> Dict[str, torch.Tensor]:
        n = self.lib.gfr_env_state_bytes(self._h)
        rec = torch.empty(n // 8, dtype=torch.float64, device=self.device)
        nat.check(self.lib, self.lib.gfr_env_state_get(self._h, rec.data_ptr(), self._stream()))
        sd = {"records": rec, "observation": self._obs.clone()}
        sd.update({k: v.clone() for k, v in self._out.items() if v is not None})
        return sd

    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        rec = self._as_device(sd["records"], (self.lib.gfr_env_state_bytes(self._h) // 8,),
                              torch.float64, "records")
        nat.check(self.lib, self.lib.gfr_env_state_set(self._h, rec.data_ptr(), self._stream()))
        for t in self._obs_bufs:
            t.copy_(sd["observation"])
        for k, v in self._out.items():
            if v is not None and k in sd:
                v.copy_(sd[k])

    # ------------------------------------------------------------------ episode statistics
    def episode_stats(self, reduce: bool = False) -> Dict[str, float]:
        """Sums over this rank's instances; ``reduce=True`` adds one NCCL all-reduce of an fp64[8]
        vector over the default process group (the only collective anywhere near this path)."""
        o = self._out
        v = torch.stack([o["reward"].sum(), o["episode_reward"].sum(),
                         o["converged"].sum(dtype=torch.float64), o["iterations"].sum(dtype=torch.float64),
                         o["violation_count"].sum(dtype=torch.float64),
                         (o["terminated"] | o["truncated"]).sum(dtype=torch.float64),
                         o["error"].sum(dtype=torch.float64),
                         torch.tensor(float(self.num_envs), dtype=torch.float64, device=self.device)])
        if reduce:
            all_reduce_stats(v)
        return stats_dict(v)

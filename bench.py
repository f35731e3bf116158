#!/usr/bin/env python
"""bench.py - env-steps/s of the fused batched GridEnvironment.step on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload ieee123|ieee13|ieee34] [--lanes L]
  python bench.py --impl reference ...      # the CPU oracle (port of the reference path) on host cores

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one batch of
instances: load / weather / battery update, Newton-Raphson (or sweep) load flow, constraints,
reward, observation - one kernel launch.  Inputs (actions) are resident in HBM for `value`;
`e2e` repeats the measurement through the public API with pinned HOST action buffers and a
device->host read of reward + done flags every step.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# workload name -> (feeder spec, per-GPU instances, solver, tolerance, max_iterations)
WORKLOADS = {
    # BASELINE.json configs[3]: the configuration the metric is quoted on; 131,072 per GPU
    # (x8 = 1,048,576 instances), weak scaling
    "ieee123": ("ieee123", 131072, "newton", 1e-6, 50),
    "ieee13": ("ieee13", 65536, "sweep", 1e-8, 50),       # configs[1]
    "ieee34": ("ieee34", 262144, "sweep", 1e-8, 50),      # configs[2]
    "ieee13_newton": ("ieee13", 65536, "newton", 1e-6, 50),
    "ieee34_newton": ("ieee34", 262144, "newton", 1e-6, 50),
    "ieee123_sweep": ("ieee123", 131072, "sweep", 1e-8, 50),
    # configs[4]: synthetic 1,000-bus radial feeder, 16,384 instances over 8 GPUs
    "synthetic1000": ("synthetic1000", 2048, "newton", 1e-6, 50),
}
ENV_KW = dict(timestep=1.0, renewable_sources=["solar", "wind"], stochastic_loads=True,
              weather_variation=True)
START_TIME = 12 * 3600.0     # daylight, so the solar branch is exercised


def make_feeder(spec):
    import grid_fed_rl_b200 as m
    if spec == "synthetic1000":
        # ScalableFeeder(1000)'s parameters (synthetic.py:236-252) but radial; loads scaled by 0.03 so that
        # the feeder is inside its loadability (sum P = 0.43 pu)
        cfg = m.NetworkConfig(num_buses=1000, connectivity=0.0, load_probability=0.9, dg_probability=0.4,
                              min_load_kw=20, max_load_kw=300, line_length_range=(0.05, 1.5))
        f = m.repair_topology(m.SyntheticFeeder(cfg, seed=1000))
        for ld in f.loads:
            ld.base_power *= 0.03; ld.active_power *= 0.03; ld.reactive_power *= 0.03
        return f
    f = {"ieee13": m.IEEE13Bus, "ieee34": lambda: m.IEEE34Bus(seed=0),
         "ieee123": lambda: m.IEEE123Bus(seed=0)}[spec]()
    return m.repair_topology(f)


def algorithmic_bytes(soa):
    """reads actions + state, writes state + observation + outputs, per env-step."""
    return 8 * soa.obs_dim + 8 * soa.act_dim + 32 * soa.n_bat + 199


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop_evt.wait(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2] if s else None, "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s)}


def physical_gpu_index(local_index):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_index])
        except Exception:
            return local_index
    return local_index


# ----------------------------------------------------------------------------- CPU oracle legs

def _cpu_worker(args):
    spec, B, steps, solver_tol, seed, threads = args
    import numpy as np
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=threads)
    except Exception:
        import contextlib
        ctx = contextlib.nullcontext()
    from oracle import port
    f = make_feeder(spec)
    with ctx:
        env = port.PortEnv(f, B, tolerance=solver_tol, **ENV_KW)
        rs = np.random.RandomState(seed)
        nz0 = np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3))], axis=1)
        env.reset(nz0, start_time=START_TIME)
        acts = [rs.uniform(-1, 1, size=(B, env.A)) for _ in range(steps + 1)]
        nzs = [np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3 + env.L))], axis=1)
               for _ in range(steps + 1)]
        env.step(acts[0], nzs[0])                       # warm-up (allocations, BLAS threads)
        t0 = time.perf_counter()
        conv = 0
        for t in range(steps):
            conv += int(env.step(acts[t + 1], nzs[t + 1])["converged"].sum())
        dt = time.perf_counter() - t0
    return B * steps, dt, conv


def cpu_baseline(spec, tol, B=32, steps=2):
    """The oracle (numpy port of the reference's step + dense Newton-Raphson) on ONE host thread."""
    n, dt, conv = _cpu_worker((spec, B, steps, tol, 0, 1))
    return {"value": n / dt, "unit": "env-steps/s", "cores": 1, "kind": "port",
            "sample": f"oracle/port.py PortEnv, {spec}, {B} instances x {steps} steps, dense NR tol {tol:g}, "
                      f"1 thread, {dt:.1f} s", "converged_frac": conv / n}


def _ref_server(conn, spec, B, tol, seed):
    """A warm oracle worker: one PortEnv over its own shard, one step per 'go'."""
    import numpy as np
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=1)
    except Exception:
        pass
    from oracle import port
    env = port.PortEnv(make_feeder(spec), B, tolerance=tol, **ENV_KW)
    rs = np.random.RandomState(seed)
    env.reset(np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3))], axis=1),
              start_time=START_TIME)
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg == "stop":
            break
        act = rs.uniform(-1, 1, size=(B, env.A))
        nz = np.concatenate([rs.random_sample((B, 1)), rs.standard_normal((B, 3 + env.L))], axis=1)
        out = env.step(act, nz)
        conn.send(int(out["converged"].sum()))


def run_reference(args):
    """--impl reference: the CPU oracle (numpy port of the reference's step with its dense
    Newton-Raphson; the reference itself is Python and does not travel to the GPU box) on every
    host core: one single-threaded process per core, each stepping its own shard of instances."""
    import multiprocessing as mp
    spec, B_gpu, solver, tol, max_it = WORKLOADS[args.workload]
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    per_proc = {"ieee123": 48, "ieee34": 384, "ieee13": 2048, "synthetic1000": 1}[spec]
    ptol = tol if solver == "newton" else 1e-8
    steps, warm = max(1, args.steps), max(0, args.warmup)
    ctx = mp.get_context("spawn")
    procs = []
    for i in range(cores):
        parent, child = ctx.Pipe()
        pr = ctx.Process(target=_ref_server, args=(child, spec, per_proc, ptol, 100 + i), daemon=True)
        pr.start()
        procs.append((pr, parent))
    for _, c in procs:
        assert c.recv() == "ready"

    def one_step():
        for _, c in procs:
            c.send("go")
        return sum(c.recv() for _, c in procs)

    for _ in range(warm):
        one_step()
    t0 = time.perf_counter()
    conv = 0
    for _ in range(steps):
        conv += one_step()
    wall = time.perf_counter() - t0
    for pr, c in procs:
        c.send("stop")
    for pr, _ in procs:
        pr.join(timeout=5)
    total = per_proc * cores * steps
    value = total / wall
    line = {"metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": warm, "ms_per_step": 1e3 * wall / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": f"{spec} GridEnvironment.step, CPU oracle (oracle/port.py: dense "
                                   f"Newton-Raphson tol {ptol:g}), bounded sample of {per_proc * cores} "
                                   f"instances per step", "instances_per_step": per_proc * cores},
            "converged_frac": conv / total,
            "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": cores, "kind": "port",
                             "sample": f"{cores} single-threaded processes x {per_proc} instances x {steps} steps"},
            "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist
    import grid_fed_rl_b200 as m
    from grid_fed_rl_b200 import _native
    from grid_fed_rl_b200.distributed import max_over_ranks

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); use --impl reference for the CPU oracle")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    spec, B, solver, tol, max_it = WORKLOADS[args.workload]
    if args.envs:
        B = args.envs
    lib = _native.load_library()
    feeder = make_feeder(spec)
    env = m.BatchedGridEnvironment(feeder, B, device=dev, solver=solver, tolerance=tol,
                                   max_iterations=max_it, lanes=args.lanes, repair=False,
                                   start_time=START_TIME, env_id_offset=rank * B, **ENV_KW)
    env.reset(seed=args.seed)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1234 + rank)
    R = 4
    actions = [env.sample_actions(gen) for _ in range(R)]
    A = env.act_dim
    K, W = args.steps, max(args.warmup, 3)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(step_fn, k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for i in range(k):
            step_fn(i)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1), dev)     # a timed region counts as its slowest rank

    # ---- device-resident arm
    conv_acc = torch.zeros((), dtype=torch.float64, device=dev)
    it_acc = torch.zeros((), dtype=torch.float64, device=dev)

    def dev_step(i):
        env.step(actions[i % R])

    for i in range(W):
        dev_step(i)
    sampler = ClockSampler(physical_gpu_index(local))
    sampler.start()
    l0 = lib.gfr_launch_count()
    ms = timed(dev_step, K)
    launches = lib.gfr_launch_count() - l0
    info = env._info()
    conv_frac = float(info["power_flow_converged"].double().mean().item())
    mean_it = float(info["iterations"].double().mean().item())
    if world > 1:
        t = torch.tensor([conv_frac, mean_it], dtype=torch.float64, device=dev)
        dist.all_reduce(t)
        conv_frac, mean_it = (t / world).tolist()
    value = B * world * K / (ms * 1e-3)
    if args.kernel_only:                 # tuning aid (tools/): the device-resident arm only
        sampler.stop()
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms / K, "mean_iterations": mean_it,
                              "converged_frac": conv_frac,
                              "config": {"launch": env.launch_info(), "instances_per_gpu": B}}), flush=True)
        env.close()
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end-to-end arm: every step copies ITS actions from pinned host memory (H2D) and its reward +
    #      done flags back (D2H), and the host reads each step's result.  Measured twice through the
    #      public API: serial (copy, step, copy, synchronise) and as HostStepper's depth-2 pipeline
    #      (the next step's action copy overlaps the running kernel on a second stream).
    from grid_fed_rl_b200.pipeline import HostStepper
    host_act = [a.cpu().pin_memory() for a in actions]

    def e2e_run(depth, observations=False, k_steps=None):
        stepper = HostStepper(env, depth=depth, observations=observations)
        checksum = [0.0]
        K = k_steps or args.steps

        def run(k):
            for i in range(k):
                stepper.submit(host_act[i % R])
                if i + 1 >= depth:
                    checksum[0] += float(stepper.result()["reward"][0])      # the host consumes the result
            while stepper._pending:
                checksum[0] += float(stepper.result()["reward"][0])

        run(3)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        run(K)
        ev1.record()
        barrier()
        return max_over_ranks(ev0.elapsed_time(ev1), dev), stepper

    # the same with every step's observation [B, D] copied to the host too (PCIe bound; fewer steps)
    K_obs = max(2, min(K, 10))
    ms_obs, stepper_obs = e2e_run(2, observations=True, k_steps=K_obs)
    obs_value = B * world * K_obs / (ms_obs * 1e-3)
    obs_d2h = stepper_obs.d2h_bytes_per_step
    del stepper_obs
    ms_serial, stepper = e2e_run(1)
    ms_e2e, stepper = e2e_run(2)
    clocks = sampler.stop()          # sampled across the timed regions
    e2e_value = B * world * K / (ms_e2e * 1e-3)
    e2e_serial = B * world * K / (ms_serial * 1e-3)

    fp64_peak = ctypes.c_double(0.0)
    if rank == 0:
        _native.check(lib, lib.gfr_fp64_peak(local, ctypes.byref(fp64_peak)))
    if rank == 0:
        soa = env.soa
        abytes = algorithmic_bytes(soa)
        peak, how = hbm_peak()
        achieved = abytes * B / (ms / K * 1e-3) / 1e9          # GB/s of the one kernel a step launches
        li = env.launch_info()
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms / K, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{spec} fused GridEnvironment.step, {solver} load flow tol {tol:g}, "
                                   f"{B} instances per GPU, in-kernel Philox noise, random U(-1,1) policy",
                       "instances_per_gpu": B, "n_bus": soa.n_bus, "obs_dim": soa.obs_dim,
                       "act_dim": soa.act_dim, "solver": solver, "parallelism": f"shard{world}",
                       "l2": f"per-step working set {B * abytes / 1e6:.0f} MB > 126 MB L2, no flush needed",
                       "launch": li},
            "converged_frac": conv_frac, "converged_solves_per_s": value * conv_frac,
            "mean_iterations": mean_it,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": TRAFFIC_BYTES.get((args.workload, B)),
                         "peak_source": how, "algorithmic_bytes_per_env_step": abytes,
                         "kernel": f"step_kernel<{li['lanes']},{solver}>", "kernel_ms": ms / K,
                         "note": "not HBM bound: ncu shows the L1 / shared-memory data pipe at 70 %, issue slots 56 %, "
                                 "FP64 pipe 28 %, DRAM 7.5 % (DESIGN.md section 5, profiles/); see also `fp64`"},
            # FP64 side (the contract's roofline bounds are hbm | tensor; this kernel is neither): FP64 pipe
            # operations per env-step (ncu, profiles/: executed DFMA / DMUL / DADD / DSETP thread
            # instructions / instances) x env-steps/s, against the DFMA issue rate measured on this
            # device just now (gfr_fp64_peak: TFLOP/s at 2 flop per DFMA, so ops/s = TFLOP/s / 2)
            "fp64": {"lane_ops_per_env_step": FP64_LANE_OPS.get(args.workload),
                     "achieved_gops": (FP64_LANE_OPS[args.workload] * value / world / 1e9)
                                      if args.workload in FP64_LANE_OPS else None,
                     "peak_gops": fp64_peak.value * 1e3 / 2.0, "peak_dfma_tflops": fp64_peak.value,
                     "frac": (FP64_LANE_OPS[args.workload] * value / world / 1e9 / (fp64_peak.value * 1e3 / 2.0))
                             if fp64_peak.value and args.workload in FP64_LANE_OPS else None},
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": stepper.h2d_bytes_per_step,
                    "d2h_bytes_per_step": stepper.d2h_bytes_per_step, "ms_per_step": ms_e2e / K,
                    "serial_value": e2e_serial, "serial_ms_per_step": ms_serial / K,
                    "with_observations": {"value": obs_value, "d2h_bytes_per_step": obs_d2h, "steps": K_obs,
                                          "gb_per_s_d2h": obs_d2h / (ms_obs / K_obs * 1e-3) / 1e9},
                    "note": "HostStepper (public API): pinned host actions in, reward + terminated + truncated "
                            "out, every step, each result read by the host; `value` = depth-2 pipeline (next "
                            "step's H2D overlaps the kernel), `serial_value` = copy-step-copy-sync; "
                            "observations stay in HBM for a device policy (`with_observations`: every "
                            "observation row copied to pinned host memory too - PCIe bound)"},
            "gpu_launches": int(launches), "clocks": clocks,
        }
        if world == 1 and not args.no_cpu:
            line["cpu_baseline"] = cpu_baseline(spec, tol if solver == "newton" else 1e-8,
                                                *{"ieee123": (256, 12), "ieee34": (2048, 12), "ieee13": (16384, 12), "synthetic1000": (2, 2)}[spec])
        print(json.dumps(line), flush=True)
    env.close()
    if world > 1:
        dist.destroy_process_group()


# FP64 lane operations per env-step: ncu smsp__sass_thread_inst_executed_op_{dfma,dmul,dadd}_pred_on summed over
# the launch / instances (profiles/r01_ncu_ieee123_v8.txt: (990 + 526 + 302) per cycle x 3.196 M cycles / 131072)
FP64_LANE_OPS = {"ieee123": 44300}

# dram__bytes_read.sum + dram__bytes_write.sum per launch of the step kernel, from the ncu --set full
# capture summarised under profiles/ (keyed by workload and instances per launch)
TRAFFIC_BYTES = {
    # profiles/r01_ncu_ieee123_v8.txt: 37.3 MB read + 985.1 MB written per launch of step_kernel<8, newton>.
    # The algorithmic bytes are 772.7 MB (of which the 2L static load columns of the observation, 199 MB, are
    # never rewritten); the rest of the writes are the solver's scratch (D^-1 U, D^-1 r: 61 MB live, 2.1 GB
    # written per launch) leaving the L2 under the observation stream
    # (final binary of the round, profiles/r01_ncu_ieee123_v10.txt: 36.9 MB + 991.1 MB)
    ("ieee123", 131072): 1028.0e6,
}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="ieee123", choices=sorted(WORKLOADS))
    ap.add_argument("--lanes", type=int, default=0)
    ap.add_argument("--envs", type=int, default=0, help="instances per GPU (default: the workload's)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--kernel-only", action="store_true", help="tuning aid: time the device-resident arm only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()

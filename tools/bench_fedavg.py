#!/usr/bin/env python
"""FedAvg over per-GPU clients as ONE weighted all-reduce: the reference's
``FedAvgAggregator.aggregate`` (federated/core.py:233-258: sum_i (n_i / N) p_i over a Python list of clients)
with the clients living on the ranks of a process group.  Run under torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/bench_fedavg.py

Parameter sets: a CQL client as the reference builds it for the IEEE-123 environment (offline.py:14-103:
actor 692 -> 256 x 3 -> 16, twin critics 700 -> 256 x 3 -> 1, target critics), and flat buckets of 64 / 256 MB
to show the bus bandwidth NVLink / NVSwitch gives once the message is large.  Prints one JSON line (rank 0):
time per aggregation, algorithm bandwidth (bytes reduced / time) and bus bandwidth 2 (N - 1) / N x that."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from grid_fed_rl_b200.distributed import fedavg_all_reduce, max_over_ranks  # noqa: E402


def cql_parameters(state_dim, action_dim, hidden, device, rank):
    g = torch.Generator(device=device).manual_seed(100 + rank)

    def mlp(prefix, dims):
        out = {}
        for i in range(len(dims) - 1):
            out[f"{prefix}.{2 * i}.weight"] = torch.randn(dims[i + 1], dims[i], device=device, generator=g) * 0.05
            out[f"{prefix}.{2 * i}.bias"] = torch.randn(dims[i + 1], device=device, generator=g) * 0.05
        return out
    p = {}
    p.update(mlp("actor", [state_dim] + hidden + [2 * action_dim]))
    for net in ("critic.q1", "critic.q2", "target_critic.q1", "target_critic.q2"):
        p.update(mlp(net, [state_dim + action_dim] + hidden + [1]))
    return p


def main():
    world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n_samples = 1000.0 * (rank + 1)
    cases = {"cql_ieee123": cql_parameters(692, 8, [256, 256, 256], dev, rank),
             "flat_64MB": {"w": torch.full((16 * 1024 * 1024,), float(rank + 1), device=dev)},
             "flat_256MB": {"w": torch.full((64 * 1024 * 1024,), float(rank + 1), device=dev)}}
    res = {}
    for name, params in cases.items():
        nbytes32 = sum(v.numel() * 4 for v in params.values())
        wire = (sum(v.numel() for v in params.values()) + 1) * 4          # the bucket travels in the parameters' fp32
        out = fedavg_all_reduce(params, n_samples)                        # warm-up + correctness
        if name.startswith("flat"):
            tot = sum(1000.0 * (r + 1) for r in range(world))
            want = sum(1000.0 * (r + 1) * (r + 1) for r in range(world)) / tot
            assert abs(float(out["w"][0]) - want) < 1e-5 and abs(float(out["w"][-1]) - want) < 1e-5
        for _ in range(3):
            fedavg_all_reduce(params, n_samples)
        K = 20 if "256" not in name else 8
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(K):
            fedavg_all_reduce(params, n_samples)
        ev1.record()
        torch.cuda.synchronize()
        ms = max_over_ranks(ev0.elapsed_time(ev1), dev) / K
        alg = wire / (ms * 1e-3) / 1e9
        res[name] = {"parameters": sum(v.numel() for v in params.values()), "tensors": len(params),
                     "param_bytes_fp32": nbytes32, "bucket_bytes": wire, "ms_per_aggregation": ms,
                     "algbw_gb_s": alg, "busbw_gb_s": alg * 2 * (world - 1) / world if world > 1 else None}
    if rank == 0:
        print(json.dumps({"what": "fedavg_all_reduce (weighted FedAvg as one NCCL all-reduce, pack + reduce + unpack timed)",
                          "n_gpus": world, "nvlink5_per_direction_gb_s": 900, "results": res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""What the HOST gives N ranks copying device buffers to pinned memory at the same time - the ceiling of the
"host-side policy" arm of bench.py (`e2e.with_observations*`: every observation row crosses PCIe).  Run under
torchrun, one rank per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
      tools/d2h_scaling.py

Every rank copies a 256 MB device buffer to pinned host memory 20 times, (a) alone (the others wait), (b) all ranks
together; with the rank pinned to the CPUs next to its GPU before the pinned buffer is allocated (what bench.py does)
and, for comparison, unpinned.  The same for host -> device.  Rank 0 prints one JSON line: GB/s per rank alone, per
rank together, the sum, and where the GPUs hang (PCI bus ids, NUMA nodes, CPU affinity sizes)."""
import json
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import bind_to_gpu_numa_node, physical_gpu_index  # noqa: E402

MB = 256
REPS = 20


def timed_copies(dst, src, reps):
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    ev0.record()
    for _ in range(reps):
        dst.copy_(src, non_blocking=True)
    ev1.record()
    torch.cuda.synchronize()
    return src.numel() * src.element_size() * reps / (ev0.elapsed_time(ev1) * 1e-3) / 1e9


def gather(value, world, dev):
    t = torch.zeros(world, dtype=torch.float64, device=dev)
    t[dist.get_rank() if world > 1 else 0] = value
    if world > 1:
        dist.all_reduce(t)
    return [round(v, 2) for v in t.tolist()]


def measure(world, rank, dev, host, devbuf, direction):
    dst, src = (host, devbuf) if direction == "d2h" else (devbuf, host)
    timed_copies(dst, src, 3)
    alone = 0.0
    for r in range(world):                       # one rank at a time
        if world > 1:
            dist.barrier()
        if r == rank:
            alone = timed_copies(dst, src, REPS)
    if world > 1:
        dist.barrier()
    together = timed_copies(dst, src, REPS)      # every rank at once
    if world > 1:
        dist.barrier()
    return gather(alone, world, dev), gather(together, world, dev)


def main():
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = MB * 1024 * 1024 // 4
    devbuf = torch.ones(n, dtype=torch.float32, device=dev)
    out = {"tool": "d2h_scaling", "n_gpus": world, "buffer_mb": MB, "copies": REPS}
    all_cpus = len(os.sched_getaffinity(0))
    # unpinned first (default placement of the pinned buffer), then bound to the GPU's CPUs as bench.py does
    for label in ("unbound", "numa_bound"):
        bound = bind_to_gpu_numa_node(physical_gpu_index(local)) if label == "numa_bound" else 0
        host = torch.empty(n, dtype=torch.float32, pin_memory=True)
        host.fill_(0.0)                          # touch every page on this rank's CPUs
        res = {}
        for direction in ("d2h", "h2d"):
            alone, together = measure(world, rank, dev, host, devbuf, direction)
            res[direction] = {"alone_gb_s": alone, "together_gb_s": together, "together_sum_gb_s": round(sum(together), 1),
                              "alone_sum_gb_s": round(sum(alone), 1)}
        res["cpus"] = gather(float(bound if bound else all_cpus), world, dev)
        out[label] = res
        del host
    # where the GPUs hang
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(physical_gpu_index(local))
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        node = -1
        p = f"/sys/bus/pci/devices/{bus[-12:].lower()}/numa_node"
        if os.path.exists(p):
            node = int(open(p).read().strip())
        gen = pynvml.nvmlDeviceGetCurrPcieLinkGeneration(h)
        width = pynvml.nvmlDeviceGetCurrPcieLinkWidth(h)
    except Exception:
        bus, node, gen, width = "?", -1, 0, 0
    out["numa_node_of_gpu"] = [int(v) for v in gather(float(node), world, dev)]
    out["pcie_gen"] = [int(v) for v in gather(float(gen), world, dev)]
    out["pcie_width"] = [int(v) for v in gather(float(width), world, dev)]
    nodes = [d for d in os.listdir("/sys/devices/system/node")] if os.path.isdir("/sys/devices/system/node") else []
    out["host_numa_nodes"] = len([d for d in nodes if d.startswith("node")])
    out["host_cpus"] = os.cpu_count()
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

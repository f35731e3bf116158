"""N > 1 host logic on CPU: world_size 2, gloo, 127.0.0.1 (the step path itself has no collective)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from grid_fed_rl_b200 import distributed as gd


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = gd.shard_range(total, rank, world)
        seeds = gd.global_seeds(7, lo, hi - lo)
        # every rank's "episode statistics" over its own shard
        local = torch.zeros(len(gd.STAT_KEYS), dtype=torch.float64)
        local[0] = float(seeds.sum())            # stands in for a per-instance quantity
        local[-1] = hi - lo
        red = gd.all_reduce_stats(local.clone())
        slowest = gd.max_over_ranks(10.0 + rank)
        # FedAvg over ranks: rank r holds parameters filled with r + 1 and has (r + 1) * 10 samples
        params = {"w": torch.full((3, 2), float(rank + 1)), "b": torch.full((4,), float(rank + 1), dtype=torch.float64)}
        avg = gd.fedavg_all_reduce(params, (rank + 1) * 10)
        out.put((rank, lo, hi, seeds.tolist(), red.tolist(), slowest,
                 {k: (v.tolist(), str(v.dtype)) for k, v in avg.items()}))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("total", [10, 11])
def test_shards_cover_and_reduce(total):
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, total, out)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # contiguous, disjoint, complete; keys are seed + global id regardless of the sharding
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == total
    assert res[0][3] + res[1][3] == [7 + i for i in range(total)]
    for r in res:
        assert r[4][-1] == total and r[4][0] == sum(7 + i for i in range(total))
        assert r[5] == 11.0
        # sum_i (n_i / N) p_i = (10 * 1 + 20 * 2) / 30 (federated/core.py:233-258)
        w, wt = r[6]["w"]
        b, bt = r[6]["b"]
        assert wt == "torch.float32" and bt == "torch.float64"
        assert all(abs(x - 50.0 / 30.0) < 1e-6 for row in w for x in row) and len(w) == 3
        assert all(abs(x - 50.0 / 30.0) < 1e-12 for x in b)


def test_shard_range_properties():
    for total in (1, 7, 131072, 1_048_576):
        for world in (1, 2, 3, 8):
            spans = [gd.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        gd.shard_range(10, 2, 2)
    # single process: reductions are the identity
    v = torch.arange(len(gd.STAT_KEYS), dtype=torch.float64)
    assert torch.equal(gd.all_reduce_stats(v.clone()), v) and gd.max_over_ranks(3.5) == 3.5
    p = {"w": torch.arange(6.0).reshape(2, 3)}
    assert torch.equal(gd.fedavg_all_reduce(p, 5)["w"], p["w"]) and gd.fedavg_all_reduce(p, 0) == {}

"""Multi-GPU logic on CPU: world_size-2 gloo run of the chunk sharding used by bench.py --gpus N
(independent chunks per rank, no data-path collective; only sizes and times are reduced)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from datagen import bench_words


def shard_plan(total_chunks, world):
    """chunk c belongs to rank c // ceil(total/world): contiguous slices, as SURVEY section 8e."""
    per = (total_chunks + world - 1) // world
    return [list(range(r * per, min((r + 1) * per, total_chunks))) for r in range(world)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    plan = shard_plan(5, world)
    mine = plan[rank]
    # each rank "compresses" its own chunks: here a deterministic stand-in size per chunk
    sizes = torch.tensor([int(bench_words(4096, start=c * 1024).sum()) % 100000 for c in mine] + [0] * (3 - len(mine)), dtype=torch.int64)
    gathered = [torch.zeros(3, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, sizes)
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        out.put(([g.tolist() for g in gathered], t.item(), plan))
    dist.destroy_process_group()


def test_shard_plan_covers_everything():
    for total in (1, 4, 5, 32):
        for world in (1, 2, 4, 8):
            plan = shard_plan(total, world)
            flat = [c for p in plan for c in p]
            assert flat == list(range(total))


def test_two_rank_gloo():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    gathered, tmax, plan = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert tmax == 2.0                                  # max over ranks
    assert plan == [[0, 1, 2], [3, 4]]
    want = [int(bench_words(4096, start=c * 1024).sum()) % 100000 for c in range(5)]
    assert gathered[0] == want[:3] and gathered[1][:2] == want[3:]

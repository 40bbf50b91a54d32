"""Multi-GPU logic on CPU: a world_size-2 gloo run of the PRODUCT's sharding code
(c-blosc_b200/sharding.py: scatter slices -> per-rank frame compress -> gather-v, and the mirror)
with the library's host code running over the emulated backend.  The root checks that every chunk
of every rank's frame is byte-identical to the oracle's chunk for that slice of the buffer."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMU = os.path.join(ROOT, "tests", "emu", "_build", "libblosc_b200_emu.so")
TOTAL, CHUNK, TS = 5 * 40000 + 1234, 40000, 4          # 6 chunks: ranks get 3 + 3, the last one short


def _data():
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from datagen import bench_words, gen
    return np.concatenate([bench_words(120000), gen("text", 50000, 5), gen("rand", TOTAL - 170000, 6)])


def _worker(rank, world, port, q, pipelined=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      BLOSC_B200_LIB=EMU, BLOSC_B200_FRAME_WORKERS="2")
    sys.path.insert(0, ROOT)
    import __graft_entry__ as g
    pkg = g.load_package()
    from cblosc_b200 import sharding
    dist.init_process_group("gloo", rank=rank, world_size=world)
    full = torch.from_numpy(_data()) if rank == 0 else None
    comp = sharding.compress_sharded_pipelined if pipelined else sharding.compress_sharded
    decomp = sharding.decompress_sharded_pipelined if pipelined else sharding.decompress_sharded
    frames, sizes = comp(pkg, dist, full, TOTAL, CHUNK, rank, world, "cpu", clevel=5, doshuffle=1, typesize=TS, compressor="lz4")
    t = torch.tensor([1.0 + rank], dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)                      # bench.py's max-over-ranks reduction
    back = decomp(pkg, dist, frames, sizes, TOTAL, CHUNK, rank, world, "cpu")
    if rank == 0:
        q.put(([f.numpy().copy() for f in frames], sizes, back.numpy().copy(), t.item()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_plan_covers_everything(pkg):
    from cblosc_b200 import sharding
    for total in (1, 4, 5, 32):
        for world in (1, 2, 4, 8):
            plan = sharding.shard_plan(total, world)
            assert [c for p in plan for c in p] == list(range(total))
    assert sharding.shard_plan(32, 8)[3] == [12, 13, 14, 15]        # GPU g gets chunks 4g..4g+3
    rg = sharding.byte_ranges(TOTAL, CHUNK, 2)
    assert rg == [(0, 3 * CHUNK), (3 * CHUNK, TOTAL)]
    assert sharding.byte_ranges(10, 4, 8)[3:] == [(10, 10)] * 5     # more ranks than chunks: empty tails


import pytest


@pytest.mark.parametrize("pipelined", [False, True])
def test_two_rank_gloo_sharded_roundtrip(emu, orc, pipelined):
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q, pipelined)) for r in range(2)]
    for p in procs:
        p.start()
    frames, sizes, back, tmax = q.get(timeout=300)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    src = _data()
    assert tmax == 2.0
    assert (back == src).all()
    # every chunk of every rank's frame == the oracle's chunk of that slice (chunks are reference chunks)
    from datagen import compress
    bounds = [(0, 3 * CHUNK), (3 * CHUNK, TOTAL)]
    for r, (lo, hi) in enumerate(bounds):
        f = frames[r]
        assert len(f) == sizes[r] and bytes(f[:4]) == b"B2FR"
        nchunks = int(np.frombuffer(f[28:32].tobytes(), "<u4")[0])
        assert nchunks == 3
        offs = np.frombuffer(f[32:32 + 8 * nchunks].tobytes(), "<u8")
        for i in range(nchunks):
            piece = src[lo + i * CHUNK:min(lo + (i + 1) * CHUNK, hi)]
            rc, want = compress(orc, "orc_compress_ctx", 5, 1, TS, piece, len(piece) + 16, "lz4")
            o = int(offs[i])
            assert (f[o:o + rc] == want[:rc]).all()

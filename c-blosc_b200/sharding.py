"""Multi-GPU sharding of the hot path: whole chunks per GPU.

Chunks are independent units (no cross-chunk state), so a large buffer is cut into
`chunk_bytes` pieces, rank r gets a contiguous run of them, compresses it locally as one frame
(``blosc_b200_frame_compress``: several chunks in flight on that GPU) and nothing is exchanged
inside the algorithm.  The only communication is the trivial scatter of input slices from the
root and the gather-v of the compressed frames back (sizes first, then payloads); decompression
is the mirror.  `dist` is ``torch.distributed`` (NCCL on GPUs; the CPU tests run the very same
code over gloo with world_size 2) or None for a single process.
"""
from __future__ import annotations

import torch


def shard_plan(total_chunks: int, world: int):
    """Chunk c belongs to rank c // ceil(total/world): contiguous runs, rank order = buffer order."""
    per = (total_chunks + world - 1) // world
    return [list(range(r * per, min((r + 1) * per, total_chunks))) for r in range(world)]


def byte_ranges(total_bytes: int, chunk_bytes: int, world: int):
    """[(lo, hi)) of the buffer owned by each rank under shard_plan()."""
    nchunks = (total_bytes + chunk_bytes - 1) // chunk_bytes
    out = []
    for mine in shard_plan(nchunks, world):
        if not mine:
            out.append((total_bytes, total_bytes))
        else:
            out.append((mine[0] * chunk_bytes, min((mine[-1] + 1) * chunk_bytes, total_bytes)))
    return out


def _wait(reqs):
    """Complete P2P requests.  With NCCL, wait() only orders the current CUDA stream after the
    transfer; the library runs on its own streams, so the host waits for the data here."""
    for q in reqs:
        q.wait()
    if reqs and torch.cuda.is_available() and torch.cuda.is_initialized():
        torch.cuda.current_stream().synchronize()


def scatter_bytes(dist, full, ranges, rank, device):
    """Root (rank 0) holds `full`; every rank returns its own [lo, hi) slice as a uint8 tensor."""
    lo, hi = ranges[rank]
    if dist is None or len(ranges) == 1:
        return full[lo:hi]
    if rank == 0:
        ops = [dist.P2POp(dist.isend, full[l:h], r) for r, (l, h) in enumerate(ranges) if r != 0 and h > l]
        _wait(dist.batch_isend_irecv(ops) if ops else [])
        return full[lo:hi]
    mine = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    if hi > lo:
        _wait(dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, 0)]))
    return mine


def gather_bytes(dist, mine, nbytes_mine, rank, world, device):
    """gather-v to rank 0: all ranks learn every size; root returns [tensor per rank], others None."""
    if dist is None or world == 1:
        return [mine[:nbytes_mine]], [int(nbytes_mine)]
    sizes_t = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(sizes_t, torch.tensor([int(nbytes_mine)], dtype=torch.int64, device=device))
    sizes = [int(s.item()) for s in sizes_t]
    if rank == 0:
        parts = [mine[:nbytes_mine]] + [torch.empty(sizes[r], dtype=torch.uint8, device=device) for r in range(1, world)]
        ops = [dist.P2POp(dist.irecv, parts[r], r) for r in range(1, world) if sizes[r] > 0]
        _wait(dist.batch_isend_irecv(ops) if ops else [])
        return parts, sizes
    if nbytes_mine > 0:
        _wait(dist.batch_isend_irecv([dist.P2POp(dist.isend, mine[:nbytes_mine], 0)]))
    return None, sizes


def compress_sharded(pkg, dist, full, total_bytes, chunk_bytes, rank, world, device, *, clevel, doshuffle, typesize,
                     compressor, blocksize=0, numinternalthreads=1):
    """scatter -> per-rank frame_compress -> gather-v.  Root gets ([frame per rank], [frame bytes]);
    the concatenation order is the buffer order."""
    ranges = byte_ranges(total_bytes, chunk_bytes, world)
    mine = scatter_bytes(dist, full, ranges, rank, device)
    n = int(mine.numel())
    bound = pkg.frame_bound(n, typesize, chunk_bytes)
    frame = torch.empty(bound, dtype=torch.uint8, device=device)
    fb = pkg.frame_compress(clevel, doshuffle, typesize, n, mine, frame, bound, compressor, blocksize, chunk_bytes,
                            numinternalthreads)
    if fb <= 0:
        raise RuntimeError(f"rank {rank}: frame_compress returned {fb}")
    return gather_bytes(dist, frame, fb, rank, world, device)


def decompress_sharded(pkg, dist, frames, sizes, total_bytes, chunk_bytes, rank, world, device, numinternalthreads=1):
    """Mirror: root sends frame r to rank r, every rank decodes its slice, slices are gathered on
    the root, which returns the reassembled buffer (others None)."""
    ranges = byte_ranges(total_bytes, chunk_bytes, world)
    if dist is None or world == 1:
        mine = frames[0]
    elif rank == 0:
        ops = [dist.P2POp(dist.isend, frames[r], r) for r in range(1, world) if sizes[r] > 0]
        _wait(dist.batch_isend_irecv(ops) if ops else [])
        mine = frames[0]
    else:
        mine = torch.empty(sizes[rank], dtype=torch.uint8, device=device)
        if sizes[rank] > 0:
            _wait(dist.batch_isend_irecv([dist.P2POp(dist.irecv, mine, 0)]))
    lo, hi = ranges[rank]
    out = torch.empty(hi - lo, dtype=torch.uint8, device=device)
    nb = pkg.frame_decompress(mine, sizes[rank], out, hi - lo, numinternalthreads)
    if nb != hi - lo:
        raise RuntimeError(f"rank {rank}: frame_decompress returned {nb}, expected {hi - lo}")
    parts, _ = gather_bytes(dist, out, hi - lo, rank, world, device)
    return torch.cat(parts) if parts is not None else None

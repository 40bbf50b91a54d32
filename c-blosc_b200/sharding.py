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


# ---------------------------------------------------------------------------------------------
# Pipelined variants: the 7/8 of the buffer that leaves (and re-enters) the root crosses its NVLink
# ports at <= ~770 GB/s per direction, which at 8 GPUs is as long as the compression itself.  So the
# scatter goes out chunk by chunk (round j = chunk j of every peer, one grouped NCCL launch) and
# every rank starts compressing chunk j while chunk j+1 is still in flight; on the way back every
# rank streams decoded chunk j to the root while it decodes chunk j+1.  The small compressed frames
# are still exchanged in one piece.  All collectives are issued from the calling thread; a few
# worker threads only drive the (re-entrant, blocking) chunk API, several chunks at a time.
# ---------------------------------------------------------------------------------------------
import contextlib
import struct
from concurrent.futures import ThreadPoolExecutor

_side_streams = {}


def _comm_ctx(device):
    """Collectives are issued under a private non-blocking CUDA stream: an operation on torch's default
    (= the legacy NULL) stream would act as a barrier between the library's blocking streams and
    serialise the chunks that are being compressed concurrently."""
    if str(device) == "cpu" or not torch.cuda.is_available():
        return contextlib.nullcontext(), None
    key = str(device)
    if key not in _side_streams:
        _side_streams[key] = torch.cuda.Stream(device=device)
    st = _side_streams[key]
    return torch.cuda.stream(st), st


def _done(req, side):
    """Host-side completion of one P2P request (NCCL's wait() only orders the issuing stream)."""
    req.wait()
    if side is not None:
        side.synchronize()


def _chunk_bounds(n, chunk_bytes):
    return [(o, min(o + chunk_bytes, n)) for o in range(0, n, chunk_bytes)]


def _frame_from_chunks(chunks, sizes, n, chunk_bytes, device):
    """A frame (32-byte header + u64 offsets + chunks, blosc_b200.c "frames") from already compressed chunks."""
    k = len(chunks)
    index = 32 + 8 * k
    offs, cur = [], index
    for s in sizes:
        offs.append(cur); cur += s
    hdr = b"B2FR" + bytes([1, 0, 0, 0]) + struct.pack("<QQII", n, cur, chunk_bytes, k) + struct.pack(f"<{k}Q", *offs)
    frame = torch.empty(cur, dtype=torch.uint8, device=device)
    frame[:index] = torch.frombuffer(bytearray(hdr), dtype=torch.uint8).to(device)
    for c, s, o in zip(chunks, sizes, offs):
        frame[o:o + s] = c[:s]
    return frame, cur


def compress_sharded_pipelined(pkg, dist, full, total_bytes, chunk_bytes, rank, world, device, *, clevel, doshuffle,
                               typesize, compressor, blocksize=0, numinternalthreads=1, workers=4):
    """Same result as compress_sharded (root: [frame per rank], [frame bytes]) with the scatter overlapped."""
    ranges = byte_ranges(total_bytes, chunk_bytes, world)
    lo, hi = ranges[rank]
    n = hi - lo
    bounds = _chunk_bounds(n, chunk_bytes)
    ctx, side = _comm_ctx(device)
    recv, sends = None, []
    with ctx:
        if dist is not None and world > 1:
            if rank == 0:
                rounds = max(len(_chunk_bounds(h - l, chunk_bytes)) for l, h in ranges)
                for j in range(rounds):
                    ops = [dist.P2POp(dist.isend, full[l + j * chunk_bytes:min(l + (j + 1) * chunk_bytes, h)], r)
                           for r, (l, h) in enumerate(ranges) if r != 0 and l + j * chunk_bytes < h]
                    if ops:
                        sends += dist.batch_isend_irecv(ops)
                mine = full[lo:hi]
            else:
                mine = torch.empty(n, dtype=torch.uint8, device=device)
                recv = [dist.irecv(mine[a:b], 0) for a, b in bounds]        # FIFO per pair: arrives in chunk order
        else:
            mine = full[lo:hi]
        slots = [torch.empty(b - a + 16, dtype=torch.uint8, device=device) for a, b in bounds]
        sizes = [0] * len(bounds)

        def one(j):
            a, b = bounds[j]
            cb = pkg.compress_ctx(clevel, doshuffle, typesize, b - a, mine[a:b], slots[j], b - a + 16, compressor, blocksize,
                                  numinternalthreads)
            if cb <= 0:
                raise RuntimeError(f"rank {rank}: chunk {j}: blosc_compress_ctx returned {cb}")
            sizes[j] = cb

        with ThreadPoolExecutor(max(1, workers)) as ex:
            futs = []
            for j in range(len(bounds)):
                if recv is not None:
                    _done(recv[j], side)
                futs.append(ex.submit(one, j))
            for f in futs:
                f.result()
        for q in sends:
            _done(q, side)
        frame, fb = _frame_from_chunks(slots, sizes, n, chunk_bytes, device)
        if side is not None:
            side.synchronize()
        return gather_bytes(dist, frame, fb, rank, world, device)


def decompress_sharded_pipelined(pkg, dist, frames, sizes, total_bytes, chunk_bytes, rank, world, device,
                                 numinternalthreads=1, workers=4):
    """Mirror of compress_sharded_pipelined: decoded chunks stream back to the root as they finish."""
    ranges = byte_ranges(total_bytes, chunk_bytes, world)
    lo, hi = ranges[rank]
    n = hi - lo
    bounds = _chunk_bounds(n, chunk_bytes)
    ctx, side = _comm_ctx(device)
    with ctx:
        multi = dist is not None and world > 1
        out_full, recvs = None, []
        if multi and rank == 0:
            out_full = torch.empty(total_bytes, dtype=torch.uint8, device=device)
            ops = [dist.P2POp(dist.isend, frames[r], r) for r in range(1, world) if sizes[r] > 0]
            for q in (dist.batch_isend_irecv(ops) if ops else []):
                _done(q, side)
            rounds = max(len(_chunk_bounds(h - l, chunk_bytes)) for l, h in ranges)
            for j in range(rounds):                                          # posted up-front, same order as the peers send
                ops = [dist.P2POp(dist.irecv, out_full[l + j * chunk_bytes:min(l + (j + 1) * chunk_bytes, h)], r)
                       for r, (l, h) in enumerate(ranges) if r != 0 and l + j * chunk_bytes < h]
                if ops:
                    recvs += dist.batch_isend_irecv(ops)
            mine, out = frames[0], out_full[lo:hi]
        elif multi:
            mine = torch.empty(sizes[rank], dtype=torch.uint8, device=device)
            if sizes[rank] > 0:
                _done(dist.irecv(mine, 0), side)
            out = torch.empty(n, dtype=torch.uint8, device=device)
        else:
            mine = frames[0]
            out = torch.empty(n, dtype=torch.uint8, device=device)
        info = pkg.frame_info(mine, sizes[rank])
        if info is None or info[0] != n or info[3] != len(bounds):
            raise RuntimeError(f"rank {rank}: not a frame of {n} bytes in {len(bounds)} chunks: {info}")
        where = [pkg.frame_chunk(mine, sizes[rank], j) for j in range(len(bounds))]

        def one(j):
            a, b = bounds[j]
            off, cb = where[j]
            nb = pkg.decompress_ctx(mine[off:off + cb], out[a:b], b - a, numinternalthreads)
            if nb != b - a:
                raise RuntimeError(f"rank {rank}: chunk {j}: blosc_decompress_ctx returned {nb}")

        sends = []
        with ThreadPoolExecutor(max(1, workers)) as ex:
            futs = [ex.submit(one, j) for j in range(len(bounds))]
            for j, f in enumerate(futs):
                f.result()
                if multi and rank != 0:
                    a, b = bounds[j]
                    sends.append(dist.isend(out[a:b], 0))                    # in chunk order, as the root posted its receives
        for q in sends + recvs:
            _done(q, side)
        if side is not None:
            side.synchronize()
        if multi:
            return out_full if rank == 0 else None
        return out

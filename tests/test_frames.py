"""Frames: buffers larger than one chunk as a sequence of ordinary
Blosc-1 chunks, several in flight at once.  Every chunk must be byte-identical to what the
reference / oracle writes for that slice, so a frame is just an index in front of reference chunks.
CPU: the product's host code over the emulated backend.  GPU: the real library, including > 2 GiB."""
import ctypes as C
import os

import numpy as np
import pytest

from datagen import bench_words, ci, compress, gen, ptr, sz

ll = C.c_longlong


def _bind(lib):
    lib.blosc_b200_frame_bound.restype = sz
    lib.blosc_b200_frame_bound.argtypes = [sz, sz, sz]
    lib.blosc_b200_frame_compress.restype = ll
    lib.blosc_b200_frame_compress.argtypes = [ci, ci, sz, sz, C.c_void_p, C.c_void_p, sz, C.c_char_p, sz, sz, ci]
    lib.blosc_b200_frame_decompress.restype = ll
    lib.blosc_b200_frame_decompress.argtypes = [C.c_void_p, sz, C.c_void_p, sz, ci]
    lib.blosc_b200_frame_getitem.restype = ll
    lib.blosc_b200_frame_getitem.argtypes = [C.c_void_p, sz, sz, sz, C.c_void_p]
    lib.blosc_b200_frame_info.restype = ci
    lib.blosc_b200_frame_info.argtypes = [C.c_void_p, sz] + [C.POINTER(sz)] * 4
    lib.blosc_b200_frame_chunk.restype = ll
    lib.blosc_b200_frame_chunk.argtypes = [C.c_void_p, sz, sz, C.POINTER(sz)]
    return lib


def _frame_roundtrip(lib, orc, src, ts, comp, shuf, chunksize, nt=1):
    n = len(src)
    bound = lib.blosc_b200_frame_bound(n, ts, chunksize)
    frame = np.full(bound + 64, 0xAA, np.uint8)
    fb = lib.blosc_b200_frame_compress(5, shuf, ts, n, ptr(src), ptr(frame), bound, comp.encode(), 0, chunksize, nt)
    assert fb > 0
    assert (frame[fb:] == 0xAA).all()                       # nothing written past the frame
    v = [sz(0) for _ in range(4)]
    assert lib.blosc_b200_frame_info(ptr(frame), fb, *[C.byref(x) for x in v]) == 0
    nbytes, cbytes, cs, nchunks = [int(x.value) for x in v]
    assert (nbytes, cbytes) == (n, fb)
    assert cs == chunksize - (chunksize % ts if ts > 1 else 0)
    assert nchunks == (n + cs - 1) // cs
    # every chunk is the reference's chunk for that slice
    end = 32 + 8 * nchunks
    for i in range(nchunks):
        cb = sz(0)
        off = lib.blosc_b200_frame_chunk(ptr(frame), fb, i, C.byref(cb))
        assert off == end
        piece = src[i * cs:(i + 1) * cs]
        r, want = compress(orc, "orc_compress_ctx", 5, shuf, ts, piece, len(piece) + 16, comp, 0, nt)
        assert r == cb.value, (i, r, cb.value)
        assert (frame[off:off + r] == want[:r]).all()
        end = off + r
    assert end == fb
    out = np.full(n + 64, 0x55, np.uint8)
    assert lib.blosc_b200_frame_decompress(ptr(frame), fb, ptr(out), n, nt) == n
    assert (out[:n] == src).all() and (out[n:] == 0x55).all()
    return frame, fb, cs


@pytest.mark.parametrize("workers", ["1", "3"])
def test_frame_chunks_equal_oracle_chunks_emu(emu, orc, workers, monkeypatch):
    monkeypatch.setenv("BLOSC_B200_FRAME_WORKERS", workers)
    lib = _bind(emu)
    src = np.concatenate([bench_words(150000), gen("rand", 70000, 1), gen("text", 90001, 2)])
    _frame_roundtrip(lib, orc, src, 4, "lz4", 1, 65536)
    _frame_roundtrip(lib, orc, src[:200003], 8, "blosclz", 2, 50000, nt=2)      # chunksize rounded to 49 992
    _frame_roundtrip(lib, orc, src[:100000], 1, "lz4", 1, 100000)              # exactly one chunk
    _frame_roundtrip(lib, orc, src[:4100], 4, "blosclz", 0, 4096)              # tiny last chunk (< 128: MEMCPYED)


def test_frame_uses_packed_tables_when_chunks_overlap_emu(emu, orc, monkeypatch):
    """The opt-in 17-bit packed LZ4 hash table (BLOSC_B200_LZ4_PACK=1, 128 KiB splits) with two chunks
    in flight: the chunks must still be the oracle's chunks."""
    monkeypatch.setenv("BLOSC_B200_FRAME_WORKERS", "2")
    monkeypatch.setenv("BLOSC_B200_LZ4_PACK", "1")
    lib = _bind(emu)
    src = np.concatenate([bench_words(1 << 20), gen("text", 1 << 20, 2)])
    _frame_roundtrip(lib, orc, src, 4, "lz4", 1, 1 << 20)


def test_frame_edge_cases_emu(emu, orc):
    lib = _bind(emu)
    # empty buffer: header only
    frame = np.zeros(64, np.uint8)
    assert lib.blosc_b200_frame_compress(5, 1, 4, 0, ptr(frame), ptr(frame), 64, b"lz4", 0, 0, 1) == 32
    assert bytes(frame[:4]) == b"B2FR"
    out = np.zeros(8, np.uint8)
    assert lib.blosc_b200_frame_decompress(ptr(frame), 32, ptr(out), 0, 1) == 0
    # argument errors carry the chunk API's codes
    src = gen("i32", 40000)
    big = np.zeros(lib.blosc_b200_frame_bound(40000, 4, 16384) + 8, np.uint8)
    assert lib.blosc_b200_frame_compress(10, 1, 4, 40000, ptr(src), ptr(big), len(big), b"lz4", 0, 16384, 1) == -10
    assert lib.blosc_b200_frame_compress(5, 3, 4, 40000, ptr(src), ptr(big), len(big), b"lz4", 0, 16384, 1) == -10
    assert lib.blosc_b200_frame_compress(5, 1, 4, 40000, ptr(src), ptr(big), len(big), b"zstd", 0, 16384, 1) == -5
    # does not fit: 0, and nothing written past destsize
    rnd = gen("rand", 40000, 7)
    small = np.full(30000 + 64, 0xEE, np.uint8)
    assert lib.blosc_b200_frame_compress(5, 1, 4, 40000, ptr(rnd), ptr(small), 30000, b"lz4", 0, 16384, 1) == 0
    assert (small[30000:] == 0xEE).all()
    assert lib.blosc_b200_frame_compress(5, 1, 4, 40000, ptr(rnd), ptr(small), 40, b"lz4", 0, 16384, 1) == 0
    # incompressible data fits exactly in the bound (every chunk MEMCPYED)
    fb = lib.blosc_b200_frame_compress(5, 1, 4, 40000, ptr(rnd), ptr(big), len(big) - 8, b"lz4", 0, 16384, 1)
    assert fb == lib.blosc_b200_frame_bound(40000, 4, 16384) == 32 + 3 * 8 + 3 * 16 + 40000
    # corrupted index / truncated frame are rejected
    out = np.zeros(40064, np.uint8)
    assert lib.blosc_b200_frame_decompress(ptr(big), fb, ptr(out), 40000, 1) == 40000
    assert lib.blosc_b200_frame_decompress(ptr(big), fb, ptr(out), 39999, 1) == -1      # dest too small
    assert lib.blosc_b200_frame_decompress(ptr(big), fb - 1, ptr(out), 40000, 1) == -1  # truncated
    for pos, val in ((0, 0x41), (4, 2), (28, 9), (32, 0xFF), (40, 1), (25, 0x41)):
        bad = big.copy(); bad[pos] = val
        assert lib.blosc_b200_frame_decompress(ptr(bad), fb, ptr(out), 40000, 1) == -1, pos
    bad = big.copy(); bad[32 + 24 + 5] ^= 0x40                                        # a chunk header's nbytes field
    assert lib.blosc_b200_frame_decompress(ptr(bad), fb, ptr(out), 40000, 1) == -1


def test_frame_getitem_spans_chunks_emu(emu, orc):
    lib = _bind(emu)
    src = gen("i32", 120000)
    frame, fb, cs = _frame_roundtrip(lib, orc, src, 4, "lz4", 1, 32768)
    items = len(src) // 4
    for start, n in ((0, 10), (8190, 5), (8191, 8193 + 4000), (0, items), (items - 1, 1), (items, 0), (100, 0)):
        out = np.full(n * 4 + 16, 0x77, np.uint8)
        assert lib.blosc_b200_frame_getitem(ptr(frame), fb, start, n, ptr(out)) == n * 4
        assert (out[:n * 4] == src[start * 4:(start + n) * 4]).all() and (out[n * 4:] == 0x77).all()
    out = np.zeros(64, np.uint8)
    assert lib.blosc_b200_frame_getitem(ptr(frame), fb, items, 1, ptr(out)) == -1
    assert lib.blosc_b200_frame_getitem(ptr(frame), fb, items + 1, 0, ptr(out)) == -1


# ------------------------------------------------------------------------------------------------
# GPU
# ------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("where", ["host", "device"])
def test_frame_chunks_equal_oracle_chunks_gpu(pkg, orc, cuda, where):
    lib = _bind(pkg.lib)
    src = np.concatenate([bench_words(3 << 20), gen("rand", 1 << 20, 1), gen("text", (1 << 20) + 13, 2)])
    if where == "host":
        _frame_roundtrip(lib, orc, src, 4, "lz4", 1, 1 << 20)
        _frame_roundtrip(lib, orc, src, 8, "blosclz", 2, 1500000, nt=4)
        return
    torch = cuda
    n, ts, cs = len(src), 4, 1 << 20
    d_src = torch.from_numpy(src).cuda()
    bound = pkg.frame_bound(n, ts, cs)
    d_frame = torch.zeros(bound, dtype=torch.uint8, device="cuda")
    fb = pkg.frame_compress(5, 1, ts, n, d_src, d_frame, bound, "lz4", 0, cs)
    h_frame = np.zeros(bound + 64, np.uint8)
    fb2 = lib.blosc_b200_frame_compress(5, 1, ts, n, ptr(src), ptr(h_frame), bound, b"lz4", 0, cs, 1)
    assert fb == fb2 and (d_frame[:fb].cpu().numpy() == h_frame[:fb]).all()
    assert pkg.frame_info(d_frame, fb) == (n, fb, cs, (n + cs - 1) // cs)
    d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    assert pkg.frame_decompress(d_frame, fb, d_out, n) == n
    assert torch.equal(d_out, d_src)
    d_it = torch.zeros(4 * 300000, dtype=torch.uint8, device="cuda")
    assert pkg.frame_getitem(d_frame, fb, 200000, 300000, d_it) == 4 * 300000      # spans chunk 0 -> 1
    assert torch.equal(d_it, d_src[800000:2000000])


@pytest.mark.gpu
def test_frame_larger_than_2gib_gpu(pkg, cuda):
    """2.25 GiB of bench.c data (more than any single Blosc-1 chunk can hold), device resident:
    9 chunks of 256 MiB whose sizes must all equal the reference's cbytes for cfg 2,
    and a bit-exact round trip."""
    torch = cuda
    chunk = 256 << 20
    n = 9 * chunk
    one = torch.from_numpy(bench_words(chunk)).cuda()
    d_src = one.repeat(9)                                     # per-chunk restart of the generator (SURVEY 8d cfg 5)
    bound = pkg.frame_bound(n, 4, chunk)
    d_frame = torch.empty(bound, dtype=torch.uint8, device="cuda")
    fb = pkg.frame_compress(5, 1, 4, n, d_src, d_frame, bound, "lz4", 0, chunk)
    assert fb == 32 + 9 * 8 + 9 * 20401680
    for i in (0, 4, 8):
        assert pkg.frame_chunk(d_frame, fb, i) == (32 + 72 + i * 20401680, 20401680)
    d_out = torch.empty(n, dtype=torch.uint8, device="cuda")
    assert pkg.frame_decompress(d_frame, fb, d_out, n) == n
    assert torch.equal(d_out, d_src)
    # host frame -> pageable host destination of the same > 2 GiB size
    h_frame = d_frame[:fb].cpu().numpy()
    h_out = np.empty(n, np.uint8)
    assert pkg.frame_decompress(h_frame, fb, h_out, n) == n
    assert (h_out.reshape(9, chunk) == one.cpu().numpy()).all()

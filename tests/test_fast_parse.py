"""The segment-parallel LZ4 parse (BLOSC_B200_PARSE=fast, csrc/dev_lz4fast.cuh).

Its chunks are not the reference's bytes, so parity is what north_star asks of a compressor: the
UNMODIFIED reference (oracle/_ref, LZ4_decompress_safe under blosc_decompress_ctx) and the oracle
decode every chunk to the original bytes, the chunk header / bstarts / split prefixes are
well-formed, nothing is written past the returned size, and the ratio on the bench.c planes is
not worse than the reference's own (7.11 / 13.16 / 36.70 / 26.32 at typesize 2/4/8/16).
CPU: the device code runs in the SIMT emulator.  GPU: through the C ABI of libblosc_b200.so."""
import os

import numpy as np
import pytest

from datagen import compress, decompress, gen

KINDS = ("bench", "rand", "zeros", "lowent", "text", "ramp", "i32", "f32", "mixed")


@pytest.fixture()
def fast_env():
    old = os.environ.get("BLOSC_B200_PARSE")
    os.environ["BLOSC_B200_PARSE"] = "fast"
    yield
    if old is None:
        os.environ.pop("BLOSC_B200_PARSE", None)
    else:
        os.environ["BLOSC_B200_PARSE"] = old


def _check(chunk, r, n, src, decoders):
    assert r > 0
    assert (chunk[r:] == 0xAA).all(), "wrote past the returned size"
    assert int.from_bytes(bytes(chunk[4:8]), "little") == n and int.from_bytes(bytes(chunk[12:16]), "little") == r
    for lib, fn in decoders:
        dn, out = decompress(lib, fn, chunk, n)
        assert dn == n and (out[:n] == src).all(), fn


@pytest.mark.parametrize("kind", KINDS)
def test_emu_fast_chunks_decode_with_the_reference(emu, ref, orc, fast_env, kind):
    for n in (1 << 20, 300001, 65536 + 77, 4096 + 5, 1000, 200, 13):
        src = gen(kind, n, seed=n & 7)
        for ts, shuf, clevel in ((4, 1, 5), (1, 0, 5), (8, 1, 9), (2, 2, 1), (16, 1, 5), (3, 1, 5)):
            r, chunk = compress(emu, "blosc_compress_ctx", clevel, shuf, ts, src, n + 16, "lz4")
            _check(chunk, r, n, src, ((ref, "blosc_decompress_ctx"), (orc, "orc_decompress_ctx"), (emu, "blosc_decompress_ctx")))


def test_emu_fast_odd_shapes(emu, ref, fast_env):
    """unaligned sources, forced block sizes (unsplit streams longer than a group of 32 segments), tiny destsize"""
    base = gen("bench", (1 << 20) + 64)
    for off in (1, 2, 3):
        src = base[off:off + 500001].copy()
        r, chunk = compress(emu, "blosc_compress_ctx", 5, 0, 1, src, len(src) + 16, "lz4")
        _check(chunk, r, len(src), src, ((ref, "blosc_decompress_ctx"),))
    src = gen("i32", 1 << 20)
    for bs in (4096, 100000, 1 << 19):
        r, chunk = compress(emu, "blosc_compress_ctx", 5, 1, 32, src, len(src) + 16, "lz4", bs)   # typesize 32: never split
        _check(chunk, r, len(src), src, ((ref, "blosc_decompress_ctx"),))
    # a destination that is too small for the compressed chunk: 0, as blosc_compress (test_maxout.c)
    src = gen("rand", 100000)
    r, _ = compress(emu, "blosc_compress_ctx", 5, 1, 4, src, 100000 + 15, "lz4")
    assert r == 0
    src = gen("bench", 1 << 20)
    full, _ = compress(emu, "blosc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4")
    r, chunk = compress(emu, "blosc_compress_ctx", 5, 1, 4, src, full, "lz4")
    assert r in (0, full)
    r, _ = compress(emu, "blosc_compress_ctx", 5, 1, 4, src, full - 1, "lz4")
    assert r == 0


def test_emu_fast_ratio_on_bench_planes(emu, ref, fast_env):
    """not worse than the reference's own ratio on the data BASELINE.json is quoted on"""
    n = 4 << 20
    src = gen("bench", n)
    for ts in (2, 4, 8, 16):
        want, _ = compress(ref, "blosc_compress_ctx", 5, 1, ts, src, n + 16, "lz4")
        got, _ = compress(emu, "blosc_compress_ctx", 5, 1, ts, src, n + 16, "lz4")
        assert 0 < got <= want * 1.02, (ts, got, want)


def test_emu_fast_is_opt_in(emu, orc):
    """without the variable the chunk is the reference's, byte for byte"""
    os.environ.pop("BLOSC_B200_PARSE", None)
    src = gen("bench", 1 << 20)
    w, want = compress(orc, "orc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4")
    g, got = compress(emu, "blosc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4")
    assert g == w and (got[:g] == want[:w]).all()


# ---------------------------------------------------------------- GPU
def _gpu_compress(pkg, clevel, shuf, ts, src, destsize, comp, bs=0):
    dest = np.full(destsize + 64, 0xAA, np.uint8)
    r = pkg.compress_ctx(clevel, shuf, ts, len(src), src, dest, destsize, comp, bs, 4)
    return r, dest


@pytest.mark.gpu
@pytest.mark.parametrize("kind", KINDS)
def test_gpu_fast_chunks_decode_with_the_oracle(pkg, orc, cuda, fast_env, kind):
    ref_path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libblosc_ref.so")
    decoders = [(orc, "orc_decompress_ctx")]
    if os.path.exists(ref_path):
        import ctypes as C
        lib = C.CDLL(ref_path)
        lib.blosc_decompress_ctx.restype = C.c_int
        decoders.append((lib, "blosc_decompress_ctx"))
    for n in (4 << 20, 1000000, 300001, 65536 + 77, 4096 + 5, 1000, 200):
        src = gen(kind, n, seed=n & 7)
        for ts, shuf, clevel in ((4, 1, 5), (1, 0, 5), (8, 1, 9), (2, 2, 1), (16, 1, 5), (3, 1, 5), (2, 1, 5)):
            r, chunk = _gpu_compress(pkg, clevel, shuf, ts, src, n + 16, "lz4")
            _check(chunk, r, n, src, decoders)
            out = np.zeros(n + 64, np.uint8)
            assert pkg.decompress_ctx(chunk, out, n) == n and (out[:n] == src).all()


@pytest.mark.gpu
def test_gpu_fast_equals_emulator(pkg, emu, cuda, fast_env):
    """the parse is deterministic: the GPU writes what the emulator writes"""
    for kind, n, ts, shuf in (("bench", 4 << 20, 4, 1), ("f32", 1000000, 8, 1), ("mixed", 1 << 20, 2, 2), ("text", 300001, 1, 0)):
        src = gen(kind, n)
        w, want = compress(emu, "blosc_compress_ctx", 5, shuf, ts, src, n + 16, "lz4")
        g, got = _gpu_compress(pkg, 5, shuf, ts, src, n + 16, "lz4")
        assert g == w and (got[:g] == want[:w]).all(), (kind, n)


@pytest.mark.gpu
def test_gpu_fast_full_size_ratio_and_round_trip(pkg, cuda, fast_env):
    """BASELINE.json configs[1] and the typesize sweep of configs[4], device resident, 256 MiB"""
    torch = cuda
    n = 256 << 20
    d_src = torch.from_numpy(gen("bench", n)).cuda()
    d_chunk = torch.zeros(n + 16, dtype=torch.uint8, device="cuda")
    d_out = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for ts, ref_cbytes in ((2, 37749776), (4, 20401680), (8, 7313680), (16, 10199056)):
        cb = pkg.compress_ctx(5, 1, ts, n, d_src, d_chunk, n + 16, "lz4")
        assert 0 < cb <= ref_cbytes * 1.02, (ts, cb, ref_cbytes)
        d_out.zero_()
        assert pkg.decompress_ctx(d_chunk, d_out, n) == n
        assert torch.equal(d_out, d_src)

"""Pins the oracle (oracle/blosc_oracle.c, our plain-C restatement) to the UNMODIFIED reference
compiled from /root/reference into oracle/_ref/libblosc_ref.so: filters, both codecs (bytes of
the compressed streams included) and the whole ctx API must agree byte for byte."""
import numpy as np
import pytest

from datagen import ci, compress, decompress, gen, ptr, sz

KINDS = ["rand", "bench", "zeros", "lowent", "text", "ramp", "i32", "f32", "mixed"]


def test_filters(orc, ref):
    for ts in [1, 2, 3, 4, 5, 7, 8, 12, 16, 17, 32, 33, 255]:
        for n in [0, 1, 7, 8, 64, 192, 500, 1000, 1792, 8000, 8192, 100000, 131072]:
            src = gen("rand", n, seed=ts + n)
            for fr, fo in (("blosc_internal_shuffle", "orc_shuffle"), ("blosc_internal_unshuffle", "orc_unshuffle")):
                a = np.zeros(n + 1, np.uint8); b = np.zeros(n + 1, np.uint8)
                getattr(ref, fr)(sz(ts), sz(n), ptr(src), ptr(a)); getattr(orc, fo)(sz(ts), sz(n), ptr(src), ptr(b))
                assert (a == b).all(), (fr, ts, n)
            if n >= ts:
                for fr, fo in (("blosc_internal_bitshuffle", "orc_bitshuffle"), ("blosc_internal_bitunshuffle", "orc_bitunshuffle")):
                    a = np.zeros(n + 1, np.uint8); b = np.zeros(n + 1, np.uint8); tmp = np.zeros(n + 64, np.uint8)
                    getattr(ref, fr)(sz(ts), sz(n), ptr(src), ptr(a), ptr(tmp)); getattr(orc, fo)(sz(ts), sz(n), ptr(src), ptr(b))
                    assert (a == b).all(), (fr, ts, n)


@pytest.mark.parametrize("kind", KINDS)
def test_lz4_streams(orc, ref, kind):
    for n in [0, 1, 5, 12, 13, 15, 16, 33, 64, 67, 128, 255, 1000, 4096, 65535, 65546, 65547, 70000, 131072, 200001]:
        src = gen(kind, n, seed=n)
        for accel in (1, 5, 9):
            for cap in sorted({n, max(n - 1, 0), n // 2, n + n // 255 + 16, 70}):
                a = np.zeros(cap + 64, np.uint8); b = np.zeros(cap + 64, np.uint8)
                ra = ref.LZ4_compress_fast(ptr(src), ptr(a), ci(n), ci(cap), ci(accel))
                rb = orc.orc_lz4_compress_fast(ptr(src), ptr(b), ci(n), ci(cap), ci(accel))
                assert ra == rb, (kind, n, accel, cap, ra, rb)
                if ra > 0:
                    assert (a[:ra] == b[:ra]).all()
                    for c2 in (n, n + 3, max(n - 1, 0)):
                        o1 = np.zeros(n + 8, np.uint8); o2 = np.zeros(n + 8, np.uint8)
                        d1 = ref.LZ4_decompress_safe(ptr(a), ptr(o1), ci(ra), ci(c2))
                        d2 = orc.orc_lz4_decompress_safe(ptr(a), ptr(o2), ci(ra), ci(c2))
                        assert (d1 < 0) == (d2 < 0) and (d1 < 0 or d1 == d2), (kind, n, c2, d1, d2)
                        if c2 == n:
                            assert d2 == n and (o2[:n] == src).all()


@pytest.mark.parametrize("kind", KINDS)
def test_blosclz_streams(orc, ref, kind):
    for n in [0, 1, 15, 16, 17, 33, 64, 66, 67, 128, 255, 1000, 4096, 16500, 65536, 70000, 131072, 200001]:
        src = gen(kind, n, seed=n)
        for clevel in (1, 2, 5, 9):
            for split in (0, 1):
                for cap in sorted({n, max(n - 1, 0), n // 2, 66, 65}):
                    a = np.zeros(cap + 64, np.uint8); b = np.zeros(cap + 64, np.uint8)
                    ra = ref.blosclz_compress(ci(clevel), ptr(src), ci(n), ptr(a), ci(cap), ci(split))
                    rb = orc.orc_blosclz_compress(ci(clevel), ptr(src), ci(n), ptr(b), ci(cap), ci(split))
                    assert ra == rb, (kind, n, clevel, split, cap, ra, rb)
                    if ra > 0:
                        assert (a[:ra] == b[:ra]).all()
                        o1 = np.zeros(n + 8, np.uint8); o2 = np.zeros(n + 8, np.uint8)
                        d1 = ref.blosclz_decompress(ptr(a), ci(ra), ptr(o1), ci(n))
                        d2 = orc.orc_blosclz_decompress(ptr(a), ci(ra), ptr(o2), ci(n))
                        assert d1 == d2 == n and (o2[:n] == src).all()


@pytest.mark.parametrize("kind", ["bench", "rand", "i32", "text", "mixed"])
def test_ctx_api(orc, ref, kind):
    """blosc_compress_ctx / blosc_decompress_ctx / blosc_getitem: same return codes, same chunk bytes."""
    for n in [0, 1, 100, 127, 128, 129, 1000, 4096, 32768, 65536, 100000, 641091, (1 << 20) + 12345]:
        src = gen(kind, n, seed=n)
        for comp in ("lz4", "blosclz"):
            for ts in ([1, 2, 3, 4, 8, 16, 17, 256] if n <= 100000 else [4, 8]):
                for shuf in (0, 1, 2):
                    for clevel in ([0, 1, 5, 9] if n <= 100000 else [5]):
                        for bs in ([0, 100, 4096] if n <= 100000 else [0]):
                            for destsize in sorted({n + 16, n + 15, max(16, n // 2), 15}):
                                ra, a = compress(ref, "blosc_compress_ctx", clevel, shuf, ts, src, destsize, comp, bs)
                                rb, b = compress(orc, "orc_compress_ctx", clevel, shuf, ts, src, destsize, comp, bs)
                                assert ra == rb, (kind, n, comp, ts, shuf, clevel, bs, destsize, ra, rb)
                                if ra <= 0:
                                    continue
                                assert (a[:ra] == b[:ra]).all()
                                d1, o1 = decompress(ref, "blosc_decompress_ctx", a, n)
                                d2, o2 = decompress(orc, "orc_decompress_ctx", a, n)
                                assert d1 == d2 == n and (o2[:n] == src).all()
                                if 0 < n <= 4096:
                                    nit = n // int(a[3])
                                    for st, cnt in ((0, nit), (nit // 3, nit // 2), (nit - 1, 1)):
                                        g1 = np.zeros(n + 8, np.uint8); g2 = np.zeros(n + 8, np.uint8)
                                        r1 = ref.blosc_getitem(ptr(a), ci(st), ci(cnt), ptr(g1))
                                        r2 = orc.orc_getitem(ptr(a), ci(st), ci(cnt), ptr(g2))
                                        assert r1 == r2 and (g1 == g2).all()


def test_reference_multithread_same_cbytes(ref):
    """ctx compress with nthreads>1 gives the same cbytes (block order may differ)."""
    src = gen("bench", 8 << 20)
    r1, a = compress(ref, "blosc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4", 0, 1)
    r4, b = compress(ref, "blosc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4", 0, 4)
    assert r1 == r4

"""The oracle against the reference's golden vectors: compat/*.cdata (chunks written by blosc
1.3.0 ... 1.18.0, compat/filegen.c:33,61-66 -> int32 data[i] = i) and the known-answer table of
SURVEY.md Appendix B / tests/test_maxout.c / tests/test_compressor.c."""
import glob
import os

import numpy as np

from datagen import compress, decompress, gen

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "compat", "*.cdata")))


def test_compat_chunks(orc):
    want = np.arange(1000000, dtype=np.int32).view(np.uint8)
    assert len(GOLD) == 29
    decoded = 0
    for f in GOLD:
        chunk = np.fromfile(f, np.uint8)
        r, out = decompress(orc, "orc_decompress_ctx", chunk, 4000000)
        if any(c in f for c in ("zlib", "zstd", "snappy")):
            assert r == -5, f
        else:
            assert r == 4000000 and (out[:4000000] == want).all(), f
            decoded += 1
    assert decoded == 17


def test_known_answers(orc):
    i32 = np.arange(1 << 18, dtype=np.int32).view(np.uint8).copy()      # 1 MiB of int32 i
    r, c = compress(orc, "orc_compress_ctx", 5, 1, 4, i32, len(i32) + 16, "lz4")
    assert r == 8848 and bytes(c[:16]) == bytes.fromhex("02012104" "00001000" "00000800" "90220000")
    r, c = compress(orc, "orc_compress_ctx", 5, 1, 4, i32[:0], 16, "lz4")
    assert r == 16 and bytes(c[:16]) == bytes.fromhex("02013304" "00000000" "01000000" "10000000")
    r, c = compress(orc, "orc_compress_ctx", 5, 1, 4, i32[:100], 116, "lz4")
    assert r == 116 and c[2] == 0x33
    r, c = compress(orc, "orc_compress_ctx", 5, 1, 4, i32[:128], 144, "lz4")
    assert r == 72 and c[2] == 0x31
    r, c = compress(orc, "orc_compress_ctx", 5, 1, 4, i32, len(i32) + 16, "lz4", 4096)
    assert int(c[8:12].view(np.int32)[0]) == 65536 and c[2] == 0x21
    rnd = gen("rand", 1 << 20)
    assert compress(orc, "orc_compress_ctx", 5, 1, 4, rnd, len(rnd) + 16, "lz4")[0] == len(rnd) + 16
    assert compress(orc, "orc_compress_ctx", 5, 1, 4, rnd, len(rnd), "lz4")[0] == 0
    assert compress(orc, "orc_compress_ctx", 10, 1, 4, rnd, len(rnd) + 16, "lz4")[0] == -10
    assert compress(orc, "orc_compress_ctx", 5, 1, 4, rnd, len(rnd) + 16, "snappy")[0] == -5


def test_baseline_md_sizes(orc):
    """BASELINE.md section 2, scaled to a size the oracle finishes in a second: the bench.c pattern is
    periodic in 2 MiB so an 8 MiB buffer compresses to exactly 1/32 of the 256 MiB body."""
    src = gen("bench", 8 << 20)
    r, _ = compress(orc, "orc_compress_ctx", 5, 1, 4, src, len(src) + 16, "lz4")
    nblocks256 = 512
    body256 = 20401680 - 16 - 4 * nblocks256
    assert (r - 16 - 4 * 16) * 32 == body256

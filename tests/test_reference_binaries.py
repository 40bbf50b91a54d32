"""Drop-in evidence: the REFERENCE'S OWN test, bench and compat programs (tests/test_*.c,
bench/bench.c, compat/filegen.c), compiled unmodified from /root/reference by `make -C oracle
reftests` and linked against libblosc_b200.so instead of libblosc, run on the GPU.
Only the built binaries travel (oracle/_ref/tests, git-ignored); parameter rows below are taken
from the reference's tests/*.csv."""
import glob
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "oracle", "_ref", "tests")


def _run(name, *args, env=None, timeout=600):
    import tempfile
    exe = os.path.join(BIN, name)
    if not os.path.exists(exe):
        pytest.skip("oracle/_ref/tests not built (needs /root/reference at build time)")
    e = dict(os.environ)
    e.update(env or {})
    with tempfile.TemporaryDirectory() as cwd:          # some of the programs write scratch .cdata files
        return subprocess.run([exe, *map(str, args)], capture_output=True, text=True, timeout=timeout, env=e, cwd=cwd)


@pytest.mark.parametrize("name", ["test_api", "test_maxout", "test_compressor", "test_nolock", "test_noinit", "test_nthreads"])
def test_minunit_suites(cuda, name):
    r = _run(name)
    assert r.returncode == 0, (name, r.stdout[-600:], r.stderr[-300:])
    assert "ALL TESTS PASSED" in r.stdout, (name, r.stdout[-600:])


def test_bitshuffle_leftovers_program(cuda):
    """tests/test_bitshuffle_leftovers.c (python-blosc#220): 641091 bytes, lz4, clevel 9, bitshuffle 4 and 8."""
    r = _run("test_bitshuffle_leftovers")
    assert r.returncode == 0, (r.stdout[-600:], r.stderr[-300:])
    assert r.stdout.count("Successful roundtrip!") == 2, r.stdout[-600:]


# rows of tests/test_compress_roundtrip.csv / test_getitem.csv: type_size, num_elements, alignment, clevel, shuffle, threads
ROWS = [(1, 7, 32, 5, 0, 1), (1, 192, 32, 5, 1, 1), (2, 1792, 32, 5, 1, 1), (4, 500, 32, 5, 1, 1), (4, 8000, 32, 5, 0, 1),
        (7, 100000, 32, 5, 1, 1), (8, 702713, 32, 5, 1, 1), (16, 100000, 32, 5, 1, 1), (22, 8000, 32, 5, 1, 1),
        (53, 1792, 32, 5, 0, 1), (80, 100000, 32, 5, 1, 1), (3, 702713, 32, 5, 1, 1)]


@pytest.mark.parametrize("row", ROWS)
def test_csv_roundtrip_and_getitem(cuda, row):
    for name in ("test_compress_roundtrip", "test_getitem"):
        r = _run(name, *row)
        assert r.returncode == 0, (name, row, r.stdout[-300:], r.stderr[-300:])


def test_compat_filegen_decompress(cuda):
    """compat/CMakeLists.txt:18-33: `filegen decompress <file>` for every blosclz/lz4/lz4hc/zlib/zstd golden."""
    files = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "compat", "*.cdata")))
    n = 0
    for f in files:
        if "snappy" in f:
            continue
        r = _run("filegen", "decompress", f)
        assert r.returncode == 0 and "Decompression successful" in r.stdout, (f, r.stdout[-200:])
        n += 1
    assert n == 25


def test_reference_bench_program(cuda):
    """bench/bench.c "test" suite (bench/CMakeLists.txt:27-104) with its own memcmp check."""
    for codec, filt in (("lz4", "shuffle"), ("blosclz", "bitshuffle")):
        r = _run("bench", codec, filt, "test", timeout=900)
        assert r.returncode == 0, (codec, filt, r.stdout[-400:], r.stderr[-300:])
        assert "OK" in r.stdout

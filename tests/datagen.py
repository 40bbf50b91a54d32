"""Deterministic test inputs (seeded) and thin ctypes helpers shared by the tests."""
import ctypes as C

import numpy as np

vp, sz, ci = C.c_void_p, C.c_size_t, C.c_int


def ptr(a):
    return a.ctypes.data_as(vp)


def bench_words(nbytes, rshift=19, start=0):
    """bench/bench.c:141-170 synthetic buffer: int32 w[i] = ((i<<26)^(i<<18)^(i<<11)^(i<<3)^i) & mask."""
    i = np.arange(start, start + (nbytes + 3) // 4, dtype=np.uint32)
    w = ((i << np.uint32(26)) ^ (i << np.uint32(18)) ^ (i << np.uint32(11)) ^ (i << np.uint32(3)) ^ i)
    if rshift < 32:
        w &= np.uint32((1 << rshift) - 1)
    return w.view(np.uint8)[:nbytes].copy()


def gen(kind, n, seed=0):
    rng = np.random.default_rng(seed)
    if kind == "rand":
        return rng.integers(0, 256, n, dtype=np.uint8)
    if kind == "bench":
        return bench_words(n)
    if kind == "zeros":
        return np.zeros(n, np.uint8)
    if kind == "lowent":
        return rng.integers(0, 4, n, dtype=np.uint8)
    if kind == "text":
        words = [bytes(rng.integers(97, 123, rng.integers(2, 9), dtype=np.uint8)) for _ in range(200)]
        out = b" ".join(words[j] for j in rng.integers(0, 200, n // 4 + 8))
        return np.frombuffer(out[:n], np.uint8).copy()
    if kind == "ramp":
        return (np.arange(n) // 7 % 251).astype(np.uint8)
    if kind == "i32":
        return np.arange(n // 4 + 1, dtype=np.int32).view(np.uint8)[:n].copy()
    if kind == "f32":
        return np.linspace(0, 100, n // 4 + 1, dtype=np.float32).view(np.uint8)[:n].copy()
    if kind == "mixed":          # compressible runs interleaved with noise: raw and compressed splits in one chunk
        a = np.zeros(n, np.uint8)
        noise = rng.integers(0, 256, n, dtype=np.uint8)
        seg = max(n // 16, 1)
        for k in range(0, n, seg):
            if (k // seg) % 2:
                a[k:k + seg] = noise[k:k + seg]
            else:
                a[k:k + seg] = (np.arange(min(seg, n - k)) // 3 % 200).astype(np.uint8)
        return a
    raise ValueError(kind)


def compress(lib, fn, clevel, shuf, ts, src, destsize, comp, bs=0, nt=1, fill=0xAA):
    dest = np.full(destsize + 64, fill, np.uint8)
    r = getattr(lib, fn)(ci(clevel), ci(shuf), sz(ts), sz(len(src)), ptr(src), ptr(dest), sz(destsize), comp.encode(), sz(bs), ci(nt))
    return r, dest


def decompress(lib, fn, chunk, destsize, nt=1):
    dest = np.zeros(destsize + 64, np.uint8)
    r = getattr(lib, fn)(ptr(chunk), ptr(dest), sz(destsize), ci(nt))
    return r, dest

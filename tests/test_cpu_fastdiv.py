"""Host logic of the tile-index decode: FastDiv (csrc/dbx_ptx.cuh) replaces n / d by umul64hi(n, m) with
m = floor((2^64 - 1) / d) + 1.  This restates make()/div() with Python integers and checks exactness over the whole
range the kernels use (0 <= n < 2^31, 1 <= d < 2^31), including powers of two, d = 1 and the extremes."""
import random


def make(d):
    d = max(d, 1)
    return (0 if d == 1 else ((2 ** 64 - 1) // d + 1) & (2 ** 64 - 1)), d


def div(n, m, d):
    return n if d == 1 else (n * m) >> 64


def test_fastdiv_exact():
    rnd = random.Random(0)
    ds = [1, 2, 3, 5, 7, 8, 15, 16, 30, 60, 74, 148, 225, 900, 3600, 14400, 57600, 131072, 2 ** 20 - 1, 2 ** 30, 2 ** 31 - 1]
    ds += [rnd.randrange(1, 2 ** 31) for _ in range(300)] + [rnd.randrange(1, 5000) for _ in range(300)]
    for d in ds:
        m, dd = make(d)
        assert m < 2 ** 64
        ns = [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1, 2 ** 31 - d] + [rnd.randrange(0, 2 ** 31) for _ in range(200)]
        for n in ns:
            if 0 <= n < 2 ** 31:
                assert div(n, m, dd) == n // d, (n, d)

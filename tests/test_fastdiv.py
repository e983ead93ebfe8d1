"""The multiply-high division of csrc/tzk_bwd.cu (`FastDiv` / `make_fast_div` / `fast_div`: v / d for 0 <= v < 2^31 with a
run-time divisor, Granlund & Montgomery): the same arithmetic restated with Python integers, checked against `//` over the
divisors the kernels see (batch sizes, F * B spans) and a random sweep.  (The CUDA side is covered by every fused-backward
GPU test; this pins the arithmetic itself.)"""
import random


def make_fast_div(d):
    if d <= 1:
        return 0, 0
    l = 0
    while (1 << l) < d:
        l += 1
    m = ((1 << (31 + l)) + d - 1) // d
    assert m < (1 << 32), (d, m)        # the multiplier fits 32 bits for every divisor below 2^31
    return m, l - 1


def fast_div(v, f):
    m, s = f
    return (((v * m) & ((1 << 64) - 1)) >> 32) >> s if m else v     # __umulhi(v, m) >> s


def test_fast_div_matches_integer_division():
    rnd = random.Random(7)
    divisors = list(range(1, 260)) + [512, 8192, 32768, 65536, 65537, 26 * 65536, 52 * 32768, 2 ** 30, 2 ** 30 + 1,
                                      2 ** 31 - 1] + [rnd.randrange(1, 2 ** 31) for _ in range(1500)]
    for d in divisors:
        f = make_fast_div(d)
        for v in [0, 1, d - 1, d, d + 1, 2 * d - 1, 2 * d, 2 ** 31 - 1] + [rnd.randrange(0, 2 ** 31) for _ in range(40)]:
            if 0 <= v < 2 ** 31:
                assert fast_div(v, f) == v // d, (d, v)

"""CPU: the exact int8 digit-plane scheme of csrc/tc_common.cuh (block contraction A p on tcgen05 kind::i8), restated
with Python integers: balanced base-256 digits, the plane / slice recombination identity, the int32 accumulator bound
and the size of the one dropped term.  Pins the arithmetic the kernel relies on (the CUDA implementation itself is
checked against the fp64 tensor-core path and the oracle in the -m gpu tests and tools/tc_check.py)."""
import numpy as np

from optimization_b200 import problems as P


def balanced_digits(v: int, n: int):
    """v = sum_i d_i 256^i with d_i in [-128, 127] (the bias trick of slice_tile_to_smem: ((v + bias) ^ bias) bytes)."""
    out = []
    for _ in range(n):
        d = ((v + 128) % 256) - 128
        out.append(d)
        v = (v - d) // 256
    assert v == 0
    return out


def test_balanced_digit_ranges_and_bias_trick():
    rng = np.random.default_rng(5)
    bias = 0x0080808080808080
    for _ in range(2000):
        F = int(rng.integers(-(1 << 54), 1 << 54))
        d = balanced_digits(F, 7)
        assert all(-128 <= x <= 127 for x in d) and sum(x * 256 ** i for i, x in enumerate(d)) == F
        u = ((F + bias) & ((1 << 64) - 1)) ^ bias                      # what the kernel computes on the u64 pattern
        byts = [((u >> (8 * i)) & 0xFF) for i in range(7)]
        assert [b - 256 if b >= 128 else b for b in byts] == d
    for A in (-(1 << 22) + 1, (1 << 22) - 1, 0, 12345, -54321):
        a = balanced_digits(A, 3)
        assert all(-128 <= x <= 127 for x in a)


def test_plane_slice_recombination_is_exact_up_to_the_dropped_term():
    rng = np.random.default_rng(7)
    K = 128
    for trial in range(20):
        Arow = [int(x) for x in rng.integers(-(1 << 22) + 1, 1 << 22, K)]          # one row of A' (|A'| < 2^22)
        Fcol = [int(x) for x in rng.integers(-(1 << 54), 1 << 54, K)]              # one column of F (55-bit signed)
        a = [balanced_digits(v, 3) for v in Arow]
        d = [balanced_digits(v, 7) for v in Fcol]
        D = []
        for u in range(8):                                                          # D_u = sum_{h + i = 8 - u} a_h d_i over k
            s = 8 - u
            acc = 0
            for k in range(K):
                for h in range(3):
                    i = s - h
                    if 0 <= i < 7:
                        acc += a[k][h] * d[k][i]
            assert abs(acc) < (1 << 31)                                             # int32 TMEM accumulators cannot wrap
            assert abs(acc) <= 3 * K * 128 * 128
            D.append(acc)
        exact = sum(x * y for x, y in zip(Arow, Fcol))
        dropped = sum(a[k][0] * d[k][0] for k in range(K))                          # the (a0, d0) pair, weight 2^0
        recombined = sum(D[u] << (8 * (8 - u)) for u in range(8))                   # 2^64 sum_u D_u 2^(-8u)
        assert recombined + dropped == exact
        assert abs(dropped) <= K * 128 * 128                                        # <= 2^21: 2^-43 of 2^64, far below fp64 eps


def test_recombination_as_the_kernel_does_it_rounds_once():
    # recombine_row16: part0 = sum_{u<4} D_u 2^(8(3-u)), part1 = sum_{u>=4} D_u 2^(8(7-u)); out = fma(part1, 2^-32, part0) 2^-24
    rng = np.random.default_rng(9)
    for _ in range(500):
        D = [int(x) for x in rng.integers(-(1 << 23), 1 << 23, 8)]
        part0 = sum(D[u] << (8 * (3 - u)) for u in range(4))
        part1 = sum(D[u] << (8 * (7 - u)) for u in range(4, 8))
        assert abs(part0) < (1 << 53) and abs(part1) < (1 << 53)                    # both convert to fp64 exactly
        exact = (part0 * (1 << 32) + part1)                                         # = 2^56 sum_u D_u 2^(-8u)
        got = np.float64(part1) * 2.0 ** -32 + np.float64(part0)                   # fma: single rounding of the exact sum
        from fractions import Fraction
        assert abs(Fraction(float(got)) - Fraction(exact, 1 << 32)) <= Fraction(abs(exact), 1 << 32) * Fraction(1, 1 << 52)


def test_synthetic_A_is_block_fixed_point():
    # the benchmark operator: every 128 x 128 block of A (bf16 storage) is an integer multiple of ONE power of two with
    # integers below 2^22 (the condition stiefel_planes_kernel checks before the tcgen05 path is taken)
    prob = P.make_stiefel_critical(2048, 32)
    A = P.from_bf16_bits(prob.A_bf16)
    for blk in A:
        e = 0
        while not np.all(blk * 2.0 ** e == np.rint(blk * 2.0 ** e)):
            e += 1
            assert e < 64
        q = blk * 2.0 ** e
        assert np.abs(q).max() < (1 << 22)
        for v in (int(q.max()), int(q.min())):
            assert all(-128 <= x <= 127 for x in balanced_digits(v, 3))

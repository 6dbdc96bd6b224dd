"""div_known (csrc/sfxm_device.cuh): q0 = RN(a rb), r = RN(a - q0 b), q = RN(q0 + r rb) with rb = RN(1 / b) must be the
correctly rounded quotient.  The same three operations in exact rational arithmetic on the CPU (an FMA is one rounding
of an exact rational), on random operands, on products nudged by an ulp and on the quotients that lie closest to a
rounding boundary (the construction of sfxm_selftest_div mode 2, which runs 2^29 of them on the GPU)."""
import math
import random
from fractions import Fraction as Fr


def fma(a, b, c):
    return float(Fr(a) * Fr(b) + Fr(c))   # Fraction -> float rounds to nearest even


def div_known(a, b):
    rb = 1.0 / b
    q0 = a * rb
    return fma(fma(-q0, b, a), rb, q0)


def exact(a, b):
    return float(Fr(a) / Fr(b))


def test_random_operands():
    rnd = random.Random(1)
    for _ in range(20000):
        a = rnd.uniform(1, 2) * 2.0 ** rnd.randint(-60, 60) * rnd.choice([-1, 1])
        b = rnd.uniform(1, 2) * 2.0 ** rnd.randint(-60, 60) * rnd.choice([-1, 1])
        assert div_known(a, b) == exact(a, b)
    assert div_known(0.0, 3.0) == 0.0


def test_products_nudged_by_one_ulp():
    rnd = random.Random(2)
    for _ in range(20000):
        b = rnd.uniform(1, 2)
        if rnd.random() < 0.3:
            b = 2.0 - rnd.randint(1, 64) * 2.0 ** -52
        q = rnd.uniform(1, 2)
        p = q * b
        for a in (p, math.nextafter(p, math.inf), math.nextafter(p, -math.inf)):
            assert div_known(a, b) == exact(a, b)


def _nearest_to_midpoint(rnd, bits):
    """(a, b) with a / b = midpoint + k / (2^105 b): M Bi + k = A 2^bits, M odd with 54 bits (a midpoint of [1, 2))."""
    while True:
        Bi = rnd.getrandbits(52) | (1 << 52) | 1
        sel = rnd.random()
        if sel < 0.3:
            Bi = (1 << 53) - 1 - 2 * rnd.randint(0, 255)
        elif sel < 0.5:
            Bi = (1 << 52) + 1 + 2 * rnd.randint(0, 255)
        k = rnd.choice([1, -1, 3, -3, 5, -5, 7, -7])
        M = (-k * pow(Bi, -1, 1 << bits)) % (1 << bits)
        if bits == 53:
            M += 1 << 53
        if M < (1 << 53):
            continue
        P = M * Bi + k
        A = P >> bits
        if P % (1 << bits) or A >= (1 << 53):
            continue
        a = A / 2.0 ** (51 if bits == 54 else 52)
        b = Bi / 2.0 ** 52
        assert abs(Fr(a) / Fr(b) - Fr(M, 1 << 53)) == Fr(abs(k), (1 << 105)) / Fr(b)
        return a, b


def test_quotients_nearest_to_a_rounding_boundary():
    rnd = random.Random(3)
    for bits in (53, 54):
        for _ in range(15000):
            a, b = _nearest_to_midpoint(rnd, bits)
            assert div_known(a, b) == exact(a, b)
            sa, sb = 2.0 ** rnd.randint(-40, 40), -(2.0 ** rnd.randint(-40, 40))
            assert div_known(a * sa, b * sb) == exact(a * sa, b * sb)


def test_corner_where_divisor_and_quotient_are_both_just_below_two():
    """The corner the error analysis has least room in: b and a/b both close to 2 (a close to 4, where a rb can be almost
    one ulp away from a/b) with a/b as close to a rounding boundary as these divisors allow.  M Bi + k = A 2^54 with
    M = 2^54 - p, Bi = 2^53 - j forces p j + k = 2^53 (mod 2^54): p and j cannot both be small, which is what keeps q0 next
    to the boundary.  For each small j the odd k with the largest quotients are searched."""
    n = 0
    for j in range(1, 2001, 2):
        Bi = (1 << 53) - j
        inv = pow(Bi, -1, 1 << 54)
        best = []
        for k in range(-4095, 4096, 2):
            M = (-k * inv) % (1 << 54)
            if M < (1 << 53):
                continue
            P = M * Bi + k
            A = P >> 54
            if P % (1 << 54) or A >= (1 << 53):
                continue
            best.append((M, k, A))
        best.sort(reverse=True)
        for M, k, A in best[:6]:                      # the quotients nearest to 2 for this divisor
            a, b = A / 2.0 ** 51, Bi / 2.0 ** 52
            assert div_known(a, b) == exact(a, b), (a.hex(), b.hex())
            n += 1
    assert n > 4000


def test_extreme_corner_constructed():
    """The same corner at its extreme: p ~ j ~ 2^26.5 (the smallest both can be, since p j + k = 2^53 mod 2^54), i.e. a/b
    and b within ~2^-26 of 2, a/b still within 1e-8 ulp of a rounding boundary.  Here a rb lies almost a full ulp from
    a/b, the largest excursion the first approximation q0 can make next to a boundary."""
    root = 94906265                      # ~ sqrt(2^53)
    n = 0
    for j in range(root - 40000, root + 40001, 2):
        for sign in (1, -1):
            rem = (1 << 53) % j
            k = rem if sign > 0 else rem - j          # p j + k = 2^53
            if k % 2 == 0:
                continue
            p = ((1 << 53) - k) // j
            if p % 2 == 0 or p * j + k != (1 << 53):
                continue
            M, Bi = (1 << 54) - p, (1 << 53) - j
            P = M * Bi + k
            if P % (1 << 54):
                continue
            A = P >> 54
            if A >= (1 << 53):
                continue
            a, b = A / 2.0 ** 51, Bi / 2.0 ** 52
            assert abs(Fr(a) / Fr(b) - Fr(M, 1 << 53)) < Fr(1, 1 << 78)      # within 2^-26 ulp of the midpoint M 2^-53
            assert div_known(a, b) == exact(a, b), (a.hex(), b.hex())
            n += 1
    assert n > 5000

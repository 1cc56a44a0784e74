"""Pure-Python big-integer oracle for the HE kernels.

Independent of the native library: used by the tests to pin down the exact semantics of
the NTT ordering, negacyclic products, CRT centring, CKKS canonical embedding and BFV
fractional encoding. Slow by design (O(N^2) where that is the clearest definition).
"""
from __future__ import annotations

import cmath
import math
from typing import List, Sequence


def is_prime(n: int) -> bool:
    if n < 2:
        return False
    small = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)
    for p in small:
        if n % p == 0:
            return n == p
    d, s = n - 1, 0
    while d % 2 == 0:
        d //= 2
        s += 1
    for a in small:
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(s - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def bit_reverse(x: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def minimal_psi(q: int, n: int) -> int:
    """Smallest primitive 2n-th root of unity mod q."""
    two_n = 2 * n
    assert (q - 1) % two_n == 0
    e = (q - 1) // two_n
    g = 2
    while True:
        c = pow(g, e, q)
        if pow(c, n, q) == q - 1:
            break
        g += 1
    best = c
    cur, sq = c, c * c % q
    for _ in range(n - 1):
        cur = cur * sq % q
        best = min(best, cur)
    return best


def ntt_by_definition(a: Sequence[int], q: int, psi: int) -> List[int]:
    """a_hat[k] = a(psi^(2*brev(k)+1)) — the ordering the kernels must reproduce."""
    n = len(a)
    bits = n.bit_length() - 1
    out = []
    for k in range(n):
        x = pow(psi, 2 * bit_reverse(k, bits) + 1, q)
        acc, xp = 0, 1
        for c in a:
            acc = (acc + c * xp) % q
            xp = xp * x % q
        out.append(acc)
    return out


def negacyclic_mul(a: Sequence[int], b: Sequence[int], q: int) -> List[int]:
    n = len(a)
    out = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            k = i + j
            if k < n:
                out[k] = (out[k] + x * y) % q
            else:
                out[k - n] = (out[k - n] - x * y) % q
    return out


def crt_centered(residues: Sequence[int], moduli: Sequence[int]) -> int:
    """Centred representative in (-Q/2, Q/2] of the CRT lift."""
    Q = 1
    for m in moduli:
        Q *= m
    x = 0
    for r, m in zip(residues, moduli):
        Qi = Q // m
        x += r * Qi * pow(Qi, -1, m)
    x %= Q
    return x - Q if x > Q // 2 else x


def ckks_slots_of(coeffs: Sequence[float], n: int) -> List[complex]:
    """Canonical embedding restricted to the orbit of 5: slot j = m(zeta^(5^j))."""
    m = 2 * n
    out = []
    g = 1
    for _ in range(n // 2):
        z = cmath.exp(2j * math.pi * g / m)
        acc, zp = 0j, 1 + 0j
        for c in coeffs:
            acc += c * zp
            zp *= z
        out.append(acc)
        g = g * 5 % m
    return out


def frac_encode(v: float, n: int, int_digits: int = 64, frac_digits: int = 32) -> List[int]:
    """SEAL 2.x FractionalEncoder, base 2 (signed-digit coefficients)."""
    out = [0] * n
    sgn = -1 if v < 0 else 1
    v = abs(v)
    ip = int(math.floor(v))
    fp = v - ip
    i = 0
    while ip and i < int_digits:
        out[i] = sgn * (ip & 1)
        ip >>= 1
        i += 1
    for i in range(1, frac_digits + 1):
        fp *= 2
        bit = 1 if fp >= 1 else 0
        fp -= bit
        out[n - i] = -sgn * bit
    return out


def frac_decode(coeffs: Sequence[int], int_digits: int = 64, frac_digits: int = 32) -> float:
    n = len(coeffs)
    acc = 0.0
    for i in range(int_digits):
        acc += coeffs[i] * 2.0 ** i
    for i in range(1, frac_digits + 1):
        acc -= coeffs[n - i] * 2.0 ** (-i)
    return acc

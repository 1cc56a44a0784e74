"""CPU tests of the HE core against the pure-Python big-integer oracle."""
import random

import pytest
import torch

from hefl_b200 import _ext
from hefl_b200.he import oracle
from hefl_b200.he.context import CKKSContext

ops = _ext.ops()


def test_prime_generation_matches_oracle():
    for logn, bits in [(10, 27), (12, 36), (13, 54), (15, 60)]:
        p = [int(x) for x in ops.gen_primes(bits, logn, 3, [])]
        assert len(set(p)) == 3
        for q in p:
            assert oracle.is_prime(q)
            assert q % (2 << logn) == 1
            assert q.bit_length() == bits


def test_tables_constants():
    logn = 6
    primes = ops.gen_primes(30, logn, 2, [])
    tables, consts = ops.build_tables(primes, logn)
    n = 1 << logn
    for l, q in enumerate(int(x) for x in primes):
        psi = oracle.minimal_psi(q, n)
        c = [int(x) & (2**64 - 1) for x in consts[l]]
        assert c[0] == q and c[5] == psi
        ratio = (c[2] << 64) | c[1]
        assert ratio == (1 << 128) // q
        assert c[3] == pow(n, -1, q)
        assert c[4] == (c[3] << 64) // q
        w = [int(x) & (2**64 - 1) for x in tables[l, 0]]
        wp = [int(x) & (2**64 - 1) for x in tables[l, 1]]
        iw = [int(x) & (2**64 - 1) for x in tables[l, 2]]
        for i in range(n):
            r = oracle.bit_reverse(i, logn)
            assert w[r] == pow(psi, i, q)
            assert wp[r] == (w[r] << 64) // q
            assert iw[r] == pow(psi, -i, q)


@pytest.mark.parametrize("logn", [4, 6, 8])
def test_ntt_matches_definition_and_inverts(logn):
    n = 1 << logn
    primes = ops.gen_primes(40, logn, 2, [])
    tables, consts = ops.build_tables(primes, logn)
    rng = random.Random(logn)
    a = torch.tensor([[rng.randrange(int(q)) for _ in range(n)] for q in primes], dtype=torch.int64)
    x = a.clone()
    ops.ntt_(x, tables, consts, 2, logn, False)
    for l, q in enumerate(int(v) for v in primes):
        psi = int(consts[l, 5])
        assert [int(v) for v in x[l]] == oracle.ntt_by_definition([int(v) for v in a[l]], q, psi)
    ops.ntt_(x, tables, consts, 2, logn, True)
    assert torch.equal(x, a)


@pytest.mark.parametrize("logn", [10, 12, 15])
def test_ntt_roundtrip_large(logn):
    n = 1 << logn
    primes = ops.gen_primes(60, logn, 2, [])
    tables, consts = ops.build_tables(primes, logn)
    g = torch.Generator().manual_seed(logn)
    a = torch.stack([torch.randint(0, int(q), (3, n), generator=g, dtype=torch.int64) for q in primes], dim=1)
    x = a.clone().contiguous()
    ops.ntt_(x, tables, consts, 2, logn, False)
    assert not torch.equal(x, a)
    ops.ntt_(x, tables, consts, 2, logn, True)
    assert torch.equal(x, a)


def test_ntt_product_is_negacyclic_convolution():
    logn, n = 5, 32
    primes = ops.gen_primes(50, logn, 1, [])
    q = int(primes[0])
    tables, consts = ops.build_tables(primes, logn)
    rng = random.Random(7)
    a = [rng.randrange(q) for _ in range(n)]
    b = [rng.randrange(q) for _ in range(n)]
    ta, tb = torch.tensor([a]), torch.tensor([b])
    ops.ntt_(ta, tables, consts, 1, logn, False)
    ops.ntt_(tb, tables, consts, 1, logn, False)
    prod = torch.empty_like(ta)
    ops.pointwise_(prod, ta, tb, 1, consts, 2)
    ops.ntt_(prod, tables, consts, 1, logn, True)
    assert [int(v) for v in prod[0]] == oracle.negacyclic_mul(a, b, q)


def test_pointwise_ops_match_bigint():
    logn, n = 4, 16
    primes = ops.gen_primes(60, logn, 3, [])
    _, consts = ops.build_tables(primes, logn)
    rng = random.Random(3)
    qs = [int(v) for v in primes]
    a = torch.tensor([[[rng.randrange(q) for _ in range(n)] for q in qs] for _ in range(2)])
    b = torch.tensor([[[rng.randrange(q) for _ in range(n)] for q in qs] for _ in range(2)])
    for op, fn in [(0, lambda x, y, q: (x + y) % q), (1, lambda x, y, q: (x - y) % q),
                   (2, lambda x, y, q: x * y % q)]:
        out = torch.empty_like(a)
        ops.pointwise_(out, a, b, 3, consts, op)
        for c in range(2):
            for l, q in enumerate(qs):
                assert [int(v) for v in out[c, l]] == [fn(int(x), int(y), q) for x, y in zip(a[c, l], b[c, l])]
    acc = a.clone()
    ops.pointwise_(acc, a, b, 3, consts, 3)  # acc = a*b + acc
    for c in range(2):
        for l, q in enumerate(qs):
            assert [int(v) for v in acc[c, l]] == [(int(x) * int(y) + int(x)) % q for x, y in zip(a[c, l], b[c, l])]
    # broadcast b over the batch (row % brows)
    out = torch.empty_like(a)
    ops.pointwise_(out, a, b[0].contiguous(), 3, consts, 2)
    assert [int(v) for v in out[1, 2]] == [int(x) * int(y) % qs[2] for x, y in zip(a[1, 2], b[0, 2])]


def test_reduce_mod_full_range():
    logn, n = 4, 16
    primes = ops.gen_primes(59, logn, 2, [])
    _, consts = ops.build_tables(primes, logn)
    qs = [int(v) for v in primes]
    rng = random.Random(11)
    vals = [[rng.randrange(2**64) for _ in range(n)] for _ in qs]
    t = torch.tensor([[v - 2**64 if v >= 2**63 else v for v in row] for row in vals], dtype=torch.int64)
    ops.reduce_mod_(t, 2, consts)
    for l, q in enumerate(qs):
        assert [int(v) for v in t[l]] == [v % q for v in vals[l]]


def test_crt_center_two_limbs():
    ctx = CKKSContext(64, prime_bits=(36, 37), scale_bits=20, enforce_security=False)
    rng = random.Random(5)
    xs = [rng.randrange(-2**70, 2**70) for _ in range(64)]
    res = torch.tensor([[[x % q for x in xs] for q in ctx.primes]], dtype=torch.int64)
    out = ops.crt_center(res, ctx.consts_cpu, ctx.q0_inv_q1)
    for x, y in zip(xs, out[0].tolist()):
        assert oracle.crt_centered([x % ctx.primes[0], x % ctx.primes[1]], ctx.primes) == x
        assert abs(y - x) <= abs(x) * 2**-52

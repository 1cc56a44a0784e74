"""CKKS scheme tests on the CPU backend: codec vs the canonical-embedding oracle,
encrypt/decrypt, additive homomorphism (FedAvg), plaintext multiply + rescale, ct*ct."""
import math

import pytest
import torch

from hefl_b200.he import oracle
from hefl_b200.he.context import CKKSContext, CtBatch


@pytest.fixture(scope="module")
def ctx():
    return CKKSContext(4096, prime_bits=(36, 36, 37), scale_bits=40)


def test_encode_is_inverse_canonical_embedding():
    c = CKKSContext(64, prime_bits=(40,), scale_bits=30, enforce_security=False)
    g = torch.Generator().manual_seed(0)
    vals = torch.randn(32, generator=g, dtype=torch.float64)
    msg = c.encode(vals)
    assert msg.shape == (1, 64)
    slots = oracle.ckks_slots_of([float(v) / c.scale for v in msg[0].tolist()], 64)
    for z, v in zip(slots, vals.tolist()):
        assert abs(z.real - v) < 1e-6 and abs(z.imag) < 1e-6


def test_decode_inverts_encode(ctx):
    g = torch.Generator().manual_seed(1)
    vals = torch.randn(3 * 2048 - 100, generator=g)
    msg = ctx.encode(vals)
    assert msg.shape == (3, 4096)
    back = ctx.decode(msg.double(), ctx.scale).reshape(-1)[: vals.numel()]
    assert (back - vals).abs().max() < 1e-6


def test_slotwise_product_property():
    """decode(a (*) b) == decode(a) * decode(b): pins the embedding to a ring homomorphism."""
    c = CKKSContext(32, prime_bits=(50,), scale_bits=20, enforce_security=False)
    a = torch.linspace(-1, 1, 16, dtype=torch.float64)
    b = torch.linspace(0.5, 2, 16, dtype=torch.float64)
    ma, mb = c.encode(a)[0].tolist(), c.encode(b)[0].tolist()
    big = (1 << 200) + 1
    prod = oracle.negacyclic_mul([int(x) for x in ma], [int(x) for x in mb], big)
    prod = [x - big if x > big // 2 else x for x in prod]
    out = c.decode(torch.tensor([prod], dtype=torch.float64), c.scale * c.scale).reshape(-1)
    assert (out - a * b).abs().max() < 1e-4


def test_encrypt_decrypt_roundtrip(ctx):
    sk, pk = ctx.keygen(seed=42)
    g = torch.Generator().manual_seed(2)
    vals = torch.randn(5000, generator=g)
    ct = ctx.encrypt(vals, pk, seed=7)
    assert ct.data.shape == (3, 2, 3, 4096)
    assert int(ct.data.min()) >= 0
    for l, q in enumerate(ctx.primes):
        assert int(ct.data[:, :, l].max()) < q
    out = ctx.decrypt(ct, sk)
    assert out.shape == (5000,)
    assert (out - vals).abs().max() < 1e-5


def test_encryption_is_randomised_and_seeded(ctx):
    sk, pk = ctx.keygen(seed=1)
    vals = torch.ones(100)
    a = ctx.encrypt(vals, pk, seed=1)
    b = ctx.encrypt(vals, pk, seed=1)
    c = ctx.encrypt(vals, pk, seed=2)
    assert torch.equal(a.data, b.data)
    assert not torch.equal(a.data, c.data)


def test_fedavg_of_eight_clients(ctx):
    """Dec(sum_p Enc(w_p)) / K == mean_p w_p (FLPyfhelin.py:381-385)."""
    sk, pk = ctx.keygen(seed=3)
    g = torch.Generator().manual_seed(3)
    K = 8
    ws = [torch.randn(4500, generator=g) * 0.1 for _ in range(K)]
    cts = [ctx.encrypt(w, pk, seed=100 + i) for i, w in enumerate(ws)]
    agg = ctx.sum_batches(cts)
    out = ctx.decrypt(agg, sk, divide_by=K)
    ref = torch.stack(ws).mean(0)
    assert (out - ref).abs().max() < 1e-5


def test_coefficient_packing_roundtrip(ctx):
    sk, pk = ctx.keygen(seed=4)
    vals = torch.linspace(-2, 2, 6000)
    ct = ctx.encrypt(vals, pk, seed=9, packing="coeff")
    assert ct.count == 2
    out = ctx.decrypt(ct, sk)
    assert (out - vals).abs().max() < 1e-5


def test_mul_scalar_with_rescale(ctx):
    sk, pk = ctx.keygen(seed=5)
    vals = torch.linspace(-1, 1, 2048)
    ct = ctx.encrypt(vals, pk, seed=11)
    ctx.mul_scalar_(ct, 1.0 / 3.0)
    assert ct.level == 2 and math.isclose(ct.scale, ctx.scale, rel_tol=1e-12)
    out = ctx.decrypt(ct, sk)
    assert (out - vals / 3).abs().max() < 1e-4


def test_add_plain_and_negate(ctx):
    sk, pk = ctx.keygen(seed=6)
    vals = torch.linspace(0, 1, 2048)
    ct = ctx.encrypt(vals, pk, seed=12)
    ctx.add_plain_(ct, torch.full((2048,), 0.25))
    ctx.negate_(ct)
    out = ctx.decrypt(ct, sk)
    assert (out + vals + 0.25).abs().max() < 1e-5


def test_mul_plain_vector(ctx):
    sk, pk = ctx.keygen(seed=7)
    a = torch.linspace(-1, 1, 2048)
    b = torch.linspace(0.5, 1.5, 2048)
    ct = ctx.encrypt(a, pk, seed=13)
    ctx.mul_plain_(ct, b)
    out = ctx.decrypt(ct, sk)
    assert (out - a * b).abs().max() < 1e-3


def test_ct_ct_multiply_with_relinearisation(ctx):
    sk, pk = ctx.keygen(seed=8)
    rlk = ctx.relin_keygen(sk, seed=8, digit_bits=12)
    a = torch.linspace(-1, 1, 2048)
    b = torch.linspace(0.5, 1.5, 2048)
    ca = ctx.encrypt(a, pk, seed=14)
    cb = ctx.encrypt(b, pk, seed=15)
    prod = ctx.multiply(ca, cb, rlk)
    assert prod.level == 2
    out = ctx.decrypt(prod, sk)
    assert (out - a * b).abs().max() < 5e-3


def test_serialization_roundtrip(ctx):
    sk, pk = ctx.keygen(seed=9)
    vals = torch.randn(3000)
    ct = ctx.encrypt(vals, pk, seed=16)
    blob = ctx.ct_to_bytes(ct)
    assert blob[:4] == b"HEFL"
    ctx2 = CKKSContext.from_bytes_context(ctx.to_bytes_context())
    assert ctx2.primes == ctx.primes and ctx2.n == ctx.n
    ct2 = ctx2.ct_from_bytes(blob)
    assert torch.equal(ct2.data, ct.data) and ct2.nvals == 3000
    out = ctx2.decrypt(ct2, sk)
    assert (out - vals).abs().max() < 1e-5
    other = CKKSContext(4096, prime_bits=(35, 36, 37), scale_bits=40)
    with pytest.raises(ValueError):
        other.ct_from_bytes(blob)


def test_security_bound_enforced():
    with pytest.raises(ValueError):
        CKKSContext(4096, prime_bits=(40, 40, 40), scale_bits=30)
    CKKSContext(4096, prime_bits=(40, 40, 40), scale_bits=30, enforce_security=False)


@pytest.mark.parametrize("preset", ["n8192_l4", "n16384_l4"])
def test_larger_presets(preset):
    from hefl_b200.config import HE_PRESETS

    p = HE_PRESETS[preset]
    c = CKKSContext(p["n"], prime_bits=p["prime_bits"], scale_bits=p["scale_bits"])
    sk, pk = c.keygen(seed=1)
    vals = torch.randn(c.n // 2 + 17)
    ct = c.encrypt(vals, pk, seed=2)
    out = c.decrypt(ct, sk)
    assert (out - vals).abs().max() < 1e-6


def test_evaluation_keys_satisfy_their_defining_equation(ctx):
    """Every digit key of relin_keygen (one batched keygen call + the message-term kernel) must satisfy
    b + a*s = e + 2^(k*w) * g_i * s^2 with a small error e: checked limb by limb against Python integers."""
    from hefl_b200 import _ext

    ops = _ext.ops()
    sk, _ = ctx.keygen(seed=31)
    rlk = ctx.relin_keygen(sk, seed=32, digit_bits=11)
    t, c = ctx._cpu["tables"], ctx._cpu["consts"]
    s2 = torch.empty_like(sk)
    ops.pointwise_(s2, sk, sk, ctx.L, c, 2)
    for i, keys in enumerate(rlk.keys):
        assert keys.shape[0] == -(-ctx.primes[i].bit_length() // 11)
        for k in range(keys.shape[0]):
            b, a = keys[k, 0].clone(), keys[k, 1]
            ops.pointwise_(b, a, sk, ctx.L, c, 3)                    # b + a*s (NTT domain)
            ops.ntt_(b, t, c, ctx.L, ctx.logn, True)
            s2c = s2.clone()
            ops.ntt_(s2c, t, c, ctx.L, ctx.logn, True)
            for l, q in enumerate(ctx.primes):
                w = pow(2, k * 11, q) if l == i else 0
                row = [int(v) for v in b[l, :64].tolist()]
                ref = [int(v) for v in s2c[l, :64].tolist()]
                for x, r in zip(row, ref):
                    e = (x - w * r) % q
                    e = e - q if e > q // 2 else e
                    assert abs(e) <= 21, (i, k, l, e)               # centred binomial, 21 coin pairs


def test_batched_public_keys_equal_one_call_per_index(ctx):
    from hefl_b200 import _ext

    ops = _ext.ops()
    sk, _ = ctx.keygen(seed=33)
    t, c = ctx._cpu["tables"], ctx._cpu["consts"]
    batch = ops.keygen_public_batch(sk, ctx.L, ctx.logn, t, c, 34, 5, 3)
    for e in range(3):
        assert torch.equal(batch[e], ops.keygen_public(sk, ctx.L, ctx.logn, t, c, 34, 5 + e))
    assert not torch.equal(batch[0], batch[1])

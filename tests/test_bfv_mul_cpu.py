"""BFV relinearisation keys and ciphertext x ciphertext (the scheme the reference runs; FLPyfhelin.py:357-364
``gen_rekey`` -> ``relinKeyGen(bitCount=1, size=5)``) against a big-integer oracle."""
import numpy as np
import pytest
import torch

from hefl_b200.compat.pyfhel_shim import Pyfhel
from hefl_b200.he.bfv import BFVFracContext


def _negacyclic(a, b, n, p):
    out = [0] * n
    for i, x in enumerate(a):
        if x == 0:
            continue
        for j, y in enumerate(b):
            if y == 0:
                continue
            k = i + j
            if k < n:
                out[k] += x * y
            else:
                out[k - n] -= x * y
    half = p // 2
    return [((v + half) % p) - half for v in out]


@pytest.mark.parametrize("bit_count", [1, 8, 16])   # base 2^30 would exceed the noise budget of a 54-bit q
def test_product_digits_equal_the_negacyclic_convolution_of_the_plaintexts(bit_count):
    ctx = BFVFracContext(p=65537, m=2048, sec=128)
    sk, pk = ctx.keygen(seed=11)
    rlk = ctx.relin_keygen(sk, seed=12, bit_count=bit_count, size=5)
    assert rlk.keys.shape[0] == -(-ctx.q.bit_length() // bit_count) and rlk.size == 5
    va = torch.tensor([1.5, -2.25, 7.0], dtype=torch.float64)
    vb = torch.tensor([-0.75, 3.0, 0.5], dtype=torch.float64)
    ca, cb = ctx.encrypt(va, pk, seed=1), ctx.encrypt(vb, pk, seed=2)
    prod = ctx.multiply(ca, cb, rlk)
    assert prod.shape == ca.shape                                   # relinearised back to two polynomials
    got = ctx.decrypt_digits(prod, sk)
    ma, mb = ctx.encode(va), ctx.encode(vb)
    for c in range(3):
        want = _negacyclic(ma[c].tolist(), mb[c].tolist(), ctx.n, ctx.p)
        assert got[c].tolist() == want
    assert torch.allclose(ctx.decrypt(prod, sk), va * vb, atol=1e-9)
    assert ctx.noise_budget_bits(prod, sk) > 0


def test_reference_gen_rekey_call_shape_and_pyctxt_product():
    he = Pyfhel()
    he.contextGen(p=65537, sec=128, m=2048)
    he.keyGen()
    with pytest.raises(RuntimeError):
        _ = he.encryptFrac(1.0) * he.encryptFrac(2.0)
    he.relinKeyGen(bitCount=1, size=5)                              # the reference's arguments
    assert "rlk:Y" in repr(he)
    a, b = he.encryptFrac(3.0), he.encryptFrac(0.125)
    assert abs(he.decryptFrac(a * b) - 0.375) < 1e-9
    assert abs(he.decryptFrac(a * b + a) - 3.375) < 1e-9
    arr = np.array([he.encryptFrac(v) for v in (1.0, -2.0)], dtype=object)
    out = arr * he.encryptFrac(0.5)                                 # NumPy object broadcasting, as the reference uses it
    assert [round(he.decryptFrac(c), 9) for c in out] == [0.5, -1.0]

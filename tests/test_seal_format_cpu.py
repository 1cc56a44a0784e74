"""Serialization: (a) golden bytes of our own HEFL / HEFB streams, (b) the SEAL-2.3-layout streams
(``format="seal2"``) against fixtures built BY HAND from the documented field order, (c) round trips through the
Pyfhel façade. Real Pyfhel 2.3.1 files cannot be produced offline: layout conformance is checked, interop is not."""
import hashlib
import os
import pickle
import struct

import numpy as np
import torch

from hefl_b200.compat import seal_format as sf
from hefl_b200.compat.pyfhel_shim import Pyfhel, PyCtxt
from hefl_b200.he.bfv import BFVFracContext
from hefl_b200.he.context import CKKSContext

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _hand_built_params(n, q, p):
    """EncryptionParameters stream written field by field, independently of seal_format's writer."""
    b = bytearray()
    b += struct.pack("<i", n + 1) + struct.pack("<i", 1)                  # BigPoly: coeff_count, coeff_bit_count
    b += (1).to_bytes(8, "little") + b"\x00" * (8 * (n - 1)) + (1).to_bytes(8, "little")
    b += struct.pack("<i", 1) + struct.pack("<Q", q)                      # one coefficient modulus
    b += struct.pack("<Q", p)                                             # plain modulus
    b += struct.pack("<d", 3.19) + struct.pack("<d", 15.95)              # noise std, noise max
    return bytes(b)


def test_seal2_parameter_stream_matches_the_hand_built_fixture():
    n, q, p = 1024, 132120577, 65537
    want = _hand_built_params(n, q, p)
    assert sf.params_to_bytes(n, [q], p) == want
    assert len(want) == 8 + 8 * (n + 1) + 4 + 8 + 8 + 16
    d, off = sf.params_from_bytes(want)
    assert off == len(want) and d["n"] == n and d["coeff_moduli"] == [q] and d["plain_modulus"] == p
    # the stored fixture (committed bytes) must never drift
    path = os.path.join(GOLD, "seal2_params_n1024.bin")
    if not os.path.exists(path):
        open(path, "wb").write(want)
    assert open(path, "rb").read() == want
    assert hashlib.sha256(want).hexdigest() == hashlib.sha256(open(path, "rb").read()).hexdigest()


def test_seal2_ciphertext_stream_layout():
    n = 8
    hb = bytes(range(32))
    polys = np.arange(2 * 1 * n, dtype=np.int64).reshape(2, 1, n) + 100
    buf = sf.polys_to_bytes(hb, polys)
    # header: hash block, size=2, poly_coeff_count=N+1, coeff_mod_count=1
    assert buf[:32] == hb and struct.unpack_from("<iii", buf, 32) == (2, n + 1, 1)
    words = np.frombuffer(buf, dtype="<u8", offset=44)
    assert words.size == 2 * (n + 1) and words[n] == 0 and words[2 * n + 1] == 0          # unused top coefficients
    assert list(words[:n]) == list(range(100, 100 + n))
    hb2, back, off = sf.polys_from_bytes(buf)
    assert hb2 == hb and off == len(buf) and np.array_equal(back, polys)


def test_pyfhel_facade_round_trips_in_the_seal2_layout(tmp_path):
    he = Pyfhel()
    he.contextGen(p=65537, sec=128, m=1024)
    he.keyGen()
    con = he.to_bytes_context(format="seal2")
    assert con[:8] == struct.pack("<ii", 1025, 1)                        # BigPoly header of x^1024 + 1
    pk, sk = he.to_bytes_publicKey(format="seal2"), he.to_bytes_secretKey(format="seal2")
    assert pk[:32] == sk[:32] == sf.params_hash(1024, [he._ctx.q], 65537)
    assert len(pk) == 32 + 12 + 2 * 1025 * 8 and len(sk) == 32 + 8 + 1025 * 8
    ct = he.encryptFrac(-1.375)
    blob = he.ctxt_to_bytes(ct, format="seal2")
    assert len(blob) == 32 + 12 + 2 * 1025 * 8
    # a second party rebuilds everything from the streams alone (the reference's get_sk sequence, FLPyfhelin.py:251-261)
    other = pickle.loads(pickle.dumps(he))
    other.from_bytes_context(con)
    other.from_bytes_publicKey(pk)
    other.from_bytes_secretKey(sk)
    assert torch.equal(other._pk, he._pk) and torch.equal(other._sk, he._sk)
    back = other.ctxt_from_bytes(blob)
    assert torch.equal(back._data, ct._data) and abs(other.decryptFrac(back) + 1.375) < 1e-9
    # the native format still round-trips and is told apart automatically
    third = Pyfhel()
    third.from_bytes_context(he.to_bytes_context())
    third.from_bytes_publicKey(he.to_bytes_publicKey())
    assert torch.equal(third._pk, he._pk)


def test_native_streams_have_golden_bytes():
    """HEFL (CKKS) and HEFB (BFV) headers and payload order are part of the on-disk contract."""
    ctx = CKKSContext(1024, primes=[132120577, 131923969], scale_bits=20, enforce_security=False)
    got = ctx.to_bytes_context()
    want = (b"HEFL" + struct.pack("<HHIII", 1, 1, 1024, 2, 20) + struct.pack("<2Q", 132120577, 131923969)
            + struct.pack("<I", 1) + struct.pack("<q", 128))
    assert got == want
    t = torch.arange(2 * 2 * 1024, dtype=torch.int64).reshape(1, 2, 2, 1024) % 1000
    from hefl_b200.he.context import CtBatch
    blob = ctx.ct_to_bytes(CtBatch(t, 2.0 ** 20, 5, "slots"))
    head = (b"HEFL" + struct.pack("<HHIII", 1, 4, 1024, 2, 20) + struct.pack("<2Q", 132120577, 131923969)
            + struct.pack("<I", 8) + struct.pack("<8q", 4, 1, 2, 2, 1024, struct.unpack("<q", struct.pack("<d", 2.0 ** 20))[0], 5, 0))
    assert blob[:len(head)] == head and blob[len(head):] == t.numpy().tobytes()
    path = os.path.join(GOLD, "hefl_ct_header.bin")
    if not os.path.exists(path):
        open(path, "wb").write(head)
    assert open(path, "rb").read() == head
    b = BFVFracContext(p=65537, m=1024, sec=128)
    assert b.to_bytes_context() == b"HEFB" + struct.pack("<HIIIIIIQ", 1, 65537, 1024, 128, 2, 64, 32, b.q)

"""SEAL-2.3-layout binary streams for parameters, keys and ciphertexts (``format="seal2"``).

The reference's wire format is whatever Pyfhel 2.3.1 -> Afseal -> SEAL 2.3 writes:
``to_bytes_context / to_bytes_publicKey / to_bytes_secretKey`` (FLPyfhelin.py:337-338, :257-259; notebook N:57-59)
and the bytes inside every pickled ``PyCtxt`` (:237, :309). Neither Pyfhel nor SEAL can be installed here, so the
layout below is written from the SEAL 2.3 sources as we know them and is **not verified against real files**;
what IS guaranteed (tests/test_seal_format_cpu.py, golden fixtures under tests/golden/) is that our writer and
reader agree byte for byte with the documented field order:

    BigPoly            int32 coeff_count | int32 coeff_bit_count | coeff_count * ceil(bits/64) little-endian u64
    SmallModulus       u64 value
    EncryptionParameters
                       BigPoly poly_modulus (x^N + 1: coeff_count N+1, bit count 1)
                       int32 coeff_mod_count | SmallModulus * count
                       SmallModulus plain_modulus
                       f64 noise_standard_deviation | f64 noise_max_deviation
    hash_block         4 * u64: SHA3-256 over the u64 words of the parameters, in the order above
    Ciphertext         hash_block | int32 size | int32 poly_coeff_count (N + 1) | int32 coeff_mod_count |
                       size * coeff_mod_count * poly_coeff_count u64 (coefficient form, top coefficient 0)
    PublicKey          hash_block | Ciphertext-shaped body of 2 polynomials
    SecretKey          hash_block | int32 poly_coeff_count | int32 coeff_mod_count | one polynomial

Known gaps (stated in README): SEAL keeps keys in its own NTT ordering with its own choice of primitive root;
we store coefficient form in these streams, so a real SEAL would read the container but not use the key. The
integer / fraction digit counts and the base that Afseal appends after the parameters are written as five
little-endian int32 (base, sec, intDigits, fracDigits, flagBatching).
"""
from __future__ import annotations

import hashlib
import struct
from typing import List, Sequence, Tuple

import numpy as np

NOISE_STD = 3.19
NOISE_MAX = 5 * 3.19


def _words(bits: int) -> int:
    return (bits + 63) // 64


def write_bigpoly(coeffs: Sequence[int], coeff_bit_count: int) -> bytes:
    w = _words(coeff_bit_count)
    out = [struct.pack("<ii", len(coeffs), coeff_bit_count)]
    for c in coeffs:
        out.append(int(c).to_bytes(8 * w, "little"))
    return b"".join(out)


def read_bigpoly(buf: bytes, off: int) -> Tuple[List[int], int, int]:
    n, bits = struct.unpack_from("<ii", buf, off)
    off += 8
    w = _words(bits)
    coeffs = [int.from_bytes(buf[off + 8 * w * i: off + 8 * w * (i + 1)], "little") for i in range(n)]
    return coeffs, bits, off + 8 * w * n


def params_to_bytes(n: int, coeff_moduli: Sequence[int], plain_modulus: int, noise_std: float = NOISE_STD,
                    noise_max: float = NOISE_MAX) -> bytes:
    poly = [1] + [0] * (n - 1) + [1]                          # x^N + 1, constant term first
    out = [write_bigpoly(poly, 1), struct.pack("<i", len(coeff_moduli))]
    out += [struct.pack("<Q", int(q)) for q in coeff_moduli]
    out.append(struct.pack("<Q", int(plain_modulus)))
    out.append(struct.pack("<dd", noise_std, noise_max))
    return b"".join(out)


def params_from_bytes(buf: bytes, off: int = 0):
    poly, bits, off = read_bigpoly(buf, off)
    n = len(poly) - 1
    if bits != 1 or poly[0] != 1 or poly[n] != 1 or any(poly[1:n]):
        raise ValueError("poly_modulus is not x^N + 1")
    (k,) = struct.unpack_from("<i", buf, off)
    off += 4
    moduli = list(struct.unpack_from(f"<{k}Q", buf, off))
    off += 8 * k
    (plain,) = struct.unpack_from("<Q", buf, off)
    off += 8
    std, mx = struct.unpack_from("<dd", buf, off)
    off += 16
    return dict(n=n, coeff_moduli=moduli, plain_modulus=plain, noise_std=std, noise_max=mx), off


def params_hash(n: int, coeff_moduli: Sequence[int], plain_modulus: int, noise_std: float = NOISE_STD,
                noise_max: float = NOISE_MAX) -> bytes:
    """32 bytes = hash_block (4 u64): SHA3-256 over the parameter words in stream order."""
    words = [1] + [0] * (n - 1) + [1] + [int(q) for q in coeff_moduli] + [int(plain_modulus)]
    blob = b"".join(struct.pack("<Q", w) for w in words) + struct.pack("<dd", noise_std, noise_max)
    return hashlib.sha3_256(blob).digest()


def polys_to_bytes(hash_block: bytes, polys: np.ndarray) -> bytes:
    """polys: uint64/int64 [size][coeff_mod_count][N] in coefficient form -> Ciphertext stream."""
    a = np.ascontiguousarray(polys).astype(np.uint64, copy=False)
    size, k, n = a.shape
    body = np.zeros((size, k, n + 1), dtype="<u8")          # SEAL 2.3 polynomials carry N + 1 coefficient slots
    body[:, :, :n] = a
    return hash_block + struct.pack("<iii", size, n + 1, k) + body.tobytes()


def polys_from_bytes(buf: bytes, off: int = 0) -> Tuple[bytes, np.ndarray, int]:
    hb = bytes(buf[off:off + 32])
    size, cc, k = struct.unpack_from("<iii", buf, off + 32)
    off += 44
    cnt = size * k * cc
    a = np.frombuffer(buf, dtype="<u8", count=cnt, offset=off).reshape(size, k, cc)
    if a[:, :, cc - 1].any():
        raise ValueError("top coefficient slot of a ciphertext polynomial must be zero")
    return hb, a[:, :, :cc - 1].astype(np.int64).copy(), off + 8 * cnt


def secret_to_bytes(hash_block: bytes, poly: np.ndarray) -> bytes:
    a = np.ascontiguousarray(poly).astype(np.uint64, copy=False)       # [coeff_mod_count][N]
    k, n = a.shape
    body = np.zeros((k, n + 1), dtype="<u8")
    body[:, :n] = a
    return hash_block + struct.pack("<ii", n + 1, k) + body.tobytes()


def secret_from_bytes(buf: bytes, off: int = 0) -> Tuple[bytes, np.ndarray, int]:
    hb = bytes(buf[off:off + 32])
    cc, k = struct.unpack_from("<ii", buf, off + 32)
    off += 40
    a = np.frombuffer(buf, dtype="<u8", count=k * cc, offset=off).reshape(k, cc)
    return hb, a[:, :cc - 1].astype(np.int64).copy(), off + 8 * k * cc


def context_to_bytes(n: int, coeff_moduli: Sequence[int], plain_modulus: int, base: int, sec: int, int_digits: int,
                     frac_digits: int, batching: bool) -> bytes:
    """Afseal::saveContext: the SEAL parameters followed by the encoder settings Pyfhel keeps next to them."""
    return params_to_bytes(n, coeff_moduli, plain_modulus) + struct.pack("<iiiii", base, sec, int_digits, frac_digits,
                                                                         1 if batching else 0)


def context_from_bytes(buf: bytes):
    p, off = params_from_bytes(buf, 0)
    base, sec, idg, fdg, bat = struct.unpack_from("<iiiii", buf, off)
    p.update(base=base, sec=sec, int_digits=idg, frac_digits=fdg, batching=bool(bat))
    return p

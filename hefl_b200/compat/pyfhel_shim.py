"""Pyfhel-compatible objects (``Pyfhel``, ``PyCtxt``, ``PyPtxt``) on top of hefl_b200.

Covers the exact surface the reference uses: the Pyfhel 2.3.1 key
generation call ``contextGen(p=65537, sec=s, m=m)`` (FLPyfhelin.py:332), ``keyGen``,
``encryptFrac``/``decryptFrac`` on scalars (:217, :295, :371), ``to_bytes_*`` /
``from_bytes_*`` (:257-259, :337-338, :352-353), ``relinKeyGen(bitCount, size)`` (:363),
``PyCtxt + int``, ``PyCtxt + PyCtxt`` (:381), ``PyCtxt * float`` (:385), a writable
``_pyfhel`` attribute (:321), and pickling that DROPS the context (which is why the reference
re-attaches it after every load, :320-321). The ``repr`` matches notebook N:44.

It also accepts the Pyfhel 3.x spelling (README R:7: ``m`` became ``n``) and a packed CKKS
mode — ``contextGen(scheme='CKKS', n=..., scale=..., qi_sizes=[...])`` with array
``encryptFrac``/``decryptFrac`` — which is what BASELINE.json configs[0] ("Pyfhel-CKKS on CPU")
asks for.

Byte-level compatibility with SEAL streams cannot be verified offline (no Pyfhel/SEAL here);
streams use the versioned hefl_b200 format (magic ``HEFB``/``HEFL``).
"""
from __future__ import annotations

import struct
from typing import Optional, Sequence, Union

import os
import secrets

import numpy as np
import torch

from . import seal_format
from ..he.bfv import BFVFracContext
from ..he.context import CKKSContext, CtBatch

ENC_FRACTIONAL = "FRACTIONAL"
ENC_CKKS = "CKKS"


class PyPtxt:
    """Plaintext handle (imported but unused by the reference, FLPyfhelin.py:27)."""

    def __init__(self, value=None, pyfhel=None, encoding: str = ENC_FRACTIONAL):
        self.value = value
        self._pyfhel = pyfhel
        self._encoding = encoding

    def __repr__(self):
        return f"<PyPtxt {self._encoding} {self.value!r}>"


def _rebuild_ctxt(blob: bytes, encoding: str, meta: tuple):
    ct = PyCtxt.__new__(PyCtxt)
    shape = meta[0]
    ct._data = torch.from_numpy(np.frombuffer(blob, dtype=np.int64).reshape(shape).copy())
    ct._encoding = encoding
    ct._scale = meta[1]
    ct._nvals = meta[2]
    ct._pyfhel = None          # context is NOT carried by the pickle (FLPyfhelin.py:320-321)
    return ct


class PyCtxt:
    """One ciphertext: int64 words ``[2, L, N]`` (fractional) or a packed CKKS batch ``[C,2,L,N]``."""

    __slots__ = ("_data", "_encoding", "_scale", "_nvals", "_pyfhel")

    def __init__(self, pyfhel: Optional["Pyfhel"] = None, data: Optional[torch.Tensor] = None,
                 encoding: str = ENC_FRACTIONAL, scale: float = 1.0, nvals: int = 1):
        self._pyfhel = pyfhel
        self._data = data
        self._encoding = encoding
        self._scale = scale
        self._nvals = nvals

    # -- helpers -----------------------------------------------------------------------------
    def _he(self) -> "Pyfhel":
        if self._pyfhel is None:
            raise RuntimeError("PyCtxt has no Pyfhel context attached; set ct._pyfhel = HE first")
        return self._pyfhel

    def _like(self, data: torch.Tensor, scale: Optional[float] = None) -> "PyCtxt":
        return PyCtxt(self._pyfhel, data, self._encoding, self._scale if scale is None else scale, self._nvals)

    def size(self) -> int:
        return 2

    def to_bytes(self) -> bytes:
        return self._data.cpu().contiguous().numpy().tobytes()

    # -- arithmetic --------------------------------------------------------------------------
    def __add__(self, other):
        he = self._he()
        if isinstance(other, PyCtxt):
            return self._like(he._add(self, other))
        if isinstance(other, (int, float, np.integer, np.floating)):
            if other == 0:
                return self._like(self._data.clone())       # add_plain(0): X1.g, elided
            return self._like(he._add_plain(self, float(other)))
        return NotImplemented

    __radd__ = __add__

    def __sub__(self, other):
        he = self._he()
        if isinstance(other, PyCtxt):
            return self._like(he._sub(self, other))
        if isinstance(other, (int, float, np.integer, np.floating)):
            return self._like(he._add_plain(self, -float(other)))
        return NotImplemented

    def __mul__(self, other):
        he = self._he()
        if isinstance(other, (int, float, np.integer, np.floating)):
            data, scale = he._mul_plain(self, float(other))
            return self._like(data, scale)
        if isinstance(other, PyCtxt):
            data, scale = he._mul_ct(self, other)
            return self._like(data, scale)
        return NotImplemented

    __rmul__ = __mul__

    def __neg__(self):
        return self * -1.0

    # -- pickle ------------------------------------------------------------------------------
    def __reduce__(self):
        return _rebuild_ctxt, (self.to_bytes(), self._encoding,
                               (tuple(self._data.shape), self._scale, self._nvals))

    def __repr__(self):
        att = "attached" if self._pyfhel is not None else "detached"
        return f"<PyCtxt {self._encoding} shape={tuple(self._data.shape)} {att}>"


class Pyfhel:
    """Pyfhel-like façade. An instance may be empty (fresh or unpickled) and is rehydrated with
    ``from_bytes_context`` / ``from_bytes_publicKey`` / ``from_bytes_secretKey``."""

    def __init__(self, device: Union[str, torch.device] = "cpu"):
        self._device = torch.device(device)
        self._ctx = None               # BFVFracContext | CKKSContext
        self._scheme = None
        self._pk = None
        self._sk = None
        self._rlk = None
        # Encryption randomness: a fresh OS-entropy seed per instance (a rehydrated get_pk() object is a new
        # instance, so two clients never draw the same (u, e0, e1)); never derived from the key seed or a
        # constant -- with predictable randomness anybody holding pk recovers m = c0 - pk0*u - e0.
        self._seed_counter = 0
        self._enc_seed = secrets.randbits(63)

    # ------------------------------------------------------------------ context / keys
    def contextGen(self, p: int = 65537, m: int = 2048, flagBatching: bool = False, base: int = 2,
                   sec: int = 128, intDigits: int = 64, fracDigits: int = 32, *, scheme: Optional[str] = None,
                   n: Optional[int] = None, scale: Optional[float] = None, scale_bits: Optional[int] = None,
                   qi_sizes: Optional[Sequence[int]] = None, **_ignored) -> None:
        if n is not None:              # Pyfhel 3.x spelling (README R:7)
            m = n
        sch = (scheme or "BFV").upper()
        if sch in ("CKKS",):
            bits = list(qi_sizes) if qi_sizes is not None else [54]
            sb = scale_bits if scale_bits is not None else (int(round(np.log2(scale))) if scale else 40)
            self._ctx = CKKSContext(m, prime_bits=bits, scale_bits=sb, device=self._device, sec=sec)
            self._scheme = ENC_CKKS
        else:
            if flagBatching:
                raise NotImplementedError("BFV batching is not part of the reference surface")
            self._ctx = BFVFracContext(p=p, m=m, sec=sec, base=base, int_digits=intDigits,
                                       frac_digits=fracDigits, device=self._device)
            self._scheme = ENC_FRACTIONAL
        self._pk = self._sk = self._rlk = None

    def keyGen(self, seed: Optional[int] = None) -> None:
        """Key pair from OS entropy (the reference seeds nothing, Q12). ``seed`` exists for reproducible tests
        only; it does not touch the encryption randomness (see ``seed_encryption``)."""
        self._need_ctx()
        if seed is None:
            seed = secrets.randbits(63)
        self._sk, self._pk = self._ctx.keygen(seed=seed)

    def seed_encryption(self, seed: int) -> None:
        """TESTS ONLY: make the following encryptions reproducible."""
        self._enc_seed = int(seed) & 0x7FFFFFFFFFFFFFFF
        self._seed_counter = 0

    def relinKeyGen(self, bitCount: int = 16, size: int = 1) -> None:
        """SEAL-2.x signature (decomposition bit count, key size); FLPyfhelin.py:362-363."""
        self._need_ctx()
        if self._sk is None:
            raise RuntimeError("relinKeyGen needs the secret key")
        if self._scheme != ENC_CKKS:
            self._rlk = self._ctx.relin_keygen(self._sk, seed=secrets.randbits(63), bit_count=int(bitCount), size=int(size))
            return
        self._rlk = self._ctx.relin_keygen(self._sk, seed=secrets.randbits(63), digit_bits=max(1, min(int(bitCount), 30)))

    def rotateKeyGen(self, *a, **k):
        raise NotImplementedError("rotations are not used by the reference (repr shows rtk:-)")

    # ------------------------------------------------------------------ (de)serialization
    def _fmt(self, format: Optional[str]) -> str:
        f = (format or os.environ.get("HEFL_SERIALIZATION", "native")).lower()
        if f not in ("native", "seal2"):
            raise ValueError(f"unknown serialization format {f!r}")
        if f == "seal2" and self._scheme != ENC_FRACTIONAL:
            raise NotImplementedError("the SEAL-2.3 layout exists for the BFV/fractional scheme of the reference")
        return f

    def _seal_hash(self) -> bytes:
        c = self._ctx
        return seal_format.params_hash(c.n, [c.q], c.p)

    def _to_coeff(self, t: torch.Tensor, inverse: bool) -> torch.Tensor:
        """[..., 1, N] between NTT form (ours) and coefficient form (the seal2 streams)."""
        x = t.detach().to(self._device).reshape(-1, self._ctx.n).clone().contiguous()
        self._ctx.ops.ntt_(x, self._ctx.tables, self._ctx.consts, 1, self._ctx.logn, inverse)
        return x.reshape(t.shape)

    def to_bytes_context(self, format: Optional[str] = None) -> bytes:
        """``format="seal2"`` (or HEFL_SERIALIZATION=seal2): SEAL-2.3 EncryptionParameters layout followed by the
        encoder settings Afseal keeps next to them (compat/seal_format.py); default: our versioned stream."""
        self._need_ctx()
        if self._fmt(format) == "seal2":
            c = self._ctx
            return seal_format.context_to_bytes(c.n, [c.q], c.p, c.base, c.sec, c.int_digits, c.frac_digits, False)
        tag = b"C" if self._scheme == ENC_CKKS else b"F"
        return tag + self._ctx.to_bytes_context()

    def from_bytes_context(self, buf: bytes) -> None:
        if buf[:1] not in (b"C", b"F"):
            # SEAL-2.3 layout: starts with BigPoly(x^N + 1) = int32 coeff_count (N + 1), int32 bit count (1)
            d = seal_format.context_from_bytes(buf)
            if len(d["coeff_moduli"]) != 1:
                raise NotImplementedError("seal2 contexts with several coefficient moduli")
            self._ctx = BFVFracContext(p=d["plain_modulus"], m=d["n"], sec=d["sec"], base=d["base"],
                                       int_digits=d["int_digits"], frac_digits=d["frac_digits"], device=self._device,
                                       q=d["coeff_moduli"][0])
            self._scheme = ENC_FRACTIONAL
            return
        if buf[:1] == b"C":
            self._ctx = CKKSContext.from_bytes_context(buf[1:], device=self._device)
            self._scheme = ENC_CKKS
        elif buf[:1] == b"F":
            self._ctx = BFVFracContext.from_bytes_context(buf[1:], device=self._device)
            self._scheme = ENC_FRACTIONAL
        else:
            raise ValueError("unknown context stream")

    def _key_bytes(self, t: torch.Tensor, kind: bytes) -> bytes:
        shape = tuple(t.shape)
        return kind + struct.pack("<I", len(shape)) + struct.pack(f"<{len(shape)}q", *shape) + \
            t.cpu().contiguous().numpy().tobytes()

    def _key_from(self, buf: bytes, kind: bytes) -> torch.Tensor:
        if buf[:1] != kind:
            raise ValueError("wrong key stream kind")
        (nd,) = struct.unpack_from("<I", buf, 1)
        shape = struct.unpack_from(f"<{nd}q", buf, 5)
        arr = np.frombuffer(buf, dtype=np.int64, offset=5 + 8 * nd).reshape(shape).copy()
        return torch.from_numpy(arr).to(self._device)

    def to_bytes_publicKey(self, format: Optional[str] = None) -> bytes:
        if self._pk is None:
            raise RuntimeError("no public key")
        if self._fmt(format) == "seal2":
            return seal_format.polys_to_bytes(self._seal_hash(), self._to_coeff(self._pk, True).cpu().numpy())
        return self._key_bytes(self._pk, b"P")

    def to_bytes_secretKey(self, format: Optional[str] = None) -> bytes:
        if self._sk is None:
            raise RuntimeError("no secret key")
        if self._fmt(format) == "seal2":
            return seal_format.secret_to_bytes(self._seal_hash(), self._to_coeff(self._sk, True).cpu().numpy())
        return self._key_bytes(self._sk, b"S")

    def _is_seal(self, buf: bytes) -> bool:
        return self._scheme == ENC_FRACTIONAL and bytes(buf[:32]) == self._seal_hash()

    def from_bytes_publicKey(self, buf: bytes) -> None:
        self._need_ctx()
        if self._is_seal(buf):
            _, polys, _ = seal_format.polys_from_bytes(buf)
            self._pk = self._to_coeff(torch.from_numpy(polys), False)
            return
        self._pk = self._key_from(buf, b"P")

    def from_bytes_secretKey(self, buf: bytes) -> None:
        self._need_ctx()
        if self._is_seal(buf):
            _, poly, _ = seal_format.secret_from_bytes(buf)
            self._sk = self._to_coeff(torch.from_numpy(poly), False)
            return
        self._sk = self._key_from(buf, b"S")

    def ctxt_to_bytes(self, ct: "PyCtxt", format: Optional[str] = None) -> bytes:
        """One fractional ciphertext as a SEAL-2.3 Ciphertext stream (hash block, size, N + 1, modulus count,
        coefficient-form words) or our raw words."""
        if self._fmt(format) == "seal2":
            return seal_format.polys_to_bytes(self._seal_hash(), self._to_coeff(ct._data, True).cpu().numpy())
        return ct.to_bytes()

    def ctxt_from_bytes(self, buf: bytes) -> "PyCtxt":
        if self._is_seal(buf):
            hb, polys, _ = seal_format.polys_from_bytes(buf)
            return PyCtxt(self, self._to_coeff(torch.from_numpy(polys), False), ENC_FRACTIONAL)
        n = self._ctx.n
        arr = np.frombuffer(buf, dtype=np.int64).reshape(2, -1, n).copy()
        return PyCtxt(self, torch.from_numpy(arr).to(self._device), self._scheme)

    # aliases used by other Pyfhel versions
    to_bytes_public_key, to_bytes_secret_key = to_bytes_publicKey, to_bytes_secretKey
    from_bytes_public_key, from_bytes_secret_key = from_bytes_publicKey, from_bytes_secretKey

    # ------------------------------------------------------------------ encrypt / decrypt
    def _next_seed(self) -> int:
        self._seed_counter += 1
        return (self._enc_seed + 0x9E3779B97F4A7C15 * self._seed_counter) & 0x7FFFFFFFFFFFFFFF

    def encryptFrac(self, value) -> PyCtxt:
        """Scalar -> one ciphertext (reference use). Arrays -> one packed CKKS ciphertext batch."""
        self._need_ctx()
        if self._pk is None:
            raise RuntimeError("encryptFrac needs the public key")
        if self._scheme == ENC_CKKS:
            vals = torch.as_tensor(np.atleast_1d(np.asarray(value, dtype=np.float64)), device=self._device)
            ct = self._ctx.encrypt(vals, self._pk, seed=self._next_seed())
            return PyCtxt(self, ct.data, ENC_CKKS, ct.scale, ct.nvals)
        v = torch.tensor([float(value)], dtype=torch.float64)
        data = self._ctx.encrypt(v, self._pk, seed=self._next_seed())[0]
        return PyCtxt(self, data, ENC_FRACTIONAL, 1.0, 1)

    def encryptFracBatch(self, values) -> np.ndarray:
        """Vectorised ``encryptFrac`` over an array: same ciphertexts-per-scalar packing as the
        reference (FLPyfhelin.py:214-219), produced by ONE kernel launch. Returns an object array
        of ``PyCtxt`` shaped like ``values``."""
        self._need_ctx()
        arr = np.asarray(values, dtype=np.float64)
        flat = torch.from_numpy(arr.reshape(-1).copy())
        if self._scheme == ENC_CKKS:
            raise TypeError("use encryptFrac(array) for packed CKKS")
        data = self._ctx.encrypt(flat, self._pk, seed=self._next_seed())
        out = np.empty(flat.numel(), dtype=object)
        for i in range(flat.numel()):
            out[i] = PyCtxt(self, data[i], ENC_FRACTIONAL, 1.0, 1)
        return out.reshape(arr.shape)

    def decryptFrac(self, ctxt: PyCtxt):
        self._need_ctx()
        if self._sk is None:
            raise RuntimeError("decryptFrac needs the secret key")
        if ctxt._encoding == ENC_CKKS:
            out = self._ctx.decrypt(CtBatch(ctxt._data.to(self._device), ctxt._scale, ctxt._nvals), self._sk)
            return out.cpu().numpy().astype(np.float64)
        return float(self._ctx.decrypt(ctxt._data.to(self._device).unsqueeze(0), self._sk)[0])

    def decryptFracBatch(self, ctxts: np.ndarray) -> np.ndarray:
        """Vectorised ``decryptFrac`` over an object array of fractional ciphertexts."""
        flat = ctxts.reshape(-1)
        data = torch.stack([c._data for c in flat]).to(self._device)
        vals = self._ctx.decrypt(data, self._sk)
        return vals.cpu().numpy().reshape(ctxts.shape)

    def noiseLevel(self, ctxt: PyCtxt) -> float:
        """Invariant noise budget in bits (X1.l; the reference only has a commented probe, :382)."""
        if self._scheme != ENC_FRACTIONAL:
            raise NotImplementedError
        return self._ctx.noise_budget_bits(ctxt._data.to(self._device).unsqueeze(0), self._sk)

    # ------------------------------------------------------------------ evaluator (used by PyCtxt)
    def _pair(self, a: PyCtxt, b: PyCtxt):
        if a._encoding != b._encoding:
            raise TypeError("ciphertext encodings differ")
        return a._data.to(self._device), b._data.to(self._device)

    def _add(self, a: PyCtxt, b: PyCtxt) -> torch.Tensor:
        x, y = self._pair(a, b)
        if a._encoding == ENC_CKKS:
            return self._ctx.add(CtBatch(x, a._scale, a._nvals), CtBatch(y, b._scale, b._nvals)).data
        return self._ctx.add(x.unsqueeze(0), y.unsqueeze(0))[0]

    def _sub(self, a: PyCtxt, b: PyCtxt) -> torch.Tensor:
        x, y = self._pair(a, b)
        if a._encoding == ENC_CKKS:
            return self._ctx.sub_(CtBatch(x.clone(), a._scale, a._nvals), CtBatch(y, b._scale, b._nvals)).data
        return self._ctx.sub(x.unsqueeze(0), y.unsqueeze(0))[0]

    def _add_plain(self, a: PyCtxt, v: float) -> torch.Tensor:
        x = a._data.to(self._device)
        if a._encoding == ENC_CKKS:
            vals = torch.full((a._nvals,), v, dtype=torch.float64, device=self._device)
            return self._ctx.add_plain_(CtBatch(x.clone(), a._scale, a._nvals), vals).data
        return self._ctx.add_plain(x.unsqueeze(0), torch.tensor([v], dtype=torch.float64))[0]

    def _mul_plain(self, a: PyCtxt, v: float):
        x = a._data.to(self._device)
        if a._encoding == ENC_CKKS:
            ct = CtBatch(x.clone(), a._scale, a._nvals)
            self._ctx.mul_scalar_(ct, v, rescale=ct.level > 1)
            return ct.data, ct.scale
        return self._ctx.mul_plain(x.unsqueeze(0), v)[0], 1.0

    def _mul_ct(self, a: PyCtxt, b: PyCtxt):
        if self._rlk is None:
            raise RuntimeError("relinKeyGen() must be called before PyCtxt * PyCtxt")
        x, y = self._pair(a, b)
        if a._encoding != ENC_CKKS:
            return self._ctx.multiply(x.unsqueeze(0), y.unsqueeze(0), self._rlk)[0], 1.0
        out = self._ctx.multiply(CtBatch(x, a._scale, a._nvals), CtBatch(y, b._scale, b._nvals), self._rlk)
        return out.data, out.scale

    # ------------------------------------------------------------------ misc
    def _need_ctx(self) -> None:
        if self._ctx is None:
            raise RuntimeError("contextGen() or from_bytes_context() must be called first")

    def __getstate__(self):
        # A pickled Pyfhel carries neither context nor keys: the reference always re-hydrates with
        # from_bytes_* after loading (FLPyfhelin.py:256-259, :351-353). In particular the secret key
        # can never leak through publickey.pickle (quirk Q11).
        return {"_device": str(self._device)}

    def __setstate__(self, st):
        self.__init__(st.get("_device", "cpu"))

    def __repr__(self) -> str:
        pk = "Y" if self._pk is not None else "-"
        sk = "Y" if self._sk is not None else "-"
        rlk = "Y" if self._rlk is not None else "-"
        if self._ctx is None:
            cx = "contx(-)"
        elif self._scheme == ENC_CKKS:
            cx = f"contx(scheme=CKKS, n={self._ctx.n}, qi={[p.bit_length() for p in self._ctx.primes]}, sec={self._ctx.sec})"
        else:
            c = self._ctx
            cx = (f"contx(p={c.p}, m={c.n}, base={c.base}, sec={c.sec}, "
                  f"dig={c.int_digits}i.{c.frac_digits}f, batch=False)")
        return f"<Pyfhel obj at {hex(id(self))}, [pk:{pk}, sk:{sk}, rtk:-, rlk:{rlk}, {cx}]>"

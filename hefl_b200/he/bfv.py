"""BFV with SEAL-2.x-style FractionalEncoder — the scheme the reference actually runs through
Pyfhel 2.3.1 (``contextGen(p=65537, sec=128, m=1024)``, ``encryptFrac``/``decryptFrac``,
FLPyfhelin.py:332, :217, :295; repr ``dig=64i.32f, batch=False`` at notebook N:44).

One ciphertext per scalar (the reference's packing), but stored and processed as batches
``[C, 2, 1, N]`` on the same kernels as the CKKS path (sampler, NTT, pointwise, decrypt):

  encrypt : c0 = pk0*u + e0 + floor(q/p) * m,  c1 = pk1*u + e1     (NTT form)
  add     : coefficient-wise mod q               (PyCtxt + PyCtxt, FLPyfhelin.py:381)
  mul pt  : ct * NTT(lift(m_plain))              (PyCtxt * float, FLPyfhelin.py:385)
  decrypt : round(p/q * INTT(c0 + c1*s)) mod p -> centred digits -> fractional decode

The coefficient modulus is a single NTT prime whose size follows SEAL's 128/192/256-bit
security tables for the ring degree (27 bits at m=1024, 54 at m=2048; 60 beyond).
"""
from __future__ import annotations

import struct
from typing import Optional, Tuple

import torch

from .. import _ext
from ..config import SEC_MAX_LOGQ


import dataclasses


@dataclasses.dataclass
class BFVRelinKey:
    keys: torch.Tensor        # [digits, 2, 1, N], NTT form
    bit_count: int
    size: int


class BFVFracContext:
    def __init__(self, p: int = 65537, m: int = 2048, sec: int = 128, base: int = 2,
                 int_digits: int = 64, frac_digits: int = 32, device: str | torch.device = "cpu",
                 q: Optional[int] = None):
        if base != 2:
            raise NotImplementedError("only base 2 fractional encoding is implemented")
        self.ops = _ext.ops()
        self.p, self.n, self.sec, self.base = int(p), int(m), int(sec), base
        self.logn = self.n.bit_length() - 1
        if 1 << self.logn != self.n:
            raise ValueError("m must be a power of two")
        self.int_digits, self.frac_digits = int_digits, frac_digits
        if int_digits + frac_digits > self.n:
            raise ValueError("intDigits + fracDigits must not exceed m")
        if q is None:
            bits = min(60, SEC_MAX_LOGQ.get(sec, SEC_MAX_LOGQ[128]).get(self.n, 60))
            q = int(self.ops.gen_primes(bits, self.logn, 1, [])[0])
        self.q = int(q)
        self.primes = [self.q]
        self.L = 1
        self.delta = self.q // self.p
        tables, consts = self.ops.build_tables(torch.tensor([self.q], dtype=torch.int64), self.logn)
        self._cpu = dict(tables=tables, consts=consts)
        self.consts_cpu = consts
        self.device = torch.device("cpu")
        self.tables, self.consts = tables, consts
        self.delta_t = torch.tensor([self.delta], dtype=torch.int64)
        self.to(device)

    def to(self, device) -> "BFVFracContext":
        self.device = torch.device(device)
        self.tables = self._cpu["tables"].to(self.device)
        self.consts = self._cpu["consts"].to(self.device)
        self.delta_t = torch.tensor([self.delta], dtype=torch.int64, device=self.device)
        return self

    # ---- keys ---------------------------------------------------------------------------
    def _key_tables(self):
        """Key generation runs where the context lives: CUDA kernels or the host twin (same keys for a seed)."""
        if self.device.type == "cuda":
            return self.tables, self.consts
        return self._cpu["tables"], self._cpu["consts"]

    def keygen(self, seed: int = 0) -> Tuple[torch.Tensor, torch.Tensor]:
        t, c = self._key_tables()
        sk = self.ops.keygen_secret(1, self.logn, t, c, int(seed))
        pk = self.ops.keygen_public(sk, 1, self.logn, t, c, int(seed), 0)
        return sk.to(self.device), pk.to(self.device)

    # ---- codec + encryption ---------------------------------------------------------------
    def encode(self, vals: torch.Tensor) -> torch.Tensor:
        """float64 [C] -> signed digit polynomials int64 [C, N] (SEAL FractionalEncoder, base 2)."""
        v = vals.reshape(-1).to(self.device, torch.float64).contiguous()
        return self.ops.frac_encode(v, self.n, self.int_digits, self.frac_digits)

    def decode(self, coeffs: torch.Tensor) -> torch.Tensor:
        return self.ops.frac_decode(coeffs.contiguous(), self.int_digits, self.frac_digits)

    def encrypt(self, vals: torch.Tensor, pk: torch.Tensor, seed: int, ct_offset: int = 0) -> torch.Tensor:
        """[C] reals -> ciphertext batch int64 [C, 2, 1, N]."""
        msg = self.encode(vals)
        C = msg.shape[0]
        return self.ops.encrypt(msg, pk, C, 1, self.logn, self.tables, self.consts, self.delta_t,
                                int(seed), int(ct_offset))

    def decrypt_digits(self, ct: torch.Tensor, sk: torch.Tensor) -> torch.Tensor:
        res = self.ops.decrypt(ct.contiguous(), sk, 1, self.logn, self.tables, self.consts)   # [C,1,N]
        return self.ops.bfv_scale_round(res.view(-1, self.n), self.q, self.p)

    def decrypt(self, ct: torch.Tensor, sk: torch.Tensor) -> torch.Tensor:
        return self.decode(self.decrypt_digits(ct, sk))

    # ---- homomorphic ops ------------------------------------------------------------------
    def add(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(a)
        self.ops.pointwise_(out, a.contiguous(), b.contiguous(), 1, self.consts, 0)
        return out

    def sub(self, a: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
        out = torch.empty_like(a)
        self.ops.pointwise_(out, a.contiguous(), b.contiguous(), 1, self.consts, 1)
        return out

    def _plain_ntt(self, vals: torch.Tensor, scale_delta: bool) -> torch.Tensor:
        msg = self.encode(vals)                                   # [C,N] signed digits
        pt = torch.remainder(msg, self.q).contiguous()
        if scale_delta:
            self.ops.pointwise_(pt, pt, self.delta_t, 1, self.consts, 5)
        self.ops.ntt_(pt, self.tables, self.consts, 1, self.logn, False)
        return pt

    def add_plain(self, a: torch.Tensor, vals: torch.Tensor) -> torch.Tensor:
        """ct + plaintext (Evaluator::add_plain; ``PyCtxt + 0`` for the first client, FLPyfhelin.py:380)."""
        out = a.clone()
        pt = self._plain_ntt(vals, True)                          # [C or 1, N]
        c0 = out[:, 0, 0].contiguous()
        self.ops.pointwise_(c0, c0, pt, 1, self.consts, 0)
        out[:, 0, 0] = c0
        return out

    def mul_plain(self, a: torch.Tensor, value: float) -> torch.Tensor:
        """ct * encoded(value) (multiply_plain; ``dct * denom``, FLPyfhelin.py:385)."""
        pt = self._plain_ntt(torch.tensor([value], dtype=torch.float64), False)   # [1,N]
        out = torch.empty_like(a)
        self.ops.pointwise_(out, a.contiguous(), pt, 1, self.consts, 2)
        return out

    # ---- relinearisation keys and ciphertext x ciphertext (FLPyfhelin.py:357-364 gen_rekey) -------------
    def relin_keygen(self, sk: torch.Tensor, seed: int, bit_count: int = 16, size: int = 1) -> "BFVRelinKey":
        """SEAL-2.x ``relinKeyGen(bitCount, size)``: evaluation keys for s^2 -> s with decomposition base
        w = 2^bitCount:  evk[k] = (-(a_k s + e_k) + w^k s^2, a_k), k < ceil(log2 q / bitCount), NTT form.
        ``size`` (how many powers of s SEAL 2.x could relinearise; 5 in the reference) is recorded: products here
        are relinearised immediately, so only the s^2 key is ever needed."""
        bit_count = max(1, min(int(bit_count), 60))
        t, c = self._key_tables()
        sk_d = sk.to(t.device).contiguous()
        s2 = torch.empty_like(sk_d)
        self.ops.pointwise_(s2, sk_d, sk_d, 1, c, 2)
        nd = (self.q.bit_length() + bit_count - 1) // bit_count
        w = [pow(2, k * bit_count, self.q) for k in range(nd)]
        evk = self.ops.keygen_public_batch(sk_d, 1, self.logn, t, c, int(seed), 1, nd)          # [nd, 2, 1, N]
        self.ops.relin_message_(evk, s2, torch.zeros(nd, dtype=torch.int32, device=t.device),
                                torch.tensor(w, dtype=torch.int64, device=t.device), 1, c)
        return BFVRelinKey(evk.to(self.device).contiguous(), bit_count, int(size))

    def _ext_basis(self):
        """Auxiliary RNS basis {q, p1, p2, p3}: wide enough (q * 2^174) to hold the integer tensor product of two
        ciphertexts, whose coefficients reach N * (q/2)^2, exactly."""
        if getattr(self, "_ext", None) is None:
            aux = []
            for _ in range(3):
                aux.append(int(self.ops.gen_primes(58, self.logn, 1, [self.q] + aux)[0]))
            primes = [self.q] + aux
            tables, consts = self.ops.build_tables(torch.tensor(primes, dtype=torch.int64), self.logn)
            self._ext = dict(primes=primes, tables=tables, consts=consts, dev={})
        e = self._ext
        key = str(self.device)
        if key not in e["dev"]:
            e["dev"][key] = (e["tables"].to(self.device), e["consts"].to(self.device))
        return e["primes"], e["dev"][key][0], e["dev"][key][1]

    def multiply(self, a: torch.Tensor, b: torch.Tensor, rlk: "BFVRelinKey") -> torch.Tensor:
        """BFV ciphertext product with scale-and-round and relinearisation (``PyCtxt * PyCtxt``):

        1. both ciphertexts to coefficient form, centred, and into the wide basis {q, p1, p2, p3};
        2. tensor product (d0, d1, d2) under every prime of that basis (NTT, point-wise, INTT) -- exact over Z;
        3. CRT-compose each coefficient, y = round(p * x / q) mod q  (host big integers: this is the API-parity
           path of dead code in the reference, not a throughput path);
        4. relinearise d2 with the base-2^bitCount evaluation keys (fused key-switch kernel on CUDA)."""
        import numpy as np

        primes, xt, xc = self._ext_basis()
        K = len(primes)
        C = a.shape[0]
        dev = self.device

        def lift(ct):
            x = ct.reshape(C * 2, self.n).clone().contiguous()
            self.ops.ntt_(x, self.tables, self.consts, 1, self.logn, True)                 # coefficients in [0, q)
            signed = torch.where(x > self.q // 2, x - self.q, x)                           # centred
            limbs = torch.stack([torch.remainder(signed, p) for p in primes], dim=1).contiguous()   # [2C, K, N]
            self.ops.ntt_(limbs, xt, xc, K, self.logn, False)
            return limbs.view(C, 2, K, self.n)

        A, B = lift(a.to(dev)), lift(b.to(dev))
        prods = []
        for (i, j, acc) in ((0, 0, None), (0, 1, None), (1, 0, 1), (1, 1, None)):
            t = torch.empty(C, K, self.n, dtype=torch.int64, device=dev)
            if acc is None:
                self.ops.pointwise_(t, A[:, i].contiguous(), B[:, j].contiguous(), K, xc, 2)
                prods.append(t)
            else:
                self.ops.pointwise_(prods[acc], A[:, i].contiguous(), B[:, j].contiguous(), K, xc, 3)   # d1 += a1*b0
        d = torch.stack(prods, dim=1).contiguous()                                        # [C, 3, K, N]
        self.ops.ntt_(d, xt, xc, K, self.logn, True)
        # exact CRT -> scale-and-round, coefficient by coefficient (Python integers)
        Q = 1
        for p in primes:
            Q *= p
        res = d.cpu().numpy().astype(object)                                              # [C,3,K,N]
        x = np.zeros(res.shape[:2] + (self.n,), dtype=object)
        for k, p in enumerate(primes):
            Qk = Q // p
            x = x + res[:, :, k, :] * ((Qk * pow(Qk, -1, p)) % Q)
        x = x % Q
        x = np.where(x > Q // 2, x - Q, x)
        y = (2 * self.p * x + self.q) // (2 * self.q)                                     # round(p x / q), exact
        y = (y % self.q).astype(np.int64)
        dq = torch.from_numpy(y).to(dev)                                                  # [C,3,N] coefficient form mod q
        out = torch.empty(C, 2, 1, self.n, dtype=torch.int64, device=dev)
        d01 = dq[:, :2].reshape(C * 2, self.n).clone().contiguous()
        self.ops.ntt_(d01, self.tables, self.consts, 1, self.logn, False)
        out[:, :, 0] = d01.view(C, 2, self.n)
        coef2 = dq[:, 2].reshape(C, 1, self.n).contiguous()
        evk = rlk.keys.to(dev)
        nd = evk.shape[0]
        done = False
        if out.is_cuda:
            done = self.ops.keyswitch_fused_(out, coef2, evk, [nd], [0], rlk.bit_count, self.tables, self.consts, self.logn)
        if not done:
            r0 = out[:, 0, 0].contiguous()
            r1 = out[:, 1, 0].contiguous()
            src = coef2.view(C, self.n)
            for k in range(nd):
                dig = self.ops.digit_extract(src, k * rlk.bit_count, rlk.bit_count)
                self.ops.ntt_(dig, self.tables, self.consts, 1, self.logn, False)
                self.ops.pointwise_(r0, dig, evk[k, 0, 0].contiguous(), 1, self.consts, 3)
                self.ops.pointwise_(r1, dig, evk[k, 1, 0].contiguous(), 1, self.consts, 3)
            out[:, 0, 0], out[:, 1, 0] = r0, r1
        return out

    def noise_budget_bits(self, ct: torch.Tensor, sk: torch.Tensor) -> float:
        """Invariant noise budget of the worst ciphertext in the batch."""
        import math

        res = self.ops.decrypt(ct.contiguous(), sk, 1, self.logn, self.tables, self.consts).view(-1, self.n)
        digits = self.ops.bfv_scale_round(res, self.q, self.p)
        err = torch.remainder(res - digits * self.delta, self.q)
        err = torch.minimum(err, self.q - err)
        worst = int(err.max().item())
        return math.log2(self.delta / 2) - math.log2(max(worst, 1))

    # ---- serialization ---------------------------------------------------------------------
    def to_bytes_context(self) -> bytes:
        return b"HEFB" + struct.pack("<HIIIIIIQ", 1, self.p, self.n, self.sec, self.base, self.int_digits,
                                     self.frac_digits, self.q)

    @classmethod
    def from_bytes_context(cls, buf: bytes, device="cpu") -> "BFVFracContext":
        if buf[:4] != b"HEFB":
            raise ValueError("not a BFV context stream")
        ver, p, n, sec, base, idg, fdg, q = struct.unpack_from("<HIIIIIIQ", buf, 4)
        return cls(p=p, m=n, sec=sec, base=base, int_digits=idg, frac_digits=fdg, device=device, q=q)

    def __repr__(self) -> str:
        return (f"<BFVFracContext p={self.p} m={self.n} q~2^{self.q.bit_length()} sec={self.sec} "
                f"dig={self.int_digits}i.{self.frac_digits}f device={self.device}>")

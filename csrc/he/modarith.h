// Modular arithmetic primitives shared by the host (g++) and device (nvcc) builds.
//
// Words are u64, primes are < 2^61 so that (a) Harvey lazy butterflies can keep
// values in [0, 4q) without wrapping and (b) the sum of 8 residues cannot wrap a
// u64 inside the fused all-reduce.
//
// Replaces the arithmetic the reference reaches through Pyfhel -> SEAL
// (FLPyfhelin.py:217, :295, :381, :385).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HEFL_HD __host__ __device__ __forceinline__
#else
#define HEFL_HD inline
#endif

namespace hefl {

HEFL_HD uint64_t mul_hi(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  return __umul64hi(a, b);
#else
  return (uint64_t)(((unsigned __int128)a * b) >> 64);
#endif
}

HEFL_HD void mul_wide(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
  lo = a * b;
  hi = __umul64hi(a, b);
#else
  unsigned __int128 p = (unsigned __int128)a * b;
  lo = (uint64_t)p;
  hi = (uint64_t)(p >> 64);
#endif
}

// Per-prime constants. ratio = floor(2^128 / q) as two words (lo, hi).
struct Modulus {
  uint64_t q;
  uint64_t ratio_lo;
  uint64_t ratio_hi;
};

HEFL_HD uint64_t add_mod(uint64_t a, uint64_t b, uint64_t q) {
  uint64_t s = a + b;
  return s >= q ? s - q : s;
}

HEFL_HD uint64_t sub_mod(uint64_t a, uint64_t b, uint64_t q) {
  return a >= b ? a - b : a + q - b;
}

HEFL_HD uint64_t neg_mod(uint64_t a, uint64_t q) { return a == 0 ? 0 : q - a; }

// x (any u64) mod q using the high ratio word; at most two corrections.
HEFL_HD uint64_t barrett_reduce_64(uint64_t x, const Modulus& m) {
  uint64_t qhat = mul_hi(x, m.ratio_hi);
  uint64_t r = x - qhat * m.q;
  if (r >= m.q) r -= m.q;
  if (r >= m.q) r -= m.q;
  return r;
}

// (hi:lo) mod q, valid for any 128-bit input when q < 2^63.
HEFL_HD uint64_t barrett_reduce_128(uint64_t hi, uint64_t lo, const Modulus& m) {
  // qhat = floor((hi:lo) * ratio / 2^128), low 64 bits only.
  uint64_t carry = mul_hi(lo, m.ratio_lo);
  uint64_t t2hi, t2lo;
  mul_wide(lo, m.ratio_hi, t2hi, t2lo);
  uint64_t t1 = t2lo + carry;
  uint64_t t3 = t2hi + (t1 < t2lo ? 1 : 0);
  mul_wide(hi, m.ratio_lo, t2hi, t2lo);
  uint64_t t1b = t1 + t2lo;
  carry = t2hi + (t1b < t1 ? 1 : 0);
  uint64_t qhat = hi * m.ratio_hi + t3 + carry;
  uint64_t r = lo - qhat * m.q;
  if (r >= m.q) r -= m.q;
  return r;
}

HEFL_HD uint64_t mul_mod(uint64_t a, uint64_t b, const Modulus& m) {
  uint64_t hi, lo;
  mul_wide(a, b, hi, lo);
  return barrett_reduce_128(hi, lo, m);
}

// a*b + c mod q (c < q).
HEFL_HD uint64_t mad_mod(uint64_t a, uint64_t b, uint64_t c, const Modulus& m) {
  uint64_t hi, lo;
  mul_wide(a, b, hi, lo);
  uint64_t lo2 = lo + c;
  hi += (lo2 < lo) ? 1 : 0;
  return barrett_reduce_128(hi, lo2, m);
}

// Shoup multiplication by a constant w with companion w' = floor(w * 2^64 / q).
// Lazy form: result in [0, 2q) for any x < 2^64.
HEFL_HD uint64_t mul_shoup_lazy(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
  uint64_t qhat = mul_hi(x, wp);
  return x * w - qhat * q;
}

HEFL_HD uint64_t mul_shoup(uint64_t x, uint64_t w, uint64_t wp, uint64_t q) {
  uint64_t r = mul_shoup_lazy(x, w, wp, q);
  return r >= q ? r - q : r;
}

// Signed small integer -> residue.
HEFL_HD uint64_t lift_signed(int64_t v, uint64_t q) {
  return v >= 0 ? (uint64_t)v : q - (uint64_t)(-v);
}

// Signed 64-bit value of arbitrary magnitude -> residue.
HEFL_HD uint64_t reduce_signed(int64_t v, const Modulus& m) {
  if (v >= 0) return barrett_reduce_64((uint64_t)v, m);
  uint64_t r = barrett_reduce_64((uint64_t)(-(v + 1)) + 1ull, m);
  return r == 0 ? 0 : m.q - r;
}

HEFL_HD uint32_t bit_reverse(uint32_t x, int bits) {
#if defined(__CUDA_ARCH__)
  return __brev(x) >> (32 - bits);
#else
  uint32_t r = 0;
  for (int i = 0; i < bits; ++i) { r = (r << 1) | (x & 1); x >>= 1; }
  return r;
#endif
}

}  // namespace hefl

// Counter-based Philox4x32-10 generator and the samplers built on it.
//
// Counter-based so that the CPU oracle and the CUDA kernels draw the *same*
// randomness for (seed, stream, ciphertext, coefficient): encrypt on the GPU is
// then bit-exact against the host implementation, and the same
// (u, e0, e1) are produced for every RNS limb without storing them.
//
// Replaces SEAL's sampler reached from Pyfhel keyGen / encryptFrac
// (FLPyfhelin.py:333, :217).
#pragma once
#include <cstdint>
#include "modarith.h"

namespace hefl {

struct Philox4 {
  uint32_t x, y, z, w;
};

HEFL_HD uint32_t mulhi32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}

HEFL_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                              uint32_t k0, uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = mulhi32(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = mulhi32(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return Philox4{c0, c1, c2, c3};
}

HEFL_HD int popc32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __popc(v);
#else
  return __builtin_popcount(v);
#endif
}

// Streams (c3 of the counter).
enum : uint32_t {
  STREAM_ENC_A = 0,   // u, e0, e1 (one call per coefficient)
  STREAM_ENC_B = 1,   // (unused since the single-call sampler; kept so stream numbers stay stable)
  STREAM_SK = 2,      // secret key
  STREAM_PK_E = 3,    // public-key error
  STREAM_UNIFORM = 4  // uniform mod q (c2 carries the limb)
};

// Uniform ternary in {-1, 0, 1}.
HEFL_HD int ternary_from(uint32_t r) { return (int)mulhi32(r, 3u) - 1; }

// Centered binomial with k = 21 (variance 10.5, sigma = 3.24).
HEFL_HD int cbd21_from(uint32_t a, uint32_t b) {
  return popc32(a & 0x1FFFFFu) - popc32(b & 0x1FFFFFu);
}

struct EncNoise {
  int u, e0, e1;
};

// All encryption randomness of one coefficient from ONE Philox call (128 bits): u from word x,
// e0 from the low 21 bits of y and z, e1 from the low 21 bits of w against the 21 bits made of
// w[21..31] and y[21..31] -- 32 + 42 + 42 disjoint bits. Philox is 65 instructions per call; with
// one call the sampler costs ~1/7 of the three NTTs it feeds instead of ~1/2.
HEFL_HD EncNoise enc_noise_from(const Philox4& a) {
  EncNoise n;
  n.u = ternary_from(a.x);
  n.e0 = cbd21_from(a.y, a.z);
  n.e1 = cbd21_from(a.w, (a.w >> 21) | ((a.y >> 21) << 11));
  return n;
}

// Encryption randomness for coefficient i of ciphertext ct.
HEFL_HD EncNoise sample_enc_noise(uint64_t seed, uint32_t ct, uint32_t i) {
  return enc_noise_from(philox4x32_10(i, ct, 0u, STREAM_ENC_A, (uint32_t)seed, (uint32_t)(seed >> 32)));
}

HEFL_HD int sample_ternary(uint64_t seed, uint32_t stream, uint32_t i) {
  Philox4 a = philox4x32_10(i, 0u, 0u, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  return ternary_from(a.x);
}

HEFL_HD int sample_cbd(uint64_t seed, uint32_t stream, uint32_t idx, uint32_t i) {
  Philox4 a = philox4x32_10(i, idx, 0u, stream, (uint32_t)seed, (uint32_t)(seed >> 32));
  return cbd21_from(a.x, a.y);
}

// Uniform residue mod q for (index idx, limb, coefficient i). 128 random bits
// reduced mod q: bias < 2^-64.
HEFL_HD uint64_t sample_uniform(uint64_t seed, uint32_t idx, uint32_t limb, uint32_t i,
                                const Modulus& m) {
  Philox4 a = philox4x32_10(i, idx, limb, STREAM_UNIFORM, (uint32_t)seed, (uint32_t)(seed >> 32));
  uint64_t hi = ((uint64_t)a.x << 32) | a.y;
  uint64_t lo = ((uint64_t)a.z << 32) | a.w;
  return barrett_reduce_128(hi, lo, m);
}

}  // namespace hefl

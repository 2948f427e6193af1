// CPU verifier (host C++, no CUDA): BN254 Fr / Fq in 4x64-bit-limb Montgomery form (R = 2^256), CIOS multiplication,
// add / sub with conditional subtraction, `to_repr` = canonical little-endian 32 bytes — the layout the reference's
// field type has in memory (halo2curves bn256, pb/util/arithmetic.rs:15-22), so proofs, points and scalars cross the
// C ABI unconverted. The verifier side of the reference runs on the CPU too.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>

namespace b200v {

typedef unsigned __int128 u128;

struct FrParams {
  static constexpr uint64_t MOD[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL,
                                      0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr uint64_t R[4] = {0xac96341c4ffffffbULL, 0x36fc76959f60cd29ULL,
                                    0x666ea36f7879462eULL, 0x0e0a77c19a07df2fULL};
  static constexpr uint64_t R2[4] = {0x1bb8e645ae216da7ULL, 0x53fe3ab1e35c59e3ULL,
                                     0x8c49833d53bb8085ULL, 0x0216d0b17f4e44a5ULL};
  static constexpr uint64_t INV = 0xc2e1f593efffffffULL;  // -r^{-1} mod 2^64
};

struct FqParams {
  static constexpr uint64_t MOD[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL,
                                      0xb85045b68181585dULL, 0x30644e72e131a029ULL};
  static constexpr uint64_t R[4] = {0xd35d438dc58f0d9dULL, 0x0a78eb28f5c70b3dULL,
                                    0x666ea36f7879462cULL, 0x0e0a77c19a07df2fULL};
  static constexpr uint64_t R2[4] = {0xf32cfc5b538afa89ULL, 0xb5e71911d44501fbULL,
                                     0x47ab1eff0a417ff6ULL, 0x06d89f71cab8351fULL};
  static constexpr uint64_t INV = 0x87d20782e4866389ULL;  // -q^{-1} mod 2^64
};

// Element of Z_p in Montgomery form, little-endian u64 limbs: the in-memory layout of
// halo2curves' `Fr([u64; 4])` / `Fq([u64; 4])`, so buffers cross the C ABI unconverted.
template <class P>
struct Fp {
  uint64_t l[4];

  static Fp zero() { return Fp{{0, 0, 0, 0}}; }
  static Fp one() { return Fp{{P::R[0], P::R[1], P::R[2], P::R[3]}}; }

  bool is_zero() const { return (l[0] | l[1] | l[2] | l[3]) == 0; }
  bool operator==(const Fp& o) const {
    return l[0] == o.l[0] && l[1] == o.l[1] && l[2] == o.l[2] && l[3] == o.l[3];
  }
  bool operator!=(const Fp& o) const { return !(*this == o); }

  static inline bool geq_mod(const uint64_t a[4]) {
    for (int i = 3; i >= 0; --i) {
      if (a[i] > P::MOD[i]) return true;
      if (a[i] < P::MOD[i]) return false;
    }
    return true;
  }
  static inline void sub_mod(uint64_t a[4]) {
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
      u128 d = (u128)a[i] - P::MOD[i] - borrow;
      a[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
  }

  Fp operator+(const Fp& o) const {
    Fp r;
    u128 c = 0;
    for (int i = 0; i < 4; ++i) {
      c += (u128)l[i] + o.l[i];
      r.l[i] = (uint64_t)c;
      c >>= 64;
    }
    // both moduli are < 2^254, so the sum of two reduced elements never carries out of 256 bits
    if (geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fp operator-(const Fp& o) const {
    Fp r;
    u128 borrow = 0;
    for (int i = 0; i < 4; ++i) {
      u128 d = (u128)l[i] - o.l[i] - borrow;
      r.l[i] = (uint64_t)d;
      borrow = (d >> 64) & 1;
    }
    if (borrow) {
      u128 c = 0;
      for (int i = 0; i < 4; ++i) {
        c += (u128)r.l[i] + P::MOD[i];
        r.l[i] = (uint64_t)c;
        c >>= 64;
      }
    }
    return r;
  }
  Fp operator-() const { return is_zero() ? *this : zero() - *this; }
  Fp dbl() const { return *this + *this; }

  // CIOS Montgomery product a*b*R^{-1} mod p. Valid for any a < 2^256 when b < p.
  static inline Fp mont_mul(const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; ++i) {
      u128 c = 0;
      for (int j = 0; j < 4; ++j) {
        c += (u128)a[j] * b[i] + t[j];
        t[j] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[4] = (uint64_t)c;
      t[5] = (uint64_t)(c >> 64);
      uint64_t m = t[0] * P::INV;
      c = (u128)m * P::MOD[0] + t[0];
      c >>= 64;
      for (int j = 1; j < 4; ++j) {
        c += (u128)m * P::MOD[j] + t[j];
        t[j - 1] = (uint64_t)c;
        c >>= 64;
      }
      c += t[4];
      t[3] = (uint64_t)c;
      t[4] = t[5] + (uint64_t)(c >> 64);
    }
    Fp r{{t[0], t[1], t[2], t[3]}};
    if (t[4] || geq_mod(r.l)) sub_mod(r.l);
    return r;
  }
  Fp operator*(const Fp& o) const { return mont_mul(l, o.l); }
  Fp sqr() const { return mont_mul(l, l); }
  Fp& operator+=(const Fp& o) { return *this = *this + o; }
  Fp& operator-=(const Fp& o) { return *this = *this - o; }
  Fp& operator*=(const Fp& o) { return *this = *this * o; }

  // canonical integer (any value < 2^256, reduced mod p) -> Montgomery form
  static Fp from_raw(const uint64_t v[4]) { return mont_mul(v, P::R2); }
  static Fp from_u64(uint64_t v) {
    uint64_t t[4] = {v, 0, 0, 0};
    return from_raw(t);
  }
  // Montgomery form -> canonical integer limbs
  void to_raw(uint64_t out[4]) const {
    const uint64_t one[4] = {1, 0, 0, 0};
    Fp r = mont_mul(l, one);
    memcpy(out, r.l, 32);
  }
  // `PrimeField::to_repr`: canonical little-endian bytes
  void to_repr(uint8_t out[32]) const {
    uint64_t raw[4];
    to_raw(raw);
    memcpy(out, raw, 32);  // host is little-endian
  }
  // little-endian 32 bytes, reduced mod p (fe_mod_from_le_bytes, pb/util/arithmetic.rs:150-152)
  static Fp from_le_bytes_mod(const uint8_t in[32]) {
    uint64_t raw[4];
    memcpy(raw, in, 32);
    return from_raw(raw);
  }

  Fp pow(const uint64_t e[4]) const {
    Fp acc = one();
    for (int i = 255; i >= 0; --i) {
      acc = acc.sqr();
      if ((e[i / 64] >> (i % 64)) & 1) acc = acc * *this;
    }
    return acc;
  }
  // Fermat inverse; inverse of zero is zero (callers that follow `BatchInvert` skip zeros).
  Fp inv() const {
    uint64_t e[4] = {P::MOD[0] - 2, P::MOD[1], P::MOD[2], P::MOD[3]};
    return pow(e);
  }
};

typedef Fp<FrParams> Fr;
typedef Fp<FqParams> Fq;

// `ff::BatchInvert` semantics (used at pb/util/arithmetic.rs:121,128): in-place, zeros skipped.
template <class F>
inline void batch_invert(F* v, size_t n) {
  if (n == 0) return;
  F* prefix = new F[n];
  F acc = F::one();
  for (size_t i = 0; i < n; ++i) {
    prefix[i] = acc;
    if (!v[i].is_zero()) acc = acc * v[i];
  }
  acc = acc.inv();
  for (size_t i = n; i-- > 0;) {
    if (v[i].is_zero()) continue;
    F t = v[i];
    v[i] = acc * prefix[i];
    acc = acc * t;
  }
  delete[] prefix;
}

}  // namespace b200v

// BN254 Fr / Fq arithmetic for sm_100a: 8 x 32-bit limbs, Montgomery form (R = 2^256), written as
// PTX mad.lo.cc / madc.hi.cc carry chains on the IMAD pipe. Replaces what the reference gets from
// halo2_curves::bn256::{Fr,Fq} (pb/util/arithmetic.rs:15-22); the in-memory layout (little-endian
// limbs of the Montgomery residue) is identical, so Rust `Vec<Fr>` crosses the C ABI unconverted.
//
// Multiplication is an operand-scanning Montgomery product that keeps the running value as
// T = even + 2^32 * odd: products a[j]*b_i for even j land on limb pairs (j, j+1) that do not
// overlap, so the four of them form ONE carry chain of pure mad instructions (no separate adds);
// odd j go to a second accumulator shifted by one limb. Dividing by 2^32 swaps the roles of the two
// accumulators, which costs nothing after unrolling.
//
// Every primitive is also implemented for the host with an emulated carry flag, so the exact limb
// algorithm is unit-tested on the CPU (tests/test_ff32_host.py) before it ever runs on a GPU.
//
// Attribution: the even / odd accumulator scheme and the helper decomposition used below (ff_mul_n,
// ff_cmad_n, ff_madc_n_rshift, ff_mad_n_redc, fe_final_sub) follow the publicly documented design of
// Supranational's sppark library (`ff/mont_t.cuh`, Apache-2.0): the technique and the helper names
// are theirs; this file is an independent implementation written for this project (fixed 8-limb
// BN254 moduli, host emulation of the carry flag, Kaliski inversion), no sppark source is included.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FF_HD __host__ __device__ __forceinline__
#define FF_D __device__ __forceinline__
#else
#define FF_HD inline
#define FF_D inline
#endif

namespace b200 {

// ---------------------------------------------------------------------------------------------
// carry-chain primitives
// ---------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
FF_HD void add_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void addc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void subc(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void mul_lo(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void mul_hi(uint32_t& d, uint32_t a, uint32_t b) { asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); }
FF_HD void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
FF_HD void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
FF_HD void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
FF_HD void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); }
#else
// host emulation of the PTX condition-code register
static thread_local uint32_t ff_cf = 0;
inline void add_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 32); }
inline void addc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + ff_cf; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 32); }
inline void addc(uint32_t& d, uint32_t a, uint32_t b) { d = a + b + ff_cf; }
inline void sub_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 63); }
inline void subc_cc(uint32_t& d, uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - ff_cf; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 63); }
inline void subc(uint32_t& d, uint32_t a, uint32_t b) { d = a - b - ff_cf; }
inline void mul_lo(uint32_t& d, uint32_t a, uint32_t b) { d = a * b; }
inline void mul_hi(uint32_t& d, uint32_t a, uint32_t b) { d = (uint32_t)(((uint64_t)a * b) >> 32); }
inline void mad_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 32); }
inline void madc_lo_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (uint64_t)(uint32_t)(a * b) + c + ff_cf; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 32); }
inline void madc_hi_cc(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { uint64_t t = (((uint64_t)a * b) >> 32) + c + ff_cf; d = (uint32_t)t; ff_cf = (uint32_t)(t >> 32); }
inline void madc_hi(uint32_t& d, uint32_t a, uint32_t b, uint32_t c) { d = (uint32_t)(((uint64_t)a * b) >> 32) + c + ff_cf; }
#endif

// ---------------------------------------------------------------------------------------------
// field parameters (32-bit limbs, little-endian)
// ---------------------------------------------------------------------------------------------
struct FrP {
  static constexpr uint32_t M0 = 0xefffffffu;  // -r^{-1} mod 2^32
  FF_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t M[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return M[i];
  }
  FF_HD static constexpr uint32_t one(int i) {  // R mod r
    constexpr uint32_t M[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                               0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return M[i];
  }
  FF_HD static constexpr uint32_t r2(int i) {  // R^2 mod r
    constexpr uint32_t M[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                               0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
    return M[i];
  }
};
struct FqP {
  static constexpr uint32_t M0 = 0xe4866389u;  // -q^{-1} mod 2^32
  FF_HD static constexpr uint32_t mod(int i) {
    constexpr uint32_t M[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                               0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
    return M[i];
  }
  FF_HD static constexpr uint32_t one(int i) {
    constexpr uint32_t M[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                               0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
    return M[i];
  }
  FF_HD static constexpr uint32_t r2(int i) {
    constexpr uint32_t M[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                               0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
    return M[i];
  }
};

template <class P>
struct alignas(32) Fe {
  uint32_t v[8];
};
typedef Fe<FrP> Fr;
typedef Fe<FqP> Fq;

template <class P>
FF_HD Fe<P> fe_zero() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = 0;
  return r;
}
template <class P>
FF_HD Fe<P> fe_one() {
  Fe<P> r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = P::one(i);
  return r;
}
template <class P>
FF_HD bool fe_is_zero(const Fe<P>& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t |= a.v[i];
  return t == 0;
}
template <class P>
FF_HD bool fe_eq(const Fe<P>& a, const Fe<P>& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) t |= a.v[i] ^ b.v[i];
  return t == 0;
}

// r = a - p if a >= p else a   (a < 2p)
template <class P>
FF_HD void fe_final_sub(uint32_t a[8]) {
  uint32_t t[8], borrow;
  sub_cc(t[0], a[0], P::mod(0));
#pragma unroll
  for (int i = 1; i < 8; ++i) subc_cc(t[i], a[i], P::mod(i));
  subc(borrow, 0, 0);  // 0xffffffff when a < p
#pragma unroll
  for (int i = 0; i < 8; ++i) a[i] = borrow ? a[i] : t[i];
}

template <class P>
FF_HD Fe<P> fe_add(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  add_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 7; ++i) addc_cc(r.v[i], a.v[i], b.v[i]);
  addc(r.v[7], a.v[7], b.v[7]);  // both moduli < 2^254: no carry out of 256 bits
  fe_final_sub<P>(r.v);
  return r;
}

template <class P>
FF_HD Fe<P> fe_sub(const Fe<P>& a, const Fe<P>& b) {
  Fe<P> r;
  uint32_t borrow;
  sub_cc(r.v[0], a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; ++i) subc_cc(r.v[i], a.v[i], b.v[i]);
  subc(borrow, 0, 0);  // all-ones when a < b
  // add back (p & borrow)
  add_cc(r.v[0], r.v[0], P::mod(0) & borrow);
#pragma unroll
  for (int i = 1; i < 7; ++i) addc_cc(r.v[i], r.v[i], P::mod(i) & borrow);
  addc(r.v[7], r.v[7], P::mod(7) & borrow);
  return r;
}

template <class P>
FF_HD Fe<P> fe_neg(const Fe<P>& a) {
  return fe_sub<P>(fe_zero<P>(), a);
}
template <class P>
FF_HD Fe<P> fe_dbl(const Fe<P>& a) {
  return fe_add<P>(a, a);
}

// ---- Montgomery multiplication ---------------------------------------------------------------
// acc[0..7] = Σ_{j=0,2,4,6} a[j] * bi * 2^{32 j}     (a may point at a+1 for the odd limbs)
FF_HD void ff_mul_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) {
    mul_lo(acc[j], a[j], bi);
    mul_hi(acc[j + 1], a[j], bi);
  }
}
// acc += Σ_{j=0,2,4,6} a[j] * bi * 2^{32 j}; leaves the carry-out in CF
FF_HD void ff_cmad_n(uint32_t* acc, const uint32_t* a, uint32_t bi) {
  mad_lo_cc(acc[0], a[0], bi, acc[0]);
  madc_hi_cc(acc[1], a[0], bi, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    madc_lo_cc(acc[j], a[j], bi, acc[j]);
    madc_hi_cc(acc[j + 1], a[j], bi, acc[j + 1]);
  }
}
// same with the modulus limbs as compile-time constants: acc += Σ_j p[j+OFF] * mi * 2^{32 j}
template <class P, int OFF>
FF_HD void ff_cmad_mod(uint32_t* acc, uint32_t mi) {
  mad_lo_cc(acc[0], P::mod(OFF), mi, acc[0]);
  madc_hi_cc(acc[1], P::mod(OFF), mi, acc[1]);
#pragma unroll
  for (int j = 2; j < 8; j += 2) {
    madc_lo_cc(acc[j], P::mod(j + OFF), mi, acc[j]);
    madc_hi_cc(acc[j + 1], P::mod(j + OFF), mi, acc[j + 1]);
  }
}
// odd'[j] = odd[j+2] + (a[j]*bi) pairs, continuing the carry in CF; top pair starts from zero
FF_HD void ff_madc_n_rshift(uint32_t* odd, const uint32_t* a, uint32_t bi) {
#pragma unroll
  for (int j = 0; j < 6; j += 2) {
    madc_lo_cc(odd[j], a[j], bi, odd[j + 2]);
    madc_hi_cc(odd[j + 1], a[j], bi, odd[j + 3]);
  }
  madc_lo_cc(odd[6], a[6], bi, 0);
  madc_hi(odd[7], a[6], bi, 0);
}
// one operand-scanning step: T = (T + a*bi + m*p) / 2^32 with T = even + 2^32*odd (see header)
template <class P, bool FIRST>
FF_HD void ff_mad_n_redc(uint32_t* even, uint32_t* odd, const uint32_t* a, uint32_t bi) {
  if (FIRST) {
    ff_mul_n(odd, a + 1, bi);
    ff_mul_n(even, a, bi);
  } else {
    add_cc(even[0], even[0], odd[1]);
    ff_madc_n_rshift(odd, a + 1, bi);
    ff_cmad_n(even, a, bi);
    addc(odd[7], odd[7], 0);
  }
  uint32_t mi;
  mul_lo(mi, even[0], P::M0);
  ff_cmad_mod<P, 1>(odd, mi);
  ff_cmad_mod<P, 0>(even, mi);
  addc(odd[7], odd[7], 0);
}

template <class P>
FF_HD Fe<P> fe_mul(const Fe<P>& a, const Fe<P>& b) {
  uint32_t even[8], odd[8];
  ff_mad_n_redc<P, true>(even, odd, a.v, b.v[0]);
  ff_mad_n_redc<P, false>(odd, even, a.v, b.v[1]);
#pragma unroll
  for (int i = 2; i < 8; i += 2) {
    ff_mad_n_redc<P, false>(even, odd, a.v, b.v[i]);
    ff_mad_n_redc<P, false>(odd, even, a.v, b.v[i + 1]);
  }
  // result = even + (odd >> 32)
  Fe<P> r;
  add_cc(r.v[0], even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < 7; ++i) addc_cc(r.v[i], even[i], odd[i + 1]);
  addc(r.v[7], even[7], 0);
  fe_final_sub<P>(r.v);
  return r;
}
template <class P>
FF_HD Fe<P> fe_sqr(const Fe<P>& a) {
  return fe_mul<P>(a, a);
}

// any 256-bit integer -> Montgomery residue of (a mod p). The running value of fe_mul(x, y) stays
// below x + p, so the possibly non-reduced operand must be the SCANNED one (second argument).
template <class P>
FF_HD Fe<P> fe_from_canonical(const Fe<P>& a) {
  Fe<P> r2;
#pragma unroll
  for (int i = 0; i < 8; ++i) r2.v[i] = P::r2(i);
  return fe_mul<P>(r2, a);
}
template <class P>
FF_HD Fe<P> fe_to_canonical(const Fe<P>& a) {
  Fe<P> one = fe_zero<P>();
  one.v[0] = 1;
  return fe_mul<P>(a, one);
}
template <class P>
FF_HD Fe<P> fe_from_u64(uint64_t x) {
  Fe<P> t = fe_zero<P>();
  t.v[0] = (uint32_t)x;
  t.v[1] = (uint32_t)(x >> 32);
  return fe_from_canonical<P>(t);
}

// Fermat inverse a^(p-2); inv(0) = 0 (kept as the reference for the unit test of fe_inv).
template <class P>
FF_HD Fe<P> fe_inv_fermat(const Fe<P>& a) {
  Fe<P> acc = fe_one<P>();
  for (int i = 255; i >= 0; --i) {
    // exponent p-2: p[0] >= 2 in both fields, so only limb 0 differs from p
    uint32_t limb;
    switch (i >> 5) {
      case 0: limb = P::mod(0) - 2; break;
      case 1: limb = P::mod(1); break;
      case 2: limb = P::mod(2); break;
      case 3: limb = P::mod(3); break;
      case 4: limb = P::mod(4); break;
      case 5: limb = P::mod(5); break;
      case 6: limb = P::mod(6); break;
      default: limb = P::mod(7); break;
    }
    acc = fe_sqr<P>(acc);
    if ((limb >> (i & 31)) & 1) acc = fe_mul<P>(acc, a);
  }
  return acc;
}

// Montgomery inverse by Kaliski's binary algorithm: inv(0) = 0. ~1.4 * 254 shift / add / subtract steps on the ALU
// pipe plus two Montgomery products, instead of the ~384 products of the Fermat ladder — 4-5x fewer issue cycles, and
// it leaves the multiplier pipe alone. Phase 1 ("almost inverse") returns x^-1 * 2^k mod p with 254 <= k <= 508 for
// the stored residue x = a R; the Montgomery form of a^-1 is x^-1 R^2 = (x^-1 2^k) * R^2 * 2^(512-k) / R / R.
template <class P>
FF_HD Fe<P> fe_inv(const Fe<P>& a) {
  if (fe_is_zero<P>(a)) return a;
  uint32_t u[8], v[8], r[8], s[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    u[i] = P::mod(i);
    v[i] = a.v[i];
    r[i] = 0;
    s[i] = 0;
  }
  s[0] = 1;
  int k = 0;
  for (;;) {
    uint32_t nz = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) nz |= v[i];
    if (!nz) break;
    // e = v - u (borrow: u > v), d = u - v = -e, sum = r + s   (r, s < 2p < 2^255: no overflow)
    uint32_t e[8], d[8], sum[8];
    uint64_t bw = 0, cy = 0, ng = 1;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint64_t t = (uint64_t)v[i] - u[i] - bw;
      e[i] = (uint32_t)t;
      bw = (t >> 32) & 1;
      const uint64_t w = (uint64_t)r[i] + s[i] + cy;
      sum[i] = (uint32_t)w;
      cy = w >> 32;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint64_t t = (uint64_t)(~e[i]) + ng;
      d[i] = (uint32_t)t;
      ng = t >> 32;
    }
    const bool ue = !(u[0] & 1), ve = !(v[0] & 1), ugt = bw != 0;
    // A: u even | B: v even | C: u > v | D: otherwise        (first match wins)
    const uint32_t mA = ue ? 0xffffffffu : 0u;
    const uint32_t mB = (!ue && ve) ? 0xffffffffu : 0u;
    const uint32_t mC = (!ue && !ve && ugt) ? 0xffffffffu : 0u;
    const uint32_t mD = ~(mA | mB | mC);
    // u' = A ? u >> 1 : C ? d >> 1 : u;   v' = B ? v >> 1 : D ? e >> 1 : v
    // r' = C ? r + s : (B | D) ? r << 1 : r;   s' = D ? r + s : (A | C) ? s << 1 : s
    uint32_t nu[8], nv[8], nr[8], ns[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t uh = i < 7 ? u[i + 1] : 0u, dh = i < 7 ? d[i + 1] : 0u;
      const uint32_t vh = i < 7 ? v[i + 1] : 0u, eh = i < 7 ? e[i + 1] : 0u;
      const uint32_t rl = i ? r[i - 1] : 0u, sl = i ? s[i - 1] : 0u;
      const uint32_t us = (u[i] >> 1) | (uh << 31), ds = (d[i] >> 1) | (dh << 31);
      const uint32_t vs = (v[i] >> 1) | (vh << 31), es = (e[i] >> 1) | (eh << 31);
      const uint32_t r2 = (r[i] << 1) | (rl >> 31), s2 = (s[i] << 1) | (sl >> 31);
      nu[i] = (mA & us) | (mC & ds) | (~(mA | mC) & u[i]);
      nv[i] = (mB & vs) | (mD & es) | (~(mB | mD) & v[i]);
      nr[i] = (mC & sum[i]) | ((mB | mD) & r2) | (mA & r[i]);
      ns[i] = (mD & sum[i]) | ((mA | mC) & s2) | (mB & s[i]);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      u[i] = nu[i];
      v[i] = nv[i];
      r[i] = nr[i];
      s[i] = ns[i];
    }
    ++k;
  }
  // r < 2p: reduce once, then negate: x^-1 2^k = p - r
  Fe<P> t;
  {
    uint64_t bw = 0;
    uint32_t w[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint64_t x = (uint64_t)r[i] - P::mod(i) - bw;
      w[i] = (uint32_t)x;
      bw = (x >> 32) & 1;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) t.v[i] = bw ? r[i] : w[i];
  }
  t = fe_neg<P>(t);
  // * R^2 / R = * R, then * 2^(512-k) / R
  Fe<P> r2c;
#pragma unroll
  for (int i = 0; i < 8; ++i) r2c.v[i] = P::r2(i);
  t = fe_mul<P>(t, r2c);
  int j = 512 - k;  // 4 .. 258
  int extra = 0;
  if (j > 253) {
    extra = j - 253;
    j = 253;
  }
  Fe<P> pw = fe_zero<P>();
#pragma unroll
  for (int i = 0; i < 8; ++i) pw.v[i] = (j >> 5) == i ? (1u << (j & 31)) : 0u;
  t = fe_mul<P>(t, pw);
  for (int i = 0; i < extra; ++i) t = fe_dbl<P>(t);
  return t;
}

// ---- 256-bit global memory access --------------------------------------------------------------
#if defined(__CUDA_ARCH__)
template <class P>
FF_D Fe<P> fe_ldg(const Fe<P>* p) {  // read-only, streaming (one LDG.E.256)
  Fe<P> r;
  asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p));
  return r;
}
template <class P>
FF_D Fe<P> fe_ld(const Fe<P>* p) {  // coherent 256-bit load (data written earlier in the same kernel)
  Fe<P> r;
  asm volatile("ld.global.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3]), "=r"(r.v[4]), "=r"(r.v[5]),
                 "=r"(r.v[6]), "=r"(r.v[7])
               : "l"(p)
               : "memory");
  return r;
}
template <class P>
FF_D void fe_st(Fe<P>* p, const Fe<P>& a) {  // one STG.E.256
  asm volatile("st.global.v8.u32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "r"(a.v[0]), "r"(a.v[1]),
               "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7])
               : "memory");
}
#else
template <class P>
inline Fe<P> fe_ldg(const Fe<P>* p) { return *p; }
template <class P>
inline Fe<P> fe_ld(const Fe<P>* p) { return *p; }
template <class P>
inline void fe_st(Fe<P>* p, const Fe<P>& a) { *p = a; }
#endif

// convenience operators for Fr (the sum-check / MLE kernels read much better with them)
FF_HD Fr operator+(const Fr& a, const Fr& b) { return fe_add<FrP>(a, b); }
FF_HD Fr operator-(const Fr& a, const Fr& b) { return fe_sub<FrP>(a, b); }
FF_HD Fr operator*(const Fr& a, const Fr& b) { return fe_mul<FrP>(a, b); }
FF_HD Fq operator+(const Fq& a, const Fq& b) { return fe_add<FqP>(a, b); }
FF_HD Fq operator-(const Fq& a, const Fq& b) { return fe_sub<FqP>(a, b); }
FF_HD Fq operator*(const Fq& a, const Fq& b) { return fe_mul<FqP>(a, b); }

}  // namespace b200

// BN254 G1 arithmetic for the MSM kernels: affine inputs (halo2_curves::bn256::G1Affine layout) and
// extended-Jacobian "XYZZ" accumulators (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity: ZZ = 0).
// Mixed addition costs 8M + 2S — the cheapest form for bucket accumulation, which is where an MSM
// spends its time (pb/util/arithmetic/msm.rs:168-173 does the same job with Jacobian mixed adds).
// Group results are unique, so only the final affine (x, y) has to match the reference.
// FF_HD: the same code is unit-tested on the host (tests/test_ff32_host.py) against the oracle.
#pragma once
#include "ff32.cuh"

namespace b200 {

struct G1Aff {  // 64 bytes, (x, y) Montgomery; identity = (0, 0)
  Fq x, y;
};
struct G1Xyzz {
  Fq x, y, zz, zzz;
};

FF_HD bool g1_aff_is_identity(const G1Aff& p) { return fe_is_zero<FqP>(p.x) && fe_is_zero<FqP>(p.y); }
FF_HD bool g1_is_identity(const G1Xyzz& p) { return fe_is_zero<FqP>(p.zz); }
FF_HD G1Xyzz g1_identity() {
  G1Xyzz r;
  r.x = fe_zero<FqP>();
  r.y = fe_one<FqP>();
  r.zz = fe_zero<FqP>();
  r.zzz = fe_zero<FqP>();
  return r;
}
FF_HD G1Xyzz g1_from_affine(const G1Aff& p) {
  if (g1_aff_is_identity(p)) return g1_identity();
  G1Xyzz r;
  r.x = p.x;
  r.y = p.y;
  r.zz = fe_one<FqP>();
  r.zzz = fe_one<FqP>();
  return r;
}

// dbl-2008-s-1 (a = 0)
FF_HD G1Xyzz g1_dbl(const G1Xyzz& p) {
  if (g1_is_identity(p)) return p;
  const Fq u = fe_dbl<FqP>(p.y);
  const Fq v = fe_sqr<FqP>(u);
  const Fq w = u * v;
  const Fq s = p.x * v;
  const Fq xx = fe_sqr<FqP>(p.x);
  const Fq m = fe_dbl<FqP>(xx) + xx;
  G1Xyzz r;
  r.x = fe_sqr<FqP>(m) - fe_dbl<FqP>(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v * p.zz;
  r.zzz = w * p.zzz;
  return r;
}

// mdbl-2008-s-1: double an affine point
FF_HD G1Xyzz g1_dbl_affine(const G1Aff& p) {
  const Fq u = fe_dbl<FqP>(p.y);
  const Fq v = fe_sqr<FqP>(u);
  const Fq w = u * v;
  const Fq s = p.x * v;
  const Fq xx = fe_sqr<FqP>(p.x);
  const Fq m = fe_dbl<FqP>(xx) + xx;
  G1Xyzz r;
  r.x = fe_sqr<FqP>(m) - fe_dbl<FqP>(s);
  r.y = m * (s - r.x) - w * p.y;
  r.zz = v;
  r.zzz = w;
  return r;
}

// madd-2008-s: acc += (x2, y2) affine, y2 negated when `neg`
FF_HD G1Xyzz g1_add_affine(const G1Xyzz& a, const G1Aff& b_in, bool neg) {
  if (g1_aff_is_identity(b_in)) return a;
  G1Aff b = b_in;
  if (neg) b.y = fe_neg<FqP>(b.y);
  if (g1_is_identity(a)) return g1_from_affine(b);
  const Fq u2 = b.x * a.zz;
  const Fq s2 = b.y * a.zzz;
  const Fq p = u2 - a.x;
  const Fq r = s2 - a.y;
  if (fe_is_zero<FqP>(p)) {
    if (fe_is_zero<FqP>(r)) return g1_dbl_affine(b);
    return g1_identity();
  }
  const Fq pp = fe_sqr<FqP>(p);
  const Fq ppp = p * pp;
  const Fq q = a.x * pp;
  G1Xyzz o;
  o.x = fe_sqr<FqP>(r) - ppp - fe_dbl<FqP>(q);
  o.y = r * (q - o.x) - a.y * ppp;
  o.zz = a.zz * pp;
  o.zzz = a.zzz * ppp;
  return o;
}

// add-2008-s
FF_HD G1Xyzz g1_add(const G1Xyzz& a, const G1Xyzz& b) {
  if (g1_is_identity(a)) return b;
  if (g1_is_identity(b)) return a;
  const Fq u1 = a.x * b.zz;
  const Fq u2 = b.x * a.zz;
  const Fq s1 = a.y * b.zzz;
  const Fq s2 = b.y * a.zzz;
  const Fq p = u2 - u1;
  const Fq r = s2 - s1;
  if (fe_is_zero<FqP>(p)) {
    if (fe_is_zero<FqP>(r)) return g1_dbl(a);
    return g1_identity();
  }
  const Fq pp = fe_sqr<FqP>(p);
  const Fq ppp = p * pp;
  const Fq q = u1 * pp;
  G1Xyzz o;
  o.x = fe_sqr<FqP>(r) - ppp - fe_dbl<FqP>(q);
  o.y = r * (q - o.x) - s1 * ppp;
  o.zz = a.zz * b.zz * pp;
  o.zzz = a.zzz * b.zzz * ppp;
  return o;
}

FF_HD G1Xyzz g1_neg(const G1Xyzz& a) {
  G1Xyzz r = a;
  r.y = fe_neg<FqP>(a.y);
  return r;
}

// k * P for a small non-negative k (double-and-add, MSB first)
FF_HD G1Xyzz g1_mul_small(const G1Xyzz& p, uint32_t k) {
  G1Xyzz acc = g1_identity();
  int top = 31;
  while (top >= 0 && !((k >> top) & 1)) --top;  // skip the leading zero bits
  for (int i = top; i >= 0; --i) {
    acc = g1_dbl(acc);
    if ((k >> i) & 1) acc = g1_add(acc, p);
  }
  return acc;
}

FF_HD G1Aff g1_to_affine(const G1Xyzz& p) {
  G1Aff r;
  if (g1_is_identity(p)) {
    r.x = fe_zero<FqP>();
    r.y = fe_zero<FqP>();
    return r;
  }
  const Fq w = fe_inv<FqP>(p.zz * p.zzz);  // 1/(ZZ*ZZZ)
  r.x = p.x * (w * p.zzz);                  // X / ZZ
  r.y = p.y * (w * p.zz);                   // Y / ZZZ
  return r;
}

}  // namespace b200

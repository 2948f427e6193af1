// CPU verifier: BN254 optimal-ate pairing for `M::pairings_product_is_identity` (pb/util/arithmetic.rs:25-32 =
// multi_miller_loop + final_exponentiation + is_identity), the check of MultilinearKzg::verify (kzg.rs:330-361).
//
//   Fq2 = Fq[u]/(u^2+1), xi = 9+u, Fq6 = Fq2[v]/(v^3-xi), Fq12 = Fq6[w]/(w^2-v)
//   G2 = E'(Fq2)[r], E': y^2 = x^3 + 3/xi (D-type twist), untwist (x', y') -> (x' w^2, y' w^3)
//   e(P, Q) = f_{6u+2,Q}(P) * l_{[6u+2]Q, pi(Q)}(P) * l_{[6u+2]Q + pi(Q), -pi^2(Q)}(P)  raised to (p^12-1)/r,
//   u = 4965661367192848881.
//
// Written for clarity first: affine G2 steps and dense Fq12 products in the Miller loop (~3 ms per term on a host core),
// one final exponentiation per product check (easy part by conjugation / inversion / p^2-Frobenius, then a 761-bit power).
#pragma once
#include <utility>
#include <vector>

#include "g1.hpp"

namespace b200v {

struct Fq2 {
  Fq c0, c1;
  static Fq2 zero() { return {Fq::zero(), Fq::zero()}; }
  static Fq2 one() { return {Fq::one(), Fq::zero()}; }
  bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
  bool operator==(const Fq2& o) const { return c0 == o.c0 && c1 == o.c1; }
  Fq2 operator+(const Fq2& o) const { return {c0 + o.c0, c1 + o.c1}; }
  Fq2 operator-(const Fq2& o) const { return {c0 - o.c0, c1 - o.c1}; }
  Fq2 operator-() const { return {-c0, -c1}; }
  Fq2 operator*(const Fq2& o) const {  // (a + bu)(c + du) = ac - bd + (ad + bc)u
    const Fq ac = c0 * o.c0, bd = c1 * o.c1;
    return {ac - bd, (c0 + c1) * (o.c0 + o.c1) - ac - bd};
  }
  Fq2 scale(const Fq& k) const { return {c0 * k, c1 * k}; }
  Fq2 sqr() const { return *this * *this; }
  Fq2 dbl() const { return *this + *this; }
  Fq2 conj() const { return {c0, -c1}; }
  Fq2 inv() const {
    const Fq n = (c0.sqr() + c1.sqr()).inv();
    return {c0 * n, -(c1 * n)};
  }
  Fq2 mul_xi() const {  // (a + bu)(9 + u) = 9a - b + (a + 9b)u
    const Fq nine = Fq::from_u64(9);
    return {c0 * nine - c1, c0 + c1 * nine};
  }
};

struct Fq6 {
  Fq2 a0, a1, a2;  // a0 + a1 v + a2 v^2
  static Fq6 zero() { return {Fq2::zero(), Fq2::zero(), Fq2::zero()}; }
  static Fq6 one() { return {Fq2::one(), Fq2::zero(), Fq2::zero()}; }
  bool operator==(const Fq6& o) const { return a0 == o.a0 && a1 == o.a1 && a2 == o.a2; }
  Fq6 operator+(const Fq6& o) const { return {a0 + o.a0, a1 + o.a1, a2 + o.a2}; }
  Fq6 operator-(const Fq6& o) const { return {a0 - o.a0, a1 - o.a1, a2 - o.a2}; }
  Fq6 operator*(const Fq6& o) const {  // schoolbook with v^3 = xi
    const Fq2 t0 = a0 * o.a0, t1 = a1 * o.a1, t2 = a2 * o.a2;
    const Fq2 c0 = t0 + (a1 * o.a2 + a2 * o.a1).mul_xi();
    const Fq2 c1 = a0 * o.a1 + a1 * o.a0 + t2.mul_xi();
    const Fq2 c2 = a0 * o.a2 + a2 * o.a0 + t1;
    return {c0, c1, c2};
  }
  Fq6 mul_v() const { return {a2.mul_xi(), a0, a1}; }  // (a0 + a1 v + a2 v^2) v
  Fq6 operator-() const { return {-a0, -a1, -a2}; }
  Fq6 scale(const Fq2& k) const { return {a0 * k, a1 * k, a2 * k}; }
  Fq6 inv() const {  // adjugate over Fq2 with v^3 = xi
    const Fq2 t0 = a0.sqr() - (a1 * a2).mul_xi(), t1 = a2.sqr().mul_xi() - a0 * a1, t2 = a1.sqr() - a0 * a2;
    const Fq2 d = (a0 * t0 + (a2 * t1 + a1 * t2).mul_xi()).inv();
    return {t0 * d, t1 * d, t2 * d};
  }
};

struct Fq12 {
  Fq6 c0, c1;  // c0 + c1 w, w^2 = v
  static Fq12 one() { return {Fq6::one(), Fq6::zero()}; }
  bool operator==(const Fq12& o) const { return c0 == o.c0 && c1 == o.c1; }
  Fq12 operator*(const Fq12& o) const {
    const Fq6 aa = c0 * o.c0, bb = c1 * o.c1;
    return {aa + bb.mul_v(), (c0 + c1) * (o.c0 + o.c1) - aa - bb};
  }
  Fq12 sqr() const { return *this * *this; }
  Fq12 conj() const { return {c0, -c1}; }  // x^(p^6)
  Fq12 inv() const {                       // (c0 - c1 w) / (c0^2 - v c1^2)
    const Fq6 d = (c0 * c0 - (c1 * c1).mul_v()).inv();
    return {c0 * d, -(c1 * d)};
  }
  // x^(p^2): coefficient of w^k times gamma^k, gamma = xi^((p^2-1)/6) in Fq (a primitive 6th root of unity)
  Fq12 frobenius_p2() const {
    static const uint64_t G1_[4] = {0xe4bd44e5607cfd49ULL, 0xc28f069fbb966e3dULL, 0x5e6dd9e7e0acccb0ULL, 0x30644e72e131a029ULL};
    static const uint64_t G2_[4] = {0xe4bd44e5607cfd48ULL, 0xc28f069fbb966e3dULL, 0x5e6dd9e7e0acccb0ULL, 0x30644e72e131a029ULL};
    static const uint64_t G3_[4] = {0x3c208c16d87cfd46ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const uint64_t G4_[4] = {0x5763473177fffffeULL, 0xd4f263f1acdb5c4fULL, 0x59e26bcea0d48bacULL, 0x0000000000000000ULL};
    static const uint64_t G5_[4] = {0x5763473177ffffffULL, 0xd4f263f1acdb5c4fULL, 0x59e26bcea0d48bacULL, 0x0000000000000000ULL};
    const Fq g1 = Fq::from_raw(G1_), g2 = Fq::from_raw(G2_), g3 = Fq::from_raw(G3_), g4 = Fq::from_raw(G4_), g5 = Fq::from_raw(G5_);
    // w^0 -> c0.a0, w^1 -> c1.a0, w^2 -> c0.a1, w^3 -> c1.a1, w^4 -> c0.a2, w^5 -> c1.a2
    return {{c0.a0, c0.a1.scale(g2), c0.a2.scale(g4)}, {c1.a0.scale(g1), c1.a1.scale(g3), c1.a2.scale(g5)}};
  }
};

struct G2Affine {
  Fq2 x, y;
  bool inf;
  static G2Affine identity() { return {Fq2::zero(), Fq2::zero(), true}; }
  static G2Affine generator() {
    static const uint64_t X0[4] = {0x46debd5cd992f6edULL, 0x674322d4f75edaddULL, 0x426a00665e5c4479ULL, 0x1800deef121f1e76ULL};
    static const uint64_t X1[4] = {0x97e485b7aef312c2ULL, 0xf1aa493335a9e712ULL, 0x7260bfb731fb5d25ULL, 0x198e9393920d483aULL};
    static const uint64_t Y0[4] = {0x4ce6cc0166fa7daaULL, 0xe3d1e7690c43d37bULL, 0x4aab71808dcb408fULL, 0x12c85ea5db8c6debULL};
    static const uint64_t Y1[4] = {0x55acdadcd122975bULL, 0xbc4b313370b38ef3ULL, 0xec9e99ad690c3395ULL, 0x090689d0585ff075ULL};
    return {{Fq::from_raw(X0), Fq::from_raw(X1)}, {Fq::from_raw(Y0), Fq::from_raw(Y1)}, false};
  }
  static Fq2 b() { return Fq2{Fq::from_u64(3), Fq::zero()} * Fq2{Fq::from_u64(9), Fq::one()}.inv(); }  // 3 / xi
  bool on_curve() const { return inf || y.sqr() == x.sqr() * x + b(); }
  bool operator==(const G2Affine& o) const { return inf == o.inf && (inf || (x == o.x && y == o.y)); }
  G2Affine neg() const { return inf ? *this : G2Affine{x, -y, false}; }
  // chord / tangent step; `slope` receives lambda when the result is finite
  G2Affine add(const G2Affine& o, Fq2* slope = nullptr) const {
    if (inf) return o;
    if (o.inf) return *this;
    Fq2 lam;
    if (x == o.x) {
      if (!(y == o.y) || y.is_zero()) return identity();
      lam = (x.sqr().dbl() + x.sqr()) * y.dbl().inv();
    } else {
      lam = (o.y - y) * (o.x - x).inv();
    }
    if (slope) *slope = lam;
    const Fq2 x3 = lam.sqr() - x - o.x;
    return {x3, lam * (x - x3) - y, false};
  }
  G2Affine mul(const Fr& k) const {
    uint64_t e[4];
    k.to_raw(e);
    G2Affine acc = identity();
    for (int i = 255; i >= 0; --i) {
      acc = acc.add(acc);
      if ((e[i >> 6] >> (i & 63)) & 1) acc = acc.add(*this);
    }
    return acc;
  }
};

// line through the untwisted T with twist-slope lambda, evaluated at P = (xp, yp) in G1:
//   l = yp + (-lambda xp) w + (lambda x_T - y_T) w^3,   w^3 = v w
inline Fq12 line_eval(const G2Affine& t, const Fq2& lam, const G1Affine& p) {
  Fq12 l;
  l.c0 = {Fq2{p.y, Fq::zero()}, Fq2::zero(), Fq2::zero()};
  l.c1 = {-(lam.scale(p.x)), lam * t.x - t.y, Fq2::zero()};
  return l;
}
// one Miller step: f *= l_{T, Q}(P), T += Q (Q == T: tangent). Vertical lines (T + Q = O) contribute 1: they lie
// in a proper subfield and are wiped out by the final exponentiation; they do not occur for points of order r.
inline void miller_step(Fq12* f, G2Affine* t, const G2Affine& q, const G1Affine& p) {
  Fq2 lam;
  const G2Affine sum = t->add(q, &lam);
  if (!sum.inf) *f = *f * line_eval(*t, lam, p);
  *t = sum;
}

inline Fq12 miller_loop(const G1Affine& p, const G2Affine& q) {
  if (p.is_identity() || q.inf) return Fq12::one();
  static const uint64_t FROB_X0[4] = {0x99e39557176f553dULL, 0xb78cc310c2c3330cULL, 0x4c0bec3cf559b143ULL, 0x2fb347984f7911f7ULL};
  static const uint64_t FROB_X1[4] = {0x1665d51c640fcba2ULL, 0x32ae2a1d0b7c9dceULL, 0x4ba4cc8bd75a0794ULL, 0x16c9e55061ebae20ULL};
  static const uint64_t FROB_Y0[4] = {0xdc54014671a0135aULL, 0xdbaae0eda9c95998ULL, 0xdc5ec698b6e2f9b9ULL, 0x063cf305489af5dcULL};
  static const uint64_t FROB_Y1[4] = {0x82d37f632623b0e3ULL, 0x21807dc98fa25bd2ULL, 0x0704b5a7ec796f2bULL, 0x07c03cbcac41049aULL};
  static const uint64_t FROB2_X[4] = {0xe4bd44e5607cfd48ULL, 0xc28f069fbb966e3dULL, 0x5e6dd9e7e0acccb0ULL, 0x30644e72e131a029ULL};
  const unsigned __int128 loop = ((unsigned __int128)1 << 64) | 0x9d797039be763ba8ULL;  // 6u + 2 = 29793968203157093288
  Fq12 f = Fq12::one();
  G2Affine t = q;
  for (int i = 63; i >= 0; --i) {  // bit 64 is the leading one
    f = f.sqr();
    miller_step(&f, &t, t, p);
    if ((loop >> i) & 1) miller_step(&f, &t, q, p);
  }
  // pi(Q) = (conj(x) xi^((p-1)/3), conj(y) xi^((p-1)/2));  -pi^2(Q) = (x xi^((p^2-1)/3), y)
  const Fq2 gx{Fq::from_raw(FROB_X0), Fq::from_raw(FROB_X1)}, gy{Fq::from_raw(FROB_Y0), Fq::from_raw(FROB_Y1)};
  const G2Affine q1{q.x.conj() * gx, q.y.conj() * gy, false};
  const G2Affine q2n{q.x.scale(Fq::from_raw(FROB2_X)), q.y, false};
  miller_step(&f, &t, q1, p);
  miller_step(&f, &t, q2n, p);
  return f;
}

// f^((p^12-1)/r) by one plain square-and-multiply: the definition, kept as the cross-check of the split version below
inline Fq12 final_exponentiation_plain(const Fq12& f) {
  static const uint64_t E[44] = {  // (p^12 - 1) / r, little-endian
      0x86964b64ca86f120ULL, 0x40a4efb7e54523a4ULL, 0x837fa97896e84abbULL, 0x361102b6b9b2b918ULL, 0xc0de81def35692daULL,
      0xbe04c7e8a6c3c760ULL, 0xd766f9c9d570bb7fULL, 0xc230974d83561841ULL, 0x5bba1668c3be69a3ULL, 0x7f3811c410526294ULL,
      0x29baee7ddadda71cULL, 0xbf813b8d145da900ULL, 0x641bbadf423f9a2cULL, 0xa80bb4ea44eacc5eULL, 0xcd65664814fde37cULL,
      0x4a0364b9580291d2ULL, 0xee93dfb10826f0ddULL, 0x6b42db8dc5514724ULL, 0xbb10cf430b0f3785ULL, 0x40494e406f804216ULL,
      0x55cfe107acf3aafbULL, 0x2088ec80e0ebae87ULL, 0x846a3ed011a337a0ULL, 0x48a45a4a1e3a5195ULL, 0xe5664568dfc50e16ULL,
      0xab6a41294c0cc4ebULL, 0x82d0d602d268c7daULL, 0x6668449aed3cc48aULL, 0x5062cd0fb2015dfcULL, 0x7f2940a8b1ddb3d1ULL,
      0x77f5b63a2a226448ULL, 0xfef0781361e443aeULL, 0xf977870e88d5c6c8ULL, 0x790364a61f676baaULL, 0x5887e72eceaddea3ULL,
      0x1377e563a09a1b70ULL, 0x0c54efee1bd8c3b2ULL, 0x3ec3d15ad524d8f7ULL, 0xdaf15466b2383a5dULL, 0xe1e30a73bb94fec0ULL,
      0x6a1c71015f3f7be2ULL, 0x842d43bf6369b1ffULL, 0x20fddadf107d20bcULL, 0x0000002f4b6dc970ULL};
  Fq12 acc = Fq12::one();
  for (int i = 44 * 64 - 1; i >= 0; --i) {
    acc = acc.sqr();
    if ((E[i >> 6] >> (i & 63)) & 1) acc = acc * f;
  }
  return acc;
}


// f^((p^12-1)/r) = ((f^(p^6-1))^(p^2+1))^((p^4-p^2+1)/r): the first two factors cost a conjugation, one inversion and
// one p^2-Frobenius; only the last one is an exponentiation (761 bits instead of 2790)
inline Fq12 final_exponentiation(const Fq12& f) {
  static const uint64_t H[12] = {  // (p^4 - p^2 + 1) / r, little-endian
      0xe81bb482ccdf42b1ULL, 0x5abf5cc4f49c36d4ULL, 0xf1154e7e1da014fdULL, 0xdcc7b44c87cdbacfULL, 0xaaa441e3954bcf8aULL,
      0x6b887d56d5095f23ULL, 0x79581e16f3fd90c6ULL, 0x3b1b1355d189227dULL, 0x4e529a5861876f6bULL, 0x6c0eb522d5b12278ULL,
      0x331ec15183177fafULL, 0x01baaa710b0759adULL};
  const Fq12 t = f.conj() * f.inv();        // f^(p^6 - 1); f != 0 for Miller-loop outputs
  const Fq12 u = t.frobenius_p2() * t;      // ^(p^2 + 1)
  Fq12 acc = Fq12::one();
  for (int i = 12 * 64 - 1; i >= 0; --i) {
    acc = acc.sqr();
    if ((H[i >> 6] >> (i & 63)) & 1) acc = acc * u;
  }
  return acc;
}

inline Fq12 pairing(const G1Affine& p, const G2Affine& q) { return final_exponentiation(miller_loop(p, q)); }

// pb/util/arithmetic.rs:25-32
inline bool pairings_product_is_identity(const std::vector<std::pair<G1Affine, G2Affine>>& terms) {
  // the Miller loops are independent: one per host thread when built with OpenMP, then ONE final exponentiation
  std::vector<Fq12> ml(terms.size(), Fq12::one());
  const long n = (long)terms.size();
#pragma omp parallel for schedule(dynamic, 1) if (n > 1)
  for (long i = 0; i < n; ++i) ml[i] = miller_loop(terms[i].first, terms[i].second);
  Fq12 f = Fq12::one();
  for (auto& m : ml) f = f * m;
  return final_exponentiation(f) == Fq12::one();
}

}  // namespace b200v

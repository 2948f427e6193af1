// CPU verifier: BN254 G1 (y^2 = x^3 + 3, generator (1, 2)) with Jacobian arithmetic; only the handful of scalar
// multiplications of `MultilinearKzg::verify` / `additive::batch_verify` (pb/pcs/multilinear/kzg.rs:330-361,
// pb/pcs/multilinear.rs:237-275) run here.
#pragma once
#include <vector>

#include "field.hpp"

namespace b200v {

struct G1Affine {
  Fq x, y;  // identity is encoded as (0, 0), as halo2curves does
  bool is_identity() const { return x.is_zero() && y.is_zero(); }
  static G1Affine identity() { return G1Affine{Fq::zero(), Fq::zero()}; }
  static G1Affine generator() { return G1Affine{Fq::from_u64(1), Fq::from_u64(2)}; }
  bool operator==(const G1Affine& o) const { return x == o.x && y == o.y; }
  bool on_curve() const {
    if (is_identity()) return true;
    return y.sqr() == x.sqr() * x + Fq::from_u64(3);
  }
  G1Affine neg() const { return is_identity() ? *this : G1Affine{x, -y}; }
};

struct G1 {
  Fq x, y, z;  // Jacobian: (x/z^2, y/z^3); identity has z == 0
  static G1 identity() { return G1{Fq::zero(), Fq::one(), Fq::zero()}; }
  static G1 from_affine(const G1Affine& p) {
    if (p.is_identity()) return identity();
    return G1{p.x, p.y, Fq::one()};
  }
  bool is_identity() const { return z.is_zero(); }

  G1 dbl() const {
    if (is_identity()) return *this;
    // dbl-2009-l (a = 0)
    Fq a = x.sqr(), b = y.sqr(), c = b.sqr();
    Fq d = ((x + b).sqr() - a - c).dbl();
    Fq e = a.dbl() + a, f = e.sqr();
    G1 r;
    r.x = f - d.dbl();
    r.y = e * (d - r.x) - c.dbl().dbl().dbl();
    r.z = (y * z).dbl();
    return r;
  }

  G1 add(const G1& o) const {
    if (is_identity()) return o;
    if (o.is_identity()) return *this;
    Fq z1z1 = z.sqr(), z2z2 = o.z.sqr();
    Fq u1 = x * z2z2, u2 = o.x * z1z1;
    Fq s1 = y * o.z * z2z2, s2 = o.y * z * z1z1;
    if (u1 == u2) {
      if (s1 == s2) return dbl();
      return identity();
    }
    Fq h = u2 - u1, i = h.dbl().sqr(), j = h * i, rr = (s2 - s1).dbl(), v = u1 * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (s1 * j).dbl();
    r.z = ((z + o.z).sqr() - z1z1 - z2z2) * h;
    return r;
  }

  G1 add_affine(const G1Affine& o) const {
    if (o.is_identity()) return *this;
    if (is_identity()) return from_affine(o);
    Fq z1z1 = z.sqr();
    Fq u2 = o.x * z1z1, s2 = o.y * z * z1z1;
    if (x == u2) {
      if (y == s2) return dbl();
      return identity();
    }
    Fq h = u2 - x, hh = h.sqr(), i = hh.dbl().dbl(), j = h * i, rr = (s2 - y).dbl(), v = x * i;
    G1 r;
    r.x = rr.sqr() - j - v.dbl();
    r.y = rr * (v - r.x) - (y * j).dbl();
    r.z = (z + h).sqr() - z1z1 - hh;
    return r;
  }

  G1 neg() const { return G1{x, -y, z}; }

  // scalar given as canonical integer limbs
  G1 mul_raw(const uint64_t k[4]) const {
    G1 acc = identity();
    for (int i = 255; i >= 0; --i) {
      acc = acc.dbl();
      if ((k[i / 64] >> (i % 64)) & 1) acc = acc.add(*this);
    }
    return acc;
  }
  G1 mul(const Fr& k) const {
    uint64_t raw[4];
    k.to_raw(raw);
    return mul_raw(raw);
  }

  G1Affine to_affine() const {
    if (is_identity()) return G1Affine::identity();
    Fq zi = z.inv(), zi2 = zi.sqr();
    return G1Affine{x * zi2, y * zi2 * zi};
  }
  bool eq(const G1& o) const { return to_affine() == o.to_affine(); }
};

// `Curve::batch_normalize` (pb/pcs/multilinear/kzg.rs:206): one shared inversion.
inline void batch_normalize(const G1* in, G1Affine* out, size_t n) {
  std::vector<Fq> zs(n);
  for (size_t i = 0; i < n; ++i) zs[i] = in[i].z;
  batch_invert(zs.data(), n);
  for (size_t i = 0; i < n; ++i) {
    if (in[i].is_identity()) {
      out[i] = G1Affine::identity();
    } else {
      Fq zi2 = zs[i].sqr();
      out[i] = G1Affine{in[i].x * zi2, in[i].y * zi2 * zs[i]};
    }
  }
}

}  // namespace b200v

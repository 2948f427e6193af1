// CPU verifier: the verifying halves of the reference's traits, restated in host C++ —
//   SumCheck::verify                      pb/piop/sum_check/classic.rs:175-194, 242-263 (Evaluations and Coefficients messages)
//   MultilinearKzg::verify                pb/pcs/multilinear/kzg.rs:330-361 (pairing product)
//   additive::batch_verify                pb/pcs/multilinear.rs:237-275
//   HyperPlonk::verify                    pb/backend/hyperplonk.rs:293-363, hyperplonk/verifier.rs:39-182
//   evaluate / lagrange_eval / eq_xy_eval pb/piop/sum_check.rs:60-125
//   rotation_eval(_points)                pb/poly/multilinear.rs:433-570
// and the verifier of the Lasso argument this package proves. No prover code lives here: a proof is
// bytes in, accept / reject out.
#pragma once
#include <algorithm>
#include <cmath>
#include <map>
#include <memory>
#include <set>
#include <vector>

#include "pairing.hpp"
#include "transcript.hpp"

namespace b200v {

typedef std::vector<Fr> Poly;


inline Poly eq_xy(const std::vector<Fr>& y) {
  Poly evals(1, Fr::one());
  for (size_t k = y.size(); k-- > 0;) {
    Poly next(2 * evals.size());
    const Fr yk = y[k];
    const long n = (long)evals.size();
    for (long i = 0; i < n; ++i) {
      next[2 * i + 1] = evals[i] * yk;
      next[2 * i] = evals[i] - next[2 * i + 1];
    }
    evals.swap(next);
  }
  return evals;
}

// pb/piop/sum_check.rs:111-121
inline Fr eq_xy_eval(const std::vector<Fr>& x, const std::vector<Fr>& y) {
  Fr acc = Fr::one();
  for (size_t i = 0; i < x.size(); ++i) acc = acc * ((x[i] * y[i]).dbl() + Fr::one() - x[i] - y[i]);
  return acc;
}

// pb/piop/sum_check.rs:123-125: Σ 2^i x_i — the MLE of the map b -> b
inline Fr identity_eval(const std::vector<Fr>& x) {
  Fr acc = Fr::zero(), pw = Fr::one();
  for (size_t i = 0; i < x.size(); ++i) {
    acc = acc + x[i] * pw;
    pw = pw.dbl();
  }
  return acc;
}

// arithmetic.rs:108-123
inline std::vector<Fr> barycentric_weights(const std::vector<Fr>& points) {
  std::vector<Fr> w(points.size());
  for (size_t j = 0; j < points.size(); ++j) {
    Fr acc = Fr::one();
    bool any = false;
    for (size_t i = 0; i < points.size(); ++i) {
      if (i == j) continue;
      acc = any ? acc * (points[j] - points[i]) : points[j] - points[i];
      any = true;
    }
    w[j] = acc;
  }
  batch_invert(w.data(), w.size());
  return w;
}

// arithmetic.rs:125-136 (kept quirk-for-quirk: BatchInvert skips zeros)
inline Fr barycentric_interpolate(const std::vector<Fr>& weights, const std::vector<Fr>& points,
                                  const std::vector<Fr>& evals, const Fr& x) {
  std::vector<Fr> coeffs(points.size());
  for (size_t i = 0; i < points.size(); ++i) coeffs[i] = x - points[i];
  batch_invert(coeffs.data(), coeffs.size());
  Fr sum = Fr::zero();
  for (size_t i = 0; i < points.size(); ++i) {
    coeffs[i] = coeffs[i] * weights[i];
    sum = sum + coeffs[i];
  }
  Fr ip = Fr::zero();
  for (size_t i = 0; i < points.size(); ++i) ip = ip + coeffs[i] * evals[i];
  return ip * sum.inv();
}

inline std::vector<Fr> points_0_to_d(int d) {
  std::vector<Fr> p(d + 1);
  p[0] = Fr::zero();
  for (int i = 1; i <= d; ++i) p[i] = p[i - 1] + Fr::one();
  return p;
}

// ClassicSumCheck::verify (classic.rs:242-263) + verify_consistency (:175-194).
// coeffs == false: Evaluations messages; true: Coefficients messages.
inline bool sumcheck_verify(int num_vars, int degree, Fr sum, bool coeffs, Transcript& tr,
                            Fr* final_claim, std::vector<Fr>* challenges) {
  std::vector<std::vector<Fr>> msgs(num_vars, std::vector<Fr>(degree + 1));
  challenges->clear();
  for (int i = 0; i < num_vars; ++i) {
    for (int k = 0; k <= degree; ++k)
      if (!tr.read_field_element(&msgs[i][k])) return false;
    challenges->push_back(tr.squeeze_challenge());
  }
  const std::vector<Fr> points = points_0_to_d(degree);
  const std::vector<Fr> weights = coeffs ? std::vector<Fr>() : barycentric_weights(points);
  for (int i = 0; i < num_vars; ++i) {
    const std::vector<Fr>& m = msgs[i];
    Fr msum;
    if (coeffs) {
      msum = m[0].dbl();
      for (int k = 1; k <= degree; ++k) msum = msum + m[k];
    } else {
      msum = m[0] + m[1];
    }
    if (sum != msum) return false;
    const Fr& r = (*challenges)[i];
    if (coeffs) {
      Fr acc = Fr::zero();
      for (int k = degree; k >= 0; --k) acc = acc * r + m[k];
      sum = acc;
    } else {
      sum = barycentric_interpolate(weights, points, m, r);
    }
  }
  *final_claim = sum;
  return true;
}

// verify_fractional_sum_check (pb/piop/gkr/fractional_sum_check.rs:192-265): GKR verifier for Σ_i p_b[i] / q_b[i] over a
// batch of B (p, q) pairs. claimed_*: Some(value) -> absorbed as common input, None (nullptr) -> read from the proof.
// Per layer v = 0 .. num_vars - 1: v = 0 reads the four single entries and checks p = p_l q_r + p_r q_l, q = q_l q_r;
// v > 0: gamma, the degree-3 sum-check of eq * Σ_b [gamma^(2b) (p_l q_r + p_r q_l) + gamma^(2b+1) q_l q_r] for the claim
// Σ gamma^i (p_0, q_0, p_1, q_1, ...) (sum_check_claim :279-284), the 4B evaluations, and the final check against the
// expression at x (:246-249); then mu and layer_down_claim (:290-296). Returns the claims at the input layer and x.
struct FractionalClaims {
  std::vector<Fr> p_xs, q_xs, x, p_0s, q_0s;
};
inline bool fractional_sum_check_verify(int num_vars, const std::vector<const Fr*>& claimed_p_0s,
                                        const std::vector<const Fr*>& claimed_q_0s, Transcript& tr, FractionalClaims* out) {
  const int B = (int)claimed_p_0s.size();
  std::vector<Fr> cp(B), cq(B);
  for (int pass = 0; pass < 2; ++pass) {
    const std::vector<const Fr*>& claimed = pass ? claimed_q_0s : claimed_p_0s;
    std::vector<Fr>& dst = pass ? cq : cp;
    for (int b = 0; b < B; ++b) {
      if (claimed[b]) {
        dst[b] = *claimed[b];
        tr.common_field_element(dst[b]);
      } else if (!tr.read_field_element(&dst[b])) {
        return false;
      }
    }
  }
  out->p_0s = cp;
  out->q_0s = cq;
  std::vector<Fr> y;
  for (int v = 0; v < num_vars; ++v) {
    std::vector<Fr> x, evals(4 * B);
    if (v == 0) {
      for (auto& e : evals)
        if (!tr.read_field_element(&e)) return false;
      for (int b = 0; b < B; ++b) {
        const Fr &pl = evals[4 * b], &pr = evals[4 * b + 1], &ql = evals[4 * b + 2], &qr = evals[4 * b + 3];
        if (cp[b] != pl * qr + pr * ql || cq[b] != ql * qr) return false;
      }
    } else {
      const Fr gamma = tr.squeeze_challenge();
      Fr pw = Fr::one(), claim = Fr::zero();
      for (int b = 0; b < B; ++b) {
        claim = claim + pw * cp[b];
        pw = pw * gamma;
        claim = claim + pw * cq[b];
        pw = pw * gamma;
      }
      Fr fin;
      if (!sumcheck_verify(v, 3, claim, false, tr, &fin, &x)) return false;
      for (auto& e : evals)
        if (!tr.read_field_element(&e)) return false;
      Fr s = Fr::zero();
      pw = Fr::one();
      for (int b = 0; b < B; ++b) {
        const Fr &pl = evals[4 * b], &pr = evals[4 * b + 1], &ql = evals[4 * b + 2], &qr = evals[4 * b + 3];
        s = s + pw * (pl * qr + pr * ql);
        pw = pw * gamma;
        s = s + pw * (ql * qr);
        pw = pw * gamma;
      }
      if (fin != s * eq_xy_eval(x, y)) return false;
    }
    const Fr mu = tr.squeeze_challenge();
    for (int b = 0; b < B; ++b) {
      cp[b] = evals[4 * b] + mu * (evals[4 * b + 1] - evals[4 * b]);
      cq[b] = evals[4 * b + 2] + mu * (evals[4 * b + 3] - evals[4 * b + 2]);
    }
    x.push_back(mu);
    y = x;
  }
  out->p_xs = cp;
  out->q_xs = cq;
  out->x = y;
  return true;
}

// MultilinearKzgVerifierParam (kzg.rs:79-84): g1, g2 are the curve generators, ss_g2[i] = g2 * s_i
struct KzgVerifierParam {
  std::vector<G2Affine> ss_g2;
  int num_vars() const { return (int)ss_g2.size(); }
};
// the verifier half of the seeded test setup (kzg.rs:215-225): ss_g2 from the trapdoor scalars
inline KzgVerifierParam kzg_verifier_setup(const std::vector<Fr>& ss) {
  KzgVerifierParam vp;
  for (const Fr& s : ss) vp.ss_g2.push_back(G2Affine::generator().mul(s));
  return vp;
}

// kzg.rs:330-361: e(C - g1 * eval, -g2) * Π_i e(Q_i, s_i g2 - x_i g2) == 1, evaluated as
// e(C - g1 * eval + Σ_i x_i Q_i, -g2) * Π_i e(Q_i, s_i g2) == 1
inline bool kzg_verify(const KzgVerifierParam& vp, const G1Affine& comm, const std::vector<Fr>& point, const Fr& eval,
                       Transcript& tr) {
  const int n = (int)point.size();
  if (n > vp.num_vars()) return false;  // "Too many variates"
  std::vector<G1Affine> qs(n);
  for (int i = 0; i < n; ++i)
    if (!tr.read_commitment(&qs[i])) return false;
  // e(Q_i, s_i g2 - x_i g2) = e(Q_i, s_i g2) * e(-x_i Q_i, g2): the x_i move to G1 (cheap scalar multiplications),
  // the same n + 1 pairings remain and no G2 arithmetic is needed — the verdict is the reference's by bilinearity
  G1 lhs = G1::from_affine(comm).add(G1::from_affine(G1Affine::generator()).mul(eval).neg());
  for (int i = 0; i < n; ++i) lhs = lhs.add(G1::from_affine(qs[i]).mul(point[i]));
  std::vector<std::pair<G1Affine, G2Affine>> terms;
  terms.push_back({lhs.to_affine(), G2Affine::generator().neg()});
  for (int i = 0; i < n; ++i) terms.push_back({qs[i], vp.ss_g2[i]});
  return pairings_product_is_identity(terms);
}


struct Evaluation {
  int poly, point;
  Fr value;
};

inline int ceil_log2(size_t n) {
  int k = 0;
  while (((size_t)1 << k) < n) ++k;
  return k;
}

// additive::batch_verify, pb/pcs/multilinear.rs:237-275
inline bool kzg_batch_verify(const KzgVerifierParam& vp, int num_vars, const std::vector<G1Affine>& comms,
                             const std::vector<std::vector<Fr>>& points,
                             const std::vector<Evaluation>& evals, Transcript& tr) {
  const int ell = ceil_log2(evals.size());
  std::vector<Fr> t = tr.squeeze_challenges(ell);
  Poly eq_xt = ell ? eq_xy(t) : Poly();
  if (eq_xt.size() < evals.size()) return false;
  Fr tilde = Fr::zero();
  for (size_t k = 0; k < evals.size(); ++k) tilde = tilde + evals[k].value * eq_xt[k];
  Fr g_eval;
  std::vector<Fr> ch;
  if (!sumcheck_verify(num_vars, 2, tilde, true, tr, &g_eval, &ch)) return false;
  std::vector<Fr> eqe(points.size());
  for (size_t i = 0; i < points.size(); ++i) eqe[i] = eq_xy_eval(ch, points[i]);
  G1 gc = G1::identity();
  for (size_t k = 0; k < evals.size(); ++k)
    gc = gc.add(G1::from_affine(comms[evals[k].poly]).mul(eqe[evals[k].point] * eq_xt[k]));
  return kzg_verify(vp, gc.to_affine(), ch, g_eval, tr);
}

// ---- Lasso: table description as far as the verifier needs it ---------------------------------
enum TableKind { TABLE_RANGE = 0, TABLE_AND = 1, TABLE_XOR = 2, TABLE_CUSTOM = 3 };
static const int SUBTABLE_VARS = 16;

struct LassoTable {
  int kind;    // TableKind
  int chunks;  // c; every chunk addresses one 2^16-entry subtable
  // TABLE_CUSTOM (b200_lasso_table of the prover side): the table is data — 2^16 values, operand layout, output stride
  int num_operands = 1, operand_bits = 16, custom_out_bits = 16;
  const uint32_t* values = nullptr;
  // bits of the lookup output contributed by one chunk (16 for range, 8 for and/xor)
  int out_bits() const { return kind == TABLE_CUSTOM ? custom_out_bits : (kind == TABLE_RANGE ? 16 : 8); }
  // MLE of the subtable at a 16-variate point: the verifier evaluates the subtable itself (closed forms for the
  // structured tables, <values, eq(., x)> for a table given as data)
  Fr subtable_mle(const std::vector<Fr>& x) const {
    if (kind == TABLE_CUSTOM) {
      const Poly eq = eq_xy(x);
      Fr acc = Fr::zero();
      for (size_t i = 0; i < eq.size(); ++i)
        if (values[i]) acc = acc + eq[i] * Fr::from_u64(values[i]);
      return acc;
    }
    if (kind == TABLE_RANGE) return identity_eval(x);
    Fr acc = Fr::zero(), pw = Fr::one();
    for (int k = 0; k < 8; ++k) {
      Fr q = x[k], p = x[8 + k];
      Fr bit = kind == TABLE_AND ? p * q : p + q - (p * q).dbl();
      acc = acc + bit * pw;
      pw = pw.dbl();
    }
    return acc;
  }
  // statement binding of a table given as data: Keccak-256 of the values (LE u32 words) as an LE integer mod r
  Fr digest() const {
    Keccak256 h;
    h.update(reinterpret_cast<const uint8_t*>(values), ((size_t)4) << 16);
    uint8_t out[32];
    h.finalize_reset(out);
    return Fr::from_le_bytes_mod(out);
  }
};

struct GrandProductOutput {
  std::vector<Fr> claims;                  // per tree: leaf-layer MLE at points[height of that tree]
  std::vector<std::vector<Fr>> points;     // points[h] = the point after h layers (h variables)
  std::vector<Fr> point;                   // == points[max height]
};

// Batched layered product argument over trees of different heights (verifier side; layer k batches the trees with
// height > k with weights gamma^i, a tree's claim freezes at its leaf layer).
inline bool grand_product_verify(const std::vector<int>& hs, Transcript& tr, std::vector<Fr>* roots,
                                 GrandProductOutput* out) {
  const int T = (int)hs.size();
  int h = 0;
  for (int t = 0; t < T; ++t) h = hs[t] > h ? hs[t] : h;
  std::vector<Fr> claims(T);
  for (int t = 0; t < T; ++t)
    if (!tr.read_field_element(&claims[t])) return false;
  *roots = claims;
  out->points.assign(h + 1, std::vector<Fr>());
  std::vector<Fr> y;
  for (int k = 0; k < h; ++k) {
    std::vector<int> act;
    for (int t = 0; t < T; ++t)
      if (hs[t] > k) act.push_back(t);
    const int A = (int)act.size();
    std::vector<Fr> x, evals(2 * A);
    if (k == 0) {
      for (auto& e : evals)
        if (!tr.read_field_element(&e)) return false;
      for (int i = 0; i < A; ++i)
        if (claims[act[i]] != evals[2 * i] * evals[2 * i + 1]) return false;
    } else {
      Fr gamma = tr.squeeze_challenge();
      Fr pw = Fr::one(), claim = Fr::zero();
      for (int i = 0; i < A; ++i) {
        claim = claim + pw * claims[act[i]];
        pw = pw * gamma;
      }
      Fr fin;
      if (!sumcheck_verify(k, 3, claim, false, tr, &fin, &x)) return false;
      for (auto& e : evals)
        if (!tr.read_field_element(&e)) return false;
      Fr s = Fr::zero();
      pw = Fr::one();
      for (int i = 0; i < A; ++i) {
        s = s + pw * evals[2 * i] * evals[2 * i + 1];
        pw = pw * gamma;
      }
      if (fin != s * eq_xy_eval(x, y)) return false;
    }
    Fr mu = tr.squeeze_challenge();
    for (int i = 0; i < A; ++i) claims[act[i]] = evals[2 * i] + mu * (evals[2 * i + 1] - evals[2 * i]);
    x.push_back(mu);
    y = x;
    out->points[k + 1] = y;
  }
  out->claims = claims;
  out->point = y;
  return true;
}

inline void lasso_absorb_statement(const LassoTable& tb, int mu, Transcript& tr) {
  tr.common_field_element(Fr::from_u64((uint64_t)tb.kind));
  tr.common_field_element(Fr::from_u64((uint64_t)tb.chunks));
  tr.common_field_element(Fr::from_u64((uint64_t)mu));
  if (tb.kind == TABLE_CUSTOM) {
    tr.common_field_element(Fr::from_u64((uint64_t)tb.num_operands));
    tr.common_field_element(Fr::from_u64((uint64_t)tb.operand_bits));
    tr.common_field_element(Fr::from_u64((uint64_t)tb.custom_out_bits));
    tr.common_field_element(tb.digest());
  }
}

// The STATEMENT of a Lasso proof is "the committed polynomial a holds table values at the committed addresses dim_t":
// ACCEPT alone only says that SOME committed a decomposes into table entries. A caller that cares which lookups were
// proven must bind the proof to its own commitments: `expect_a` (the outputs) and / or `expect_dims` (c commitments to
// the chunked operands) are compared with what the proof carries, and `out_comms` (1 + 4c points: a | dim | E |
// read_ts | final_cts) returns the proof's commitments for linking to an outer protocol.
struct LassoStatement {
  const G1Affine* expect_a = nullptr;
  const G1Affine* expect_dims = nullptr;
  G1Affine* out_comms = nullptr;
};

inline bool lasso_verify(const KzgVerifierParam& vp, const LassoTable& tb, int mu, Transcript& tr,
                         const LassoStatement& stm = LassoStatement()) {
  const int c = tb.chunks;
  lasso_absorb_statement(tb, mu, tr);
  std::vector<G1Affine> mcomms(1 + 3 * c), scomms(c);
  for (auto& p : mcomms)
    if (!tr.read_commitment(&p)) return false;
  for (auto& p : scomms)
    if (!tr.read_commitment(&p)) return false;
  if (stm.out_comms) {
    for (int i = 0; i < 1 + 3 * c; ++i) stm.out_comms[i] = mcomms[i];
    for (int i = 0; i < c; ++i) stm.out_comms[1 + 3 * c + i] = scomms[i];
  }
  if (stm.expect_a && !(*stm.expect_a == mcomms[0])) return false;
  if (stm.expect_dims)
    for (int t = 0; t < c; ++t)
      if (!(stm.expect_dims[t] == mcomms[1 + t])) return false;
  std::vector<Fr> r = tr.squeeze_challenges(mu);
  Fr v_a;
  if (!tr.read_field_element(&v_a)) return false;
  Fr fin;
  std::vector<Fr> x_p;
  if (!sumcheck_verify(mu, 2, v_a, false, tr, &fin, &x_p)) return false;
  std::vector<Fr> e_p(c);
  for (auto& e : e_p)
    if (!tr.read_field_element(&e)) return false;
  Fr g = Fr::zero();
  for (int t = 0; t < c; ++t) g = g + Fr::from_u64((uint64_t)1 << (tb.out_bits() * t)) * e_p[t];
  if (fin != g * eq_xy_eval(x_p, r)) return false;

  Fr gamma = tr.squeeze_challenge(), tau = tr.squeeze_challenge();
  Fr gamma2 = gamma.sqr();
  std::vector<Fr> roots;
  GrandProductOutput gp, gm, gs;
  std::vector<int> hs(4 * c, mu);
  for (int t = 2 * c; t < 4 * c; ++t) hs[t] = SUBTABLE_VARS;
  if (!grand_product_verify(hs, tr, &roots, &gp)) return false;
  std::vector<Fr> mroots(roots.begin(), roots.begin() + 2 * c), sroots(roots.begin() + 2 * c, roots.end());
  gm.point = gp.points[mu];
  gs.point = gp.points[SUBTABLE_VARS];
  gm.claims.assign(gp.claims.begin(), gp.claims.begin() + 2 * c);
  gs.claims.assign(gp.claims.begin() + 2 * c, gp.claims.end());
  // multiset equality  Init * Write == Read * Final  per memory
  for (int t = 0; t < c; ++t)
    if (sroots[2 * t] * mroots[2 * t + 1] != mroots[2 * t] * sroots[2 * t + 1]) return false;

  std::vector<Fr> ev_dim(c), ev_e(c), ev_ts(c), ev_cts(c);
  for (auto* v : {&ev_dim, &ev_e, &ev_ts, &ev_cts})
    for (auto& e : *v)
      if (!tr.read_field_element(&e)) return false;
  Fr id_s = identity_eval(gs.point), t_s = tb.subtable_mle(gs.point);
  for (int t = 0; t < c; ++t) {
    Fr rd = ev_dim[t] * gamma2 + ev_e[t] * gamma + ev_ts[t] - tau;
    if (gm.claims[2 * t] != rd || gm.claims[2 * t + 1] != rd + Fr::one()) return false;
    Fr in = id_s * gamma2 + t_s * gamma - tau;
    if (gs.claims[2 * t] != in || gs.claims[2 * t + 1] != in + ev_cts[t]) return false;
  }
  std::vector<std::vector<Fr>> pts = {r, x_p, gm.point};
  std::vector<Evaluation> evs;
  evs.push_back(Evaluation{0, 0, v_a});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + c + t, 1, e_p[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + t, 2, ev_dim[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + c + t, 2, ev_e[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + 2 * c + t, 2, ev_ts[t]});
  if (!kzg_batch_verify(vp, mu, mcomms, pts, evs, tr)) return false;
  std::vector<Evaluation> sevs;
  for (int t = 0; t < c; ++t) sevs.push_back(Evaluation{t, 0, ev_cts[t]});
  if (!kzg_batch_verify(vp, SUBTABLE_VARS, scomms, {gs.point}, sevs, tr)) return false;
  return true;  // whether bytes may follow is the caller's decision (b200v_transcript_done)
}

// ---- expressions and the boolean hypercube ------------------------------------------------------------------
// bh.rs:5-74 (primitive polynomials over GF(2), index = num_vars) and x^-1 constants
static const uint32_t BH_PRIMITIVES[32] = {
    1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771, 65581, 131081, 262183, 524327,
    1048585, 2097157, 4194307, 8388641, 16777243, 33554441, 67108935, 134217767, 268435465, 536870917,
    1073741907, 2147483657u};

struct BooleanHypercube {
  int num_vars;
  uint64_t primitive, x_inv;
  explicit BooleanHypercube(int n) : num_vars(n), primitive(BH_PRIMITIVES[n]) {
    x_inv = primitive >> 1;  // bh.rs X_INVS: (primitive - 1) / x  ==  primitive >> 1 for polynomials with constant term 1
  }
  uint64_t next(uint64_t b) const {
    b <<= 1;
    b ^= (b >> num_vars) * primitive;
    return b;
  }
  uint64_t prev(uint64_t b) const { return (b >> 1) ^ ((b & 1) * x_inv); }
  uint64_t rotate(uint64_t b, int rotation) const {
    for (int i = 0; i < rotation; ++i) b = next(b);
    for (int i = 0; i > rotation; --i) b = prev(b);
    return b;
  }
  std::vector<uint64_t> iter() const {  // bh.rs:123-130: 0, 1, next(1), ...
    std::vector<uint64_t> out;
    out.push_back(0);
    uint64_t b = 1;
    while (out.size() < ((size_t)1 << num_vars)) {
      out.push_back(b);
      b = next(b);
    }
    return out;
  }
};

struct Expr {
  enum Kind { CONST, IDENTITY, LAGRANGE, EQXY, POLY, CHALLENGE, NEG, SUM, PROD, SCALED, DPOW } kind;
  Fr scalar;      // CONST, SCALED
  int a = 0, b = 0;  // LAGRANGE i / EQXY idx / CHALLENGE idx / POLY (poly, rotation)
  std::vector<std::shared_ptr<Expr>> ch;  // children; DPOW: exprs..., base last
};
typedef std::shared_ptr<Expr> ExprP;

// prefix token stream (include/b200_lasso.h) -> tree; nullptr for a malformed stream
inline ExprP parse_expr(const int32_t*& t, const int32_t* end, const Fr* consts, int nconsts, int depth = 0) {
  if (t >= end || depth > 4096) return nullptr;
  ExprP e = std::make_shared<Expr>();
  const int k = *t++;
  if (k < 0 || k > (int)Expr::DPOW) return nullptr;
  e->kind = (Expr::Kind)k;
  auto need = [&](int n) { return end - t >= n; };
  auto child = [&]() {
    ExprP c = parse_expr(t, end, consts, nconsts, depth + 1);
    if (c) e->ch.push_back(c);
    return (bool)c;
  };
  switch (k) {
    case Expr::CONST:
    case Expr::SCALED: {
      if (!need(1)) return nullptr;
      const int idx = *t++;
      if (idx < 0 || idx >= nconsts) return nullptr;
      e->scalar = consts[idx];
      if (k == Expr::SCALED && !child()) return nullptr;
      break;
    }
    case Expr::IDENTITY: break;
    case Expr::LAGRANGE:
    case Expr::EQXY:
    case Expr::CHALLENGE:
      if (!need(1)) return nullptr;
      e->a = *t++;
      break;
    case Expr::POLY:
      if (!need(2)) return nullptr;
      e->a = *t++;
      e->b = *t++;
      break;
    case Expr::NEG:
      if (!child()) return nullptr;
      break;
    case Expr::SUM:
    case Expr::PROD:
      if (!child() || !child()) return nullptr;
      break;
    case Expr::DPOW: {
      if (!need(1)) return nullptr;
      const int n = *t++;
      if (n < 1 || n > 1 << 20) return nullptr;
      for (int i = 0; i <= n; ++i)
        if (!child()) return nullptr;
      break;
    }
  }
  return e;
}

inline int expr_degree(const ExprP& e) {  // expression.rs:171-182
  switch (e->kind) {
    case Expr::CONST: case Expr::CHALLENGE: return 0;
    case Expr::IDENTITY: case Expr::LAGRANGE: case Expr::EQXY: case Expr::POLY: return 1;
    case Expr::NEG: case Expr::SCALED: return expr_degree(e->ch[0]);
    case Expr::SUM: return std::max(expr_degree(e->ch[0]), expr_degree(e->ch[1]));
    case Expr::PROD: return expr_degree(e->ch[0]) + expr_degree(e->ch[1]);
    case Expr::DPOW: {
      int d = 0;
      for (size_t i = 0; i + 1 < e->ch.size(); ++i) d = std::max(d, expr_degree(e->ch[i]));
      return d + expr_degree(e->ch.back());  // sum(acc, product(scalar, expr)) with a degree-0 base
    }
  }
  return 0;
}

struct LeafValues {
  Fr identity;
  std::map<int, Fr> lagrange;
  std::vector<Fr> eq;
  std::map<std::pair<int, int>, Fr> poly;  // (poly, rotation)
  const Fr* challenges;
};

inline Fr expr_eval(const ExprP& e, const LeafValues& lv) {  // expression.rs:109-169
  switch (e->kind) {
    case Expr::CONST: return e->scalar;
    case Expr::IDENTITY: return lv.identity;
    case Expr::LAGRANGE: return lv.lagrange.at(e->a);
    case Expr::EQXY: return lv.eq[e->a];
    case Expr::POLY: return lv.poly.at({e->a, e->b});
    case Expr::CHALLENGE: return lv.challenges[e->a];
    case Expr::NEG: return -expr_eval(e->ch[0], lv);
    case Expr::SUM: return expr_eval(e->ch[0], lv) + expr_eval(e->ch[1], lv);
    case Expr::PROD: return expr_eval(e->ch[0], lv) * expr_eval(e->ch[1], lv);
    case Expr::SCALED: return expr_eval(e->ch[0], lv) * e->scalar;
    case Expr::DPOW: {
      const size_t n = e->ch.size() - 1;
      if (n == 1) return expr_eval(e->ch[0], lv);
      const Fr base = expr_eval(e->ch[n], lv);
      Fr acc = expr_eval(e->ch[0], lv), pw = base;
      for (size_t i = 1; i < n; ++i) {
        acc = acc + pw * expr_eval(e->ch[i], lv);
        pw = pw * base;
      }
      return acc;
    }
  }
  return Fr::zero();
}

inline void expr_collect(const ExprP& e, std::vector<std::pair<int, int>>* queries, std::vector<int>* lagranges) {
  if (e->kind == Expr::POLY) queries->push_back({e->a, e->b});
  if (e->kind == Expr::LAGRANGE) lagranges->push_back(e->a);
  for (auto& c : e->ch) expr_collect(c, queries, lagranges);
}

typedef std::pair<int, int> Query;  // (poly, rotation) — `Query` orders by (poly, rotation), expression.rs:40-44

inline void collect_queries(const ExprP& e, std::set<Query>* out) {
  if (e->kind == Expr::POLY) out->insert({e->a, e->b});
  for (auto& c : e->ch) collect_queries(c, out);
}

// multilinear.rs:519-541
inline std::vector<uint64_t> rotation_eval_point_pattern(bool next, int num_vars, int distance) {
  BooleanHypercube bh(num_vars);
  const uint64_t rem = next ? bh.primitive : bh.x_inv;
  std::vector<uint64_t> pat((size_t)1 << distance, 0);
  for (int depth = 0; depth < distance; ++depth) {
    const size_t step = (size_t)1 << (distance - depth);
    for (size_t e = 0; e < pat.size(); e += step) {
      const size_t o = e + (step >> 1);
      const uint64_t rotated = next ? pat[e] << 1 : pat[e] >> 1;
      pat[o] = rotated ^ rem;
      pat[e] = rotated;
    }
  }
  return pat;
}

// multilinear.rs:543-566
inline std::vector<uint64_t> rotation_eval_coeff_pattern(bool next, int num_vars, int distance) {
  BooleanHypercube bh(num_vars);
  const uint64_t rem = next ? bh.primitive - ((uint64_t)1 << num_vars) : bh.x_inv << distance;
  std::vector<uint64_t> pat((size_t)1 << (distance - 1), 0);
  for (int depth = 0; depth + 1 < distance; ++depth) {
    const size_t step = (size_t)1 << (distance - depth - 1);
    for (size_t e = 0; e < pat.size(); e += step) {
      const size_t o = e + (step >> 1);
      const uint64_t rotated = next ? pat[e] << 1 : pat[e] >> 1;
      pat[o] = rotated ^ rem;
      pat[e] = rotated;
    }
  }
  return pat;
}

// multilinear.rs:475-517
inline std::vector<std::vector<Fr>> rotation_eval_points(const std::vector<Fr>& x, int rotation) {
  if (rotation == 0) return {x};
  const int n = (int)x.size(), distance = std::abs(rotation), num_x = n - distance;
  std::vector<std::vector<Fr>> out;
  if (rotation < 0) {
    for (uint64_t pat : rotation_eval_point_pattern(false, n, distance)) {
      std::vector<Fr> p;
      for (int i = 0; i < num_x; ++i) p.push_back(((pat >> i) & 1) ? Fr::one() - x[distance + i] : x[distance + i]);
      for (int i = 0; i < distance; ++i) p.push_back(((pat >> (i + num_x)) & 1) ? Fr::one() : Fr::zero());
      out.push_back(p);
    }
  } else {
    for (uint64_t pat : rotation_eval_point_pattern(true, n, distance)) {
      std::vector<Fr> p;
      for (int i = 0; i < distance; ++i) p.push_back(((pat >> i) & 1) ? Fr::one() : Fr::zero());
      for (int i = 0; i < num_x; ++i) p.push_back(((pat >> (i + distance)) & 1) ? Fr::one() - x[i] : x[i]);
      out.push_back(p);
    }
  }
  return out;
}

// multilinear.rs:433-473
inline Fr rotation_eval(const std::vector<Fr>& x, int rotation, const std::vector<Fr>& evals_for_rotation) {
  if (rotation == 0) return evals_for_rotation[0];
  const int n = (int)x.size(), distance = std::abs(rotation);
  std::vector<uint64_t> pattern;
  std::vector<int> nths;
  std::vector<Fr> xs;
  if (rotation < 0) {
    pattern = rotation_eval_coeff_pattern(false, n, distance);
    for (int i = distance; i >= 1; --i) nths.push_back(i);
    for (int i = distance - 1; i >= 0; --i) xs.push_back(x[i]);
  } else {
    pattern = rotation_eval_coeff_pattern(true, n, distance);
    for (int i = 0; i < distance; ++i) nths.push_back(n - 1 + i);
    for (int i = n - distance; i < n; ++i) xs.push_back(x[i]);
  }
  std::vector<Fr> evals = evals_for_rotation;
  for (int idx = 0; idx < distance; ++idx) {
    std::vector<Fr> next;
    for (size_t k = 0; 2 * k + 1 < evals.size(); ++k) {
      const uint64_t pat = pattern[k << idx];
      const bool bit = (pat >> nths[idx]) & 1;
      const Fr &e0 = evals[2 * k], &e1 = evals[2 * k + 1];
      next.push_back(bit ? (e0 - e1) * xs[idx] + e1 : (e1 - e0) * xs[idx] + e0);
    }
    evals.swap(next);
  }
  return evals[0];
}

inline Fr lagrange_eval(const std::vector<Fr>& x, uint64_t b) {  // sum_check.rs:97-109
  Fr acc = Fr::one();
  for (size_t i = 0; i < x.size(); ++i) acc = acc * (((b >> i) & 1) ? x[i] : Fr::one() - x[i]);
  return acc;
}

struct PcsQueryPlan {
  std::vector<Query> queries;        // pcs_query (BTreeSet order)
  std::vector<int> rotations;        // distinct rotations, sorted
  std::map<int, int> point_offset;   // verifier.rs:164-182
};

inline PcsQueryPlan pcs_query_plan(const ExprP& e, int num_instance_poly) {
  PcsQueryPlan pl;
  std::set<Query> qs;
  collect_queries(e, &qs);
  std::set<int> rots;
  for (auto& q : qs)
    if (q.first >= num_instance_poly) {
      pl.queries.push_back(q);
      rots.insert(q.second);
    }
  pl.rotations.assign(rots.begin(), rots.end());
  int off = 0;
  for (int r : pl.rotations) {
    pl.point_offset[r] = off;
    off += 1 << std::abs(r);
  }
  return pl;
}

// HyperPlonkVerifierParam (hyperplonk.rs:58-74)
struct HyperPlonkVerifierParam {
  KzgVerifierParam kzg;
  int num_vars = 0;
  std::vector<int> num_instances;        // per instance column
  std::vector<int> phase_witness_polys;  // per phase
  std::vector<int> phase_challenges;
  int num_lookups = 0, num_permutation_z_polys = 0;
  ExprP expression;                      // the composed zero-check expression (preprocessor.rs:25-60)
  std::vector<G1Affine> preprocess_comms, permutation_comms;
};

// hyperplonk.rs:293-363 + verifier.rs:39-145
inline bool hyperplonk_verify(const HyperPlonkVerifierParam& vp, const std::vector<std::vector<Fr>>& instances, Transcript& tr) {
  const int n = vp.num_vars;
  for (auto& inst : instances)
    for (auto& v : inst) tr.common_field_element(v);
  if (instances.size() != vp.num_instances.size()) return false;
  for (size_t i = 0; i < instances.size(); ++i)
    if ((int)instances[i].size() != vp.num_instances[i]) return false;  // hyperplonk.rs:299-305
  // rounds 0..n (hyperplonk.rs:307-315)
  const std::vector<int> phase_w = vp.phase_witness_polys;
  const std::vector<int> phase_c = vp.phase_challenges;
  std::vector<G1Affine> witness_comms;
  std::vector<Fr> challenges;
  for (size_t round = 0; round < phase_w.size(); ++round) {
    for (int i = 0; i < phase_w[round]; ++i) {
      G1Affine c;
      if (!tr.read_commitment(&c)) return false;
      witness_comms.push_back(c);
    }
    for (auto& c : tr.squeeze_challenges(phase_c[round])) challenges.push_back(c);
  }
  const Fr beta = tr.squeeze_challenge();
  std::vector<G1Affine> m_comms((size_t)vp.num_lookups);
  for (auto& c : m_comms)
    if (!tr.read_commitment(&c)) return false;
  const Fr gamma = tr.squeeze_challenge();
  std::vector<G1Affine> z_comms((size_t)vp.num_lookups + vp.num_permutation_z_polys);  // h polys, then z polys
  for (auto& c : z_comms)
    if (!tr.read_commitment(&c)) return false;
  const Fr alpha = tr.squeeze_challenge();
  std::vector<Fr> y = tr.squeeze_challenges(n);
  challenges.insert(challenges.end(), {beta, gamma, alpha});
  const int d = expr_degree(vp.expression);
  Fr x_eval;
  std::vector<Fr> x;
  if (!sumcheck_verify(n, d, Fr::zero(), false, tr, &x_eval, &x)) return false;
  PcsQueryPlan pl = pcs_query_plan(vp.expression, (int)instances.size());
  LeafValues lv;
  lv.challenges = challenges.data();
  std::vector<std::vector<Fr>> evals_for_rotation;
  for (auto& q : pl.queries) {
    std::vector<Fr> ev((size_t)1 << std::abs(q.second));
    for (auto& e : ev)
      if (!tr.read_field_element(&e)) return false;
    lv.poly[q] = rotation_eval(x, q.second, ev);
    evals_for_rotation.push_back(ev);
  }
  // instance_evals (verifier.rs:92-145): Σ_j inst[j] * L_{bh[is_j]}(x) with is = 1 - rot, 2 - rot, ... for rot <= 0
  // and -rot, ..., -1, 1, 2, ... for rot > 0 (row 0 of the LFSR order is skipped)
  std::vector<uint64_t> order = BooleanHypercube(n).iter();
  std::set<Query> qs;
  collect_queries(vp.expression, &qs);
  for (auto& q : qs)
    if (q.first < (int)instances.size()) {
      const long Nrows = 1L << n;
      Fr acc = Fr::zero();
      long i = q.second > 0 ? -(long)q.second : 1 - (long)q.second;
      for (size_t j = 0; j < instances[q.first].size(); ++j, ++i) {
        if (q.second > 0 && i == 0) i = 1;
        acc = acc + instances[q.first][j] * lagrange_eval(x, order[((i % Nrows) + Nrows) % Nrows]);
      }
      lv.poly[q] = acc;
    }
  // evaluate (sum_check.rs:60-95)
  lv.identity = identity_eval(x);
  std::vector<std::pair<int, int>> dummy;
  std::vector<int> lag_ids;
  expr_collect(vp.expression, &dummy, &lag_ids);
  const long N = 1L << n;
  for (int i : lag_ids) lv.lagrange[i] = lagrange_eval(x, order[((i % N) + N) % N]);
  lv.eq.push_back(eq_xy_eval(x, y));
  if (expr_eval(vp.expression, lv) != x_eval) return false;
  std::vector<std::vector<Fr>> points;
  for (int r : pl.rotations)
    for (auto& p : rotation_eval_points(x, r)) points.push_back(p);
  std::vector<Evaluation> evals;
  for (size_t k = 0; k < pl.queries.size(); ++k) {
    int pt = pl.point_offset[pl.queries[k].second];
    for (auto& e : evals_for_rotation[k]) evals.push_back(Evaluation{pl.queries[k].first, pt++, e});
  }
  std::vector<G1Affine> comms(instances.size(), G1Affine::identity());
  comms.insert(comms.end(), vp.preprocess_comms.begin(), vp.preprocess_comms.end());
  comms.insert(comms.end(), witness_comms.begin(), witness_comms.end());
  comms.insert(comms.end(), vp.permutation_comms.begin(), vp.permutation_comms.end());
  comms.insert(comms.end(), m_comms.begin(), m_comms.end());
  comms.insert(comms.end(), z_comms.begin(), z_comms.end());
  if (!kzg_batch_verify(vp.kzg, n, comms, points, evals, tr)) return false;
  return true;  // whether bytes may follow is the caller's decision (b200v_transcript_done)
}

}  // namespace b200v

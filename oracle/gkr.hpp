// ORACLE (test infrastructure only — never linked into the product path).
//
// GKR for fractional sum-checks, restated statement by statement from
//   pb/piop/gkr/fractional_sum_check.rs
//     Layer::bottom / Layer::up                :41-85   (halves of a table; p = p_l q_r + p_r q_l, q = q_l q_r)
//     prove_fractional_sum_check               :87-190
//     verify_fractional_sum_check              :192-265
//     sum_check_expression / sum_check_claim   :267-288 (distribute_powers: Σ_i expr_i γ^i, pb/util/expression.rs:155-167)
//     layer_down_claim                         :290-296
// The per-layer sum-check is ClassicSumCheck<EvaluationsProver> over the expression
//   eq(x, y) * Σ_b [ γ^(2b) (p_l q_r + p_r q_l) + γ^(2b+1) q_l q_r ],
// i.e. the EVAL shape of sumcheck.hpp with three terms per batch element that SHARE the four tables of that element;
// its final evaluations come back in poly order [p_l, p_r, q_l, q_r] per element (classic.rs:143-149).
//
// parity: UNPINNED against the Rust crate (it cannot be built here and holds no numeric vectors for this module; its
// own test is a prove -> verify property test, restated in tests/test_oracle_gkr.py). Pinned instead by an independent
// pure-Python prover written from the same source (tests/golden/pymodel_gkr.py -> gkr_golden.json, byte for byte).
#pragma once
#include <vector>

#include "ff.hpp"
#include "mle.hpp"
#include "sumcheck.hpp"
#include "transcript.hpp"

namespace oracle {

struct FractionalOutput {
  std::vector<Fr> p_xs, q_xs;  // p_b(x), q_b(x) per batch element
  std::vector<Fr> x;           // num_vars coordinates
  std::vector<Fr> p_0s, q_0s;  // the layer-0 values Σ p/q = p_0 / q_0 (written or absorbed, :121-146)
};

// claimed_*: pointer to a value = Some(claimed) (absorbed with common_field_element), nullptr = None (written)
inline FractionalOutput fractional_sum_check_prove(const std::vector<const Fr*>& claimed_p_0s,
                                                   const std::vector<const Fr*>& claimed_q_0s,
                                                   const std::vector<const Poly*>& ps, const std::vector<const Poly*>& qs,
                                                   Transcript& tr) {
  const int B = (int)ps.size();
  const int n = log2_exact(ps[0]->size());
  // tabs[b][v]: (p, q) tables with 2^v entries; v = n is the input; Layer with num_vars v = halves of level v + 1
  std::vector<std::vector<Poly>> P(B), Q(B);
  for (int b = 0; b < B; ++b) {
    P[b].resize(n + 1);
    Q[b].resize(n + 1);
    P[b][n] = *ps[b];
    Q[b][n] = *qs[b];
    for (int v = n - 1; v >= 0; --v) {
      const long half = 1L << v;
      const Poly &pc = P[b][v + 1], &qc = Q[b][v + 1];
      Poly pu(half), qu(half);
#pragma omp parallel for if (half >= 4096)
      for (long i = 0; i < half; ++i) {
        pu[i] = pc[i] * qc[i + half] + pc[i + half] * qc[i];  // :79
        qu[i] = qc[i] * qc[i + half];                        // :80
      }
      P[b][v] = pu;
      Q[b][v] = qu;
    }
  }
  FractionalOutput out;
  auto hash = [&](const std::vector<const Fr*>& claimed, const std::vector<Fr>& computed) {
    for (size_t i = 0; i < computed.size(); ++i) {
      if (claimed[i]) tr.common_field_element(computed[i]);
      else tr.write_field_elements(&computed[i], 1);
    }
  };
  for (int b = 0; b < B; ++b) out.p_0s.push_back(P[b][0][0]);
  for (int b = 0; b < B; ++b) out.q_0s.push_back(Q[b][0][0]);
  hash(claimed_p_0s, out.p_0s);
  hash(claimed_q_0s, out.q_0s);

  std::vector<Fr> cp = out.p_0s, cq = out.q_0s, y;
  for (int v = 0; v < n; ++v) {  // the layer with num_vars = v
    const size_t half = (size_t)1 << v;
    std::vector<Poly> tabs(4 * B);
    for (int b = 0; b < B; ++b) {
      const Poly &pc = P[b][v + 1], &qc = Q[b][v + 1];
      tabs[4 * b + 0].assign(pc.begin(), pc.begin() + half);
      tabs[4 * b + 1].assign(pc.begin() + half, pc.end());
      tabs[4 * b + 2].assign(qc.begin(), qc.begin() + half);
      tabs[4 * b + 3].assign(qc.begin() + half, qc.end());
    }
    std::vector<Fr> x, evals;
    if (v == 0) {
      for (auto& t : tabs) evals.push_back(t[0]);
    } else {
      const Fr gamma = tr.squeeze_challenge();
      VirtualPoly vp;
      vp.has_eq = true;
      vp.y = y;
      for (auto& t : tabs) vp.polys.push_back(&t);
      Fr pw = Fr::one(), claim = Fr::zero();
      for (int b = 0; b < B; ++b) {
        vp.terms.push_back(Term{pw, {4 * b + 0, 4 * b + 3}});  // p_l q_r
        vp.terms.push_back(Term{pw, {4 * b + 1, 4 * b + 2}});  // p_r q_l
        claim = claim + pw * cp[b];
        pw = pw * gamma;
        vp.terms.push_back(Term{pw, {4 * b + 2, 4 * b + 3}});  // q_l q_r
        claim = claim + pw * cq[b];
        pw = pw * gamma;
      }
      SumCheckOutput sc = sumcheck_prove_evals(v, vp, claim, tr);
      x = sc.challenges;
      evals = sc.evals;
    }
    tr.write_field_elements(evals.data(), evals.size());
    const Fr mu = tr.squeeze_challenge();
    for (int b = 0; b < B; ++b) {  // layer_down_claim :290-296
      cp[b] = evals[4 * b] + mu * (evals[4 * b + 1] - evals[4 * b]);
      cq[b] = evals[4 * b + 2] + mu * (evals[4 * b + 3] - evals[4 * b + 2]);
    }
    x.push_back(mu);
    y = x;
  }
  out.p_xs = cp;
  out.q_xs = cq;
  out.x = y;
  return out;
}

// :192-265. claimed_*: as above (Some -> absorbed, None -> read from the proof).
inline bool fractional_sum_check_verify(int num_vars, const std::vector<const Fr*>& claimed_p_0s,
                                        const std::vector<const Fr*>& claimed_q_0s, Transcript& tr, FractionalOutput* out) {
  const int B = (int)claimed_p_0s.size();
  std::vector<Fr> cp(B), cq(B);
  auto take = [&](const std::vector<const Fr*>& claimed, std::vector<Fr>& dst) {
    for (int b = 0; b < B; ++b) {
      if (claimed[b]) {
        dst[b] = *claimed[b];
        tr.common_field_element(dst[b]);
      } else if (!tr.read_field_element(&dst[b])) {
        return false;
      }
    }
    return true;
  };
  if (!take(claimed_p_0s, cp) || !take(claimed_q_0s, cq)) return false;
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

}  // namespace oracle

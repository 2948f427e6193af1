// ORACLE (test infrastructure only — never linked into the product path).
//
// ClassicSumCheck restated:
//   driver loop            pb/piop/sum_check/classic.rs:208-263
//   ProverState / bind     classic.rs:41-150
//   EvaluationsProver      classic/eval.rs:21-57, 92-131, 210-323 (message = p(0..d); p(0) DERIVED
//                          as sum - p(1); per pair: eval = t[2b+1], step = t[2b+1]-t[2b], then += step)
//   CoefficientsProver     classic/coeff.rs:19-39, 136-203 (degree 2; c1 derived)
//   barycentric            pb/util/arithmetic.rs:108-136
// The expression is restricted to the shapes the Lasso/PCS hot path uses (SURVEY §8c allows fixed
// shapes as long as the derived evaluations match):
//   EVAL shape :  F(x) = [eq(x, y)] * Σ_t coeff_t * Π_{i in polys_t} P_i(x)
//   COEFF shape:  F(x) = Σ_i scalar_i * eq(x, y_i) * P_{poly_i}(x)
// Evaluations of a polynomial at 0..d are unique field elements, so the generic DAG evaluator of
// pb/util/expression/evaluator.rs yields byte-identical messages for these expressions.
#pragma once
#include <vector>

#include "ff.hpp"
#include "mle.hpp"
#include "transcript.hpp"

namespace oracle {

struct Term {
  Fr coeff;
  std::vector<int> polys;
};

struct VirtualPoly {
  bool has_eq = false;
  std::vector<Fr> y;  // eq point when has_eq
  std::vector<Term> terms;
  std::vector<const Poly*> polys;
  int degree() const {
    size_t d = 0;
    for (auto& t : terms) d = t.polys.size() > d ? t.polys.size() : d;
    return (int)d + (has_eq ? 1 : 0);
  }
};

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

struct SumCheckOutput {
  std::vector<Fr> challenges;  // x
  std::vector<Fr> evals;       // every input table bound at x (classic.rs:143-149)
};

// ClassicSumCheck::<EvaluationsProver>::prove for the EVAL shape.
inline SumCheckOutput sumcheck_prove_evals(int num_vars, const VirtualPoly& vp, Fr sum,
                                           Transcript& tr) {
  const int d = vp.degree();
  const int np = (int)vp.polys.size();
  std::vector<Poly> tabs(np);
  for (int i = 0; i < np; ++i) tabs[i] = *vp.polys[i];
  Poly eq;
  if (vp.has_eq) eq = eq_xy(vp.y);
  const std::vector<Fr> points = points_0_to_d(d);
  const std::vector<Fr> weights = barycentric_weights(points);

  SumCheckOutput out;
  for (int round = 0; round < num_vars; ++round) {
    const long size = 1L << (num_vars - round - 1);
    std::vector<Fr> evals(d + 1, Fr::zero());
#pragma omp parallel if (size >= 1024)
    {
      std::vector<Fr> acc(d + 1, Fr::zero());
      std::vector<Fr> ev(np), st(np);
#pragma omp for schedule(static)
      for (long b = 0; b < size; ++b) {
        Fr eq_ev = Fr::one(), eq_st = Fr::zero();
        if (vp.has_eq) {
          eq_ev = eq[2 * b + 1];
          eq_st = eq[2 * b + 1] - eq[2 * b];
        }
        for (int i = 0; i < np; ++i) {
          ev[i] = tabs[i][2 * b + 1];
          st[i] = tabs[i][2 * b + 1] - tabs[i][2 * b];
        }
        for (int x = 1; x <= d; ++x) {
          if (x > 1) {
            eq_ev = eq_ev + eq_st;
            for (int i = 0; i < np; ++i) ev[i] = ev[i] + st[i];
          }
          Fr s = Fr::zero();
          for (auto& t : vp.terms) {
            Fr p = t.coeff;
            bool first = true;
            for (int i : t.polys) {
              p = (first && t.coeff == Fr::one()) ? ev[i] : p * ev[i];
              first = false;
            }
            s = s + p;
          }
          if (vp.has_eq) s = s * eq_ev;
          acc[x] = acc[x] + s;
        }
      }
#pragma omp critical
      for (int x = 1; x <= d; ++x) evals[x] = evals[x] + acc[x];
    }
    evals[0] = sum - evals[1];  // eval.rs:129
    tr.write_field_elements(evals.data(), evals.size());
    Fr r = tr.squeeze_challenge();
    out.challenges.push_back(r);
    sum = barycentric_interpolate(weights, points, evals, r);
    if (vp.has_eq) fix_var_in_place(eq, r);
    for (int i = 0; i < np; ++i) fix_var_in_place(tabs[i], r);
  }
  for (int i = 0; i < np; ++i) out.evals.push_back(tabs[i][0]);
  return out;
}

struct CoeffProduct {
  Fr scalar;
  std::vector<Fr> y;  // eq point
  int poly;
};

// ClassicSumCheck::<CoefficientsProver>::prove for Σ scalar_i * eq(x,y_i) * P_i(x) (+ constant 0).
inline SumCheckOutput sumcheck_prove_coeffs(int num_vars, const std::vector<CoeffProduct>& prods,
                                            const std::vector<const Poly*>& polys, Fr sum,
                                            Transcript& tr) {
  const int np = (int)polys.size();
  std::vector<Poly> tabs(np);
  for (int i = 0; i < np; ++i) tabs[i] = *polys[i];
  std::vector<Poly> eqs(prods.size());
  for (size_t k = 0; k < prods.size(); ++k) eqs[k] = eq_xy(prods[k].y);

  SumCheckOutput out;
  for (int round = 0; round < num_vars; ++round) {
    const long size = 1L << (num_vars - round - 1);
    Fr c[3] = {Fr::zero(), Fr::zero(), Fr::zero()};
    for (size_t k = 0; k < prods.size(); ++k) {
      const Poly& l = eqs[k];
      const Poly& rp = tabs[prods[k].poly];
      Fr c0 = Fr::zero(), c2 = Fr::zero();
#pragma omp parallel if (size >= 1024)
      {
        Fr a0 = Fr::zero(), a2 = Fr::zero();
#pragma omp for schedule(static)
        for (long b = 0; b < size; ++b) {
          a0 = a0 + l[2 * b] * rp[2 * b];
          a2 = a2 + (l[2 * b + 1] - l[2 * b]) * (rp[2 * b + 1] - rp[2 * b]);
        }
#pragma omp critical
        {
          c0 = c0 + a0;
          c2 = c2 + a2;
        }
      }
      // coeff.rs:49-62: scalar == 1 adds as is, otherwise scaled
      c[0] = c[0] + prods[k].scalar * c0;
      c[2] = c[2] + prods[k].scalar * c2;
    }
    c[1] = sum - (c[0].dbl() + c[2]);  // coeff.rs:147 with coeffs[1] == 0 inside sum()
    tr.write_field_elements(c, 3);
    Fr r = tr.squeeze_challenge();
    out.challenges.push_back(r);
    sum = (c[2] * r + c[1]) * r + c[0];  // horner, coeff.rs:36-38
    for (auto& e : eqs) fix_var_in_place(e, r);
    for (auto& t : tabs) fix_var_in_place(t, r);
  }
  for (int i = 0; i < np; ++i) out.evals.push_back(tabs[i][0]);
  return out;
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

}  // namespace oracle

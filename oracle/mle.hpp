// ORACLE (test infrastructure only — never linked into the product path).
//
// MultilinearPolynomial pieces of pb/poly/multilinear.rs used on the hot path:
//   eq_xy          :91-127   (doubling; y[0] lands on bit 0 because y is walked in reverse)
//   evaluate       :137-156  (successive LSB-first binds; the boolean short-cuts of the
//                             reference only skip work, the value is the same field element)
//   fix_var(_in_place) / merge_into :179-189, 599-618   out[b] = (t[2b+1]-t[2b])*r + t[2b]
//   fix_last_vars  :158-177
// plus the verifier-side closed forms of pb/piop/sum_check.rs:97-125.
#pragma once
#include <vector>

#include "ff.hpp"

namespace oracle {

typedef std::vector<Fr> Poly;

inline int log2_exact(size_t n) {
  int k = 0;
  while (((size_t)1 << k) < n) ++k;
  return k;
}

inline Poly eq_xy(const std::vector<Fr>& y) {
  Poly evals(1, Fr::one());
  for (size_t k = y.size(); k-- > 0;) {
    Poly next(2 * evals.size());
    const Fr yk = y[k];
    const long n = (long)evals.size();
#pragma omp parallel for if (n >= 4096)
    for (long i = 0; i < n; ++i) {
      next[2 * i + 1] = evals[i] * yk;
      next[2 * i] = evals[i] - next[2 * i + 1];
    }
    evals.swap(next);
  }
  return evals;
}

inline Poly fix_var(const Poly& p, const Fr& r) {
  Poly out(p.size() / 2);
  const long n = (long)out.size();
#pragma omp parallel for if (n >= 4096)
  for (long b = 0; b < n; ++b) out[b] = (p[2 * b + 1] - p[2 * b]) * r + p[2 * b];
  return out;
}

inline void fix_var_in_place(Poly& p, const Fr& r) {
  Poly out = fix_var(p, r);
  p.swap(out);
}

inline Fr evaluate(const Poly& p, const std::vector<Fr>& x) {
  Poly cur = p;
  for (size_t i = 0; i < x.size(); ++i) fix_var_in_place(cur, x[i]);
  return cur[0];
}

// multilinear.rs:158-177: binds the TOP |x| variables (x.last() is the highest one)
inline Poly fix_last_vars(const Poly& p, const std::vector<Fr>& x) {
  Poly out = p;
  size_t len = p.size();
  for (size_t k = x.size(); k-- > 0;) {
    len >>= 1;
    for (size_t i = 0; i < len; ++i) out[i] = out[i] + (out[i + len] - out[i]) * x[k];
  }
  out.resize(len);
  return out;
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

}  // namespace oracle

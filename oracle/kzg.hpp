// ORACLE (test infrastructure only — never linked into the product path).
//
// Pippenger MSM      pb/util/arithmetic/msm.rs:8-14 (window_size), :33-48 (windowed_scalar),
//                    :84-115 (chunk per thread, then sum), :117-181 (serial bucket method)
// MultilinearKzg     pb/pcs/multilinear/kzg.rs:166-228 (setup), :252-274 (commit),
//                    :276-302 (open), :330-361 (verify)
// quotients          pb/pcs/multilinear.rs:72-107
// additive batch     pb/pcs/multilinear.rs:134-275 (batch_open / batch_verify)
//
// Deviations, stated: (1) SRS scalars `ss` come from a caller-supplied seed list instead of
// `F::random(StdRng)` (ChaCha is not restated; SURVEY §7 hard part g). (2) `verify` has two forms: the
// reference's pairing product e(C - v*g1, -g2) * Π e(Q_i, (s_i - x_i) g2) == 1 (pairing.hpp, SURVEY §8(f) N2;
// enabled with KzgParams::pairing_check) and, by default because it is ~100x faster, the same equation in the
// exponent with the trapdoor `ss`:  C - v*g1 == Σ (s_i - x_i) Q_i.
#pragma once
#include <cmath>
#include <vector>

#include <omp.h>

#include "ff.hpp"
#include "g1.hpp"
#include "pairing.hpp"
#include "mle.hpp"
#include "sumcheck.hpp"
#include "transcript.hpp"

namespace oracle {

inline int msm_window_size(size_t n) { return n < 32 ? 3 : (int)std::floor(std::log((double)n)); }

inline size_t windowed_scalar(int window, size_t mask, int idx, const uint8_t repr[32]) {
  size_t skip_bits = (size_t)idx * window, skip_bytes = skip_bits / 8;
  uint8_t v[8] = {0};
  for (size_t i = 0; i < 8 && skip_bytes + i < 32; ++i) v[i] = repr[skip_bytes + i];
  uint64_t w;
  memcpy(&w, v, 8);
  return (size_t)(w >> (skip_bits - skip_bytes * 8)) & mask;
}

// msm.rs:117-181; scalars are canonical LE reprs
inline G1 msm_serial(const uint8_t (*reprs)[32], const G1Affine* bases, size_t n) {
  G1 result = G1::identity();
  if (n == 0) return result;
  const int c = msm_window_size(n);
  const size_t num_buckets = ((size_t)1 << c) - 1;
  const int num_windows = (256 + c - 1) / c;
  std::vector<G1> buckets(num_buckets);
  std::vector<uint8_t> used(num_buckets);
  for (int idx = num_windows - 1; idx >= 0; --idx) {
    for (int k = 0; k < c; ++k) result = result.dbl();
    std::fill(used.begin(), used.end(), 0);
    for (size_t i = 0; i < n; ++i) {
      size_t s = windowed_scalar(c, num_buckets, idx, reprs[i]);
      if (s != 0) {
        if (!used[s - 1]) {
          buckets[s - 1] = G1::from_affine(bases[i]);
          used[s - 1] = 1;
        } else {
          buckets[s - 1] = buckets[s - 1].add_affine(bases[i]);
        }
      }
    }
    G1 running = G1::identity();
    for (size_t b = num_buckets; b-- > 0;) {
      if (used[b]) running = running.add(buckets[b]);
      result = result.add(running);
    }
  }
  return result;
}

// msm.rs:84-115: contiguous chunk per thread, each a full serial Pippenger, results summed
inline G1 variable_base_msm(const Fr* scalars, const G1Affine* bases, size_t n) {
  if (n == 0) return G1::identity();
  std::vector<uint8_t> reprs(n * 32);
  uint8_t(*rp)[32] = (uint8_t(*)[32])reprs.data();
  const long ln = (long)n;
#pragma omp parallel for if (ln >= 1024)
  for (long i = 0; i < ln; ++i) scalars[i].to_repr(rp[i]);
  const size_t threads = (size_t)omp_get_max_threads();
  if (n <= threads) return msm_serial(rp, bases, n);
  const size_t chunk = (n + threads - 1) / threads;
  const size_t nchunks = (n + chunk - 1) / chunk;
  std::vector<G1> results(nchunks, G1::identity());
#pragma omp parallel for schedule(static, 1)
  for (long t = 0; t < (long)nchunks; ++t) {
    size_t lo = t * chunk, hi = lo + chunk < n ? lo + chunk : n;
    results[t] = msm_serial(rp + lo, bases + lo, hi - lo);
  }
  G1 acc = G1::identity();
  for (auto& r : results) acc = acc.add(r);
  return acc;
}

struct KzgParams {
  int num_vars = 0;
  std::vector<Fr> ss;                      // trapdoor (test SRS; also used by the fast form of `verify`)
  std::vector<std::vector<G1Affine>> eqs;  // eqs[k]: 2^k points, eqs[k][b] = g1 * Π_j (b_j ? s_j : 1-s_j)
  // verifier side (kzg.rs:215-225, 240-249): ss_g2[i] = g2 * s_i, built on first use
  mutable std::vector<G2Affine> ss_g2;
  bool pairing_check = false;  // verify with the pairing product (as the reference) instead of the trapdoor identity
  const std::vector<G2Affine>& g2_powers() const {
    if (ss_g2.size() != ss.size()) {
      ss_g2.clear();
      for (const Fr& s : ss) ss_g2.push_back(G2Affine::generator().mul(s));
    }
    return ss_g2;
  }
};

// kzg.rs:166-213. The newest variable s_i lands on the TOP bit (evals_hi = s_i * last).
inline KzgParams kzg_setup(const std::vector<Fr>& ss) {
  KzgParams p;
  p.num_vars = (int)ss.size();
  p.ss = ss;
  std::vector<Poly> eqs(1, Poly(1, Fr::one()));
  for (const Fr& s : ss) {
    const Poly& last = eqs.back();
    Poly ev(2 * last.size());
    for (size_t i = 0; i < last.size(); ++i) {
      ev[last.size() + i] = s * last[i];
      ev[i] = last[i] - ev[last.size() + i];
    }
    eqs.push_back(ev);
  }
  // fixed_base_msm(g1, scalar) == g1 * scalar; computed with a window table (8-bit digits)
  const G1 g = G1::from_affine(G1Affine::generator());
  std::vector<std::vector<G1Affine>> table(32);  // table[w][d-1] = g * d * 2^{8w}
  {
    G1 base = g;
    for (int w = 0; w < 32; ++w) {
      std::vector<G1> row(255);
      G1 acc = base;
      for (int d = 0; d < 255; ++d) {
        row[d] = acc;
        acc = acc.add(base);
      }
      table[w].resize(255);
      batch_normalize(row.data(), table[w].data(), 255);
      base = acc;  // 256 * base
    }
  }
  p.eqs.resize(eqs.size());
  for (size_t k = 0; k < eqs.size(); ++k) {
    const long n = (long)eqs[k].size();
    std::vector<G1> proj(n);
#pragma omp parallel for if (n >= 64)
    for (long i = 0; i < n; ++i) {
      uint8_t repr[32];
      eqs[k][i].to_repr(repr);
      G1 acc = G1::identity();
      for (int w = 0; w < 32; ++w)
        if (repr[w]) acc = acc.add_affine(table[w][repr[w] - 1]);
      proj[i] = acc;
    }
    p.eqs[k].resize(n);
    batch_normalize(proj.data(), p.eqs[k].data(), n);
  }
  return p;
}

// kzg.rs:252-257
inline G1Affine kzg_commit(const KzgParams& pp, const Poly& poly) {
  int nv = log2_exact(poly.size());
  return variable_base_msm(poly.data(), pp.eqs[nv].data(), poly.size()).to_affine();
}

// multilinear.rs (pcs) :72-107 + kzg.rs:291-299: writes n quotient commitments, returns remainder
inline Fr kzg_open(const KzgParams& pp, const Poly& poly, const std::vector<Fr>& point,
                   Transcript& tr, bool* ok) {
  const int n = (int)point.size();
  Poly rem = poly;
  std::vector<G1Affine> comms(n);
  for (int nv = n - 1; nv >= 0; --nv) {
    const size_t half = (size_t)1 << nv;
    Poly q(half);
    for (size_t i = 0; i < half; ++i) q[i] = rem[half + i] - rem[i];
    for (size_t i = 0; i < half; ++i) rem[i] = rem[i] + (rem[half + i] - rem[i]) * point[nv];
    rem.resize(half);
    comms[nv] = variable_base_msm(q.data(), pp.eqs[nv].data(), half).to_affine();
  }
  *ok = true;
  for (int i = 0; i < n; ++i)
    if (!tr.write_commitment(comms[i])) *ok = false;
  return rem[0];
}

// kzg.rs:330-361. With vp.pairing_check the reference's check itself:
//   e(comm - g1 * eval, -g2) * Π_i e(quotient_i, g2 * s_i - g2 * x_i) == 1     (pairings_product_is_identity);
// otherwise the same equation in the exponent, evaluated with the trapdoor (fast; the SRS of the tests is generated
// from known scalars). Both forms accept / reject the same openings (tests/test_oracle_pairing.py).
inline bool kzg_verify(const KzgParams& vp, const G1Affine& comm, const std::vector<Fr>& point,
                       const Fr& eval, Transcript& tr) {
  const int n = (int)point.size();
  std::vector<G1Affine> qs(n);
  for (int i = 0; i < n; ++i)
    if (!tr.read_commitment(&qs[i])) return false;
  G1 lhs = G1::from_affine(comm).add(G1::from_affine(G1Affine::generator()).mul(eval).neg());
  if (vp.pairing_check) {
    const std::vector<G2Affine>& sg2 = vp.g2_powers();
    const G2Affine g2 = G2Affine::generator();
    std::vector<std::pair<G1Affine, G2Affine>> terms;
    terms.push_back({lhs.to_affine(), g2.neg()});
    for (int i = 0; i < n; ++i) terms.push_back({qs[i], sg2[i].add(g2.mul(point[i]).neg())});
    return pairings_product_is_identity(terms);
  }
  G1 rhs = G1::identity();
  for (int i = 0; i < n; ++i) rhs = rhs.add(G1::from_affine(qs[i]).mul(vp.ss[i] - point[i]));
  return lhs.eq(rhs);
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

// additive::batch_open, pb/pcs/multilinear.rs:134-235 (sanity-check feature off)
inline bool kzg_batch_open(const KzgParams& pp, int num_vars, const std::vector<const Poly*>& polys,
                           const std::vector<std::vector<Fr>>& points,
                           const std::vector<Evaluation>& evals, Transcript& tr) {
  const int ell = ceil_log2(evals.size());
  std::vector<Fr> t = tr.squeeze_challenges(ell);
  Poly eq_xt = ell ? eq_xy(t) : Poly();  // eq_xy(&[]) is the zero polynomial (multilinear.rs:92-94)
  // merged_polys: (scalar, poly) per point; a single contribution is kept borrowed with its scalar
  struct Merged {
    Fr scalar;
    const Poly* borrowed;
    Poly owned;
    bool empty;
  };
  std::vector<Merged> merged(points.size(), Merged{Fr::one(), nullptr, Poly(), true});
  for (size_t k = 0; k < evals.size(); ++k) {
    if (k >= eq_xt.size()) return false;  // reference would panic on the zip; not reachable for ell>0
    Merged& m = merged[evals[k].point];
    const Fr& e = eq_xt[k];
    if (m.empty) {
      m.scalar = e;
      m.borrowed = polys[evals[k].poly];
      m.empty = false;
    } else {
      if (m.borrowed) {
        m.owned = *m.borrowed;
        m.borrowed = nullptr;
      }
      if (m.scalar != Fr::one()) {
        for (auto& v : m.owned) v = v * m.scalar;
        m.scalar = Fr::one();
      }
      const Poly& p = *polys[evals[k].poly];
      const long n = (long)p.size();
#pragma omp parallel for if (n >= 4096)
      for (long i = 0; i < n; ++i) m.owned[i] = m.owned[i] + e * p[i];
    }
  }
  // unique_by address: borrowed polys that alias share one table (pb/pcs/multilinear.rs:173-181)
  std::vector<const Poly*> uniq;
  std::vector<CoeffProduct> prods;
  for (size_t i = 0; i < merged.size(); ++i) {
    const Poly* addr = merged[i].borrowed ? merged[i].borrowed : &merged[i].owned;
    int idx = -1;
    for (size_t u = 0; u < uniq.size(); ++u)
      if (uniq[u] == addr) idx = (int)u;
    if (idx < 0) {
      idx = (int)uniq.size();
      uniq.push_back(addr);
    }
    prods.push_back(CoeffProduct{merged[i].scalar, points[i], idx});
  }
  Fr tilde = Fr::zero();
  for (size_t k = 0; k < evals.size(); ++k) tilde = tilde + evals[k].value * eq_xt[k];
  SumCheckOutput sc = sumcheck_prove_coeffs(num_vars, prods, uniq, tilde, tr);
  // g' = Σ (scalar_i * eq(challenges, point_i)) * merged_i
  Poly g(((size_t)1) << num_vars, Fr::zero());
  for (size_t i = 0; i < merged.size(); ++i) {
    Fr s = merged[i].scalar * eq_xy_eval(sc.challenges, points[i]);
    const Poly& p = merged[i].borrowed ? *merged[i].borrowed : merged[i].owned;
    const long n = (long)p.size();
#pragma omp parallel for if (n >= 4096)
    for (long k = 0; k < n; ++k) g[k] = g[k] + s * p[k];
  }
  bool ok;
  kzg_open(pp, g, sc.challenges, tr, &ok);
  return ok;
}

// additive::batch_verify, pb/pcs/multilinear.rs:237-275
inline bool kzg_batch_verify(const KzgParams& vp, int num_vars, const std::vector<G1Affine>& comms,
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

}  // namespace oracle

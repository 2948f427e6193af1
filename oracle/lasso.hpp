// ORACLE (test infrastructure only — never linked into the product path).
//
// Lasso / Surge lookup argument. The mounted reference snapshot contains NO Lasso code
// (SURVEY §0 F1/F2), so this file is the *specification* (Lasso paper eprint 2023/1216 §Surge +
// BASELINE.json north_star; written down in DESIGN.md §Lasso protocol) built on the reference
// primitives that do exist: transcript (pb/util/transcript.rs), ClassicSumCheck
// (pb/piop/sum_check/classic*.rs), MultilinearKzg + additive batch open (pb/pcs/multilinear*.rs).
// The layered grand-product argument follows the in-tree template
// pb/piop/gkr/fractional_sum_check.rs:41-190,272-296 specialised from fractions (p,q) to plain
// products: `Layer::bottom` splits on the TOP bit (l = lower half, r = upper half), one batched
// degree-3 sum-check per layer with powers of a fresh gamma, 2 evals per tree written, mu squeezed,
// claims folded l + mu*(r-l), point extended x.push(mu).
//
// parity: UNPINNED (no reference implementation, tests or vectors exist). Validated by
// prove -> verify round trips and soundness negatives (tests/test_oracle_lasso.py).
#pragma once
#include <vector>

#include "kzg.hpp"

namespace oracle {

enum TableKind { TABLE_RANGE = 0, TABLE_AND = 1, TABLE_XOR = 2, TABLE_CUSTOM = 3 };

// A decomposable table in Surge form (the role of a `DecomposableTable` implementation): every lookup splits into
// `chunks` chunks, chunk t addresses ONE 2^16-entry subtable T, output g(E) = Σ_t 2^(out_bits t) E_t.
// TABLE_CUSTOM carries the subtable as data: `values` (2^16 entries), `num_operands` (1: dim_t = chunk t of x;
// 2: dim_t = chunk t of x << operand_bits | chunk t of y), `operand_bits` per operand chunk, `out_bits`.
struct LassoTable {
  int kind;    // TableKind
  int chunks;  // c; every chunk addresses one 2^16-entry subtable
  int num_operands = 1, operand_bits = 16, custom_out_bits = 16;  // TABLE_CUSTOM only
  const uint32_t* values = nullptr;                               // TABLE_CUSTOM only: 2^16 entries
  // bits of the lookup output contributed by one chunk (16 for range, 8 for and/xor)
  int out_bits() const { return kind == TABLE_CUSTOM ? custom_out_bits : (kind == TABLE_RANGE ? 16 : 8); }
  uint64_t subtable(uint32_t x) const {
    if (kind == TABLE_CUSTOM) return values[x];
    if (kind == TABLE_RANGE) return x;
    uint32_t p = x >> 8, q = x & 0xff;
    return kind == TABLE_AND ? (p & q) : (p ^ q);
  }
  // chunk t of lookup j; for and/xor the index interleaves operand bytes: (x_t << 8) | y_t
  uint32_t dim(uint64_t x, uint64_t y, int t) const {
    if (kind == TABLE_CUSTOM) {
      const uint64_t mask = ((uint64_t)1 << operand_bits) - 1;
      const uint32_t xt = (uint32_t)((x >> (operand_bits * t)) & mask), yt = (uint32_t)((y >> (operand_bits * t)) & mask);
      return num_operands == 1 ? xt : (xt << operand_bits) | yt;
    }
    if (kind == TABLE_RANGE) return (uint32_t)((x >> (16 * t)) & 0xffff);
    return (uint32_t)((((x >> (8 * t)) & 0xff) << 8) | ((y >> (8 * t)) & 0xff));
  }
  // MLE of the subtable at a 16-variate point (verifier side)
  Fr subtable_mle(const std::vector<Fr>& x) const {
    if (kind == TABLE_CUSTOM) {  // no closed form: <values, eq(., x)>
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
  // TABLE_CUSTOM: the table is part of the statement — Keccak-256 of the 2^16 values as little-endian u32 words,
  // read as a little-endian integer mod r (the same map squeeze_challenge uses, transcript.rs:127-131)
  Fr digest() const {
    Keccak256 h;
    h.update(reinterpret_cast<const uint8_t*>(values), ((size_t)4) << 16);
    uint8_t out[32];
    h.finalize_reset(out);
    return Fr::from_le_bytes_mod(out);
  }
  bool valid() const {
    if (chunks < 2 || chunks > 8) return false;
    if (kind != TABLE_CUSTOM) return kind >= 0 && kind <= 2 && (kind != TABLE_RANGE || chunks <= 4);
    if (!values || num_operands < 1 || num_operands > 2 || operand_bits < 1 || num_operands * operand_bits > 16) return false;
    if (operand_bits * chunks > 64 || custom_out_bits < 1 || custom_out_bits > 32) return false;
    uint32_t mx = 0;
    for (size_t i = 0; i < ((size_t)1 << 16); ++i) mx = values[i] > mx ? values[i] : mx;
    int eb = 0;
    while (eb < 32 && (mx >> eb)) ++eb;
    return custom_out_bits * (chunks - 1) + eb <= 64;  // the lookup output fits a u64
  }
};

static const int SUBTABLE_VARS = 16;

struct LassoWitness {
  int mu;  // log2(#lookups)
  Poly a;                        // lookup outputs
  std::vector<Poly> dim, e, read_ts;  // c polys each, 2^mu
  std::vector<Poly> final_cts;        // c polys, 2^16
};

// E_t[j] = T[dim_t[j]], read_ts_t[j] = #{j' < j : dim_t[j'] == dim_t[j]},
// final_cts_t[x] = #{j : dim_t[j] == x}.
inline LassoWitness lasso_witness(const LassoTable& tb, int mu, const uint64_t* xs,
                                  const uint64_t* ys) {
  const size_t m = (size_t)1 << mu, S = (size_t)1 << SUBTABLE_VARS;
  const int c = tb.chunks;
  LassoWitness w;
  w.mu = mu;
  w.a.assign(m, Fr::zero());
  w.dim.assign(c, Poly(m));
  w.e.assign(c, Poly(m));
  w.read_ts.assign(c, Poly(m));
  w.final_cts.assign(c, Poly(S));
  for (int t = 0; t < c; ++t) {
    std::vector<uint64_t> cnt(S, 0);
    for (size_t j = 0; j < m; ++j) {
      uint32_t d = tb.dim(xs[j], ys ? ys[j] : 0, t);
      w.dim[t][j] = Fr::from_u64(d);
      w.e[t][j] = Fr::from_u64(tb.subtable(d));
      w.read_ts[t][j] = Fr::from_u64(cnt[d]++);
    }
    for (size_t x = 0; x < S; ++x) w.final_cts[t][x] = Fr::from_u64(cnt[x]);
  }
  for (size_t j = 0; j < m; ++j) {
    uint64_t out = 0;
    for (int t = 0; t < c; ++t)
      out += tb.subtable(tb.dim(xs[j], ys ? ys[j] : 0, t)) << (tb.out_bits() * t);  // g = Σ_t 2^(out_bits t) E_t
    w.a[j] = Fr::from_u64(out);
  }
  return w;
}

struct GrandProductOutput {
  std::vector<Fr> claims;                  // per tree: leaf-layer MLE at points[height of that tree]
  std::vector<std::vector<Fr>> points;     // points[h] = the point after h layers (h variables)
  std::vector<Fr> point;                   // == points[max height]
};

// Batched layered product argument over T trees (template: fractional_sum_check.rs). Trees may have
// DIFFERENT heights h_t = log2(#leaves): all roots sit at layer 0, layer k batches the trees that are
// still running (h_t > k), in input order, with weights gamma^i over that active list. A tree's claim
// freezes when its leaf layer is reached; its point is the running point after h_t layers.
inline GrandProductOutput grand_product_prove(const std::vector<Poly>& leaves, Transcript& tr,
                                              std::vector<Fr>* roots_out) {
  const int T = (int)leaves.size();
  std::vector<int> hs(T);
  int h = 0;
  for (int t = 0; t < T; ++t) {
    hs[t] = log2_exact(leaves[t].size());
    h = hs[t] > h ? hs[t] : h;
  }
  // layers[t][k]: 2^k nodes; node i = child[i] * child[i + 2^k]
  std::vector<std::vector<Poly>> layers(T);
  for (int t = 0; t < T; ++t) {
    layers[t].resize(hs[t] + 1);
    layers[t][hs[t]] = leaves[t];
    for (int k = hs[t] - 1; k >= 0; --k) {
      const Poly& ch = layers[t][k + 1];
      const long half = 1L << k;
      Poly up(half);
#pragma omp parallel for if (half >= 4096)
      for (long i = 0; i < half; ++i) up[i] = ch[i] * ch[i + half];
      layers[t][k] = up;
    }
  }
  std::vector<Fr> claims(T);
  for (int t = 0; t < T; ++t) claims[t] = layers[t][0][0];
  tr.write_field_elements(claims.data(), T);
  if (roots_out) *roots_out = claims;

  GrandProductOutput out;
  out.points.resize(h + 1);
  std::vector<Fr> y;
  for (int k = 0; k < h; ++k) {
    const size_t half = (size_t)1 << k;
    std::vector<int> act;
    for (int t = 0; t < T; ++t)
      if (hs[t] > k) act.push_back(t);
    const int A = (int)act.size();
    std::vector<Poly> ls(A), rs(A);
    for (int i = 0; i < A; ++i) {
      const Poly& ch = layers[act[i]][k + 1];
      ls[i].assign(ch.begin(), ch.begin() + half);
      rs[i].assign(ch.begin() + half, ch.end());
    }
    std::vector<Fr> x, evals;
    if (k == 0) {
      for (int i = 0; i < A; ++i) {
        evals.push_back(ls[i][0]);
        evals.push_back(rs[i][0]);
      }
    } else {
      Fr gamma = tr.squeeze_challenge();
      VirtualPoly vp;
      vp.has_eq = true;
      vp.y = y;
      Fr pw = Fr::one(), claim = Fr::zero();
      for (int i = 0; i < A; ++i) {
        vp.polys.push_back(&ls[i]);
        vp.polys.push_back(&rs[i]);
        vp.terms.push_back(Term{pw, {2 * i, 2 * i + 1}});
        claim = claim + pw * claims[act[i]];
        pw = pw * gamma;
      }
      SumCheckOutput sc = sumcheck_prove_evals(k, vp, claim, tr);
      x = sc.challenges;
      evals = sc.evals;
    }
    tr.write_field_elements(evals.data(), evals.size());
    Fr mu = tr.squeeze_challenge();
    for (int i = 0; i < A; ++i) claims[act[i]] = evals[2 * i] + mu * (evals[2 * i + 1] - evals[2 * i]);
    x.push_back(mu);
    y = x;
    out.points[k + 1] = y;
  }
  out.claims = claims;
  out.point = y;
  return out;
}

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
  if (tb.kind == TABLE_CUSTOM) {  // a table given as data is part of the statement
    tr.common_field_element(Fr::from_u64((uint64_t)tb.num_operands));
    tr.common_field_element(Fr::from_u64((uint64_t)tb.operand_bits));
    tr.common_field_element(Fr::from_u64((uint64_t)tb.custom_out_bits));
    tr.common_field_element(tb.digest());
  }
}

// Full Lasso proof (DESIGN.md §Lasso protocol, steps 1-10). Returns false when a commitment is the
// identity (the reference transcript cannot encode it, transcript.rs:174-179).
inline bool lasso_prove(const KzgParams& pp, const LassoTable& tb, int mu, const uint64_t* xs,
                        const uint64_t* ys, Transcript& tr) {
  const int c = tb.chunks;
  const size_t m = (size_t)1 << mu, S = (size_t)1 << SUBTABLE_VARS;
  LassoWitness w = lasso_witness(tb, mu, xs, ys);
  lasso_absorb_statement(tb, mu, tr);

  // 2. commitments: a, dim_*, E_*, read_ts_*, final_cts_*
  std::vector<const Poly*> mpolys;
  mpolys.push_back(&w.a);
  for (auto& p : w.dim) mpolys.push_back(&p);
  for (auto& p : w.e) mpolys.push_back(&p);
  for (auto& p : w.read_ts) mpolys.push_back(&p);
  for (auto* p : mpolys)
    if (!tr.write_commitment(kzg_commit(pp, *p))) return false;
  for (auto& p : w.final_cts)
    if (!tr.write_commitment(kzg_commit(pp, p))) return false;

  // 3-5. primary Surge sum-check  Σ_j eq(r,j) * Σ_t 2^{out_bits*t} E_t(j) = a(r)
  std::vector<Fr> r = tr.squeeze_challenges(mu);
  Fr v_a = evaluate(w.a, r);
  tr.write_field_element(v_a);
  VirtualPoly vp;
  vp.has_eq = true;
  vp.y = r;
  for (int t = 0; t < c; ++t) {
    vp.polys.push_back(&w.e[t]);
    vp.terms.push_back(Term{Fr::from_u64((uint64_t)1 << (tb.out_bits() * t)), {t}});
  }
  SumCheckOutput prim = sumcheck_prove_evals(mu, vp, v_a, tr);
  tr.write_field_elements(prim.evals.data(), prim.evals.size());

  // 6-7. fingerprints h(a,v,t) = a*gamma^2 + v*gamma + t - tau
  Fr gamma = tr.squeeze_challenge(), tau = tr.squeeze_challenge();
  Fr gamma2 = gamma.sqr();
  std::vector<Poly> mleaves(2 * c, Poly(m)), sleaves(2 * c, Poly(S));
  for (int t = 0; t < c; ++t) {
#pragma omp parallel for
    for (long j = 0; j < (long)m; ++j) {
      Fr rd = w.dim[t][j] * gamma2 + w.e[t][j] * gamma + w.read_ts[t][j] - tau;
      mleaves[2 * t][j] = rd;
      mleaves[2 * t + 1][j] = rd + Fr::one();
    }
#pragma omp parallel for
    for (long x = 0; x < (long)S; ++x) {
      Fr in = Fr::from_u64((uint64_t)x) * gamma2 + Fr::from_u64(tb.subtable((uint32_t)x)) * gamma - tau;
      sleaves[2 * t][x] = in;
      sleaves[2 * t + 1][x] = in + w.final_cts[t][x];
    }
  }
  // 8. ONE batched grand product over all 4c trees: [Read_t, Write_t]_t (height mu) then
  //    [Init_t, Final_t]_t (height 16); x_m / x_s are the running points after mu / 16 layers
  std::vector<Poly> all_leaves = mleaves;
  all_leaves.insert(all_leaves.end(), sleaves.begin(), sleaves.end());
  GrandProductOutput gp = grand_product_prove(all_leaves, tr, nullptr);
  GrandProductOutput gm, gs;
  gm.point = gp.points[mu];
  gs.point = gp.points[SUBTABLE_VARS];

  // 9. leaf openings
  std::vector<Fr> ev_dim(c), ev_e(c), ev_ts(c), ev_cts(c);
  for (int t = 0; t < c; ++t) {
    ev_dim[t] = evaluate(w.dim[t], gm.point);
    ev_e[t] = evaluate(w.e[t], gm.point);
    ev_ts[t] = evaluate(w.read_ts[t], gm.point);
    ev_cts[t] = evaluate(w.final_cts[t], gs.point);
  }
  tr.write_field_elements(ev_dim.data(), c);
  tr.write_field_elements(ev_e.data(), c);
  tr.write_field_elements(ev_ts.data(), c);
  tr.write_field_elements(ev_cts.data(), c);

  // 10. batch openings (mu-variate: points r, x_p, x_m; 16-variate: x_s)
  std::vector<std::vector<Fr>> pts = {r, prim.challenges, gm.point};
  std::vector<Evaluation> evs;
  evs.push_back(Evaluation{0, 0, v_a});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + c + t, 1, prim.evals[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + t, 2, ev_dim[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + c + t, 2, ev_e[t]});
  for (int t = 0; t < c; ++t) evs.push_back(Evaluation{1 + 2 * c + t, 2, ev_ts[t]});
  if (!kzg_batch_open(pp, mu, mpolys, pts, evs, tr)) return false;
  std::vector<const Poly*> spolys;
  for (auto& p : w.final_cts) spolys.push_back(&p);
  std::vector<Evaluation> sevs;
  for (int t = 0; t < c; ++t) sevs.push_back(Evaluation{t, 0, ev_cts[t]});
  return kzg_batch_open(pp, SUBTABLE_VARS, spolys, {gs.point}, sevs, tr);
}

inline bool lasso_verify(const KzgParams& vp, const LassoTable& tb, int mu, Transcript& tr) {
  const int c = tb.chunks;
  lasso_absorb_statement(tb, mu, tr);
  std::vector<G1Affine> mcomms(1 + 3 * c), scomms(c);
  for (auto& p : mcomms)
    if (!tr.read_commitment(&p)) return false;
  for (auto& p : scomms)
    if (!tr.read_commitment(&p)) return false;
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
  return tr.rpos == tr.stream.size();
}

}  // namespace oracle

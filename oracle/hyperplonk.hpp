// ORACLE (test infrastructure only — never linked into the product path).
//
// HyperPlonk over MultilinearKzg for circuits WITHOUT lookups (the LogUp branch of the snapshot is empty when
// `lookups` is empty, prover.rs:56-58), restating
//   HyperPlonk::{preprocess, prove, verify}     pb/backend/hyperplonk.rs:97-363
//   instance_polys, permutation_z_polys,        pb/backend/hyperplonk/prover.rs:32-48, 252-345, 348-409
//   prove_sum_check
//   permutation_polys                           pb/backend/hyperplonk/preprocessor.rs:172-203
//   verify_sum_check, instance_evals,           pb/backend/hyperplonk/verifier.rs:39-182
//   pcs_query, points, point_offset
//   rotation_eval(_points/_patterns)            pb/poly/multilinear.rs:433-570
//   evaluate / lagrange_eval                    pb/piop/sum_check.rs:60-125
// The zero-check expression itself is composed by the caller (tests build it with the product's
// expression.py::compose, which is KAT-tested against preprocessor.rs:216-256) and arrives as a token stream.
#pragma once
#include <algorithm>
#include <array>
#include <functional>
#include <map>
#include <set>

#include "expression.hpp"
#include "kzg.hpp"

namespace oracle {

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

struct HyperPlonkParams {
  const KzgParams* kzg;
  int num_vars;
  std::vector<int> num_instances;     // per instance column (backend.rs:50-51)
  int num_witness_polys;              // over all phases
  std::vector<int> phase_witness_polys = {};  // per phase (backend.rs:55-60); empty = one phase without challenges
  std::vector<int> phase_challenges = {};
  ExprP expression;
  std::vector<Poly> preprocess_polys;
  std::vector<G1Affine> preprocess_comms;
  std::vector<int> permutation_poly_idx;
  std::vector<Poly> permutation_polys;
  std::vector<G1Affine> permutation_comms;
  int num_permutation_z_polys;
  std::vector<std::vector<std::pair<ExprP, ExprP>>> lookups;  // LogUp: per lookup the (input, table) column pairs
};

// ---- LogUp helper polynomials (pb/backend/hyperplonk/prover.rs:50-250) ----------------------------------
// Expression::evaluate on one hypercube row (prover.rs:96-117): queries read poly[bh.rotate(b, rotation)],
// Lagrange(i) is 1 on row bh[i mod 2^n], the identity polynomial is F::from(b).
inline Fr expr_eval_row(const ExprP& e, uint64_t b, const BooleanHypercube& bh, const std::vector<uint64_t>& order,
                        const std::vector<const Poly*>& polys, const std::vector<Fr>& challenges) {
  switch (e->kind) {
    case Expr::CONST: return e->scalar;
    case Expr::IDENTITY: return Fr::from_u64(b);
    case Expr::LAGRANGE: {
      const long N = (long)order.size();
      return order[((e->a % N) + N) % N] == b ? Fr::one() : Fr::zero();
    }
    case Expr::EQXY: return Fr::zero();  // unreachable!() in the reference
    case Expr::POLY: return (*polys[e->a])[bh.rotate(b, e->b)];
    case Expr::CHALLENGE: return challenges[e->a];
    case Expr::NEG: return -expr_eval_row(e->ch[0], b, bh, order, polys, challenges);
    case Expr::SUM:
      return expr_eval_row(e->ch[0], b, bh, order, polys, challenges) + expr_eval_row(e->ch[1], b, bh, order, polys, challenges);
    case Expr::PROD:
      return expr_eval_row(e->ch[0], b, bh, order, polys, challenges) * expr_eval_row(e->ch[1], b, bh, order, polys, challenges);
    case Expr::SCALED: return expr_eval_row(e->ch[0], b, bh, order, polys, challenges) * e->scalar;
    case Expr::DPOW: {
      const size_t n = e->ch.size() - 1;
      const Fr base = expr_eval_row(e->ch[n], b, bh, order, polys, challenges);
      Fr acc = expr_eval_row(e->ch[0], b, bh, order, polys, challenges), pw = base;
      for (size_t i = 1; i < n; ++i) {
        acc = acc + pw * expr_eval_row(e->ch[i], b, bh, order, polys, challenges);
        pw = pw * base;
      }
      return acc;
    }
  }
  return Fr::zero();
}

// lookup_compressed_poly (prover.rs:78-134): [Σ_j beta^j input_j, Σ_j beta^j table_j]
inline std::array<Poly, 2> lookup_compressed_poly(const std::vector<std::pair<ExprP, ExprP>>& lookup, int num_vars,
                                                  const std::vector<const Poly*>& polys,
                                                  const std::vector<Fr>& challenges, const Fr& beta) {
  const size_t N = (size_t)1 << num_vars;
  BooleanHypercube bh(num_vars);
  const std::vector<uint64_t> order = bh.iter();
  std::array<Poly, 2> out = {Poly(N, Fr::zero()), Poly(N, Fr::zero())};
  Fr pw = Fr::one();
  for (auto& col : lookup) {
    for (size_t b = 0; b < N; ++b) {
      out[0][b] = out[0][b] + pw * expr_eval_row(col.first, b, bh, order, polys, challenges);
      out[1][b] = out[1][b] + pw * expr_eval_row(col.second, b, bh, order, polys, challenges);
    }
    pw = pw * beta;
  }
  return out;
}

// lookup_m_poly (prover.rs:143-192): multiplicities; a value that occurs on several table rows is counted on
// the LAST of them (HashMap::from_iter keeps the latest entry). false = "Invalid lookup input".
inline bool lookup_m_poly(const std::array<Poly, 2>& compressed, Poly* m) {
  const Poly &input = compressed[0], &table = compressed[1];
  std::map<std::array<uint64_t, 4>, size_t> index;
  auto key = [](const Fr& f) { return std::array<uint64_t, 4>{f.l[0], f.l[1], f.l[2], f.l[3]}; };
  for (size_t i = 0; i < table.size(); ++i) index[key(table[i])] = i;
  std::vector<uint64_t> counts(input.size(), 0);
  for (auto& v : input) {
    auto it = index.find(key(v));
    if (it == index.end()) return false;
    counts[it->second] += 1;
  }
  m->resize(input.size());
  for (size_t i = 0; i < counts.size(); ++i) (*m)[i] = Fr::from_u64(counts[i]);
  return true;
}

// lookup_h_poly (prover.rs:206-250): h = 1/(gamma + input) - m/(gamma + table)
inline Poly lookup_h_poly(const std::array<Poly, 2>& compressed, const Poly& m, const Fr& gamma) {
  const size_t N = m.size();
  Poly hi(N), ht(N);
  for (size_t b = 0; b < N; ++b) {
    hi[b] = gamma + compressed[0][b];
    ht[b] = gamma + compressed[1][b];
  }
  batch_invert(hi.data(), N);
  batch_invert(ht.data(), N);
  for (size_t b = 0; b < N; ++b) hi[b] = hi[b] - ht[b] * m[b];
  return hi;
}

// hyperplonk.rs:365-369 + prover.rs:32-48
inline std::vector<Poly> instance_polys(int num_vars, const std::vector<std::vector<Fr>>& instances) {
  std::vector<uint64_t> row = BooleanHypercube(num_vars).iter();
  std::vector<Poly> out;
  for (auto& inst : instances) {
    Poly p((size_t)1 << num_vars, Fr::zero());
    for (size_t i = 0; i < inst.size(); ++i) p[i + 1 < row.size() ? row[i + 1] : 0] = inst[i];
    out.push_back(p);
  }
  return out;
}

// preprocessor.rs:172-203; cycles: list of cycles of (poly, row)
inline std::vector<Poly> permutation_polys(int num_vars, const std::vector<int>& perm_idx,
                                           const std::vector<std::vector<std::pair<int, int>>>& cycles) {
  std::map<int, int> poly_index;
  for (size_t i = 0; i < perm_idx.size(); ++i) poly_index[perm_idx[i]] = (int)i;
  const size_t N = (size_t)1 << num_vars;
  std::vector<Poly> perms(perm_idx.size(), Poly(N));
  for (size_t i = 0; i < perm_idx.size(); ++i)
    for (size_t j = 0; j < N; ++j) perms[i][j] = Fr::from_u64(((uint64_t)i << num_vars) + j);
  for (auto& cyc : cycles) {
    Fr last = perms[poly_index[cyc[0].first]][cyc[0].second];
    for (size_t k = 1; k <= cyc.size(); ++k) {
      auto& ij = cyc[k % cyc.size()];
      std::swap(perms[poly_index[ij.first]][ij.second], last);
    }
  }
  return perms;
}

// prover.rs:252-345
inline std::vector<Poly> permutation_z_polys(int num_chunks, const std::vector<int>& perm_idx,
                                             const std::vector<Poly>& perm_polys, const std::vector<const Poly*>& polys,
                                             const Fr& beta, const Fr& gamma) {
  if (perm_idx.empty()) return {};
  const int chunk = ((int)perm_idx.size() + num_chunks - 1) / num_chunks;
  const int num_vars = log2_exact(polys[0]->size());
  const size_t N = (size_t)1 << num_vars;
  std::vector<Poly> products;
  for (int c = 0; c * chunk < (int)perm_idx.size(); ++c) {
    Poly prod(N, Fr::one());
    const int lo = c * chunk, hi = std::min((int)perm_idx.size(), lo + chunk);
    for (int i = lo; i < hi; ++i)
      for (size_t b = 0; b < N; ++b) prod[b] = prod[b] * (beta * perm_polys[i][b] + gamma + (*polys[perm_idx[i]])[b]);
    batch_invert(prod.data(), N);
    for (int i = lo; i < hi; ++i) {
      const uint64_t id_offset = (uint64_t)i << num_vars;
      for (size_t b = 0; b < N; ++b)
        prod[b] = prod[b] * (Fr::from_u64(id_offset + b) * beta + gamma + (*polys[perm_idx[i]])[b]);
    }
    products.push_back(prod);
  }
  std::vector<uint64_t> order = BooleanHypercube(num_vars).iter();
  std::vector<Fr> z((size_t)num_chunks << num_vars, Fr::zero());
  z[num_chunks] = Fr::one();
  {
    Fr state = Fr::one();
    size_t pos = num_chunks + 1;
    for (size_t k = 1; k < order.size() && pos < z.size(); ++k)
      for (int c = 0; c < num_chunks && pos < z.size(); ++c) {
        state = state * products[c][order[k]];
        z[pos++] = state;
      }
  }
  std::vector<size_t> nth(N);
  for (size_t i = 0; i < order.size(); ++i) nth[order[i]] = i;
  std::vector<Poly> out(num_chunks, Poly(N));
  for (int c = 0; c < num_chunks; ++c)
    for (size_t b = 0; b < N; ++b) out[c][b] = z[c + num_chunks * nth[b]];
  return out;
}

inline HyperPlonkParams hyperplonk_preprocess(const KzgParams& kzg, int num_vars, const ExprP& expression,
                                              const std::vector<int>& num_instances, int num_witness_polys,
                                              const std::vector<Poly>& preprocess_polys,
                                              const std::vector<int>& perm_idx,
                                              const std::vector<std::vector<std::pair<int, int>>>& cycles,
                                              int num_permutation_z_polys,
                                              const std::vector<std::vector<std::pair<ExprP, ExprP>>>& lookups = {}) {
  HyperPlonkParams pp;
  pp.lookups = lookups;
  pp.kzg = &kzg;
  pp.num_vars = num_vars;
  pp.num_instances = num_instances;
  pp.num_witness_polys = num_witness_polys;
  pp.expression = expression;
  pp.preprocess_polys = preprocess_polys;
  for (auto& p : preprocess_polys) pp.preprocess_comms.push_back(kzg_commit(kzg, p));
  pp.permutation_poly_idx = perm_idx;
  pp.permutation_polys = permutation_polys(num_vars, perm_idx, cycles);
  for (auto& p : pp.permutation_polys) pp.permutation_comms.push_back(kzg_commit(kzg, p));
  pp.num_permutation_z_polys = num_permutation_z_polys;
  return pp;
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

// `PlonkishCircuit::synthesize(round, challenges)` (backend.rs:100-110): the witness polynomials of one phase
typedef std::function<std::vector<Poly>(int, const std::vector<Fr>&)> Synthesize;

// hyperplonk.rs:164-291
inline bool hyperplonk_prove_phased(const HyperPlonkParams& pp, const std::vector<std::vector<Fr>>& instances,
                                    const Synthesize& synthesize, Transcript& tr) {
  const int n = pp.num_vars;
  if (instances.size() != pp.num_instances.size()) return false;
  for (size_t i = 0; i < instances.size(); ++i)
    if ((int)instances[i].size() != pp.num_instances[i]) return false;  // assert_eq! in the reference
  for (auto& inst : instances)
    for (auto& v : inst) tr.common_field_element(v);
  std::vector<Poly> inst_polys = instance_polys(n, instances);
  // rounds 0..n (hyperplonk.rs:183-204): synthesize, commit, squeeze the phase's challenges
  const std::vector<int> phase_w = pp.phase_witness_polys.empty() ? std::vector<int>{pp.num_witness_polys} : pp.phase_witness_polys;
  const std::vector<int> phase_c = pp.phase_challenges.empty() ? std::vector<int>{0} : pp.phase_challenges;
  std::vector<Poly> witness_polys;
  std::vector<Fr> challenges;
  for (size_t round = 0; round < phase_w.size(); ++round) {
    std::vector<Poly> ps = synthesize((int)round, challenges);
    if ((int)ps.size() != phase_w[round]) return false;
    for (auto& w : ps) {
      if (w.size() != (size_t)1 << n) return false;
      if (!tr.write_commitment(kzg_commit(*pp.kzg, w))) return false;
    }
    for (auto& w : ps) witness_polys.push_back(std::move(w));
    for (auto& c : tr.squeeze_challenges(phase_c[round])) challenges.push_back(c);
  }
  std::vector<const Poly*> polys;
  for (auto& p : inst_polys) polys.push_back(&p);
  for (auto& p : pp.preprocess_polys) polys.push_back(&p);
  for (auto& p : witness_polys) polys.push_back(&p);
  const Fr beta = tr.squeeze_challenge();
  std::vector<std::array<Poly, 2>> compressed;
  std::vector<Poly> ms(pp.lookups.size()), hs;
  for (size_t l = 0; l < pp.lookups.size(); ++l) {
    compressed.push_back(lookup_compressed_poly(pp.lookups[l], n, polys, challenges, beta));
    if (!lookup_m_poly(compressed[l], &ms[l])) return false;  // Error::InvalidSnark("Invalid lookup input")
  }
  for (auto& m : ms)
    if (!tr.write_commitment(kzg_commit(*pp.kzg, m))) return false;
  const Fr gamma = tr.squeeze_challenge();
  for (size_t l = 0; l < pp.lookups.size(); ++l) hs.push_back(lookup_h_poly(compressed[l], ms[l], gamma));
  std::vector<Poly> zs = permutation_z_polys(pp.num_permutation_z_polys, pp.permutation_poly_idx, pp.permutation_polys,
                                             polys, beta, gamma);
  for (auto& h : hs)
    if (!tr.write_commitment(kzg_commit(*pp.kzg, h))) return false;
  for (auto& z : zs)
    if (!tr.write_commitment(kzg_commit(*pp.kzg, z))) return false;
  const Fr alpha = tr.squeeze_challenge();
  std::vector<Fr> y = tr.squeeze_challenges(n);
  for (auto& p : pp.permutation_polys) polys.push_back(&p);
  for (auto& m : ms) polys.push_back(&m);
  for (auto& h : hs) polys.push_back(&h);
  for (auto& z : zs) polys.push_back(&z);
  challenges.insert(challenges.end(), {beta, gamma, alpha});
  // prove_zero_check (prover.rs:348-409)
  SumCheckOutput sc = sumcheck_prove_generic(n, pp.expression, polys, challenges, {y}, Fr::zero(), tr);
  PcsQueryPlan pl = pcs_query_plan(pp.expression, (int)instances.size());
  std::vector<std::vector<Fr>> points;
  for (int r : pl.rotations)
    for (auto& p : rotation_eval_points(sc.challenges, r)) points.push_back(p);
  std::vector<Evaluation> evals;
  for (auto& q : pl.queries) {
    if (q.second == 0) {
      evals.push_back(Evaluation{q.first, pl.point_offset[0], sc.evals[q.first]});
    } else {  // evaluate_for_rotation == evaluations at rotation_eval_points (multilinear.rs:191-263)
      int pt = pl.point_offset[q.second];
      for (auto& p : rotation_eval_points(sc.challenges, q.second)) evals.push_back(Evaluation{q.first, pt++, evaluate(*polys[q.first], p)});
    }
  }
  for (auto& e : evals) tr.write_field_element(e.value);
  return kzg_batch_open(*pp.kzg, n, polys, points, evals, tr);
}

// single-phase circuits: the witness is known up front
inline bool hyperplonk_prove(const HyperPlonkParams& pp, const std::vector<std::vector<Fr>>& instances,
                             const std::vector<Poly>& witness_polys, Transcript& tr) {
  if (pp.phase_witness_polys.size() > 1) return false;
  return hyperplonk_prove_phased(pp, instances, [&](int, const std::vector<Fr>&) { return witness_polys; }, tr);
}

// hyperplonk.rs:293-363 + verifier.rs:39-145
inline bool hyperplonk_verify(const HyperPlonkParams& vp, const std::vector<std::vector<Fr>>& instances, Transcript& tr) {
  const int n = vp.num_vars;
  for (auto& inst : instances)
    for (auto& v : inst) tr.common_field_element(v);
  if (instances.size() != vp.num_instances.size()) return false;
  for (size_t i = 0; i < instances.size(); ++i)
    if ((int)instances[i].size() != vp.num_instances[i]) return false;  // hyperplonk.rs:299-305
  // rounds 0..n (hyperplonk.rs:307-315)
  const std::vector<int> phase_w = vp.phase_witness_polys.empty() ? std::vector<int>{vp.num_witness_polys} : vp.phase_witness_polys;
  const std::vector<int> phase_c = vp.phase_challenges.empty() ? std::vector<int>{0} : vp.phase_challenges;
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
  std::vector<G1Affine> m_comms(vp.lookups.size());
  for (auto& c : m_comms)
    if (!tr.read_commitment(&c)) return false;
  const Fr gamma = tr.squeeze_challenge();
  std::vector<G1Affine> z_comms(vp.lookups.size() + vp.num_permutation_z_polys);  // h polys, then z polys
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
  if (!kzg_batch_verify(*vp.kzg, n, comms, points, evals, tr)) return false;
  return tr.rpos == tr.stream.size();
}

}  // namespace oracle

// ORACLE (test infrastructure only — never linked into the product path).
//
// Generic ClassicSumCheck<EvaluationsProver> over an arbitrary `Expression`, restating
//   Expression / Query / Rotation / CommonPolynomial   pb/util/expression.rs:13-182
//   BooleanHypercube (LFSR row order, rotate)           pb/util/arithmetic/bh.rs:5-153
//   ProverState::{new,next_round,into_evals}            pb/piop/sum_check/classic.rs:41-150
//   SumCheckEvaluator::evaluate_polys_next / evaluate   pb/piop/sum_check/classic/eval.rs:210-323
// The reference compiles the expression into a CSE'd straight-line program (ExpressionRegistry) and
// splits Lagrange terms into sparse evaluators; both only change HOW the values p(1..d) are computed,
// not the field elements, so this restatement walks the expression tree directly (expression.rs:109-169)
// with the same leaf semantics: identity = bound part + 2^round * X + b * 2^(round+1), Lagrange = (b, value)
// pairs halved every round, eq tables bound every round, rotated queries read through bh.rotate in
// round 0 and materialised + bound afterwards.
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "sumcheck.hpp"

namespace oracle {

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

// prefix token stream -> tree (see oracle.py serialize_expression)
inline ExprP parse_expr(const int*& t, const Fr* consts) {
  ExprP e = std::make_shared<Expr>();
  const int k = *t++;
  e->kind = (Expr::Kind)k;
  switch (k) {
    case Expr::CONST: e->scalar = consts[*t++]; break;
    case Expr::IDENTITY: break;
    case Expr::LAGRANGE: e->a = *t++; break;
    case Expr::EQXY: e->a = *t++; break;
    case Expr::POLY: e->a = *t++; e->b = *t++; break;
    case Expr::CHALLENGE: e->a = *t++; break;
    case Expr::NEG: e->ch.push_back(parse_expr(t, consts)); break;
    case Expr::SUM:
    case Expr::PROD:
      e->ch.push_back(parse_expr(t, consts));
      e->ch.push_back(parse_expr(t, consts));
      break;
    case Expr::SCALED:
      e->scalar = consts[*t++];
      e->ch.push_back(parse_expr(t, consts));
      break;
    case Expr::DPOW: {
      const int n = *t++;
      for (int i = 0; i <= n; ++i) e->ch.push_back(parse_expr(t, consts));
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

// ClassicSumCheck::<EvaluationsProver>::prove for an arbitrary expression (classic.rs:208-240)
inline SumCheckOutput sumcheck_prove_generic(int num_vars, const ExprP& expr, const std::vector<const Poly*>& polys,
                                             const std::vector<Fr>& challenges, const std::vector<std::vector<Fr>>& ys,
                                             Fr sum, Transcript& tr) {
  const int d = expr_degree(expr);
  std::vector<std::pair<int, int>> queries;
  std::vector<int> lag_ids;
  expr_collect(expr, &queries, &lag_ids);
  // ProverState::new (classic.rs:41-84)
  BooleanHypercube bh(num_vars);
  const std::vector<uint64_t> bh_order = bh.iter();
  std::map<int, std::pair<uint64_t, Fr>> lagranges;
  for (int i : lag_ids) {
    const long N = 1L << num_vars;
    lagranges[i] = {bh_order[((i % N) + N) % N], Fr::one()};
  }
  std::vector<Poly> eqs;
  for (auto& y : ys) eqs.push_back(eq_xy(y));
  std::map<std::pair<int, int>, Poly> tabs;  // bound tables per (poly, rotation); rotation 0 starts as the input
  for (auto& q : queries)
    if (q.second == 0) tabs[q] = *polys[q.first];
  for (size_t p = 0; p < polys.size(); ++p) tabs[{(int)p, 0}] = *polys[p];  // every poly is bound and returned
  Fr identity = Fr::zero();
  const std::vector<Fr> points = points_0_to_d(d);
  const std::vector<Fr> weights = barycentric_weights(points);

  SumCheckOutput out;
  for (int round = 0; round < num_vars; ++round) {
    const long size = 1L << (num_vars - round - 1);
    std::vector<Fr> evals(d + 1, Fr::zero());
    for (long b = 0; b < size; ++b) {
      // evaluate_polys_next::<_, true>: x = 1 loads eval = t[b1], step = t[b1] - t[b0]
      LeafValues lv, st;
      lv.challenges = challenges.data();
      lv.identity = identity + Fr::from_u64(((uint64_t)1 << round) + ((uint64_t)b << (round + 1)));
      st.identity = Fr::from_u64((uint64_t)1 << round);
      for (auto& kv : lagranges) {
        Fr ev = Fr::zero(), sp = Fr::zero();
        if ((uint64_t)b == (kv.second.first >> 1)) {
          if ((kv.second.first & 1) == 0) {
            sp = -kv.second.second;
          } else {
            ev = kv.second.second;
            sp = kv.second.second;
          }
        }
        lv.lagrange[kv.first] = ev;
        st.lagrange[kv.first] = sp;
      }
      for (auto& eq : eqs) {
        lv.eq.push_back(eq[2 * b + 1]);
        st.eq.push_back(eq[2 * b + 1] - eq[2 * b]);
      }
      for (auto& q : queries) {
        uint64_t b0 = 2 * b, b1 = 2 * b + 1;
        const Poly* t;
        if (round == 0) {  // rotated queries read the ORIGINAL table through the LFSR map (eval.rs:216-226, 258-263)
          b0 = bh.rotate(b0, q.second);
          b1 = bh.rotate(b1, q.second);
          t = polys[q.first];
        } else {
          t = &tabs.at(q);
        }
        lv.poly[q] = (*t)[b1];
        st.poly[q] = (*t)[b1] - (*t)[b0];
      }
      for (int x = 1; x <= d; ++x) {
        if (x > 1) {  // eval += step (eval.rs:275-286)
          lv.identity = lv.identity + st.identity;
          for (auto& kv : lv.lagrange) kv.second = kv.second + st.lagrange[kv.first];
          for (size_t i = 0; i < lv.eq.size(); ++i) lv.eq[i] = lv.eq[i] + st.eq[i];
          for (auto& kv : lv.poly) kv.second = kv.second + st.poly[kv.first];
        }
        evals[x] = evals[x] + expr_eval(expr, lv);
      }
    }
    evals[0] = sum - evals[1];
    tr.write_field_elements(evals.data(), evals.size());
    const Fr r = tr.squeeze_challenge();
    out.challenges.push_back(r);
    sum = barycentric_interpolate(weights, points, evals, r);
    // ProverState::next_round (classic.rs:90-141)
    identity = identity + Fr::from_u64((uint64_t)1 << round) * r;
    for (auto& kv : lagranges) {
      kv.second.second = kv.second.second * ((kv.second.first & 1) ? r : Fr::one() - r);
      kv.second.first >>= 1;
    }
    for (auto& eq : eqs) fix_var_in_place(eq, r);
    if (round == 0) {
      for (auto& q : queries)
        if (q.second != 0 && !tabs.count(q)) {
          Poly rot(polys[q.first]->size());
          for (size_t b = 0; b < rot.size(); ++b) rot[b] = (*polys[q.first])[bh.rotate(b, q.second)];
          tabs[q] = rot;
        }
    }
    for (auto& kv : tabs) fix_var_in_place(kv.second, r);
    bh = BooleanHypercube(num_vars - round - 1 > 0 ? num_vars - round - 1 : 1);
  }
  for (size_t p = 0; p < polys.size(); ++p) out.evals.push_back(tabs.at({(int)p, 0})[0]);
  return out;
}

}  // namespace oracle

// Host side of the generic sum-check: the `Expression` tree (pb/util/expression.rs:14-182, 488-560), its
// compilation into the straight-line program the kernels of generic.cu / lookup.cu interpret (the role of
// ExpressionRegistry + Calculation, pb/util/expression/evaluator.rs:22-228), and HyperPlonk's `compose`
// (pb/backend/hyperplonk/preprocessor.rs:25-170). Plain C++ (no device code): it runs inside the library on the
// host, and `b200_expression_compile` exposes it so that it is testable without a GPU.
//
// Wire format of an expression (prefix tokens, int32): 0 Constant(const_idx) | 1 Identity | 2 Lagrange(i) |
// 3 EqXY(idx) | 4 Polynomial(poly, rotation) | 5 Challenge(idx) | 6 Negated e | 7 Sum a b | 8 Product a b |
// 9 Scaled(const_idx) e | 10 DistributePowers(n) e_1 .. e_n base.
#pragma once
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <tuple>
#include <utility>
#include <vector>

#include "ff32.cuh"

namespace b200 {

struct Expr;
typedef std::shared_ptr<const Expr> ExprP;
struct Expr {
  enum Kind { CONST, IDENTITY, LAGRANGE, EQXY, POLY, CHALLENGE, NEG, SUM, PROD, SCALED, DPOW } kind;
  Fr scalar;         // CONST, SCALED (Montgomery)
  int a = 0, b = 0;  // LAGRANGE i | EQXY idx | CHALLENGE idx | POLY (poly, rotation)
  std::vector<ExprP> ch;  // children; DPOW: terms..., base last
};

// ---- constructors / operators (expression.rs:67-106, 488-560) ----------------------------------------
inline ExprP e_node(Expr::Kind k, int a = 0, int b = 0) {
  auto e = std::make_shared<Expr>();
  e->kind = k;
  e->a = a;
  e->b = b;
  e->scalar = fe_zero<FrP>();
  return e;
}
inline ExprP e_const(const Fr& v) {
  auto e = std::make_shared<Expr>();
  e->kind = Expr::CONST;
  e->scalar = v;
  return e;
}
inline ExprP e_const_u64(uint64_t v) { return e_const(fe_from_u64<FrP>(v)); }
inline ExprP e_identity() { return e_node(Expr::IDENTITY); }
inline ExprP e_lagrange(int i) { return e_node(Expr::LAGRANGE, i); }
inline ExprP e_eq(int idx) { return e_node(Expr::EQXY, idx); }
inline ExprP e_poly(int poly, int rotation = 0) { return e_node(Expr::POLY, poly, rotation); }
inline ExprP e_chal(int idx) { return e_node(Expr::CHALLENGE, idx); }
inline ExprP e_unary(Expr::Kind k, const ExprP& x) {
  auto e = std::make_shared<Expr>();
  e->kind = k;
  e->scalar = fe_zero<FrP>();
  e->ch = {x};
  return e;
}
inline ExprP e_binary(Expr::Kind k, const ExprP& x, const ExprP& y) {
  auto e = std::make_shared<Expr>();
  e->kind = k;
  e->scalar = fe_zero<FrP>();
  e->ch = {x, y};
  return e;
}
inline ExprP operator-(const ExprP& x) { return e_unary(Expr::NEG, x); }
inline ExprP operator+(const ExprP& x, const ExprP& y) { return e_binary(Expr::SUM, x, y); }
inline ExprP operator-(const ExprP& x, const ExprP& y) { return e_binary(Expr::SUM, x, -y); }
inline ExprP operator*(const ExprP& x, const ExprP& y) { return e_binary(Expr::PROD, x, y); }
inline ExprP e_distribute_powers(const std::vector<ExprP>& terms, const ExprP& base) {  // expression.rs:93-106
  if (terms.size() == 1) return terms[0];
  auto e = std::make_shared<Expr>();
  e->kind = Expr::DPOW;
  e->scalar = fe_zero<FrP>();
  e->ch = terms;
  e->ch.push_back(base);
  return e;
}
inline ExprP e_product(const std::vector<ExprP>& xs) {
  ExprP acc = xs[0];
  for (size_t i = 1; i < xs.size(); ++i) acc = acc * xs[i];
  return acc;
}

// prefix tokens -> tree; returns null on a malformed stream
// Nesting deeper than E_MAX_DEPTH is rejected: every later pass (compile, degree, leaves, serialize) recurses as deep.
static const int E_MAX_DEPTH = 2048;
inline ExprP e_parse(const int32_t*& t, const int32_t* end, const Fr* consts, int nconsts, int depth = 0) {
  if (t >= end || depth > E_MAX_DEPTH) return nullptr;
  const int k = *t++;
  auto need = [&](int n) { return end - t >= n; };
  switch (k) {
    case Expr::CONST:
      if (!need(1) || *t < 0 || *t >= nconsts) return nullptr;
      return e_const(consts[*t++]);
    case Expr::IDENTITY: return e_identity();
    case Expr::LAGRANGE:
    case Expr::EQXY:
    case Expr::CHALLENGE: {
      if (!need(1)) return nullptr;
      const int a = *t++;
      return e_node((Expr::Kind)k, a);
    }
    case Expr::POLY: {
      if (!need(2)) return nullptr;
      const int a = *t++, b = *t++;
      return e_poly(a, b);
    }
    case Expr::NEG: {
      ExprP x = e_parse(t, end, consts, nconsts, depth + 1);
      return x ? e_unary(Expr::NEG, x) : nullptr;
    }
    case Expr::SUM:
    case Expr::PROD: {
      ExprP x = e_parse(t, end, consts, nconsts, depth + 1);
      ExprP y = x ? e_parse(t, end, consts, nconsts, depth + 1) : nullptr;
      return y ? e_binary((Expr::Kind)k, x, y) : nullptr;
    }
    case Expr::SCALED: {
      if (!need(1) || *t < 0 || *t >= nconsts) return nullptr;
      const Fr s = consts[*t++];
      ExprP x = e_parse(t, end, consts, nconsts, depth + 1);
      if (!x) return nullptr;
      auto e = std::make_shared<Expr>();
      e->kind = Expr::SCALED;
      e->scalar = s;
      e->ch = {x};
      return e;
    }
    case Expr::DPOW: {
      if (!need(1) || *t < 1) return nullptr;
      const int n = *t++;
      std::vector<ExprP> terms;
      for (int i = 0; i <= n; ++i) {
        ExprP x = e_parse(t, end, consts, nconsts, depth + 1);
        if (!x) return nullptr;
        terms.push_back(x);
      }
      ExprP base = terms.back();
      terms.pop_back();
      return e_distribute_powers(terms, base);
    }
  }
  return nullptr;
}

// tree -> prefix tokens + constants (the inverse of e_parse; constants are appended in visiting order)
inline void e_serialize(const ExprP& e, std::vector<int32_t>* tokens, std::vector<Fr>* consts) {
  tokens->push_back((int32_t)e->kind);
  switch (e->kind) {
    case Expr::CONST:
    case Expr::SCALED:
      tokens->push_back((int32_t)consts->size());
      consts->push_back(e->scalar);
      break;
    case Expr::LAGRANGE:
    case Expr::EQXY:
    case Expr::CHALLENGE: tokens->push_back(e->a); break;
    case Expr::POLY:
      tokens->push_back(e->a);
      tokens->push_back(e->b);
      break;
    case Expr::DPOW: tokens->push_back((int32_t)e->ch.size() - 1); break;
    default: break;
  }
  for (auto& c : e->ch) e_serialize(c, tokens, consts);
}

inline int e_degree(const ExprP& e) {  // expression.rs:171-182
  switch (e->kind) {
    case Expr::CONST:
    case Expr::CHALLENGE: return 0;
    case Expr::IDENTITY:
    case Expr::LAGRANGE:
    case Expr::EQXY:
    case Expr::POLY: return 1;
    case Expr::NEG:
    case Expr::SCALED: return e_degree(e->ch[0]);
    case Expr::SUM: return std::max(e_degree(e->ch[0]), e_degree(e->ch[1]));
    case Expr::PROD: return e_degree(e->ch[0]) + e_degree(e->ch[1]);
    case Expr::DPOW: {
      int d = 0;
      for (size_t i = 0; i + 1 < e->ch.size(); ++i) d = std::max(d, e_degree(e->ch[i]));
      return d + e_degree(e->ch.back());
    }
  }
  return 0;
}

// A leaf of the expression that becomes a dense device table.
struct Leaf {
  Expr::Kind kind;  // IDENTITY | LAGRANGE | EQXY | POLY
  int a, b;
  bool operator==(const Leaf& o) const { return kind == o.kind && a == o.a && b == o.b; }
};
inline void e_leaves(const ExprP& e, std::vector<Leaf>* out) {  // ordered, unique, in evaluation order
  switch (e->kind) {
    case Expr::IDENTITY:
    case Expr::LAGRANGE:
    case Expr::EQXY:
    case Expr::POLY: {
      const Leaf l{e->kind, e->a, e->b};
      for (auto& x : *out)
        if (x == l) return;
      out->push_back(l);
      return;
    }
    default:
      for (auto& c : e->ch) e_leaves(c, out);
  }
}

// ---- compilation -----------------------------------------------------------------------------------
// Slots: [0, K) leaf tables | [K, K+C) constants | temporaries. A constant is either a literal or the value of
// challenge `chal` (>= 0), which the caller copies device-to-device: challenges never visit the host.
struct ProgConst {
  Fr value;
  int chal;  // -1: literal
};
struct Program {
  std::vector<Leaf> leaves;
  std::vector<ProgConst> consts;
  std::vector<int32_t> ops;  // (opcode, dst, a, b) quadruples; opcode 0 add, 1 sub, 2 mul, 3 neg
  int ntemps = 0;
  int degree = 0;
};

class ExprCompiler {
 public:
  // refs order like the Python mirror's tuples ("const" < "leaf" < "op"), which fixes the canonical operand order
  typedef std::pair<int, int> Ref;  // (0 const | 1 leaf | 2 op, index)

  Program compile(const ExprP& root_expr) {
    Program p;
    e_leaves(root_expr, &p.leaves);
    leaves_ = &p.leaves;
    const Ref root = walk(root_expr);
    const int K = (int)p.leaves.size();
    int C = (int)consts_.size();
    auto slot = [&](const Ref& r) { return r.first == 1 ? r.second : (r.first == 0 ? K + r.second : K + C + r.second); };
    struct Op {
      int op, d, a, b;
    };
    std::vector<Op> prog;
    for (size_t i = 0; i < ops_.size(); ++i) prog.push_back({ops_[i].op, K + C + (int)i, slot(ops_[i].a), slot(ops_[i].b)});
    // drop ops that do not feed the root (e.g. the unused last power of a DistributePowers base)
    std::set<int> live;
    std::vector<int> need = {slot(root)};
    std::map<int, Op> by_dst;
    for (auto& o : prog) by_dst[o.d] = o;
    while (!need.empty()) {
      const int s = need.back();
      need.pop_back();
      auto it = by_dst.find(s);
      if (it != by_dst.end() && !live.count(s)) {
        live.insert(s);
        need.push_back(it->second.a);
        need.push_back(it->second.b);
      }
    }
    std::vector<Op> kept;
    for (auto& o : prog)
      if (live.count(o.d)) kept.push_back(o);
    if (kept.empty()) {  // a single leaf / constant: copy it through an addition with zero
      const Ref z = const_slot(fe_zero<FrP>(), -1);
      C = (int)consts_.size();
      kept.push_back({0, K + C, root.first == 0 ? K + root.second : slot(root), K + z.second});
    }
    // liveness-based reuse of temporary slots (the kernels keep the slot file in shared memory)
    std::map<int, int> last_use, mapping;
    for (size_t i = 0; i < kept.size(); ++i) {
      last_use[kept[i].a] = (int)i;
      last_use[kept[i].b] = (int)i;
    }
    std::vector<int> free_slots;
    int next_tmp = K + C;
    for (size_t i = 0; i < kept.size(); ++i) {
      const Op& o = kept[i];
      auto mapped = [&](int s) {
        auto it = mapping.find(s);
        return it == mapping.end() ? s : it->second;
      };
      const int ra = mapped(o.a), rb = mapped(o.b);
      std::set<int> srcs = {o.a, o.b};
      for (int s : srcs)
        if (s >= K + C && last_use[s] == (int)i) free_slots.push_back(mapping[s]);
      int rd;
      if (!free_slots.empty()) {
        rd = free_slots.back();
        free_slots.pop_back();
      } else {
        rd = next_tmp++;
      }
      mapping[o.d] = rd;
      p.ops.insert(p.ops.end(), {o.op, rd, ra, rb});
    }
    p.consts = consts_;
    p.ntemps = next_tmp - (K + C);
    p.degree = e_degree(root_expr);
    return p;
  }

 private:
  struct RawOp {
    int op;
    Ref a, b;
  };
  const std::vector<Leaf>* leaves_ = nullptr;
  std::vector<ProgConst> consts_;
  std::vector<RawOp> ops_;
  std::map<std::tuple<int, Ref, Ref>, int> op_memo_;

  Ref const_slot(const Fr& v, int chal) {
    for (size_t i = 0; i < consts_.size(); ++i) {
      if (chal >= 0 ? consts_[i].chal == chal : (consts_[i].chal < 0 && fe_eq<FrP>(consts_[i].value, v)))
        return {0, (int)i};
    }
    consts_.push_back({v, chal});
    return {0, (int)consts_.size() - 1};
  }
  Ref emit(int op, Ref a, Ref b, bool unary = false) {
    if ((op == 0 || op == 2) && !unary && b < a) std::swap(a, b);  // canonical operand order (evaluator.rs:185-189)
    if (unary) b = a;
    const auto key = std::make_tuple(op, a, b);
    auto it = op_memo_.find(key);
    if (it != op_memo_.end()) return {2, it->second};
    ops_.push_back({op, a, b});
    op_memo_[key] = (int)ops_.size() - 1;
    return {2, (int)ops_.size() - 1};
  }
  Ref walk(const ExprP& e) {
    switch (e->kind) {
      case Expr::CONST: return const_slot(e->scalar, -1);
      case Expr::CHALLENGE: return const_slot(fe_zero<FrP>(), e->a);
      case Expr::IDENTITY:
      case Expr::LAGRANGE:
      case Expr::EQXY:
      case Expr::POLY: {
        const Leaf l{e->kind, e->a, e->b};
        for (size_t i = 0; i < leaves_->size(); ++i)
          if ((*leaves_)[i] == l) return {1, (int)i};
        return {1, 0};
      }
      case Expr::NEG: return emit(3, walk(e->ch[0]), Ref(), true);
      case Expr::SUM: {
        if (e->ch[1]->kind == Expr::NEG) {
          const Ref a = walk(e->ch[0]);
          const Ref b = walk(e->ch[1]->ch[0]);
          return emit(1, a, b);
        }
        const Ref a = walk(e->ch[0]);
        const Ref b = walk(e->ch[1]);
        return emit(0, a, b);
      }
      case Expr::PROD: {
        const Ref a = walk(e->ch[0]);
        const Ref b = walk(e->ch[1]);
        return emit(2, a, b);
      }
      case Expr::SCALED: {
        const Ref a = walk(e->ch[0]);
        return emit(2, a, const_slot(e->scalar, -1));
      }
      case Expr::DPOW: {  // acc = e_0 + base e_1 + base^2 e_2 + ...   (expression.rs:150-166)
        const size_t n = e->ch.size() - 1;
        const Ref base = walk(e->ch[n]);
        Ref acc = walk(e->ch[0]), pw = base;
        for (size_t i = 1; i < n; ++i) {
          const Ref term = walk(e->ch[i]);
          acc = emit(0, acc, emit(2, pw, term));
          pw = emit(2, pw, base);
        }
        return acc;
      }
    }
    return Ref();
  }
};

// ---- HyperPlonk compose (preprocessor.rs:25-170) ------------------------------------------------------------
typedef std::vector<std::pair<ExprP, ExprP>> LookupCols;  // (input, table) column pairs of one lookup

// preprocessor.rs:78-109: per lookup  h (input+γ)(table+γ) - (table+γ) + m (input+γ);  plus Σ_b h(b) = 0
inline void e_lookup_constraints(const std::vector<LookupCols>& lookups, int num_poly, int num_permutation_polys,
                                 const ExprP& beta, const ExprP& gamma, std::vector<ExprP>* constraints,
                                 std::vector<ExprP>* sum_checks) {
  const int m_off = num_poly + num_permutation_polys, h_off = m_off + (int)lookups.size();
  for (size_t i = 0; i < lookups.size(); ++i) {
    const ExprP m = e_poly(m_off + (int)i), h = e_poly(h_off + (int)i);
    std::vector<ExprP> ins, tabs;
    for (auto& col : lookups[i]) {
      ins.push_back(col.first);
      tabs.push_back(col.second);
    }
    const ExprP inp = e_distribute_powers(ins, beta), tab = e_distribute_powers(tabs, beta);
    constraints->push_back(h * (inp + gamma) * (tab + gamma) - (tab + gamma) + m * (inp + gamma));
    sum_checks->push_back(h);
  }
}

// preprocessor.rs:111-170
inline int e_permutation_constraints(int num_vars, int num_poly, const std::vector<int>& permutation_polys,
                                     int max_degree, const ExprP& beta, const ExprP& gamma,
                                     int num_builtin_witness_polys, std::vector<ExprP>* out) {
  const int chunk = max_degree - 1, np = (int)permutation_polys.size();
  const int nchunks = (np + chunk - 1) / chunk;
  const int perm_off = num_poly, z_off = perm_off + np + num_builtin_witness_polys;
  if (nchunks == 0) return 0;
  std::vector<ExprP> polys, ids, perms, zs;
  for (int i = 0; i < np; ++i) {
    polys.push_back(e_poly(permutation_polys[i]));
    ids.push_back(e_const_u64((uint64_t)i << num_vars) + e_identity());
    perms.push_back(e_poly(perm_off + i));
  }
  for (int c = 0; c < nchunks; ++c) zs.push_back(e_poly(z_off + c));
  const ExprP z0_next = e_poly(z_off, 1), one = e_const_u64(1);
  out->push_back(e_lagrange(1) * (zs[0] - one));
  for (int c = 0; c < nchunks; ++c) {
    const int lo = c * chunk, hi = std::min(np, lo + chunk);
    std::vector<ExprP> l, r;
    for (int i = lo; i < hi; ++i) {
      l.push_back(polys[i] + beta * ids[i] + gamma);
      r.push_back(polys[i] + beta * perms[i] + gamma);
    }
    const ExprP z_l = zs[c], z_r = c + 1 < nchunks ? zs[c + 1] : z0_next;
    out->push_back(z_l * e_product(l) - z_r * e_product(r));
  }
  return nchunks;
}

// preprocessor.rs:25-60: (num_permutation_z_polys, zero-check expression); challenges beta, gamma, alpha follow the
// circuit's own `num_challenges`.
inline ExprP e_compose(int num_vars, const std::vector<ExprP>& constraints, int num_poly,
                       const std::vector<int>& permutation_polys, int num_challenges, int max_degree,
                       const std::vector<LookupCols>& lookups, int* num_z, int* chunk_size = nullptr) {
  const ExprP beta = e_chal(num_challenges), gamma = e_chal(num_challenges + 1), alpha = e_chal(num_challenges + 2);
  std::vector<ExprP> lookup_cons, lookup_sums;
  e_lookup_constraints(lookups, num_poly, (int)permutation_polys.size(), beta, gamma, &lookup_cons, &lookup_sums);
  int md = std::max(max_degree, 2);
  for (auto& c : constraints) md = std::max(md, e_degree(c));
  for (auto& c : lookup_cons) md = std::max(md, e_degree(c));
  std::vector<ExprP> perm;
  if (chunk_size) *chunk_size = md - 1;
  *num_z = e_permutation_constraints(num_vars, num_poly, permutation_polys, md, beta, gamma, 2 * (int)lookups.size(), &perm);
  std::vector<ExprP> all(constraints);
  all.insert(all.end(), lookup_cons.begin(), lookup_cons.end());
  all.insert(all.end(), perm.begin(), perm.end());
  const ExprP on_every_row = e_distribute_powers(all, alpha) * e_eq(0);
  lookup_sums.push_back(on_every_row);
  return e_distribute_powers(lookup_sums, alpha);
}

// BooleanHypercube (pb/util/arithmetic/bh.rs:76-153): primitive polynomials and the i-th row of the LFSR order
static const uint32_t BH_PRIMITIVE[32] = {1,        3,        7,         11,        19,        37,         67,        131,
                                          285,      529,      1033,      2053,      4179,      8219,       16427,     32771,
                                          65581,    131081,   262183,    524327,    1048585,   2097157,    4194307,   8388641,
                                          16777243, 33554441, 67108935,  134217767, 268435465, 536870917,  1073741907, 2147483657u};
inline uint64_t bh_next(uint64_t b, int num_vars) {
  b <<= 1;
  return b ^ ((b >> num_vars) * BH_PRIMITIVE[num_vars]);
}
inline uint64_t bh_nth(int num_vars, long i) {  // iter(): 0, then 1, x, x^2, ...; i taken modulo 2^n as rem_euclid
  const long N = 1L << num_vars;
  i = ((i % N) + N) % N;
  if (i == 0) return 0;
  uint64_t b = 1;
  for (long k = 1; k < i; ++k) b = bh_next(b, num_vars);
  return b;
}

}  // namespace b200

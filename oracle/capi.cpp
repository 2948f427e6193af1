// ORACLE (test infrastructure only — never linked into the product path).
// Flat C entry points over the CPU restatement so tests/ and bench.py's cpu_baseline leg can drive
// it through ctypes. Field elements cross as 32-byte Montgomery LE limbs, G1 points as 64-byte
// (x, y) Montgomery — the halo2curves in-memory layout.
#include <chrono>
#include <cstdio>

#include "expression.hpp"
#include "gkr.hpp"
#include "hyperplonk.hpp"
#include "lasso.hpp"

using namespace oracle;

// Documented synthetic-input PRNG: stateless splitmix64.
static inline uint64_t sm64(uint64_t seed, uint64_t i) {
  uint64_t z = seed + (i + 1) * 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

extern "C" {

int orc_num_threads() { return omp_get_max_threads(); }
void orc_set_num_threads(int n) { omp_set_num_threads(n); }

uint64_t orc_rand_u64(uint64_t seed, uint64_t i) { return sm64(seed, i); }
void orc_rand_u64s(uint64_t seed, uint64_t n, uint64_t* out) {
  for (uint64_t i = 0; i < n; ++i) out[i] = sm64(seed, i);
}
// element i = limbs sm64(seed, 4i..4i+3), top limb masked to 61 bits (value < 2^253 < r), Montgomery
void orc_rand_fr(uint64_t seed, uint64_t n, Fr* out) {
#pragma omp parallel for
  for (long i = 0; i < (long)n; ++i) {
    uint64_t raw[4] = {sm64(seed, 4 * i), sm64(seed, 4 * i + 1), sm64(seed, 4 * i + 2),
                       sm64(seed, 4 * i + 3) & 0x1fffffffffffffffULL};
    out[i] = Fr::from_raw(raw);
  }
}

// ---- field ------------------------------------------------------------------------------------
#define FIELD_API(NAME, F)                                                                   \
  void orc_##NAME##_mul(const F* a, const F* b, F* o, uint64_t n) {                          \
    for (uint64_t i = 0; i < n; ++i) o[i] = a[i] * b[i];                                     \
  }                                                                                          \
  void orc_##NAME##_add(const F* a, const F* b, F* o, uint64_t n) {                          \
    for (uint64_t i = 0; i < n; ++i) o[i] = a[i] + b[i];                                     \
  }                                                                                          \
  void orc_##NAME##_sub(const F* a, const F* b, F* o, uint64_t n) {                          \
    for (uint64_t i = 0; i < n; ++i) o[i] = a[i] - b[i];                                     \
  }                                                                                          \
  void orc_##NAME##_inv(const F* a, F* o, uint64_t n) {                                      \
    for (uint64_t i = 0; i < n; ++i) o[i] = a[i].inv();                                      \
  }                                                                                          \
  void orc_##NAME##_from_raw(const uint64_t* raw, F* o, uint64_t n) {                        \
    for (uint64_t i = 0; i < n; ++i) o[i] = F::from_raw(raw + 4 * i);                        \
  }                                                                                          \
  void orc_##NAME##_to_raw(const F* a, uint64_t* raw, uint64_t n) {                          \
    for (uint64_t i = 0; i < n; ++i) a[i].to_raw(raw + 4 * i);                               \
  }
FIELD_API(fr, Fr)
FIELD_API(fq, Fq)

// ---- keccak / transcript -----------------------------------------------------------------------
void orc_keccak256(const uint8_t* data, uint64_t n, uint8_t pad, uint8_t out[32]) {
  Keccak256 k(pad);
  k.update(data, n);
  k.finalize_reset(out);
}
void* orc_tr_new() { return new Transcript(); }
void* orc_tr_from_proof(const uint8_t* proof, uint64_t n) {
  return new Transcript(std::vector<uint8_t>(proof, proof + n));
}
void orc_tr_free(void* h) { delete (Transcript*)h; }
uint64_t orc_tr_proof_len(void* h) { return ((Transcript*)h)->stream.size(); }
void orc_tr_proof(void* h, uint8_t* out) {
  auto& s = ((Transcript*)h)->stream;
  memcpy(out, s.data(), s.size());
}
void orc_tr_common_fe(void* h, const Fr* fe) { ((Transcript*)h)->common_field_element(*fe); }
void orc_tr_write_fe(void* h, const Fr* fe) { ((Transcript*)h)->write_field_element(*fe); }
int orc_tr_read_fe(void* h, Fr* fe) { return ((Transcript*)h)->read_field_element(fe) ? 0 : 1; }
void orc_tr_squeeze(void* h, Fr* out) { *out = ((Transcript*)h)->squeeze_challenge(); }
int orc_tr_write_comm(void* h, const G1Affine* p) { return ((Transcript*)h)->write_commitment(*p) ? 0 : 1; }
int orc_tr_read_comm(void* h, G1Affine* p) { return ((Transcript*)h)->read_commitment(p) ? 0 : 1; }

// ---- G1 ----------------------------------------------------------------------------------------
void orc_g1_generator(G1Affine* out) { *out = G1Affine::generator(); }
void orc_g1_mul(const G1Affine* p, const Fr* k, G1Affine* out) { *out = G1::from_affine(*p).mul(*k).to_affine(); }
void orc_g1_add(const G1Affine* a, const G1Affine* b, G1Affine* out) {
  *out = G1::from_affine(*a).add_affine(*b).to_affine();
}
int orc_g1_on_curve(const G1Affine* p) { return p->on_curve() ? 1 : 0; }

// ---- MLE ---------------------------------------------------------------------------------------
void orc_eq_xy(const Fr* y, int n, Fr* out) {
  Poly e = eq_xy(std::vector<Fr>(y, y + n));
  memcpy(out, e.data(), e.size() * sizeof(Fr));
}
void orc_evaluate(const Fr* p, int nv, const Fr* x, Fr* out) {
  *out = evaluate(Poly(p, p + ((size_t)1 << nv)), std::vector<Fr>(x, x + nv));
}
void orc_fix_var(const Fr* p, int nv, const Fr* r, Fr* out) {
  Poly o = fix_var(Poly(p, p + ((size_t)1 << nv)), *r);
  memcpy(out, o.data(), o.size() * sizeof(Fr));
}
void orc_eq_xy_eval(const Fr* x, const Fr* y, int n, Fr* out) {
  *out = eq_xy_eval(std::vector<Fr>(x, x + n), std::vector<Fr>(y, y + n));
}

// ---- sum-check ---------------------------------------------------------------------------------
// EVAL shape. polys: npolys pointers to 2^num_vars tables. term k multiplies coeffs[k] by the
// tables idx[off[k] .. off[k+1]). Returns challenges (num_vars) and evals (npolys).
void orc_sumcheck_prove_evals(void* tr, int num_vars, int npolys, const Fr* const* polys, int has_eq,
                              const Fr* y, int nterms, const Fr* coeffs, const int* off,
                              const int* idx, const Fr* sum, Fr* challenges, Fr* evals) {
  std::vector<Poly> tabs(npolys);
  VirtualPoly vp;
  for (int i = 0; i < npolys; ++i) tabs[i].assign(polys[i], polys[i] + ((size_t)1 << num_vars));
  for (int i = 0; i < npolys; ++i) vp.polys.push_back(&tabs[i]);
  vp.has_eq = has_eq != 0;
  if (has_eq) vp.y.assign(y, y + num_vars);
  for (int k = 0; k < nterms; ++k) vp.terms.push_back(Term{coeffs[k], std::vector<int>(idx + off[k], idx + off[k + 1])});
  SumCheckOutput o = sumcheck_prove_evals(num_vars, vp, *sum, *(Transcript*)tr);
  memcpy(challenges, o.challenges.data(), num_vars * sizeof(Fr));
  memcpy(evals, o.evals.data(), npolys * sizeof(Fr));
}
// COEFF shape: Σ_k scalars[k] * eq(x, ys[k]) * polys[poly_idx[k]]
void orc_sumcheck_prove_coeffs(void* tr, int num_vars, int npolys, const Fr* const* polys, int nprods,
                               const Fr* scalars, const Fr* ys, const int* poly_idx, const Fr* sum,
                               Fr* challenges, Fr* evals) {
  std::vector<Poly> tabs(npolys);
  std::vector<const Poly*> ptrs;
  for (int i = 0; i < npolys; ++i) tabs[i].assign(polys[i], polys[i] + ((size_t)1 << num_vars));
  for (int i = 0; i < npolys; ++i) ptrs.push_back(&tabs[i]);
  std::vector<CoeffProduct> prods;
  for (int k = 0; k < nprods; ++k)
    prods.push_back(CoeffProduct{scalars[k], std::vector<Fr>(ys + k * num_vars, ys + (k + 1) * num_vars), poly_idx[k]});
  SumCheckOutput o = sumcheck_prove_coeffs(num_vars, prods, ptrs, *sum, *(Transcript*)tr);
  memcpy(challenges, o.challenges.data(), num_vars * sizeof(Fr));
  memcpy(evals, o.evals.data(), npolys * sizeof(Fr));
}
int orc_sumcheck_verify(void* tr, int num_vars, int degree, const Fr* sum, int coeffs, Fr* final_claim,
                        Fr* challenges) {
  std::vector<Fr> ch;
  if (!sumcheck_verify(num_vars, degree, *sum, coeffs != 0, *(Transcript*)tr, final_claim, &ch)) return 1;
  memcpy(challenges, ch.data(), num_vars * sizeof(Fr));
  return 0;
}
// generic expression (prefix token stream, see oracle.py serialize_expression)
void orc_sumcheck_prove_generic(void* tr, int num_vars, const int* tokens, const Fr* consts, int npolys,
                                const Fr* const* polys, const Fr* challenges, int nchallenges, const Fr* ys, int nys,
                                const Fr* sum, Fr* out_challenges, Fr* out_evals, int* out_degree) {
  const int* t = tokens;
  ExprP e = parse_expr(t, consts);
  std::vector<Poly> tabs(npolys);
  std::vector<const Poly*> ptrs;
  for (int i = 0; i < npolys; ++i) tabs[i].assign(polys[i], polys[i] + ((size_t)1 << num_vars));
  for (int i = 0; i < npolys; ++i) ptrs.push_back(&tabs[i]);
  std::vector<std::vector<Fr>> yv;
  for (int k = 0; k < nys; ++k) yv.push_back(std::vector<Fr>(ys + k * num_vars, ys + (k + 1) * num_vars));
  *out_degree = expr_degree(e);
  SumCheckOutput o = sumcheck_prove_generic(num_vars, e, ptrs, std::vector<Fr>(challenges, challenges + nchallenges), yv,
                                            *sum, *(Transcript*)tr);
  memcpy(out_challenges, o.challenges.data(), num_vars * sizeof(Fr));
  memcpy(out_evals, o.evals.data(), npolys * sizeof(Fr));
}
uint64_t orc_bh_rotate(int num_vars, uint64_t b, int rotation) { return BooleanHypercube(num_vars).rotate(b, rotation); }
void orc_bh_iter(int num_vars, uint64_t* out) {
  auto v = BooleanHypercube(num_vars).iter();
  memcpy(out, v.data(), v.size() * 8);
}
// cfg2 claim: Σ_b eq(b,y) a(b) b(b)
void orc_sum_eq_ab(int num_vars, const Fr* y, const Fr* a, const Fr* b, Fr* out) {
  Poly e = eq_xy(std::vector<Fr>(y, y + num_vars));
  Fr acc = Fr::zero();
  for (size_t i = 0; i < e.size(); ++i) acc = acc + e[i] * a[i] * b[i];
  *out = acc;
}

// ---- MSM / KZG ---------------------------------------------------------------------------------
void orc_msm(const Fr* scalars, const G1Affine* bases, uint64_t n, G1Affine* out) {
  *out = variable_base_msm(scalars, bases, n).to_affine();
}
void* orc_kzg_setup(const Fr* ss, int n) { return new KzgParams(kzg_setup(std::vector<Fr>(ss, ss + n))); }
// import an SRS produced elsewhere (e.g. downloaded from the GPU): eqs_flat = eqs[0] | eqs[1] | ... | eqs[n]
void* orc_kzg_import(const Fr* ss, int n, const G1Affine* eqs_flat) {
  KzgParams* p = new KzgParams();
  p->num_vars = n;
  p->ss.assign(ss, ss + n);
  p->eqs.resize(n + 1);
  size_t off = 0;
  for (int k = 0; k <= n; ++k) {
    p->eqs[k].assign(eqs_flat + off, eqs_flat + off + ((size_t)1 << k));
    off += (size_t)1 << k;
  }
  return p;
}
void orc_kzg_free(void* h) { delete (KzgParams*)h; }
void orc_kzg_eqs(void* h, int level, G1Affine* out) {
  auto& e = ((KzgParams*)h)->eqs[level];
  memcpy(out, e.data(), e.size() * sizeof(G1Affine));
}
void orc_kzg_commit(void* h, const Fr* poly, int nv, G1Affine* out) {
  *out = kzg_commit(*(KzgParams*)h, Poly(poly, poly + ((size_t)1 << nv)));
}
int orc_kzg_open(void* h, void* tr, const Fr* poly, int nv, const Fr* point, Fr* eval) {
  bool ok;
  *eval = kzg_open(*(KzgParams*)h, Poly(poly, poly + ((size_t)1 << nv)), std::vector<Fr>(point, point + nv),
                   *(Transcript*)tr, &ok);
  return ok ? 0 : 1;
}
int orc_kzg_verify(void* h, void* tr, const G1Affine* comm, int nv, const Fr* point, const Fr* eval) {
  return kzg_verify(*(KzgParams*)h, *comm, std::vector<Fr>(point, point + nv), *eval, *(Transcript*)tr) ? 0 : 1;
}
static void unpack_batch(int nv, int npoints, const Fr* points, int nevals, const int* ev_poly,
                         const int* ev_point, const Fr* ev_value, std::vector<std::vector<Fr>>* pts,
                         std::vector<Evaluation>* evs) {
  for (int i = 0; i < npoints; ++i) pts->push_back(std::vector<Fr>(points + i * nv, points + (i + 1) * nv));
  for (int k = 0; k < nevals; ++k) evs->push_back(Evaluation{ev_poly[k], ev_point[k], ev_value[k]});
}
int orc_kzg_batch_open(void* h, void* tr, int nv, int npolys, const Fr* const* polys, int npoints,
                       const Fr* points, int nevals, const int* ev_poly, const int* ev_point,
                       const Fr* ev_value) {
  std::vector<Poly> tabs(npolys);
  std::vector<const Poly*> ptrs;
  for (int i = 0; i < npolys; ++i) tabs[i].assign(polys[i], polys[i] + ((size_t)1 << nv));
  for (int i = 0; i < npolys; ++i) ptrs.push_back(&tabs[i]);
  std::vector<std::vector<Fr>> pts;
  std::vector<Evaluation> evs;
  unpack_batch(nv, npoints, points, nevals, ev_poly, ev_point, ev_value, &pts, &evs);
  return kzg_batch_open(*(KzgParams*)h, nv, ptrs, pts, evs, *(Transcript*)tr) ? 0 : 1;
}
int orc_kzg_batch_verify(void* h, void* tr, int nv, int ncomms, const G1Affine* comms, int npoints,
                         const Fr* points, int nevals, const int* ev_poly, const int* ev_point,
                         const Fr* ev_value) {
  std::vector<std::vector<Fr>> pts;
  std::vector<Evaluation> evs;
  unpack_batch(nv, npoints, points, nevals, ev_poly, ev_point, ev_value, &pts, &evs);
  return kzg_batch_verify(*(KzgParams*)h, nv, std::vector<G1Affine>(comms, comms + ncomms), pts, evs,
                          *(Transcript*)tr) ? 0 : 1;
}

// ---- pairing (verifier side) -----------------------------------------------------------------------
void orc_kzg_set_pairing_check(void* kzg, int on) { ((KzgParams*)kzg)->pairing_check = on != 0; }
// e(a * G1, b * G2) as 12 x 4 canonical limbs (c0.a0.c0, c0.a0.c1, c0.a1.c0, ...)
static void fq12_out(const Fq12& f, uint64_t* out) {
  const Fq2* parts[6] = {&f.c0.a0, &f.c0.a1, &f.c0.a2, &f.c1.a0, &f.c1.a1, &f.c1.a2};
  for (int i = 0; i < 6; ++i) {
    parts[i]->c0.to_raw(out + 8 * i);
    parts[i]->c1.to_raw(out + 8 * i + 4);
  }
}
void orc_pairing_gen_multiples(const Fr* a, const Fr* b, uint64_t* out48) {
  const G1Affine p = G1::from_affine(G1Affine::generator()).mul(*a).to_affine();
  const G2Affine q = G2Affine::generator().mul(*b);
  fq12_out(pairing(p, q), out48);
}
// g^e for g = e(G1, G2), e a scalar
void orc_pairing_gen_pow(const Fr* e, uint64_t* out48) {
  const Fq12 g = pairing(G1Affine::generator(), G2Affine::generator());
  uint64_t k[4];
  e->to_raw(k);
  Fq12 acc = Fq12::one();
  for (int i = 255; i >= 0; --i) {
    acc = acc.sqr();
    if ((k[i >> 6] >> (i & 63)) & 1) acc = acc * g;
  }
  fq12_out(acc, out48);
}
int orc_g2_checks(const Fr* k) {  // generator on the twist, [k]G2 on the twist, [r]G2 = O (r = 0 as a scalar: use r-1 and add)
  const G2Affine g = G2Affine::generator();
  if (!g.on_curve()) return 1;
  const G2Affine kg = g.mul(*k);
  if (!kg.on_curve() || kg.inf) return 2;
  const Fr minus_one = Fr::zero() - Fr::one();
  if (!(g.mul(minus_one).add(g)).inf) return 3;  // [r-1]G + G = O
  if (!(g.mul(*k).add(g.mul(Fr::zero() - *k))).inf) return 4;
  return 0;
}
// Π e(a_i G1, b_i G2) == 1 ?
int orc_pairing_product_is_identity(const Fr* a, const Fr* b, int n) {
  std::vector<std::pair<G1Affine, G2Affine>> terms;
  for (int i = 0; i < n; ++i)
    terms.push_back({G1::from_affine(G1Affine::generator()).mul(a[i]).to_affine(), G2Affine::generator().mul(b[i])});
  return pairings_product_is_identity(terms) ? 1 : 0;
}

// ---- HyperPlonk --------------------------------------------------------------------------------------
// cycles_flat: [len, poly, row, poly, row, ..., len, ...]
// lookup_tokens: per lookup [width, input_0, table_0, input_1, table_1, ...] with every expression in prefix form
void* orc_hp_preprocess(void* kzg, int num_vars, const int* tokens, const Fr* consts, int num_instances,
                        int num_witness, int npre, const Fr* const* pre, int nperm, const int* perm_idx,
                        const int* cycles_flat, int ncycles, int num_z, int nlookups, const int* lookup_tokens,
                        const Fr* lookup_consts) {
  const int* t = tokens;
  ExprP e = parse_expr(t, consts);
  std::vector<std::vector<std::pair<ExprP, ExprP>>> lookups;
  const int* lt = lookup_tokens;
  for (int l = 0; l < nlookups; ++l) {
    const int width = *lt++;
    std::vector<std::pair<ExprP, ExprP>> cols;
    for (int j = 0; j < width; ++j) {
      ExprP in = parse_expr(lt, lookup_consts);
      ExprP tb = parse_expr(lt, lookup_consts);
      cols.push_back({in, tb});
    }
    lookups.push_back(cols);
  }
  std::vector<Poly> pre_polys(npre);
  for (int i = 0; i < npre; ++i) pre_polys[i].assign(pre[i], pre[i] + ((size_t)1 << num_vars));
  std::vector<std::vector<std::pair<int, int>>> cycles;
  const int* c = cycles_flat;
  for (int k = 0; k < ncycles; ++k) {
    int len = *c++;
    std::vector<std::pair<int, int>> cyc;
    for (int i = 0; i < len; ++i) {
      cyc.push_back({c[0], c[1]});
      c += 2;
    }
    cycles.push_back(cyc);
  }
  return new HyperPlonkParams(hyperplonk_preprocess(*(KzgParams*)kzg, num_vars, e, {num_instances}, num_witness, pre_polys,
                                                    std::vector<int>(perm_idx, perm_idx + nperm), cycles, num_z, lookups));
}
// LogUp helper polynomials on their own (kernel-level parity)
void orc_expression_rows(int num_vars, const int* tokens, const Fr* consts, int npolys, const Fr* const* polys,
                         const Fr* challenges, int nchal, Fr* out) {
  const int* t = tokens;
  ExprP e = parse_expr(t, consts);
  const size_t N = (size_t)1 << num_vars;
  std::vector<Poly> ps(npolys);
  std::vector<const Poly*> pp;
  for (int i = 0; i < npolys; ++i) ps[i].assign(polys[i], polys[i] + N);
  for (auto& p : ps) pp.push_back(&p);
  std::vector<Fr> ch(challenges, challenges + nchal);
  BooleanHypercube bh(num_vars);
  const std::vector<uint64_t> order = bh.iter();
  for (size_t b = 0; b < N; ++b) out[b] = expr_eval_row(e, b, bh, order, pp, ch);
}
int orc_lookup_m(int num_vars, const Fr* input, const Fr* table, Fr* m_out) {
  const size_t N = (size_t)1 << num_vars;
  std::array<Poly, 2> c = {Poly(input, input + N), Poly(table, table + N)};
  Poly m;
  if (!lookup_m_poly(c, &m)) return 1;
  memcpy(m_out, m.data(), N * sizeof(Fr));
  return 0;
}
void orc_lookup_h(int num_vars, const Fr* input, const Fr* table, const Fr* m, const Fr* gamma, Fr* h_out) {
  const size_t N = (size_t)1 << num_vars;
  std::array<Poly, 2> c = {Poly(input, input + N), Poly(table, table + N)};
  Poly h = lookup_h_poly(c, Poly(m, m + N), *gamma);
  memcpy(h_out, h.data(), N * sizeof(Fr));
}
void orc_hp_free(void* h) { delete (HyperPlonkParams*)h; }
void orc_hp_permutation_poly(void* h, int i, Fr* out) {
  auto& p = ((HyperPlonkParams*)h)->permutation_polys[i];
  memcpy(out, p.data(), p.size() * sizeof(Fr));
}
// PlonkishCircuitInfo::{num_instances, num_witness_polys, num_challenges} (pb/backend.rs:50-60) for circuits with
// several instance columns and / or several witness phases; without this call: one column, one phase, no challenges.
int orc_hp_set_phases(void* h, int ncols, const int* num_instances, int nphases, const int* num_witness, const int* num_challenges) {
  HyperPlonkParams* pp = (HyperPlonkParams*)h;
  int total = 0;
  for (int i = 0; i < nphases; ++i) total += num_witness[i];
  if (total != pp->num_witness_polys) return 1;
  pp->num_instances.assign(num_instances, num_instances + ncols);
  pp->phase_witness_polys.assign(num_witness, num_witness + nphases);
  pp->phase_challenges.assign(num_challenges, num_challenges + nphases);
  return 0;
}
// instance columns back to back
static std::vector<std::vector<Fr>> split_instances(const HyperPlonkParams& pp, const Fr* instances, int ninst) {
  std::vector<std::vector<Fr>> out;
  int off = 0;
  for (int n : pp.num_instances) {
    if (off + n > ninst) return {};
    out.emplace_back(instances + off, instances + off + n);
    off += n;
  }
  if (off != ninst) return {};
  return out;
}
int orc_hp_prove(void* h, void* tr, const Fr* instances, int ninst, const Fr* const* witness, int nwit) {
  HyperPlonkParams* pp = (HyperPlonkParams*)h;
  std::vector<Poly> wit(nwit);
  for (int i = 0; i < nwit; ++i) wit[i].assign(witness[i], witness[i] + ((size_t)1 << pp->num_vars));
  return hyperplonk_prove(*pp, split_instances(*pp, instances, ninst), wit, *(Transcript*)tr) ? 0 : 1;
}
// synthesize(user, round, challenges, nchallenges, out): fills out[i] (2^num_vars elements each, preallocated) with the
// witness polynomials of that phase; returns 0 on success.
typedef int (*orc_synthesize_fn)(void* user, int round, const Fr* challenges, int nchallenges, Fr* const* out);
int orc_hp_prove_phased(void* h, void* tr, const Fr* instances, int ninst, orc_synthesize_fn synth, void* user) {
  HyperPlonkParams* pp = (HyperPlonkParams*)h;
  const size_t N = (size_t)1 << pp->num_vars;
  bool synth_failed = false;
  Synthesize fn = [&](int round, const std::vector<Fr>& challenges) {
    const int nw = pp->phase_witness_polys.empty() ? pp->num_witness_polys : pp->phase_witness_polys[round];
    std::vector<Poly> polys(nw, Poly(N, Fr::zero()));
    std::vector<Fr*> ptrs;
    for (auto& p : polys) ptrs.push_back(p.data());
    if (synth(user, round, challenges.data(), (int)challenges.size(), ptrs.data()) != 0) {
      synth_failed = true;
      return std::vector<Poly>();
    }
    return polys;
  };
  const bool ok = hyperplonk_prove_phased(*pp, split_instances(*pp, instances, ninst), fn, *(Transcript*)tr);
  return ok && !synth_failed ? 0 : 1;
}
int orc_hp_verify(void* h, void* tr, const Fr* instances, int ninst) {
  HyperPlonkParams* pp = (HyperPlonkParams*)h;
  return hyperplonk_verify(*pp, split_instances(*pp, instances, ninst), *(Transcript*)tr) ? 0 : 1;
}
void orc_permutation_z(int num_vars, int nperm, const Fr* const* perm_polys, const Fr* const* wires, const Fr* beta,
                       const Fr* gamma, Fr* out) {
  const size_t N = (size_t)1 << num_vars;
  std::vector<Poly> perms(nperm), w(nperm);
  std::vector<const Poly*> polys;
  std::vector<int> idx;
  for (int i = 0; i < nperm; ++i) {
    perms[i].assign(perm_polys[i], perm_polys[i] + N);
    w[i].assign(wires[i], wires[i] + N);
  }
  for (int i = 0; i < nperm; ++i) {
    polys.push_back(&w[i]);
    idx.push_back(i);
  }
  Poly z = permutation_z_polys(1, idx, perms, polys, *beta, *gamma)[0];
  memcpy(out, z.data(), N * sizeof(Fr));
}

// ---- Lasso -------------------------------------------------------------------------------------
int orc_lasso_prove(void* kzg, void* tr, int kind, int chunks, int mu, const uint64_t* xs, const uint64_t* ys) {
  LassoTable tb{kind, chunks};
  return lasso_prove(*(KzgParams*)kzg, tb, mu, xs, ys, *(Transcript*)tr) ? 0 : 1;
}
int orc_lasso_verify(void* kzg, void* tr, int kind, int chunks, int mu) {
  LassoTable tb{kind, chunks};
  return lasso_verify(*(KzgParams*)kzg, tb, mu, *(Transcript*)tr) ? 0 : 1;
}
// TABLE_CUSTOM: the subtable crosses as data (2^16 u32 values). 2 = invalid descriptor / operand outside the table
static LassoTable custom_table(int chunks, int num_operands, int operand_bits, int out_bits, const uint32_t* values) {
  LassoTable tb{TABLE_CUSTOM, chunks};
  tb.num_operands = num_operands;
  tb.operand_bits = operand_bits;
  tb.custom_out_bits = out_bits;
  tb.values = values;
  return tb;
}
int orc_lasso_prove_custom(void* kzg, void* tr, int chunks, int num_operands, int operand_bits, int out_bits,
                           const uint32_t* values, int mu, const uint64_t* xs, const uint64_t* ys) {
  LassoTable tb = custom_table(chunks, num_operands, operand_bits, out_bits, values);
  if (!tb.valid() || (num_operands == 2 && !ys)) return 2;
  const int bits = operand_bits * chunks;
  for (size_t j = 0; j < ((size_t)1 << mu); ++j) {
    if (bits < 64 && ((xs[j] >> bits) || (num_operands == 2 && (ys[j] >> bits)))) return 2;
  }
  return lasso_prove(*(KzgParams*)kzg, tb, mu, xs, ys, *(Transcript*)tr) ? 0 : 1;
}
int orc_lasso_verify_custom(void* kzg, void* tr, int chunks, int num_operands, int operand_bits, int out_bits,
                            const uint32_t* values, int mu) {
  LassoTable tb = custom_table(chunks, num_operands, operand_bits, out_bits, values);
  if (!tb.valid()) return 2;
  return lasso_verify(*(KzgParams*)kzg, tb, mu, *(Transcript*)tr) ? 0 : 1;
}
// witness tables, flattened: a | dim[c] | e[c] | read_ts[c] (each 2^mu) then final_cts[c] (each 2^16)
void orc_lasso_witness(int kind, int chunks, int mu, const uint64_t* xs, const uint64_t* ys, Fr* mtabs, Fr* stabs) {
  LassoTable tb{kind, chunks};
  LassoWitness w = lasso_witness(tb, mu, xs, ys);
  const size_t m = (size_t)1 << mu, S = (size_t)1 << SUBTABLE_VARS;
  size_t k = 0;
  memcpy(mtabs + (k++) * m, w.a.data(), m * sizeof(Fr));
  for (auto& p : w.dim) memcpy(mtabs + (k++) * m, p.data(), m * sizeof(Fr));
  for (auto& p : w.e) memcpy(mtabs + (k++) * m, p.data(), m * sizeof(Fr));
  for (auto& p : w.read_ts) memcpy(mtabs + (k++) * m, p.data(), m * sizeof(Fr));
  for (int t = 0; t < chunks; ++t) memcpy(stabs + t * S, w.final_cts[t].data(), S * sizeof(Fr));
}
// grand product alone (for kernel-level parity): leaves = T tables of 2^h
void orc_grand_product_prove(void* tr, int T, int h, const Fr* const* leaves, Fr* claims, Fr* point) {
  std::vector<Poly> lv(T);
  for (int t = 0; t < T; ++t) lv[t].assign(leaves[t], leaves[t] + ((size_t)1 << h));
  GrandProductOutput o = grand_product_prove(lv, *(Transcript*)tr, nullptr);
  memcpy(claims, o.claims.data(), T * sizeof(Fr));
  memcpy(point, o.point.data(), h * sizeof(Fr));
}

// fractional_sum_check.rs: prove. claimed_mask bit b (p) / bit 16 + b (q): Some(claimed) -> absorbed, else written.
// outputs: p_xs[B], q_xs[B], x[n], p_0s[B], q_0s[B]
void orc_fractional_prove(void* tr, int B, int n, const Fr* const* ps, const Fr* const* qs, uint32_t claimed_mask,
                          Fr* p_xs, Fr* q_xs, Fr* x, Fr* p_0s, Fr* q_0s) {
  std::vector<Poly> P(B), Q(B);
  std::vector<const Poly*> pp(B), qp(B);
  for (int b = 0; b < B; ++b) {
    P[b].assign(ps[b], ps[b] + ((size_t)1 << n));
    Q[b].assign(qs[b], qs[b] + ((size_t)1 << n));
    pp[b] = &P[b];
    qp[b] = &Q[b];
  }
  const Fr dummy = Fr::zero();  // Some(_): only the presence matters on the prover side (sanity-check feature off)
  std::vector<const Fr*> cp(B), cq(B);
  for (int b = 0; b < B; ++b) {
    cp[b] = (claimed_mask >> b) & 1 ? &dummy : nullptr;
    cq[b] = (claimed_mask >> (16 + b)) & 1 ? &dummy : nullptr;
  }
  FractionalOutput o = fractional_sum_check_prove(cp, cq, pp, qp, *(Transcript*)tr);
  memcpy(p_xs, o.p_xs.data(), B * sizeof(Fr));
  memcpy(q_xs, o.q_xs.data(), B * sizeof(Fr));
  memcpy(x, o.x.data(), n * sizeof(Fr));
  memcpy(p_0s, o.p_0s.data(), B * sizeof(Fr));
  memcpy(q_0s, o.q_0s.data(), B * sizeof(Fr));
}
// verify: claimed values are read from claimed_p / claimed_q where the mask bit is set; 0 = accept
int orc_fractional_verify(void* tr, int B, int n, uint32_t claimed_mask, const Fr* claimed_p, const Fr* claimed_q,
                          Fr* p_xs, Fr* q_xs, Fr* x, Fr* p_0s, Fr* q_0s) {
  std::vector<const Fr*> cp(B), cq(B);
  for (int b = 0; b < B; ++b) {
    cp[b] = (claimed_mask >> b) & 1 ? claimed_p + b : nullptr;
    cq[b] = (claimed_mask >> (16 + b)) & 1 ? claimed_q + b : nullptr;
  }
  FractionalOutput o;
  if (!fractional_sum_check_verify(n, cp, cq, *(Transcript*)tr, &o)) return 1;
  memcpy(p_xs, o.p_xs.data(), B * sizeof(Fr));
  memcpy(q_xs, o.q_xs.data(), B * sizeof(Fr));
  memcpy(x, o.x.data(), n * sizeof(Fr));
  memcpy(p_0s, o.p_0s.data(), B * sizeof(Fr));
  memcpy(q_0s, o.q_0s.data(), B * sizeof(Fr));
  return 0;
}

}  // extern "C"

// extern "C" boundary of the CPU verifier (include/b200_verify.h). Every entry point validates its arguments, never
// throws, and maps a failed check to B200V_REJECT.
#include <cstring>
#include <new>

#include "../../include/b200_verify.h"
#include "verify.hpp"

using namespace b200v;

struct b200v_transcript {
  Transcript tr;
};
struct b200v_kzg {
  KzgVerifierParam vp;
};
struct b200v_hyperplonk {
  HyperPlonkVerifierParam vp;
};

namespace {

template <class F>
int guarded(F&& f) {
  try {
    return f();
  } catch (const std::bad_alloc&) {
    return B200V_ERR_ARG;
  } catch (...) {
    return B200V_REJECT;
  }
}
inline int verdict(bool ok) { return ok ? B200V_ACCEPT : B200V_REJECT; }

bool g1_ok(const G1Affine& p) { return p.on_curve(); }
// field elements cross as Montgomery limbs: every residue is below the modulus; anything else is a malformed argument
bool fr_ok(const void* fr, size_t n) {
  const Fr* f = (const Fr*)fr;
  for (size_t i = 0; i < n; ++i)
    if (Fr::geq_mod(f[i].l)) return false;
  return true;
}

// every Polynomial / Challenge / EqXY index and every rotation of the expression is usable by the verifier
bool expr_ok(const ExprP& e, int npolys, int nchallenges, int num_vars) {
  // (no abs(): abs(INT_MIN) is undefined and stays negative)
  if (e->kind == Expr::POLY && (e->a < 0 || e->a >= npolys || e->b < -num_vars || e->b > num_vars || e->b < -16 || e->b > 16))
    return false;
  if (e->kind == Expr::CHALLENGE && (e->a < 0 || e->a >= nchallenges)) return false;
  if (e->kind == Expr::EQXY && e->a != 0) return false;
  for (auto& c : e->ch)
    if (!expr_ok(c, npolys, nchallenges, num_vars)) return false;
  return true;
}

}  // namespace

extern "C" {

int b200v_transcript_new(const uint8_t* proof, uint64_t len, b200v_transcript** out) {
  if (!out || (!proof && len)) return B200V_ERR_ARG;
  return guarded([&] {
    *out = new b200v_transcript{Transcript(std::vector<uint8_t>(proof, proof + len))};
    return B200V_ACCEPT;
  });
}
void b200v_transcript_free(b200v_transcript* tr) { delete tr; }

int b200v_transcript_common_field_elements(b200v_transcript* tr, const void* fr, int n) {
  if (!tr || n < 0 || (!fr && n) || !fr_ok(fr, (size_t)n)) return B200V_ERR_ARG;
  for (int i = 0; i < n; ++i) tr->tr.common_field_element(((const Fr*)fr)[i]);
  return B200V_ACCEPT;
}
int b200v_transcript_read_field_elements(b200v_transcript* tr, void* fr_out, int n) {
  if (!tr || n < 0 || (!fr_out && n)) return B200V_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!tr->tr.read_field_element((Fr*)fr_out + i)) return B200V_REJECT;
  return B200V_ACCEPT;
}
int b200v_transcript_read_commitments(b200v_transcript* tr, void* g1_out, int n) {
  if (!tr || n < 0 || (!g1_out && n)) return B200V_ERR_ARG;
  for (int i = 0; i < n; ++i)
    if (!tr->tr.read_commitment((G1Affine*)g1_out + i)) return B200V_REJECT;
  return B200V_ACCEPT;
}
int b200v_transcript_squeeze_challenges(b200v_transcript* tr, void* fr_out, int n) {
  if (!tr || n < 0 || (!fr_out && n)) return B200V_ERR_ARG;
  for (int i = 0; i < n; ++i) ((Fr*)fr_out)[i] = tr->tr.squeeze_challenge();
  return B200V_ACCEPT;
}
int b200v_transcript_done(const b200v_transcript* tr) {
  if (!tr) return B200V_ERR_ARG;
  return verdict(tr->tr.rpos == tr->tr.stream.size());
}

int b200v_sumcheck_verify(b200v_transcript* tr, int num_vars, int degree, const void* sum_fr, int coefficients_form,
                          void* final_claim_out, void* challenges_out) {
  if (!tr || num_vars < 1 || num_vars > 40 || degree < 1 || degree > 32 || !sum_fr || !final_claim_out || !challenges_out ||
      !fr_ok(sum_fr, 1))
    return B200V_ERR_ARG;
  return guarded([&] {
    Fr fin;
    std::vector<Fr> x;
    if (!sumcheck_verify(num_vars, degree, *(const Fr*)sum_fr, coefficients_form != 0, tr->tr, &fin, &x)) return B200V_REJECT;
    *(Fr*)final_claim_out = fin;
    memcpy(challenges_out, x.data(), x.size() * sizeof(Fr));
    return B200V_ACCEPT;
  });
}

int b200v_fractional_sum_check_verify(b200v_transcript* tr, int num_batching, int num_vars, uint32_t claimed_mask,
                                      const void* claimed_p_fr, const void* claimed_q_fr, void* p_xs_out, void* q_xs_out,
                                      void* x_out, void* p_0s_out, void* q_0s_out) {
  const int B = num_batching;
  if (!tr || B < 1 || B > 16 || num_vars < 1 || num_vars > 40 || !p_xs_out || !q_xs_out || !x_out) return B200V_ERR_ARG;
  if ((claimed_mask & 0xffffu) >> B || (claimed_mask >> 16) >> B) return B200V_ERR_ARG;
  if (((claimed_mask & 0xffffu) && !claimed_p_fr) || ((claimed_mask >> 16) && !claimed_q_fr)) return B200V_ERR_ARG;
  for (int b = 0; b < B; ++b) {
    if (((claimed_mask >> b) & 1) && !fr_ok((const Fr*)claimed_p_fr + b, 1)) return B200V_ERR_ARG;
    if (((claimed_mask >> (16 + b)) & 1) && !fr_ok((const Fr*)claimed_q_fr + b, 1)) return B200V_ERR_ARG;
  }
  return guarded([&] {
    std::vector<const Fr*> cp(B), cq(B);
    for (int b = 0; b < B; ++b) {
      cp[b] = (claimed_mask >> b) & 1 ? (const Fr*)claimed_p_fr + b : nullptr;
      cq[b] = (claimed_mask >> (16 + b)) & 1 ? (const Fr*)claimed_q_fr + b : nullptr;
    }
    FractionalClaims o;
    if (!fractional_sum_check_verify(num_vars, cp, cq, tr->tr, &o)) return B200V_REJECT;
    memcpy(p_xs_out, o.p_xs.data(), B * sizeof(Fr));
    memcpy(q_xs_out, o.q_xs.data(), B * sizeof(Fr));
    memcpy(x_out, o.x.data(), num_vars * sizeof(Fr));
    if (p_0s_out) memcpy(p_0s_out, o.p_0s.data(), B * sizeof(Fr));
    if (q_0s_out) memcpy(q_0s_out, o.q_0s.data(), B * sizeof(Fr));
    return B200V_ACCEPT;
  });
}

int b200v_kzg_setup(const void* ss_fr, int num_vars, b200v_kzg** out) {
  if (!ss_fr || !out || num_vars < 1 || num_vars > 40 || !fr_ok(ss_fr, (size_t)num_vars)) return B200V_ERR_ARG;
  return guarded([&] {
    *out = new b200v_kzg{kzg_verifier_setup(std::vector<Fr>((const Fr*)ss_fr, (const Fr*)ss_fr + num_vars))};
    return B200V_ACCEPT;
  });
}
int b200v_kzg_import(const void* ss_g2, int num_vars, b200v_kzg** out) {
  if (!ss_g2 || !out || num_vars < 1 || num_vars > 40) return B200V_ERR_ARG;
  return guarded([&] {
    KzgVerifierParam vp;
    const Fq* w = (const Fq*)ss_g2;  // wire format: x.c0, x.c1, y.c0, y.c1 per point
    for (int i = 0; i < num_vars; ++i, w += 4) {
      const G2Affine p{Fq2{w[0], w[1]}, Fq2{w[2], w[3]}, false};
      if (!p.on_curve()) return B200V_ERR_ARG;
      vp.ss_g2.push_back(p);
    }
    *out = new b200v_kzg{vp};
    return B200V_ACCEPT;
  });
}
int b200v_kzg_export(const b200v_kzg* vp, void* ss_g2_out) {
  if (!vp || !ss_g2_out) return B200V_ERR_ARG;
  Fq* w = (Fq*)ss_g2_out;
  for (const G2Affine& p : vp->vp.ss_g2) {
    *w++ = p.x.c0;
    *w++ = p.x.c1;
    *w++ = p.y.c0;
    *w++ = p.y.c1;
  }
  return B200V_ACCEPT;
}
void b200v_kzg_free(b200v_kzg* vp) { delete vp; }

int b200v_kzg_verify(const b200v_kzg* vp, b200v_transcript* tr, const void* comm_g1, const void* point_fr, int num_vars,
                     const void* eval_fr) {
  if (!vp || !tr || !comm_g1 || !point_fr || !eval_fr || num_vars < 1 || num_vars > vp->vp.num_vars() ||
      !fr_ok(point_fr, (size_t)num_vars) || !fr_ok(eval_fr, 1))
    return B200V_ERR_ARG;
  return guarded([&] {
    const G1Affine c = *(const G1Affine*)comm_g1;
    if (!g1_ok(c)) return B200V_ERR_ARG;
    const std::vector<Fr> point((const Fr*)point_fr, (const Fr*)point_fr + num_vars);
    return verdict(kzg_verify(vp->vp, c, point, *(const Fr*)eval_fr, tr->tr));
  });
}

int b200v_kzg_batch_verify(const b200v_kzg* vp, b200v_transcript* tr, int num_vars, const void* comms_g1, int ncomms,
                           const void* points_fr, int npoints, const int32_t* ev_poly, const int32_t* ev_point,
                           const void* ev_values_fr, int nevals) {
  if (!vp || !tr || !comms_g1 || !points_fr || !ev_poly || !ev_point || !ev_values_fr || num_vars < 1 ||
      num_vars > vp->vp.num_vars() || ncomms < 1 || npoints < 1 || npoints > (1 << 16) || nevals < 2 || nevals > (1 << 20) ||
      !fr_ok(points_fr, (size_t)npoints * num_vars) || !fr_ok(ev_values_fr, (size_t)nevals))
    return B200V_ERR_ARG;
  return guarded([&] {
    std::vector<G1Affine> comms((const G1Affine*)comms_g1, (const G1Affine*)comms_g1 + ncomms);
    for (auto& c : comms)
      if (!g1_ok(c)) return B200V_ERR_ARG;
    std::vector<std::vector<Fr>> points(npoints);
    for (int i = 0; i < npoints; ++i)
      points[i].assign((const Fr*)points_fr + (size_t)i * num_vars, (const Fr*)points_fr + (size_t)(i + 1) * num_vars);
    std::vector<Evaluation> evals(nevals);
    for (int k = 0; k < nevals; ++k) {
      if (ev_poly[k] < 0 || ev_poly[k] >= ncomms || ev_point[k] < 0 || ev_point[k] >= npoints) return B200V_ERR_ARG;
      evals[k] = Evaluation{ev_poly[k], ev_point[k], ((const Fr*)ev_values_fr)[k]};
    }
    return verdict(kzg_batch_verify(vp->vp, num_vars, comms, points, evals, tr->tr));
  });
}

int b200v_lasso_verify(const b200v_kzg* vp, b200v_transcript* tr, int kind, int chunks, int mu) {
  return b200v_lasso_verify_statement(vp, tr, kind, chunks, mu, nullptr, nullptr, nullptr);
}

int b200v_lasso_verify_statement(const b200v_kzg* vp, b200v_transcript* tr, int kind, int chunks, int mu,
                                 const void* expect_a_g1, const void* expect_dims_g1, void* out_comms_g1) {
  if (!vp || !tr || kind < 0 || kind > 2 || chunks < 2 || chunks > 8 || (kind == TABLE_RANGE && chunks > 4) || mu < 1 ||
      mu > 30 || vp->vp.num_vars() < (mu > SUBTABLE_VARS ? mu : SUBTABLE_VARS))
    return B200V_ERR_ARG;
  LassoStatement stm;
  stm.expect_a = (const G1Affine*)expect_a_g1;
  stm.expect_dims = (const G1Affine*)expect_dims_g1;
  stm.out_comms = (G1Affine*)out_comms_g1;
  if (stm.expect_a && !g1_ok(*stm.expect_a)) return B200V_ERR_ARG;
  if (stm.expect_dims)
    for (int t = 0; t < chunks; ++t)
      if (!g1_ok(stm.expect_dims[t])) return B200V_ERR_ARG;
  return guarded([&] {
    LassoTable tb{kind, chunks};
    return verdict(lasso_verify(vp->vp, tb, mu, tr->tr, stm));
  });
}

int b200v_lasso_verify_table(const b200v_kzg* vp, b200v_transcript* tr, int chunks, int num_operands, int operand_bits,
                             int out_bits, const uint32_t* subtable, int mu, const void* expect_a_g1,
                             const void* expect_dims_g1, void* out_comms_g1) {
  if (!vp || !tr || !subtable || chunks < 2 || chunks > 8 || num_operands < 1 || num_operands > 2 || operand_bits < 1 ||
      num_operands * operand_bits > SUBTABLE_VARS || operand_bits * chunks > 64 || out_bits < 1 || out_bits > 32 || mu < 1 ||
      mu > 30 || vp->vp.num_vars() < (mu > SUBTABLE_VARS ? mu : SUBTABLE_VARS))
    return B200V_ERR_ARG;
  LassoStatement stm;
  stm.expect_a = (const G1Affine*)expect_a_g1;
  stm.expect_dims = (const G1Affine*)expect_dims_g1;
  stm.out_comms = (G1Affine*)out_comms_g1;
  if (stm.expect_a && !g1_ok(*stm.expect_a)) return B200V_ERR_ARG;
  if (stm.expect_dims)
    for (int t = 0; t < chunks; ++t)
      if (!g1_ok(stm.expect_dims[t])) return B200V_ERR_ARG;
  return guarded([&] {
    LassoTable tb{TABLE_CUSTOM, chunks};
    tb.num_operands = num_operands;
    tb.operand_bits = operand_bits;
    tb.custom_out_bits = out_bits;
    tb.values = subtable;
    return verdict(lasso_verify(vp->vp, tb, mu, tr->tr, stm));
  });
}

int b200v_hyperplonk_new(const b200v_kzg* vp, int k, int ninstance_cols, const int32_t* num_instances, int nphases,
                         const int32_t* num_witness_polys, const int32_t* num_challenges, int num_lookups,
                         int num_permutation_z_polys, const int32_t* expression_tokens, int ntokens,
                         const void* consts_fr, int nconsts, const void* preprocess_comms_g1, int npreprocess,
                         const void* permutation_comms_g1, int npermutation, b200v_hyperplonk** out) {
  if (!vp || !out || k < 1 || k > 30 || k > vp->vp.num_vars() || ninstance_cols < 0 || ninstance_cols > 64 ||
      (ninstance_cols && !num_instances) || nphases < 1 || nphases > 64 || !num_witness_polys || !num_challenges ||
      num_lookups < 0 || num_lookups > 64 || num_permutation_z_polys < 0 || num_permutation_z_polys > 64 ||
      !expression_tokens || ntokens < 1 || nconsts < 0 || (nconsts && !consts_fr) || npreprocess < 0 ||
      (npreprocess && !preprocess_comms_g1) || npermutation < 0 || (npermutation && !permutation_comms_g1) ||
      !fr_ok(consts_fr, (size_t)nconsts))
    return B200V_ERR_ARG;
  return guarded([&] {
    HyperPlonkVerifierParam hp;
    hp.kzg = vp->vp;
    hp.num_vars = k;
    int nwit = 0, nchal = 0;
    for (int i = 0; i < ninstance_cols; ++i) {
      if (num_instances[i] < 0 || ((size_t)num_instances[i] + 1) > ((size_t)1 << k)) return B200V_ERR_ARG;
      hp.num_instances.push_back(num_instances[i]);
    }
    for (int i = 0; i < nphases; ++i) {
      if (num_witness_polys[i] < 1 || num_challenges[i] < 0 || (i + 1 < nphases && num_challenges[i] == 0)) return B200V_ERR_ARG;
      hp.phase_witness_polys.push_back(num_witness_polys[i]);
      hp.phase_challenges.push_back(num_challenges[i]);
      nwit += num_witness_polys[i];
      nchal += num_challenges[i];
    }
    if (nwit > 4096 || nchal > 4096) return B200V_ERR_ARG;
    hp.num_lookups = num_lookups;
    hp.num_permutation_z_polys = num_permutation_z_polys;
    const int32_t* t = expression_tokens;
    hp.expression = parse_expr(t, expression_tokens + ntokens, (const Fr*)consts_fr, nconsts);
    if (!hp.expression || t != expression_tokens + ntokens) return B200V_ERR_ARG;
    const int npolys = ninstance_cols + npreprocess + nwit + npermutation + 2 * num_lookups + num_permutation_z_polys;
    if (!expr_ok(hp.expression, npolys, nchal + 3, k)) return B200V_ERR_ARG;
    const int degree = expr_degree(hp.expression);
    if (degree < 1 || degree > 32) return B200V_ERR_ARG;  // a zero check is eq * (...): at least degree 1
    hp.preprocess_comms.assign((const G1Affine*)preprocess_comms_g1, (const G1Affine*)preprocess_comms_g1 + npreprocess);
    hp.permutation_comms.assign((const G1Affine*)permutation_comms_g1, (const G1Affine*)permutation_comms_g1 + npermutation);
    for (auto* v : {&hp.preprocess_comms, &hp.permutation_comms})
      for (auto& c : *v)
        if (!g1_ok(c)) return B200V_ERR_ARG;
    *out = new b200v_hyperplonk{hp};
    return B200V_ACCEPT;
  });
}
void b200v_hyperplonk_free(b200v_hyperplonk* hp) { delete hp; }

int b200v_hyperplonk_verify(const b200v_hyperplonk* hp, b200v_transcript* tr, const void* instances_fr, int ninstances) {
  if (!hp || !tr || ninstances < 0 || (ninstances && !instances_fr) || !fr_ok(instances_fr, (size_t)ninstances))
    return B200V_ERR_ARG;
  return guarded([&] {
    std::vector<std::vector<Fr>> cols;
    int off = 0;
    for (int n : hp->vp.num_instances) {
      if (off + n > ninstances) return B200V_REJECT;  // hyperplonk.rs:299-305 Error::InvalidSnark
      cols.emplace_back((const Fr*)instances_fr + off, (const Fr*)instances_fr + off + n);
      off += n;
    }
    if (off != ninstances) return B200V_REJECT;
    return verdict(hyperplonk_verify(hp->vp, cols, tr->tr));
  });
}

}  // extern "C"

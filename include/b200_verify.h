/* b200_verify.h — C ABI of the CPU verifier (halo2-lasso_b200/libb200verify.so, host C++, no CUDA).
 *
 * The verifying halves of the reference's traits (the reference verifies on the CPU as well):
 *   FiatShamirTranscript (reading side)   pb/util/transcript.rs:99-238
 *   SumCheck::verify                      pb/piop/sum_check.rs:39-58, classic.rs:175-194, 242-263
 *   MultilinearKzg::verify / batch_verify pb/pcs/multilinear/kzg.rs:330-361, pb/pcs/multilinear.rs:237-275
 *   HyperPlonk::verify                    pb/backend/hyperplonk.rs:293-363, hyperplonk/verifier.rs:39-182
 * plus the verifier of the Lasso argument that b200_lasso_prove produces.
 *
 * Data crosses unconverted, as in b200_lasso.h: Fr / Fq = 32-byte little-endian Montgomery limbs, G1Affine = 64 bytes
 * (x, y), G2Affine = 128 bytes (x.c0, x.c1, y.c0, y.c1); proofs are the byte strings of b200_transcript_proof /
 * `Keccak256Transcript::into_proof`. All pointers are host pointers. Return value: B200V_ACCEPT, B200V_REJECT (the
 * reference's Err(InvalidSumcheck / InvalidPcsOpen / InvalidSnark / Transcript)) or B200V_ERR_ARG; nothing throws. */
#ifndef B200_VERIFY_H
#define B200_VERIFY_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define B200V_ACCEPT 0
#define B200V_REJECT 1
#define B200V_ERR_ARG 2

/* ---- transcript over a proof (transcript.rs:113-123 `from_proof`) ---------------------------------------- */
typedef struct b200v_transcript b200v_transcript;
int b200v_transcript_new(const uint8_t* proof, uint64_t len, b200v_transcript** out);
void b200v_transcript_free(b200v_transcript* tr);
int b200v_transcript_common_field_elements(b200v_transcript* tr, const void* fr, int n);       /* :133-136 */
int b200v_transcript_read_field_elements(b200v_transcript* tr, void* fr_out, int n);           /* :139-156; REJECT on a bad encoding */
int b200v_transcript_read_commitments(b200v_transcript* tr, void* g1_out, int n);              /* :186-210 */
int b200v_transcript_squeeze_challenges(b200v_transcript* tr, void* fr_out, int n);            /* :127-131 */
/* ACCEPT when every proof byte has been read */
int b200v_transcript_done(const b200v_transcript* tr);

/* ---- SumCheck::verify ---------------------------------------------------------------------------------- */
/* ClassicSumCheck::verify: reads num_vars messages of degree + 1 elements (coefficients_form = 0: EvaluationsProver
 * messages p(0..d), 1: CoefficientsProver), checks every round against the running claim and returns the final claim
 * and the challenge point x[num_vars]; the caller compares the final claim with its own evaluation at x. */
int b200v_sumcheck_verify(b200v_transcript* tr, int num_vars, int degree, const void* sum_fr, int coefficients_form,
                          void* final_claim_out, void* challenges_out);

/* ---- MultilinearKzg verifier ------------------------------------------------------------------------------ */
typedef struct b200v_kzg b200v_kzg; /* MultilinearKzgVerifierParam (kzg.rs:79-84): ss_g2[i] = g2 * s_i */
/* verify_fractional_sum_check (pb/piop/gkr/fractional_sum_check.rs:192-265): the GKR argument b200_fractional_sum_check_prove
 * writes, for num_batching (<= 16) pairs (p_b, q_b) of 2^num_vars-entry tables. Bit b / bit 16 + b of claimed_mask: the
 * layer-0 value p_b / q_b is a public claim taken from claimed_p_fr[b] / claimed_q_fr[b] (Some(_): absorbed), otherwise it
 * is read from the proof. ACCEPT returns the claims p_b(x), q_b(x) the caller must still check against its polynomials
 * (p_xs_out, q_xs_out: num_batching each; x_out: num_vars) and the layer-0 values (p_0s_out / q_0s_out, may be NULL;
 * sum_i p_b[i] / q_b[i] = p_0s[b] / q_0s[b]). */
int b200v_fractional_sum_check_verify(b200v_transcript* tr, int num_batching, int num_vars, uint32_t claimed_mask,
                                      const void* claimed_p_fr, const void* claimed_q_fr, void* p_xs_out, void* q_xs_out,
                                      void* x_out, void* p_0s_out, void* q_0s_out);

/* the verifier half of the seeded test setup b200_kzg_setup uses (kzg.rs:166-225) */
int b200v_kzg_setup(const void* ss_fr, int num_vars, b200v_kzg** out);
/* parameters from elsewhere: num_vars G2Affine points */
int b200v_kzg_import(const void* ss_g2, int num_vars, b200v_kzg** out);
int b200v_kzg_export(const b200v_kzg* vp, void* ss_g2_out);
void b200v_kzg_free(b200v_kzg* vp);
/* verify (kzg.rs:330-361): reads num_vars quotient commitments, pairing product check */
int b200v_kzg_verify(const b200v_kzg* vp, b200v_transcript* tr, const void* comm_g1, const void* point_fr, int num_vars,
                     const void* eval_fr);
/* batch_verify (pb/pcs/multilinear.rs:237-275); evaluation k = (ev_poly[k], ev_point[k], ev_values[k]) */
int b200v_kzg_batch_verify(const b200v_kzg* vp, b200v_transcript* tr, int num_vars, const void* comms_g1, int ncomms,
                           const void* points_fr, int npoints, const int32_t* ev_poly, const int32_t* ev_point,
                           const void* ev_values_fr, int nevals);

/* ---- Lasso ------------------------------------------------------------------------------------------------ */
/* the proof of b200_lasso_prove(ctx, kind, chunks, mu, ...): kind 0 range / 1 and / 2 xor. Consumes the whole Lasso
 * section; combine with b200v_transcript_done when nothing follows it. */
int b200v_lasso_verify(const b200v_kzg* vp, b200v_transcript* tr, int kind, int chunks, int mu);
/* The same check BOUND TO A STATEMENT. b200v_lasso_verify alone only establishes that some committed a decomposes into
 * table entries at some committed addresses — any prover can satisfy that. Linking the proof to the lookups the caller
 * cares about is mandatory: pass the expected commitment to a (the lookup outputs; NULL = not compared) and / or the
 * `chunks` expected commitments to dim_t (the chunked operands; NULL = not compared); a proof carrying other
 * commitments is REJECTED. out_comms (1 + 4 chunks points: a | dim | E | read_ts | final_cts; NULL = not wanted)
 * returns the commitments the proof carries so that an outer protocol can open / compare them itself. */
int b200v_lasso_verify_statement(const b200v_kzg* vp, b200v_transcript* tr, int kind, int chunks, int mu,
                                 const void* expect_a_g1, const void* expect_dims_g1, void* out_comms_g1);

/* The proof of b200_lasso_prove_table: the table is DATA (b200_lasso_table of b200_lasso.h — chunks, num_operands,
 * operand_bits, out_bits, 2^16 subtable values) and part of the statement: a proof for another table, operand layout or
 * output stride is REJECTED. The verifier evaluates the subtable's multilinear extension itself (2^16 products).
 * expect_a_g1 / expect_dims_g1 / out_comms_g1 as in b200v_lasso_verify_statement (NULL = not used). */
int b200v_lasso_verify_table(const b200v_kzg* vp, b200v_transcript* tr, int chunks, int num_operands, int operand_bits,
                             int out_bits, const uint32_t* subtable, int mu, const void* expect_a_g1,
                             const void* expect_dims_g1, void* out_comms_g1);

/* ---- HyperPlonk ---------------------------------------------------------------------------------------- */
typedef struct b200v_hyperplonk b200v_hyperplonk; /* HyperPlonkVerifierParam (hyperplonk.rs:58-74) */
/* expression: the composed zero-check expression (preprocessor.rs:25-60) in the prefix-token format of b200_lasso.h;
 * preprocess / permutation commitments as returned by b200_hyperplonk_commitments. The kzg parameters are borrowed. */
int b200v_hyperplonk_new(const b200v_kzg* vp, int k, int ninstance_cols, const int32_t* num_instances, int nphases,
                         const int32_t* num_witness_polys, const int32_t* num_challenges, int num_lookups,
                         int num_permutation_z_polys, const int32_t* expression_tokens, int ntokens,
                         const void* consts_fr, int nconsts, const void* preprocess_comms_g1, int npreprocess,
                         const void* permutation_comms_g1, int npermutation, b200v_hyperplonk** out);
void b200v_hyperplonk_free(b200v_hyperplonk* hp);
/* HyperPlonk::verify: instances = the instance columns back to back; consumes the HyperPlonk section of the proof */
int b200v_hyperplonk_verify(const b200v_hyperplonk* hp, b200v_transcript* tr, const void* instances_fr, int ninstances);

#ifdef __cplusplus
}
#endif
#endif

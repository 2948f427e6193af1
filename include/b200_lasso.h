/* libb200lasso — C ABI of the B200-native Lasso / HyperPlonk hot path.
 *
 * The reference (plonkish_backend) is 100% Rust with no FFI layer; its seams are the
 * generic traits cited next to each entry point below. A Rust shim binds these
 * symbols with `extern "C"` and implements `SumCheck`, `PolynomialCommitmentScheme` and the
 * transcript traits on top of them.
 *
 * Conventions
 *  - b200_fr  : 32 bytes, BN254 scalar, Montgomery form, little-endian u64 limbs — the in-memory
 *               layout of halo2_curves::bn256::Fr, so `&[Fr]` is passed as-is.
 *  - b200_g1  : 64 bytes, affine (x, y), each coordinate an Fq in the same Montgomery layout
 *               (halo2_curves::bn256::G1Affine); the identity is (0, 0).
 *  - "dev" pointers are device buffers owned by the library (b200_poly_*); "host" pointers are
 *    caller-owned. Nothing is retained after a call returns.
 *  - Every function returns 0 on success or a B200_ERR_* code; there are no panics/exceptions
 *    across the boundary. One host thread per context.
 *  - The Fiat-Shamir transcript (Keccak-256, pb/util/transcript.rs:99-238) lives ON THE DEVICE inside
 *    the context: provers append to it without host round trips; b200_transcript_* are the
 *    FieldTranscript / TranscriptWrite operations for the host-side caller.
 */
#ifndef B200_LASSO_H
#define B200_LASSO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200_OK 0
#define B200_ERR_CUDA 1       /* CUDA runtime failure (message on stderr) */
#define B200_ERR_ARG 2        /* invalid argument -> Error::InvalidPcsParam / InvalidSumcheck */
#define B200_ERR_TRANSCRIPT 3 /* identity commitment or proof overflow -> Error::Transcript */
#define B200_ERR_NOMEM 4
#define B200_ERR_LOOKUP 5     /* a lookup input is not in its table -> Error::InvalidSnark("Invalid lookup input") */
#define B200_ERR_PEER 6       /* a multi-GPU wait timed out: a peer left the collective */

typedef struct b200_ctx b200_ctx;

/* ---- context ----------------------------------------------------------------------------- */
int b200_ctx_create(int device, b200_ctx** out);
void b200_ctx_destroy(b200_ctx* ctx);
int b200_sync(b200_ctx* ctx);
/* kernels launched by this context since the last reset (bench "gpu_launches") */
uint64_t b200_launch_count(b200_ctx* ctx, int reset);
/* raw CUDA stream of the context (cudaStream_t), for event timing by the caller */
void* b200_stream(b200_ctx* ctx);
/* per-launch CUDA-event timing of the sum-check round kernels (the reference's `sum_check_prove_round-i`
 * timers, classic.rs:226): enable, run a prove, then read (ms[i], round tag[i]) pairs */
/* debug: clock64() stamps of single-CTA round kernels (needs env B200_DEBUG_CLOCKS=1 at ctx creation); out[32*16] */
int b200_debug_clocks(b200_ctx* ctx, long long* out);
int b200_profile_enable(b200_ctx* ctx, int on);
int b200_profile_read(b200_ctx* ctx, float* ms, int* tags, int cap, int* n);

/* ---- device polynomials: MultilinearPolynomial<Fr>::evals (pb/poly/multilinear.rs:20-24) ---- */
int b200_poly_alloc(b200_ctx* ctx, uint64_t len, void** dev);
int b200_poly_upload(b200_ctx* ctx, const void* host_fr, uint64_t len, void** dev);
int b200_poly_write(b200_ctx* ctx, void* dev, const void* host_fr, uint64_t len);
int b200_poly_download(b200_ctx* ctx, const void* dev, uint64_t len, void* host_fr);
int b200_poly_free(b200_ctx* ctx, void* dev);
/* canonical 256-bit integers <-> Montgomery residues, on the device (to_mont != 0: into Montgomery) */
int b200_fr_convert(b200_ctx* ctx, const void* dev_in, void* dev_out, uint64_t len, int to_mont);

/* ---- transcript: FiatShamirTranscript<Keccak256, _> (pb/util/transcript.rs:99-238) ---------- */
int b200_transcript_reset(b200_ctx* ctx);
/* FieldTranscript::common_field_elements (:133-136) */
int b200_transcript_common_field_elements(b200_ctx* ctx, const void* host_fr, int n);
/* FieldTranscriptWrite::write_field_elements (:158-165) */
int b200_transcript_write_field_elements(b200_ctx* ctx, const void* host_fr, int n);
/* FieldTranscript::squeeze_challenges (:127-131) */
int b200_transcript_squeeze_challenges(b200_ctx* ctx, int n, void* host_fr_out);
/* TranscriptWrite::write_commitments (:216-227); B200_ERR_TRANSCRIPT for the identity (:174-179) */
int b200_transcript_write_commitments(b200_ctx* ctx, const void* host_g1, int n);
/* InMemoryTranscript::into_proof (:112-114): copies the stream, *len = its length */
int b200_transcript_proof(b200_ctx* ctx, uint8_t* out, uint64_t cap, uint64_t* len);

/* ---- MultilinearPolynomial (pb/poly/multilinear.rs) ---------------------------------------- */
/* eq_xy (:91-127): dev_out[2^n] */
int b200_eq_xy(b200_ctx* ctx, const void* host_y, int n, void* dev_out);
/* fix_var (:179-183): dev_out[2^(n-1)] */
int b200_fix_var(b200_ctx* ctx, const void* dev_in, int n, const void* host_r, void* dev_out);
/* evaluate (:137-156) for `ntables` polynomials at one point */
int b200_evaluate(b200_ctx* ctx, const void* const* dev_tables, int ntables, int n, const void* host_x,
                  void* host_fr_out);

/* ---- SumCheck::prove (pb/piop/sum_check.rs:39-58), ClassicSumCheck (classic.rs:208-240) ------ */
/* EvaluationsProver (classic/eval.rs) for  F(x) = eq(x, y) * sum_t w[t] * prod_{k<np} P[t*np+k](x),
 * np in {1, 2}, degree np+1. Writes num_vars * (np+2) field elements to the context transcript and
 * squeezes num_vars challenges, exactly as the reference does for the same Expression.
 * Outputs: challenges[num_vars], evals[nterms*np] (= ProverState::into_evals). Tables untouched. */
int b200_sumcheck_prove_evals(b200_ctx* ctx, int num_vars, int nterms, int np,
                              const void* const* dev_tables, const void* host_weights,
                              const void* host_y, const void* host_sum, void* host_challenges_out,
                              void* host_evals_out);
/* same, tables in HOST memory (uploaded inside the call): the end-to-end entry a Rust caller holding
 * `Vec<Fr>` uses */
int b200_sumcheck_prove_evals_host(b200_ctx* ctx, int num_vars, int nterms, int np,
                                   const void* const* host_tables, const void* host_weights,
                                   const void* host_y, const void* host_sum,
                                   void* host_challenges_out, void* host_evals_out);
/* CoefficientsProver (classic/coeff.rs) for  F(x) = sum_k s[k] * eq(x, y_k) * P_k(x)  (degree 2).
 * host_ys: nprods * num_vars elements. */
int b200_sumcheck_prove_coeffs(b200_ctx* ctx, int num_vars, int nprods, const void* const* dev_tables,
                               const void* host_scalars, const void* host_ys, const void* host_sum,
                               void* host_challenges_out, void* host_evals_out);

/* EvaluationsProver for an ARBITRARY Expression (classic/eval.rs + util/expression/evaluator.rs): the host
 * compiles the expression into a straight-line program over slots [tables | constants | temporaries]
 * (ops = nops x {opcode 0 add / 1 sub / 2 mul / 3 neg, dst, a, b}); every leaf (polynomial query, rotated
 * query, eq_xy, identity, Lagrange) is passed as a dense device table. degree = Expression::degree().
 * evals_out[ntables] = every table bound at the challenges. */
int b200_sumcheck_prove_generic(b200_ctx* ctx, int num_vars, int degree, int ntables,
                                const void* const* dev_tables, int nconsts, const void* host_consts_fr, int nops,
                                const int32_t* host_ops, const void* host_sum, void* host_challenges_out,
                                void* host_evals_out);
/* leaf tables: identity polynomial b -> F::from(b); one-hot Lagrange table; rotated[b] = poly[bh.rotate(b, rot)]
 * (BooleanHypercube LFSR order, pb/util/arithmetic/bh.rs) */
int b200_poly_iota(b200_ctx* ctx, int num_vars, void* dev_out);
int b200_poly_onehot(b200_ctx* ctx, int num_vars, uint64_t index, void* dev_out);
int b200_poly_rotate(b200_ctx* ctx, const void* dev_in, int num_vars, int rotation, void* dev_out);

/* The same `SumCheck::prove`, with the expression itself crossing the boundary (what a `VirtualPolynomial` holds,
 * pb/piop/sum_check.rs:16-37) as prefix tokens (int32): 0 Constant(const_idx) | 1 Identity | 2 Lagrange(i) |
 * 3 EqXY(idx) | 4 Polynomial(poly, rotation) | 5 Challenge(idx) | 6 Negated e | 7 Sum a b | 8 Product a b |
 * 9 Scaled(const_idx) e | 10 DistributePowers(n) e_1..e_n base; consts_fr = Montgomery constants. The library
 * compiles it (the ExpressionRegistry role, pb/util/expression/evaluator.rs:22-228), materialises the leaf tables
 * and runs the generic round kernels. host_ys = nys points of num_vars elements; host_evals_out[npolys] = every
 * polynomial bound at the challenges (classic.rs:143-149). */
int b200_sumcheck_prove_expression(b200_ctx* ctx, int num_vars, const int32_t* tokens, int ntokens,
                                   const void* consts_fr, int nconsts, const void* const* dev_polys, int npolys,
                                   const void* host_challenges, int nchallenges, const void* host_ys, int nys,
                                   const void* host_sum, void* host_challenges_out, void* host_evals_out);
/* The compiler alone (host only, needs no GPU): leaves_out = (kind, a, b) triples in table order, consts_out /
 * const_chal_out = constant values and, where >= 0, the challenge index a constant stands for, ops_out =
 * (opcode, dst, a, b) quadruples over the slots [leaves | constants | temporaries]. */
int b200_expression_compile(const int32_t* tokens, int ntokens, const void* consts_fr, int nconsts, int32_t* leaves_out,
                            int leaves_cap, int* nleaves, void* consts_out, int32_t* const_chal_out, int consts_cap,
                            int* nconsts_out, int32_t* ops_out, int ops_cap, int* nops, int* ntemps, int* degree);

/* ---- HyperPlonk (pb/backend/hyperplonk.rs:97-291) ------------------------------------------------
 * preprocess: PlonkishCircuitInfo (pb/backend.rs:46-73) -> prover parameters. Polynomial order as in the reference:
 * instance (one polynomial) | preprocess | witness | permutation | lookup m | lookup h | permutation z.
 * dev_preprocess: device polynomials (borrowed: keep them alive); constraints: `nconstraints` expressions back to
 * back in the token format above; lookups: per lookup [width, input_0, table_0, input_1, table_1, ...];
 * cycles_flat: per copy cycle [len, poly, row, poly, row, ...] (preprocessor.rs:172-203). Commits the preprocess and
 * permutation polynomials with the SRS of the context and composes the zero-check expression (preprocessor.rs:25-170).
 * prove: instances as Montgomery field elements, witness polynomials on the device; appends the proof to the context
 * transcript. B200_ERR_LOOKUP = Error::InvalidSnark("Invalid lookup input"). */
/* compose (pb/backend/hyperplonk/preprocessor.rs:25-60) as a host-only service (no GPU needed): the circuit's
 * constraints / lookups / permutation columns (same token streams as b200_hyperplonk_preprocess; num_poly = instance +
 * preprocess + witness polynomials; num_challenges = the circuit's own) -> the composed zero-check expression as prefix
 * tokens + constants, and num_permutation_z_polys. This is the `expression` b200v_hyperplonk_new (b200_verify.h) and
 * b200_sumcheck_prove_expression take. B200_ERR_NOMEM when a capacity is too small (*ntokens / *nconsts_out then hold the
 * required sizes). */
int b200_expression_compose(int k, int num_poly, int num_challenges, int nconstraints, const int32_t* constraint_tokens,
                            int nconstraint_tokens, int nlookups, const int32_t* lookup_tokens, int nlookup_tokens,
                            const void* consts_fr, int nconsts, int nperm, const int32_t* permutation_polys,
                            int max_degree, int32_t* tokens_out, int tokens_cap, int* ntokens, void* consts_out,
                            int consts_cap, int* nconsts_out, int* num_permutation_z_polys);

typedef struct b200_hyperplonk b200_hyperplonk;
int b200_hyperplonk_preprocess(b200_ctx* ctx, int k, int num_instances, int num_witness_polys, int npreprocess,
                               const void* const* dev_preprocess, int nconstraints, const int32_t* constraint_tokens,
                               int nconstraint_tokens, int nlookups, const int32_t* lookup_tokens, int nlookup_tokens,
                               const void* consts_fr, int nconsts, int nperm, const int32_t* permutation_polys,
                               int ncycles, const int32_t* cycles_flat, int max_degree, b200_hyperplonk** out);
void b200_hyperplonk_free(b200_hyperplonk* pp);
int b200_hyperplonk_info(const b200_hyperplonk* pp, int* num_permutation_z_polys, int* degree, int* num_polys);
/* verifier parameters: preprocess_out[npreprocess], permutation_out[nperm] affine commitments */
int b200_hyperplonk_commitments(const b200_hyperplonk* pp, void* preprocess_out, void* permutation_out);
int b200_hyperplonk_permutation_poly(const b200_hyperplonk* pp, int i, void* host_out);
int b200_hyperplonk_prove(b200_hyperplonk* pp, const void* host_instances_fr, int ninstances,
                          const void* const* dev_witness);

/* Circuits with several instance columns and / or several witness phases (PlonkishCircuitInfo::{num_instances,
 * num_witness_polys, num_challenges}, pb/backend.rs:50-60; the phase loop of HyperPlonk::prove, hyperplonk.rs:183-204).
 * Polynomial order: instance columns | preprocess | witness (phase 0, phase 1, ...) | permutation | lookup m | lookup h |
 * permutation z; challenge order: the phases' challenges, then beta, gamma, alpha (preprocessor.rs:28-30).
 * preprocess_phased: num_instances[ninstance_cols], num_witness_polys[nphases], num_challenges[nphases]; every phase
 * needs witness polynomials and every phase but the last needs challenges (is_well_formed, backend.rs:76-105), else
 * B200_ERR_ARG. The other arguments are those of b200_hyperplonk_preprocess.
 * prove_phased: host_instances_fr = the instance columns back to back (ninstances = their total length). `synthesize`
 * plays PlonkishCircuit::synthesize(round, challenges) (backend.rs:100-110): it receives the challenges squeezed so far
 * (host, Montgomery) and stores the device pointers of that phase's num_witness_polys[round] witness polynomials
 * (2^k elements each, caller-owned, alive until prove returns; work enqueued on the context stream is ordered before
 * the commitments) into dev_witness_out; a non-zero return aborts the proof with B200_ERR_ARG. */
typedef int (*b200_synthesize_fn)(void* user, int round, const void* host_challenges_fr, int nchallenges,
                                  const void** dev_witness_out);
int b200_hyperplonk_preprocess_phased(b200_ctx* ctx, int k, int ninstance_cols, const int32_t* num_instances, int nphases,
                                      const int32_t* num_witness_polys, const int32_t* num_challenges, int npreprocess,
                                      const void* const* dev_preprocess, int nconstraints,
                                      const int32_t* constraint_tokens, int nconstraint_tokens, int nlookups,
                                      const int32_t* lookup_tokens, int nlookup_tokens, const void* consts_fr,
                                      int nconsts, int nperm, const int32_t* permutation_polys, int ncycles,
                                      const int32_t* cycles_flat, int max_degree, b200_hyperplonk** out);
int b200_hyperplonk_prove_phased(b200_hyperplonk* pp, const void* host_instances_fr, int ninstances,
                                 b200_synthesize_fn synthesize, void* user);

/* permutation_z_polys (pb/backend/hyperplonk/prover.rs:252-345) for one chunk of `npolys` wire columns:
 * grand-product polynomial z in BooleanHypercube order; id_offsets[i] = (index of wire i among the permuted
 * columns) << num_vars; host_beta_gamma = {beta, gamma}. dev_z_out[2^num_vars]. */
int b200_permutation_z(b200_ctx* ctx, int num_vars, int npolys, const void* const* dev_wires,
                       const void* const* dev_sigmas, const uint64_t* id_offsets, const void* host_beta_gamma,
                       void* dev_z_out);

/* LogUp helper polynomials of HyperPlonk's own lookup argument (pb/backend/hyperplonk/prover.rs:50-250).
 * b200_expression_rows: Expression::evaluate on every hypercube row (prover.rs:96-117) with the bytecode and the
 * dense leaf tables of b200_sumcheck_prove_generic; the host passes Σ_j beta^j expr_j to obtain a compressed
 * input / table polynomial (lookup_compressed_poly, :78-134). dev_out[2^num_vars].
 * b200_lookup_m: multiplicities (lookup_m_poly, :143-192); a value present on several table rows is counted on the
 * last one; B200_ERR_LOOKUP if some input value is missing from the table.
 * b200_lookup_h: h = 1/(gamma + input) - m/(gamma + table) (lookup_h_poly, :206-250). */
int b200_expression_rows(b200_ctx* ctx, int num_vars, int ntables, const void* const* dev_tables, int nconsts,
                         const void* host_consts_fr, int nops, const int32_t* host_ops, void* dev_out);
int b200_lookup_m(b200_ctx* ctx, int num_vars, const void* dev_input, const void* dev_table, void* dev_m_out);
int b200_lookup_h(b200_ctx* ctx, int num_vars, const void* dev_input, const void* dev_table, const void* dev_m,
                  const void* host_gamma, void* dev_h_out);

/* ---- variable_base_msm (pb/util/arithmetic/msm.rs:84-115) ------------------------------------- */
/* Σ scalars[i] * bases[i] with HOST inputs (the free function's signature); out = affine point.
 * An identity result is returned as (0, 0). */
int b200_variable_base_msm(b200_ctx* ctx, const void* host_scalars_fr, const void* host_bases_g1,
                           uint64_t n, void* host_out_g1);

/* ---- MultilinearKzg (pb/pcs/multilinear/kzg.rs) --------------------------------------------------- */
/* ProverParam: upload eqs[level] (2^level affine points, kzg.rs:36-53, :230-245 trim). Levels must be
 * uploaded in increasing order starting from 0. */
int b200_kzg_srs_upload(b200_ctx* ctx, int level, const void* host_g1);
/* setup + trim (kzg.rs:166-250) on the device from the trapdoor scalars ss[0..num_vars) (the reference
 * samples them from its RNG; here the caller supplies them): builds eqs[0..=num_vars]. */
int b200_kzg_setup(b200_ctx* ctx, const void* host_ss_fr, int num_vars);
/* copy eqs[level] (2^level affine points) back to the host */
int b200_kzg_srs_download(b200_ctx* ctx, int level, void* host_g1_out);
/* batch_commit (kzg.rs:259-274): one MSM per polynomial against eqs[num_vars[i]]; out[npolys] affine.
 * write_transcript != 0 also performs write_commitments (Pcs::batch_commit_and_write, pcs.rs:62-75). */
int b200_kzg_batch_commit(b200_ctx* ctx, const void* const* dev_polys, const int* num_vars, int npolys,
                          int write_transcript, void* host_out_g1);
/* open (kzg.rs:276-302): writes num_vars quotient commitments to the context transcript */
int b200_kzg_open(b200_ctx* ctx, const void* dev_poly, int num_vars, const void* host_point);
/* batch_open (kzg.rs:304-315 -> pb/pcs/multilinear.rs:134-235). host_points: npoints*num_vars elements;
 * evaluation k = (ev_poly[k], ev_point[k], host_ev_values[k]) mirrors `Evaluation` (pcs.rs:132-155). */
int b200_kzg_batch_open(b200_ctx* ctx, int num_vars, const void* const* dev_polys, int npolys,
                        const void* host_points, int npoints, const int* ev_poly, const int* ev_point,
                        const void* host_ev_values, int nevals);

/* EVAL-shape sum-checks (b200_sumcheck_prove_evals*, and those inside the Lasso / KZG provers) normally run the
 * eq-FACTORED round kernel: p_i(X) = c_i eq1(X, y_i) Q_i(X) with Q_i accumulated against the eq table of the remaining
 * variables (9 instead of 12 Montgomery products per pair for eq*a*b, no eq-table bind). Same messages, same bytes;
 * on = 0 selects the plain kernel that binds a materialised eq table (kept for A/B measurements and tests). */
int b200_sumcheck_eq_factored(b200_ctx* ctx, int on);

/* ---- multi-GPU: one process per GPU, collectives through NVLink peer memory --------- */
/* CUDA-IPC handles (128 bytes: mailbox | bulk arena) of this context; exchange the handles of all ranks out of band
 * (e.g. torch.distributed.all_gather), then call b200_dist_init with the `world` handles in rank order. world must be
 * a power of two <= 8. The mailbox carries the small in-kernel collectives (round partials, evaluations, partial
 * commitments), the arena (B200_ARENA_MB, default 320 MiB) the bulk all-gathers of bound sum-check tables. */
int b200_dist_mailbox_handle(b200_ctx* ctx, void* out_handle128);
int b200_dist_init(b200_ctx* ctx, int rank, int world, const void* handles);
/* The same group built from `world` contexts of ONE process (same GPU, or peer-accessible GPUs): plain device
 * pointers instead of IPC. Every context must then be driven from its own host thread. Used by the tests to run the
 * sharded provers on a single-GPU box; production is one process per GPU. */
int b200_dist_init_local(b200_ctx* const* ctxs, int world);
/* Every in-kernel wait for a peer is bounded (B200_PEER_TIMEOUT_S, default 20 s): returns B200_ERR_PEER if one timed
 * out since the last call (the results of that collective are garbage), else 0. Synchronises the stream. */
int b200_dist_check(b200_ctx* ctx);
/* Fully sharded Lasso prover (cfg4): after b200_dist_shard_lasso(ctx, k0) with k0 > 0, b200_lasso_prove* keeps the
 * 2^mu-sized witness tables, fingerprints, product-tree layers >= k0 and all sum-checks / openings over them on the
 * rank's 1/world slice — the entries whose index bits [k0 - log2 world, k0) equal the rank: closed under the LSB-first
 * binds (multilinear.rs:612-616) and under the top-bit tree halving (fractional_sum_check.rs:41-76). Proofs stay
 * byte-identical. Implies point-sharded commitments; mu <= k0 falls back to the replicated prover. 0 = off. */
int b200_dist_shard_lasso(b200_ctx* ctx, int k0);
/* exchange tuning knobs for experiments (tools/micro/shard_tune.py): key 0 small-message protocol (0, 1, 2), key 1
 * heartbeat CTAs (0 = off), key 2 heartbeat sleep in ns, key 3 heartbeat mode bits (1 NVLink stores, 2 HBM reads), key 4 heartbeat self-timeout in ms */
int b200_dist_tune(b200_ctx* ctx, int key, int value);
/* a sharded sum-check round is exchanged over NVLink while a rank holds at least `items` (pair, term) items; below
 * that the bound tables are all-gathered once and the remaining rounds run replicated (default 2^16) */
int b200_dist_shard_min_items(b200_ctx* ctx, int items);
/* Point-sharded commitments: after b200_dist_shard_commits(ctx, 1) every commitment MSM issued by b200_kzg_* and
 * b200_lasso_prove* (variable_base_msm call sites kzg.rs:255,271,292) is split by point range over the ranks and the
 * partial commitments are all-gathered over NVLink and added, so these calls become COLLECTIVE: all ranks run the same
 * prover on the same inputs (everything between the commitments is replicated) and produce the identical proof. */
int b200_dist_shard_commits(b200_ctx* ctx, int on);
/* Hypercube-sharded sum-checks inside the whole provers: after b200_dist_shard_sumchecks(ctx, min_vars) with
 * min_vars > 0, every EvaluationsProver sum-check of b200_lasso_prove* (the Surge primary sum-check and the per-layer
 * grand-product sum-checks) over at least min_vars variables is evaluated on the rank's 1/world slice of the (replicated)
 * tables, with the round partials exchanged over NVLink as in b200_sumcheck_prove_evals_sharded. Collective like
 * b200_dist_shard_commits: all ranks run the same prover on the same inputs and produce the identical proof.
 * min_vars = 0 switches it off. */
int b200_dist_shard_sumchecks(b200_ctx* ctx, int min_vars);

/* b200_sumcheck_prove_evals on a hypercube sharded over the TOP log2(world) variables: rank g passes the
 * slices [g*2^n/world, (g+1)*2^n/world) of every table. All ranks must call it; all receive the same
 * challenges / evals and append the same bytes to their transcripts (identical to the unsharded proof). */
int b200_sumcheck_prove_evals_sharded(b200_ctx* ctx, int num_vars_total, int nterms, int np,
                                      const void* const* dev_local_tables, const void* host_weights,
                                      const void* host_y, const void* host_sum, void* host_challenges_out,
                                      void* host_evals_out);
/* The general layout: the tables are sharded on the index bits [window_pos, window_pos + log2 world) (rank = those
 * bits, local index = high bits ‖ low window_pos bits; window_pos = -1 means the top variables as above). The first
 * `sharded_rounds` <= window_pos rounds exchange their partials over NVLink inside the round kernel (-1: as many as pay,
 * see b200_dist_shard_min_items), then the bound tables are all-gathered once and the rest runs replicated. */
int b200_sumcheck_prove_evals_windowed(b200_ctx* ctx, int num_vars_total, int window_pos, int sharded_rounds, int nterms,
                                       int np, const void* const* dev_local_tables, const void* host_weights,
                                       const void* host_y, const void* host_sum, void* host_challenges_out,
                                       void* host_evals_out);
/* variable_base_msm with the points sharded by range: every rank passes its slice, all get the full sum */
int b200_variable_base_msm_sharded(b200_ctx* ctx, const void* host_scalars_fr, const void* host_bases_g1,
                                   uint64_t n_local, void* host_out_g1);

/* ---- Lasso / Surge lookup argument (north_star; no counterpart in the mounted snapshot).
 * Specification: DESIGN.md "Lasso protocol"; CPU restatement: oracle/lasso.hpp. ---------------------- */
#define B200_TABLE_RANGE 0 /* chunks x 16-bit limbs, identity subtable, g = sum 2^(16 t) E_t */
#define B200_TABLE_AND 1   /* chunks x (8|8)-bit operand bytes, subtable p & q, g = sum 2^(8 t) E_t */
#define B200_TABLE_XOR 2
/* Proves 2^mu lookups (host_xs[, host_ys] operands, u64 each) and appends the whole proof to the context
 * transcript: commitments to a, dim_*, E_*, read_ts_*, final_cts_*; primary Surge sum-check; memory-
 * checking grand products; leaf evaluations; two batch openings. The SRS must cover max(mu, 16) levels. */
int b200_lasso_prove(b200_ctx* ctx, int table_kind, int chunks, int mu, const uint64_t* host_xs,
                     const uint64_t* host_ys);
/* same with operands already on the device (u64 arrays) */
int b200_lasso_prove_dev(b200_ctx* ctx, int table_kind, int chunks, int mu, const void* dev_xs, const void* dev_ys);
/* Decomposable tables as DATA instead of a built-in kind (the role of a `DecomposableTable` implementation in the Lasso
 * frontend north_star names: chunk bits, subtable values, the combiner g): every lookup splits into `chunks` chunks, chunk
 * t addresses ONE 2^16-entry subtable T, and the lookup output is g(E) = sum_t 2^(out_bits t) T[dim_t].
 *   num_operands = 1: dim_t = chunk t (operand_bits bits) of x;
 *   num_operands = 2: dim_t = (chunk t of x) << operand_bits | (chunk t of y)   (num_operands * operand_bits <= 16).
 * The table is part of the proved statement: (3, chunks, mu, num_operands, operand_bits, out_bits, digest) is absorbed,
 * digest = Keccak-256 of the 2^16 values as little-endian u32 words, as a little-endian integer mod r. Operands that do not
 * fit operand_bits * chunks bits are rejected with B200_ERR_LOOKUP before anything reaches the transcript. */
typedef struct {
  int chunks;               /* 2..8: memories = chunks */
  int num_operands;         /* 1 or 2 */
  int operand_bits;         /* bits of one operand chunk */
  int out_bits;             /* g = sum_t 2^(out_bits t) E_t; out_bits * (chunks - 1) + bits(max T) <= 64 */
  const uint32_t* subtable; /* host, 2^16 values (all of them: the whole table is committed to by its digest) */
} b200_lasso_table;
typedef struct b200_lasso_tab b200_lasso_tab; /* uploaded table (device values + digest) */
int b200_lasso_table_create(b200_ctx* ctx, const b200_lasso_table* table, b200_lasso_tab** out);
void b200_lasso_table_free(b200_lasso_tab* tab);
/* b200_lasso_prove / b200_lasso_prove_dev for an uploaded table (host_ys / dev_ys may be NULL when num_operands = 1) */
int b200_lasso_prove_table(b200_ctx* ctx, const b200_lasso_tab* tab, int mu, const uint64_t* host_xs, const uint64_t* host_ys);
int b200_lasso_prove_table_dev(b200_ctx* ctx, const b200_lasso_tab* tab, int mu, const void* dev_xs, const void* dev_ys);
/* prove_fractional_sum_check (pb/piop/gkr/fractional_sum_check.rs:87-190): GKR argument for Σ_i p_b[i] / q_b[i] over
 * num_batching (<= 10) pairs of 2^num_vars-entry device tables (Montgomery Fr), on the context's transcript: writes (or,
 * where bit b / bit 16 + b of claimed_mask marks p_b / q_b as a public claim = Some(_), absorbs) the layer-0 values, then
 * per layer gamma, the degree-3 ClassicSumCheck<EvaluationsProver> rounds, the 4 * num_batching evaluations and mu, exactly
 * as the reference. Returns (p_xs, q_xs, x) of the reference plus the layer-0 values (host buffers, may be NULL). */
int b200_fractional_sum_check_prove(b200_ctx* ctx, int num_batching, int num_vars, const void* const* dev_ps,
                                    const void* const* dev_qs, uint32_t claimed_mask, void* host_p_xs, void* host_q_xs,
                                    void* host_x, void* host_p_0s, void* host_q_0s);
/* witness tables only: dev_mtabs = a | dim[c] | E[c] | read_ts[c] (2^mu each), dev_stabs = final_cts[c] (2^16 each) */
int b200_lasso_witness(b200_ctx* ctx, int table_kind, int chunks, int mu, const uint64_t* host_xs,
                       const uint64_t* host_ys, void* dev_mtabs, void* dev_stabs);

#ifdef __cplusplus
}
#endif
#endif

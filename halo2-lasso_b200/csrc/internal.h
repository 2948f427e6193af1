// Host-side internal interface between the translation units of libb200lasso.so.
// Everything here is asynchronous on ctx->stream; scalars (claims, challenges, evaluations) stay in
// device memory so that no step of a proof needs a host round trip.
#pragma once
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include <vector>

#include "common.cuh"
#include "g1.cuh"
#include "peer.cuh"

namespace b200 {

struct ScState {
  Fr claim;  // running claimed sum
  Fr r;      // challenge of the previous round
  Fr eqc;    // eq-factored rounds: c_i = scale * Π_{j<i} eq1(r_j, y_j)
  unsigned int counter;
  unsigned int pad[7];
};

struct Ctx {
  int device;
  cudaStream_t stream;
  Transcript* d_tr;
  uint8_t* d_proof;
  uint32_t proof_cap;
  BaryTable* d_bary;
  ScState* d_sc;
  Fr* d_partial;       // per-CTA partial sums of the round kernels
  size_t partial_elems;
  uint64_t launches;   // kernels launched since the last reset (bench "gpu_launches")
  // optional per-launch CUDA-event timing of the sum-check round kernels (bench roofline leg)
  bool profile;
  long long* dbg_clocks;  // device, 32 rounds x 16 stamps (null unless B200_DEBUG_CLOCKS is set)
  std::vector<cudaEvent_t> prof_events;  // pairs (start, stop)
  std::vector<int> prof_tags;            // round index per pair
  // SRS: eqs[k] = 2^k affine points (MultilinearKzgProverParams::eqs, kzg.rs:36-53)
  std::vector<G1Aff*> srs;
  // precomputed window multiples: srs_ext[k][w * 2^k + i] = 2^(16 w) * srs[k][i], w < EXT_WINDOWS.
  // With them every window of a full-width scalar lands in ONE shared bucket set, which removes the
  // per-window reduction and the 254 dependent doublings of the final Horner pass.
  std::vector<G1Aff*> srs_ext;
  // multi-GPU (one process per GPU): peer mailboxes mapped over NVLink, see peer.cuh
  PeerCtx peer;            // world == 1 when not initialised
  Mailbox* my_mailbox;     // cudaMalloc'ed, exported through CUDA IPC
  bool peer_ipc = false;   // peer pointers are CUDA-IPC mappings (closed on destroy)
  unsigned int peer_seq;   // small collectives issued so far (identical on every rank)
  unsigned int bulk_seq = 0;  // bulk all-gathers issued so far
  unsigned char* my_arena = nullptr;  // bulk arena (2 halves), exported like the mailbox
  unsigned int* d_peer_err = nullptr;  // device word raised by a timed-out wait (B200_ERR_PEER)
  // heartbeat (B200_HEARTBEAT="ctas,sleep_ns,write_peers"): a low-priority side kernel that keeps SMs and NVLink links
  // from idling while a sharded proof is latency-bound (shard.cu: HeartbeatScope)
  cudaStream_t hb_stream = nullptr;
  cudaEvent_t hb_event = nullptr;
  unsigned int* hb_stop = nullptr;  // device word: the running heartbeat kernel leaves when it equals its generation
  unsigned int hb_gen = 0;
  int hb_ctas = 0, hb_sleep_ns = 1000, hb_write = 1, hb_depth = 0, hb_max_ms = 2000;
  int shard_min_items = 1 << 16;  // a sum-check round stays sharded while a rank has at least this many (pair, term) items
  int shard_lasso_k0 = 0;  // > 0: the Lasso prover shards its tables / trees on the index window [k0 - g, k0) (lasso.cu)
  bool eq_factored = true;  // EVAL-shape sum-checks use the eq-factored round kernel (b200_sumcheck_eq_factored)
  bool shard_commits = false;  // every commitment MSM is split by point range over the ranks (all ranks call)
  // value-sorted address lists of the built-in and / xor subtables for grouped E commitments (MsmJob::group_*), device
  uint32_t* d_group_perm[3] = {nullptr, nullptr, nullptr};
  uint32_t* d_group_off[3] = {nullptr, nullptr, nullptr};
  int shard_sumcheck_min_vars = 0;  // > 0: sum-checks of the whole provers with at least that many variables run
                                    // hypercube-sharded over the ranks (sumcheck_prove_evals_dist, shard.cu)
};
static const int EXT_C = 16;        // window bits of the precomputed tables
static const int EXT_WINDOWS = 16;  // ceil(255 / 16)

static const int SC_MAX_TERMS = 32;
static const int SC_MAX_TABLES = 64;

// EVAL shape  F(x) = eq(x, y) * Σ_t w_t * Π_{k<NP} P_{t,k}(x)       (NP = 1 or 2)
struct ScEvalJob {
  int num_vars, T, NP;
  const Fr* tables[SC_MAX_TABLES];  // [t*NP + k], each 2^num_vars, read-only
  const Fr* weights;                // device, T
  const Fr* eq_point;               // device, num_vars
  const Fr* claim;                  // device, 1
  Fr* challenges_out;               // device, num_vars
  Fr* evals_out;                    // device, T*NP (+1 when want_eq_eval: the bound eq value last)
  const Fr* eq_table = nullptr;     // optional prebuilt eq table (2^num_vars); eq_point is then unused
  const Fr* eq_scale = nullptr;     // optional device scalar multiplied into the eq table built from eq_point
  bool sharded = false;             // sum the per-round partials over all ranks (peer mailboxes)
  bool want_eq_eval = false;
  // stop_after > 0: run only that many rounds and hand the state over (sharded drivers, shard.cu): carry->cur[i] are
  // the tables as the NEXT round would read them (ntab tables then eq; the last challenge, still unbound, is in
  // Ctx::d_sc->r), scratch is allocated from carry->scope so that it outlives this call
  int stop_after = 0;
  struct ScCarry* carry = nullptr;
};
struct ScCarry {
  struct DevScope* scope;
  const Fr* cur[2 * SC_MAX_TABLES + 2];
  uint64_t len;  // entries per table in cur[]
};
int sumcheck_prove_evals(Ctx* c, const ScEvalJob& job);
// shard.cu — hypercube-sharded sum-check and point-sharded MSM over peer memory
int sumcheck_prove_evals_sharded(Ctx* c, const ScEvalJob& job_local, int num_vars_total, int p = -1, int rounds = -1);
// replicated full tables in, sharded evaluation when enabled (b200_dist_shard_sumchecks); else sumcheck_prove_evals
int sumcheck_prove_evals_dist(Ctx* c, const ScEvalJob& job);


// COEFF shape  F(x) = Σ_k s_k * eq(x, y_k) * P_k(x)   (CoefficientsProver, degree 2)
struct ScCoeffJob {
  int num_vars, K;
  const Fr* tables[SC_MAX_TERMS];     // P_k
  const Fr* eq_points[SC_MAX_TERMS];  // device, num_vars each
  const Fr* scalars;                  // device, K
  const Fr* claim;
  Fr* challenges_out;
  Fr* evals_out;  // K
  const Fr* eq_tables[SC_MAX_TERMS] = {nullptr};  // optional prebuilt eq tables (all or none); eq_points unused then
  bool sharded = false;                 // as ScEvalJob
  int stop_after = 0;                   // carry->cur = K tables then K eq tables
  struct ScCarry* carry = nullptr;
};
int sumcheck_prove_coeffs(Ctx* c, const ScCoeffJob& job);
int sumcheck_prove_coeffs_sharded(Ctx* c, const ScCoeffJob& job_local, int num_vars_total, int p = -1, int rounds = -1);
// shard.cu helpers: bulk all-gather into the peer arenas, small all-reduce, evaluate() of sharded polynomials
int shard_allgather(Ctx* c, const Fr* const* src, int ntab, uint32_t len, int q, bool bind, const Fr** full_out);
int shard_allreduce(Ctx* c, Fr* d_vals, int cnt);
int mle_eval_many_sharded(Ctx* c, const Fr* const* h_tables_loc, int ntables, int n, int p, const Fr* d_point, Fr* d_out);

// generic.cu — EvaluationsProver for an arbitrary Expression compiled to bytecode on the host
struct GenericJob {
  int num_vars, ntables, nconsts, nops, degree, ntemps;
  const Fr* tables[64];  // every leaf of the expression as a dense 2^num_vars table
  const Fr* consts;      // device, nconsts (Montgomery)
  const int4* ops;       // device, nops x (opcode, dst, a, b)
  const Fr* claim;       // device
  Fr* challenges_out;    // device, num_vars
  Fr* evals_out;         // device, ntables
};
int sumcheck_prove_generic(Ctx* c, const GenericJob& job);
int poly_iota(Ctx* c, int num_vars, Fr* d_out);
int poly_onehot(Ctx* c, int num_vars, uint64_t index, Fr* d_out);
int poly_rotate(Ctx* c, const Fr* d_in, int num_vars, int rotation, Fr* d_out);

// perm.cu — permutation_z_polys (prover.rs:252-345), one chunk
int permutation_z(Ctx* c, int num_vars, int npolys, const Fr* const* wires, const Fr* const* sigmas,
                  const uint64_t* id_offsets, const Fr* d_beta_gamma, Fr* d_z);
// all chunks: nz z polynomials, chunk zi = wire columns [zi * chunk_size, min(npolys, (zi + 1) * chunk_size))
int permutation_z_chunks(Ctx* c, int num_vars, int nz, int chunk_size, int npolys, const Fr* const* wires,
                         const Fr* const* sigmas, const uint64_t* id_offsets, const Fr* d_beta_gamma, Fr* const* d_z);

// mle.cu
int eq_build(Ctx* c, const Fr* d_y, int n, Fr* d_out);                         // eq_xy
int fix_var(Ctx* c, const Fr* d_in, int n, const Fr* d_r, Fr* d_out);          // one bind
int mle_eval_many(Ctx* c, const Fr* const* h_tables, int ntables, int n, const Fr* d_point,
                  Fr* d_out);                                                   // evaluate
int mle_dot_many(Ctx* c, const Fr* const* h_tables, int ntables, size_t len, const Fr* d_eq, Fr* d_out);  // <P_t, eq>
int fr_lincomb(Ctx* c, const Fr* const* h_tables, int k, const Fr* d_scalars, size_t len,
               Fr* d_out);                                                      // Σ s_i P_i
int fr_convert(Ctx* c, const Fr* d_in, Fr* d_out, size_t n, int to_mont);
int fr_scale(Ctx* c, Fr* d_tab, size_t n, const Fr* d_scalar);  // tab[i] *= scalar
int fr_from_u64(Ctx* c, const uint64_t* d_in, Fr* d_out, size_t n);
int quotient_step(Ctx* c, Fr* d_rem, size_t half, const Fr* d_x, Fr* d_q);     // pcs/multilinear.rs:72-107
int eq_xy_eval_dev(Ctx* c, const Fr* d_x, const Fr* d_y, int n, Fr* d_out);    // sum_check.rs:111-121
int transcript_op(Ctx* c, int op, const Fr* d_in, Fr* d_out, int n);           // 0 common, 1 write, 2 squeeze
int transcript_write_points(Ctx* c, const G1Aff* d_pts, int n);

enum { TR_COMMON = 0, TR_WRITE = 1, TR_SQUEEZE = 2 };

// msm.cu — variable_base_msm (pb/util/arithmetic/msm.rs:84-181), batched
enum MsmScalarKind { MSM_FR_MONT = 0, MSM_FR_CANON = 1, MSM_U64 = 2, MSM_U32 = 3 };
struct MsmJob {
  const void* scalars;  // device
  const G1Aff* bases;   // device
  uint64_t n;
  int kind;             // MsmScalarKind
  int bits;             // significant bits of the largest scalar (254 for arbitrary Fr)
  const G1Aff* ext;     // optional precomputed window multiples of `bases` (Ctx::srs_ext layout), or null
  uint64_t ext_stride = 0;  // points per window in `ext` (0: n) — larger than n when the job is a point range
  // map_g > 0: the scalars are one rank's compact slice of a polynomial sharded on the index bits [map_p, map_p + map_g):
  // scalar i multiplies base ((i >> p) << (p + g)) | (rank << p) | (i & (2^p - 1))
  int map_p = 0, map_g = 0, map_rank = 0;
  // GROUPED job (no scalars of its own: scalars = nullptr, n = 0): result = Σ_v v * (Σ_{k in group v} B_k) over the bucket
  // sums B_k of job `group_src` of the same batch, whose scalars are 16-bit addresses d (bucket k = address k + 1) — the
  // commitment to E = T[dim] falls out of the buckets of the commitment to dim (same bases, E constant per address), so
  // its 2^mu mixed additions are replaced by one sum per table value. group_perm = the addresses' bucket indices sorted
  // by table value, group_off[v - 1 .. v] = the range of value v (v = 1 .. ngroups); addresses with T = 0 are left out.
  int group_src = -1, ngroups = 0;
  const uint32_t* group_perm = nullptr;
  const uint32_t* group_off = nullptr;
};
// a grouped E commitment pays from this many points on (below, forcing 2^16 buckets on the dim job costs more than it saves)
uint64_t msm_group_min_points();
// extra result J + i = Σ_t 2^(shift t) * result(src[t]): a commitment that is a linear combination of other commitments
// of the same batch (Lasso: a = Σ_t 2^(w t) E_t) costs doublings instead of an MSM
struct MsmDerive {
  int nsrc, shift, src[8];
};
int msm_batch(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out, const MsmDerive* derive = nullptr, int nderive = 0);
// shard.cu: msm_batch, point-sharded over the ranks when commit sharding is on (collective), else local
int msm_batch_dist(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out, const MsmDerive* derive = nullptr, int nderive = 0);
int msm_sharded(Ctx* c, const MsmJob& local, G1Aff* d_out);  // shard.cu

// kzg.cu — MultilinearKzg (pb/pcs/multilinear/kzg.rs) + additive::batch_open (pb/pcs/multilinear.rs:134-235)
struct BatchOpenJob {
  int num_vars, npolys, npoints, nevals;
  const Fr* const* polys;  // host array of device pointers
  const Fr* points;        // device, npoints * num_vars
  const int* ev_poly;      // host
  const int* ev_point;     // host
  const Fr* ev_values;     // device, nevals
  int shard_p = -1;        // >= 0: polys are the rank's slices of polynomials sharded on the window [shard_p, shard_p + g)
};
int kzg_commit_batch(Ctx* c, const MsmJob* jobs, int J, bool write_transcript, G1Aff* d_out);
int kzg_open(Ctx* c, const Fr* d_poly, int n, const Fr* d_point, int shard_p = -1);
int kzg_batch_open(Ctx* c, const BatchOpenJob& job);
int kzg_setup(Ctx* c, const Fr* d_ss, int n);
int kzg_build_ext(Ctx* c, int level);  // fills c->srs_ext[level]

// gkr.cu — GKR for fractional sum-checks (pb/piop/gkr/fractional_sum_check.rs:87-190)
// d_out: p_xs[B] | q_xs[B] | x[n] | p_0s[B] | q_0s[B]; claimed_mask bit b / 16 + b: the layer-0 value p_b / q_b is a
// public claim (absorbed) instead of being written to the proof
int fractional_sum_check_prove(Ctx* c, int B, int n, const Fr* const* d_ps, const Fr* const* d_qs, uint32_t claimed_mask,
                               Fr* d_out);

// lasso.cu — Lasso / Surge prover (DESIGN.md §Lasso protocol; oracle/lasso.hpp)
// A decomposable table given as DATA (b200_lasso_table, include/b200_lasso.h): one 2^16-entry subtable in device memory,
// 1 or 2 operands of operand_bits bits per chunk, g = Σ_t 2^(out_bits t) E_t; digest = Keccak-256 of the values (LE u32
// words) as a little-endian integer mod r, absorbed with the statement; value_bits = bit length of the largest value.
struct LassoTableDesc {
  int chunks, num_operands, operand_bits, out_bits, value_bits;
  const uint32_t* d_values;
  Fr digest;  // Montgomery
  // grouped E commitments (MsmJob::group_*): set when T[0] = 0 and the values fit 12 bits, else null
  const uint32_t* d_group_perm = nullptr;
  const uint32_t* d_group_off = nullptr;
  int ngroups = 0;
};
// host: value-sorted bucket indices of a 2^16-entry table (perm: 65535 entries at most, off: ngroups + 1); false when
// the table is not eligible (T[0] != 0 or a value above 4095)
bool lasso_group_lists(const uint32_t* values, std::vector<uint32_t>* perm, std::vector<uint32_t>* off);
int lasso_prove(Ctx* c, int kind, int chunks, int mu, const uint64_t* d_xs, const uint64_t* d_ys,
                const LassoTableDesc* desc = nullptr);
int lasso_witness(Ctx* c, int kind, int chunks, int mu, const uint64_t* d_xs, const uint64_t* d_ys, Fr* d_mt,
                  Fr* d_st);

inline void count_launch(Ctx* c, int n = 1) { c->launches += n; }

// B200_TRACE=1: host-side timeline of a rank (stderr, wall-clock ms) — for chasing stalls between the ranks of a group
void trace_point(Ctx* c, const char* what);

// CUDA loads kernels lazily, on their first launch, and that load may synchronise the whole context. A proof should not
// stall on it, and ranks that share one context (b200_dist_init_local) would deadlock on it: a rank's kernel waiting in
// a collective for a peer whose own kernel cannot be loaded before the context drains. cudaFuncGetAttributes forces the
// load; b200_ctx_create calls preload_all_kernels() once per process.
#define B200_PRELOAD(...)                                     \
  do {                                                        \
    cudaFuncAttributes fa_;                                   \
    cudaFuncGetAttributes(&fa_, (const void*)(__VA_ARGS__)); \
  } while (0)
void preload_generic();
void preload_gkr();
void preload_hyperplonk();
void preload_kzg();
void preload_lasso();
void preload_lookup();
void preload_mle();
void preload_msm();
void preload_perm();
void preload_shard();
void preload_sumcheck();

// NVTX range named like the reference's timer at the same site (pb/util/timer.rs:19-60: start_timer / end_timer around
// variable_base_msm-N, sum_check_prove-n-d, sum_check_prove_round-i, pcs_batch_open-N, merged_polys, g_prime, quotients,
// witness_collector-i, lookup_*_polys-N, permutation_z_polys-N, evals-N): shows up on the host timeline of nsys / ncu
// --nvtx; costs two library calls when no tool is attached.
struct NvtxRange {
  explicit NvtxRange(const char* fmt, ...) {
    char name[96];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(name, sizeof(name), fmt, ap);
    va_end(ap);
    nvtxRangePushA(name);
  }
  NvtxRange(const NvtxRange&) = delete;
  ~NvtxRange() { nvtxRangePop(); }
};

// Stream-ordered scratch released on EVERY exit path of a host function (error returns included).
struct DevScope {
  cudaStream_t s;
  std::vector<void*> ptrs;
  explicit DevScope(cudaStream_t stream) : s(stream) {}
  DevScope(const DevScope&) = delete;
  DevScope& operator=(const DevScope&) = delete;
  template <class T>
  cudaError_t alloc(T** p, size_t bytes) {
    void* q = nullptr;
    cudaError_t e = cudaMallocAsync(&q, bytes ? bytes : 32, s);
    *p = reinterpret_cast<T*>(q);
    if (e == cudaSuccess) ptrs.push_back(q);
    return e;
  }
  ~DevScope() {
    for (void* p : ptrs) cudaFreeAsync(p, s);
  }
};
// Keeps the heartbeat kernel (if configured) running for the lifetime of the scope, in STREAM order: started behind what
// is already queued on ctx->stream, stopped by a tiny kernel queued when the scope ends. Nested scopes share one.
struct HeartbeatScope {
  Ctx* c;
  explicit HeartbeatScope(Ctx* ctx);
  ~HeartbeatScope();
  HeartbeatScope(const HeartbeatScope&) = delete;
};
// returns the index of the (start, stop) event pair, or -1 when profiling is off; pairs may nest
inline int prof_begin(Ctx* c, int tag) {
  if (!c->profile) return -1;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  c->prof_events.push_back(e0);
  c->prof_events.push_back(e1);
  c->prof_tags.push_back(tag);
  cudaEventRecord(e0, c->stream);
  return (int)c->prof_tags.size() - 1;
}
inline void prof_end(Ctx* c, int idx) {
  if (idx < 0) return;
  cudaEventRecord(c->prof_events[2 * idx + 1], c->stream);
}
enum {  // phase tags (>= 1000) of lasso_prove; tags < 1000 are sum-check round indices
  PH_WITNESS = 1000, PH_COMMIT, PH_PRIMARY, PH_TREES_M, PH_GKR_M, PH_TREES_S, PH_GKR_S, PH_LEAF_EVALS,
  PH_OPEN_M, PH_OPEN_S, PH_MSM_SORT = 1100, PH_MSM_ACC, PH_MSM_REDUCE
};

}  // namespace b200

// opaque handle of include/b200_lasso.h
struct b200_ctx {
  b200::Ctx c;
};

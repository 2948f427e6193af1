// HyperPlonk host orchestration inside the library (pb/backend/hyperplonk.rs:97-291, prover.rs:32-48, 348-409,
// verifier.rs:147-182): preprocess (commit the preprocess and permutation polynomials, compose the zero-check
// expression) and prove (instance polynomials, witness commitments, LogUp m / h polynomials, permutation grand
// product, zero check over the composed expression, rotated evaluations, additive batch opening). Every step is
// enqueued on the context stream; challenges, evaluation points and evaluations stay in device memory (the only host
// read is the "Invalid lookup input" flag of lookup_m).
//
// Also the `SumCheck::prove` entry point for an arbitrary `Expression` given as prefix tokens
// (b200_sumcheck_prove_expression): compile (expr.hpp), materialise the leaf tables, run generic.cu.
#include <algorithm>
#include <cstdlib>
#include <map>
#include <set>

#include "../../include/b200_lasso.h"
#include "expr.hpp"
#include "internal.h"

namespace b200 {

// lookup.cu
int lookup_m(Ctx* c, int num_vars, const Fr* d_input, const Fr* d_table, Fr* d_m);
int lookup_h(Ctx* c, int num_vars, const Fr* d_input, const Fr* d_table, const Fr* d_m, const Fr* d_gamma, Fr* d_h);
int expression_rows_prog(Ctx* c, int num_vars, const Fr* const* tables, int ntables, const Fr* d_consts, int nconsts,
                         const int4* d_ops, int nops, int ntemps, Fr* d_out);

// dst[idx[i]] = src[i]
__global__ void scatter_fr_kernel(const Fr* src, const uint64_t* idx, int n, Fr* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(dst + idx[i], fe_ld(src + i));
}
// dst[i] = Fr::from(src[i])
__global__ void u64_rows_to_fr_kernel(const uint64_t* __restrict__ src, size_t n, Fr* __restrict__ dst) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) fe_st(dst + i, fe_from_u64<FrP>(src[i]));
}
// rotation_eval_points (pb/poly/multilinear.rs:475-517): point p, coordinate i, from the sum-check point x and the
// pattern of that point (multilinear.rs:519-541, computed on the host — it depends on the rotation only).
__global__ void rotation_points_kernel(const Fr* x, int n, int rotation, const uint64_t* patterns, int npoints, Fr* out) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= npoints * n) return;
  const int p = t / n, i = t % n;
  const uint64_t pat = patterns[p];
  const int d = rotation < 0 ? -rotation : rotation, num_x = n - d;
  const Fr one = fe_one<FrP>();
  Fr v;
  if (rotation < 0) {
    if (i < num_x) {
      const Fr xi = fe_ld(x + d + i);
      v = ((pat >> i) & 1) ? one - xi : xi;
    } else {
      v = ((pat >> i) & 1) ? one : fe_zero<FrP>();
    }
  } else {
    if (i < d) {
      v = ((pat >> i) & 1) ? one : fe_zero<FrP>();
    } else {
      const Fr xi = fe_ld(x + i - d);
      v = ((pat >> i) & 1) ? one - xi : xi;
    }
  }
  fe_st(out + (size_t)p * n + i, v);
}
__global__ void gather_evals_kernel(const Fr* src, const int* idx, int n, Fr* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(dst + i, fe_ld(src + idx[i]));
}

// multilinear.rs:519-541
static std::vector<uint64_t> rotation_eval_point_pattern(bool next, int num_vars, int distance) {
  const uint64_t prim = BH_PRIMITIVE[num_vars], x_inv = prim >> 1;
  const uint64_t rem = next ? prim : x_inv;
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

// Device buffers that live until the end of one host call; freed stream-ordered.
struct Arena {
  Ctx* c;
  std::vector<void*> bufs;
  explicit Arena(Ctx* ctx) : c(ctx) {}
  ~Arena() {
    for (void* p : bufs) cudaFreeAsync(p, c->stream);
  }
  template <class T>
  T* alloc(size_t n) {
    void* p = nullptr;
    if (cudaMallocAsync(&p, (n ? n : 1) * sizeof(T), c->stream) != cudaSuccess) return nullptr;
    bufs.push_back(p);
    return (T*)p;
  }
  template <class T>
  T* upload(const T* host, size_t n) {  // pageable source: staged before cudaMemcpyAsync returns
    T* d = alloc<T>(n);
    if (d && n) cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, c->stream);
    return d;
  }
};

// Constants of a compiled program on the device: literals uploaded, challenge values copied device-to-device.
static Fr* program_consts(Ctx* c, Arena& ar, const Program& p, const Fr* d_challenges, std::vector<Fr>* host_keep) {
  const size_t C = p.consts.size();
  host_keep->resize(C ? C : 1);
  for (size_t i = 0; i < C; ++i) (*host_keep)[i] = p.consts[i].value;
  Fr* d = ar.upload<Fr>(host_keep->data(), C);
  if (!d) return nullptr;
  for (size_t i = 0; i < C; ++i)
    if (p.consts[i].chal >= 0)
      cudaMemcpyAsync(d + i, d_challenges + p.consts[i].chal, sizeof(Fr), cudaMemcpyDeviceToDevice, c->stream);
  return d;
}

// Every leaf as a dense table (rotated queries through the BooleanHypercube map, eq_xy, identity, Lagrange)
static int leaf_tables(Ctx* c, Arena& ar, int num_vars, const std::vector<Leaf>& leaves, const Fr* const* polys, int npolys,
                       const Fr* d_ys, int nys, std::vector<const Fr*>* out) {
  const size_t N = (size_t)1 << num_vars;
  for (auto& l : leaves) {
    int rc = B200_OK;
    if (l.kind == Expr::POLY) {
      if (l.a < 0 || l.a >= npolys) return B200_ERR_ARG;
      if (l.b == 0) {
        out->push_back(polys[l.a]);
        continue;
      }
      Fr* t = ar.alloc<Fr>(N);
      if (!t) return B200_ERR_NOMEM;
      rc = poly_rotate(c, polys[l.a], num_vars, l.b, t);
      out->push_back(t);
    } else {
      Fr* t = ar.alloc<Fr>(N);
      if (!t) return B200_ERR_NOMEM;
      if (l.kind == Expr::EQXY) {
        if (l.a < 0 || l.a >= nys) return B200_ERR_ARG;
        rc = eq_build(c, d_ys + (size_t)l.a * num_vars, num_vars, t);
      } else if (l.kind == Expr::IDENTITY) {
        rc = poly_iota(c, num_vars, t);
      } else {  // Lagrange(i): one-hot at the i-th row in BooleanHypercube order (classic.rs:44-55)
        rc = poly_onehot(c, num_vars, bh_nth(num_vars, l.a), t);
      }
      out->push_back(t);
    }
    if (rc) return rc;
  }
  return B200_OK;
}

// `ClassicSumCheck::<EvaluationsProver>::prove` for an arbitrary expression; d_evals_out[npolys] = every polynomial
// bound at the challenges (also those the expression never queries at rotation 0, classic.rs:143-149).
static int prove_expression(Ctx* c, int num_vars, const ExprP& expr, const Fr* const* polys, int npolys,
                            const Fr* d_challenges, const Fr* d_ys, int nys, const Fr* d_sum, Fr* d_x_out,
                            Fr* d_evals_out) {
  Arena ar(c);
  ExprCompiler comp;
  Program p = comp.compile(expr);
  std::vector<Fr> keep;
  Fr* d_consts = program_consts(c, ar, p, d_challenges, &keep);
  if (!d_consts) return B200_ERR_NOMEM;
  std::vector<const Fr*> tables;
  int rc = leaf_tables(c, ar, num_vars, p.leaves, polys, npolys, d_ys, nys, &tables);
  if (rc) return rc;
  const int K = (int)tables.size();
  std::vector<int> pos(npolys, -1);
  for (int i = 0; i < K; ++i)
    if (p.leaves[i].kind == Expr::POLY && p.leaves[i].b == 0) pos[p.leaves[i].a] = i;
  int extra = 0;
  for (int q = 0; q < npolys; ++q)
    if (pos[q] < 0) {
      tables.push_back(polys[q]);
      pos[q] = K + extra++;
    }
  std::vector<int32_t> ops(p.ops);
  if (extra)  // the appended tables shift the constant and temporary slots
    for (size_t i = 0; i < ops.size(); i += 4)
      for (int j = 1; j < 4; ++j)
        if (ops[i + j] >= K) ops[i + j] += extra;
  const int KT = K + extra;
  if (KT > 40) return B200_ERR_ARG;
  int4* d_ops = (int4*)ar.upload<int32_t>(ops.data(), ops.size());
  Fr* d_ev = ar.alloc<Fr>(KT);
  int* d_pos = ar.upload<int>(pos.data(), pos.size());
  if (!d_ops || !d_ev || !d_pos) return B200_ERR_NOMEM;
  GenericJob job;
  job.num_vars = num_vars;
  job.ntables = KT;
  job.nconsts = (int)p.consts.size();
  job.nops = (int)ops.size() / 4;
  job.degree = p.degree;
  job.ntemps = p.ntemps;
  for (int i = 0; i < KT; ++i) job.tables[i] = tables[i];
  job.consts = d_consts;
  job.ops = d_ops;
  job.claim = d_sum;
  job.challenges_out = d_x_out;
  job.evals_out = d_ev;
  rc = sumcheck_prove_generic(c, job);
  if (rc) return rc;
  gather_evals_kernel<<<(npolys + 63) / 64, 64, 0, c->stream>>>(d_ev, d_pos, npolys, d_evals_out);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// Expression::evaluate on every row (prover.rs:96-117) -> d_out
static int expression_rows(Ctx* c, int num_vars, const ExprP& expr, const Fr* const* polys, int npolys,
                           const Fr* d_challenges, Fr* d_out) {
  Arena ar(c);
  ExprCompiler comp;
  Program p = comp.compile(expr);
  if (p.leaves.empty()) return B200_ERR_ARG;
  std::vector<Fr> keep;
  Fr* d_consts = program_consts(c, ar, p, d_challenges, &keep);
  if (!d_consts) return B200_ERR_NOMEM;
  std::vector<const Fr*> tables;
  int rc = leaf_tables(c, ar, num_vars, p.leaves, polys, npolys, nullptr, 0, &tables);
  if (rc) return rc;
  int4* d_ops = (int4*)ar.upload<int32_t>(p.ops.data(), p.ops.size());
  if (!d_ops) return B200_ERR_NOMEM;
  rc = expression_rows_prog(c, num_vars, tables.data(), (int)tables.size(), d_consts, (int)p.consts.size(), d_ops,
                            (int)p.ops.size() / 4, p.ntemps, d_out);
  return rc;
}

struct HyperPlonk {
  Ctx* c;
  int k, num_witness, num_poly, num_z, chunk_size;
  std::vector<int> num_instances;                     // per instance column (pb/backend.rs:50-51)
  std::vector<int> phase_witness, phase_challenges;   // per witness phase (pb/backend.rs:55-60)
  std::vector<const Fr*> preprocess;  // device, borrowed from the caller
  std::vector<int> perm_idx;
  std::vector<Fr*> perm;              // device, owned
  std::vector<LookupCols> lookups;
  ExprP expression;
  std::vector<G1Aff> preprocess_comms, permutation_comms;  // host copies (the verifier parameters)
};

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_hyperplonk() {
  B200_PRELOAD(scatter_fr_kernel);
  B200_PRELOAD(u64_rows_to_fr_kernel);
  B200_PRELOAD(rotation_points_kernel);
  B200_PRELOAD(gather_evals_kernel);
}

}  // namespace b200

struct b200_hyperplonk {
  b200::HyperPlonk hp;
};

using namespace b200;

// every Polynomial / Challenge index of the tree is below the given counts
static bool indices_in_range(const ExprP& e, int npolys, int nchallenges) {
  if (e->kind == Expr::POLY && (e->a < 0 || e->a >= npolys)) return false;
  if (e->kind == Expr::CHALLENGE && (e->a < 0 || e->a >= nchallenges)) return false;
  for (auto& c : e->ch)
    if (!indices_in_range(c, npolys, nchallenges)) return false;
  return true;
}

static int commit_polys(Ctx* c, const std::vector<const Fr*>& polys, int k, bool write, G1Aff* d_out) {
  if (polys.empty()) return B200_OK;
  if ((int)c->srs.size() <= k) return B200_ERR_ARG;
  std::vector<MsmJob> jobs;
  for (auto* p : polys) jobs.push_back(MsmJob{p, c->srs[k], (uint64_t)1 << k, MSM_FR_MONT, 254, c->srs_ext[k]});
  return kzg_commit_batch(c, jobs.data(), (int)jobs.size(), write, d_out);
}

extern "C" {

int b200_expression_compile(const int32_t* tokens, int ntokens, const void* consts_fr, int nconsts, int32_t* leaves_out,
                            int leaves_cap, int* nleaves, void* consts_out, int32_t* const_chal_out, int consts_cap,
                            int* nconsts_out, int32_t* ops_out, int ops_cap, int* nops, int* ntemps, int* degree) {
  const int32_t* t = tokens;
  ExprP e = e_parse(t, tokens + ntokens, (const Fr*)consts_fr, nconsts);
  if (!e || t != tokens + ntokens) return B200_ERR_ARG;
  ExprCompiler comp;
  Program p = comp.compile(e);
  if ((int)p.leaves.size() > leaves_cap || (int)p.consts.size() > consts_cap || (int)p.ops.size() / 4 > ops_cap)
    return B200_ERR_NOMEM;
  for (size_t i = 0; i < p.leaves.size(); ++i) {
    leaves_out[3 * i] = p.leaves[i].kind;
    leaves_out[3 * i + 1] = p.leaves[i].a;
    leaves_out[3 * i + 2] = p.leaves[i].b;
  }
  for (size_t i = 0; i < p.consts.size(); ++i) {
    ((Fr*)consts_out)[i] = p.consts[i].value;
    const_chal_out[i] = p.consts[i].chal;
  }
  memcpy(ops_out, p.ops.data(), p.ops.size() * sizeof(int32_t));
  *nleaves = (int)p.leaves.size();
  *nconsts_out = (int)p.consts.size();
  *nops = (int)p.ops.size() / 4;
  *ntemps = p.ntemps;
  *degree = p.degree;
  return B200_OK;
}

// preprocessor.rs:25-60 as a host-only service: circuit expressions in, composed zero-check expression out
int b200_expression_compose(int k, int num_poly, int num_challenges, int nconstraints, const int32_t* constraint_tokens,
                            int nconstraint_tokens, int nlookups, const int32_t* lookup_tokens, int nlookup_tokens,
                            const void* consts_fr, int nconsts, int nperm, const int32_t* permutation_polys,
                            int max_degree, int32_t* tokens_out, int tokens_cap, int* ntokens, void* consts_out,
                            int consts_cap, int* nconsts_out, int* num_permutation_z_polys) {
  if (k < 1 || k > 30 || num_poly < 1 || num_challenges < 0 || nconstraints < 1 || !constraint_tokens || nlookups < 0 ||
      (nlookups && !lookup_tokens) || nconsts < 0 || (nconsts && !consts_fr) || nperm < 0 || nperm > 8 ||
      (nperm && !permutation_polys) || max_degree < 2 || !tokens_out || !ntokens || !consts_out || !nconsts_out ||
      !num_permutation_z_polys)
    return B200_ERR_ARG;
  std::vector<ExprP> constraints;
  const int32_t *t = constraint_tokens, *tend = constraint_tokens + nconstraint_tokens;
  for (int i = 0; i < nconstraints; ++i) {
    ExprP e = e_parse(t, tend, (const Fr*)consts_fr, nconsts);
    if (!e || !indices_in_range(e, num_poly, num_challenges)) return B200_ERR_ARG;
    constraints.push_back(e);
  }
  if (t != tend) return B200_ERR_ARG;
  std::vector<LookupCols> lookups;
  t = lookup_tokens;
  tend = lookup_tokens + nlookup_tokens;
  for (int l = 0; l < nlookups; ++l) {
    if (t >= tend) return B200_ERR_ARG;
    const int width = *t++;
    if (width < 1 || width > 64) return B200_ERR_ARG;
    LookupCols cols;
    for (int j = 0; j < width; ++j) {
      ExprP in = e_parse(t, tend, (const Fr*)consts_fr, nconsts);
      ExprP tb = in ? e_parse(t, tend, (const Fr*)consts_fr, nconsts) : nullptr;
      if (!tb || !indices_in_range(in, num_poly, num_challenges) || !indices_in_range(tb, num_poly, num_challenges))
        return B200_ERR_ARG;
      cols.push_back({in, tb});
    }
    lookups.push_back(cols);
  }
  std::vector<int> perm(permutation_polys, permutation_polys + nperm);
  for (int p : perm)
    if (p < 0 || p >= num_poly) return B200_ERR_ARG;
  int num_z = 0;
  ExprP composed = e_compose(k, constraints, num_poly, perm, num_challenges, max_degree, lookups, &num_z);
  std::vector<int32_t> tok;
  std::vector<Fr> cs;
  e_serialize(composed, &tok, &cs);
  *ntokens = (int)tok.size();
  *nconsts_out = (int)cs.size();
  *num_permutation_z_polys = num_z;
  if ((int)tok.size() > tokens_cap || (int)cs.size() > consts_cap) return B200_ERR_NOMEM;  // sizes reported above
  memcpy(tokens_out, tok.data(), tok.size() * sizeof(int32_t));
  if (!cs.empty()) memcpy(consts_out, cs.data(), cs.size() * sizeof(Fr));
  return B200_OK;
}

int b200_sumcheck_prove_expression(b200_ctx* h, int num_vars, const int32_t* tokens, int ntokens, const void* consts_fr,
                                   int nconsts, const void* const* dev_polys, int npolys, const void* host_challenges,
                                   int nchallenges, const void* host_ys, int nys, const void* host_sum,
                                   void* host_challenges_out, void* host_evals_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || npolys < 1 || npolys > 40 || nchallenges < 0 || nys < 0) return B200_ERR_ARG;
  const int32_t* t = tokens;
  ExprP e = e_parse(t, tokens + ntokens, (const Fr*)consts_fr, nconsts);
  if (!e || t != tokens + ntokens) return B200_ERR_ARG;
  Arena ar(c);
  const size_t nin = (size_t)nchallenges + (size_t)nys * num_vars + 1, nout = (size_t)num_vars + npolys;
  std::vector<Fr> in(nin);
  if (nchallenges) memcpy(in.data(), host_challenges, nchallenges * sizeof(Fr));
  if (nys) memcpy(in.data() + nchallenges, host_ys, (size_t)nys * num_vars * sizeof(Fr));
  memcpy(in.data() + nin - 1, host_sum, sizeof(Fr));
  Fr* d_in = ar.upload<Fr>(in.data(), nin);
  Fr* d_out = ar.alloc<Fr>(nout);
  if (!d_in || !d_out) return B200_ERR_NOMEM;
  int rc = prove_expression(c, num_vars, e, (const Fr* const*)dev_polys, npolys, d_in, d_in + nchallenges, nys,
                            d_in + nin - 1, d_out, d_out + num_vars);
  if (rc) return rc;
  std::vector<Fr> out(nout);
  CUDA_TRY(cudaMemcpyAsync(out.data(), d_out, nout * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(host_challenges_out, out.data(), num_vars * sizeof(Fr));
  memcpy(host_evals_out, out.data() + num_vars, npolys * sizeof(Fr));
  return B200_OK;
}

int b200_hyperplonk_preprocess_phased(b200_ctx* h, int k, int ninstance_cols, const int32_t* num_instances, int nphases,
                                      const int32_t* num_witness_polys, const int32_t* num_challenges, int npreprocess,
                                      const void* const* dev_preprocess, int nconstraints,
                                      const int32_t* constraint_tokens, int nconstraint_tokens, int nlookups,
                                      const int32_t* lookup_tokens, int nlookup_tokens, const void* consts_fr,
                                      int nconsts, int nperm, const int32_t* permutation_polys, int ncycles,
                                      const int32_t* cycles_flat, int max_degree, b200_hyperplonk** out) {
  Ctx* c = &h->c;
  if (k < 1 || k > 30 || (int)c->srs.size() <= k || ninstance_cols < 0 || ninstance_cols > 8 || nphases < 1 ||
      nphases > 8 || npreprocess < 0 || nconstraints < 1 || nperm < 0 || nperm > 8 || max_degree < 2)
    return B200_ERR_ARG;
  const size_t N = (size_t)1 << k;
  // PlonkishCircuitInfo::is_well_formed (pb/backend.rs:76-105): every phase has witness polynomials, every phase but
  // the last one has challenges; an instance column fits the rows bh[1..] (prover.rs:32-48)
  int total_witness = 0, total_challenges = 0;
  for (int i = 0; i < nphases; ++i) {
    if (num_witness_polys[i] < 1 || num_challenges[i] < 0 || (i + 1 < nphases && num_challenges[i] == 0)) return B200_ERR_ARG;
    total_witness += num_witness_polys[i];
    total_challenges += num_challenges[i];
  }
  for (int i = 0; i < ninstance_cols; ++i)
    if (num_instances[i] < 0 || (size_t)num_instances[i] + 1 > N) return B200_ERR_ARG;
  if (total_witness > 32 || total_challenges > 64) return B200_ERR_ARG;
  b200_hyperplonk* obj = new b200_hyperplonk();
  HyperPlonk& hp = obj->hp;
  hp.c = c;
  hp.k = k;
  hp.num_instances.assign(num_instances, num_instances + ninstance_cols);
  hp.phase_witness.assign(num_witness_polys, num_witness_polys + nphases);
  hp.phase_challenges.assign(num_challenges, num_challenges + nphases);
  hp.num_witness = total_witness;
  hp.num_poly = ninstance_cols + npreprocess + total_witness;
  for (int i = 0; i < npreprocess; ++i) hp.preprocess.push_back((const Fr*)dev_preprocess[i]);
  hp.perm_idx.assign(permutation_polys, permutation_polys + nperm);
  uint64_t* d_u64 = nullptr;
  auto fail = [&](int rc) {
    for (Fr* p : hp.perm) cudaFree(p);
    if (d_u64) cudaFree(d_u64);
    delete obj;
    return rc;
  };
  for (int i = 0; i < nperm; ++i)
    if (hp.perm_idx[i] < 0 || hp.perm_idx[i] >= hp.num_poly) return fail(B200_ERR_ARG);
  // constraints and lookups
  std::vector<ExprP> constraints;
  const int32_t *t = constraint_tokens, *tend = constraint_tokens + nconstraint_tokens;
  for (int i = 0; i < nconstraints; ++i) {
    ExprP e = e_parse(t, tend, (const Fr*)consts_fr, nconsts);
    if (!e) return fail(B200_ERR_ARG);
    constraints.push_back(e);
  }
  t = lookup_tokens;
  tend = lookup_tokens + nlookup_tokens;
  for (int l = 0; l < nlookups; ++l) {
    if (t >= tend) return fail(B200_ERR_ARG);
    const int width = *t++;
    LookupCols cols;
    for (int j = 0; j < width; ++j) {
      ExprP in = e_parse(t, tend, (const Fr*)consts_fr, nconsts);
      ExprP tb = in ? e_parse(t, tend, (const Fr*)consts_fr, nconsts) : nullptr;
      if (!tb) return fail(B200_ERR_ARG);
      cols.push_back({in, tb});
    }
    hp.lookups.push_back(cols);
  }
  hp.expression = e_compose(k, constraints, hp.num_poly, hp.perm_idx, total_challenges, max_degree, hp.lookups, &hp.num_z,
                            &hp.chunk_size);
  {  // polynomial and challenge indices in range (is_well_formed, pb/backend.rs:94-97)
    bool ok = true;
    for (auto& e : constraints) ok = ok && indices_in_range(e, hp.num_poly, total_challenges);
    for (auto& cols : hp.lookups)
      for (auto& col : cols)
        ok = ok && indices_in_range(col.first, hp.num_poly, total_challenges) &&
             indices_in_range(col.second, hp.num_poly, total_challenges);
    if (!ok) return fail(B200_ERR_ARG);
  }
  if (hp.num_z > 8 || hp.chunk_size > 8) return fail(B200_ERR_ARG);
  // permutation_polys (preprocessor.rs:172-203): identity (i << k) + j, then every cycle rotated by one
  std::vector<std::vector<uint64_t>> perms(nperm, std::vector<uint64_t>(N));
  std::map<int, int> index;
  for (int i = 0; i < nperm; ++i) {
    index[hp.perm_idx[i]] = i;
    for (size_t j = 0; j < N; ++j) perms[i][j] = ((uint64_t)i << k) + j;
  }
  const int32_t* cy = cycles_flat;
  for (int q = 0; q < ncycles; ++q) {
    const int len = *cy++;
    if (len < 1) return fail(B200_ERR_ARG);
    for (int s = 0; s < len; ++s)
      if (!index.count(cy[2 * s]) || cy[2 * s + 1] < 0 || (size_t)cy[2 * s + 1] >= N) return fail(B200_ERR_ARG);
    uint64_t last = perms[index[cy[0]]][cy[1]];
    for (int s = 1; s <= len; ++s) {
      const int32_t* ij = cy + 2 * (s % len);
      std::swap(perms[index[ij[0]]][ij[1]], last);
    }
    cy += 2 * len;
  }
  if (nperm) {
    if (cudaMalloc(&d_u64, N * sizeof(uint64_t)) != cudaSuccess) return fail(B200_ERR_NOMEM);
    for (int i = 0; i < nperm; ++i) {
      Fr* d = nullptr;
      if (cudaMalloc(&d, N * sizeof(Fr)) != cudaSuccess) return fail(B200_ERR_NOMEM);
      hp.perm.push_back(d);
      cudaMemcpyAsync(d_u64, perms[i].data(), N * sizeof(uint64_t), cudaMemcpyHostToDevice, c->stream);
      u64_rows_to_fr_kernel<<<2 * NUM_SMS, 256, 0, c->stream>>>(d_u64, N, d);
      count_launch(c);
      if (cudaStreamSynchronize(c->stream) != cudaSuccess) return fail(B200_ERR_CUDA);
    }
    cudaFree(d_u64);
    d_u64 = nullptr;
  }
  // commitments (hyperplonk.rs:127-150)
  const int ncomm = npreprocess + nperm;
  if (ncomm) {
    G1Aff* d_comms = nullptr;
    if (cudaMalloc(&d_comms, ncomm * sizeof(G1Aff)) != cudaSuccess) return fail(B200_ERR_NOMEM);
    std::vector<const Fr*> all(hp.preprocess);
    for (Fr* p : hp.perm) all.push_back(p);
    int rc = commit_polys(c, all, k, false, d_comms);
    std::vector<G1Aff> hc(ncomm);
    if (!rc && cudaMemcpyAsync(hc.data(), d_comms, ncomm * sizeof(G1Aff), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess)
      rc = B200_ERR_CUDA;
    if (!rc && cudaStreamSynchronize(c->stream) != cudaSuccess) rc = B200_ERR_CUDA;
    cudaFree(d_comms);
    if (rc) return fail(rc);
    hp.preprocess_comms.assign(hc.begin(), hc.begin() + npreprocess);
    hp.permutation_comms.assign(hc.begin() + npreprocess, hc.end());
  }
  *out = obj;
  return B200_OK;
}

// one instance column, one witness phase without challenges (the reference's own test circuits, util.rs:30-98)
int b200_hyperplonk_preprocess(b200_ctx* h, int k, int num_instances, int num_witness_polys, int npreprocess,
                               const void* const* dev_preprocess, int nconstraints, const int32_t* constraint_tokens,
                               int nconstraint_tokens, int nlookups, const int32_t* lookup_tokens, int nlookup_tokens,
                               const void* consts_fr, int nconsts, int nperm, const int32_t* permutation_polys,
                               int ncycles, const int32_t* cycles_flat, int max_degree, b200_hyperplonk** out) {
  const int32_t ni = num_instances, nw = num_witness_polys, nc = 0;
  return b200_hyperplonk_preprocess_phased(h, k, 1, &ni, 1, &nw, &nc, npreprocess, dev_preprocess, nconstraints,
                                           constraint_tokens, nconstraint_tokens, nlookups, lookup_tokens, nlookup_tokens,
                                           consts_fr, nconsts, nperm, permutation_polys, ncycles, cycles_flat, max_degree,
                                           out);
}

void b200_hyperplonk_free(b200_hyperplonk* obj) {
  if (!obj) return;
  for (Fr* p : obj->hp.perm) cudaFree(p);
  delete obj;
}

int b200_hyperplonk_info(const b200_hyperplonk* obj, int* num_permutation_z_polys, int* degree, int* num_polys) {
  *num_permutation_z_polys = obj->hp.num_z;
  *degree = e_degree(obj->hp.expression);
  *num_polys = obj->hp.num_poly + (int)obj->hp.perm.size() + 2 * (int)obj->hp.lookups.size() + obj->hp.num_z;
  return B200_OK;
}

int b200_hyperplonk_commitments(const b200_hyperplonk* obj, void* preprocess_out, void* permutation_out) {
  const HyperPlonk& hp = obj->hp;
  if (!hp.preprocess_comms.empty()) memcpy(preprocess_out, hp.preprocess_comms.data(), hp.preprocess_comms.size() * sizeof(G1Aff));
  if (!hp.permutation_comms.empty()) memcpy(permutation_out, hp.permutation_comms.data(), hp.permutation_comms.size() * sizeof(G1Aff));
  return B200_OK;
}

int b200_hyperplonk_permutation_poly(const b200_hyperplonk* obj, int i, void* host_out) {
  const HyperPlonk& hp = obj->hp;
  if (i < 0 || i >= (int)hp.perm.size()) return B200_ERR_ARG;
  CUDA_TRY(cudaMemcpyAsync(host_out, hp.perm[i], ((size_t)1 << hp.k) * sizeof(Fr), cudaMemcpyDeviceToHost, hp.c->stream));
  CUDA_TRY(cudaStreamSynchronize(hp.c->stream));
  return B200_OK;
}

// hyperplonk.rs:164-291; appends to the context transcript
int b200_hyperplonk_prove_phased(b200_hyperplonk* obj, const void* host_instances_fr, int ninstances,
                                 b200_synthesize_fn synthesize, void* user) {
  HyperPlonk& hp = obj->hp;
  Ctx* c = hp.c;
  cudaStream_t s = c->stream;
  const int k = hp.k, nwit = hp.num_witness, nlk = (int)hp.lookups.size(), nper = (int)hp.perm.size();
  const int ncols = (int)hp.num_instances.size(), nphases = (int)hp.phase_witness.size();
  const size_t N = (size_t)1 << k;
  int ninst_total = 0, nc = 0;  // nc: the circuit's own challenges; beta, gamma, alpha follow them (preprocessor.rs:28-30)
  for (int n : hp.num_instances) ninst_total += n;
  for (int n : hp.phase_challenges) nc += n;
  if (ninstances != ninst_total || !synthesize) return B200_ERR_ARG;
  Arena ar(c);
  int rc;
  // instances: absorbed column by column, then the instance polynomials (prover.rs:32-48): instance i of a column
  // sits on row bh[i + 1]
  Fr* d_inst = ar.upload<Fr>((const Fr*)host_instances_fr, ninstances);
  std::vector<uint64_t> rows(ninstances ? ninstances : 1);
  {
    int o = 0;
    for (int n : hp.num_instances) {
      uint64_t b = 1;
      for (int i = 0; i < n; ++i) {
        rows[o++] = b;
        b = bh_next(b, k);
      }
    }
  }
  uint64_t* d_rows = ar.upload<uint64_t>(rows.data(), rows.size());
  if (!d_inst || !d_rows) return B200_ERR_NOMEM;
  rc = transcript_op(c, TR_COMMON, d_inst, nullptr, ninstances);
  if (rc) return rc;
  std::vector<const Fr*> polys;
  {
    int o = 0;
    for (int n : hp.num_instances) {
      Fr* inst_poly = ar.alloc<Fr>(N);
      if (!inst_poly) return B200_ERR_NOMEM;
      CUDA_TRY(cudaMemsetAsync(inst_poly, 0, N * sizeof(Fr), s));
      if (n) {
        scatter_fr_kernel<<<(n + 63) / 64, 64, 0, s>>>(d_inst + o, d_rows + o, n, inst_poly);
        count_launch(c);
      }
      polys.push_back(inst_poly);
      o += n;
    }
  }
  polys.insert(polys.end(), hp.preprocess.begin(), hp.preprocess.end());
  // rounds 0..n (hyperplonk.rs:183-204): synthesize the phase's witness polynomials (the callback sees the challenges
  // squeezed so far — the one host read-back of the phase loop), commit, squeeze the phase's challenges
  G1Aff* d_comms = ar.alloc<G1Aff>(nwit + 2 * nlk + hp.num_z + 1);
  Fr* d_chal = ar.alloc<Fr>(nc + 3 + k);  // circuit challenges, beta, gamma, alpha, y[k]
  Fr* d_zero = ar.alloc<Fr>(1);
  if (!d_comms || !d_chal || !d_zero) return B200_ERR_NOMEM;
  Fr* d_ch = d_chal + nc;  // beta, gamma, alpha, y[k]
  CUDA_TRY(cudaMemsetAsync(d_zero, 0, sizeof(Fr), s));
  {
    std::vector<Fr> host_chal(nc ? nc : 1);
    int have = 0;
    for (int round = 0; round < nphases; ++round) {
      const int nw = hp.phase_witness[round];
      NvtxRange nvtx_w("witness_collector-%d", round);  // hyperplonk.rs:192
      if (have) {
        CUDA_TRY(cudaMemcpyAsync(host_chal.data(), d_chal, have * sizeof(Fr), cudaMemcpyDeviceToHost, s));
        CUDA_TRY(cudaStreamSynchronize(s));
      }
      std::vector<const void*> dev(nw, nullptr);
      if (synthesize(user, round, host_chal.data(), have, dev.data()) != 0) return B200_ERR_ARG;
      std::vector<const Fr*> wit(nw);
      for (int i = 0; i < nw; ++i) {
        if (!dev[i]) return B200_ERR_ARG;
        wit[i] = (const Fr*)dev[i];
      }
      rc = commit_polys(c, wit, k, true, d_comms);
      if (rc) return rc;
      polys.insert(polys.end(), wit.begin(), wit.end());
      if (hp.phase_challenges[round]) {
        rc = transcript_op(c, TR_SQUEEZE, nullptr, d_chal + have, hp.phase_challenges[round]);
        if (rc) return rc;
        have += hp.phase_challenges[round];
      }
    }
  }
  // round n: beta; LogUp compressed polys and multiplicities (prover.rs:50-192)
  rc = transcript_op(c, TR_SQUEEZE, nullptr, d_ch, 1);
  if (rc) return rc;
  std::vector<Fr*> comp_in(nlk), comp_tab(nlk), ms(nlk), hs(nlk);
  nvtxRangePushA("lookup_compressed_polys+lookup_m_polys");  // hyperplonk.rs:215,223
  struct Pop {
    bool armed = true;
    ~Pop() { if (armed) nvtxRangePop(); }
  } pop_lookup;
  for (int l = 0; l < nlk; ++l) {
    comp_in[l] = ar.alloc<Fr>(N);
    comp_tab[l] = ar.alloc<Fr>(N);
    ms[l] = ar.alloc<Fr>(N);
    hs[l] = ar.alloc<Fr>(N);
    if (!comp_in[l] || !comp_tab[l] || !ms[l] || !hs[l]) return B200_ERR_NOMEM;
    std::vector<ExprP> ins, tabs;
    for (auto& col : hp.lookups[l]) {
      ins.push_back(col.first);
      tabs.push_back(col.second);
    }
    // Σ_j beta^j column_j (prover.rs:78-134): beta is challenge `nc` of this little program, behind the circuit's own
    rc = expression_rows(c, k, e_distribute_powers(ins, e_chal(nc)), polys.data(), (int)polys.size(), d_chal, comp_in[l]);
    if (rc) return rc;
    rc = expression_rows(c, k, e_distribute_powers(tabs, e_chal(nc)), polys.data(), (int)polys.size(), d_chal, comp_tab[l]);
    if (rc) return rc;
    rc = lookup_m(c, k, comp_in[l], comp_tab[l], ms[l]);
    if (rc) return rc;
  }
  nvtxRangePop();
  pop_lookup.armed = false;
  if (nlk) {
    rc = commit_polys(c, std::vector<const Fr*>(ms.begin(), ms.end()), k, true, d_comms);
    if (rc) return rc;
  }
  // round n+1: gamma; h polys, permutation z
  rc = transcript_op(c, TR_SQUEEZE, nullptr, d_ch + 1, 1);
  if (rc) return rc;
  for (int l = 0; l < nlk; ++l) {
    NvtxRange nvtx_h("lookup_h_polys-%d", nlk);  // hyperplonk.rs:233
    rc = lookup_h(c, k, comp_in[l], comp_tab[l], ms[l], d_ch + 1, hs[l]);
    if (rc) return rc;
  }
  std::vector<const Fr*> hz(hs.begin(), hs.end());
  if (hp.num_z) {
    std::vector<Fr*> zs(hp.num_z);
    for (auto& z : zs) {
      z = ar.alloc<Fr>(N);
      if (!z) return B200_ERR_NOMEM;
    }
    std::vector<const Fr*> wires(nper), sigmas(nper);
    std::vector<uint64_t> offs(nper);
    for (int i = 0; i < nper; ++i) {
      wires[i] = polys[hp.perm_idx[i]];
      sigmas[i] = hp.perm[i];
      offs[i] = (uint64_t)i << k;
    }
    {
      NvtxRange nvtx_z("permutation_z_polys-%d", nper);  // hyperplonk.rs:237
      rc = permutation_z_chunks(c, k, hp.num_z, hp.chunk_size, nper, wires.data(), sigmas.data(), offs.data(), d_ch, zs.data());
    }
    if (rc) return rc;
    for (Fr* z : zs) hz.push_back(z);
  }
  rc = commit_polys(c, hz, k, true, d_comms);
  if (rc) return rc;
  // round n+2: alpha, y; zero check (prover.rs:348-387)
  rc = transcript_op(c, TR_SQUEEZE, nullptr, d_ch + 2, 1 + k);
  if (rc) return rc;
  for (Fr* p : hp.perm) polys.push_back(p);
  for (Fr* p : ms) polys.push_back(p);
  for (auto* p : hz) polys.push_back(p);
  const int npolys = (int)polys.size();
  Fr* d_x = ar.alloc<Fr>(k);
  Fr* d_evals = ar.alloc<Fr>(npolys);
  if (!d_x || !d_evals) return B200_ERR_NOMEM;
  rc = prove_expression(c, k, hp.expression, polys.data(), npolys, d_chal, d_ch + 3, 1, d_zero, d_x, d_evals);
  if (rc) return rc;
  // pcs_query / points / evaluations (prover.rs:388-409, verifier.rs:147-182): queries in BTreeSet order
  std::vector<Leaf> leaves;
  e_leaves(hp.expression, &leaves);
  std::set<std::pair<int, int>> queries;
  std::set<int> rotations;
  for (auto& l : leaves)
    if (l.kind == Expr::POLY && l.a >= ncols) {  // instance polynomials: evaluated by the verifier itself (verifier.rs:92-145)
      queries.insert({l.a, l.b});
      rotations.insert(l.b);
    }
  std::map<int, int> offset;
  int npoints = 0;
  for (int r : rotations) {
    offset[r] = npoints;
    npoints += 1 << std::abs(r);
  }
  Fr* d_points = ar.alloc<Fr>((size_t)npoints * k);
  if (!d_points) return B200_ERR_NOMEM;
  std::vector<std::vector<uint64_t>> pattern_keep;
  for (int r : rotations) {
    Fr* dst = d_points + (size_t)offset[r] * k;
    if (r == 0) {
      CUDA_TRY(cudaMemcpyAsync(dst, d_x, k * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
      continue;
    }
    pattern_keep.push_back(rotation_eval_point_pattern(r > 0, k, std::abs(r)));
    const int np = (int)pattern_keep.back().size();
    uint64_t* d_pat = ar.upload<uint64_t>(pattern_keep.back().data(), np);
    if (!d_pat) return B200_ERR_NOMEM;
    rotation_points_kernel<<<(np * k + 127) / 128, 128, 0, s>>>(d_x, k, r, d_pat, np, dst);
    count_launch(c);
  }
  std::vector<int> ev_poly, ev_point;
  for (auto& q : queries)
    for (int j = 0; j < (1 << std::abs(q.second)); ++j) {
      ev_poly.push_back(q.first);
      ev_point.push_back(offset[q.second] + j);
    }
  const int nevals = (int)ev_poly.size();
  Fr* d_vals = ar.alloc<Fr>(nevals);
  if (!d_vals) return B200_ERR_NOMEM;
  {
    NvtxRange nvtx_e("evals-%d", nevals);  // prover.rs:391
    int e = 0;
    for (auto& q : queries)
      for (int j = 0; j < (1 << std::abs(q.second)); ++j, ++e) {
        if (q.second == 0) {
          CUDA_TRY(cudaMemcpyAsync(d_vals + e, d_evals + q.first, sizeof(Fr), cudaMemcpyDeviceToDevice, s));
        } else {  // evaluate_for_rotation == plain evaluations at the rotation_eval_points (multilinear.rs:191-263)
          const Fr* tab[1] = {polys[q.first]};
          rc = mle_eval_many(c, tab, 1, k, d_points + (size_t)ev_point[e] * k, d_vals + e);
          if (rc) return rc;
        }
      }
  }
  rc = transcript_op(c, TR_WRITE, d_vals, nullptr, nevals);
  if (rc) return rc;
  BatchOpenJob job;
  job.num_vars = k;
  job.npolys = npolys;
  job.npoints = npoints;
  job.nevals = nevals;
  job.polys = polys.data();
  job.points = d_points;
  job.ev_poly = ev_poly.data();
  job.ev_point = ev_point.data();
  job.ev_values = d_vals;
  rc = kzg_batch_open(c, job);
  if (rc) return rc;
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// single-phase circuits: the witness is known up front
static int witness_given(void* user, int, const void*, int, const void** dev_witness_out) {
  auto* w = (std::pair<const void* const*, int>*)user;
  for (int i = 0; i < w->second; ++i) dev_witness_out[i] = w->first[i];
  return 0;
}
int b200_hyperplonk_prove(b200_hyperplonk* obj, const void* host_instances_fr, int ninstances,
                          const void* const* dev_witness) {
  if (obj->hp.phase_witness.size() != 1) return B200_ERR_ARG;
  std::pair<const void* const*, int> w{dev_witness, obj->hp.num_witness};
  return b200_hyperplonk_prove_phased(obj, host_instances_fr, ninstances, witness_given, &w);
}

}  // extern "C"

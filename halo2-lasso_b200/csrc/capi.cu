// extern "C" boundary (include/b200_lasso.h): context, device polynomials, transcript, MLE and
// sum-check entry points. Host buffers are staged with stream-ordered copies; the only host syncs
// are the ones that return results to the caller.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/b200_lasso.h"
#include "internal.h"

using namespace b200;

static const uint32_t PROOF_CAP = 8u << 20;
static const size_t PARTIAL_ELEMS = 1u << 16;

__global__ void tr_init_kernel(Transcript* tr, uint8_t* proof, uint32_t cap) {
  if (threadIdx.x == 0 && blockIdx.x == 0) tr_init(tr, proof, cap);
}

static void host_bary(BaryTable* t) {
  memset(t, 0, sizeof(*t));
  for (int d = 1; d <= 6; ++d)
    for (int i = 0; i <= d; ++i) {
      Fr acc = fe_one<FrP>();
      for (int j = 0; j <= d; ++j) {
        if (j == i) continue;
        Fr fi = fe_from_u64<FrP>((uint64_t)i), fj = fe_from_u64<FrP>((uint64_t)j);
        acc = acc * (fi - fj);
      }
      t->w[d][i] = fe_inv<FrP>(acc);
    }
}

// stage `n` host field elements into a fresh device buffer (freed by the caller with cudaFreeAsync)
static int stage(Ctx* c, const void* host, size_t bytes, void** dev) {
  CUDA_TRY(cudaMallocAsync(dev, bytes ? bytes : 32, c->stream));
  if (bytes) CUDA_TRY(cudaMemcpyAsync(*dev, host, bytes, cudaMemcpyHostToDevice, c->stream));
  return B200_OK;
}
static int fetch(Ctx* c, void* host, const void* dev, size_t bytes) {
  CUDA_TRY(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return B200_OK;
}
static int check_transcript(Ctx* c) {
  Transcript t;
  CUDA_TRY(cudaMemcpyAsync(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return t.error ? B200_ERR_TRANSCRIPT : B200_OK;
}

static void preload_all_kernels() {
  B200_PRELOAD(tr_init_kernel);
  preload_generic();
  preload_gkr();
  preload_hyperplonk();
  preload_kzg();
  preload_lasso();
  preload_lookup();
  preload_mle();
  preload_msm();
  preload_perm();
  preload_shard();
  preload_sumcheck();
  cudaGetLastError();
}

extern "C" {

int b200_ctx_create(int device, b200_ctx** out) {
  if (!out) return B200_ERR_ARG;
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return B200_ERR_ARG;
  CUDA_TRY(cudaSetDevice(device));
  preload_all_kernels();  // per device: the load is a no-op once a kernel is resident
  b200_ctx* h = new b200_ctx();
  Ctx* c = &h->c;
  c->device = device;
  c->launches = 0;
  c->profile = false;
  memset(&c->peer, 0, sizeof(c->peer));
  c->peer.rank = 0;
  c->peer.world = 1;
  c->my_mailbox = nullptr;
  c->peer_seq = 0;
  c->dbg_clocks = nullptr;
  if (getenv("B200_DEBUG_CLOCKS")) {
    CUDA_TRY(cudaMalloc(&c->dbg_clocks, 32 * 16 * sizeof(long long)));
    CUDA_TRY(cudaMemset(c->dbg_clocks, 0, 32 * 16 * sizeof(long long)));
  }
  CUDA_TRY(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  for (int kind = 1; kind <= 2; ++kind) {  // value-sorted address lists of the and / xor subtables (grouped E commitments)
    std::vector<uint32_t> vals((size_t)1 << 16), perm, off;
    for (uint32_t x = 0; x < (1u << 16); ++x) vals[x] = kind == 1 ? ((x >> 8) & (x & 0xff)) : ((x >> 8) ^ (x & 0xff));
    if (!lasso_group_lists(vals.data(), &perm, &off)) return B200_ERR_ARG;
    CUDA_TRY(cudaMalloc(&c->d_group_perm[kind], perm.size() * 4));
    CUDA_TRY(cudaMalloc(&c->d_group_off[kind], off.size() * 4));
    CUDA_TRY(cudaMemcpy(c->d_group_perm[kind], perm.data(), perm.size() * 4, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(c->d_group_off[kind], off.data(), off.size() * 4, cudaMemcpyHostToDevice));
  }
  // keep freed blocks cached in the stream-ordered pool: proofs reuse the same sizes over and over
  cudaMemPool_t pool;
  CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  CUDA_TRY(cudaMalloc(&c->d_tr, sizeof(Transcript)));
  CUDA_TRY(cudaMalloc(&c->d_proof, PROOF_CAP));
  c->proof_cap = PROOF_CAP;
  CUDA_TRY(cudaMalloc(&c->d_bary, sizeof(BaryTable)));
  CUDA_TRY(cudaMalloc(&c->d_sc, sizeof(ScState)));
  CUDA_TRY(cudaMemset(c->d_sc, 0, sizeof(ScState)));
  CUDA_TRY(cudaMalloc(&c->d_partial, PARTIAL_ELEMS * sizeof(Fr)));
  c->partial_elems = PARTIAL_ELEMS;
  BaryTable bt;
  host_bary(&bt);
  CUDA_TRY(cudaMemcpy(c->d_bary, &bt, sizeof(bt), cudaMemcpyHostToDevice));
  tr_init_kernel<<<1, 32, 0, c->stream>>>(c->d_tr, c->d_proof, c->proof_cap);
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  *out = h;
  return B200_OK;
}

void b200_ctx_destroy(b200_ctx* h) {
  if (!h) return;
  Ctx* c = &h->c;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto p : c->srs) cudaFree(p);
  for (auto p : c->srs_ext) cudaFree(p);
  for (int r = 0; r < c->peer.world; ++r)
    if (c->peer_ipc && r != c->peer.rank) {
      cudaIpcCloseMemHandle(c->peer.box[r]);
      cudaIpcCloseMemHandle(c->peer.arena[r]);
    }
  if (c->my_mailbox) cudaFree(c->my_mailbox);
  if (c->my_arena) cudaFree(c->my_arena);
  if (c->d_peer_err) cudaFree(c->d_peer_err);
  for (int kind = 1; kind <= 2; ++kind) {
    cudaFree(c->d_group_perm[kind]);
    cudaFree(c->d_group_off[kind]);
  }
  if (c->hb_stream) cudaStreamDestroy(c->hb_stream);
  if (c->hb_event) cudaEventDestroy(c->hb_event);
  if (c->hb_stop) cudaFree(c->hb_stop);
  cudaFree(c->d_tr);
  cudaFree(c->d_proof);
  cudaFree(c->d_bary);
  cudaFree(c->d_sc);
  cudaFree(c->d_partial);
  cudaStreamDestroy(c->stream);
  delete h;
}

int b200_sync(b200_ctx* h) {
  CUDA_TRY(cudaStreamSynchronize(h->c.stream));
  return B200_OK;
}

uint64_t b200_launch_count(b200_ctx* h, int reset) {
  uint64_t v = h->c.launches;
  if (reset) h->c.launches = 0;
  return v;
}

void* b200_stream(b200_ctx* h) { return (void*)h->c.stream; }

int b200_sumcheck_eq_factored(b200_ctx* h, int on) {
  h->c.eq_factored = on != 0;
  return B200_OK;
}

int b200_debug_clocks(b200_ctx* h, long long* out) {
  Ctx* c = &h->c;
  if (!c->dbg_clocks) return B200_ERR_ARG;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(out, c->dbg_clocks, 32 * 16 * sizeof(long long), cudaMemcpyDeviceToHost));
  return B200_OK;
}
int b200_profile_enable(b200_ctx* h, int on) {
  Ctx* c = &h->c;
  for (auto e : c->prof_events) cudaEventDestroy(e);
  c->prof_events.clear();
  c->prof_tags.clear();
  c->profile = on != 0;
  return B200_OK;
}
int b200_profile_read(b200_ctx* h, float* ms, int* tags, int cap, int* n) {
  Ctx* c = &h->c;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  int cnt = (int)c->prof_tags.size();
  if (n) *n = cnt;
  for (int i = 0; i < cnt && i < cap; ++i) {
    CUDA_TRY(cudaEventElapsedTime(&ms[i], c->prof_events[2 * i], c->prof_events[2 * i + 1]));
    tags[i] = c->prof_tags[i];
  }
  return B200_OK;
}

// ---- device polynomials ---------------------------------------------------------------------
int b200_poly_alloc(b200_ctx* h, uint64_t len, void** dev) {
  CUDA_TRY(cudaMallocAsync(dev, (len ? len : 1) * sizeof(Fr), h->c.stream));
  return B200_OK;
}
int b200_poly_write(b200_ctx* h, void* dev, const void* host_fr, uint64_t len) {
  CUDA_TRY(cudaMemcpyAsync(dev, host_fr, len * sizeof(Fr), cudaMemcpyHostToDevice, h->c.stream));
  return B200_OK;
}
int b200_poly_upload(b200_ctx* h, const void* host_fr, uint64_t len, void** dev) {
  int rc = b200_poly_alloc(h, len, dev);
  if (rc) return rc;
  return b200_poly_write(h, *dev, host_fr, len);
}
int b200_poly_download(b200_ctx* h, const void* dev, uint64_t len, void* host_fr) {
  return fetch(&h->c, host_fr, dev, len * sizeof(Fr));
}
int b200_poly_free(b200_ctx* h, void* dev) {
  CUDA_TRY(cudaFreeAsync(dev, h->c.stream));
  return B200_OK;
}
int b200_fr_convert(b200_ctx* h, const void* dev_in, void* dev_out, uint64_t len, int to_mont) {
  return fr_convert(&h->c, (const Fr*)dev_in, (Fr*)dev_out, len, to_mont);
}

// ---- transcript -----------------------------------------------------------------------------
int b200_transcript_reset(b200_ctx* h) {
  Ctx* c = &h->c;
  tr_init_kernel<<<1, 32, 0, c->stream>>>(c->d_tr, c->d_proof, c->proof_cap);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
static int tr_host_op(b200_ctx* h, int op, const void* host_fr, int n) {
  Ctx* c = &h->c;
  if (n < 0) return B200_ERR_ARG;
  if (n == 0) return B200_OK;
  void* d = nullptr;
  int rc = stage(c, host_fr, (size_t)n * sizeof(Fr), &d);
  if (rc) return rc;
  rc = transcript_op(c, op, (const Fr*)d, nullptr, n);
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return rc;
}
int b200_transcript_common_field_elements(b200_ctx* h, const void* host_fr, int n) {
  return tr_host_op(h, TR_COMMON, host_fr, n);
}
int b200_transcript_write_field_elements(b200_ctx* h, const void* host_fr, int n) {
  return tr_host_op(h, TR_WRITE, host_fr, n);
}
int b200_transcript_squeeze_challenges(b200_ctx* h, int n, void* host_fr_out) {
  Ctx* c = &h->c;
  if (n < 0) return B200_ERR_ARG;
  if (n == 0) return B200_OK;
  Fr* d = nullptr;
  CUDA_TRY(cudaMallocAsync(&d, (size_t)n * sizeof(Fr), c->stream));
  int rc = transcript_op(c, TR_SQUEEZE, nullptr, d, n);
  if (rc) return rc;
  rc = fetch(c, host_fr_out, d, (size_t)n * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return rc;
}
int b200_transcript_write_commitments(b200_ctx* h, const void* host_g1, int n) {
  Ctx* c = &h->c;
  if (n < 0) return B200_ERR_ARG;
  if (n == 0) return B200_OK;
  void* d = nullptr;
  int rc = stage(c, host_g1, (size_t)n * sizeof(G1Aff), &d);
  if (rc) return rc;
  rc = transcript_write_points(c, (const G1Aff*)d, n);
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  if (rc) return rc;
  return check_transcript(c);
}
int b200_transcript_proof(b200_ctx* h, uint8_t* out, uint64_t cap, uint64_t* len) {
  Ctx* c = &h->c;
  Transcript t;
  CUDA_TRY(cudaMemcpyAsync(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (t.error) return B200_ERR_TRANSCRIPT;
  if (len) *len = t.proof_len;
  if (t.proof_len > cap) return B200_ERR_NOMEM;
  if (t.proof_len) return fetch(c, out, c->d_proof, t.proof_len);
  return B200_OK;
}

// ---- MLE ------------------------------------------------------------------------------------
int b200_eq_xy(b200_ctx* h, const void* host_y, int n, void* dev_out) {
  Ctx* c = &h->c;
  if (n < 1 || n > 30) return B200_ERR_ARG;
  void* dy = nullptr;
  int rc = stage(c, host_y, (size_t)n * sizeof(Fr), &dy);
  if (rc) return rc;
  rc = eq_build(c, (const Fr*)dy, n, (Fr*)dev_out);
  CUDA_TRY(cudaFreeAsync(dy, c->stream));
  return rc;
}
int b200_fix_var(b200_ctx* h, const void* dev_in, int n, const void* host_r, void* dev_out) {
  Ctx* c = &h->c;
  void* dr = nullptr;
  int rc = stage(c, host_r, sizeof(Fr), &dr);
  if (rc) return rc;
  rc = fix_var(c, (const Fr*)dev_in, n, (const Fr*)dr, (Fr*)dev_out);
  CUDA_TRY(cudaFreeAsync(dr, c->stream));
  return rc;
}
int b200_evaluate(b200_ctx* h, const void* const* dev_tables, int ntables, int n, const void* host_x,
                  void* host_fr_out) {
  Ctx* c = &h->c;
  if (ntables < 1 || ntables > SC_MAX_TABLES || n < 1 || n > 30) return B200_ERR_ARG;
  void* dx = nullptr;
  Fr* dout = nullptr;
  int rc = stage(c, host_x, (size_t)n * sizeof(Fr), &dx);
  if (rc) return rc;
  CUDA_TRY(cudaMallocAsync(&dout, ntables * sizeof(Fr), c->stream));
  rc = mle_eval_many(c, (const Fr* const*)dev_tables, ntables, n, (const Fr*)dx, dout);
  if (rc) return rc;
  rc = fetch(c, host_fr_out, dout, ntables * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(dx, c->stream));
  CUDA_TRY(cudaFreeAsync(dout, c->stream));
  return rc;
}

// ---- sum-check ------------------------------------------------------------------------------
int b200_sumcheck_prove_evals(b200_ctx* h, int num_vars, int nterms, int np,
                              const void* const* dev_tables, const void* host_weights,
                              const void* host_y, const void* host_sum, void* host_challenges_out,
                              void* host_evals_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || nterms < 1 || nterms > SC_MAX_TERMS || (np != 1 && np != 2))
    return B200_ERR_ARG;
  const int ntab = nterms * np;
  // one staging buffer: weights | y | sum | challenges | evals
  const size_t nin = (size_t)nterms + num_vars + 1, nout = (size_t)num_vars + ntab;
  std::vector<Fr> hbuf(nin);
  memcpy(hbuf.data(), host_weights, nterms * sizeof(Fr));
  memcpy(hbuf.data() + nterms, host_y, num_vars * sizeof(Fr));
  memcpy(hbuf.data() + nterms + num_vars, host_sum, sizeof(Fr));
  Fr* d = nullptr;
  CUDA_TRY(cudaMallocAsync(&d, (nin + nout) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(d, hbuf.data(), nin * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  ScEvalJob job;
  job.num_vars = num_vars;
  job.T = nterms;
  job.NP = np;
  for (int i = 0; i < ntab; ++i) job.tables[i] = (const Fr*)dev_tables[i];
  job.weights = d;
  job.eq_point = d + nterms;
  job.claim = d + nterms + num_vars;
  job.challenges_out = d + nin;
  job.evals_out = d + nin + num_vars;
  int rc = sumcheck_prove_evals(c, job);
  if (rc) return rc;
  std::vector<Fr> hout(nout);
  rc = fetch(c, hout.data(), d + nin, nout * sizeof(Fr));
  if (rc) return rc;
  memcpy(host_challenges_out, hout.data(), num_vars * sizeof(Fr));
  memcpy(host_evals_out, hout.data() + num_vars, ntab * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return B200_OK;
}

int b200_sumcheck_prove_evals_host(b200_ctx* h, int num_vars, int nterms, int np,
                                   const void* const* host_tables, const void* host_weights,
                                   const void* host_y, const void* host_sum,
                                   void* host_challenges_out, void* host_evals_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || nterms < 1 || nterms > SC_MAX_TERMS || (np != 1 && np != 2))
    return B200_ERR_ARG;
  const int ntab = nterms * np;
  const size_t N = (size_t)1 << num_vars;
  Fr* dt = nullptr;
  CUDA_TRY(cudaMallocAsync(&dt, (size_t)ntab * N * sizeof(Fr), c->stream));
  std::vector<const void*> ptrs(ntab);
  for (int i = 0; i < ntab; ++i) {
    CUDA_TRY(cudaMemcpyAsync(dt + (size_t)i * N, host_tables[i], N * sizeof(Fr), cudaMemcpyHostToDevice,
                             c->stream));
    ptrs[i] = dt + (size_t)i * N;
  }
  int rc = b200_sumcheck_prove_evals(h, num_vars, nterms, np, ptrs.data(), host_weights, host_y, host_sum,
                                     host_challenges_out, host_evals_out);
  CUDA_TRY(cudaFreeAsync(dt, c->stream));
  return rc;
}

int b200_sumcheck_prove_coeffs(b200_ctx* h, int num_vars, int nprods, const void* const* dev_tables,
                               const void* host_scalars, const void* host_ys, const void* host_sum,
                               void* host_challenges_out, void* host_evals_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || nprods < 1 || nprods > SC_MAX_TERMS) return B200_ERR_ARG;
  const size_t nin = (size_t)nprods + (size_t)nprods * num_vars + 1, nout = (size_t)num_vars + nprods;
  std::vector<Fr> hbuf(nin);
  memcpy(hbuf.data(), host_scalars, nprods * sizeof(Fr));
  memcpy(hbuf.data() + nprods, host_ys, (size_t)nprods * num_vars * sizeof(Fr));
  memcpy(hbuf.data() + nprods + (size_t)nprods * num_vars, host_sum, sizeof(Fr));
  Fr* d = nullptr;
  CUDA_TRY(cudaMallocAsync(&d, (nin + nout) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(d, hbuf.data(), nin * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  ScCoeffJob job;
  job.num_vars = num_vars;
  job.K = nprods;
  for (int k = 0; k < nprods; ++k) {
    job.tables[k] = (const Fr*)dev_tables[k];
    job.eq_points[k] = d + nprods + (size_t)k * num_vars;
  }
  job.scalars = d;
  job.claim = d + nprods + (size_t)nprods * num_vars;
  job.challenges_out = d + nin;
  job.evals_out = d + nin + num_vars;
  int rc = sumcheck_prove_coeffs(c, job);
  if (rc) return rc;
  std::vector<Fr> hout(nout);
  rc = fetch(c, hout.data(), d + nin, nout * sizeof(Fr));
  if (rc) return rc;
  memcpy(host_challenges_out, hout.data(), num_vars * sizeof(Fr));
  memcpy(host_evals_out, hout.data() + num_vars, nprods * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return B200_OK;
}

}  // extern "C"

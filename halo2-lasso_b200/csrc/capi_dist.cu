// extern "C" boundary, part 3: multi-GPU (one process per GPU; peer mailboxes over CUDA IPC / NVLink).
#include <string.h>

#include <vector>

#include "../../include/b200_lasso.h"
#include "internal.h"

using namespace b200;

extern "C" {

int b200_dist_mailbox_handle(b200_ctx* h, void* out_handle64) {
  Ctx* c = &h->c;
  if (!c->my_mailbox) {
    CUDA_TRY(cudaMalloc(&c->my_mailbox, sizeof(Mailbox)));
    CUDA_TRY(cudaMemset(c->my_mailbox, 0, sizeof(Mailbox)));
  }
  cudaIpcMemHandle_t hd;
  CUDA_TRY(cudaIpcGetMemHandle(&hd, c->my_mailbox));
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  memcpy(out_handle64, &hd, 64);
  return B200_OK;
}

int b200_dist_init(b200_ctx* h, int rank, int world, const void* handles) {
  Ctx* c = &h->c;
  if (world < 1 || world > PEER_MAX_WORLD || rank < 0 || rank >= world || !c->my_mailbox) return B200_ERR_ARG;
  c->peer.rank = rank;
  c->peer.world = world;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->peer.box[r] = c->my_mailbox;
      continue;
    }
    cudaIpcMemHandle_t hd;
    memcpy(&hd, (const char*)handles + 64 * r, 64);
    void* p = nullptr;
    CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c->peer.box[r] = (Mailbox*)p;
  }
  c->peer_seq = 0;
  return B200_OK;
}

int b200_dist_shard_commits(b200_ctx* h, int on) {
  if (on && h->c.peer.world < 2) return B200_ERR_ARG;
  h->c.shard_commits = on != 0;
  return B200_OK;
}

int b200_dist_shard_sumchecks(b200_ctx* h, int min_vars) {
  if (min_vars < 0 || (min_vars && h->c.peer.world < 2)) return B200_ERR_ARG;
  h->c.shard_sumcheck_min_vars = min_vars;
  return B200_OK;
}

int b200_sumcheck_prove_evals_sharded(b200_ctx* h, int num_vars_total, int nterms, int np,
                                      const void* const* dev_local_tables, const void* host_weights,
                                      const void* host_y, const void* host_sum, void* host_challenges_out,
                                      void* host_evals_out) {
  Ctx* c = &h->c;
  const int n = num_vars_total;
  if (n < 2 || n > 34 || nterms < 1 || nterms > SC_MAX_TERMS || (np != 1 && np != 2)) return B200_ERR_ARG;
  const int ntab = nterms * np;
  const size_t nin = (size_t)nterms + n + 1, nout = (size_t)n + ntab;
  std::vector<Fr> hbuf(nin);
  memcpy(hbuf.data(), host_weights, nterms * sizeof(Fr));
  memcpy(hbuf.data() + nterms, host_y, n * sizeof(Fr));
  memcpy(hbuf.data() + nterms + n, host_sum, sizeof(Fr));
  Fr* d = nullptr;
  CUDA_TRY(cudaMallocAsync(&d, (nin + nout) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(d, hbuf.data(), nin * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  ScEvalJob job;
  job.T = nterms;
  job.NP = np;
  job.num_vars = 0;  // set by the sharded driver
  for (int i = 0; i < ntab; ++i) job.tables[i] = (const Fr*)dev_local_tables[i];
  job.weights = d;
  job.eq_point = d + nterms;
  job.claim = d + nterms + n;
  job.challenges_out = d + nin;
  job.evals_out = d + nin + n;
  int rc = sumcheck_prove_evals_sharded(c, job, n);
  if (rc) return rc;
  std::vector<Fr> hout(nout);
  CUDA_TRY(cudaMemcpyAsync(hout.data(), d + nin, nout * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(host_challenges_out, hout.data(), n * sizeof(Fr));
  memcpy(host_evals_out, hout.data() + n, ntab * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return B200_OK;
}

int b200_variable_base_msm_sharded(b200_ctx* h, const void* host_scalars_fr, const void* host_bases_g1,
                                   uint64_t n_local, void* host_out_g1) {
  Ctx* c = &h->c;
  if (n_local == 0) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  Fr* ds = nullptr;
  G1Aff *db = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMallocAsync(&ds, n_local * sizeof(Fr), s));
  CUDA_TRY(cudaMallocAsync(&db, n_local * sizeof(G1Aff), s));
  CUDA_TRY(cudaMallocAsync(&dout, sizeof(G1Aff), s));
  CUDA_TRY(cudaMemcpyAsync(ds, host_scalars_fr, n_local * sizeof(Fr), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(db, host_bases_g1, n_local * sizeof(G1Aff), cudaMemcpyHostToDevice, s));
  MsmJob job{ds, db, n_local, MSM_FR_MONT, 254, nullptr};
  int rc = msm_sharded(c, job, dout);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(host_out_g1, dout, sizeof(G1Aff), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaFreeAsync(ds, s));
  CUDA_TRY(cudaFreeAsync(db, s));
  CUDA_TRY(cudaFreeAsync(dout, s));
  return B200_OK;
}

// ---- generic expression sum-check -------------------------------------------------------------------
int b200_sumcheck_prove_generic(b200_ctx* h, int num_vars, int degree, int ntables, const void* const* dev_tables,
                                int nconsts, const void* host_consts_fr, int nops, const int32_t* host_ops,
                                const void* host_sum, void* host_challenges_out, void* host_evals_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || ntables < 1 || ntables > 40 || nconsts < 0 || nops < 1) return B200_ERR_ARG;
  const size_t nin = (size_t)nconsts + 1, nout = (size_t)num_vars + ntables;
  std::vector<Fr> hbuf(nin);
  if (nconsts) memcpy(hbuf.data(), host_consts_fr, nconsts * sizeof(Fr));
  memcpy(hbuf.data() + nconsts, host_sum, sizeof(Fr));
  Fr* d = nullptr;
  int4* dops = nullptr;
  CUDA_TRY(cudaMallocAsync(&d, (nin + nout) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMallocAsync(&dops, (size_t)nops * sizeof(int4), c->stream));
  CUDA_TRY(cudaMemcpyAsync(d, hbuf.data(), nin * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(dops, host_ops, (size_t)nops * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
  GenericJob job;
  job.num_vars = num_vars;
  job.ntables = ntables;
  job.nconsts = nconsts;
  job.nops = nops;
  job.degree = degree;
  int max_dst = ntables + nconsts;
  for (int i = 0; i < nops; ++i) {
    const int32_t* o = host_ops + 4 * i;
    const int lim = ntables + nconsts;
    if (o[0] < 0 || o[0] > 3 || o[1] < lim || o[2] < 0 || o[3] < 0 || o[2] > o[1] + 64 || o[3] > o[1] + 64) return B200_ERR_ARG;
    if (o[1] > max_dst) max_dst = o[1];
  }
  job.ntemps = max_dst - (ntables + nconsts) + 1;
  for (int i = 0; i < nops; ++i)
    if (host_ops[4 * i + 2] > max_dst || host_ops[4 * i + 3] > max_dst) return B200_ERR_ARG;
  for (int i = 0; i < ntables; ++i) job.tables[i] = (const Fr*)dev_tables[i];
  job.consts = d;
  job.ops = dops;
  job.claim = d + nconsts;
  job.challenges_out = d + nin;
  job.evals_out = d + nin + num_vars;
  int rc = sumcheck_prove_generic(c, job);
  if (rc) return rc;
  std::vector<Fr> hout(nout);
  CUDA_TRY(cudaMemcpyAsync(hout.data(), d + nin, nout * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  memcpy(host_challenges_out, hout.data(), num_vars * sizeof(Fr));
  memcpy(host_evals_out, hout.data() + num_vars, ntables * sizeof(Fr));
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  CUDA_TRY(cudaFreeAsync(dops, c->stream));
  return B200_OK;
}
int b200_permutation_z(b200_ctx* h, int num_vars, int npolys, const void* const* dev_wires, const void* const* dev_sigmas,
                       const uint64_t* id_offsets, const void* host_beta_gamma, void* dev_z_out) {
  Ctx* c = &h->c;
  Fr* bg = nullptr;
  CUDA_TRY(cudaMallocAsync(&bg, 2 * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(bg, host_beta_gamma, 2 * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  int rc = permutation_z(c, num_vars, npolys, (const Fr* const*)dev_wires, (const Fr* const*)dev_sigmas, id_offsets, bg,
                         (Fr*)dev_z_out);
  CUDA_TRY(cudaFreeAsync(bg, c->stream));
  return rc;
}
int b200_poly_iota(b200_ctx* h, int num_vars, void* dev_out) { return poly_iota(&h->c, num_vars, (Fr*)dev_out); }
int b200_poly_onehot(b200_ctx* h, int num_vars, uint64_t index, void* dev_out) {
  return poly_onehot(&h->c, num_vars, index, (Fr*)dev_out);
}
int b200_poly_rotate(b200_ctx* h, const void* dev_in, int num_vars, int rotation, void* dev_out) {
  return poly_rotate(&h->c, (const Fr*)dev_in, num_vars, rotation, (Fr*)dev_out);
}

}  // extern "C"

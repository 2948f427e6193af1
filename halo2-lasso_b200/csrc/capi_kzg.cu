// extern "C" boundary, part 2: MSM and MultilinearKzg (include/b200_lasso.h).
#include <vector>

#include "../../include/b200_lasso.h"
#include "internal.h"

using namespace b200;

extern "C" {

int b200_variable_base_msm(b200_ctx* h, const void* host_scalars_fr, const void* host_bases_g1, uint64_t n,
                           void* host_out_g1) {
  Ctx* c = &h->c;
  if (n == 0) {  // empty sum = identity, as variable_base_msm returns
    memset(host_out_g1, 0, sizeof(G1Aff));
    return B200_OK;
  }
  cudaStream_t s = c->stream;
  Fr* ds = nullptr;
  G1Aff *db = nullptr, *dout = nullptr;
  CUDA_TRY(cudaMallocAsync(&ds, n * sizeof(Fr), s));
  CUDA_TRY(cudaMallocAsync(&db, n * sizeof(G1Aff), s));
  CUDA_TRY(cudaMallocAsync(&dout, sizeof(G1Aff), s));
  CUDA_TRY(cudaMemcpyAsync(ds, host_scalars_fr, n * sizeof(Fr), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(db, host_bases_g1, n * sizeof(G1Aff), cudaMemcpyHostToDevice, s));
  MsmJob job{ds, db, n, MSM_FR_MONT, 254, nullptr};
  int rc = msm_batch(c, &job, 1, dout);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(host_out_g1, dout, sizeof(G1Aff), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaFreeAsync(ds, s));
  CUDA_TRY(cudaFreeAsync(db, s));
  CUDA_TRY(cudaFreeAsync(dout, s));
  return B200_OK;
}

int b200_kzg_srs_upload(b200_ctx* h, int level, const void* host_g1) {
  Ctx* c = &h->c;
  if (level < 0 || level > 30 || level != (int)c->srs.size()) return B200_ERR_ARG;
  G1Aff* d = nullptr;
  const size_t bytes = ((size_t)1 << level) * sizeof(G1Aff);
  CUDA_TRY(cudaMalloc(&d, bytes));
  CUDA_TRY(cudaMemcpy(d, host_g1, bytes, cudaMemcpyHostToDevice));
  c->srs.push_back(d);
  return kzg_build_ext(c, level);
}

int b200_kzg_setup(b200_ctx* h, const void* host_ss_fr, int num_vars) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 28) return B200_ERR_ARG;
  Fr* dss = nullptr;
  CUDA_TRY(cudaMallocAsync(&dss, num_vars * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(dss, host_ss_fr, num_vars * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  int rc = kzg_setup(c, dss, num_vars);
  CUDA_TRY(cudaFreeAsync(dss, c->stream));
  return rc;
}

int b200_kzg_srs_download(b200_ctx* h, int level, void* host_g1_out) {
  Ctx* c = &h->c;
  if (level < 0 || level >= (int)c->srs.size()) return B200_ERR_ARG;
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaMemcpy(host_g1_out, c->srs[level], ((size_t)1 << level) * sizeof(G1Aff), cudaMemcpyDeviceToHost));
  return B200_OK;
}

int b200_kzg_batch_commit(b200_ctx* h, const void* const* dev_polys, const int* num_vars, int npolys,
                          int write_transcript, void* host_out_g1) {
  Ctx* c = &h->c;
  if (npolys < 1 || npolys > 64) return B200_ERR_ARG;
  std::vector<MsmJob> jobs(npolys);
  for (int i = 0; i < npolys; ++i) {
    if (num_vars[i] < 0 || num_vars[i] >= (int)c->srs.size()) return B200_ERR_ARG;  // "Too many variates"
    jobs[i] = MsmJob{dev_polys[i], c->srs[num_vars[i]], (uint64_t)1 << num_vars[i], MSM_FR_MONT, 254,
                     c->srs_ext[num_vars[i]]};
  }
  G1Aff* dout = nullptr;
  CUDA_TRY(cudaMallocAsync(&dout, npolys * sizeof(G1Aff), c->stream));
  int rc = kzg_commit_batch(c, jobs.data(), npolys, write_transcript != 0, dout);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(host_out_g1, dout, npolys * sizeof(G1Aff), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  CUDA_TRY(cudaFreeAsync(dout, c->stream));
  if (write_transcript) {
    Transcript t;
    CUDA_TRY(cudaMemcpy(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost));
    if (t.error) return B200_ERR_TRANSCRIPT;
  }
  return B200_OK;
}

int b200_kzg_open(b200_ctx* h, const void* dev_poly, int num_vars, const void* host_point) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30) return B200_ERR_ARG;
  Fr* dp = nullptr;
  CUDA_TRY(cudaMallocAsync(&dp, num_vars * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(dp, host_point, num_vars * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  int rc = kzg_open(c, (const Fr*)dev_poly, num_vars, dp);
  CUDA_TRY(cudaFreeAsync(dp, c->stream));
  return rc;
}

int b200_kzg_batch_open(b200_ctx* h, int num_vars, const void* const* dev_polys, int npolys,
                        const void* host_points, int npoints, const int* ev_poly, const int* ev_point,
                        const void* host_ev_values, int nevals) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || npolys < 1 || npoints < 1 || nevals < 2) return B200_ERR_ARG;
  for (int k = 0; k < nevals; ++k)
    if (ev_poly[k] < 0 || ev_poly[k] >= npolys || ev_point[k] < 0 || ev_point[k] >= npoints) return B200_ERR_ARG;
  Fr* d = nullptr;
  const size_t npt = (size_t)npoints * num_vars;
  CUDA_TRY(cudaMallocAsync(&d, (npt + nevals) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(d, host_points, npt * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(d + npt, host_ev_values, nevals * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  BatchOpenJob job;
  job.num_vars = num_vars;
  job.npolys = npolys;
  job.npoints = npoints;
  job.nevals = nevals;
  job.polys = (const Fr* const*)dev_polys;
  job.points = d;
  job.ev_poly = ev_poly;
  job.ev_point = ev_point;
  job.ev_values = d + npt;
  int rc = kzg_batch_open(c, job);
  CUDA_TRY(cudaFreeAsync(d, c->stream));
  return rc;
}

}  // extern "C"

// ---- Lasso ------------------------------------------------------------------------------------
extern "C" {

static int upload_operands(Ctx* c, int mu, const uint64_t* host_xs, const uint64_t* host_ys, uint64_t** dx,
                           uint64_t** dy) {
  const size_t bytes = ((size_t)1 << mu) * 8;
  *dy = nullptr;
  CUDA_TRY(cudaMallocAsync(dx, bytes, c->stream));
  CUDA_TRY(cudaMemcpyAsync(*dx, host_xs, bytes, cudaMemcpyHostToDevice, c->stream));
  if (host_ys) {
    CUDA_TRY(cudaMallocAsync(dy, bytes, c->stream));
    CUDA_TRY(cudaMemcpyAsync(*dy, host_ys, bytes, cudaMemcpyHostToDevice, c->stream));
  }
  return B200_OK;
}

int b200_lasso_prove_dev(b200_ctx* h, int kind, int chunks, int mu, const void* dev_xs, const void* dev_ys) {
  if (kind != 0 && !dev_ys) return B200_ERR_ARG;
  return lasso_prove(&h->c, kind, chunks, mu, (const uint64_t*)dev_xs, (const uint64_t*)dev_ys);
}

int b200_lasso_prove(b200_ctx* h, int kind, int chunks, int mu, const uint64_t* host_xs, const uint64_t* host_ys) {
  Ctx* c = &h->c;
  if (mu < 1 || mu > 26 || !host_xs || (kind != 0 && !host_ys)) return B200_ERR_ARG;
  uint64_t *dx, *dy;
  int rc = upload_operands(c, mu, host_xs, host_ys, &dx, &dy);
  if (rc) return rc;
  rc = lasso_prove(c, kind, chunks, mu, dx, dy);
  CUDA_TRY(cudaFreeAsync(dx, c->stream));
  if (dy) CUDA_TRY(cudaFreeAsync(dy, c->stream));
  if (rc) return rc;
  Transcript t;
  CUDA_TRY(cudaMemcpyAsync(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return t.error ? B200_ERR_TRANSCRIPT : B200_OK;
}

// ---- decomposable tables as data (the role of a DecomposableTable implementation) --------------------------------
struct b200_lasso_tab {
  LassoTableDesc d;
  uint32_t* d_values;
  uint32_t *d_perm = nullptr, *d_off = nullptr;
};

int b200_lasso_table_create(b200_ctx* h, const b200_lasso_table* t, b200_lasso_tab** out) {
  if (!h || !t || !out || !t->subtable) return B200_ERR_ARG;
  if (t->chunks < 2 || t->chunks > 8 || t->num_operands < 1 || t->num_operands > 2 || t->operand_bits < 1 ||
      t->num_operands * t->operand_bits > 16 || t->operand_bits * t->chunks > 64 || t->out_bits < 1 || t->out_bits > 32)
    return B200_ERR_ARG;
  const size_t S = (size_t)1 << 16;
  uint32_t mx = 0;
  for (size_t i = 0; i < S; ++i) mx = t->subtable[i] > mx ? t->subtable[i] : mx;
  int vb = 0;
  while (vb < 32 && (mx >> vb)) ++vb;
  if (t->out_bits * (t->chunks - 1) + vb > 64) return B200_ERR_ARG;  // the lookup output must fit 64 bits
  // digest: Keccak-256 (rate 136, pad 0x01 .. 0x80) over the values as little-endian u32 words
  uint64_t st[25] = {0};
  uint32_t pos = 0;
  for (size_t i = 0; i < S; ++i) {
    st[pos >> 3] ^= (uint64_t)t->subtable[i] << (8 * (pos & 7));  // pos is a multiple of 4: a word never straddles a lane
    pos += 4;
    if (pos == 136) {
      keccak_f1600(st);
      pos = 0;
    }
  }
  st[pos >> 3] ^= (uint64_t)0x01 << (8 * (pos & 7));
  st[16] ^= 0x8000000000000000ULL;
  keccak_f1600(st);
  Fr raw;
  for (int i = 0; i < 4; ++i) {
    raw.v[2 * i] = (uint32_t)st[i];
    raw.v[2 * i + 1] = (uint32_t)(st[i] >> 32);
  }
  b200_lasso_tab* tab = new b200_lasso_tab();
  tab->d_values = nullptr;
  if (cudaMalloc(&tab->d_values, S * 4) != cudaSuccess) {
    delete tab;
    return B200_ERR_NOMEM;
  }
  if (cudaMemcpy(tab->d_values, t->subtable, S * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
    cudaFree(tab->d_values);
    delete tab;
    return B200_ERR_CUDA;
  }
  tab->d = LassoTableDesc{t->chunks, t->num_operands, t->operand_bits, t->out_bits, vb, tab->d_values,
                          fe_from_canonical<FrP>(raw)};  // reduces the 256-bit integer mod r
  std::vector<uint32_t> perm, off;
  if (lasso_group_lists(t->subtable, &perm, &off) && !perm.empty()) {  // grouped E commitments (MsmJob::group_*)
    if (cudaMalloc(&tab->d_perm, perm.size() * 4) == cudaSuccess && cudaMalloc(&tab->d_off, off.size() * 4) == cudaSuccess &&
        cudaMemcpy(tab->d_perm, perm.data(), perm.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(tab->d_off, off.data(), off.size() * 4, cudaMemcpyHostToDevice) == cudaSuccess) {
      tab->d.d_group_perm = tab->d_perm;
      tab->d.d_group_off = tab->d_off;
      tab->d.ngroups = (int)off.size() - 1;
    }
    cudaGetLastError();
  }
  *out = tab;
  return B200_OK;
}
void b200_lasso_table_free(b200_lasso_tab* tab) {
  if (!tab) return;
  cudaFree(tab->d_values);
  cudaFree(tab->d_perm);
  cudaFree(tab->d_off);
  delete tab;
}
int b200_lasso_prove_table_dev(b200_ctx* h, const b200_lasso_tab* tab, int mu, const void* dev_xs, const void* dev_ys) {
  if (!h || !tab) return B200_ERR_ARG;
  return lasso_prove(&h->c, 3, tab->d.chunks, mu, (const uint64_t*)dev_xs, (const uint64_t*)dev_ys, &tab->d);
}
int b200_lasso_prove_table(b200_ctx* h, const b200_lasso_tab* tab, int mu, const uint64_t* host_xs, const uint64_t* host_ys) {
  if (!h || !tab) return B200_ERR_ARG;
  Ctx* c = &h->c;
  if (mu < 1 || mu > 26 || !host_xs || (tab->d.num_operands == 2 && !host_ys)) return B200_ERR_ARG;
  uint64_t *dx, *dy;
  int rc = upload_operands(c, mu, host_xs, tab->d.num_operands == 2 ? host_ys : nullptr, &dx, &dy);
  if (rc) return rc;
  rc = lasso_prove(c, 3, tab->d.chunks, mu, dx, dy, &tab->d);
  CUDA_TRY(cudaFreeAsync(dx, c->stream));
  if (dy) CUDA_TRY(cudaFreeAsync(dy, c->stream));
  if (rc) return rc;
  Transcript t;
  CUDA_TRY(cudaMemcpyAsync(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  return t.error ? B200_ERR_TRANSCRIPT : B200_OK;
}

int b200_fractional_sum_check_prove(b200_ctx* h, int num_batching, int num_vars, const void* const* dev_ps,
                                    const void* const* dev_qs, uint32_t claimed_mask, void* host_p_xs, void* host_q_xs,
                                    void* host_x, void* host_p_0s, void* host_q_0s) {
  Ctx* c = &h->c;
  const int B = num_batching, n = num_vars;
  if (B < 1 || B > 10 || n < 1 || n > 28 || !dev_ps || !dev_qs) return B200_ERR_ARG;
  for (int b = 0; b < B; ++b)
    if (!dev_ps[b] || !dev_qs[b]) return B200_ERR_ARG;
  DevScope mem(c->stream);
  Fr* d_out = nullptr;
  const size_t nout = (size_t)4 * B + n;
  CUDA_TRY(mem.alloc(&d_out, nout * sizeof(Fr)));
  int rc = fractional_sum_check_prove(c, B, n, (const Fr* const*)dev_ps, (const Fr* const*)dev_qs, claimed_mask, d_out);
  if (rc) return rc;
  std::vector<Fr> hout(nout);
  Transcript t;
  CUDA_TRY(cudaMemcpyAsync(hout.data(), d_out, nout * sizeof(Fr), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaMemcpyAsync(&t, c->d_tr, sizeof(t), cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (t.error) return B200_ERR_TRANSCRIPT;
  if (host_p_xs) memcpy(host_p_xs, hout.data(), B * sizeof(Fr));
  if (host_q_xs) memcpy(host_q_xs, hout.data() + B, B * sizeof(Fr));
  if (host_x) memcpy(host_x, hout.data() + 2 * B, n * sizeof(Fr));
  if (host_p_0s) memcpy(host_p_0s, hout.data() + 2 * B + n, B * sizeof(Fr));
  if (host_q_0s) memcpy(host_q_0s, hout.data() + 3 * B + n, B * sizeof(Fr));
  return B200_OK;
}

int b200_lasso_witness(b200_ctx* h, int kind, int chunks, int mu, const uint64_t* host_xs, const uint64_t* host_ys,
                       void* dev_mtabs, void* dev_stabs) {
  Ctx* c = &h->c;
  if (mu < 1 || mu > 26 || !host_xs || (kind != 0 && !host_ys)) return B200_ERR_ARG;
  uint64_t *dx, *dy;
  int rc = upload_operands(c, mu, host_xs, host_ys, &dx, &dy);
  if (rc) return rc;
  rc = lasso_witness(c, kind, chunks, mu, dx, dy, (Fr*)dev_mtabs, (Fr*)dev_stabs);
  CUDA_TRY(cudaFreeAsync(dx, c->stream));
  if (dy) CUDA_TRY(cudaFreeAsync(dy, c->stream));
  return rc;
}

}  // extern "C"

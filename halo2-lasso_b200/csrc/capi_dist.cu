// extern "C" boundary, part 3: multi-GPU (one process per GPU; peer mailboxes over CUDA IPC / NVLink).
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <vector>

#include "../../include/b200_lasso.h"
#include "internal.h"

using namespace b200;

namespace b200 {
void trace_point(Ctx* c, const char* what) {
  static const bool on = getenv("B200_TRACE") != nullptr;
  if (!on) return;
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  fprintf(stderr, "[b200 trace] rank %d %-28s %.3f ms\n", c->peer.rank, what, ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6);
}
}  // namespace b200

extern "C" {

static const size_t ARENA_HALF_DEFAULT = (size_t)160 << 20;  // per parity; B200_ARENA_MB overrides (both halves)

static int dist_alloc_local(Ctx* c) {
  if (!c->my_mailbox) CUDA_TRY(cudaMalloc(&c->my_mailbox, sizeof(Mailbox)));
  CUDA_TRY(cudaMemset(c->my_mailbox, 0, sizeof(Mailbox)));  // stale sequence numbers of an earlier group must not match
  if (!c->my_arena) {
    size_t half = ARENA_HALF_DEFAULT;
    if (const char* e = getenv("B200_ARENA_MB")) half = ((size_t)atol(e) << 20) / 2;
    if (half < ((size_t)1 << 20)) half = (size_t)1 << 20;
    CUDA_TRY(cudaMalloc(&c->my_arena, 2 * half));
    c->peer.arena_half = half;
  }
  if (!c->d_peer_err) {
    CUDA_TRY(cudaMalloc(&c->d_peer_err, sizeof(unsigned int)));
    CUDA_TRY(cudaMemset(c->d_peer_err, 0, sizeof(unsigned int)));
  }
  CUDA_TRY(cudaDeviceSynchronize());
  return B200_OK;
}
static void dist_set_common(Ctx* c, int rank, int world) {
  c->peer.rank = rank;
  c->peer.world = world;
  c->peer.err = c->d_peer_err;
  double secs = 20.0;
  if (const char* e = getenv("B200_PEER_TIMEOUT_S")) secs = atof(e);
  c->peer.timeout_ns = (unsigned long long)(secs * 1e9);
  // small messages: payload + release flag + acquire fence (1, default: 9 us per sharded round at 2 GPUs) or LL words
  // (0: measured 105 us — relaxed system-scope stores are not pushed out promptly without a release)
  c->peer.proto = 1;
  if (const char* e = getenv("B200_PEER_PROTO")) c->peer.proto = atoi(e) < 0 || atoi(e) > 3 ? 1 : atoi(e);
  c->peer_seq = 0;
  c->bulk_seq = 0;
  // A sharded round is worth its exchange while the (pair, term) items it takes off every other rank cost more than the
  // exchange itself: measured per sharded round 5 us at 2 GPUs, 15 us at 4, 115 us at 8 (tools/micro/shard_tune.py,
  // profiles/r02_shard_tune_*), at ~0.24 ns per item on a whole GPU: (world - 1) * items * 0.24 ns >= cost.
  // 115 us <-> 2^16 items per rank at 8 GPUs; kept for every world size (the 2- and 4-GPU optimum is lower, but flat).
  c->shard_min_items = 1 << 16;
  c->hb_ctas = 0;
  if (const char* e = getenv("B200_HEARTBEAT")) {  // "ctas,sleep_ns,write_peers"
    int a = 0, b = 1000, w = 1;
    if (sscanf(e, "%d,%d,%d", &a, &b, &w) >= 1 && a > 0) {
      c->hb_ctas = a > 148 ? 148 : a;
      c->hb_sleep_ns = b < 0 ? 0 : b;
      c->hb_write = w;
      if (!c->hb_stream) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        cudaStreamCreateWithPriority(&c->hb_stream, cudaStreamNonBlocking, lo);
        cudaEventCreateWithFlags(&c->hb_event, cudaEventDisableTiming);
        cudaMalloc(&c->hb_stop, sizeof(unsigned int));
        cudaMemset(c->hb_stop, 0, sizeof(unsigned int));
      }
    }
  }
}

// out_handle: 128 bytes = CUDA-IPC handle of the mailbox | CUDA-IPC handle of the bulk arena
int b200_dist_mailbox_handle(b200_ctx* h, void* out_handle128) {
  Ctx* c = &h->c;
  int rc = dist_alloc_local(c);
  if (rc) return rc;
  cudaIpcMemHandle_t hd;
  static_assert(sizeof(hd) == 64, "cudaIpcMemHandle_t is 64 bytes");
  CUDA_TRY(cudaIpcGetMemHandle(&hd, c->my_mailbox));
  memcpy(out_handle128, &hd, 64);
  CUDA_TRY(cudaIpcGetMemHandle(&hd, c->my_arena));
  memcpy((char*)out_handle128 + 64, &hd, 64);
  return B200_OK;
}

int b200_dist_init(b200_ctx* h, int rank, int world, const void* handles) {
  Ctx* c = &h->c;
  if (world < 1 || world > PEER_MAX_WORLD || (world & (world - 1)) || rank < 0 || rank >= world || !c->my_mailbox || !c->my_arena)
    return B200_ERR_ARG;
  dist_set_common(c, rank, world);
  c->peer_ipc = true;
  for (int r = 0; r < world; ++r) {
    if (r == rank) {
      c->peer.box[r] = c->my_mailbox;
      c->peer.arena[r] = c->my_arena;
      continue;
    }
    cudaIpcMemHandle_t hd;
    void* p = nullptr;
    memcpy(&hd, (const char*)handles + 128 * r, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c->peer.box[r] = (Mailbox*)p;
    memcpy(&hd, (const char*)handles + 128 * r + 64, 64);
    CUDA_TRY(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    c->peer.arena[r] = (unsigned char*)p;
  }
  return B200_OK;
}

// Several contexts of ONE process as the ranks of a group (contexts on the same GPU, or on GPUs with peer access
// enabled by the caller): mailboxes and arenas are shared as plain device pointers. Drive every context from its own
// host thread — collective calls wait for each other on the device. This is how the sharded provers are tested on a
// single-GPU box (tests/test_gpu_sharded.py); production uses one process per GPU (b200_dist_init).
int b200_dist_init_local(b200_ctx* const* ctxs, int world) {
  if (world < 1 || world > PEER_MAX_WORLD || (world & (world - 1))) return B200_ERR_ARG;
  for (int r = 0; r < world; ++r) {
    CUDA_TRY(cudaSetDevice(ctxs[r]->c.device));
    int rc = dist_alloc_local(&ctxs[r]->c);
    if (rc) return rc;
  }
  {
    // ... and the pool must not have to GROW while collectives are in flight (growing it can wait for the device, i.e.
    // for a kernel that is itself waiting for this rank): keep a reserve in the pool from the start. Freed blocks stay
    // cached (the release threshold is unlimited, b200_ctx_create).
    size_t reserve = (size_t)8 << 30;
    if (const char* e = getenv("B200_LOCAL_POOL_RESERVE_MB")) reserve = (size_t)atol(e) << 20;
    Ctx* c0 = &ctxs[0]->c;
    void* p = nullptr;
    if (reserve && cudaMallocAsync(&p, reserve, c0->stream) == cudaSuccess) {
      cudaFreeAsync(p, c0->stream);
      cudaStreamSynchronize(c0->stream);
    }
    cudaGetLastError();
  }
  for (int r = 0; r < world; ++r) {
    Ctx* c = &ctxs[r]->c;
    // ranks sharing a device share its stream-ordered memory pool: reusing a block another rank has freed but not yet
    // passed in ITS stream would make this rank's stream wait for that rank — which may be waiting for this one
    cudaMemPool_t pool;
    int zero = 0;
    CUDA_TRY(cudaDeviceGetDefaultMemPool(&pool, c->device));
    CUDA_TRY(cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &zero));
    dist_set_common(c, r, world);
    c->peer_ipc = false;
    g_use_pdl = false;
    c->peer.arena_half = ctxs[0]->c.peer.arena_half;
    for (int k = 0; k < world; ++k) {
      c->peer.box[k] = ctxs[k]->c.my_mailbox;
      c->peer.arena[k] = ctxs[k]->c.my_arena;
    }
  }
  return B200_OK;
}

// 0 when no peer wait has timed out since the last call (the flag is cleared), else B200_ERR_PEER
int b200_dist_check(b200_ctx* h) {
  Ctx* c = &h->c;
  if (!c->d_peer_err) return B200_OK;
  unsigned int v = 0;
  CUDA_TRY(cudaMemcpyAsync(&v, c->d_peer_err, 4, cudaMemcpyDeviceToHost, c->stream));
  CUDA_TRY(cudaStreamSynchronize(c->stream));
  if (v) {
    if (getenv("B200_PEER_DEBUG"))
      fprintf(stderr, "[b200] rank %d/%d: first timed-out wait: kind %u, source rank %u, sequence %u (small collectives issued %u, bulk %u)\n",
              c->peer.rank, c->peer.world, v >> 28, (v >> 24) & 15u, v & 0xffffffu, c->peer_seq, c->bulk_seq);
    CUDA_TRY(cudaMemsetAsync(c->d_peer_err, 0, 4, c->stream));
  }
  return v ? B200_ERR_PEER : B200_OK;
}

// Fully sharded Lasso prover: the m-sized witness tables, fingerprints, product-tree layers >= k0 and every sum-check
// over them live on the rank's 1/world slice (index window [k0 - log2 world, k0)); lookups with mu <= k0 fall back to
// the replicated prover. k0 = 0 switches it off. Collective like b200_dist_shard_commits (which it implies).
int b200_dist_shard_lasso(b200_ctx* h, int k0) {
  Ctx* c = &h->c;
  int g = 0;
  while ((1 << g) < c->peer.world) ++g;
  if (k0 < 0 || (k0 && (c->peer.world < 2 || k0 - g < 1 || k0 > 28))) return B200_ERR_ARG;
  c->shard_lasso_k0 = k0;
  return B200_OK;
}
// runtime knobs of the exchange (experiments, tools/micro): key 0 = small-message protocol (0 LL words, 1 release flag +
// fence, 2 release flag + acquire polls), key 1 = heartbeat CTAs (0 = off; needs a context created with B200_HEARTBEAT),
// key 2 = heartbeat sleep ns, key 3 = heartbeat mode bits
int b200_dist_tune(b200_ctx* h, int key, int value) {
  Ctx* c = &h->c;
  switch (key) {
    case 0:
      if (value < 0 || value > 3) return B200_ERR_ARG;
      c->peer.proto = value;
      return B200_OK;
    case 1:
      if (value < 0 || value > 148 || (value && !c->hb_stream)) return B200_ERR_ARG;
      c->hb_ctas = value;
      return B200_OK;
    case 2:
      c->hb_sleep_ns = value < 0 ? 0 : value;
      return B200_OK;
    case 3:
      c->hb_write = value;
      return B200_OK;
    case 4:
      if (value < 1 || value > 10000) return B200_ERR_ARG;
      c->hb_max_ms = value;  // the heartbeat kernel leaves by itself after this many ms
      return B200_OK;
  }
  return B200_ERR_ARG;
}
int b200_dist_shard_min_items(b200_ctx* h, int items) {
  if (items < 1) return B200_ERR_ARG;
  h->c.shard_min_items = items;
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
  return b200_sumcheck_prove_evals_windowed(h, num_vars_total, -1, -1, nterms, np, dev_local_tables, host_weights, host_y,
                                            host_sum, host_challenges_out, host_evals_out);
}

int b200_sumcheck_prove_evals_windowed(b200_ctx* h, int num_vars_total, int window_pos, int sharded_rounds, int nterms,
                                       int np, const void* const* dev_local_tables, const void* host_weights,
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
  int rc = sumcheck_prove_evals_sharded(c, job, n, window_pos, sharded_rounds);
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

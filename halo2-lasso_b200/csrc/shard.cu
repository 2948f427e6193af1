// Multi-GPU sharding of the prover, one process per GPU.
//
// Sum-check: rank r owns the entries of every table whose index bits [p, p + g) equal r (p = n - g: the contiguous
// top-variable slices of north_star). The reference binds bit 0 first (multilinear.rs:612-616), so the first
// R <= p rounds only need the D per-rank partial sums of each round message: they are exchanged INSIDE the round
// kernel through peer memory (peer.cuh) and every rank then runs the identical Fiat-Shamir step on its own device
// transcript. Once a round no longer pays for its exchange, ONE bulk all-gather (fused with the pending bind, stores
// straight into the peers' arenas over NVLink) rebuilds the bound tables everywhere and the remaining rounds run
// redundantly on every rank. The transcript is byte-identical to the single-GPU / reference proof.
//
// MSM: each rank commits its own points (a point range, or the points of its slice through MsmJob::map_*); the G
// affine partial results are all-gathered and added in rank order.
#include <algorithm>
#include <vector>

#include "internal.h"

namespace b200 {

// ---- geometry --------------------------------------------------------------------------------------------------
// A polynomial over n variables is sharded on the index bits [p, p + g), G = 2^g ranks: rank r holds, compactly, the
// entries whose window bits equal r; local index = (high bits above the window) ‖ (low p bits). p = n - g is the
// "top variables" layout of north_star. The window is closed under the reference's LSB-first binds of the first p
// rounds (multilinear.rs:612-616) and, for p + g <= k, under the top-bit halving of the product trees
// (fractional_sum_check.rs:41-76), which is why the Lasso prover shards on a window in the middle.

// y_loc = y[0..p) ‖ y[p+g..n);  factor = Π_j (rank_j ? y[p+j] : 1 - y[p+j])   (eq factor of the fixed window bits)
__global__ void shard_local_point_kernel(const Fr* y, int n, int p, int g, int rank, Fr* y_loc, Fr* factor) {
  const int i = threadIdx.x;
  if (i < n - g) fe_st(y_loc + i, fe_ld(y + (i < p ? i : i + g)));
  if (i == 0) {
    const Fr one = fe_one<FrP>();
    Fr acc = one;
    for (int j = 0; j < g; ++j) {
      const Fr yj = fe_ld(y + p + j);
      acc = acc * (((rank >> j) & 1) ? yj : one - yj);
    }
    fe_st(factor, acc);
  }
}

// Bulk all-gather fused with the pending bind: every rank stores its (bound) local tables straight into every
// rank's arena in natural index order (full index = ((hi G + rank) << q) | lo for local index (hi << q) | lo), the
// last CTA publishes / waits (peer.cuh). After the kernel every rank holds the full tables, table i at i * len * G.
struct BulkArgs {
  PeerCtx pc;
  unsigned int bseq;
  const Fr* src[2 * SC_MAX_TABLES + 2];
  int ntab, q, bind;
  uint32_t len;  // entries per table written by this rank (after the bind)
  const ScState* st;
  unsigned int* counter;
};
__global__ void __launch_bounds__(256) shard_push_kernel(BulkArgs a) {
  const int G = a.pc.world;
  Fr r = fe_zero<FrP>();
  if (a.bind) r = fe_ld(&a.st->r);
  const size_t total = (size_t)a.ntab * a.len;
  const uint32_t mask = (1u << a.q) - 1;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)(e / a.len), j = (uint32_t)(e % a.len);
    Fr v;
    if (a.bind) {
      const Fr x0 = fe_ldg(a.src[i] + 2 * (size_t)j), x1 = fe_ldg(a.src[i] + 2 * (size_t)j + 1);
      v = (x1 - x0) * r + x0;
    } else {
      v = fe_ldg(a.src[i] + j);
    }
    const size_t didx = (size_t)i * a.len * G + ((((size_t)(j >> a.q) * G + a.pc.rank) << a.q) | (j & mask));
    for (int k = 0; k < G; ++k) fe_st(peer_bulk_dst(a.pc, a.bseq, (a.pc.rank + k) % G, didx), v);
  }
  __threadfence_system();
  if (!last_cta_ticket(a.counter)) return;
  if (threadIdx.x < 32) peer_bulk_commit_and_wait(a.pc, a.bseq);
}

// gathers `ntab` local tables of `len` entries (after the optional bind of 2 len entries with d_sc->r) into the
// arena of every rank; full_out[i] = where table i (len * G entries) now lives on THIS rank
int shard_allgather(Ctx* c, const Fr* const* src, int ntab, uint32_t len, int q, bool bind, const Fr** full_out) {
  const int G = c->peer.world;
  if (ntab < 1 || ntab > 2 * SC_MAX_TABLES + 2 || !c->peer.arena[c->peer.rank]) return B200_ERR_ARG;
  if ((size_t)ntab * len * G * sizeof(Fr) > c->peer.arena_half) return B200_ERR_NOMEM;
  BulkArgs a;
  a.pc = c->peer;
  a.bseq = ++c->bulk_seq;
  for (int i = 0; i < ntab; ++i) a.src[i] = src[i];
  a.ntab = ntab;
  a.q = q;
  a.bind = bind ? 1 : 0;
  a.len = len;
  a.st = c->d_sc;
  a.counter = &c->d_sc->counter;
  const size_t total = (size_t)ntab * len;
  int blocks = (int)((total + 255) / 256);
  if (blocks > NUM_SMS * 4) blocks = NUM_SMS * 4;
  if (blocks < 1) blocks = 1;
  shard_push_kernel<<<blocks, 256, 0, c->stream>>>(a);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  const Fr* base = reinterpret_cast<const Fr*>(c->peer.arena[c->peer.rank] + (size_t)(a.bseq & 1) * c->peer.arena_half);
  for (int i = 0; i < ntab; ++i) full_out[i] = base + (size_t)i * len * G;
  return B200_OK;
}

// all-reduce (sum) of cnt <= 32 field elements: vals[i] <- Σ_ranks vals[i]
__global__ void shard_allreduce_kernel(PeerCtx pc, unsigned int seq, Fr* vals, int cnt, unsigned int* sink) {
  __shared__ volatile int s_busy;
  if (threadIdx.x == 0) s_busy = 1;
  __syncthreads();
  if (threadIdx.x >= 32) {  // warps 1-3 keep the SM busy during the exchange (peer.cuh)
    peer_spin_while(&s_busy, sink);
    return;
  }
  const int i = threadIdx.x;
  peer_publish(pc, seq, i < cnt ? fe_ld(vals + i) : fe_zero<FrP>(), cnt);
  if (i < cnt) {
    Fr sum = fe_zero<FrP>();
    for (int r = 0; r < pc.world; ++r) sum = sum + peer_read(pc, seq, r, i);
    fe_st(vals + i, sum);
  }
  __syncwarp();
  if (i == 0) s_busy = 0;
}
int shard_allreduce(Ctx* c, Fr* d_vals, int cnt) {
  if (c->peer.world < 2) return B200_OK;
  for (int at = 0; at < cnt; at += 32) {
    const int k = std::min(32, cnt - at);
    shard_allreduce_kernel<<<1, 128, 0, c->stream>>>(c->peer, ++c->peer_seq, d_vals + at, k, &c->d_sc->pad[0]);
    count_launch(c);
  }
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// ---- heartbeat -------------------------------------------------------------------------------------------------
// Experiment / mitigation for the 8-GPU round latency: between two exchanges a rank's GPU runs ONE warp
// and its NVLink links carry nothing for tens of microseconds. `ctas` single-warp CTAs issue a few FMAs every
// `sleep_ns`; lane r of CTA 0 also stores a word into rank r's mailbox pad, so every link stays trained. The kernel
// leaves when *stop == gen (set in stream order by hb_stop_kernel) or after 2 s, whichever comes first. Keep `ctas`
// small: a resident heartbeat CTA pins its SM's shared-memory carveout, and a kernel that needs another carveout cannot
// start there until the heartbeat leaves (measured: 148 CTAs stalled every proof for the full 2 s).
// mode bit 0: store a word into every rank's mailbox pad (NVLink traffic); bit 1: stream 128-byte reads through the
// rank's own bulk arena (HBM traffic that misses L2 sooner or later)
__global__ void __launch_bounds__(32) peer_heartbeat_kernel(PeerCtx pc, const unsigned int* stop, unsigned int gen,
                                                            unsigned int sleep_ns, int mode, unsigned int* sink,
                                                            unsigned long long max_ns) {
  const int lane = threadIdx.x;
  const unsigned long long t0 = peer_now_ns();
  float x = (float)lane;
  unsigned int beat = 0, junk = 0;
  const unsigned int* arena = reinterpret_cast<const unsigned int*>(pc.arena[pc.rank]);
  const size_t arena_words = (size_t)(2 * pc.arena_half / 4);
  size_t at = (size_t)blockIdx.x * 4096 + lane;
  for (;;) {
    unsigned int v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(stop) : "memory");
    if (v == gen || peer_now_ns() - t0 > max_ns) break;
#pragma unroll
    for (int i = 0; i < 32; ++i) x = fmaf(x, 1.0001f, 0.5f);
    if ((mode & 1) && blockIdx.x == 0 && lane < pc.world) {
      unsigned int* q = &pc.box[lane]->pad[pc.rank & 7];
      asm volatile("st.relaxed.sys.global.u32 [%0], %1;" ::"l"(q), "r"(++beat) : "memory");
    }
    if (mode & 2) {
      unsigned int w;
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(w) : "l"(arena + at) : "memory");
      junk ^= w;
      at += 32 * 1024;  // a new 128-byte line (and DRAM page) every beat
      if (at >= arena_words) at = lane;
    }
    if (sleep_ns) __nanosleep(sleep_ns);
  }
  if (x == 12345.678f || junk == 0x12345678u) *sink = 0;
}
__global__ void hb_stop_kernel(unsigned int* stop, unsigned int gen) {
  if (threadIdx.x == 0) asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(stop), "r"(gen) : "memory");
}
HeartbeatScope::HeartbeatScope(Ctx* ctx) : c(ctx) {
  if (c->hb_ctas <= 0 || c->peer.world < 2 || !c->hb_stream) return;
  if (c->hb_depth++ > 0) return;
  ++c->hb_gen;
  cudaEventRecord(c->hb_event, c->stream);
  cudaStreamWaitEvent(c->hb_stream, c->hb_event, 0);
  peer_heartbeat_kernel<<<c->hb_ctas, 32, 0, c->hb_stream>>>(c->peer, c->hb_stop, c->hb_gen, (unsigned)c->hb_sleep_ns,
                                                             c->hb_write, &c->d_sc->pad[1],
                                                             (unsigned long long)c->hb_max_ms * 1000000ull);
  count_launch(c);
}
HeartbeatScope::~HeartbeatScope() {
  if (c->hb_ctas <= 0 || c->peer.world < 2 || !c->hb_stream) return;
  if (--c->hb_depth > 0) return;
  hb_stop_kernel<<<1, 32, 0, c->stream>>>(c->hb_stop, c->hb_gen);
  count_launch(c);
}

static int ilog2_exact(int G) {
  int g = 0;
  while ((1 << g) < G) ++g;
  return (1 << g) == G ? g : -1;
}
// sharded rounds worth their exchange: while a rank still has >= SHARD_MIN_ITEMS (pair, term) items per round
static int auto_rounds(Ctx* c, int n_loc, int terms, int p) {
  int R = 0;
  while (R < p && R < n_loc && ((uint64_t)terms << (n_loc - 1 - R)) >= (uint64_t)c->shard_min_items) ++R;
  return R;
}

// local eq table of a sharded polynomial: eq(y_loc, .) * factor
static int shard_eq_build(Ctx* c, DevScope& mem, const Fr* y, int n, int p, int g, Fr** eq_out, Fr** yloc_out) {
  const int n_loc = n - g;
  Fr *y_loc = nullptr, *eq = nullptr;
  CUDA_TRY(mem.alloc(&y_loc, (size_t)(n_loc + 1) * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&eq, ((size_t)1 << n_loc) * sizeof(Fr)));
  shard_local_point_kernel<<<1, 64, 0, c->stream>>>(y, n, p, g, c->peer.rank, y_loc, y_loc + n_loc);
  count_launch(c);
  int rc = n_loc >= 1 ? eq_build(c, y_loc, n_loc, eq) : B200_ERR_ARG;
  if (rc) return rc;
  rc = fr_scale(c, eq, (size_t)1 << n_loc, y_loc + n_loc);
  if (rc) return rc;
  *eq_out = eq;
  if (yloc_out) *yloc_out = y_loc;
  return B200_OK;
}

// EVAL shape on local slices. job_local.tables = compact local slices (2^(n-g) entries), eq_point = the FULL point
// (n coordinates). p < 0: top-variable layout (p = n - g). rounds < 0: as many sharded rounds as pay (auto_rounds).
int sumcheck_prove_evals_sharded(Ctx* c, const ScEvalJob& job_local, int n, int p, int rounds) {
  const int G = c->peer.world, g = ilog2_exact(G);
  if (G < 2 || g < 0 || n - g < 1) return B200_ERR_ARG;
  const int n_loc = n - g, ntab = job_local.T * job_local.NP;
  if (p < 0) p = n_loc;
  if (p > n_loc || ntab > SC_MAX_TABLES) return B200_ERR_ARG;
  int R = rounds < 0 ? auto_rounds(c, n_loc, job_local.T, p) : rounds;
  if (R > p) R = p;
  cudaStream_t s = c->stream;
  HeartbeatScope hb(c);
  DevScope mem(s);
  const Fr* full[SC_MAX_TABLES + 1];
  int rc;
  if (R > 0) {
    // phase 1: R rounds on the local slices, round partials exchanged inside the round kernel. The local eq table is
    // eq(y_loc, .) * factor: handed over as point + scale so that the eq-factored round kernel applies.
    Fr* y_loc = nullptr;
    CUDA_TRY(mem.alloc(&y_loc, (size_t)(n_loc + 1) * sizeof(Fr)));
    shard_local_point_kernel<<<1, 64, 0, s>>>(job_local.eq_point, n, p, g, c->peer.rank, y_loc, y_loc + n_loc);
    count_launch(c);
    ScCarry carry;
    carry.scope = &mem;
    ScEvalJob j1 = job_local;
    j1.num_vars = n_loc;
    j1.eq_point = y_loc;
    j1.eq_scale = y_loc + n_loc;
    j1.eq_table = nullptr;
    j1.sharded = true;
    j1.stop_after = R;
    j1.carry = &carry;
    rc = sumcheck_prove_evals(c, j1);
    if (rc) return rc;
    rc = shard_allgather(c, carry.cur, ntab + 1, (uint32_t)(carry.len >> 1), p - R, true, full);
  } else {
    Fr* eq_loc = nullptr;
    rc = shard_eq_build(c, mem, job_local.eq_point, n, p, g, &eq_loc, nullptr);
    if (rc) return rc;
    const Fr* src[SC_MAX_TABLES + 1];
    for (int i = 0; i < ntab; ++i) src[i] = job_local.tables[i];
    src[ntab] = eq_loc;
    rc = shard_allgather(c, src, ntab + 1, (uint32_t)1 << n_loc, p, false, full);
  }
  if (rc) return rc;
  // phase 2: the remaining n - R rounds on the gathered tables, identical on every rank
  ScEvalJob j2 = job_local;
  j2.num_vars = n - R;
  for (int i = 0; i < ntab; ++i) j2.tables[i] = full[i];
  j2.eq_table = full[ntab];
  j2.eq_scale = nullptr;
  j2.sharded = false;
  if (R > 0) j2.claim = &c->d_sc->claim;  // the running claim after phase 1
  j2.challenges_out = job_local.challenges_out + R;
  return sumcheck_prove_evals(c, j2);
}

// COEFF shape (additive::batch_open's sum-check) on local slices; eq_points are the FULL points
int sumcheck_prove_coeffs_sharded(Ctx* c, const ScCoeffJob& job_local, int n, int p, int rounds) {
  const int G = c->peer.world, g = ilog2_exact(G);
  if (G < 2 || g < 0 || n - g < 1) return B200_ERR_ARG;
  const int n_loc = n - g, K = job_local.K;
  if (p < 0) p = n_loc;
  if (p > n_loc || K < 1 || K > SC_MAX_TERMS) return B200_ERR_ARG;
  int R = rounds < 0 ? auto_rounds(c, n_loc, K, p) : rounds;
  if (R > p) R = p;
  cudaStream_t s = c->stream;
  DevScope mem(s);
  ScCoeffJob j1 = job_local;
  for (int k = 0; k < K; ++k) {
    Fr* eq = nullptr;
    int rc = shard_eq_build(c, mem, job_local.eq_points[k], n, p, g, &eq, nullptr);
    if (rc) return rc;
    j1.eq_tables[k] = eq;
  }
  const Fr* full[2 * SC_MAX_TERMS];
  int rc;
  if (R > 0) {
    ScCarry carry;
    carry.scope = &mem;
    j1.num_vars = n_loc;
    j1.sharded = true;
    j1.stop_after = R;
    j1.carry = &carry;
    rc = sumcheck_prove_coeffs(c, j1);
    if (rc) return rc;
    rc = shard_allgather(c, carry.cur, 2 * K, (uint32_t)(carry.len >> 1), p - R, true, full);
  } else {
    const Fr* src[2 * SC_MAX_TERMS];
    for (int k = 0; k < K; ++k) {
      src[k] = job_local.tables[k];
      src[K + k] = j1.eq_tables[k];
    }
    rc = shard_allgather(c, src, 2 * K, (uint32_t)1 << n_loc, p, false, full);
  }
  if (rc) return rc;
  ScCoeffJob j2 = job_local;
  j2.num_vars = n - R;
  for (int k = 0; k < K; ++k) {
    j2.tables[k] = full[k];
    j2.eq_tables[k] = full[K + k];
  }
  j2.sharded = false;
  if (R > 0) j2.claim = &c->d_sc->claim;
  j2.challenges_out = job_local.challenges_out + R;
  return sumcheck_prove_coeffs(c, j2);
}

// evaluate() of sharded polynomials: local <P, eq> with the rank's eq slice, then an all-reduce of the values
int mle_eval_many_sharded(Ctx* c, const Fr* const* h_tables_loc, int ntables, int n, int p, const Fr* d_point, Fr* d_out) {
  const int G = c->peer.world, g = ilog2_exact(G);
  if (G < 2 || g < 0 || n - g < 1 || p < 0 || p > n - g) return B200_ERR_ARG;
  DevScope mem(c->stream);
  Fr* eq = nullptr;
  int rc = shard_eq_build(c, mem, d_point, n, p, g, &eq, nullptr);
  if (rc) return rc;
  rc = mle_dot_many(c, h_tables_loc, ntables, (size_t)1 << (n - g), eq, d_out);
  if (rc) return rc;
  return shard_allreduce(c, d_out, ntables);
}

// Sum-checks issued by the replicated whole provers on REPLICATED tables (b200_dist_shard_sumchecks): rank r passes the
// sub-arrays [r 2^n / G, (r + 1) 2^n / G) of each table to the sharded driver above (top-variable layout): the
// per-pair work of the big rounds is divided by G and every rank still ends with the same transcript, challenges and
// evaluations. Collective: the decision depends on the job shape only.
int sumcheck_prove_evals_dist(Ctx* c, const ScEvalJob& job) {
  const int G = c->peer.world, g = ilog2_exact(G);
  if (G < 2 || g < 0 || c->shard_sumcheck_min_vars <= 0 || job.num_vars < c->shard_sumcheck_min_vars ||
      job.num_vars - g < 1 || job.eq_table || job.eq_scale || job.want_eq_eval || job.sharded)
    return sumcheck_prove_evals(c, job);
  ScEvalJob loc = job;
  const size_t off = (size_t)c->peer.rank << (job.num_vars - g);
  for (int i = 0; i < job.T * job.NP; ++i) loc.tables[i] = job.tables[i] + off;
  return sumcheck_prove_evals_sharded(c, loc, job.num_vars, -1, -1);
}

__global__ void shard_point_sum_kernel(PeerCtx pc, unsigned int seq, const G1Aff* mine, G1Aff* out, unsigned int* sink) {
  __shared__ volatile int s_busy;
  if (threadIdx.x == 0) s_busy = 1;
  __syncthreads();
  if (threadIdx.x >= 32) {
    peer_spin_while(&s_busy, sink);
    return;
  }
  // the affine point travels as two field-sized values (x, y); Fq and Fr share the 8x32-bit layout
  const int lane = threadIdx.x;
  Fr v = fe_zero<FrP>();
  if (lane < 2) {
    const Fq c = fe_ld(lane == 0 ? &mine->x : &mine->y);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] = c.v[i];
  }
  peer_publish(pc, seq, v, 2);
  if (lane == 0) {
    G1Xyzz acc = g1_identity();
    for (int r = 0; r < pc.world; ++r) {
      const Fr x = peer_read(pc, seq, r, 0), y = peer_read(pc, seq, r, 1);
      G1Aff p;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p.x.v[i] = x.v[i];
        p.y.v[i] = y.v[i];
      }
      acc = g1_add_affine(acc, p, false);
    }
    const G1Aff a = g1_to_affine(acc);
    fe_st(&out->x, a.x);
    fe_st(&out->y, a.y);
  }
  __syncwarp();
  if (lane == 0) s_busy = 0;
}

// all-gather + add up to 16 partial commitments at once: lane 2i / 2i+1 carry x / y of point idx[i]; afterwards
// lane i adds the `world` partial points of its commitment in rank order (identical result on every rank).
struct PointIdx {
  int v[16];
};
__global__ void shard_points_sum_kernel(PeerCtx pc, unsigned int seq, G1Aff* pts, PointIdx pidx, int cnt,
                                        unsigned int* sink) {
  __shared__ volatile int s_busy;
  if (threadIdx.x == 0) s_busy = 1;
  __syncthreads();
  if (threadIdx.x >= 32) {
    peer_spin_while(&s_busy, sink);
    return;
  }
  const int lane = threadIdx.x;
  const int* idx = pidx.v;
  Fr v = fe_zero<FrP>();
  if (lane < 2 * cnt) {
    const G1Aff* p = pts + idx[lane >> 1];
    const Fq c = fe_ld((lane & 1) ? &p->y : &p->x);
#pragma unroll
    for (int i = 0; i < 8; ++i) v.v[i] = c.v[i];
  }
  peer_publish(pc, seq, v, 2 * cnt);
  if (lane < cnt) {
    G1Xyzz acc = g1_identity();
    for (int r = 0; r < pc.world; ++r) {
      const Fr x = peer_read(pc, seq, r, 2 * lane), y = peer_read(pc, seq, r, 2 * lane + 1);
      G1Aff p;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        p.x.v[i] = x.v[i];
        p.y.v[i] = y.v[i];
      }
      acc = g1_add_affine(acc, p, false);
    }
    const G1Aff a = g1_to_affine(acc);
    fe_st(&pts[idx[lane]].x, a.x);
    fe_st(&pts[idx[lane]].y, a.y);
  }
  __syncwarp();
  if (lane == 0) s_busy = 0;
}

// Commitment MSMs split by point range: rank r takes scalars / bases [r n/G, (r+1) n/G) of every job that is large
// enough, the G partial commitments are all-gathered through the peer mailboxes and added. Small jobs (the low
// quotient levels of an opening) are computed redundantly on every rank. Collective: all ranks call it with
// identical job lists (the provers run replicated between the commitments).
static const uint64_t MSM_SHARD_MIN = 1u << 14;
int msm_batch_dist(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out, const MsmDerive* derive, int nderive) {
  const int G = c->peer.world;
  bool any_mapped = false;
  for (int j = 0; j < J; ++j) any_mapped |= jobs[j].map_g > 0;
  if (G < 2 || !(c->shard_commits || any_mapped)) return msm_batch(c, jobs, J, d_out, derive, nderive);
  std::vector<MsmJob> loc(jobs, jobs + J);
  std::vector<int> sharded;
  std::vector<char> is_sharded(J, 0);
  for (int j = 0; j < J; ++j) {
    MsmJob& m = loc[j];
    if (m.map_g > 0) {  // already the rank's own points (a slice of a sharded polynomial)
      sharded.push_back(j);
      is_sharded[j] = 1;
      continue;
    }
    if (!c->shard_commits || m.n < MSM_SHARD_MIN || m.n % G) continue;
    const uint64_t len = m.n / G, off = (uint64_t)c->peer.rank * len;
    const size_t esz = m.kind == MSM_U32 ? 4 : (m.kind == MSM_U64 ? 8 : sizeof(Fr));
    m.scalars = (const char*)m.scalars + off * esz;
    m.bases += off;
    if (m.ext) {
      m.ext_stride = m.ext_stride ? m.ext_stride : m.n;
      m.ext += off;
    }
    m.n = len;
    sharded.push_back(j);
    is_sharded[j] = 1;
  }
  for (int j = 0; j < J; ++j)  // a grouped job sums bucket sums of its source: partial exactly when the source is
    if (loc[j].group_src >= 0 && is_sharded[loc[j].group_src]) {
      sharded.push_back(j);
      is_sharded[j] = 1;
    }
  // a derived result is a partial sum exactly when its sources are (all of them or none: same shape by construction)
  for (int i = 0; i < nderive; ++i) {
    int cnt = 0;
    for (int t = 0; t < derive[i].nsrc; ++t) cnt += is_sharded[derive[i].src[t]];
    if (cnt != 0 && cnt != derive[i].nsrc) return B200_ERR_ARG;
    if (cnt) sharded.push_back(J + i);
  }
  int rc = msm_batch(c, loc.data(), J, d_out, derive, nderive);
  if (rc) return rc;
  if (sharded.empty()) return B200_OK;
  for (size_t at = 0; at < sharded.size(); at += 16) {
    const int cnt = (int)std::min<size_t>(16, sharded.size() - at);
    PointIdx pidx;
    for (int i = 0; i < 16; ++i) pidx.v[i] = i < cnt ? sharded[at + i] : 0;
    shard_points_sum_kernel<<<1, 128, 0, c->stream>>>(c->peer, ++c->peer_seq, d_out, pidx, cnt, &c->d_sc->pad[0]);
    count_launch(c);
  }
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int msm_sharded(Ctx* c, const MsmJob& local, G1Aff* d_out) {
  if (c->peer.world < 2) return B200_ERR_ARG;
  G1Aff* part = nullptr;
  CUDA_TRY(cudaMallocAsync(&part, sizeof(G1Aff), c->stream));
  int rc = msm_batch(c, &local, 1, part);
  if (rc) return rc;
  shard_point_sum_kernel<<<1, 128, 0, c->stream>>>(c->peer, ++c->peer_seq, part, d_out, &c->d_sc->pad[0]);
  count_launch(c);
  CUDA_TRY(cudaFreeAsync(part, c->stream));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_shard() {
  B200_PRELOAD(shard_local_point_kernel);
  B200_PRELOAD(shard_push_kernel);
  B200_PRELOAD(shard_allreduce_kernel);
  B200_PRELOAD(shard_point_sum_kernel);
  B200_PRELOAD(shard_points_sum_kernel);
  // the heartbeat must not pin a small shared-memory carveout on the SMs it sits on (kernels that need more shared memory
  // could not start there until it leaves): ask for the largest one
  cudaFuncSetAttribute(peer_heartbeat_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  B200_PRELOAD(peer_heartbeat_kernel);
  B200_PRELOAD(hb_stop_kernel);
}

}  // namespace b200

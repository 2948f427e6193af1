// Multi-GPU sharding of the two kernels BASELINE cfg4 names, one process per GPU.
//
// Sum-check: rank g owns the contiguous slice evals[g*N/G .. (g+1)*N/G) of every table, i.e. the TOP
// log2 G variables are fixed to the bits of g. The reference binds bit 0 first (multilinear.rs:612-616), so
// the first n - log2 G rounds only need the D per-rank partial sums of each round message: they are
// exchanged INSIDE the round kernel through peer memory (peer.cuh) and every rank then runs the identical
// Fiat-Shamir step on its own device transcript. After those rounds each table is down to one value per
// rank; one more peer all-gather rebuilds the G-entry tables everywhere and the last log2 G rounds run
// redundantly on every rank. The transcript is byte-identical to the single-GPU / reference proof.
//
// MSM: each rank commits its own point range; the G affine partial results are all-gathered and added.
#include <algorithm>
#include <vector>

#include "internal.h"

namespace b200 {

// eq factor of the fixed top variables: Π_j (rank_j ? y_j : 1 - y_j)
__global__ void shard_eq_factor_kernel(const Fr* y_top, int g, int rank, Fr* out) {
  if (threadIdx.x || blockIdx.x) return;
  const Fr one = fe_one<FrP>();
  Fr acc = one;
  for (int j = 0; j < g; ++j) {
    const Fr yj = fe_ld(y_top + j);
    acc = acc * (((rank >> j) & 1) ? yj : one - yj);
  }
  fe_st(out, acc);
}

// all-gather `cnt` (<= 32) values per rank; gathered[i * world + r] = value i of rank r
__global__ void shard_gather_kernel(PeerCtx pc, unsigned int seq, const Fr* vals, int cnt, Fr* gathered,
                                    unsigned int* sink) {
  __shared__ volatile int s_busy;  // warps 1-3 keep the SM busy during the exchange (peer.cuh)
  if (threadIdx.x == 0) s_busy = 1;
  __syncthreads();
  if (threadIdx.x >= 32) {
    peer_spin_while(&s_busy, sink);
    return;
  }
  const int lane = threadIdx.x;
  const Fr mine = lane < cnt ? fe_ld(vals + lane) : fe_zero<FrP>();
  peer_publish(pc, seq, mine, cnt);
  if (lane < cnt)
    for (int r = 0; r < pc.world; ++r) fe_st(gathered + (size_t)lane * pc.world + r, peer_read(pc, seq, r, lane));
  __syncwarp();
  if (lane == 0) s_busy = 0;
}

int sumcheck_prove_evals_sharded(Ctx* c, const ScEvalJob& job_local, int n) {
  const int G = c->peer.world;
  int g = 0;
  while ((1 << g) < G) ++g;
  if (G < 2 || (1 << g) != G || n - g < 1) return B200_ERR_ARG;
  const int n_loc = n - g, ntab = job_local.T * job_local.NP;
  if (ntab + 1 > SC_MAX_TABLES) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  Fr* scratch = nullptr;  // factor | finals[ntab+1] | gathered[(ntab+1)*G] | evals2[ntab]
  const size_t nscr = 1 + (ntab + 1) + (size_t)(ntab + 1) * G + ntab + 1;
  CUDA_TRY(cudaMallocAsync(&scratch, nscr * sizeof(Fr), s));
  Fr* factor = scratch;
  Fr* finals = factor + 1;
  Fr* gathered = finals + ntab + 1;
  shard_eq_factor_kernel<<<1, 32, 0, s>>>(job_local.eq_point + n_loc, g, c->peer.rank, factor);
  count_launch(c);

  // phase 1: n_loc rounds on the local slices, partial sums exchanged inside the round kernel
  ScEvalJob j1 = job_local;
  j1.num_vars = n_loc;
  j1.eq_scale = factor;
  j1.sharded = true;
  j1.want_eq_eval = true;
  j1.evals_out = finals;
  int rc = sumcheck_prove_evals(c, j1);
  if (rc) return rc;

  // phase 2: rebuild the G-entry tables everywhere and finish redundantly
  for (int at = 0; at < ntab + 1; at += 32) {  // one mailbox message carries up to 32 values per rank
    const int cnt = std::min(32, ntab + 1 - at);
    shard_gather_kernel<<<1, 128, 0, s>>>(c->peer, ++c->peer_seq, finals + at, cnt, gathered + (size_t)at * G,
                                          &c->d_sc->pad[0]);
    count_launch(c);
  }
  ScEvalJob j2 = job_local;
  j2.num_vars = g;
  for (int i = 0; i < ntab; ++i) j2.tables[i] = gathered + (size_t)i * G;
  j2.eq_table = gathered + (size_t)ntab * G;
  j2.claim = &c->d_sc->claim;  // the running claim after phase 1
  j2.challenges_out = job_local.challenges_out + n_loc;
  j2.evals_out = job_local.evals_out;
  rc = sumcheck_prove_evals(c, j2);
  if (rc) return rc;
  CUDA_TRY(cudaFreeAsync(scratch, s));
  return B200_OK;
}

// Sum-checks issued INSIDE the replicated whole provers (Lasso: the Surge primary sum-check and the per-layer
// grand-product sum-checks). Every rank holds the full tables, so rank g simply passes the sub-arrays
// [g 2^n / G, (g + 1) 2^n / G) of each table to the sharded driver above: the per-pair work of the big rounds is
// divided by G, nothing moves between GPUs but the D round partials, and every rank still ends with the same
// transcript, challenges and evaluations. Collective: the decision depends on the job shape only.
int sumcheck_prove_evals_dist(Ctx* c, const ScEvalJob& job) {
  const int G = c->peer.world;
  int g = 0;
  while ((1 << g) < G) ++g;
  if (G < 2 || (1 << g) != G || c->shard_sumcheck_min_vars <= 0 || job.num_vars < c->shard_sumcheck_min_vars ||
      job.num_vars - g < 1 || job.eq_table || job.eq_scale || job.want_eq_eval || job.sharded)
    return sumcheck_prove_evals(c, job);
  ScEvalJob loc = job;
  const size_t off = (size_t)c->peer.rank << (job.num_vars - g);
  for (int i = 0; i < job.T * job.NP; ++i) loc.tables[i] = job.tables[i] + off;
  return sumcheck_prove_evals_sharded(c, loc, job.num_vars);
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
int msm_batch_dist(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out) {
  const int G = c->peer.world;
  if (G < 2 || !c->shard_commits) return msm_batch(c, jobs, J, d_out);
  std::vector<MsmJob> loc(jobs, jobs + J);
  std::vector<int> sharded;
  for (int j = 0; j < J; ++j) {
    MsmJob& m = loc[j];
    if (m.n < MSM_SHARD_MIN || m.n % G) continue;
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
  }
  int rc = msm_batch(c, loc.data(), J, d_out);
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

}  // namespace b200

// MultilinearKzg on the GPU (pb/pcs/multilinear/kzg.rs:252-315) and the additive batch opening
// (pb/pcs/multilinear.rs:72-107 quotients, :134-235 batch_open). Every step stays on the device:
// challenges are squeezed into device memory, merged polynomials / g' are built by the linear-
// combination kernel, the degree-2 CoefficientsProver sum-check runs with the fused round kernel,
// and all n quotient commitments of an opening go through ONE batched MSM.
#include "internal.h"

namespace b200 {

int kzg_commit_batch(Ctx* c, const MsmJob* jobs, int J, bool write_transcript, G1Aff* d_out) {
  int rc = msm_batch_dist(c, jobs, J, d_out);
  if (rc) return rc;
  if (write_transcript) return transcript_write_points(c, d_out, J);
  return B200_OK;
}

// kzg.rs:276-302 (sanity-check off): quotients top variable first, commit q_i with eqs[i], write them.
// shard_p >= 0: d_poly is the rank's slice (window [p, p + g)): the quotient steps of the variables above the window
// are local, their commitments are MSMs over the rank's points (MsmJob::map_*) summed over the ranks; then the 2^(p+g)
// remainder is all-gathered once and the low levels run replicated.
int kzg_open(Ctx* c, const Fr* d_poly, int n, const Fr* d_point, int shard_p) {
  if (n < 1 || n > 30 || (int)c->srs.size() <= n - 1) return B200_ERR_ARG;
  NvtxRange nvtx("quotients");  // pcs/multilinear.rs:85
  cudaStream_t s = c->stream;
  const bool sh = shard_p >= 0 && c->peer.world > 1;
  int g = 0;
  while (sh && (1 << g) < c->peer.world) ++g;
  const int K0 = sh ? shard_p + g : 0;
  if (sh && K0 > n) return B200_ERR_ARG;
  const size_t N_loc = (size_t)1 << (n - g);
  DevScope mem(s);
  Fr *rem = nullptr, *q = nullptr, *rem_full = nullptr, *q_full = nullptr;
  G1Aff* comms = nullptr;
  CUDA_TRY(mem.alloc(&rem, N_loc * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&q, N_loc * sizeof(Fr)));  // level nv lives at [2^(nv-g), 2^(nv-g+1))
  CUDA_TRY(mem.alloc(&comms, n * sizeof(G1Aff)));
  CUDA_TRY(cudaMemcpyAsync(rem, d_poly, N_loc * sizeof(Fr), cudaMemcpyDeviceToDevice, s));
  MsmJob jobs[32];
  for (int nv = n - 1; nv >= K0; --nv) {
    const size_t half = (size_t)1 << (nv - g);
    int rc = quotient_step(c, rem, half, d_point + nv, q + half);
    if (rc) return rc;
    jobs[nv] = MsmJob{q + half, c->srs[nv], half, MSM_FR_MONT, 254, c->srs_ext[nv], (uint64_t)1 << nv};
    if (sh) {
      jobs[nv].map_p = shard_p;
      jobs[nv].map_g = g;
      jobs[nv].map_rank = c->peer.rank;
    }
  }
  if (sh && K0 > 0) {
    CUDA_TRY(mem.alloc(&rem_full, sizeof(Fr) << K0));
    CUDA_TRY(mem.alloc(&q_full, sizeof(Fr) << K0));
    const Fr* src[1] = {rem};
    const Fr* full[1];
    int rc = shard_allgather(c, src, 1, (uint32_t)1 << shard_p, shard_p, false, full);
    if (rc) return rc;
    CUDA_TRY(cudaMemcpyAsync(rem_full, full[0], sizeof(Fr) << K0, cudaMemcpyDeviceToDevice, s));
    for (int nv = K0 - 1; nv >= 0; --nv) {
      const size_t half = (size_t)1 << nv;
      rc = quotient_step(c, rem_full, half, d_point + nv, q_full + half);
      if (rc) return rc;
      jobs[nv] = MsmJob{q_full + half, c->srs[nv], half, MSM_FR_MONT, 254, c->srs_ext[nv]};
    }
  }
  return kzg_commit_batch(c, jobs, n, true, comms);
}

__global__ void gather_fr_kernel(const Fr* src, const int* idx, int n, Fr* dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(dst + i, fe_ld(src + idx[i]));
}
// out = Σ_k a[k] * b[k]  (tiny, single thread)
__global__ void dot_small_kernel(const Fr* a, const Fr* b, int n, Fr* out) {
  if (threadIdx.x || blockIdx.x) return;
  Fr acc = fe_zero<FrP>();
  for (int k = 0; k < n; ++k) acc = acc + fe_ld(a + k) * fe_ld(b + k);
  fe_st(out, acc);
}
__global__ void fill_one_kernel(Fr* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(out + i, fe_one<FrP>());
}

int kzg_batch_open(Ctx* c, const BatchOpenJob& job) {
  const int n = job.num_vars, P = job.npoints, E = job.nevals;
  if (n < 1 || n > 30 || P < 1 || P > SC_MAX_TERMS || E < 2 || E > SC_MAX_TABLES) return B200_ERR_ARG;
  NvtxRange nvtx("pcs_batch_open-%d", E);  // hyperplonk.rs:286
  cudaStream_t s = c->stream;
  const bool sh = job.shard_p >= 0 && c->peer.world > 1;
  int g = 0;
  while (sh && (1 << g) < c->peer.world) ++g;
  const size_t N = (size_t)1 << (n - g);  // entries per (local) table
  int ell = 0;
  while ((1 << ell) < E) ++ell;
  // scalar arena: t[ell] | eq_xt[2^ell] | gathered[E] | tilde | ones[P] | challenges[n] | evals[P] | eqev[P]
  const size_t arena_n = (size_t)ell + ((size_t)1 << ell) + E + 1 + P + n + P + P;
  DevScope mem(s);
  Fr* arena = nullptr;
  CUDA_TRY(mem.alloc(&arena, arena_n * sizeof(Fr)));
  Fr* t = arena;
  Fr* eq_xt = t + ell;
  Fr* gathered = eq_xt + ((size_t)1 << ell);
  Fr* tilde = gathered + E;
  Fr* ones = tilde + 1;
  Fr* challenges = ones + P;
  Fr* sc_evals = challenges + n;
  Fr* eqev = sc_evals + P;
  int rc = transcript_op(c, TR_SQUEEZE, nullptr, t, ell);
  if (rc) return rc;
  rc = eq_build(c, t, ell, eq_xt);
  if (rc) return rc;

  // merged_i = Σ_{k : point(k) = i} eq_xt[k] * poly(k)      (pb/pcs/multilinear.rs:155-170)
  Fr* merged = nullptr;
  CUDA_TRY(mem.alloc(&merged, (size_t)P * N * sizeof(Fr)));
  int h_idx[SC_MAX_TABLES];
  int* d_idx = nullptr;
  CUDA_TRY(mem.alloc(&d_idx, E * sizeof(int)));
  int pos = 0, start[SC_MAX_TERMS + 1];
  const Fr* tabs[SC_MAX_TABLES];
  std::vector<std::vector<const Fr*>> per_point(P);
  for (int i = 0; i < P; ++i) {
    start[i] = pos;
    for (int k = 0; k < E; ++k)
      if (job.ev_point[k] == i) {
        h_idx[pos++] = k;
        per_point[i].push_back(job.polys[job.ev_poly[k]]);
      }
    if (pos == start[i]) return B200_ERR_ARG;  // a point nobody is evaluated at
  }
  start[P] = pos;
  CUDA_TRY(cudaMemcpyAsync(d_idx, h_idx, E * sizeof(int), cudaMemcpyHostToDevice, s));
  gather_fr_kernel<<<1, 64, 0, s>>>(eq_xt, d_idx, E, gathered);
  count_launch(c);
  {
    NvtxRange nvtx_m("merged_polys");  // pcs/multilinear.rs:153
    for (int i = 0; i < P; ++i) {
      const int k = start[i + 1] - start[i];
      for (int j = 0; j < k; ++j) tabs[j] = per_point[i][j];
      rc = fr_lincomb(c, tabs, k, gathered + start[i], N, merged + (size_t)i * N);
      if (rc) return rc;
    }
  }
  // tilde_gs_sum = <evals.value, eq_xt[..E]>   (:195-196)
  dot_small_kernel<<<1, 32, 0, s>>>(job.ev_values, eq_xt, E, tilde);
  fill_one_kernel<<<1, 64, 0, s>>>(ones, P);
  count_launch(c, 2);

  ScCoeffJob sj;
  sj.num_vars = n;
  sj.K = P;
  for (int i = 0; i < P; ++i) {
    sj.tables[i] = merged + (size_t)i * N;
    sj.eq_points[i] = job.points + (size_t)i * n;
  }
  sj.scalars = ones;
  sj.claim = tilde;
  sj.challenges_out = challenges;
  sj.evals_out = sc_evals;
  rc = sh ? sumcheck_prove_coeffs_sharded(c, sj, n, job.shard_p, -1) : sumcheck_prove_coeffs(c, sj);
  if (rc) return rc;

  // g' = Σ_i eq(challenges, point_i) * merged_i      (:203-212)
  for (int i = 0; i < P; ++i) {
    rc = eq_xy_eval_dev(c, challenges, job.points + (size_t)i * n, n, eqev + i);
    if (rc) return rc;
    tabs[i] = merged + (size_t)i * N;
  }
  Fr* gprime = nullptr;
  CUDA_TRY(mem.alloc(&gprime, N * sizeof(Fr)));
  {
    NvtxRange nvtx_g("g_prime");  // pcs/multilinear.rs:203
    rc = fr_lincomb(c, tabs, P, eqev, N, gprime);
  }
  if (rc) return rc;
  return kzg_open(c, gprime, n, challenges, sh ? job.shard_p : -1);
}

// ---------------------------------------------------------------------------------------------
// MultilinearKzg::setup (kzg.rs:166-213): eqs[k][b] = g1 * Π_{j<k} (b_j ? s_j : 1 - s_j), k = 0..n.
// The reference builds the scalar tables by doubling and runs `fixed_base_msm` with a window table;
// here eq_build produces each level's scalars and one thread per point adds 32 byte-window table
// entries (mixed adds) and normalises. One-off cost, not on the prove path.
// ---------------------------------------------------------------------------------------------
__global__ void srs_window_bases_kernel(G1Xyzz* bases) {  // bases[w] = 2^(8w) * G, w < 32
  const int w = threadIdx.x;
  if (w >= 32) return;
  G1Aff g;
  g.x = fe_from_u64<FqP>(1);
  g.y = fe_from_u64<FqP>(2);
  G1Xyzz acc = g1_from_affine(g);
  for (int k = 0; k < 8 * w; ++k) acc = g1_dbl(acc);
  bases[w] = acc;
}
__global__ void srs_window_table_kernel(const G1Xyzz* bases, G1Aff* table) {  // table[w*255 + d-1] = d * bases[w]
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 32 * 255) return;
  const int w = i / 255, d = i % 255 + 1;
  table[i] = g1_to_affine(g1_mul_small(bases[w], (uint32_t)d));
}
__global__ void __launch_bounds__(128) srs_fixed_base_kernel(const Fr* __restrict__ scalars, size_t n,
                                                             const G1Aff* __restrict__ table, G1Aff* __restrict__ out) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Fr s = fe_to_canonical<FrP>(fe_ldg(scalars + i));
    G1Xyzz acc = g1_identity();
    for (int w = 0; w < 32; ++w) {
      const uint32_t byte = (s.v[w >> 2] >> (8 * (w & 3))) & 0xff;
      if (byte) {
        G1Aff p;
        p.x = fe_ldg(&table[w * 255 + byte - 1].x);
        p.y = fe_ldg(&table[w * 255 + byte - 1].y);
        acc = g1_add_affine(acc, p, false);
      }
    }
    const G1Aff a = g1_to_affine(acc);
    fe_st(&out[i].x, a.x);
    fe_st(&out[i].y, a.y);
  }
}
// ext[w*n + i] = 2^(16 w) * base[i]: 16 doublings per window, then ONE shared inversion per base point
// (Montgomery's trick over its 15 multiples).
__global__ void __launch_bounds__(128) srs_ext_kernel(const G1Aff* __restrict__ base, size_t n, G1Aff* __restrict__ ext) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    G1Aff p;
    p.x = fe_ldg(&base[i].x);
    p.y = fe_ldg(&base[i].y);
    fe_st(&ext[i].x, p.x);
    fe_st(&ext[i].y, p.y);
    if (g1_aff_is_identity(p)) {
      for (int w = 1; w < EXT_WINDOWS; ++w) {
        fe_st(&ext[(size_t)w * n + i].x, p.x);
        fe_st(&ext[(size_t)w * n + i].y, p.y);
      }
      continue;
    }
    G1Xyzz pts[EXT_WINDOWS - 1];
    Fq pre[EXT_WINDOWS - 1];  // prefix products of d_w = zz_w * zzz_w
    G1Xyzz acc = g1_from_affine(p);
    Fq run = fe_one<FqP>();
    for (int w = 0; w < EXT_WINDOWS - 1; ++w) {
      for (int k = 0; k < EXT_C; ++k) acc = g1_dbl(acc);
      pts[w] = acc;
      pre[w] = run;
      run = run * (acc.zz * acc.zzz);
    }
    Fq inv = fe_inv<FqP>(run);  // 1 / Π d_w   (points of prime order never double to the identity)
    for (int w = EXT_WINDOWS - 2; w >= 0; --w) {
      const Fq d = pts[w].zz * pts[w].zzz;
      const Fq dinv = inv * pre[w];  // 1 / d_w
      inv = inv * d;
      fe_st(&ext[(size_t)(w + 1) * n + i].x, pts[w].x * (dinv * pts[w].zzz));
      fe_st(&ext[(size_t)(w + 1) * n + i].y, pts[w].y * (dinv * pts[w].zz));
    }
  }
}

int kzg_build_ext(Ctx* c, int level) {
  if (level < 0 || level >= (int)c->srs.size()) return B200_ERR_ARG;
  if ((int)c->srs_ext.size() <= level) c->srs_ext.resize(level + 1, nullptr);
  if (c->srs_ext[level]) return B200_OK;
  const size_t n = (size_t)1 << level;
  G1Aff* ext = nullptr;
  CUDA_TRY(cudaMalloc(&ext, (size_t)EXT_WINDOWS * n * sizeof(G1Aff)));
  int blocks = (int)((n + 127) / 128);
  if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
  srs_ext_kernel<<<blocks, 128, 0, c->stream>>>(c->srs[level], n, ext);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  c->srs_ext[level] = ext;
  return B200_OK;
}

__global__ void fill_one_fr_kernel(Fr* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) fe_st(out, fe_one<FrP>());
}

int kzg_setup(Ctx* c, const Fr* d_ss, int n) {
  if (n < 1 || n > 28 || !c->srs.empty()) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  G1Xyzz* bases = nullptr;
  G1Aff* table = nullptr;
  Fr* eq = nullptr;
  CUDA_TRY(cudaMallocAsync(&bases, 32 * sizeof(G1Xyzz), s));
  CUDA_TRY(cudaMallocAsync(&table, 32 * 255 * sizeof(G1Aff), s));
  CUDA_TRY(cudaMallocAsync(&eq, ((size_t)1 << n) * sizeof(Fr), s));
  srs_window_bases_kernel<<<1, 32, 0, s>>>(bases);
  srs_window_table_kernel<<<(32 * 255 + 127) / 128, 128, 0, s>>>(bases, table);
  count_launch(c, 2);
  for (int k = 0; k <= n; ++k) {
    const size_t N = (size_t)1 << k;
    G1Aff* lvl = nullptr;
    CUDA_TRY(cudaMalloc(&lvl, N * sizeof(G1Aff)));
    if (k == 0) {
      fill_one_fr_kernel<<<1, 32, 0, s>>>(eq);
    } else {
      int rc = eq_build(c, d_ss, k, eq);
      if (rc) return rc;
    }
    int blocks = (int)((N + 127) / 128);
    if (blocks > NUM_SMS * 16) blocks = NUM_SMS * 16;
    srs_fixed_base_kernel<<<blocks, 128, 0, s>>>(eq, N, table, lvl);
    count_launch(c, 2);
    c->srs.push_back(lvl);
    int rc = kzg_build_ext(c, k);
    if (rc) return rc;
  }
  CUDA_TRY(cudaFreeAsync(bases, s));
  CUDA_TRY(cudaFreeAsync(table, s));
  CUDA_TRY(cudaFreeAsync(eq, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_kzg() {
  B200_PRELOAD(gather_fr_kernel);
  B200_PRELOAD(dot_small_kernel);
  B200_PRELOAD(fill_one_kernel);
  B200_PRELOAD(srs_window_bases_kernel);
  B200_PRELOAD(srs_window_table_kernel);
  B200_PRELOAD(srs_fixed_base_kernel);
  B200_PRELOAD(srs_ext_kernel);
  B200_PRELOAD(fill_one_fr_kernel);
}

}  // namespace b200

// Lasso / Surge lookup argument on the GPU. The mounted reference has no Lasso code,
// so the protocol is the one written down in DESIGN.md §"Lasso protocol" and restated on the CPU in
// oracle/lasso.hpp; the layered grand-product argument follows the in-tree template
// pb/piop/gkr/fractional_sum_check.rs:41-190,272-296 (top-bit halves, batched with powers of gamma).
//
// Everything is enqueued on one stream with no host round trip:
//   witness   : chunk extraction, subtable gather, DETERMINISTIC read/final counters
//               (per-chunk shared-memory histograms -> column scan -> in-order warp ranking; no
//               order-dependent atomics, so read_ts is bit-identical run to run)
//   commit    : one batched MSM over the raw integer witnesses (16-/8-/21-/64-bit scalars populate only
//               the windows they need)
//   sum-checks: primary Surge (degree 2) and one degree-3 batched sum-check per tree layer, all with
//               the fused bind+round kernel and the on-device transcript
//   openings  : leaf evaluations + two additive batch openings (mu-variate and 16-variate)
#include "internal.h"

namespace b200 {

static const int SUB_VARS = 16;
static const uint32_t SUB_SIZE = 1u << SUB_VARS;

// The table as the kernels see it: the three built-in kinds have closed forms; kind 3 (LassoTableDesc, internal.h) is a
// table given as data — 2^16 subtable values in device memory, 1 or 2 operands of op_bits bits per chunk.
struct LassoTab {
  int kind, c, nops, op_bits, out_bits;
  const uint32_t* values;
};
__device__ __forceinline__ uint32_t lasso_subtable(const LassoTab& tb, uint32_t x) {
  if (tb.kind == 3) return __ldg(tb.values + x);
  if (tb.kind == 0) return x;
  const uint32_t p = x >> 8, q = x & 0xff;
  return tb.kind == 1 ? (p & q) : (p ^ q);
}
__device__ __forceinline__ uint32_t lasso_dim(const LassoTab& tb, uint64_t x, uint64_t y, int t) {
  if (tb.kind == 3) {
    const uint64_t mask = ((uint64_t)1 << tb.op_bits) - 1;
    const uint32_t xt = (uint32_t)((x >> (tb.op_bits * t)) & mask), yt = (uint32_t)((y >> (tb.op_bits * t)) & mask);
    return tb.nops == 1 ? xt : (xt << tb.op_bits) | yt;
  }
  if (tb.kind == 0) return (uint32_t)((x >> (16 * t)) & 0xffff);
  return (uint32_t)((((x >> (8 * t)) & 0xff) << 8) | ((y >> (8 * t)) & 0xff));
}

// dims[t][j], e[t][j] (u32) and the lookup output a[j] = Σ_t 2^(out_bits t) e[t][j] (u64)
// *bad is raised when an operand does not fit the table (x >= 2^(16 c) for the range table, x or y >= 2^(op_bits c)
// otherwise): such a lookup is NOT in the decomposed table and must not be proven modulo the chunk width.
__global__ void lasso_chunks_kernel(LassoTab tb, uint32_t m, const uint64_t* __restrict__ xs,
                                    const uint64_t* __restrict__ ys, uint32_t* __restrict__ dims,
                                    uint32_t* __restrict__ es, uint64_t* __restrict__ a, unsigned int* bad) {
  const uint32_t stride = gridDim.x * blockDim.x;
  const int c = tb.c, out_bits = tb.out_bits;
  const int op_bits = tb.op_bits * c;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    const uint64_t x = xs[j], y = ys ? ys[j] : 0;
    if (op_bits < 64 && (((x | y) >> op_bits) != 0)) *bad = 1;
    uint64_t out = 0;
    for (int t = 0; t < c; ++t) {
      const uint32_t d = lasso_dim(tb, x, y, t);
      const uint32_t e = lasso_subtable(tb, d);
      dims[(size_t)t * m + j] = d;
      es[(size_t)t * m + j] = e;
      out += (uint64_t)e << (out_bits * t);
    }
    a[j] = out;
  }
}

// pass 1: per (chunk-of-lookups, dim) histogram over the 2^16 addresses, 16-bit counters packed in
// 32-bit shared-memory words (a chunk holds < 65536 lookups)
// (memories: grid row i handles memory t0 + i * tstep; hist / base / read_ts / final_cts are indexed by the row, so a rank
// that owns every tstep-th memory keeps compact outputs)
__global__ void __launch_bounds__(256) lasso_hist_kernel(uint32_t m, uint32_t chunk_len,
                                                         const uint32_t* __restrict__ dims,
                                                         uint32_t* __restrict__ hist, int t0, int tstep) {
  extern __shared__ uint32_t sh[];  // 32768 words
  const uint32_t ch = blockIdx.x, t = blockIdx.y, nch = gridDim.x;
  for (uint32_t i = threadIdx.x; i < SUB_SIZE / 2; i += blockDim.x) sh[i] = 0;
  __syncthreads();
  const uint32_t* d = dims + (size_t)(t0 + t * tstep) * m + (size_t)ch * chunk_len;
  for (uint32_t i = threadIdx.x; i < chunk_len; i += blockDim.x) {
    const uint32_t addr = d[i];
    atomicAdd(&sh[addr >> 1], 1u << (16 * (addr & 1)));
  }
  __syncthreads();
  uint32_t* out = hist + ((size_t)t * nch + ch) * (SUB_SIZE / 2);
  for (uint32_t i = threadIdx.x; i < SUB_SIZE / 2; i += blockDim.x) out[i] = sh[i];
}

// pass 2: per (dim, address) exclusive scan over chunks -> base[t][chunk][addr]; total -> final_cts
__global__ void lasso_colscan_kernel(uint32_t nch, const uint32_t* __restrict__ hist, uint32_t* __restrict__ base,
                                     uint32_t* __restrict__ final_cts) {
  const uint32_t addr = blockIdx.x * blockDim.x + threadIdx.x, t = blockIdx.y;
  if (addr >= SUB_SIZE) return;
  uint32_t run = 0;
  for (uint32_t ch = 0; ch < nch; ++ch) {
    const uint32_t w = hist[((size_t)t * nch + ch) * (SUB_SIZE / 2) + (addr >> 1)];
    const uint32_t cnt = (w >> (16 * (addr & 1))) & 0xffff;
    base[((size_t)t * nch + ch) * SUB_SIZE + addr] = run;
    run += cnt;
  }
  final_cts[(size_t)t * SUB_SIZE + addr] = run;
}

// pass 3: ONE warp per (chunk, dim) walks its lookups in order, 32 at a time:
// read_ts[j] = base + (#earlier in chunk) + (#earlier lanes of this step with the same address)
__global__ void __launch_bounds__(32) lasso_rank_kernel(uint32_t m, uint32_t chunk_len,
                                                        const uint32_t* __restrict__ dims,
                                                        const uint32_t* __restrict__ base,
                                                        uint32_t* __restrict__ read_ts, int t0, int tstep) {
  extern __shared__ uint32_t sh[];
  uint16_t* local = reinterpret_cast<uint16_t*>(sh);
  const uint32_t ch = blockIdx.x, t = blockIdx.y, nch = gridDim.x, lane = threadIdx.x;
  for (uint32_t i = lane; i < SUB_SIZE / 2; i += 32) sh[i] = 0;
  __syncwarp();
  const size_t off = (size_t)t * m + (size_t)ch * chunk_len;                          // compact output row
  const size_t doff = (size_t)(t0 + t * tstep) * m + (size_t)ch * chunk_len;          // the memory's addresses
  const uint32_t* b = base + ((size_t)t * nch + ch) * SUB_SIZE;
  constexpr int U = 8;  // steps whose address + base loads are issued together (hides the global latency)
  for (uint32_t i0 = 0; i0 < chunk_len; i0 += 32 * U) {
    uint32_t addr[U], bs[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t i = i0 + u * 32 + lane;
      addr[u] = i < chunk_len ? dims[doff + i] : (0x80000000u | lane);  // invalid lanes never match
    }
#pragma unroll
    for (int u = 0; u < U; ++u) bs[u] = addr[u] < SUB_SIZE ? b[addr[u]] : 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const bool valid = addr[u] < SUB_SIZE;
      const uint32_t mask = __match_any_sync(0xffffffffu, addr[u]);
      const uint32_t before = __popc(mask & ((1u << lane) - 1));
      uint32_t seen = 0;
      if (valid) seen = local[addr[u]];
      __syncwarp();
      if (valid) {
        read_ts[off + i0 + u * 32 + lane] = bs[u] + seen + before;
        if (before == 0) local[addr[u]] = (uint16_t)(seen + __popc(mask));
      }
      __syncwarp();
    }
  }
}

__global__ void u32_to_fr_kernel(const uint32_t* __restrict__ in, Fr* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    fe_st(out + i, fe_from_u64<FrP>(in[i]));
}

// Sharded layout (shard.cu): rank r keeps the entries whose index bits [p, p + g) equal r, compactly; local index i
// stands for the global index below. g = 0 is the identity (single GPU).
__device__ __forceinline__ uint32_t shard_map(uint32_t i, int p, int g, uint32_t rank) {
  return g ? (((((i >> p) << g) | rank) << p) | (i & ((1u << p) - 1))) : i;
}
// out[t][i] = in[t][map(i)] for ntab integer tables of n entries each (n_loc = n >> g local entries)
__global__ void u32_to_fr_map_kernel(const uint32_t* __restrict__ in, Fr* __restrict__ out, int ntab, uint32_t n,
                                     uint32_t n_loc, int p, int g, uint32_t rank) {
  const uint32_t t = blockIdx.y, stride = gridDim.x * blockDim.x;  // grid row = table
  const uint32_t* __restrict__ src = in + (size_t)t * n;
  Fr* __restrict__ dst = out + (size_t)t * n_loc;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_loc; i += stride)
    fe_st(dst + i, fe_from_u64<FrP>(src[shard_map(i, p, g, rank)]));
}
__global__ void u64_to_fr_map_kernel(const uint64_t* __restrict__ in, Fr* __restrict__ out, uint32_t n_loc, int p, int g,
                                     uint32_t rank) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_loc; i += stride)
    fe_st(out + i, fe_from_u64<FrP>(in[shard_map((uint32_t)i, p, g, rank)]));
}

// m-sized leaves: read = dim*g^2 + e*g + ts - tau, write = read + 1 (trees 2t, 2t+1), stored in the
// leaf layer [2^h, 2^(h+1)) of each tree array
__global__ void __launch_bounds__(256) lasso_leaves_m_kernel(int c, uint32_t m, const Fr* __restrict__ dim_fr,
                                                             const Fr* __restrict__ e_fr, const Fr* __restrict__ ts_fr,
                                                             const Fr* __restrict__ gt, Fr* __restrict__ trees) {
  const Fr g = fe_ld(gt), tau = fe_ld(gt + 1);
  const Fr g2 = g * g, one = fe_one<FrP>();
  const int t = blockIdx.y;
  const uint32_t stride = gridDim.x * blockDim.x;
  Fr* rd = trees + (size_t)(2 * t) * 2 * m + m;
  Fr* wr = trees + (size_t)(2 * t + 1) * 2 * m + m;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < m; j += stride) {
    const size_t k = (size_t)t * m + j;
    const Fr v = fe_ldg(dim_fr + k) * g2 + fe_ldg(e_fr + k) * g + fe_ldg(ts_fr + k) - tau;
    fe_st(rd + j, v);
    fe_st(wr + j, v + one);
  }
}
// S-sized leaves: init = x*g^2 + T[x]*g - tau, final = init + final_cts
// (the rank's S_loc = 2^16 >> sg entries when the subtable trees are sharded; cts_fr is always the full table)
__global__ void __launch_bounds__(256) lasso_leaves_s_kernel(LassoTab tb, const Fr* __restrict__ cts_fr,
                                                             const Fr* __restrict__ gt, Fr* __restrict__ trees,
                                                             uint32_t S_loc, int sp, int sg, uint32_t rank) {
  const Fr g = fe_ld(gt), tau = fe_ld(gt + 1);
  const Fr g2 = g * g;
  const int t = blockIdx.y;
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S_loc) return;
  const uint32_t x = shard_map(i, sp, sg, rank);
  Fr* in = trees + (size_t)(2 * t) * 2 * S_loc + S_loc;
  Fr* fi = trees + (size_t)(2 * t + 1) * 2 * S_loc + S_loc;
  const Fr v = fe_from_u64<FrP>(x) * g2 + fe_from_u64<FrP>(lasso_subtable(tb, x)) * g - tau;
  fe_st(in + i, v);
  fe_st(fi + i, v + fe_ldg(cts_fr + (size_t)t * SUB_SIZE + x));
}

// one tree layer for a group of equally sized trees: V_k[i] = V_{k+1}[i] * V_{k+1}[i + 2^k]; tree
// arrays are heap-ordered (layer k at [2^k, 2^(k+1)))
__global__ void __launch_bounds__(256) tree_up_kernel(Fr* __restrict__ trees, size_t tree_stride, uint32_t half) {
  pdl_prologue();
  Fr* tr = trees + (size_t)blockIdx.y * tree_stride;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride)
    fe_st(tr + half + i, fe_ld(tr + 2 * (size_t)half + i) * fe_ld(tr + 3 * (size_t)half + i));
}

// ---- grand-product bookkeeping kernels (single warp, warp-cooperative transcript) ----------------
struct GpState {
  Fr claims[SC_MAX_TERMS];   // indexed by TREE id
  Fr weights[SC_MAX_TERMS];  // indexed by slot in the active list of the current layer
  Fr claim;
  Fr y[32];
  Fr evals[2 * SC_MAX_TERMS];  // (l, r) per active slot
};
struct GpTrees {
  const Fr* base[SC_MAX_TERMS];  // heap-ordered tree arrays (sharded trees: layers <= k0 only, replicated)
  int height[SC_MAX_TERMS];
  int T;
  // sharded provers: trees higher than k0 keep their layers >= k0 on the rank's slice only — loc[t] is the local heap
  // (layer k at [2^(k-g), 2^(k-g+1))), and the sum-checks of the layers k >= k0 run on those slices (shard.cu)
  const Fr* loc[SC_MAX_TERMS];
  int k0 = 0, g = 0;
};

// write the roots, which are the first claims
__global__ void gp_roots_kernel(Transcript* tr, GpTrees trees, GpState* st) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  trw_copy(&sh_tr, tr);
  for (int t = 0; t < trees.T; ++t) {
    const Fr root = fe_ld(trees.base[t] + 1);
    if (threadIdx.x == 0) fe_st(&st->claims[t], root);
    trw_write_fe(&sh_tr, root);
  }
  trw_copy(tr, &sh_tr);
}

// Layer k, before the sum-check. k == 0: the two children are the evaluations. k > 0: squeeze gamma,
// weights = gamma^slot over the ACTIVE trees (height > k, input order), claim = Σ weights * claims.
__global__ void gp_before_kernel(Transcript* tr, GpTrees trees, int k, GpState* st) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  const int lane = threadIdx.x;
  if (k == 0) {
    int slot = 0;
    for (int t = 0; t < trees.T; ++t) {
      if (trees.height[t] <= 0) continue;
      if (lane < 2) fe_st(&st->evals[2 * slot + lane], fe_ld(trees.base[t] + 2 + lane));
      ++slot;
    }
    return;
  }
  trw_copy(&sh_tr, tr);
  const Fr gamma = trw_squeeze(&sh_tr);
  Fr pw = fe_one<FrP>(), claim = fe_zero<FrP>();
  int slot = 0;
  for (int t = 0; t < trees.T; ++t) {
    if (trees.height[t] <= k) continue;
    if (lane == 0) fe_st(&st->weights[slot], pw);
    claim = claim + fr_mul_ni(pw, fe_ld(&st->claims[t]));
    pw = fr_mul_ni(pw, gamma);
    ++slot;
  }
  trw_copy(tr, &sh_tr);
  if (lane == 0) fe_st(&st->claim, claim);
}
// after the sum-check: write the 2A evaluations, squeeze mu, fold the active claims, y = x || mu
__global__ void gp_after_kernel(Transcript* tr, GpTrees trees, int k, const Fr* x /* k challenges */, GpState* st) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  const int lane = threadIdx.x;
  trw_copy(&sh_tr, tr);
  int A = 0;
  for (int t = 0; t < trees.T; ++t) A += trees.height[t] > k;
  // canonical forms of up to 32 evaluations per pass in parallel lanes
  for (int base = 0; base < 2 * A; base += 32) {
    const int i = base + lane;
    const Fr canon = fr_canon_ni(i < 2 * A ? fe_ld(&st->evals[i]) : fe_zero<FrP>());
    const int cnt = 2 * A - base < 32 ? 2 * A - base : 32;
    for (int j = 0; j < cnt; ++j) trw_write_canon_from_lane(&sh_tr, canon, j, true);
  }
  const Fr mu = trw_squeeze(&sh_tr);
  int slot = 0;
  for (int t = 0; t < trees.T; ++t) {
    if (trees.height[t] <= k) continue;
    if (lane == (slot & 31)) {
      const Fr l = fe_ld(&st->evals[2 * slot]), r = fe_ld(&st->evals[2 * slot + 1]);
      fe_st(&st->claims[t], l + (r - l) * mu);
    }
    ++slot;
  }
  for (int i = lane; i < k; i += 32) fe_st(&st->y[i], fe_ld(x + i));
  trw_copy(tr, &sh_tr);
  if (lane == 0) fe_st(&st->y[k], mu);
}

__global__ void gather_points_kernel(const G1Aff* src, const int* idx, int n, G1Aff* dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  fe_st(&dst[i].x, fe_ld(&src[idx[i]].x));
  fe_st(&dst[i].y, fe_ld(&src[idx[i]].y));
}
__global__ void copy_fr_kernel(const Fr* src, Fr* dst, int n) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(dst + i, fe_ld(src + i));
}

// Batched product argument over heap-ordered trees of possibly DIFFERENT heights (oracle/lasso.hpp
// grand_product_prove). After layer h-1 the running point (h coordinates) is copied to point_out[h]
// when that pointer is non-null; the per-tree claims stay in st->claims.
static int grand_product_prove(Ctx* c, const GpTrees& trees, GpState* st, Fr* scratch_x, Fr* const* point_out) {
  cudaStream_t s = c->stream;
  int h = 0;
  for (int t = 0; t < trees.T; ++t) h = trees.height[t] > h ? trees.height[t] : h;
  CUDA_TRY(launch_pdl(gp_roots_kernel, dim3(1), dim3(32), 0, s, c->d_tr, trees, st));
  count_launch(c);
  for (int k = 0; k < h; ++k) {
    CUDA_TRY(launch_pdl(gp_before_kernel, dim3(1), dim3(32), 0, s, c->d_tr, trees, k, st));
    count_launch(c);
    if (k > 0) {
      ScEvalJob job;
      job.num_vars = k;
      job.NP = 2;
      int A = 0;
      for (int t = 0; t < trees.T; ++t) {
        if (trees.height[t] <= k) continue;
        const Fr* l = trees.base[t] + ((size_t)2 << k);
        job.tables[2 * A] = l;
        job.tables[2 * A + 1] = l + ((size_t)1 << k);
        ++A;
      }
      job.T = A;
      job.weights = st->weights;
      job.eq_point = st->y;
      job.claim = &st->claim;
      job.challenges_out = scratch_x;
      job.evals_out = st->evals;
      int rc;
      if (trees.k0 > 0 && k >= trees.k0) {  // every active tree is sharded: children halves of the LOCAL layer k + 1
        A = 0;
        for (int t = 0; t < trees.T; ++t) {
          if (trees.height[t] <= k) continue;
          const Fr* l = trees.loc[t] + ((size_t)2 << (k - trees.g));
          job.tables[2 * A] = l;
          job.tables[2 * A + 1] = l + ((size_t)1 << (k - trees.g));
          ++A;
        }
        rc = sumcheck_prove_evals_sharded(c, job, k, trees.k0 - trees.g, -1);
      } else {
        rc = sumcheck_prove_evals_dist(c, job);
      }
      if (rc) return rc;
    }
    CUDA_TRY(launch_pdl(gp_after_kernel, dim3(1), dim3(32), 0, s, c->d_tr, trees, k, scratch_x, st));
    count_launch(c);
    if (point_out[k + 1]) {
      CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, st->y, point_out[k + 1], k + 1));
      count_launch(c);
    }
  }
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}


// integer witness of all 2^mu lookups (dims / E / read_ts / final_cts / a), validated operands
struct LassoWitness {
  uint32_t *dims, *es, *ts, *cts;
  uint64_t* a_u64;
};
// shard_counters: the read / final counters of the c memories are independent, so rank r computes those of the memories
// r, r + G, ... only and the integer arrays are all-gathered into the bulk arenas (they are needed by every rank for the
// point-sharded commitment MSMs and for its slice of the field-element tables)
static int lasso_witness_ints(Ctx* c, DevScope& mem, const LassoTab& tb, int mu, const uint64_t* d_xs, const uint64_t* d_ys,
                              LassoWitness* w, bool shard_counters = false) {
  const int C_ = tb.c;
  cudaStream_t s = c->stream;
  const uint32_t m = 1u << mu;
  const size_t S = SUB_SIZE;
  uint32_t *hist, *base;
  unsigned int* bad;
  const uint32_t nch = m >= (1u << 13) ? (m / 32768 > 128 ? m / 32768 : 128) : 1;
  const uint32_t chunk_len = m / nch;
  const int G = c->peer.world;
  if (shard_counters && (G < 2 || C_ % G || m < 8 || (size_t)C_ * m * 4 > c->peer.arena_half)) shard_counters = false;
  const int own = shard_counters ? C_ / G : C_, t0 = shard_counters ? c->peer.rank : 0, tstep = shard_counters ? G : 1;
  CUDA_TRY(mem.alloc(&w->dims, (size_t)C_ * m * 4));
  CUDA_TRY(mem.alloc(&w->es, (size_t)C_ * m * 4));
  CUDA_TRY(mem.alloc(&w->ts, (size_t)own * m * 4));
  CUDA_TRY(mem.alloc(&w->cts, (size_t)own * S * 4));
  CUDA_TRY(mem.alloc(&w->a_u64, (size_t)m * 8));
  CUDA_TRY(mem.alloc(&hist, (size_t)own * nch * (S / 2) * 4));
  CUDA_TRY(mem.alloc(&base, (size_t)own * nch * S * 4));
  CUDA_TRY(mem.alloc(&bad, 4));
  CUDA_TRY(cudaMemsetAsync(bad, 0, 4, s));
  int bx = (int)((m + 255) / 256);
  if (bx > NUM_SMS * 8) bx = NUM_SMS * 8;
  trace_point(c, "witness: allocated");
  lasso_chunks_kernel<<<bx, 256, 0, s>>>(tb, m, d_xs, d_ys, w->dims, w->es, w->a_u64, bad);
  // an operand outside the table is an error BEFORE anything reaches the transcript (one 4-byte read-back, the only
  // host round trip of the proof; the witness kernels below are already queued behind it)
  unsigned int h_bad = 0;
  CUDA_TRY(cudaMemcpyAsync(&h_bad, bad, 4, cudaMemcpyDeviceToHost, s));
  trace_point(c, "witness: flag copy queued");
  const int smem = (int)(S / 2) * 4;  // 128 KiB (opted in once per device, preload_lasso)
  lasso_hist_kernel<<<dim3(nch, own), 256, smem, s>>>(m, chunk_len, w->dims, hist, t0, tstep);
  lasso_colscan_kernel<<<dim3(S / 256, own), 256, 0, s>>>(nch, hist, base, w->cts);
  lasso_rank_kernel<<<dim3(nch, own), 32, smem, s>>>(m, chunk_len, w->dims, base, w->ts, t0, tstep);
  count_launch(c, 4);
  CUDA_TRY(cudaGetLastError());
  if (shard_counters) {
    // 8 counters travel as one field-element-sized word; row i of rank r is memory r + G i, and the gathered layout
    // (row, rank, entry) is exactly (memory, entry)
    const Fr* src[8];
    const Fr* full[8];
    int lq = 0;
    while ((8u << lq) < m) ++lq;
    for (int i = 0; i < own; ++i) src[i] = reinterpret_cast<const Fr*>(w->ts + (size_t)i * m);
    int rc = shard_allgather(c, src, own, m / 8, lq, false, full);
    if (rc) return rc;
    w->ts = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(full[0]));
    for (int i = 0; i < own; ++i) src[i] = reinterpret_cast<const Fr*>(w->cts + (size_t)i * S);
    rc = shard_allgather(c, src, own, (uint32_t)(S / 8), 13, false, full);
    if (rc) return rc;
    w->cts = const_cast<uint32_t*>(reinterpret_cast<const uint32_t*>(full[0]));
  }
  trace_point(c, "witness: kernels queued");
  CUDA_TRY(cudaStreamSynchronize(s));
  trace_point(c, "witness: synchronised");
  return h_bad ? B200_ERR_LOOKUP : B200_OK;
}

// built-in kinds as a LassoTab; a descriptor (kind 3) is validated here
static int lasso_tab(int kind, int chunks, const LassoTableDesc* desc, LassoTab* tb) {
  if (desc) {
    if (desc->chunks < 2 || desc->chunks > 8 || !desc->d_values || desc->num_operands < 1 || desc->num_operands > 2 ||
        desc->operand_bits < 1 || desc->num_operands * desc->operand_bits > SUB_VARS || desc->operand_bits * desc->chunks > 64 ||
        desc->out_bits < 1 || desc->out_bits > 32 || desc->out_bits * (desc->chunks - 1) + desc->value_bits > 64)
      return B200_ERR_ARG;
    *tb = LassoTab{3, desc->chunks, desc->num_operands, desc->operand_bits, desc->out_bits, desc->d_values};
    return B200_OK;
  }
  if (kind < 0 || kind > 2 || chunks < 1 || chunks > 8 || (kind == 0 && chunks > 4)) return B200_ERR_ARG;
  *tb = LassoTab{kind, chunks, kind == 0 ? 1 : 2, kind == 0 ? 16 : 8, kind == 0 ? 16 : 8, nullptr};
  return B200_OK;
}

int lasso_prove(Ctx* c, int kind, int chunks, int mu, const uint64_t* d_xs, const uint64_t* d_ys,
                const LassoTableDesc* desc) {
  LassoTab tb;
  if (lasso_tab(kind, chunks, desc, &tb)) return B200_ERR_ARG;
  kind = tb.kind;
  chunks = tb.c;
  if (mu < 1 || mu > 26) return B200_ERR_ARG;
  if (tb.nops == 2 && !d_ys) return B200_ERR_ARG;  // two-operand tables need both operands
  if (chunks < 2) return B200_ERR_ARG;  // additive batch_open needs >= 2 evaluations (pcs/multilinear.rs:150)
  if ((int)c->srs.size() <= (mu > SUB_VARS ? mu : SUB_VARS)) return B200_ERR_ARG;
  NvtxRange nvtx("lasso_prove-%d", mu);
  trace_point(c, "lasso_prove: enter");
  cudaStream_t s = c->stream;
  const int C_ = chunks;
  const uint32_t m = 1u << mu;
  const size_t S = SUB_SIZE;
  // ---- sharding geometry (b200_dist_shard_lasso): index window [p, p + g) -> rank -----------------------------
  const int G = c->peer.world;
  int g = 0;
  while ((1 << g) < G) ++g;
  const int K0 = c->shard_lasso_k0;
  const bool sh = K0 > 0 && G > 1 && mu > K0;        // m-sized tables / trees live on the rank's slice
  const bool s_sh = sh && SUB_VARS > K0;              // so do the subtable-sized trees (small K0: tests)
  const int p = sh ? K0 - g : 0, lg = sh ? g : 0;
  const uint32_t rank = sh ? (uint32_t)c->peer.rank : 0;
  const uint32_t m_loc = m >> lg;
  const uint32_t S_loc = s_sh ? (uint32_t)(S >> g) : (uint32_t)S;
  const bool saved_shard_commits = c->shard_commits;
  if (sh) c->shard_commits = true;
  struct Restore {
    Ctx* c;
    bool v;
    ~Restore() { c->shard_commits = v; }
  } restore{c, saved_shard_commits};
  HeartbeatScope hb(c);
  DevScope mem(s);

  // ---- 1. witness (integers, all lookups) ------------------------------------------------------------
  int ph = prof_begin(c, PH_WITNESS);
  LassoWitness wit;
  int rc = lasso_witness_ints(c, mem, tb, mu, d_xs, d_ys, &wit, sh);
  if (rc) return rc;
  uint32_t *dims = wit.dims, *es = wit.es, *ts = wit.ts, *cts = wit.cts;
  uint64_t* a_u64 = wit.a_u64;
  // field-element tables of the rank's slice: a | dim[c] | e[c] | ts[c] (m_loc each); cts[c] (S each, always full)
  const int NM = 1 + 3 * C_;
  Fr *mt, *st_tabs;
  CUDA_TRY(mem.alloc(&mt, (size_t)NM * m_loc * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&st_tabs, (size_t)C_ * S * sizeof(Fr)));
  Fr* a_fr = mt;
  Fr* dim_fr = mt + (size_t)m_loc;
  Fr* e_fr = dim_fr + (size_t)C_ * m_loc;
  Fr* ts_fr = e_fr + (size_t)C_ * m_loc;
  {
    const int bx = NUM_SMS * 8;
    int by = (int)((m_loc + 255) / 256);
    if (by > NUM_SMS * 2) by = NUM_SMS * 2;
    u64_to_fr_map_kernel<<<bx, 256, 0, s>>>(a_u64, a_fr, m_loc, p, lg, rank);
    u32_to_fr_map_kernel<<<dim3(by, C_), 256, 0, s>>>(dims, dim_fr, C_, m, m_loc, p, lg, rank);
    u32_to_fr_map_kernel<<<dim3(by, C_), 256, 0, s>>>(es, e_fr, C_, m, m_loc, p, lg, rank);
    u32_to_fr_map_kernel<<<dim3(by, C_), 256, 0, s>>>(ts, ts_fr, C_, m, m_loc, p, lg, rank);
    u32_to_fr_kernel<<<bx, 256, 0, s>>>(cts, st_tabs, (size_t)C_ * S);
    count_launch(c, 5);
  }

  // ---- scalar arena -------------------------------------------------------------------------------
  // stmt[7] | r[mu] | v_a | pw[c] | x_p[mu] | e_p[c] | gt[2] | x_scratch[32] | ev_m[3c] | ev_s[c] | pts[3*mu]
  const size_t arena_n = 7 + mu + 1 + C_ + mu + C_ + 2 + 32 + 3 * C_ + C_ + 3 * (size_t)mu + SUB_VARS;
  Fr* arena;
  CUDA_TRY(mem.alloc(&arena, arena_n * sizeof(Fr)));
  Fr* stmt = arena;
  Fr* r = stmt + 7;
  Fr* v_a = r + mu;
  Fr* pw = v_a + 1;
  Fr* x_p = pw + C_;
  Fr* e_p = x_p + mu;
  Fr* gt = e_p + C_;
  Fr* x_scratch = gt + 2;
  Fr* ev_m = x_scratch + 32;
  Fr* ev_s = ev_m + 3 * C_;
  Fr* pts = ev_s + C_;
  Fr* pt_s = pts + 3 * (size_t)mu;
  GpState* gp;
  CUDA_TRY(mem.alloc(&gp, sizeof(GpState)));
  const int out_bits = tb.out_bits;
  const int nstmt = desc ? 7 : 3;  // a table given as data is part of the statement
  {
    Fr h[7 + 8];
    h[0] = fe_from_u64<FrP>((uint64_t)kind);
    h[1] = fe_from_u64<FrP>((uint64_t)C_);
    h[2] = fe_from_u64<FrP>((uint64_t)mu);
    if (desc) {
      h[3] = fe_from_u64<FrP>((uint64_t)tb.nops);
      h[4] = fe_from_u64<FrP>((uint64_t)tb.op_bits);
      h[5] = fe_from_u64<FrP>((uint64_t)tb.out_bits);
      h[6] = desc->digest;
    }
    for (int t = 0; t < C_; ++t) h[7 + t] = fe_from_u64<FrP>((uint64_t)1 << (out_bits * t));
    CUDA_TRY(cudaMemcpyAsync(stmt, h, nstmt * sizeof(Fr), cudaMemcpyHostToDevice, s));
    CUDA_TRY(cudaMemcpyAsync(pw, h + 7, C_ * sizeof(Fr), cudaMemcpyHostToDevice, s));
  }
  rc = transcript_op(c, TR_COMMON, stmt, nullptr, nstmt);
  if (rc) return rc;
  prof_end(c, ph);
  ph = prof_begin(c, PH_COMMIT);

  // ---- 2. commitments: a, dim_*, E_*, read_ts_* (level mu), final_cts_* (level 16) -----------------
  // MSMs run on the integer witness (small scalars populate few windows). Two commitments are free:
  //  * a = Σ_t 2^(w t) E_t as polynomials, so Com(a) = Σ_t 2^(w t) Com(E_t): doublings instead of an MSM;
  //  * identity subtable (range): E_t == dim_t, the same point is written twice.
  {
    MsmJob jobs[4 * 8];
    int J = 0;
    const bool e_is_dim = kind == 0;
    const int j_dim = J;
    for (int t = 0; t < C_; ++t) jobs[J++] = MsmJob{dims + (size_t)t * m, c->srs[mu], m, MSM_U32, 16, nullptr};
    const int j_e = e_is_dim ? j_dim : J;
    // E_t = T[dim_t] over the same bases as dim_t: its commitment is a regrouping of dim_t's bucket sums by table value
    // (MsmJob::group_*) — no pass over the points at all. Needs T[0] = 0 and enough points per rank to pay.
    const uint32_t* gperm = desc ? desc->d_group_perm : c->d_group_perm[kind];
    const uint32_t* goff = desc ? desc->d_group_off : c->d_group_off[kind];
    const int ngroups = desc ? desc->ngroups : (kind == 0 ? 0 : 255);
    const uint64_t pts_per_rank = c->shard_commits && c->peer.world > 1 && m % c->peer.world == 0 ? m / c->peer.world : m;
    const bool grouped = !e_is_dim && gperm && goff && ngroups > 0 && pts_per_rank >= msm_group_min_points();
    if (!e_is_dim)
      for (int t = 0; t < C_; ++t) {
        MsmJob ej{es + (size_t)t * m, c->srs[mu], m, MSM_U32, desc ? (desc->value_bits > 0 ? desc->value_bits : 1) : out_bits, nullptr};
        if (grouped) {
          ej = MsmJob{nullptr, nullptr, 0, MSM_U32, 1, nullptr};
          ej.group_src = j_dim + t;
          ej.ngroups = ngroups;
          ej.group_perm = gperm;
          ej.group_off = goff;
        }
        jobs[J++] = ej;
      }
    const int j_ts = J;
    for (int t = 0; t < C_; ++t) jobs[J++] = MsmJob{ts + (size_t)t * m, c->srs[mu], m, MSM_U32, mu + 1, nullptr};
    const int j_cts = J;
    for (int t = 0; t < C_; ++t) jobs[J++] = MsmJob{cts + (size_t)t * S, c->srs[SUB_VARS], S, MSM_U32, mu + 1, nullptr};
    MsmDerive da;
    da.nsrc = C_;
    da.shift = out_bits;
    for (int t = 0; t < C_; ++t) da.src[t] = j_e + t;
    G1Aff *comms, *all;
    const int NC = 1 + 4 * C_;
    CUDA_TRY(mem.alloc(&comms, (J + 1) * sizeof(G1Aff)));
    CUDA_TRY(mem.alloc(&all, NC * sizeof(G1Aff)));
    trace_point(c, "commit: msm enqueue");
    rc = msm_batch_dist(c, jobs, J, comms, &da, 1);
    if (rc) return rc;
    trace_point(c, "commit: msm enqueued");
    // transcript order: a | dim[c] | E[c] | read_ts[c] | final_cts[c]
    int h_src[1 + 4 * 8];
    h_src[0] = J;
    for (int t = 0; t < C_; ++t) {
      h_src[1 + t] = j_dim + t;
      h_src[1 + C_ + t] = j_e + t;
      h_src[1 + 2 * C_ + t] = j_ts + t;
      h_src[1 + 3 * C_ + t] = j_cts + t;
    }
    int* d_src;
    CUDA_TRY(mem.alloc(&d_src, NC * sizeof(int)));
    CUDA_TRY(cudaMemcpyAsync(d_src, h_src, NC * sizeof(int), cudaMemcpyHostToDevice, s));
    gather_points_kernel<<<1, 64, 0, s>>>(comms, d_src, NC, all);
    count_launch(c);
    rc = transcript_write_points(c, all, NC);
    if (rc) return rc;
  }

  prof_end(c, ph);
  ph = prof_begin(c, PH_PRIMARY);
  // ---- 3-5. primary Surge sum-check ----------------------------------------------------------------
  rc = transcript_op(c, TR_SQUEEZE, nullptr, r, mu);
  if (rc) return rc;
  {
    const Fr* tab[1] = {a_fr};
    rc = sh ? mle_eval_many_sharded(c, tab, 1, mu, p, r, v_a) : mle_eval_many(c, tab, 1, mu, r, v_a);
    if (rc) return rc;
  }
  rc = transcript_op(c, TR_WRITE, v_a, nullptr, 1);
  if (rc) return rc;
  {
    ScEvalJob job;
    job.num_vars = mu;
    job.T = C_;
    job.NP = 1;
    for (int t = 0; t < C_; ++t) job.tables[t] = e_fr + (size_t)t * m_loc;
    job.weights = pw;
    job.eq_point = r;
    job.claim = v_a;
    job.challenges_out = x_p;
    job.evals_out = e_p;
    rc = sh ? sumcheck_prove_evals_sharded(c, job, mu, p, -1) : sumcheck_prove_evals_dist(c, job);
    if (rc) return rc;
  }
  rc = transcript_op(c, TR_WRITE, e_p, nullptr, C_);
  if (rc) return rc;

  prof_end(c, ph);
  ph = prof_begin(c, PH_GKR_M);
  // ---- 6-8. memory checking -------------------------------------------------------------------------
  rc = transcript_op(c, TR_SQUEEZE, nullptr, gt, 2);
  if (rc) return rc;
  const int T = 2 * C_;
  // heaps: the rank's layers (all layers when not sharded), and for sharded trees a replicated heap of the layers <= K0
  Fr *mtrees, *strees, *mfull = nullptr, *sfull = nullptr;
  CUDA_TRY(mem.alloc(&mtrees, (size_t)T * 2 * m_loc * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&strees, (size_t)T * 2 * S_loc * sizeof(Fr)));
  if (sh) CUDA_TRY(mem.alloc(&mfull, ((size_t)T * 2 << K0) * sizeof(Fr)));
  if (s_sh) CUDA_TRY(mem.alloc(&sfull, ((size_t)T * 2 << K0) * sizeof(Fr)));
  {
    int bx = (int)((m_loc + 255) / 256);
    int cap = (8 * NUM_SMS + C_ - 1) / C_;
    if (bx > cap) bx = cap;
    lasso_leaves_m_kernel<<<dim3(bx, C_), 256, 0, s>>>(C_, m_loc, dim_fr, e_fr, ts_fr, gt, mtrees);
    lasso_leaves_s_kernel<<<dim3((S_loc + 255) / 256, C_), 256, 0, s>>>(tb, st_tabs, gt, strees, S_loc,
                                                                        s_sh ? p : 0, s_sh ? g : 0, rank);
    count_launch(c, 2);
  }
  {
    // product trees, then ONE batched GKR over the 4c trees. Sharded trees: local layers down to K0, one bulk
    // all-gather of layer K0 (2^K0 entries per tree), replicated layers above.
    auto build = [&](Fr* loc, uint32_t n_loc, int height, bool sharded, Fr* full) -> int {
      const int lgg = sharded ? g : 0, kmin = sharded ? K0 : 0;
      for (int k = height - 1; k >= kmin; --k) {
        const uint32_t half = 1u << (k - lgg);
        int bx = (int)((half + 255) / 256), cap = (4 * NUM_SMS + T - 1) / T;
        if (bx > cap) bx = cap;
        CUDA_TRY(launch_pdl(tree_up_kernel, dim3(dim3(bx, T)), dim3(256), 0, s, loc, (size_t)2 * n_loc, half));
        count_launch(c);
      }
      if (!sharded) return B200_OK;
      const Fr* src[SC_MAX_TERMS];
      const Fr* gathered[SC_MAX_TERMS];
      const uint32_t len = 1u << (K0 - g);
      for (int t = 0; t < T; ++t) src[t] = loc + (size_t)t * 2 * n_loc + len;  // local layer K0
      int rc2 = shard_allgather(c, src, T, len, K0 - g, false, gathered);
      if (rc2) return rc2;
      for (int t = 0; t < T; ++t)
        CUDA_TRY(cudaMemcpyAsync(full + ((size_t)t * 2 << K0) + ((size_t)1 << K0), gathered[t], sizeof(Fr) << K0,
                                 cudaMemcpyDeviceToDevice, s));
      for (int k = K0 - 1; k >= 0; --k) {
        const uint32_t half = 1u << k;
        int bx = (int)((half + 255) / 256), cap = (4 * NUM_SMS + T - 1) / T;
        if (bx > cap) bx = cap;
        CUDA_TRY(launch_pdl(tree_up_kernel, dim3(dim3(bx, T)), dim3(256), 0, s, full, (size_t)2 << K0, half));
        count_launch(c);
      }
      return B200_OK;
    };
    rc = build(mtrees, m_loc, mu, sh, mfull);
    if (rc) return rc;
    rc = build(strees, S_loc, SUB_VARS, s_sh, sfull);
    if (rc) return rc;
    GpTrees trees;
    trees.T = 2 * T;
    trees.k0 = sh ? K0 : 0;
    trees.g = sh ? g : 0;
    for (int t = 0; t < T; ++t) {
      trees.base[t] = sh ? mfull + ((size_t)t * 2 << K0) : mtrees + (size_t)t * 2 * m;
      trees.loc[t] = sh ? mtrees + (size_t)t * 2 * m_loc : nullptr;
      trees.height[t] = mu;
      trees.base[T + t] = s_sh ? sfull + ((size_t)t * 2 << K0) : strees + (size_t)t * 2 * S;
      trees.loc[T + t] = s_sh ? strees + (size_t)t * 2 * S_loc : nullptr;
      trees.height[T + t] = SUB_VARS;
    }
    Fr* point_out[33] = {nullptr};
    point_out[mu] = pts + 2 * (size_t)mu;  // x_m
    if (mu == SUB_VARS) {  // both leaf layers are reached at the same point
      rc = grand_product_prove(c, trees, gp, x_scratch, point_out);
      if (rc) return rc;
      CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, pts + 2 * (size_t)mu, pt_s, SUB_VARS));
      count_launch(c);
    } else {
      point_out[SUB_VARS] = pt_s;  // x_s
      rc = grand_product_prove(c, trees, gp, x_scratch, point_out);
      if (rc) return rc;
    }
  }

  prof_end(c, ph);
  trace_point(c, "gkr enqueued");
  ph = prof_begin(c, PH_LEAF_EVALS);
  // ---- 9. leaf openings -----------------------------------------------------------------------------
  {
    const Fr* tabs[3 * 8];
    for (int i = 0; i < 3 * C_; ++i) tabs[i] = dim_fr + (size_t)i * m_loc;  // dim*, e*, ts* are contiguous
    rc = sh ? mle_eval_many_sharded(c, tabs, 3 * C_, mu, p, pts + 2 * (size_t)mu, ev_m)
            : mle_eval_many(c, tabs, 3 * C_, mu, pts + 2 * (size_t)mu, ev_m);
    if (rc) return rc;
    for (int t = 0; t < C_; ++t) tabs[t] = st_tabs + (size_t)t * S;
    rc = mle_eval_many(c, tabs, C_, SUB_VARS, pt_s, ev_s);
    if (rc) return rc;
  }
  rc = transcript_op(c, TR_WRITE, ev_m, nullptr, 3 * C_);  // dims, E, read_ts
  if (rc) return rc;
  rc = transcript_op(c, TR_WRITE, ev_s, nullptr, C_);
  if (rc) return rc;

  prof_end(c, ph);
  ph = prof_begin(c, PH_OPEN_M);
  // ---- 10. batch openings ---------------------------------------------------------------------------
  {
    CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, r, pts, mu));
    CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, x_p, pts + mu, mu));
    count_launch(c, 2);
    const int E = 1 + 4 * C_;
    Fr* vals;
    CUDA_TRY(mem.alloc(&vals, E * sizeof(Fr)));
    CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, v_a, vals, 1));
    CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, e_p, vals + 1, C_));
    CUDA_TRY(launch_pdl(copy_fr_kernel, dim3(1), dim3(64), 0, s, ev_m, vals + 1 + C_, 3 * C_));
    count_launch(c, 3);
    const Fr* polys[1 + 3 * 8];
    for (int i = 0; i < NM; ++i) polys[i] = mt + (size_t)i * m_loc;
    int ev_poly[1 + 4 * 8], ev_point[1 + 4 * 8], k = 0;
    ev_poly[k] = 0, ev_point[k++] = 0;
    for (int t = 0; t < C_; ++t) ev_poly[k] = 1 + C_ + t, ev_point[k++] = 1;
    for (int t = 0; t < C_; ++t) ev_poly[k] = 1 + t, ev_point[k++] = 2;
    for (int t = 0; t < C_; ++t) ev_poly[k] = 1 + C_ + t, ev_point[k++] = 2;
    for (int t = 0; t < C_; ++t) ev_poly[k] = 1 + 2 * C_ + t, ev_point[k++] = 2;
    BatchOpenJob bj{mu, NM, 3, E, polys, pts, ev_poly, ev_point, vals};
    bj.shard_p = sh ? p : -1;
    rc = kzg_batch_open(c, bj);
    if (rc) return rc;
    prof_end(c, ph);
    ph = prof_begin(c, PH_OPEN_S);
    const Fr* spolys[8];
    int sp[8], spt[8];
    for (int t = 0; t < C_; ++t) spolys[t] = st_tabs + (size_t)t * S, sp[t] = t, spt[t] = 0;
    BatchOpenJob sj{SUB_VARS, C_, 1, C_, spolys, pt_s, sp, spt, ev_s};
    rc = kzg_batch_open(c, sj);
    if (rc) return rc;
    prof_end(c, ph);
  }
  CUDA_TRY(cudaGetLastError());
  trace_point(c, "lasso_prove: all enqueued");
  return B200_OK;
}

// witness only (parity of dims / E / read_ts / final_cts / a against the oracle)
int lasso_witness(Ctx* c, int kind, int chunks, int mu, const uint64_t* d_xs, const uint64_t* d_ys, Fr* d_mt,
                  Fr* d_st) {
  LassoTab tb;
  if (lasso_tab(kind, chunks, nullptr, &tb) || mu < 1 || mu > 26) return B200_ERR_ARG;
  if (kind != 0 && !d_ys) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  const int C_ = chunks;
  const uint32_t m = 1u << mu;
  const size_t S = SUB_SIZE;
  DevScope mem(s);
  LassoWitness w;
  int rc = lasso_witness_ints(c, mem, tb, mu, d_xs, d_ys, &w);
  if (rc) return rc;
  rc = fr_from_u64(c, w.a_u64, d_mt, m);
  if (rc) return rc;
  const int gx = NUM_SMS * 8;
  u32_to_fr_kernel<<<gx, 256, 0, s>>>(w.dims, d_mt + (size_t)m, (size_t)C_ * m);
  u32_to_fr_kernel<<<gx, 256, 0, s>>>(w.es, d_mt + (size_t)(1 + C_) * m, (size_t)C_ * m);
  u32_to_fr_kernel<<<gx, 256, 0, s>>>(w.ts, d_mt + (size_t)(1 + 2 * C_) * m, (size_t)C_ * m);
  u32_to_fr_kernel<<<gx, 256, 0, s>>>(w.cts, d_st, (size_t)C_ * S);
  count_launch(c, 4);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_lasso() {
  // cudaFuncSetAttribute waits for the device when kernels are in flight (measured: a rank of an in-process group
  // stalled here until its peers' collective timed out), so the 128 KiB opt-in happens once, at context creation
  cudaFuncSetAttribute(lasso_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SUB_SIZE / 2) * 4);
  cudaFuncSetAttribute(lasso_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(SUB_SIZE / 2) * 4);
  B200_PRELOAD(lasso_chunks_kernel);
  B200_PRELOAD(lasso_hist_kernel);
  B200_PRELOAD(lasso_colscan_kernel);
  B200_PRELOAD(lasso_rank_kernel);
  B200_PRELOAD(u32_to_fr_kernel);
  B200_PRELOAD(u32_to_fr_map_kernel);
  B200_PRELOAD(u64_to_fr_map_kernel);
  B200_PRELOAD(lasso_leaves_m_kernel);
  B200_PRELOAD(lasso_leaves_s_kernel);
  B200_PRELOAD(tree_up_kernel);
  B200_PRELOAD(gp_roots_kernel);
  B200_PRELOAD(gp_before_kernel);
  B200_PRELOAD(gp_after_kernel);
  B200_PRELOAD(gather_points_kernel);
  B200_PRELOAD(copy_fr_kernel);
}

}  // namespace b200

// Batched Pippenger MSM on G1 for MultilinearKzg commit / open
// (variable_base_msm, pb/util/arithmetic/msm.rs:84-181; call sites kzg.rs:255,271,292).
//
// The reference splits the input per thread and runs a serial bucket method per chunk. On the GPU
// the whole batch of MSMs of one protocol phase (all commitments, or all n quotient commitments of
// an opening) is processed by one sequence of launches:
//   1. digits   : signed c-bit digits per (scalar, window); per-bucket population via atomics, whose
//                 return value is the element's rank inside its bucket
//   2. scan     : exclusive scan of the bucket populations (and of the per-bucket task counts)
//   3. scatter  : counting-sort the (point index, sign) pairs by bucket
//   4. accumulate: one thread per TASK (<= TASK_LEN points of one bucket) sums with XYZZ mixed adds —
//                 splitting heavy buckets keeps skewed scalars (Lasso counters) load-balanced
//   5. buckets  : per-bucket sum of its task partials (one warp per heavy bucket)
//   6. windows  : Σ (k+1) B_k without scalar multiplications: running sums over 8 buckets per thread, then two
//                 warp-per-32-entries levels (suffix scan by shuffles) and a one-warp final combine per window
//   7. finish   : one warp per commitment: windows shifted in parallel lanes, shuffle sum, projective -> affine;
//                 derived commitments (linear combinations of other results of the batch) cost only doublings
// Scalars with few significant bits (dims, counters, subtable values) only populate the windows they
// need, which the reference cannot exploit. The group element is unique, so any schedule yields the
// reference's commitment bytes.
#include <stdlib.h>

#include "internal.h"

namespace b200 {

static const int TASK_LEN = 64;       // minimum task length; a batch uses a power of two in [64, 512]: about two tasks
                                      // per resident thread of the accumulate grid
static const int SEQ_TASKS = 8;     // buckets with more task partials than this go to the warp kernel
static const int MSM_MAX_JOBS = 64;
static const int MSM_MAX_WINDOWS = 128;

struct MsmJobDev {
  const void* scalars;
  const G1Aff* bases;
  uint32_t n;
  int kind;           // MsmScalarKind
  int c, W;           // window bits, number of windows
  uint32_t B;         // buckets per window = 2^(c-1)
  int hist;           // 1: small-integer scalars into <= HIST_MAX_B buckets per window: digits pass with CTA histograms
  int precomp;        // 1: bases = precomputed window multiples, all windows share one bucket set
  uint32_t ext_stride;  // precomp: entries per window of the extended table
  int Wred;           // windows that need a bucket reduction (1 when precomp, else W)
  uint32_t bucket_base;
  uint64_t pair_base;
  int map_p, map_g;     // map_g > 0: scalar i belongs to base point ((i >> p) << (p + g)) | (rank << p) | (i & (2^p - 1))
  uint32_t map_rank;    // (the rank's slice of a polynomial sharded on the index bits [p, p + g), shard.cu)
  uint32_t win_base;    // first window slot of this job
  // grouped job (MsmJob::group_src >= 0): its single reduction window has `B` = ngroups buckets filled by msm_group_kernel
  int group_src;
  const uint32_t* group_perm;
  const uint32_t* group_off;
};
struct MsmPlanDev {
  int J;
  uint32_t task_len;
  MsmJobDev job[MSM_MAX_JOBS];
};

// canonical little-endian limbs of scalar i
__device__ __forceinline__ void load_scalar(const MsmJobDev& jb, uint32_t i, uint32_t s[8]) {
  switch (jb.kind) {
    case MSM_FR_MONT: {
      Fr x = fe_to_canonical<FrP>(fe_ldg(reinterpret_cast<const Fr*>(jb.scalars) + i));
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = x.v[k];
      break;
    }
    case MSM_FR_CANON: {
      Fr x = fe_ldg(reinterpret_cast<const Fr*>(jb.scalars) + i);
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = x.v[k];
      break;
    }
    case MSM_U64: {
      uint64_t x = reinterpret_cast<const uint64_t*>(jb.scalars)[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = 0;
      s[0] = (uint32_t)x;
      s[1] = (uint32_t)(x >> 32);
      break;
    }
    default: {  // MSM_U32
      uint32_t x = reinterpret_cast<const uint32_t*>(jb.scalars)[i];
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] = 0;
      s[0] = x;
      break;
    }
  }
}

__device__ __forceinline__ uint32_t bits_at(const uint32_t s[8], int pos, int c) {
  // c <= 17 bits starting at bit `pos` (may straddle two limbs; bits >= 256 read as zero)
  if (pos >= 256) return 0;
  const int limb = pos >> 5, sh = pos & 31;
  uint64_t lo = s[limb];
  uint64_t hi = limb + 1 < 8 ? s[limb + 1] : 0;
  return (uint32_t)(((hi << 32) | lo) >> sh) & ((1u << c) - 1);
}

// pass 0: count + rank; pass 1: scatter with the scanned offsets
template <int PASS>
__global__ void __launch_bounds__(256) msm_digits_kernel(MsmPlanDev plan, uint32_t* __restrict__ cnt,
                                                         const uint32_t* __restrict__ boff,
                                                         uint32_t* __restrict__ ranks,
                                                         uint32_t* __restrict__ sorted) {
  const MsmJobDev& jb = plan.job[blockIdx.y];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (PASS == 0) {
    if (jb.hist) return;  // msm_digits_hist_kernel
    // Small-integer scalars (Lasso counters, 8-/16-bit subtable values) hit a handful of buckets: aggregate
    // the population atomics per warp (one atomicAdd per distinct bucket in the warp) instead of 32 colliding ones.
    const bool aggregate = jb.kind == MSM_U32 || jb.kind == MSM_U64;
    if (!aggregate && i >= jb.n) return;
    if ((i & ~31u) >= jb.n) return;  // whole warp past the end (warp-uniform)
    const bool valid = i < jb.n;
    uint32_t s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (valid) load_scalar(jb, i, s);
    const uint32_t lane = threadIdx.x & 31;
    uint32_t carry = 0;
    for (int w = 0; w < jb.W; ++w) {
      uint32_t raw = bits_at(s, w * jb.c, jb.c) + carry;
      int32_t d;
      if (raw > jb.B) {  // B = 2^(c-1): digits live in [-B, B]
        d = (int32_t)raw - (int32_t)(2 * jb.B);
        carry = 1;
      } else {
        d = (int32_t)raw;
        carry = 0;
      }
      uint32_t code = 0xffffffffu, gb = 0xffffffffu;
      if (valid && d != 0) {
        const uint32_t mag = d < 0 ? (uint32_t)(-d) : (uint32_t)d;
        gb = jb.bucket_base + (jb.precomp ? 0u : (uint32_t)w * jb.B) + (mag - 1);
      }
      uint32_t rank = 0;
      if (aggregate) {
        const uint32_t peers = __match_any_sync(0xffffffffu, gb);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (gb != 0xffffffffu && (int)lane == leader) base = atomicAdd(&cnt[gb], (uint32_t)__popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        rank = base + __popc(peers & ((1u << lane) - 1));
      } else if (gb != 0xffffffffu) {
        rank = atomicAdd(&cnt[gb], 1u);
      }
      if (valid) {
        if (gb != 0xffffffffu) code = rank | (d < 0 ? 0x80000000u : 0u);
        // 2 words per pair: bucket, rank|sign
        ranks[2 * (jb.pair_base + (uint64_t)w * jb.n + i)] = gb;
        ranks[2 * (jb.pair_base + (uint64_t)w * jb.n + i) + 1] = code;
      }
    }
  } else {
    if (i >= jb.n) return;
    for (int w = 0; w < jb.W; ++w) {
      const uint64_t p = 2 * (jb.pair_base + (uint64_t)w * jb.n + i);
      const uint32_t gb = ranks[p];
      if (gb == 0xffffffffu) continue;
      const uint32_t code = ranks[p + 1];
      // with precomputed tables the point of window w is entry i + w*n of the extended table
      const uint32_t bi = jb.map_g ? ((((i >> jb.map_p) << jb.map_g) | jb.map_rank) << jb.map_p) | (i & ((1u << jb.map_p) - 1)) : i;
      sorted[boff[gb] + (code & 0x7fffffffu)] = (jb.precomp ? bi + (uint32_t)w * jb.ext_stride : bi) | (code & 0x80000000u);
    }
  }
}

// Digits pass for small-integer scalars that fall into few buckets (Lasso counters, 8-bit subtable values: 2^22
// scalars into a few hundred buckets). Per-bucket global atomics, even warp-aggregated, serialise on those buckets;
// here a CTA takes 2048 scalars, ranks them inside the CTA with shared-memory atomics (one histogram per window) and
// claims its range of every non-empty bucket with ONE global atomic. Ranks inside a bucket are unique, not ordered —
// the bucket sum does not depend on the order.
static const uint32_t HIST_MAX_B = 4096;
static const int HIST_K = 8;  // scalars per thread
__global__ void __launch_bounds__(256) msm_digits_hist_kernel(MsmPlanDev plan, uint32_t* __restrict__ cnt,
                                                              uint32_t* __restrict__ ranks) {
  __shared__ uint32_t hist[HIST_MAX_B];
  const MsmJobDev& jb = plan.job[blockIdx.y];
  if (!jb.hist) return;
  const uint32_t base_i = blockIdx.x * (256 * HIST_K), tid = threadIdx.x;
  if (base_i >= jb.n) return;
  uint64_t sc[HIST_K];
#pragma unroll
  for (int k = 0; k < HIST_K; ++k) {
    const uint32_t i = base_i + k * 256 + tid;
    sc[k] = 0;
    if (i < jb.n)
      sc[k] = jb.kind == MSM_U64 ? reinterpret_cast<const uint64_t*>(jb.scalars)[i]
                                 : (uint64_t) reinterpret_cast<const uint32_t*>(jb.scalars)[i];
  }
  uint32_t carry = 0;  // bit k: carry of scalar k into the next window
  for (int w = 0; w < jb.W; ++w) {
    for (uint32_t b = tid; b < jb.B; b += 256) hist[b] = 0;
    __syncthreads();
    uint32_t lr[HIST_K];
    int32_t dd[HIST_K];
    const int sh = w * jb.c;
#pragma unroll
    for (int k = 0; k < HIST_K; ++k) {
      const uint32_t raw = (sh < 64 ? (uint32_t)((sc[k] >> sh) & ((1u << jb.c) - 1)) : 0u) + ((carry >> k) & 1u);
      int32_t d;
      if (raw > jb.B) {
        d = (int32_t)raw - (int32_t)(2 * jb.B);
        carry |= 1u << k;
      } else {
        d = (int32_t)raw;
        carry &= ~(1u << k);
      }
      dd[k] = d;
      lr[k] = 0;
      if (d != 0 && base_i + k * 256 + tid < jb.n) lr[k] = atomicAdd(&hist[(d < 0 ? -d : d) - 1], 1u);
    }
    __syncthreads();
    const uint32_t gb0 = jb.bucket_base + (uint32_t)w * jb.B;
    for (uint32_t b = tid; b < jb.B; b += 256) {
      const uint32_t h = hist[b];
      if (h) hist[b] = atomicAdd(&cnt[gb0 + b], h);
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < HIST_K; ++k) {
      const uint32_t i = base_i + k * 256 + tid;
      if (i >= jb.n) continue;
      uint32_t gb = 0xffffffffu, code = 0xffffffffu;
      if (dd[k] != 0) {
        const uint32_t mag = dd[k] < 0 ? (uint32_t)(-dd[k]) : (uint32_t)dd[k];
        gb = gb0 + mag - 1;
        code = (hist[mag - 1] + lr[k]) | (dd[k] < 0 ? 0x80000000u : 0u);
      }
      ranks[2 * (jb.pair_base + (uint64_t)w * jb.n + i)] = gb;
      ranks[2 * (jb.pair_base + (uint64_t)w * jb.n + i) + 1] = code;
    }
    __syncthreads();
  }
}

// ---- exclusive scan of uint32 (three launches; 1024 elements per block) ------------------------
__global__ void __launch_bounds__(256) scan_local_kernel(const uint32_t* __restrict__ in, uint32_t n,
                                                         uint32_t* __restrict__ out,
                                                         uint32_t* __restrict__ block_sums, int op_tasks) {
  __shared__ uint32_t warp_sums[8];
  const uint32_t base = blockIdx.x * 1024 + threadIdx.x * 4;
  uint32_t v[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    uint32_t x = base + k < n ? in[base + k] : 0;
    if (op_tasks) x = (x + op_tasks - 1) / op_tasks;  // task count of a bucket population (op_tasks = task length)
    v[k] = x;
  }
  uint32_t tsum = v[0] + v[1] + v[2] + v[3];
  uint32_t incl = tsum;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    uint32_t y = __shfl_up_sync(0xffffffffu, incl, off);
    if (lane >= off) incl += y;
  }
  if (lane == 31) warp_sums[warp] = incl;
  __syncthreads();
  uint32_t wprefix = 0;
  for (int k = 0; k < warp; ++k) wprefix += warp_sums[k];
  uint32_t excl = wprefix + incl - tsum;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    if (base + k < n) out[base + k] = excl;
    excl += v[k];
  }
  if (threadIdx.x == 255) block_sums[blockIdx.x] = wprefix + incl;
}
__global__ void scan_blocks_kernel(uint32_t* block_sums, uint32_t nblocks, uint32_t* total) {
  // single CTA, serial over chunks of 1024 (nblocks is at most a few thousand)
  __shared__ uint32_t sh[1024];
  __shared__ uint32_t carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t base = 0; base < nblocks; base += 1024) {
    const uint32_t i = base + threadIdx.x;
    uint32_t x = i < nblocks ? block_sums[i] : 0;
    sh[threadIdx.x] = x;
    __syncthreads();
    for (int off = 1; off < 1024; off <<= 1) {
      uint32_t y = threadIdx.x >= off ? sh[threadIdx.x - off] : 0;
      __syncthreads();
      sh[threadIdx.x] += y;
      __syncthreads();
    }
    if (i < nblocks) block_sums[i] = carry + sh[threadIdx.x] - x;
    __syncthreads();
    if (threadIdx.x == 0) carry += sh[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t* __restrict__ out, uint32_t n,
                                                       const uint32_t* __restrict__ block_sums,
                                                       const uint32_t* __restrict__ total) {
  const uint32_t base = blockIdx.x * 1024 + threadIdx.x * 4;
  const uint32_t add = block_sums[blockIdx.x];
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (base + k < n) out[base + k] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) out[n] = *total;  // sentinel: out has n+1 entries
}

static int exclusive_scan(Ctx* c, const uint32_t* in, uint32_t n, uint32_t* out, uint32_t* scratch, int op_tasks) {
  const uint32_t nblocks = (n + 1023) / 1024;
  uint32_t* block_sums = scratch;
  uint32_t* total = scratch + nblocks;
  scan_local_kernel<<<nblocks, 256, 0, c->stream>>>(in, n, out, block_sums, op_tasks);
  scan_blocks_kernel<<<1, 1024, 0, c->stream>>>(block_sums, nblocks, total);
  scan_add_kernel<<<nblocks, 256, 0, c->stream>>>(out, n, block_sums, total);
  count_launch(c, 3);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// ---- accumulate --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t upper_bound_u32(const uint32_t* __restrict__ a, uint32_t n, uint32_t key) {
  // largest idx in [0, n) with a[idx] <= key  (a non-decreasing, a[0] == 0)
  uint32_t lo = 0, hi = n;
  while (hi - lo > 1) {
    const uint32_t mid = (lo + hi) >> 1;
    if (a[mid] <= key) lo = mid;
    else hi = mid;
  }
  return lo;
}
__device__ __forceinline__ int job_of_bucket(const MsmPlanDev& plan, uint32_t gb) {
  int j = 0;
  while (j + 1 < plan.J && plan.job[j + 1].bucket_base <= gb) ++j;
  return j;
}
__device__ __forceinline__ void st_xyzz(G1Xyzz* p, const G1Xyzz& v) {
  fe_st(&p->x, v.x);
  fe_st(&p->y, v.y);
  fe_st(&p->zz, v.zz);
  fe_st(&p->zzz, v.zzz);
}
__device__ __forceinline__ G1Xyzz ld_xyzz(const G1Xyzz* p) {
  G1Xyzz v;
  v.x = fe_ld(&p->x);
  v.y = fe_ld(&p->y);
  v.zz = fe_ld(&p->zz);
  v.zzz = fe_ld(&p->zzz);
  return v;
}
__device__ __forceinline__ G1Aff ldg_aff(const G1Aff* p) {
  G1Aff v;
  v.x = fe_ldg(&p->x);
  v.y = fe_ldg(&p->y);
  return v;
}

__global__ void __launch_bounds__(128) msm_accumulate_kernel(MsmPlanDev plan, uint32_t nbuckets,
                                                             const uint32_t* __restrict__ cnt,
                                                             const uint32_t* __restrict__ boff,
                                                             const uint32_t* __restrict__ toff,
                                                             const uint32_t* __restrict__ sorted,
                                                             G1Xyzz* __restrict__ partial) {
  const uint32_t ntasks = toff[nbuckets];
  for (uint32_t task = blockIdx.x * blockDim.x + threadIdx.x; task < ntasks; task += gridDim.x * blockDim.x) {
    // buckets with zero tasks share their offset with the next one: take the LAST bucket whose offset <= task
    uint32_t gb = upper_bound_u32(toff, nbuckets + 1, task);
    const MsmJobDev& jb = plan.job[job_of_bucket(plan, gb)];
    const uint32_t first = boff[gb] + (task - toff[gb]) * plan.task_len;
    const uint32_t end_b = boff[gb] + cnt[gb];
    const uint32_t last = first + plan.task_len < end_b ? first + plan.task_len : end_b;
    G1Xyzz acc = g1_identity();
    for (uint32_t k = first; k < last; ++k) {
      const uint32_t v = sorted[k];
      acc = g1_add_affine(acc, ldg_aff(jb.bases + (v & 0x7fffffffu)), (v >> 31) != 0);
    }
    st_xyzz(partial + task, acc);
  }
}

// per-bucket sum of task partials; heavy buckets are queued for the warp kernel
__global__ void __launch_bounds__(128) msm_bucket_kernel(uint32_t nbuckets, const uint32_t* __restrict__ toff,
                                                         const G1Xyzz* __restrict__ partial,
                                                         G1Xyzz* __restrict__ bucket_sum,
                                                         uint32_t* __restrict__ heavy, uint32_t* heavy_count) {
  const uint32_t gb = blockIdx.x * blockDim.x + threadIdx.x;
  if (gb >= nbuckets) return;
  const uint32_t t0 = toff[gb], nt = toff[gb + 1] - t0;
  if (nt > SEQ_TASKS) {
    heavy[atomicAdd(heavy_count, 1u)] = gb;
    return;
  }
  G1Xyzz acc = g1_identity();
  for (uint32_t k = 0; k < nt; ++k) acc = g1_add(acc, ld_xyzz(partial + t0 + k));
  st_xyzz(bucket_sum + gb, acc);
}

// One shared, NON-inlined copy of the group operations for the latency-bound reduction kernels (few warps, long
// dependent chains): inlined, every call site carries 14 Montgomery products (~35 KB of code) and a single warp
// becomes instruction-fetch bound.
static __device__ __noinline__ G1Xyzz g1_add_ni(G1Xyzz a, G1Xyzz b) {
  if (g1_is_identity(a)) return b;
  if (g1_is_identity(b)) return a;
  const Fq u1 = fq_mul_ni(a.x, b.zz), u2 = fq_mul_ni(b.x, a.zz);
  const Fq s1 = fq_mul_ni(a.y, b.zzz), s2 = fq_mul_ni(b.y, a.zzz);
  const Fq p = u2 - u1, r = s2 - s1;
  if (fe_is_zero<FqP>(p)) {
    if (fe_is_zero<FqP>(r)) return g1_dbl(a);
    return g1_identity();
  }
  const Fq pp = fq_mul_ni(p, p), ppp = fq_mul_ni(p, pp), q = fq_mul_ni(u1, pp);
  G1Xyzz o;
  o.x = fq_mul_ni(r, r) - ppp - fe_dbl<FqP>(q);
  o.y = fq_mul_ni(r, q - o.x) - fq_mul_ni(s1, ppp);
  o.zz = fq_mul_ni(fq_mul_ni(a.zz, b.zz), pp);
  o.zzz = fq_mul_ni(fq_mul_ni(a.zzz, b.zzz), ppp);
  return o;
}
static __device__ __noinline__ G1Xyzz g1_dbl_ni(G1Xyzz p) {
  if (g1_is_identity(p)) return p;
  const Fq u = fe_dbl<FqP>(p.y);
  const Fq v = fq_mul_ni(u, u), w = fq_mul_ni(u, fq_mul_ni(u, u));
  const Fq s = fq_mul_ni(p.x, v), xx = fq_mul_ni(p.x, p.x);
  const Fq m = fe_dbl<FqP>(xx) + xx;
  G1Xyzz r;
  r.x = fq_mul_ni(m, m) - fe_dbl<FqP>(s);
  r.y = fq_mul_ni(m, s - r.x) - fq_mul_ni(w, p.y);
  r.zz = fq_mul_ni(v, p.zz);
  r.zzz = fq_mul_ni(w, p.zzz);
  return r;
}

__device__ __forceinline__ G1Xyzz shfl_down_xyzz(const G1Xyzz& p, int off) {
  G1Xyzz r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.x.v[i] = __shfl_down_sync(0xffffffffu, p.x.v[i], off);
    r.y.v[i] = __shfl_down_sync(0xffffffffu, p.y.v[i], off);
    r.zz.v[i] = __shfl_down_sync(0xffffffffu, p.zz.v[i], off);
    r.zzz.v[i] = __shfl_down_sync(0xffffffffu, p.zzz.v[i], off);
  }
  return r;
}
__device__ __forceinline__ G1Xyzz warp_sum_xyzz(G1Xyzz acc) {
  for (int off = 16; off > 0; off >>= 1) {
    const G1Xyzz o = shfl_down_xyzz(acc, off);
    acc = g1_add_ni(acc, o);
  }
  return acc;  // valid in lane 0
}

// one WARP per heavy bucket: lanes stride over the task partials, then a shuffle reduction
__global__ void __launch_bounds__(128) msm_heavy_kernel(const uint32_t* __restrict__ toff,
                                                        const G1Xyzz* __restrict__ partial,
                                                        G1Xyzz* __restrict__ bucket_sum,
                                                        const uint32_t* __restrict__ heavy,
                                                        const uint32_t* __restrict__ heavy_count) {
  const uint32_t nheavy = *heavy_count, lane = threadIdx.x & 31;
  const uint32_t nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t h = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; h < nheavy; h += nwarps) {
    const uint32_t gb = heavy[h];
    const uint32_t t0 = toff[gb], nt = toff[gb + 1] - t0;
    G1Xyzz acc = g1_identity();
    for (uint32_t k = lane; k < nt; k += 32) acc = g1_add_ni(acc, ld_xyzz(partial + t0 + k));
    acc = warp_sum_xyzz(acc);
    if (lane == 0) st_xyzz(bucket_sum + gb, acc);
  }
}

// Grouped jobs: bucket v - 1 of the job = Σ_{k in group v} B_k over the bucket sums of its source job. One CTA per
// (value, job): the threads stride over the member list, then a shuffle + shared-memory reduction.
__global__ void __launch_bounds__(128) msm_group_kernel(MsmPlanDev plan, G1Xyzz* __restrict__ bucket_sum) {
  __shared__ G1Xyzz s_part[4];
  const MsmJobDev& jb = plan.job[blockIdx.y];
  if (jb.group_src < 0 || blockIdx.x >= jb.B) return;
  const uint32_t v = blockIdx.x;  // weight v + 1
  const uint32_t lo = jb.group_off[v], hi = jb.group_off[v + 1];
  const G1Xyzz* __restrict__ src = bucket_sum + plan.job[jb.group_src].bucket_base;
  G1Xyzz acc = g1_identity();
  for (uint32_t i = lo + threadIdx.x; i < hi; i += blockDim.x) acc = g1_add_ni(acc, ld_xyzz(src + jb.group_perm[i]));
  acc = warp_sum_xyzz(acc);
  if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    acc = s_part[0];
    for (int w = 1; w < 4; ++w) acc = g1_add_ni(acc, s_part[w]);
    st_xyzz(bucket_sum + jb.bucket_base + v, acc);
  }
}

// ---- window reduction ----------------------------------------------------------------------------
// S_w = Σ_k (k+1) B_k over the B buckets of a window, without any scalar multiplication. Write
// k = lo + 8 (l1 + 8 (l2 + 32 l3)); then  k + 1 = (lo + 1) + 8 l1 + 64 l2 + 2048 l3  and
//   S_w = Σ Y + 8 (Σ A + 8 (Σ Bq + 32 C)),   Y = Σ_lo (lo+1) B,  A = Σ l1 R1,  Bq = Σ l2 R2,  C = Σ l3 R3,
// where R1 / R2 / R3 are the plain sums of 8 / 64 / 2048 consecutive buckets. Four launches:
//   l0, l1: one THREAD per 8 entries, running sums (16 dependent additions; work-efficient while entries are many)
//                                                                                     -> R1, Y1 ; R2, A2, Y2
//   tree : one WARP per 32 entries: suffix scan by shuffles (Σ l R_l = Σ_{l>=1} suffix_l, 10 add steps) and plain
//          shuffle sums of the carried streams                                        -> R3, Bq3, A3, Y3
//   final: one warp per window: the last weighted sum, the stream totals and 11 doublings
// All windows of all jobs of the batch go through the same launches (the reference reduces each window with a serial
// running sum, msm.rs:160-181; the group element is the same).
static const int L0 = 8;
struct MsmRedDesc {       // per reduction window (device arrays, nwin + 1 prefix entries where noted)
  const uint32_t* wb;     // first bucket of the window
  const uint32_t* wB;     // buckets in the window
  const uint32_t* off1;   // prefix of level-1 entries (groups of 8 buckets)
  const uint32_t* off2;   // prefix of level-2 entries (32 level-1 entries each)
  const uint32_t* off3;   // prefix of level-3 entries
  uint32_t nwin;
};
__device__ __forceinline__ uint32_t seg_of(const uint32_t* __restrict__ off, uint32_t n, uint32_t key) {
  return upper_bound_u32(off, n + 1, key);  // off[seg] <= key < off[seg + 1] (empty segments never match: sizes >= 1)
}

__global__ void __launch_bounds__(128) msm_red_l0_kernel(MsmRedDesc d, uint32_t n1tot,
                                                         const G1Xyzz* __restrict__ bucket_sum,
                                                         G1Xyzz* __restrict__ R1, G1Xyzz* __restrict__ Y1) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n1tot) return;
  const uint32_t w = seg_of(d.off1, d.nwin, g);
  const uint32_t B = d.wB[w], k0 = (g - d.off1[w]) * L0;
  const uint32_t k1 = k0 + L0 < B ? k0 + L0 : B;
  const G1Xyzz* b = bucket_sum + d.wb[w];
  G1Xyzz run = g1_identity(), wsum = g1_identity();
  for (uint32_t k = k1; k-- > k0;) {
    run = g1_add(run, ld_xyzz(b + k));
    wsum = g1_add(wsum, run);
  }
  st_xyzz(R1 + g, run);
  st_xyzz(Y1 + g, wsum);  // Σ (k - k0 + 1) B_k
}

__device__ __forceinline__ G1Xyzz shfl_xyzz(const G1Xyzz& p, int src) {
  G1Xyzz r;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.x.v[i] = __shfl_sync(0xffffffffu, p.x.v[i], src);
    r.y.v[i] = __shfl_sync(0xffffffffu, p.y.v[i], src);
    r.zz.v[i] = __shfl_sync(0xffffffffu, p.zz.v[i], src);
    r.zzz.v[i] = __shfl_sync(0xffffffffu, p.zzz.v[i], src);
  }
  return r;
}
// lane l holds x_l: returns (in lane 0) plain = Σ_l x_l and weighted = Σ_l l x_l
__device__ __forceinline__ void warp_weighted_sum(G1Xyzz x, G1Xyzz& plain, G1Xyzz& weighted) {
  const int lane = threadIdx.x & 31;
  // inclusive suffix scan: x_l <- Σ_{j >= l} x_j
  for (int off = 1; off < 32; off <<= 1) {
    const G1Xyzz o = shfl_down_xyzz(x, off);
    if (lane + off < 32) x = g1_add_ni(x, o);
  }
  plain = x;  // lane 0: the total
  G1Xyzz s = lane >= 1 ? x : g1_identity();
  weighted = warp_sum_xyzz(s);
}

// level 1 -> level 2: TWO threads per 8 consecutive level-1 entries (work-efficient while there are still many entries):
// the even thread runs the dependent chain R2 = Σ R1, A2 = Σ l R1_l (l = 0..7), the odd thread the independent plain sum
// Y2 = Σ Y1 — 16 instead of 24 group operations on the critical path of this latency-bound launch
__global__ void __launch_bounds__(128) msm_red_l1_kernel(MsmRedDesc d, uint32_t n2tot, const G1Xyzz* __restrict__ in,
                                                         uint32_t n1tot, G1Xyzz* __restrict__ out) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t g = tid >> 1, role = tid & 1;
  if (g >= n2tot) return;
  const uint32_t w = seg_of(d.off2, d.nwin, g);
  const uint32_t i0 = d.off1[w] + (g - d.off2[w]) * 8, iend = d.off1[w + 1];
  const uint32_t i1 = i0 + 8 < iend ? i0 + 8 : iend;
  if (role) {
    G1Xyzz ysum = g1_identity();
    for (uint32_t i = i1; i-- > i0;) ysum = g1_add(ysum, ld_xyzz(in + n1tot + i));
    st_xyzz(out + (size_t)2 * n2tot + g, ysum);
    return;
  }
  G1Xyzz run = g1_identity(), wsum = g1_identity();
  for (uint32_t i = i1; i-- > i0;) {
    run = g1_add(run, ld_xyzz(in + i));
    if (i > i0) wsum = g1_add(wsum, run);
  }
  st_xyzz(out + g, run);
  st_xyzz(out + n2tot + g, wsum);
}

// 1 + NP warps per 32 consecutive entries of a window at level LV (1 or 2). Streams in: R (weighted), P[0..NP) plain.
// Streams out: R', Wt' (= Σ l R_l) by the first warp (suffix scan + sum: 10 add steps) and the NP plain sums by one warp
// each (5 add steps) — the streams are independent, so they do not queue up behind each other in one warp.
template <int NP>
__global__ void __launch_bounds__(128) msm_red_tree_kernel(MsmRedDesc d, int lv, uint32_t nout_tot,
                                                           const G1Xyzz* __restrict__ in, uint32_t nin_tot,
                                                           G1Xyzz* __restrict__ out) {
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t wid = gw / (1 + NP), role = gw % (1 + NP);
  if (wid >= nout_tot) return;
  const uint32_t* offo = lv == 1 ? d.off2 : d.off3;
  const uint32_t* offi = lv == 1 ? d.off1 : d.off2;
  const uint32_t w = seg_of(offo, d.nwin, wid);
  const uint32_t i0 = offi[w] + (wid - offo[w]) * 32, iend = offi[w + 1];
  const bool valid = i0 + lane < iend;
  if (role == 0) {
    G1Xyzz plain, wt;
    warp_weighted_sum(valid ? ld_xyzz(in + i0 + lane) : g1_identity(), plain, wt);
    if (lane == 0) {
      st_xyzz(out + wid, plain);
      st_xyzz(out + nout_tot + wid, wt);
    }
  } else {
    const uint32_t sidx = role - 1;
    const G1Xyzz v = warp_sum_xyzz(valid ? ld_xyzz(in + (size_t)(1 + sidx) * nin_tot + i0 + lane) : g1_identity());
    if (lane == 0) st_xyzz(out + (size_t)(2 + sidx) * nout_tot + wid, v);
  }
}

// One CTA of four warps per window: level-3 streams R3, Bq3, A3, Y3 (<= 32 entries each) -> S_w. Warp 0 builds C (weighted
// sum of R3), warps 1-3 the plain sums of the carried streams in parallel; lane 0 of warp 0 then runs the Horner tail
// (11 doublings, 3 additions).
__global__ void __launch_bounds__(128) msm_red_final_kernel(MsmRedDesc d, const G1Xyzz* __restrict__ in, uint32_t n3tot,
                                                            G1Xyzz* __restrict__ window_sum) {
  __shared__ G1Xyzz s_sum[3];
  const uint32_t w = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (w >= d.nwin) return;
  const uint32_t i0 = d.off3[w], n = d.off3[w + 1] - i0;
  const bool valid = lane < n;
  G1Xyzz r3, c;
  if (warp == 0) {
    warp_weighted_sum(valid ? ld_xyzz(in + i0 + lane) : g1_identity(), r3, c);
  } else {
    const G1Xyzz v = warp_sum_xyzz(valid ? ld_xyzz(in + (size_t)warp * n3tot + i0 + lane) : g1_identity());
    if (lane == 0) s_sum[warp - 1] = v;  // 0: Bq, 1: A, 2: Y
  }
  __syncthreads();
  if (threadIdx.x == 0) {  // S = Y + 8 (A + 8 (Bq + 32 C))
    G1Xyzz acc = c;
    for (int k = 0; k < 5; ++k) acc = g1_dbl_ni(acc);
    acc = g1_add_ni(acc, s_sum[0]);
    for (int k = 0; k < 3; ++k) acc = g1_dbl_ni(acc);
    acc = g1_add_ni(acc, s_sum[1]);
    for (int k = 0; k < 3; ++k) acc = g1_dbl_ni(acc);
    acc = g1_add_ni(acc, s_sum[2]);
    st_xyzz(window_sum + w, acc);
  }
}

// Derived results: out[J + i] = Σ_t 2^(shift t) * result(src[t]) — commitments that are linear combinations of other
// commitments of the same batch (Lasso: a = Σ_t 2^(w t) E_t) cost a few doublings instead of an MSM.
struct MsmDeriveDev {
  int n;
  struct {
    int nsrc, shift, src[8];
  } d[4];
};

// warp per job (or derived result): lane w folds the windows w, w + 32, ... (Horner), shifts by w * c doublings,
// the lanes are summed by shuffles and lane 0 normalises. Warps >= J handle the derived results.
__global__ void __launch_bounds__(32) msm_finish_kernel(MsmPlanDev plan, MsmDeriveDev dv,
                                                        const G1Xyzz* __restrict__ window_sum, G1Aff* __restrict__ out) {
  const int j = blockIdx.x, lane = threadIdx.x;
  G1Xyzz acc = g1_identity();
  if (j < plan.J) {
    const MsmJobDev& jb = plan.job[j];
    int top = -1;
    for (int w = lane; w < jb.Wred; w += 32) top = w;
    for (int w = top; w >= 0; w -= 32) {  // Σ_i 2^(32 c i) S_{lane + 32 i}
      if (w != top)
        for (int k = 0; k < 32 * jb.c; ++k) acc = g1_dbl_ni(acc);
      acc = g1_add_ni(acc, ld_xyzz(window_sum + jb.win_base + w));
    }
    if (jb.Wred > 1)
      for (int k = 0; k < lane * jb.c; ++k) acc = g1_dbl_ni(acc);
  } else {
    const auto& d = dv.d[j - plan.J];
    if (lane < d.nsrc) {  // sources are single-window jobs or short Horner chains: recomputed by this lane
      const MsmJobDev& jb = plan.job[d.src[lane]];
      for (int w = jb.Wred - 1; w >= 0; --w) {
        for (int k = 0; k < jb.c; ++k) acc = g1_dbl_ni(acc);
        acc = g1_add_ni(acc, ld_xyzz(window_sum + jb.win_base + w));
      }
      for (int k = 0; k < lane * d.shift; ++k) acc = g1_dbl_ni(acc);
    }
  }
  acc = warp_sum_xyzz(acc);
  if (lane == 0) {
    const G1Aff a = g1_to_affine(acc);
    fe_st(&out[j].x, a.x);
    fe_st(&out[j].y, a.y);
  }
}

static int ilog2_floor(uint64_t n) {
  int k = 0;
  while ((n >> (k + 1)) != 0) ++k;
  return k;
}

int msm_batch(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out, const MsmDerive* derive, int nderive) {
  if (J < 1 || J > MSM_MAX_JOBS || nderive < 0 || nderive > 4) return B200_ERR_ARG;
  NvtxRange nvtx("variable_base_msm-%llu x%d", (unsigned long long)jobs[0].n, J);  // msm.rs:92 (batched here)
  cudaStream_t s = c->stream;
  MsmPlanDev plan;
  plan.J = J;
  uint64_t pairs = 0;
  uint32_t nbuckets = 0, nwin = 0, max_n = 0, max_n_hist = 0;
  std::vector<uint32_t> desc;  // wb | wB | off1 | off2 | off3, each nwin (+1 for the prefixes)
  std::vector<uint32_t> wb, wB;
  bool any_grouped = false;
  std::vector<char> is_group_src(J, 0);
  for (int j = 0; j < J; ++j) {
    if (jobs[j].group_src < 0) continue;
    const MsmJob& g = jobs[j];
    // the source comes earlier in the batch, carries 16-bit u32 addresses and is not grouped / precomputed itself
    if (g.group_src >= j || g.ngroups < 1 || g.ngroups > 4095 || !g.group_perm || !g.group_off) return B200_ERR_ARG;
    const MsmJob& src = jobs[g.group_src];
    if (src.group_src >= 0 || src.kind != MSM_U32 || src.bits > 16 || src.ext) return B200_ERR_ARG;
    is_group_src[g.group_src] = 1;
    any_grouped = true;
  }
  for (int j = 0; j < J; ++j) {
    const MsmJob& in = jobs[j];
    MsmJobDev& jb = plan.job[j];
    jb.group_src = in.group_src;
    jb.group_perm = in.group_perm;
    jb.group_off = in.group_off;
    if (in.group_src >= 0) {  // no points of its own: one reduction window of ngroups buckets, filled by msm_group_kernel
      jb.scalars = nullptr;
      jb.bases = nullptr;
      jb.n = 0;
      jb.kind = MSM_U32;
      jb.map_p = jb.map_g = 0;
      jb.map_rank = 0;
      jb.precomp = 0;
      jb.ext_stride = 0;
      jb.c = 1;
      jb.W = 0;
      jb.Wred = 1;
      jb.B = (uint32_t)in.ngroups;
      jb.hist = 0;
      jb.bucket_base = nbuckets;
      jb.pair_base = pairs;
      jb.win_base = nwin;
      wb.push_back(nbuckets);
      wB.push_back(jb.B);
      nbuckets += jb.B;
      nwin += 1;
      continue;
    }
    if (in.n == 0 || in.n > (1u << 30) || in.bits < 1 || in.bits > 256) return B200_ERR_ARG;
    jb.scalars = in.scalars;
    jb.bases = in.bases;
    jb.n = (uint32_t)in.n;
    jb.kind = in.kind;
    jb.map_p = in.map_p;
    jb.map_g = in.map_g;
    jb.map_rank = (uint32_t)in.map_rank;
    const int need = in.bits + 1;  // signed digits may carry one bit past the top
    jb.precomp = (in.ext != nullptr && in.bits > 2 * EXT_C + 2) ? 1 : 0;
    jb.ext_stride = (uint32_t)(in.ext_stride ? in.ext_stride : in.n);
    if (jb.precomp) {
      jb.bases = in.ext;
      jb.c = EXT_C;
      jb.W = (need + EXT_C - 1) / EXT_C;
      if (jb.W > EXT_WINDOWS) return B200_ERR_ARG;
      jb.Wred = 1;
    } else {
      int cmax = ilog2_floor(in.n) - 3;
      if (cmax < 3) cmax = 3;
      if (cmax > 17) cmax = 17;
      if (is_group_src[j]) cmax = 17;  // ONE window whose bucket k is address k + 1, whatever the number of points
      jb.W = (need + cmax - 1) / cmax;
      jb.c = (need + jb.W - 1) / jb.W;
      if (jb.c < 2) jb.c = 2;
      jb.Wred = jb.W;
    }
    if (jb.W > MSM_MAX_WINDOWS) return B200_ERR_ARG;
    jb.B = 1u << (jb.c - 1);
    jb.hist = (!jb.precomp && (in.kind == MSM_U32 || in.kind == MSM_U64) && jb.B <= HIST_MAX_B && in.n >= 4096) ? 1 : 0;
    if (jb.hist && jb.n > max_n_hist) max_n_hist = jb.n;
    jb.bucket_base = nbuckets;
    jb.pair_base = pairs;
    jb.win_base = nwin;
    for (int w = 0; w < jb.Wred; ++w) {
      wb.push_back(nbuckets + (uint32_t)w * jb.B);
      wB.push_back(jb.B);
    }
    nbuckets += jb.B * jb.Wred;
    pairs += (uint64_t)jb.n * jb.W;
    nwin += jb.Wred;
    if (jb.n > max_n) max_n = jb.n;
  }
  if (pairs >= (1ull << 31)) return B200_ERR_ARG;
  MsmDeriveDev dv;
  dv.n = nderive;
  for (int i = 0; i < nderive; ++i) {
    if (derive[i].nsrc < 1 || derive[i].nsrc > 8 || derive[i].shift < 0 || derive[i].shift > 32) return B200_ERR_ARG;
    dv.d[i].nsrc = derive[i].nsrc;
    dv.d[i].shift = derive[i].shift;
    for (int t = 0; t < derive[i].nsrc; ++t) {
      if (derive[i].src[t] < 0 || derive[i].src[t] >= J) return B200_ERR_ARG;
      dv.d[i].src[t] = derive[i].src[t];
    }
  }
  // task length: about two tasks per resident thread of the accumulate grid (148 x 16 x 128), within [64, 512]
  uint32_t task_len = TASK_LEN;
  while (task_len < 512 && (uint64_t)task_len * 2 * NUM_SMS * 16 * 128 < pairs) task_len <<= 1;
  plan.task_len = task_len;
  const uint32_t max_tasks = nbuckets + (uint32_t)(pairs / task_len) + 1;
  const uint32_t scan_blocks = (nbuckets + 1023) / 1024;
  // reduction levels (see msm_red_*): entries per window at level 1 / 2 / 3
  std::vector<uint32_t> off1(nwin + 1, 0), off2(nwin + 1, 0), off3(nwin + 1, 0);
  for (uint32_t w = 0; w < nwin; ++w) {
    const uint32_t n1 = (wB[w] + L0 - 1) / L0, n2 = (n1 + 7) / 8, n3 = (n2 + 31) / 32;
    if (n3 > 32) return B200_ERR_ARG;
    off1[w + 1] = off1[w] + n1;
    off2[w + 1] = off2[w] + n2;
    off3[w + 1] = off3[w] + n3;
  }
  const uint32_t n1tot = off1[nwin], n2tot = off2[nwin], n3tot = off3[nwin];
  desc.insert(desc.end(), wb.begin(), wb.end());
  desc.insert(desc.end(), wB.begin(), wB.end());
  desc.insert(desc.end(), off1.begin(), off1.end());
  desc.insert(desc.end(), off2.begin(), off2.end());
  desc.insert(desc.end(), off3.begin(), off3.end());

  DevScope mem(s);
  uint32_t *cnt, *boff, *toff, *ranks, *sorted, *scratch, *heavy, *heavy_count, *d_desc;
  G1Xyzz *partial, *bucket_sum, *lvl1, *lvl2, *lvl3, *window_sum;
  CUDA_TRY(mem.alloc(&cnt, (size_t)nbuckets * 4));
  CUDA_TRY(mem.alloc(&boff, ((size_t)nbuckets + 1) * 4));
  CUDA_TRY(mem.alloc(&toff, ((size_t)nbuckets + 1) * 4));
  CUDA_TRY(mem.alloc(&ranks, (size_t)pairs * 8));
  CUDA_TRY(mem.alloc(&sorted, (size_t)pairs * 4 + 4));
  CUDA_TRY(mem.alloc(&scratch, ((size_t)scan_blocks + 8) * 4));
  CUDA_TRY(mem.alloc(&heavy, ((size_t)nbuckets + 1) * 4));
  CUDA_TRY(mem.alloc(&partial, (size_t)max_tasks * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&bucket_sum, (size_t)nbuckets * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&lvl1, (size_t)2 * n1tot * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&lvl2, (size_t)3 * n2tot * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&lvl3, (size_t)4 * n3tot * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&window_sum, (size_t)nwin * sizeof(G1Xyzz)));
  CUDA_TRY(mem.alloc(&d_desc, desc.size() * 4));
  CUDA_TRY(cudaMemcpyAsync(d_desc, desc.data(), desc.size() * 4, cudaMemcpyHostToDevice, s));
  MsmRedDesc rd;
  rd.wb = d_desc;
  rd.wB = d_desc + nwin;
  rd.off1 = d_desc + 2 * nwin;
  rd.off2 = rd.off1 + nwin + 1;
  rd.off3 = rd.off2 + nwin + 1;
  rd.nwin = nwin;
  heavy_count = scratch + scan_blocks + 4;
  CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)nbuckets * 4, s));
  CUDA_TRY(cudaMemsetAsync(heavy_count, 0, 4, s));

  dim3 grid((max_n + 255) / 256, J);
  int pi = prof_begin(c, PH_MSM_SORT);
  msm_digits_kernel<0><<<grid, 256, 0, s>>>(plan, cnt, nullptr, ranks, nullptr);
  count_launch(c);
  if (max_n_hist) {
    msm_digits_hist_kernel<<<dim3((max_n_hist + 256 * HIST_K - 1) / (256 * HIST_K), J), 256, 0, s>>>(plan, cnt, ranks);
    count_launch(c);
  }
  int rc = exclusive_scan(c, cnt, nbuckets, boff, scratch, 0);
  if (rc) return rc;
  rc = exclusive_scan(c, cnt, nbuckets, toff, scratch, (int)task_len);
  if (rc) return rc;
  msm_digits_kernel<1><<<grid, 256, 0, s>>>(plan, nullptr, boff, ranks, sorted);
  prof_end(c, pi);
  pi = prof_begin(c, PH_MSM_ACC);
  // persistent-style grids: a multiple of the SM count, tasks are grid-strided
  msm_accumulate_kernel<<<NUM_SMS * 16, 128, 0, s>>>(plan, nbuckets, cnt, boff, toff, sorted, partial);
  prof_end(c, pi);
  pi = prof_begin(c, PH_MSM_REDUCE);
  msm_bucket_kernel<<<(nbuckets + 127) / 128, 128, 0, s>>>(nbuckets, toff, partial, bucket_sum, heavy, heavy_count);
  msm_heavy_kernel<<<NUM_SMS * 8, 128, 0, s>>>(toff, partial, bucket_sum, heavy, heavy_count);
  if (any_grouped) {
    uint32_t maxg = 0;
    for (int j = 0; j < J; ++j)
      if (jobs[j].group_src >= 0 && (uint32_t)jobs[j].ngroups > maxg) maxg = (uint32_t)jobs[j].ngroups;
    msm_group_kernel<<<dim3(maxg, J), 128, 0, s>>>(plan, bucket_sum);
    count_launch(c);
  }
  msm_red_l0_kernel<<<(n1tot + 127) / 128, 128, 0, s>>>(rd, n1tot, bucket_sum, lvl1, lvl1 + n1tot);
  msm_red_l1_kernel<<<(2 * n2tot + 127) / 128, 128, 0, s>>>(rd, n2tot, lvl1, n1tot, lvl2);
  msm_red_tree_kernel<2><<<(3 * n3tot + 3) / 4, 128, 0, s>>>(rd, 2, n3tot, lvl2, n2tot, lvl3);
  msm_red_final_kernel<<<nwin, 128, 0, s>>>(rd, lvl3, n3tot, window_sum);
  msm_finish_kernel<<<J + nderive, 32, 0, s>>>(plan, dv, window_sum, d_out);
  prof_end(c, pi);
  count_launch(c, 10);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

uint64_t msm_group_min_points() {
  static const uint64_t v = [] {
    const char* e = getenv("B200_MSM_GROUP_MIN_POINTS");  // tests lower it to exercise the grouped path at small sizes
    return e ? (uint64_t)atoll(e) : (uint64_t)1 << 18;
  }();
  return v;
}

bool lasso_group_lists(const uint32_t* values, std::vector<uint32_t>* perm, std::vector<uint32_t>* off) {
  const uint32_t S = 1u << 16;
  if (values[0] != 0) return false;  // address 0 has no bucket (digit 0 is skipped)
  uint32_t mx = 0;
  for (uint32_t d = 0; d < S; ++d) mx = values[d] > mx ? values[d] : mx;
  if (mx < 1 || mx > 4095) return false;
  std::vector<uint32_t> cnt(mx + 2, 0);
  for (uint32_t d = 1; d < S; ++d)
    if (values[d]) ++cnt[values[d] + 1];
  off->assign(mx + 1, 0);  // off[v - 1] = start of value v, off[mx] = end
  for (uint32_t v = 1; v <= mx; ++v) (*off)[v] = (*off)[v - 1] + cnt[v + 1];
  perm->assign((*off)[mx], 0);
  std::vector<uint32_t> at(off->begin(), off->end());
  for (uint32_t d = 1; d < S; ++d)
    if (values[d]) (*perm)[at[values[d] - 1]++] = d - 1;  // bucket k holds address k + 1
  return true;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_msm() {
  B200_PRELOAD(msm_digits_hist_kernel);
  B200_PRELOAD(msm_digits_kernel<0>);
  B200_PRELOAD(msm_digits_kernel<1>);
  B200_PRELOAD(scan_local_kernel);
  B200_PRELOAD(scan_blocks_kernel);
  B200_PRELOAD(scan_add_kernel);
  B200_PRELOAD(msm_accumulate_kernel);
  B200_PRELOAD(msm_bucket_kernel);
  B200_PRELOAD(msm_heavy_kernel);
  B200_PRELOAD(msm_group_kernel);
  B200_PRELOAD(msm_red_l0_kernel);
  B200_PRELOAD(msm_red_l1_kernel);
  B200_PRELOAD(msm_red_tree_kernel<2>);
  B200_PRELOAD(msm_red_final_kernel);
  B200_PRELOAD(msm_finish_kernel);
}

}  // namespace b200

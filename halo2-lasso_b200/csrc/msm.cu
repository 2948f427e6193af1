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
//   5. buckets  : per-bucket sum of its task partials (warp-cooperative for heavy buckets)
//   6. windows  : running-sum reduction Σ (k+1) B_k in groups, then a block-wide point sum
//   7. finish   : Horner over windows (c doublings each) and projective -> affine
// Scalars with few significant bits (dims, counters, subtable values) only populate the windows they
// need, which the reference cannot exploit. The group element is unique, so any schedule yields the
// reference's commitment bytes.
#include "internal.h"

namespace b200 {

static const int TASK_LEN = 64;       // minimum task length; a batch uses a power of two in [64, 1024] so that an
                                      // average bucket splits into ~8 tasks (keeps the per-bucket reduction short at 2^22+)
static const int SEQ_TASKS = 8;     // buckets with more task partials than this go to the warp kernel
static const int GROUP = 16;       // buckets per thread in the window reduction
static const int MSM_MAX_JOBS = 64;
static const int MSM_MAX_WINDOWS = 128;

struct MsmJobDev {
  const void* scalars;
  const G1Aff* bases;
  uint32_t n;
  int kind;           // MsmScalarKind
  int c, W;           // window bits, number of windows
  uint32_t B;         // buckets per window = 2^(c-1)
  int precomp;        // 1: bases = precomputed window multiples, all windows share one bucket set
  uint32_t ext_stride;  // precomp: entries per window of the extended table
  int Wred;           // windows that need a bucket reduction (1 when precomp, else W)
  uint32_t bucket_base;
  uint64_t pair_base;
  uint32_t group_base;  // first reduction group of this job
  uint32_t win_base;    // first window slot of this job
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
      sorted[boff[gb] + (code & 0x7fffffffu)] = (jb.precomp ? i + (uint32_t)w * jb.ext_stride : i) | (code & 0x80000000u);
    }
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
    acc = g1_add(acc, o);
  }
  return acc;  // valid in lane 0
}

// one CTA per heavy bucket: threads stride over the task partials, then warp + shared-memory reduction
__global__ void __launch_bounds__(128) msm_heavy_kernel(const uint32_t* __restrict__ toff,
                                                        const G1Xyzz* __restrict__ partial,
                                                        G1Xyzz* __restrict__ bucket_sum,
                                                        const uint32_t* __restrict__ heavy,
                                                        const uint32_t* __restrict__ heavy_count) {
  __shared__ G1Xyzz sh[4];
  const uint32_t nheavy = *heavy_count;
  for (uint32_t h = blockIdx.x; h < nheavy; h += gridDim.x) {
    const uint32_t gb = heavy[h];
    const uint32_t t0 = toff[gb], nt = toff[gb + 1] - t0;
    G1Xyzz acc = g1_identity();
    for (uint32_t k = threadIdx.x; k < nt; k += blockDim.x) acc = g1_add(acc, ld_xyzz(partial + t0 + k));
    acc = warp_sum_xyzz(acc);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < 4; ++k) acc = g1_add(acc, sh[k]);
      st_xyzz(bucket_sum + gb, acc);
    }
    __syncthreads();
  }
}

// ---- window reduction ----------------------------------------------------------------------------
// thread per group of GROUP buckets: C_g = Σ_{k in group} (k+1) B_k   (k = bucket index inside window)
__global__ void __launch_bounds__(128) msm_group_kernel(MsmPlanDev plan, uint32_t ngroups,
                                                        const G1Xyzz* __restrict__ bucket_sum,
                                                        G1Xyzz* __restrict__ group_sum) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= ngroups) return;
  int j = 0;
  while (j + 1 < plan.J && plan.job[j + 1].group_base <= g) ++j;
  const MsmJobDev& jb = plan.job[j];
  const uint32_t gpw = (jb.B + GROUP - 1) / GROUP;  // groups per window
  const uint32_t local = g - jb.group_base;
  const uint32_t w = local / gpw, gi = local % gpw;
  const uint32_t k0 = gi * GROUP;
  const uint32_t k1 = k0 + GROUP < jb.B ? k0 + GROUP : jb.B;
  const G1Xyzz* b = bucket_sum + jb.bucket_base + w * jb.B;
  G1Xyzz run = g1_identity(), wsum = g1_identity();
  for (uint32_t k = k1; k-- > k0;) {
    run = g1_add(run, ld_xyzz(b + k));
    wsum = g1_add(wsum, run);
  }
  // wsum = Σ (k - k0 + 1) B_k ; add k0 * Σ B_k
  if (k0) wsum = g1_add(wsum, g1_mul_small(run, k0));
  st_xyzz(group_sum + g, wsum);
}

// CTA per (job, window): sum its groups
__global__ void __launch_bounds__(128) msm_window_kernel(MsmPlanDev plan, const G1Xyzz* __restrict__ group_sum,
                                                         G1Xyzz* __restrict__ window_sum) {
  __shared__ G1Xyzz sh[4];
  const uint32_t slot = blockIdx.x;
  int j = 0;
  while (j + 1 < plan.J && plan.job[j + 1].win_base <= slot) ++j;
  const MsmJobDev& jb = plan.job[j];
  const uint32_t w = slot - jb.win_base;
  const uint32_t gpw = (jb.B + GROUP - 1) / GROUP;
  const G1Xyzz* g = group_sum + jb.group_base + w * gpw;
  G1Xyzz acc = g1_identity();
  for (uint32_t k = threadIdx.x; k < gpw; k += blockDim.x) acc = g1_add(acc, ld_xyzz(g + k));
  acc = warp_sum_xyzz(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 4; ++k) acc = g1_add(acc, sh[k]);
    st_xyzz(window_sum + slot, acc);
  }
}

// thread per job: Horner over windows, then affine
__global__ void msm_finish_kernel(MsmPlanDev plan, const G1Xyzz* __restrict__ window_sum, G1Aff* __restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= plan.J) return;
  const MsmJobDev& jb = plan.job[j];
  G1Xyzz acc = g1_identity();
  for (int w = jb.Wred - 1; w >= 0; --w) {
    for (int k = 0; k < jb.c; ++k) acc = g1_dbl(acc);
    acc = g1_add(acc, ld_xyzz(window_sum + jb.win_base + w));
  }
  const G1Aff a = g1_to_affine(acc);
  fe_st(&out[j].x, a.x);
  fe_st(&out[j].y, a.y);
}

static int ilog2_floor(uint64_t n) {
  int k = 0;
  while ((n >> (k + 1)) != 0) ++k;
  return k;
}

int msm_batch(Ctx* c, const MsmJob* jobs, int J, G1Aff* d_out) {
  if (J < 1 || J > MSM_MAX_JOBS) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  MsmPlanDev plan;
  plan.J = J;
  uint64_t pairs = 0;
  uint32_t nbuckets = 0, ngroups = 0, nwin = 0, max_n = 0;
  for (int j = 0; j < J; ++j) {
    const MsmJob& in = jobs[j];
    if (in.n == 0 || in.n > (1u << 30) || in.bits < 1 || in.bits > 256) return B200_ERR_ARG;
    MsmJobDev& jb = plan.job[j];
    jb.scalars = in.scalars;
    jb.bases = in.bases;
    jb.n = (uint32_t)in.n;
    jb.kind = in.kind;
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
      jb.W = (need + cmax - 1) / cmax;
      jb.c = (need + jb.W - 1) / jb.W;
      if (jb.c < 2) jb.c = 2;
      jb.Wred = jb.W;
    }
    if (jb.W > MSM_MAX_WINDOWS) return B200_ERR_ARG;
    jb.B = 1u << (jb.c - 1);
    jb.bucket_base = nbuckets;
    jb.pair_base = pairs;
    jb.group_base = ngroups;
    jb.win_base = nwin;
    nbuckets += jb.B * jb.Wred;
    pairs += (uint64_t)jb.n * jb.W;
    ngroups += ((jb.B + GROUP - 1) / GROUP) * jb.Wred;
    nwin += jb.Wred;
    if (jb.n > max_n) max_n = jb.n;
  }
  if (pairs >= (1ull << 31)) return B200_ERR_ARG;
  uint32_t task_len = TASK_LEN;
  while (task_len < 1024 && (uint64_t)task_len * 8 * nbuckets < pairs) task_len <<= 1;
  plan.task_len = task_len;
  const uint32_t max_tasks = nbuckets + (uint32_t)(pairs / task_len) + 1;
  const uint32_t scan_blocks = (nbuckets + 1023) / 1024;

  uint32_t *cnt, *boff, *toff, *ranks, *sorted, *scratch, *heavy, *heavy_count;
  G1Xyzz *partial, *bucket_sum, *group_sum, *window_sum;
  CUDA_TRY(cudaMallocAsync(&cnt, (size_t)nbuckets * 4, s));
  CUDA_TRY(cudaMallocAsync(&boff, ((size_t)nbuckets + 1) * 4, s));
  CUDA_TRY(cudaMallocAsync(&toff, ((size_t)nbuckets + 1) * 4, s));
  CUDA_TRY(cudaMallocAsync(&ranks, (size_t)pairs * 8, s));
  CUDA_TRY(cudaMallocAsync(&sorted, (size_t)pairs * 4 + 4, s));
  CUDA_TRY(cudaMallocAsync(&scratch, ((size_t)scan_blocks + 8) * 4, s));
  CUDA_TRY(cudaMallocAsync(&heavy, ((size_t)nbuckets + 1) * 4, s));
  CUDA_TRY(cudaMallocAsync(&partial, (size_t)max_tasks * sizeof(G1Xyzz), s));
  CUDA_TRY(cudaMallocAsync(&bucket_sum, (size_t)nbuckets * sizeof(G1Xyzz), s));
  CUDA_TRY(cudaMallocAsync(&group_sum, (size_t)ngroups * sizeof(G1Xyzz), s));
  CUDA_TRY(cudaMallocAsync(&window_sum, (size_t)nwin * sizeof(G1Xyzz), s));
  heavy_count = scratch + scan_blocks + 4;
  CUDA_TRY(cudaMemsetAsync(cnt, 0, (size_t)nbuckets * 4, s));
  CUDA_TRY(cudaMemsetAsync(heavy_count, 0, 4, s));

  dim3 grid((max_n + 255) / 256, J);
  int pi = prof_begin(c, PH_MSM_SORT);
  msm_digits_kernel<0><<<grid, 256, 0, s>>>(plan, cnt, nullptr, ranks, nullptr);
  count_launch(c);
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
  msm_heavy_kernel<<<NUM_SMS * 4, 128, 0, s>>>(toff, partial, bucket_sum, heavy, heavy_count);
  msm_group_kernel<<<(ngroups + 127) / 128, 128, 0, s>>>(plan, ngroups, bucket_sum, group_sum);
  msm_window_kernel<<<nwin, 128, 0, s>>>(plan, group_sum, window_sum);
  msm_finish_kernel<<<(J + 31) / 32, 32, 0, s>>>(plan, window_sum, d_out);
  prof_end(c, pi);
  count_launch(c, 7);
  CUDA_TRY(cudaGetLastError());
  for (void* p : {(void*)cnt, (void*)boff, (void*)toff, (void*)ranks, (void*)sorted, (void*)scratch, (void*)heavy,
                  (void*)partial, (void*)bucket_sum, (void*)group_sum, (void*)window_sum})
    CUDA_TRY(cudaFreeAsync(p, s));
  return B200_OK;
}

}  // namespace b200

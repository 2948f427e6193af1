// Generic EvaluationsProver round kernel (pb/piop/sum_check/classic/eval.rs:92-131, 210-323) for an
// arbitrary `Expression`: the role of ExpressionRegistry's straight-line `Calculation` program
// (pb/util/expression/evaluator.rs:22-228) is played by a bytecode program produced on the host
// (halo2-lasso_b200/expression.py::compile_expression) and interpreted per hypercube pair.
//
// Every leaf of the expression is a dense device table (polynomial queries — rotated ones gathered
// through the BooleanHypercube LFSR map, pb/util/arithmetic/bh.rs:105-141 —, eq_xy tables, the identity
// polynomial, one-hot Lagrange tables), so one code path binds and evaluates all of them: per pair and
// table (eval, step) = (t[2b+1], t[2b+1]-t[2b]), x -> x+1 adds the step, the program runs on the slot
// file [tables | constants | temporaries] and its last slot is accumulated into p(x). The bind of the
// previous challenge is fused exactly as in sumcheck.cu. Slots live in local memory (L1): the program for
// the vanilla-plonk zero check has 13+ tables at degree 5, far beyond the register file.
#include "internal.h"

namespace b200 {

static const int GEN_MAX_TABLES = 40;
static const int GEN_MAX_SLOTS = 200;
static const int GEN_MAX_DEG = 6;

struct GenArgs {
  const Fr* in[GEN_MAX_TABLES];
  Fr* out[GEN_MAX_TABLES];
  const Fr* consts;
  const int4* ops;  // (opcode, dst, a, b)
  int K, C, nops, D;
  ScState* st;
  Fr* partial;
  Transcript* tr;
  const BaryTable* bary;
  Fr* challenges_out;
  uint32_t pairs;
  int round;
};

// Last CTA of a round: total the per-CTA partials, derive p(0), Fiat-Shamir, fold the claim (same scheme as
// sc_eval_round_kernel). Called by every thread of that CTA; smem holds blockDim.x / 32 elements.
__device__ __forceinline__ void gen_finalize(const GenArgs& a, Fr* smem, Fr* s_tot) {
  const int D = a.D;
  for (int x = 0; x < D; ++x) {
    Fr t[1] = {fe_zero<FrP>()};
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) t[0] = t[0] + fr_ld_cg(a.partial + (size_t)i * D + x);
    block_reduce_fr<1>(t, smem);
    if (threadIdx.x == 0) s_tot[x] = t[0];
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // warp 0: lane i owns p(i), i <= D (same scheme as sc_eval_round_kernel)
    const int lane = threadIdx.x;
    __shared__ Transcript sh_tr;
    trw_copy(&sh_tr, a.tr);
    Fr mine = fe_zero<FrP>();
    if (lane >= 1 && lane <= D) mine = s_tot[lane - 1];
    if (lane == 0) mine = fe_ld(&a.st->claim) - s_tot[0];  // p(0) = sum - p(1)
    const Fr canon = fr_canon_ni(mine);
    for (int x = 0; x <= D; ++x) trw_write_canon_from_lane(&sh_tr, canon, x, true);
    const Fr ch = trw_squeeze(&sh_tr);
    const Fr one = fe_one<FrP>();
    Fr num = lane <= D ? a.bary->w[D][lane <= D ? lane : 0] : fe_zero<FrP>();
    Fr jf = fe_zero<FrP>();
    for (int j = 0; j <= D; ++j) {
      const Fr f = (j == lane) ? one : ch - jf;
      num = fr_mul_ni(num, f);
      jf = jf + one;
    }
    Fr term = fr_mul_ni(num, mine);
    if (lane > D) term = fe_zero<FrP>();
#pragma unroll
    for (int off = 1; off < 8; off <<= 1) {
      Fr o;
#pragma unroll
      for (int i = 0; i < 8; ++i) o.v[i] = __shfl_xor_sync(0xffffffffu, term.v[i], off);
      term = term + o;
    }
    trw_copy(a.tr, &sh_tr);
    if (lane == 0) {
      fe_st(a.challenges_out + a.round, ch);
      fe_st(&a.st->r, ch);
      fe_st(&a.st->claim, term);
    }
  }
}

// slot file in SHARED memory, one column per thread: [K values | T temporaries] x blockDim.x. Only the CURRENT
// evaluation of every table is kept: moving from x to x+1 re-reads the pair (L1/L2 hit: this thread read or wrote it
// a moment ago) and adds the step, instead of holding K more step slots per thread — the slot file is what limits
// the resident warps (19 polynomials + eq + rotated z + identity + Lagrange = 23 tables for plonk-with-lookup).
// Constants are read from global memory (uniform address -> broadcast). The host's liveness-based slot reuse
// keeps T tiny (5 for the vanilla-plonk zero check).
static const int GEN_THREADS = 256;
template <bool BIND>
__global__ void __launch_bounds__(GEN_THREADS) sc_generic_round_kernel(GenArgs a) {
  extern __shared__ __align__(32) unsigned char gen_smem_raw[];
  Fr* sm = reinterpret_cast<Fr*>(gen_smem_raw);
  __shared__ Fr smem[GEN_THREADS / 32];
  __shared__ Fr s_tot[GEN_MAX_DEG];
  Fr acc[GEN_MAX_DEG];
  const int K = a.K, D = a.D, KC = a.K + a.C;
  const int nth = blockDim.x, tid = threadIdx.x;
#pragma unroll
  for (int x = 0; x < GEN_MAX_DEG; ++x) acc[x] = fe_zero<FrP>();
  Fr r = fe_zero<FrP>();
  if (BIND) r = fe_ld(&a.st->r);
  const int last = a.ops[a.nops - 1].y;
  auto rd = [&](int idx) -> Fr {
    if (idx < K) return sm[idx * nth + tid];
    if (idx < KC) return fe_ld(a.consts + (idx - K));
    return sm[(K + idx - KC) * nth + tid];
  };

  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < a.pairs; b += gridDim.x * blockDim.x) {
    if (BIND) {  // fused bind of the previous challenge: 4 -> 2 elements per table, stored for the next round
      for (int k = 0; k < K; ++k) {
        const Fr* p = a.in[k] + 4 * (size_t)b;
        const Fr x0 = fe_ldg(p), x1 = fe_ldg(p + 1), x2 = fe_ldg(p + 2), x3 = fe_ldg(p + 3);
        fe_st(a.out[k] + 2 * (size_t)b, (x1 - x0) * r + x0);
        fe_st(a.out[k] + 2 * (size_t)b + 1, (x3 - x2) * r + x2);
      }
    }
#pragma unroll 1
    for (int x = 0; x < D; ++x) {
      // eval(1) = t[2b+1]; eval(x+1) = eval(x) + (t[2b+1] - t[2b])   (eval.rs:228-286)
      for (int k = 0; k < K; ++k) {
        const Fr* p = (BIND ? (const Fr*)a.out[k] : a.in[k]) + 2 * (size_t)b;
        const Fr u0 = fe_ld(p), u1 = fe_ld(p + 1);  // coherent loads: with BIND this thread has just written them
        sm[k * nth + tid] = x == 0 ? u1 : sm[k * nth + tid] + (u1 - u0);
      }
      for (int i = 0; i < a.nops; ++i) {
        const int4 op = a.ops[i];
        const Fr lhs = rd(op.z);
        Fr res;
        switch (op.x) {
          case 0: res = lhs + rd(op.w); break;
          case 1: res = lhs - rd(op.w); break;
          case 2: res = lhs * rd(op.w); break;
          default: res = fe_neg<FrP>(lhs); break;
        }
        sm[(K + op.y - KC) * nth + tid] = res;
      }
      const Fr v = sm[(K + last - KC) * nth + tid];
#pragma unroll
      for (int xx = 0; xx < GEN_MAX_DEG; ++xx)
        if (xx == x) acc[xx] = acc[xx] + v;
    }
  }
  // per-CTA partials
  for (int x = 0; x < D; ++x) {
    Fr t[1] = {fe_zero<FrP>()};
#pragma unroll
    for (int xx = 0; xx < GEN_MAX_DEG; ++xx)
      if (xx == x) t[0] = acc[xx];
    block_reduce_fr<1>(t, smem);
    if (threadIdx.x == 0) fe_st(a.partial + (size_t)blockIdx.x * D + x, t[0]);
  }
  if (!last_cta_ticket(&a.st->counter)) return;
  gen_finalize(a, smem, s_tot);
}

// Small rounds (pairs <= GEN_SMALL_PAIRS): one thread per pair would run K binds and D program evaluations back to
// back (~200 dependent multiplications). Here a CTA owns 32 pairs: the binds are spread over (pair, table) items,
// then warp x evaluates the program at point x+1 for the CTA's pairs, so a round is ~one program evaluation deep.
static const uint32_t GEN_SMALL_PAIRS = 16384;
template <bool BIND>
__global__ void __launch_bounds__(32 * GEN_MAX_DEG) sc_generic_small_kernel(GenArgs a) {
  extern __shared__ __align__(32) unsigned char gen_smem_raw[];
  Fr* sm = reinterpret_cast<Fr*>(gen_smem_raw);
  __shared__ Fr smem[GEN_MAX_DEG];
  __shared__ Fr s_tot[GEN_MAX_DEG];
  const int K = a.K, D = a.D, KC = a.K + a.C;
  const int nth = blockDim.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const uint32_t b0 = blockIdx.x * 32u;
  const uint32_t here = a.pairs - b0 < 32u ? a.pairs - b0 : 32u;
  if (BIND) {
    const Fr r = fe_ld(&a.st->r);
    for (uint32_t it = tid; it < here * (uint32_t)K; it += nth) {
      const uint32_t k = it / here, b = b0 + it % here;
      const Fr* p = a.in[k] + 4 * (size_t)b;
      const Fr x0 = fe_ldg(p), x1 = fe_ldg(p + 1), x2 = fe_ldg(p + 2), x3 = fe_ldg(p + 3);
      fe_st(a.out[k] + 2 * (size_t)b, (x1 - x0) * r + x0);
      fe_st(a.out[k] + 2 * (size_t)b + 1, (x3 - x2) * r + x2);
    }
    __syncthreads();  // the bound pairs are read back by other threads of this CTA
  }
  Fr acc[1] = {fe_zero<FrP>()};
  if ((uint32_t)lane < here) {
    const uint32_t b = b0 + lane;
    const int last = a.ops[a.nops - 1].y;
    auto rd = [&](int idx) -> Fr {
      if (idx < K) return sm[idx * nth + tid];
      if (idx < KC) return fe_ld(a.consts + (idx - K));
      return sm[(K + idx - KC) * nth + tid];
    };
    for (int k = 0; k < K; ++k) {
      const Fr* p = (BIND ? (const Fr*)a.out[k] : a.in[k]) + 2 * (size_t)b;
      const Fr u0 = fe_ld(p), u1 = fe_ld(p + 1);
      const Fr step = u1 - u0;
      Fr v = u1;
      for (int j = 0; j < w; ++j) v = v + step;  // evaluation point w + 1
      sm[k * nth + tid] = v;
    }
    for (int i = 0; i < a.nops; ++i) {
      const int4 op = a.ops[i];
      const Fr lhs = rd(op.z);
      Fr res;
      switch (op.x) {
        case 0: res = lhs + rd(op.w); break;
        case 1: res = lhs - rd(op.w); break;
        case 2: res = lhs * rd(op.w); break;
        default: res = fe_neg<FrP>(lhs); break;
      }
      sm[(K + op.y - KC) * nth + tid] = res;
    }
    acc[0] = sm[(K + last - KC) * nth + tid];
  }
  warp_reduce_fr<1>(acc);
  if (lane == 0) fe_st(a.partial + (size_t)blockIdx.x * D + w, acc[0]);
  if (!last_cta_ticket(&a.st->counter)) return;
  gen_finalize(a, smem, s_tot);
}

__global__ void gen_final_bind_kernel(const Fr* const* tabs, int ntabs, const ScState* st, Fr* evals_out) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntabs) return;
  const Fr r = fe_ld(&st->r);
  const Fr x0 = fe_ld(tabs[i]), x1 = fe_ld(tabs[i] + 1);
  fe_st(evals_out + i, (x1 - x0) * r + x0);
}
__global__ void gen_init_kernel(ScState* st, const Fr* claim) {
  if (threadIdx.x == 0) {
    fe_st(&st->claim, fe_ld(claim));
    fe_st(&st->r, fe_zero<FrP>());
    st->counter = 0;
  }
}

int sumcheck_prove_generic(Ctx* c, const GenericJob& job) {
  const int n = job.num_vars, K = job.ntables;
  if (n < 1 || n > 30 || K < 1 || K > GEN_MAX_TABLES || job.degree < 1 || job.degree > GEN_MAX_DEG || job.nops < 1 ||
      job.ntemps < 1 || job.ntemps > 64)
    return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  const size_t N = (size_t)1 << n;
  const size_t szA = N / 2, szB = N / 4 ? N / 4 : 1;
  Fr *bufA = nullptr, *bufB = nullptr;
  CUDA_TRY(cudaMallocAsync(&bufA, (size_t)K * szA * sizeof(Fr), s));
  CUDA_TRY(cudaMallocAsync(&bufB, (size_t)K * szB * sizeof(Fr), s));
  gen_init_kernel<<<1, 32, 0, s>>>(c->d_sc, job.claim);
  count_launch(c);
  GenArgs a;
  a.consts = job.consts;
  a.ops = job.ops;
  a.K = K;
  a.C = job.nconsts;
  a.nops = job.nops;
  a.D = job.degree;
  a.st = c->d_sc;
  a.partial = c->d_partial;
  a.tr = c->d_tr;
  a.bary = c->d_bary;
  a.challenges_out = job.challenges_out;
  // shared-memory slot file: (K + T) field elements per thread
  const int nslots = K + job.ntemps;
  int nth = GEN_THREADS;
  while (nth > 32 && (size_t)nslots * nth * sizeof(Fr) > 200 * 1024) nth -= 32;
  const size_t smem_bytes = (size_t)nslots * nth * sizeof(Fr);
  if (smem_bytes > 220 * 1024) return B200_ERR_ARG;
  const size_t small_bytes = (size_t)nslots * 32 * job.degree * sizeof(Fr);
  if (small_bytes > 220 * 1024) return B200_ERR_ARG;
  const Fr* cur[GEN_MAX_TABLES];
  for (int i = 0; i < K; ++i) cur[i] = job.tables[i];
  for (int round = 0; round < n; ++round) {
    a.round = round;
    a.pairs = (uint32_t)(N >> (round + 1));
    Fr* dst_base = (round & 1) ? bufA : bufB;
    const size_t dst_sz = (round & 1) ? szA : szB;
    for (int i = 0; i < K; ++i) {
      a.in[i] = cur[i];
      a.out[i] = dst_base + (size_t)i * dst_sz;
    }
    const bool small = a.pairs <= GEN_SMALL_PAIRS;
    int blocks = small ? (int)((a.pairs + 31) / 32) : (int)((a.pairs + nth - 1) / nth);
    if (!small && blocks > NUM_SMS) blocks = NUM_SMS;  // one CTA per SM (the slot file takes most of the shared memory)
    if ((size_t)blocks * job.degree > c->partial_elems) return B200_ERR_NOMEM;
    const int pi = prof_begin(c, round);
    if (small) {
      const int sth = 32 * job.degree;
      const size_t sbytes = (size_t)nslots * sth * sizeof(Fr);
      if (round == 0) sc_generic_small_kernel<false><<<blocks, sth, sbytes, s>>>(a);
      else sc_generic_small_kernel<true><<<blocks, sth, sbytes, s>>>(a);
    } else if (round == 0) {
      sc_generic_round_kernel<false><<<blocks, nth, smem_bytes, s>>>(a);
    } else {
      sc_generic_round_kernel<true><<<blocks, nth, smem_bytes, s>>>(a);
    }
    if (round > 0)
      for (int i = 0; i < K; ++i) cur[i] = a.out[i];
    prof_end(c, pi);
    count_launch(c);
  }
  const Fr** d_ptrs = nullptr;
  CUDA_TRY(cudaMallocAsync(&d_ptrs, K * sizeof(Fr*), s));
  CUDA_TRY(cudaMemcpyAsync(d_ptrs, cur, K * sizeof(Fr*), cudaMemcpyHostToDevice, s));
  gen_final_bind_kernel<<<(K + 63) / 64, 64, 0, s>>>(d_ptrs, K, c->d_sc, job.evals_out);
  count_launch(c);
  CUDA_TRY(cudaFreeAsync(d_ptrs, s));
  CUDA_TRY(cudaFreeAsync(bufA, s));
  CUDA_TRY(cudaFreeAsync(bufB, s));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// ---- leaf tables -----------------------------------------------------------------------------------
// identity polynomial: out[b] = F::from(b)      (CommonPolynomial::Identity, sum_check.rs:123-125)
__global__ void poly_iota_kernel(Fr* out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) fe_st(out + b, fe_from_u64<FrP>(b));
}
// Lagrange_i: one-hot at hypercube point `index`
__global__ void poly_onehot_kernel(Fr* out, size_t n, size_t index) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride)
    fe_st(out + b, b == index ? fe_one<FrP>() : fe_zero<FrP>());
}
// rotated[b] = poly[bh.rotate(b, rotation)]   (classic.rs:105-126, bh.rs:105-153)
__global__ void poly_rotate_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, int num_vars, uint32_t primitive,
                                   int rotation) {
  const size_t n = (size_t)1 << num_vars, stride = (size_t)gridDim.x * blockDim.x;
  const uint64_t x_inv = primitive >> 1;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
    uint64_t s = b;
    for (int i = 0; i < rotation; ++i) {
      s <<= 1;
      s ^= (s >> num_vars) * primitive;
    }
    for (int i = 0; i > rotation; --i) s = (s >> 1) ^ ((s & 1) * x_inv);
    fe_st(out + b, fe_ldg(in + s));
  }
}
static int grid_for(size_t n) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  return blocks < 1 ? 1 : blocks;
}
int poly_iota(Ctx* c, int num_vars, Fr* d_out) {
  const size_t n = (size_t)1 << num_vars;
  poly_iota_kernel<<<grid_for(n), 256, 0, c->stream>>>(d_out, n);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
int poly_onehot(Ctx* c, int num_vars, uint64_t index, Fr* d_out) {
  const size_t n = (size_t)1 << num_vars;
  if (index >= n) return B200_ERR_ARG;
  poly_onehot_kernel<<<grid_for(n), 256, 0, c->stream>>>(d_out, n, index);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}
int poly_rotate(Ctx* c, const Fr* d_in, int num_vars, int rotation, Fr* d_out) {
  static const uint32_t PRIM[32] = {1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771,
                                    65581, 131081, 262183, 524327, 1048585, 2097157, 4194307, 8388641, 16777243,
                                    33554441, 67108935, 134217767, 268435465, 536870917, 1073741907, 2147483657u};
  if (num_vars < 1 || num_vars > 31 || rotation < -num_vars || rotation > num_vars) return B200_ERR_ARG;
  poly_rotate_kernel<<<grid_for((size_t)1 << num_vars), 256, 0, c->stream>>>(d_in, d_out, num_vars, PRIM[num_vars], rotation);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_generic() {
  // dynamic shared memory opt-in (up to the 220 KiB the launch sites allow), once per device: cudaFuncSetAttribute waits
  // for kernels in flight, so it must not sit in the middle of a proof
  const int max_smem = 220 * 1024;
  cudaFuncSetAttribute(sc_generic_round_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(sc_generic_round_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(sc_generic_small_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  cudaFuncSetAttribute(sc_generic_small_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, max_smem);
  B200_PRELOAD(sc_generic_round_kernel<false>);
  B200_PRELOAD(sc_generic_round_kernel<true>);
  B200_PRELOAD(sc_generic_small_kernel<false>);
  B200_PRELOAD(sc_generic_small_kernel<true>);
  B200_PRELOAD(gen_final_bind_kernel);
  B200_PRELOAD(gen_init_kernel);
  B200_PRELOAD(poly_iota_kernel);
  B200_PRELOAD(poly_onehot_kernel);
  B200_PRELOAD(poly_rotate_kernel);
}

}  // namespace b200

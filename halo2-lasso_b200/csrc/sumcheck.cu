// ClassicSumCheck on the GPU: one kernel per round that FUSES
//   (1) the bind of the previous round's challenge (ProverState::next_round / fix_var_in_place,
//       pb/piop/sum_check/classic.rs:90-141, pb/poly/multilinear.rs:599-618),
//   (2) the round evaluation at x = 1..d over hypercube pairs (EvaluationsProver::evals,
//       pb/piop/sum_check/classic/eval.rs:102-131; p(0) DERIVED as sum - p(1), :129), or the
//       Karatsuba coefficients of CoefficientsProver (classic/coeff.rs:136-203),
//   (3) the Fiat-Shamir step (write d+1 elements, squeeze r: classic.rs:225-236) and the claim fold
//       (barycentric_interpolate / horner), executed by the last CTA to finish.
// Tables are read with 256-bit loads (4 consecutive elements per thread when binding), bound values
// are written once (ping-pong scratch), so a table element moves 3 times over a whole sum-check
// instead of the reference's 5 (32*P*(4*2^n - 3) bytes).
#include "internal.h"

namespace b200 {

struct ScEvalArgs {
  const Fr* eq_in;
  Fr* eq_out;
  const Fr* in[SC_MAX_TABLES];
  Fr* out[SC_MAX_TABLES];
  const Fr* weights;
  ScState* st;
  Fr* partial;
  Transcript* tr;
  const BaryTable* bary;
  Fr* challenges_out;
  uint32_t pairs;
  int round;
  long long* dbg;  // optional clock64() trace of the last CTA (debug builds of the bench only)
  PeerCtx peer;    // world > 1: the round totals are summed over all ranks through the peer mailboxes
  unsigned int seq;
};
#define DBG_CLK(i)                                                     \
  do {                                                                 \
    if (a.dbg && threadIdx.x == 0) {                                   \
      a.dbg[a.round * 16 + (i)] = clock64();                           \
      /* wall-clock (ns) copies of stamps 0, 6, 7, 11 in the spare slots 12..15 */ \
      if ((i) == 0 || (i) == 6 || (i) == 7 || (i) == 11) {             \
        unsigned long long ns_;                                        \
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ns_));        \
        a.dbg[a.round * 16 + ((i) == 0 ? 12 : (i) == 6 ? 13 : (i) == 7 ? 14 : 15)] = (long long)ns_; \
      }                                                                \
    }                                                                  \
  } while (0)

// Load the pair (u0, u1) = (t[2b], t[2b+1]) of the CURRENT round. With BIND the table still has the
// previous round's size: bind 4 consecutive elements with r first and store the bound pair.
template <bool BIND>
__device__ __forceinline__ void load_pair(const Fr* __restrict__ in, Fr* __restrict__ out, uint32_t b,
                                          const Fr& r, bool store, Fr& u0, Fr& u1) {
  if (BIND) {
    const Fr* p = in + 4 * (size_t)b;
    Fr x0 = fe_ldg(p), x1 = fe_ldg(p + 1), x2 = fe_ldg(p + 2), x3 = fe_ldg(p + 3);
    u0 = (x1 - x0) * r + x0;
    u1 = (x3 - x2) * r + x2;
    if (store) {
      fe_st(out + 2 * (size_t)b, u0);
      fe_st(out + 2 * (size_t)b + 1, u1);
    }
  } else {
    const Fr* p = in + 2 * (size_t)b;
    u0 = fe_ldg(p);
    u1 = fe_ldg(p + 1);
  }
}

// EQPRE: the eq table of this round was already bound by a separate (tiny) kernel, so the T grid rows do not
// each redo its two binding products per pair (used for the batched grand-product layers, T >= 4).
template <int NP, bool BIND, bool EQPRE = false>
__global__ void __launch_bounds__(SC_THREADS) sc_eval_round_kernel(ScEvalArgs a) {
  constexpr int D = NP + 1;  // degree; evaluations at 1..D are computed, p(0) derived
  __shared__ Fr smem[(SC_THREADS / 32) * D];
  pdl_prologue();
  const int t = blockIdx.y;
  Fr acc[D];
#pragma unroll
  for (int x = 0; x < D; ++x) acc[x] = fe_zero<FrP>();
  Fr r = fe_zero<FrP>();
  if (BIND) r = fe_ld(&a.st->r);
  const bool dbg_cta = a.dbg && gridDim.x * gridDim.y == 1;
  if (dbg_cta) DBG_CLK(0);

  const Fr* __restrict__ in0 = a.in[t * NP];
  Fr* __restrict__ out0 = a.out[t * NP];
  const Fr* __restrict__ in1 = NP == 2 ? a.in[t * NP + 1] : nullptr;
  Fr* __restrict__ out1 = NP == 2 ? a.out[t * NP + 1] : nullptr;

  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < a.pairs; b += gridDim.x * blockDim.x) {
    Fr e0, e1, p0, p1, q0, q1;
    if (EQPRE) load_pair<false>(a.eq_out, nullptr, b, r, false, e0, e1);
    else load_pair<BIND>(a.eq_in, a.eq_out, b, r, t == 0, e0, e1);
    load_pair<BIND>(in0, out0, b, r, true, p0, p1);
    if (NP == 2) load_pair<BIND>(in1, out1, b, r, true, q0, q1);
    // eval = t[2b+1], step = t[2b+1] - t[2b]; x -> x+1 adds the step (eval.rs:228-286)
    e0 = e1 - e0;
    p0 = p1 - p0;
    if (NP == 2) q0 = q1 - q0;
#pragma unroll
    for (int x = 0; x < D; ++x) {
      Fr prod = e1 * p1;
      if (NP == 2) prod = prod * q1;
      acc[x] = acc[x] + prod;
      if (x + 1 < D) {
        e1 = e1 + e0;
        p1 = p1 + p0;
        if (NP == 2) q1 = q1 + q0;
      }
    }
  }
  if (dbg_cta) DBG_CLK(1);
  block_reduce_fr<D>(acc, smem);
  if (dbg_cta) DBG_CLK(2);
  if (threadIdx.x < 32) {  // lane x scales and stores partial x (one multiplication deep, not D)
    const int lane = threadIdx.x;
    Fr v = fe_zero<FrP>();
#pragma unroll
    for (int x = 0; x < D; ++x) {
      const Fr tmp = fr_bcast(acc[x], 0);
      if (lane == x) v = tmp;
    }
    v = fr_mul_ni(v, fe_ld(a.weights + t));
    if (lane < D) fe_st(a.partial + ((size_t)t * gridDim.x + blockIdx.x) * D + lane, v);
  }
  if (dbg_cta) DBG_CLK(3);
  if (!last_cta_ticket(&a.st->counter)) return;
  if (dbg_cta) DBG_CLK(4);

  // ---- last CTA: total, derive p(0), Fiat-Shamir, fold the claim -------------------------------
  const uint32_t nparts = gridDim.x * gridDim.y;
#pragma unroll
  for (int x = 0; x < D; ++x) acc[x] = fe_zero<FrP>();
  const bool keep_busy = a.peer.world > 1;
  if (nparts <= 32) {  // small rounds: one warp sums the partials, no block-wide barrier
    if (threadIdx.x >= 32) {
      if (!keep_busy) return;
    } else {
      if (threadIdx.x < nparts) {
#pragma unroll
        for (int x = 0; x < D; ++x) acc[x] = fr_ld_cg(a.partial + (size_t)threadIdx.x * D + x);
      }
      warp_reduce_fr<D>(acc);
    }
  } else {
    for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
#pragma unroll
      for (int x = 0; x < D; ++x) acc[x] = acc[x] + fr_ld_cg(a.partial + (size_t)i * D + x);
    }
    block_reduce_fr<D>(acc, smem);
  }
  // Sharded rounds: warps 1-3 keep the SM busy while warp 0 runs the exchange and the finalize (peer.cuh)
  __shared__ volatile int s_busy;
  if (keep_busy) {
    if (threadIdx.x == 0) s_busy = 1;
    __syncthreads();
    if (threadIdx.x >= 128) return;
    if (threadIdx.x >= 32) {
      peer_spin_while(&s_busy, &a.st->pad[0]);
      return;
    }
  }
  if (threadIdx.x < 32) {  // warp 0, warp-uniform control flow; lane i owns p(i)
    const int lane = threadIdx.x;
    __shared__ Transcript sh_tr;
    if (dbg_cta) DBG_CLK(5);
    trw_copy(&sh_tr, a.tr);
    if (dbg_cta) DBG_CLK(6);
    Fr tot = fe_zero<FrP>();  // lane x < D owns the total of evaluation point x+1
#pragma unroll
    for (int x = 0; x < D; ++x) {
      const Fr tmp = fr_bcast(acc[x], 0);
      if (lane == x) tot = tmp;
    }
    if (a.peer.world > 1) {  // fused collective: all-gather the D partials over NVLink and add them
      peer_publish(a.peer, a.seq, tot, D);
      Fr sum = fe_zero<FrP>();
      if (lane < D)
        for (int r = 0; r < a.peer.world; ++r) sum = sum + peer_read(a.peer, a.seq, r, lane);
      tot = sum;
    }
    const Fr p1 = fr_bcast(tot, 0);
    Fr mine = fe_zero<FrP>();
#pragma unroll
    for (int x = 0; x < D; ++x) {
      const Fr tmp = fr_bcast(tot, x);
      if (lane == x + 1) mine = tmp;
    }
    if (lane == 0) mine = fe_ld(&a.st->claim) - p1;  // p(0) = sum - p(1)   (eval.rs:129)
    const Fr canon = fr_canon_ni(mine);              // D+1 conversions in parallel lanes
    if (dbg_cta) DBG_CLK(7);
    for (int x = 0; x <= D; ++x) trw_write_canon_from_lane(&sh_tr, canon, x, true);
    if (dbg_cta) DBG_CLK(8);
    const Fr ch = trw_squeeze(&sh_tr);
    if (dbg_cta) DBG_CLK(9);
    // next claim p(ch) = Σ_i p(i) w_i Π_{j != i} (ch - j): lane i builds its own term
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
    if (dbg_cta) DBG_CLK(10);
    trw_copy(a.tr, &sh_tr);
    if (lane == 0) {
      fe_st(a.challenges_out + a.round, ch);
      fe_st(&a.st->r, ch);
      fe_st(&a.st->claim, term);
    }
    if (dbg_cta) DBG_CLK(11);
    if (keep_busy && lane == 0) s_busy = 0;
  }
}


// ---------------------------------------------------------------------------------------------
// Eq-FACTORED round kernel. eq(x, y) is a product over the variables, so the round polynomial of
// F = eq * G factors as  p_i(X) = c_i * eq1(X, y_i) * Q_i(X),  c_i = Π_{j<i} eq1(r_j, y_j),
// Q_i(X) = Σ_x' E_i[x'] * G(r, X, x'),  E_i = eq table of the REMAINING variables i+1..n-1 (a fixed table, never
// bound). Q_i has degree NP, so per pair the kernel accumulates Q(1), Q(2) and the leading coefficient
// (Q(3) = 2 Q(2) - Q(1) + 2 lead) with  E*p1, E*dp  (E*p2 is their sum)  then  *q1, *q2, *dq : 5 products plus the
// 4 binding products = 9 per pair instead of 12, no eq-table bind, and one element read instead of four (and none
// written) for the eq factor. The message is assembled exactly as the reference's: p(x) for x = 1..D from the
// formula above — the same field elements as Σ eq * G — and p(0) = claim - p(1) (eval.rs:129); verified against the
// reference messages, also for inconsistent claims, by tests/test_oracle_protocols.py::test_eq_factored_...
// c_i and the D factors c_i * eq1(x, y_i) are computed by warp 1 of the last CTA while warp 0 sums the partials.
// ---------------------------------------------------------------------------------------------
struct ScFactArgs {
  const Fr* esuf;      // E_round: 2^(n-1-round) entries
  const Fr* y;         // device: the eq point (y[round] is this round's coordinate)
  const Fr* eq_scale;  // optional device scalar folded into c_0
  const Fr* in[SC_MAX_TABLES];
  Fr* out[SC_MAX_TABLES];
  const Fr* weights;
  ScState* st;
  Fr* partial;
  Transcript* tr;
  const BaryTable* bary;
  Fr* challenges_out;
  uint32_t pairs;
  int round;
  PeerCtx peer;
  unsigned int seq;
};

template <int NP, bool BIND>
__global__ void __launch_bounds__(SC_THREADS) sc_eval_fact_kernel(ScFactArgs a) {
  constexpr int D = NP + 1;  // accumulators: NP = 2: Q(1), Q(2), lead;  NP = 1: Q(1), slope
  __shared__ Fr smem[(SC_THREADS / 32) * D];
  pdl_prologue();
  const int t = blockIdx.y;
  Fr acc[D];
#pragma unroll
  for (int x = 0; x < D; ++x) acc[x] = fe_zero<FrP>();
  Fr r = fe_zero<FrP>();
  if (BIND) r = fe_ld(&a.st->r);
  const Fr* __restrict__ in0 = a.in[t * NP];
  Fr* __restrict__ out0 = a.out[t * NP];
  const Fr* __restrict__ in1 = NP == 2 ? a.in[t * NP + 1] : nullptr;
  Fr* __restrict__ out1 = NP == 2 ? a.out[t * NP + 1] : nullptr;
  const Fr* __restrict__ esuf = a.esuf;

  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < a.pairs; b += gridDim.x * blockDim.x) {
    Fr p0, p1, q0, q1;
    const Fr e = fe_ldg(esuf + b);
    load_pair<BIND>(in0, out0, b, r, true, p0, p1);
    if (NP == 2) load_pair<BIND>(in1, out1, b, r, true, q0, q1);
    const Fr dp = p1 - p0;
    const Fr ep1 = e * p1, edp = e * dp;
    if (NP == 1) {
      acc[0] = acc[0] + ep1;
      acc[1] = acc[1] + edp;
    } else {
      const Fr dq = q1 - q0;
      acc[0] = acc[0] + ep1 * q1;
      acc[1] = acc[1] + (ep1 + edp) * (q1 + dq);
      acc[NP] = acc[NP] + edp * dq;
    }
  }
  block_reduce_fr<D>(acc, smem);
  if (threadIdx.x < 32) {  // lane x scales and stores partial x (one multiplication deep, not D)
    const int lane = threadIdx.x;
    Fr v = fe_zero<FrP>();
#pragma unroll
    for (int x = 0; x < D; ++x) {
      const Fr tmp = fr_bcast(acc[x], 0);
      if (lane == x) v = tmp;
    }
    v = fr_mul_ni(v, fe_ld(a.weights + t));
    if (lane < D) fe_st(a.partial + ((size_t)t * gridDim.x + blockIdx.x) * D + lane, v);
  }
  if (!last_cta_ticket(&a.st->counter)) return;

  // ---- last CTA: total, eq factor, derive p(0), Fiat-Shamir, fold the claim ---------------------
  const uint32_t nparts = gridDim.x * gridDim.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __shared__ Fr s_f[4];
  const bool keep_busy = a.peer.world > 1;
#pragma unroll
  for (int x = 0; x < D; ++x) acc[x] = fe_zero<FrP>();
  if (warp == 1) {  // off the critical path: c_i = c_(i-1) * eq1(r_(i-1), y_(i-1)) and f_x = c_i * eq1(x, y_i)
    const Fr one = fe_one<FrP>();
    Fr c;
    if (a.round == 0) {
      c = a.eq_scale ? fe_ld(a.eq_scale) : one;
    } else {
      const Fr yp = fe_ld(a.y + a.round - 1);
      const Fr ry = fr_mul_ni(r, yp);
      c = fr_mul_ni(fe_ld(&a.st->eqc), ry + ry + one - r - yp);
    }
    const Fr yi = fe_ld(a.y + a.round);
    const Fr step = yi + yi - one;  // eq1(x + 1, y) - eq1(x, y)
    Fr ex = yi;                     // eq1(1, y) = y
    for (int x = 1; x <= lane && x < D; ++x) ex = ex + step;
    const Fr f = fr_mul_ni(c, ex);
    if (lane < D) s_f[lane] = f;
    if (lane == 0) fe_st(&a.st->eqc, c);
  }
  if (nparts <= 32) {  // small rounds: one warp sums the partials
    if (warp == 0) {
      if (threadIdx.x < nparts) {
#pragma unroll
        for (int x = 0; x < D; ++x) acc[x] = fr_ld_cg(a.partial + (size_t)threadIdx.x * D + x);
      }
      warp_reduce_fr<D>(acc);
    }
    if (warp < 2) asm volatile("bar.sync 1, 64;" ::: "memory");  // s_f of warp 1 -> warp 0
    if (warp >= 1 && !keep_busy) return;
  } else {
    if (warp != 1) {  // warp 1 is busy with the eq factor: the other 7 warps stride over the partials
      const uint32_t tid7 = warp == 0 ? threadIdx.x : threadIdx.x - 32;
      for (uint32_t i = tid7; i < nparts; i += blockDim.x - 32) {
#pragma unroll
        for (int x = 0; x < D; ++x) acc[x] = acc[x] + fr_ld_cg(a.partial + (size_t)i * D + x);
      }
    }
    block_reduce_fr<D>(acc, smem);  // its barriers also publish s_f
  }
  // Sharded rounds: warps 1-3 keep the SM busy while warp 0 runs the exchange and the finalize (peer.cuh)
  __shared__ volatile int s_busy;
  if (keep_busy) {
    if (threadIdx.x == 0) s_busy = 1;
    __syncthreads();
    if (threadIdx.x >= 128) return;
    if (threadIdx.x >= 32) {
      peer_spin_while(&s_busy, &a.st->pad[0]);
      return;
    }
  }
  if (threadIdx.x < 32) {  // warp 0, warp-uniform control flow; lane i owns p(i)
    __shared__ Transcript sh_tr;
    trw_copy(&sh_tr, a.tr);
    // Q(x) for x = 1..D into lane x - 1 (this rank's part)
    const Fr q1 = fr_bcast(acc[0], 0), q2raw = fr_bcast(acc[1], 0);
    Fr qx;
    if (NP == 1) {
      qx = lane == 1 ? q1 + q2raw : q1;  // Q(2) = Q(1) + slope
    } else {
      const Fr lead = fr_bcast(acc[NP], 0);
      const Fr q3 = q2raw + q2raw - q1 + lead + lead;
      qx = lane == 0 ? q1 : (lane == 1 ? q2raw : q3);
    }
    Fr px = fe_zero<FrP>();
    if (lane < D) px = s_f[lane];
    px = fr_mul_ni(px, qx);  // p(x) = c eq1(x, y_i) Q(x); c carries the rank's own eq factor of the window bits
    if (a.peer.world > 1) {  // fused collective: all-gather the D values over NVLink and add them
      peer_publish(a.peer, a.seq, px, D);
      Fr sum = fe_zero<FrP>();
      if (lane < D)
        for (int rr = 0; rr < a.peer.world; ++rr) sum = sum + peer_read(a.peer, a.seq, rr, lane);
      px = sum;
    }
    Fr mine = fe_zero<FrP>();  // lane i owns p(i)
#pragma unroll
    for (int x = 0; x < D; ++x) {
      const Fr tmp = fr_bcast(px, x);
      if (lane == x + 1) mine = tmp;
    }
    const Fr p1 = fr_bcast(mine, 1);
    if (lane == 0) mine = fe_ld(&a.st->claim) - p1;  // p(0) = sum - p(1)   (eval.rs:129)
    const Fr canon = fr_canon_ni(mine);              // D+1 conversions in parallel lanes
    for (int x = 0; x <= D; ++x) trw_write_canon_from_lane(&sh_tr, canon, x, true);
    const Fr ch = trw_squeeze(&sh_tr);
    // next claim p(ch) = Σ_i p(i) w_i Π_{j != i} (ch - j): lane i builds its own term
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
    if (keep_busy && lane == 0) s_busy = 0;
  }
}

// E_i tables of all rounds in one heap: level k (2^k entries, the table of round i = n - 1 - k over the variables
// n-k..n-1) at [2^k, 2^(k+1)). Levels of at most SUF_DIRECT variables are computed entry by entry; a larger level is
// L[x' & mask] * H[x' >> shift] with H the SUF_DIRECT-variable level and L the (small) table over the variables between.
static const int SUF_DIRECT = 11;
__device__ __forceinline__ Fr eq_entry(const Fr* __restrict__ y, int first_var, int nv, uint32_t bits) {
  const Fr one = fe_one<FrP>();
  Fr acc = one;
  for (int j = 0; j < nv; ++j) {
    const Fr yj = fe_ld(y + first_var + j);
    acc = acc * (((bits >> j) & 1) ? yj : one - yj);
  }
  return acc;
}
// heap[e] for e in [1, 2^(nh+1)) and lheap[e] for e in [2, 2^(iH+1)): lheap level kk (2^kk entries at [2^kk, 2^(kk+1))) is
// the table over the variables iH+1-kk..iH
__global__ void __launch_bounds__(128) eq_suffix_small_kernel(const Fr* __restrict__ y, int n, int nh, int iH,
                                                              Fr* __restrict__ heap, Fr* __restrict__ lheap) {
  pdl_prologue();
  const uint32_t nsmall = 1u << (nh + 1), nl = iH >= 0 ? (1u << (iH + 1)) : 0;
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= 1 && e < nsmall) {
    const int k = 31 - __clz(e);
    fe_st(heap + e, eq_entry(y, n - k, k, e - (1u << k)));
  } else if (e >= nsmall + 2 && e < nsmall + nl) {
    const uint32_t w = e - nsmall;
    const int kk = 31 - __clz(w);
    fe_st(lheap + w, eq_entry(y, iH + 1 - kk, kk, w - (1u << kk)));
  }
}
__global__ void __launch_bounds__(256) eq_suffix_big_kernel(int n, int nh, int iH, Fr* __restrict__ heap,
                                                            const Fr* __restrict__ lheap) {
  pdl_prologue();
  const size_t total = (size_t)1 << n, stride = (size_t)gridDim.x * blockDim.x;
  const Fr* __restrict__ H = heap + ((size_t)1 << nh);
  for (size_t e = ((size_t)1 << (nh + 1)) + (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
    const int k = 63 - __clzll((unsigned long long)e);  // level: k variables, n-k..n-1
    const uint32_t x = (uint32_t)(e - ((size_t)1 << k));
    const int kk = k - nh;                               // low variables n-k..iH
    const Fr l = fe_ld(lheap + ((1u << kk) + (x & ((1u << kk) - 1))));
    fe_st(heap + e, l * fe_ld(H + (x >> kk)));
  }
}
static int eq_suffix_build(Ctx* c, DevScope& mem, const Fr* d_y, int n, Fr** heap_out) {
  cudaStream_t s = c->stream;
  const int nh = n - 1 < SUF_DIRECT ? n - 1 : SUF_DIRECT;  // variables of the largest directly computed level
  const int iH = n - 1 - nh;                               // its round; rounds < iH use the product form
  Fr *heap = nullptr, *lheap = nullptr;
  CUDA_TRY(mem.alloc(&heap, sizeof(Fr) << n));
  CUDA_TRY(mem.alloc(&lheap, sizeof(Fr) << (iH + 1)));
  const uint32_t nthreads = (1u << (nh + 1)) + (iH > 0 ? (1u << (iH + 1)) : 0);
  CUDA_TRY(launch_pdl(eq_suffix_small_kernel, dim3((nthreads + 127) / 128), dim3(128), 0, s, d_y, n, nh, iH > 0 ? iH : -1,
                      heap, lheap));
  count_launch(c);
  if (iH > 0) {
    const size_t big = ((size_t)1 << n) - ((size_t)1 << (nh + 1));
    int blocks = (int)((big + 255) / 256);
    if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
    CUDA_TRY(launch_pdl(eq_suffix_big_kernel, dim3(blocks), dim3(256), 0, s, n, nh, iH, heap, (const Fr*)lheap));
    count_launch(c);
  }
  *heap_out = heap;
  return B200_OK;
}
// out[j] = scalar * in[j]
__global__ void __launch_bounds__(256) scale_copy_kernel(const Fr* __restrict__ in, const Fr* __restrict__ scalar,
                                                         Fr* __restrict__ out, size_t n) {
  pdl_prologue();
  const Fr sc = fe_ld(scalar);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) fe_st(out + i, fe_ld(in + i) * sc);
}

// ---------------------------------------------------------------------------------------------
// Tail of a sum-check in ONE launch. Once a round has at most TAIL_ITEMS (pair, term) items its kernel is pure
// latency: launch, per-CTA partials through global memory, the last-CTA ticket, transcript copies in and out. This
// kernel runs ALL remaining rounds in a single CTA: the tables are bound into shared memory once (<= 32 KB), each
// thread owns one (pair, term) item per round, warp 0 runs the Fiat-Shamir step on a transcript that stays in shared
// memory, and the final evaluations (sc_final_bind_kernel's job) are written at the end. Same field operations in
// the same association as the per-round kernels, hence the same transcript bytes.
// ---------------------------------------------------------------------------------------------
static const int TAIL_ITEMS = 256;     // pairs * T of the first tail round
static const int TAIL_ENTRIES = 1024;  // (ntab + 1) * 2 * pairs: shared-memory table entries (32 KB)
struct ScTailArgs {
  const Fr* in[SC_MAX_TABLES + 1];  // tables BEFORE the bind of the first tail round; index ntab = eq table
  const Fr* weights;
  ScState* st;
  Transcript* tr;
  const BaryTable* bary;
  Fr* challenges_out;
  Fr* evals_out;
  uint32_t pairs;  // pairs of the first tail round
  int T, first_round, num_rounds, want_eq_eval;
  const Fr* eq_c;  // eq-factored rounds before the tail: in[ntab] is the unscaled E table, multiplied by *eq_c here
};

template <int NP>
__global__ void __launch_bounds__(256) sc_eval_tail_kernel(ScTailArgs a) {
  constexpr int D = NP + 1;
  __shared__ Fr tab[TAIL_ENTRIES];
  __shared__ Fr red[(256 / 32) * D];
  __shared__ Transcript sh_tr;
  __shared__ Fr s_ch, s_claim;
  pdl_prologue();
  const int tid = threadIdx.x, lane = tid & 31;
  const int ntab = a.T * NP, ntabs = ntab + 1;
  uint32_t pairs = a.pairs;
  Fr r = fe_ld(&a.st->r);
  if (tid == 0) s_claim = fe_ld(&a.st->claim);
  // bind of the previous challenge: 4 entries -> 2 per pair and table, into shared memory
  for (uint32_t e = tid; e < (uint32_t)ntabs * 2 * pairs; e += blockDim.x) {
    const uint32_t i = e / (2 * pairs), idx = e % (2 * pairs);
    const Fr x0 = fe_ldg(a.in[i] + 2 * (size_t)idx), x1 = fe_ldg(a.in[i] + 2 * (size_t)idx + 1);
    Fr v = (x1 - x0) * r + x0;
    if (a.eq_c && i == (uint32_t)ntab) v = v * fe_ld(a.eq_c);
    tab[e] = v;
  }
  if (tid < 32) trw_copy(&sh_tr, a.tr);
  __syncthreads();
  for (int round = a.first_round; round < a.num_rounds; ++round) {
    // ---- evaluate: thread = (term, pair) item ------------------------------------------------------
    Fr acc[D];
#pragma unroll
    for (int x = 0; x < D; ++x) acc[x] = fe_zero<FrP>();
    if ((uint32_t)tid < pairs * (uint32_t)a.T) {
      const uint32_t t = tid / pairs, b = tid % pairs;
      const Fr* et = tab + (size_t)ntab * 2 * pairs;
      const Fr* pt = tab + (size_t)(t * NP) * 2 * pairs;
      Fr e0 = et[2 * b], e1 = et[2 * b + 1], p0 = pt[2 * b], p1 = pt[2 * b + 1], q0, q1;
      if (NP == 2) {
        const Fr* qt = pt + 2 * pairs;
        q0 = qt[2 * b];
        q1 = qt[2 * b + 1];
      }
      e0 = e1 - e0;
      p0 = p1 - p0;
      if (NP == 2) q0 = q1 - q0;
      const Fr w = fe_ld(a.weights + t);
#pragma unroll
      for (int x = 0; x < D; ++x) {
        Fr prod = e1 * p1;
        if (NP == 2) prod = prod * q1;
        acc[x] = prod * w;
        if (x + 1 < D) {
          e1 = e1 + e0;
          p1 = p1 + p0;
          if (NP == 2) q1 = q1 + q0;
        }
      }
    }
    block_reduce_fr<D>(acc, red);
    // ---- warp 0: message, challenge, next claim (same scheme as sc_eval_round_kernel) -----------------
    if (tid < 32) {
      Fr mine = fe_zero<FrP>();  // lane i owns p(i)
#pragma unroll
      for (int x = 0; x < D; ++x) {
        const Fr tmp = fr_bcast(acc[x], 0);
        if (lane == x + 1) mine = tmp;
      }
      const Fr p1v = fr_bcast(mine, 1);
      if (lane == 0) mine = s_claim - p1v;  // p(0) = sum - p(1)   (eval.rs:129)
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
      if (lane == 0) {
        s_ch = ch;
        s_claim = term;
        fe_st(a.challenges_out + round, ch);
      }
    }
    __syncthreads();
    r = s_ch;
    if (round + 1 == a.num_rounds) break;
    // ---- bind for the next round: 2 * pairs entries -> pairs entries per table (in place, two phases) ----
    const uint32_t half = pairs;  // new entries per table
    Fr nv[4];
    int cnt = 0;
    for (uint32_t e = tid; e < (uint32_t)ntabs * half; e += blockDim.x) {
      const uint32_t i = e / half, idx = e % half;
      const Fr x0 = tab[(size_t)i * 2 * pairs + 2 * idx], x1 = tab[(size_t)i * 2 * pairs + 2 * idx + 1];
      nv[cnt++] = (x1 - x0) * r + x0;
    }
    __syncthreads();
    cnt = 0;
    for (uint32_t e = tid; e < (uint32_t)ntabs * half; e += blockDim.x) tab[e] = nv[cnt++];  // table i at i * half
    __syncthreads();
    pairs >>= 1;
  }
  // ---- final bind (tables have 2 entries) and state write-back ------------------------------------------
  const int nfinal = ntab + (a.want_eq_eval ? 1 : 0);
  for (int i = tid; i < nfinal; i += blockDim.x) {
    const Fr x0 = tab[2 * i], x1 = tab[2 * i + 1];
    fe_st(a.evals_out + i, (x1 - x0) * r + x0);
  }
  if (tid < 32) {
    trw_copy(a.tr, &sh_tr);
    if (lane == 0) {
      fe_st(&a.st->r, r);
      fe_st(&a.st->claim, s_claim);
    }
  }
}

// After the last round every table has 2 entries: bind them with the last challenge.
__global__ void sc_final_bind_kernel(const Fr* const* tabs, int ntabs, const ScState* st, Fr* evals_out) {
  pdl_prologue();
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ntabs) return;
  const Fr r = fe_ld(&st->r);
  const Fr x0 = fe_ld(tabs[i]), x1 = fe_ld(tabs[i] + 1);
  fe_st(evals_out + i, (x1 - x0) * r + x0);
}

__global__ void sc_init_kernel(ScState* st, const Fr* claim) {
  pdl_prologue();
  if (threadIdx.x == 0) {
    fe_st(&st->claim, fe_ld(claim));
    fe_st(&st->r, fe_zero<FrP>());
    st->counter = 0;
  }
}

// Grid of a round kernel: `rows` rows (terms) of equal work, at most ONE wave. `occ` = CTAs of that kernel an SM can hold
// (register-limited: 2 for the NP = 2 kernels, 3 for NP = 1; measured at preload, g_occ). Round 1 sized the grid as
// ceil(2 * 148 / rows) * rows, which for 16 rows is 304 CTAs on 296 slots: the 8 CTAs of the second wave ran alone and
// stretched every big grand-product round by half (38 G products/s at T = 16 against 47 at T = 1 on the same kernel).
enum { OCC_EVAL = 0, OCC_FACT = 4, OCC_COEFF = 8 };  // + (NP - 1) * 2 + BIND
static int g_occ[10] = {2, 2, 2, 2, 2, 2, 2, 2, 2, 2};
static inline int blocks_for(uint32_t pairs, int rows, int occ) {
  if (occ < 1) occ = 1;
  int per_row = (occ * NUM_SMS) / rows;
  if (per_row < 1) per_row = 1;
  int need = (int)((pairs + SC_THREADS - 1) / SC_THREADS);
  if (need < 1) need = 1;
  return need < per_row ? need : per_row;
}

int sumcheck_prove_evals(Ctx* c, const ScEvalJob& job) {
  const int n = job.num_vars, T = job.T, NP = job.NP, ntab = T * NP;
  if (n < 1 || n > 30 || T < 1 || T > SC_MAX_TERMS || (NP != 1 && NP != 2)) return B200_ERR_ARG;
  NvtxRange nvtx("sum_check_prove-%d-%d", n, NP + 1);  // classic.rs:215-218
  cudaStream_t s = c->stream;
  const size_t N = (size_t)1 << n;
  // scratch: eq table (N) + ping-pong halves for eq and every table
  const size_t szA = N / 2, szB = N / 4 ? N / 4 : 1;
  Fr *eq0 = nullptr, *bufA = nullptr, *bufB = nullptr;
  if ((job.stop_after > 0) != (job.carry != nullptr) || job.stop_after > n) return B200_ERR_ARG;
  DevScope own(s);
  DevScope& mem = job.carry ? *job.carry->scope : own;
  CUDA_TRY(mem.alloc(&bufA, (size_t)(ntab + 1) * szA * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&bufB, (size_t)(ntab + 1) * szB * sizeof(Fr)));
  int rc;
  // eq-factored rounds (sc_eval_fact_kernel) whenever the eq table is built here from its point
  const bool fact = c->eq_factored && !job.eq_table && !job.want_eq_eval && n >= 3;
  Fr* suf = nullptr;  // heap of the E_i tables
  if (fact) {
    rc = eq_suffix_build(c, mem, job.eq_point, n, &suf);
    if (rc) return rc;
  } else if (!job.eq_table) {
    CUDA_TRY(mem.alloc(&eq0, N * sizeof(Fr)));
    rc = eq_build(c, job.eq_point, n, eq0);
    if (rc) return rc;
    if (job.eq_scale) {
      rc = fr_scale(c, eq0, N, job.eq_scale);
      if (rc) return rc;
    }
  }
  CUDA_TRY(launch_pdl(sc_init_kernel, dim3(1), dim3(32), 0, s, c->d_sc, job.claim));
  count_launch(c);
  ScFactArgs fa;
  fa.y = job.eq_point;
  fa.eq_scale = job.eq_scale;
  fa.weights = job.weights;
  fa.st = c->d_sc;
  fa.partial = c->d_partial;
  fa.tr = c->d_tr;
  fa.bary = c->d_bary;
  fa.challenges_out = job.challenges_out;
  fa.peer = c->peer;
  if (!job.sharded) fa.peer.world = 1;
  fa.seq = 0;

  ScEvalArgs a;
  a.weights = job.weights;
  a.st = c->d_sc;
  a.partial = c->d_partial;
  a.tr = c->d_tr;
  a.bary = c->d_bary;
  a.challenges_out = job.challenges_out;
  a.dbg = c->dbg_clocks;
  a.peer = c->peer;
  if (!job.sharded) a.peer.world = 1;
  a.seq = 0;
  const Fr* cur[SC_MAX_TABLES + 1];  // current (unbound) tables; slot ntab = eq
  for (int i = 0; i < ntab; ++i) cur[i] = job.tables[i];
  cur[ntab] = job.eq_table ? job.eq_table : eq0;
  bool tail_done = false;
  const int nrounds = job.stop_after > 0 ? job.stop_after : n;
  for (int round = 0; round < nrounds; ++round) {
    a.round = round;
    a.pairs = (uint32_t)(N >> (round + 1));
    if (a.peer.world > 1) a.seq = ++c->peer_seq;
    Fr* dst_base = (round & 1) ? bufA : bufB;  // round 1 writes A, round 2 writes B, ...
    const size_t dst_sz = (round & 1) ? szA : szB;
    a.eq_in = cur[ntab];
    a.eq_out = dst_base + (size_t)ntab * dst_sz;
    for (int i = 0; i < ntab; ++i) {
      a.in[i] = cur[i];
      a.out[i] = dst_base + (size_t)i * dst_sz;
    }
    if (round >= (fact ? 2 : 1) && !c->profile && !c->dbg_clocks && a.peer.world == 1 && !job.carry &&
        (size_t)a.pairs * T <= TAIL_ITEMS && (size_t)(ntab + 1) * 2 * a.pairs <= TAIL_ENTRIES) {
      // all remaining rounds (and the final bind) in one single-CTA launch
      ScTailArgs ta;
      for (int i = 0; i <= ntab; ++i) ta.in[i] = cur[i];
      ta.eq_c = nullptr;
      if (fact) {  // the eq table before the bind of challenge round-1 is c_(round-1) * E_(round-2)
        ta.in[ntab] = suf + ((size_t)1 << (n + 1 - round));
        ta.eq_c = &c->d_sc->eqc;
      }
      ta.weights = job.weights;
      ta.st = c->d_sc;
      ta.tr = c->d_tr;
      ta.bary = c->d_bary;
      ta.challenges_out = job.challenges_out;
      ta.evals_out = job.evals_out;
      ta.pairs = a.pairs;
      ta.T = T;
      ta.first_round = round;
      ta.num_rounds = n;
      ta.want_eq_eval = job.want_eq_eval ? 1 : 0;
      if (NP == 1) CUDA_TRY(launch_pdl(sc_eval_tail_kernel<1>, dim3(1), dim3(256), 0, s, ta));
      else CUDA_TRY(launch_pdl(sc_eval_tail_kernel<2>, dim3(1), dim3(256), 0, s, ta));
      count_launch(c);
      tail_done = true;
      break;
    }
    const int bind = round > 0 ? 1 : 0;
    dim3 grid(blocks_for(a.pairs, T, g_occ[(fact ? OCC_FACT : OCC_EVAL) + (NP - 1) * 2 + bind]), T);
    if ((size_t)grid.x * grid.y * (NP + 1) > c->partial_elems) return B200_ERR_NOMEM;
    NvtxRange nvtx_round("sum_check_prove_round-%d", round);  // classic.rs:226 (+ next_round :234, fused into the launch)
    const int pi = prof_begin(c, round);
    if (fact) {
      fa.esuf = suf + ((size_t)1 << (n - 1 - round));
      fa.pairs = a.pairs;
      fa.round = round;
      fa.seq = a.seq;
      for (int i = 0; i < ntab; ++i) {
        fa.in[i] = cur[i];
        fa.out[i] = a.out[i];
      }
      if (round == 0) {
        if (NP == 1) CUDA_TRY(launch_pdl(sc_eval_fact_kernel<1, false>, grid, SC_THREADS, 0, s, fa));
        else CUDA_TRY(launch_pdl(sc_eval_fact_kernel<2, false>, grid, SC_THREADS, 0, s, fa));
      } else {
        if (NP == 1) CUDA_TRY(launch_pdl(sc_eval_fact_kernel<1, true>, grid, SC_THREADS, 0, s, fa));
        else CUDA_TRY(launch_pdl(sc_eval_fact_kernel<2, true>, grid, SC_THREADS, 0, s, fa));
        for (int i = 0; i < ntab; ++i) cur[i] = a.out[i];
      }
    } else if (round == 0) {
      if (NP == 1) CUDA_TRY(launch_pdl(sc_eval_round_kernel<1, false, false>, grid, SC_THREADS, 0, s, a));
      else CUDA_TRY(launch_pdl(sc_eval_round_kernel<2, false, false>, grid, SC_THREADS, 0, s, a));
    } else if (NP == 2 && T >= 4 && a.pairs >= 2048) {
      int lg = 0;
      while (((size_t)1 << lg) < 4 * (size_t)a.pairs) ++lg;
      rc = fix_var(c, a.eq_in, lg, &c->d_sc->r, a.eq_out);
      if (rc) return rc;
      CUDA_TRY(launch_pdl(sc_eval_round_kernel<2, true, true>, grid, SC_THREADS, 0, s, a));
      for (int i = 0; i <= ntab; ++i) cur[i] = (i < ntab) ? a.out[i] : a.eq_out;
    } else {
      if (NP == 1) CUDA_TRY(launch_pdl(sc_eval_round_kernel<1, true, false>, grid, SC_THREADS, 0, s, a));
      else CUDA_TRY(launch_pdl(sc_eval_round_kernel<2, true, false>, grid, SC_THREADS, 0, s, a));
      for (int i = 0; i <= ntab; ++i) cur[i] = (i < ntab) ? a.out[i] : a.eq_out;
    }
    prof_end(c, pi);
    count_launch(c);
  }
  if (job.carry) {  // hand over: tables as the next round would read them, last challenge pending in d_sc->r
    const size_t len = nrounds == 1 ? N : (N >> (nrounds - 1));
    if (fact) {  // materialise the eq table in the same (pre-bind) state: c_(R-1) * E_(R-2); R = 1: scale * eq(y, .)
      Fr* eq_pre = nullptr;
      CUDA_TRY(mem.alloc(&eq_pre, len * sizeof(Fr)));
      if (nrounds == 1) {
        rc = eq_build(c, job.eq_point, n, eq_pre);
        if (rc) return rc;
        if (job.eq_scale) {
          rc = fr_scale(c, eq_pre, N, job.eq_scale);
          if (rc) return rc;
        }
      } else {
        int blocks = (int)((len + 255) / 256);
        if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
        CUDA_TRY(launch_pdl(scale_copy_kernel, dim3(blocks), dim3(256), 0, s, (const Fr*)(suf + ((size_t)1 << (n + 1 - nrounds))),
                            (const Fr*)&c->d_sc->eqc, eq_pre, len));
        count_launch(c);
      }
      cur[ntab] = eq_pre;
    }
    for (int i = 0; i <= ntab; ++i) job.carry->cur[i] = cur[i];
    job.carry->len = len;
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
  }
  // final bind -> evals (eq excluded: ProverState::into_evals returns the polys only, classic.rs:143-149)
  if (!tail_done) {
    const Fr** d_ptrs = nullptr;
    const int nfinal = ntab + (job.want_eq_eval ? 1 : 0);
    CUDA_TRY(own.alloc(&d_ptrs, nfinal * sizeof(Fr*)));
    CUDA_TRY(cudaMemcpyAsync(d_ptrs, cur, nfinal * sizeof(Fr*), cudaMemcpyHostToDevice, s));
    CUDA_TRY(launch_pdl(sc_final_bind_kernel, dim3((nfinal + 63) / 64), dim3(64), 0, s, d_ptrs, nfinal, c->d_sc, job.evals_out));
    count_launch(c);
  }
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// ---------------------------------------------------------------------------------------------
// CoefficientsProver: per product k, c0 = Σ l0*r0, c2 = Σ (l1-l0)(r1-r0) with l = eq_k, r = P_k;
// message (c0, c1, c2) with c1 = sum - 2 c0 - c2 (coeff.rs:136-149); next claim by Horner (:36-38).
// ---------------------------------------------------------------------------------------------
struct ScCoeffArgs {
  const Fr* eq_in[SC_MAX_TERMS];
  Fr* eq_out[SC_MAX_TERMS];
  const Fr* in[SC_MAX_TERMS];
  Fr* out[SC_MAX_TERMS];
  const Fr* scalars;
  ScState* st;
  Fr* partial;
  Transcript* tr;
  Fr* challenges_out;
  uint32_t pairs;
  int round;
  PeerCtx peer;  // world > 1: (c0, c2) are summed over all ranks through the peer mailboxes
  unsigned int seq;
};

template <bool BIND>
__global__ void __launch_bounds__(SC_THREADS) sc_coeff_round_kernel(ScCoeffArgs a) {
  __shared__ Fr smem[(SC_THREADS / 32) * 2];
  pdl_prologue();
  const int k = blockIdx.y;
  Fr acc[2] = {fe_zero<FrP>(), fe_zero<FrP>()};
  Fr r = fe_zero<FrP>();
  if (BIND) r = fe_ld(&a.st->r);
  const Fr* __restrict__ eq_in = a.eq_in[k];
  Fr* __restrict__ eq_out = a.eq_out[k];
  const Fr* __restrict__ in = a.in[k];
  Fr* __restrict__ out = a.out[k];
  for (uint32_t b = blockIdx.x * blockDim.x + threadIdx.x; b < a.pairs; b += gridDim.x * blockDim.x) {
    Fr l0, l1, r0, r1;
    load_pair<BIND>(eq_in, eq_out, b, r, true, l0, l1);
    load_pair<BIND>(in, out, b, r, true, r0, r1);
    acc[0] = acc[0] + l0 * r0;
    acc[1] = acc[1] + (l1 - l0) * (r1 - r0);
  }
  block_reduce_fr<2>(acc, smem);
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    const Fr t0 = fr_bcast(acc[0], 0), t1 = fr_bcast(acc[1], 0);
    const Fr v = fr_mul_ni(lane == 0 ? t0 : t1, fe_ld(a.scalars + k));
    if (lane < 2) fe_st(a.partial + ((size_t)k * gridDim.x + blockIdx.x) * 2 + lane, v);
  }
  if (!last_cta_ticket(&a.st->counter)) return;
  const uint32_t nparts = gridDim.x * gridDim.y;
  acc[0] = fe_zero<FrP>();
  acc[1] = fe_zero<FrP>();
  const bool keep_busy = a.peer.world > 1;
  if (nparts <= 32) {
    if (threadIdx.x >= 32) {
      if (!keep_busy) return;
    } else if (threadIdx.x < nparts) {
      acc[0] = fr_ld_cg(a.partial + (size_t)threadIdx.x * 2);
      acc[1] = fr_ld_cg(a.partial + (size_t)threadIdx.x * 2 + 1);
    }
    if (threadIdx.x < 32) warp_reduce_fr<2>(acc);
  } else {
    for (uint32_t i = threadIdx.x; i < nparts; i += blockDim.x) {
      acc[0] = acc[0] + fr_ld_cg(a.partial + (size_t)i * 2);
      acc[1] = acc[1] + fr_ld_cg(a.partial + (size_t)i * 2 + 1);
    }
    block_reduce_fr<2>(acc, smem);
  }
  __shared__ volatile int s_busy;
  if (keep_busy) {  // warps 1-3 keep the SM busy while warp 0 runs the exchange and the finalize (peer.cuh)
    if (threadIdx.x == 0) s_busy = 1;
    __syncthreads();
    if (threadIdx.x >= 128) return;
    if (threadIdx.x >= 32) {
      peer_spin_while(&s_busy, &a.st->pad[0]);
      return;
    }
  }
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    __shared__ Transcript sh_tr;
    trw_copy(&sh_tr, a.tr);
    const Fr claim = fe_ld(&a.st->claim);
    Fr c0 = fr_bcast(acc[0], 0), c2 = fr_bcast(acc[1], 0);
    if (a.peer.world > 1) {  // fused collective: all-gather (c0, c2) over NVLink and add them
      peer_publish(a.peer, a.seq, lane == 0 ? c0 : c2, 2);
      Fr sum = fe_zero<FrP>();
      if (lane < 2)
        for (int r = 0; r < a.peer.world; ++r) sum = sum + peer_read(a.peer, a.seq, r, lane);
      c0 = fr_bcast(sum, 0);
      c2 = fr_bcast(sum, 1);
    }
    const Fr c1 = claim - (c0 + c0 + c2);  // coeff.rs:147
    const Fr canon = fr_canon_ni(lane == 0 ? c0 : (lane == 1 ? c1 : c2));
    for (int x = 0; x < 3; ++x) trw_write_canon_from_lane(&sh_tr, canon, x, true);
    const Fr ch = trw_squeeze(&sh_tr);
    const Fr next = fr_mul_ni(fr_mul_ni(c2, ch) + c1, ch) + c0;  // horner (coeff.rs:36-38)
    trw_copy(a.tr, &sh_tr);
    if (lane == 0) {
      fe_st(a.challenges_out + a.round, ch);
      fe_st(&a.st->r, ch);
      fe_st(&a.st->claim, next);
    }
    if (keep_busy && lane == 0) s_busy = 0;
  }
}

int sumcheck_prove_coeffs(Ctx* c, const ScCoeffJob& job) {
  const int n = job.num_vars, K = job.K;
  if (n < 1 || n > 30 || K < 1 || K > SC_MAX_TERMS) return B200_ERR_ARG;
  if ((job.stop_after > 0) != (job.carry != nullptr) || job.stop_after > n) return B200_ERR_ARG;
  NvtxRange nvtx("sum_check_prove-%d-2", n);
  cudaStream_t s = c->stream;
  const size_t N = (size_t)1 << n;
  const size_t szA = N / 2, szB = N / 4 ? N / 4 : 1;
  Fr *eq0 = nullptr, *bufA = nullptr, *bufB = nullptr;
  DevScope own(s);
  DevScope& mem = job.carry ? *job.carry->scope : own;
  CUDA_TRY(mem.alloc(&bufA, (size_t)2 * K * szA * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&bufB, (size_t)2 * K * szB * sizeof(Fr)));
  if (!job.eq_tables[0]) {
    CUDA_TRY(mem.alloc(&eq0, (size_t)K * N * sizeof(Fr)));
    for (int k = 0; k < K; ++k) {
      int rc = eq_build(c, job.eq_points[k], n, eq0 + (size_t)k * N);
      if (rc) return rc;
    }
  }
  CUDA_TRY(launch_pdl(sc_init_kernel, dim3(1), dim3(32), 0, s, c->d_sc, job.claim));
  count_launch(c);
  ScCoeffArgs a;
  a.scalars = job.scalars;
  a.st = c->d_sc;
  a.partial = c->d_partial;
  a.tr = c->d_tr;
  a.challenges_out = job.challenges_out;
  a.peer = c->peer;
  if (!job.sharded) a.peer.world = 1;
  a.seq = 0;
  const Fr* cur[2 * SC_MAX_TERMS];  // [k] poly, [K + k] eq
  for (int k = 0; k < K; ++k) {
    cur[k] = job.tables[k];
    cur[K + k] = job.eq_tables[0] ? job.eq_tables[k] : eq0 + (size_t)k * N;
  }
  const int nrounds = job.stop_after > 0 ? job.stop_after : n;
  for (int round = 0; round < nrounds; ++round) {
    a.round = round;
    a.pairs = (uint32_t)(N >> (round + 1));
    if (a.peer.world > 1) a.seq = ++c->peer_seq;
    Fr* dst_base = (round & 1) ? bufA : bufB;
    const size_t dst_sz = (round & 1) ? szA : szB;
    for (int k = 0; k < K; ++k) {
      a.in[k] = cur[k];
      a.eq_in[k] = cur[K + k];
      a.out[k] = dst_base + (size_t)k * dst_sz;
      a.eq_out[k] = dst_base + (size_t)(K + k) * dst_sz;
    }
    dim3 grid(blocks_for(a.pairs, K, g_occ[OCC_COEFF + (round > 0 ? 1 : 0)]), K);
    if ((size_t)grid.x * grid.y * 2 > c->partial_elems) return B200_ERR_NOMEM;
    if (round == 0) {
      CUDA_TRY(launch_pdl(sc_coeff_round_kernel<false>, grid, SC_THREADS, 0, s, a));
    } else {
      CUDA_TRY(launch_pdl(sc_coeff_round_kernel<true>, grid, SC_THREADS, 0, s, a));
      for (int k = 0; k < K; ++k) {
        cur[k] = a.out[k];
        cur[K + k] = a.eq_out[k];
      }
    }
    count_launch(c);
  }
  if (job.carry) {
    for (int i = 0; i < 2 * K; ++i) job.carry->cur[i] = cur[i];
    job.carry->len = nrounds == 1 ? N : (N >> (nrounds - 1));
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
  }
  const Fr** d_ptrs = nullptr;
  CUDA_TRY(own.alloc(&d_ptrs, K * sizeof(Fr*)));
  CUDA_TRY(cudaMemcpyAsync(d_ptrs, cur, K * sizeof(Fr*), cudaMemcpyHostToDevice, s));
  CUDA_TRY(launch_pdl(sc_final_bind_kernel, dim3(1), dim3(64), 0, s, d_ptrs, K, c->d_sc, job.evals_out));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_sumcheck() {
  auto occ_of = [](const void* k) {
    int o = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, k, SC_THREADS, 0) != cudaSuccess || o < 1) o = 2;
    return o;
  };
  g_occ[OCC_EVAL + 0] = occ_of((const void*)sc_eval_round_kernel<1, false, false>);
  g_occ[OCC_EVAL + 1] = occ_of((const void*)sc_eval_round_kernel<1, true, false>);
  g_occ[OCC_EVAL + 2] = occ_of((const void*)sc_eval_round_kernel<2, false, false>);
  g_occ[OCC_EVAL + 3] = occ_of((const void*)sc_eval_round_kernel<2, true, false>);  // <2, true, true> has the same footprint
  g_occ[OCC_FACT + 0] = occ_of((const void*)sc_eval_fact_kernel<1, false>);
  g_occ[OCC_FACT + 1] = occ_of((const void*)sc_eval_fact_kernel<1, true>);
  g_occ[OCC_FACT + 2] = occ_of((const void*)sc_eval_fact_kernel<2, false>);
  g_occ[OCC_FACT + 3] = occ_of((const void*)sc_eval_fact_kernel<2, true>);
  g_occ[OCC_COEFF + 0] = occ_of((const void*)sc_coeff_round_kernel<false>);
  g_occ[OCC_COEFF + 1] = occ_of((const void*)sc_coeff_round_kernel<true>);
  cudaGetLastError();
  B200_PRELOAD(sc_eval_round_kernel<1, false, false>);
  B200_PRELOAD(sc_eval_round_kernel<2, false, false>);
  B200_PRELOAD(sc_eval_round_kernel<1, true, false>);
  B200_PRELOAD(sc_eval_round_kernel<2, true, false>);
  B200_PRELOAD(sc_eval_round_kernel<2, true, true>);
  B200_PRELOAD(sc_eval_fact_kernel<1, false>);
  B200_PRELOAD(sc_eval_fact_kernel<2, false>);
  B200_PRELOAD(sc_eval_fact_kernel<1, true>);
  B200_PRELOAD(sc_eval_fact_kernel<2, true>);
  B200_PRELOAD(eq_suffix_small_kernel);
  B200_PRELOAD(eq_suffix_big_kernel);
  B200_PRELOAD(scale_copy_kernel);
  B200_PRELOAD(sc_eval_tail_kernel<1>);
  B200_PRELOAD(sc_eval_tail_kernel<2>);
  B200_PRELOAD(sc_final_bind_kernel);
  B200_PRELOAD(sc_init_kernel);
  B200_PRELOAD(sc_coeff_round_kernel<false>);
  B200_PRELOAD(sc_coeff_round_kernel<true>);
}

}  // namespace b200

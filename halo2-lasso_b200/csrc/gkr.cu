// GKR for fractional sum-checks on the GPU: prove_fractional_sum_check of
// pb/piop/gkr/fractional_sum_check.rs:87-190 (Layer::bottom / Layer::up :41-85, sum_check_expression :267-277,
// sum_check_claim :279-284, layer_down_claim :290-296).
//
// Layout: per batch element two heaps (numerators p, denominators q): level v (2^v entries) at [2^v, 2^(v+1)), level
// num_vars is the caller's table itself (borrowed, never written). The reference's Layer with v variables is the pair
// of halves of level v + 1 — contiguous sub-arrays, so every per-layer sum-check runs on the heaps without a copy.
// The per-layer sum-check  eq(x, y) * Σ_b [γ^(2b) (p_l q_r + p_r q_l) + γ^(2b+1) q_l q_r]  is the EVAL shape of
// sumcheck.cu with three (weight, table pair) terms per batch element; all Fiat-Shamir steps (claims, γ, the 4·batch
// evaluations, μ) run in single-warp kernels on the device transcript, so a whole argument is enqueued without a host
// round trip.
#include "internal.h"

namespace b200 {

static const int FRAC_MAX_BATCH = 10;  // 3 terms per element <= SC_MAX_TERMS

struct FracState {
  Fr cp[FRAC_MAX_BATCH], cq[FRAC_MAX_BATCH];  // running claims
  Fr weights[3 * FRAC_MAX_BATCH];
  Fr claim;
  Fr y[32];
  Fr sc_evals[6 * FRAC_MAX_BATCH];  // sum-check output, term order: p_l q_r | p_r q_l | q_l q_r
  Fr evals[4 * FRAC_MAX_BATCH];     // p_l p_r q_l q_r per element (the order the reference writes, :165)
};
struct FracTabs {
  const Fr* p[FRAC_MAX_BATCH];  // level v + 1 of the p / q heaps for the current layer
  const Fr* q[FRAC_MAX_BATCH];
  int B;
};

// Layer::up (:62-85) for all batch elements: blockIdx.y = element
__global__ void __launch_bounds__(256) frac_up_kernel(FracTabs child, Fr* const* __restrict__ p_heaps,
                                                      Fr* const* __restrict__ q_heaps, uint32_t half) {
  pdl_prologue();
  const int b = blockIdx.y;
  const Fr* __restrict__ pc = child.p[b];
  const Fr* __restrict__ qc = child.q[b];
  Fr* __restrict__ po = p_heaps[b] + half;
  Fr* __restrict__ qo = q_heaps[b] + half;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    const Fr pl = fe_ld(pc + i), pr = fe_ld(pc + half + i), ql = fe_ld(qc + i), qr = fe_ld(qc + half + i);
    fe_st(po + i, pl * qr + pr * ql);
    fe_st(qo + i, ql * qr);
  }
}

// layer-0 values: absorbed when claimed (Some), written otherwise (:121-146); p's first, then q's
__global__ void frac_roots_kernel(Transcript* tr, Fr* const* p_heaps, Fr* const* q_heaps, int B, uint32_t claimed_mask,
                                  FracState* st, Fr* out_p0, Fr* out_q0) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  trw_copy(&sh_tr, tr);
  for (int pass = 0; pass < 2; ++pass)
    for (int b = 0; b < B; ++b) {
      const Fr v = fe_ld((pass ? q_heaps[b] : p_heaps[b]) + 1);
      if (threadIdx.x == 0) {
        fe_st(pass ? &st->cq[b] : &st->cp[b], v);
        fe_st(pass ? out_q0 + b : out_p0 + b, v);
      }
      if ((claimed_mask >> (pass * 16 + b)) & 1) trw_common_fe(&sh_tr, v);
      else trw_write_fe(&sh_tr, v);
    }
  trw_copy(tr, &sh_tr);
}

// before the sum-check of the layer with v variables. v == 0: the four single entries are the evaluations.
// v > 0: gamma, weights over the terms (p_l q_r, p_r q_l share gamma^(2b); q_l q_r has gamma^(2b+1)), claim (:279-284)
__global__ void frac_before_kernel(Transcript* tr, FracTabs tabs, int v, FracState* st) {
  pdl_prologue();
  const int lane = threadIdx.x;
  if (v == 0) {
    for (int b = 0; b < tabs.B; ++b)
      if (lane < 4) fe_st(&st->evals[4 * b + lane], fe_ld((lane < 2 ? tabs.p[b] : tabs.q[b]) + (lane & 1)));
    return;
  }
  __shared__ Transcript sh_tr;
  trw_copy(&sh_tr, tr);
  const Fr gamma = trw_squeeze(&sh_tr);
  Fr pw = fe_one<FrP>(), claim = fe_zero<FrP>();
  for (int b = 0; b < tabs.B; ++b) {
    if (lane == 0) {
      fe_st(&st->weights[3 * b], pw);
      fe_st(&st->weights[3 * b + 1], pw);
    }
    claim = claim + fr_mul_ni(pw, fe_ld(&st->cp[b]));
    pw = fr_mul_ni(pw, gamma);
    if (lane == 0) fe_st(&st->weights[3 * b + 2], pw);
    claim = claim + fr_mul_ni(pw, fe_ld(&st->cq[b]));
    pw = fr_mul_ni(pw, gamma);
  }
  trw_copy(tr, &sh_tr);
  if (lane == 0) fe_st(&st->claim, claim);
}

// after the sum-check: write the 4B evaluations (:165), squeeze mu, layer_down_claim (:290-296), y = x || mu
__global__ void frac_after_kernel(Transcript* tr, int B, int v, const Fr* x /* v challenges */, FracState* st) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  const int lane = threadIdx.x;
  trw_copy(&sh_tr, tr);
  if (v > 0) {  // sum-check order (p_l, q_r, p_r, q_l, q_l, q_r) -> (p_l, p_r, q_l, q_r)
    for (int i = lane; i < 4 * B; i += 32) {
      const int b = i >> 2, k = i & 3;
      const int src = k == 0 ? 0 : (k == 1 ? 2 : (k == 2 ? 3 : 1));
      fe_st(&st->evals[i], fe_ld(&st->sc_evals[6 * b + src]));
    }
    __syncwarp();
  }
  for (int base = 0; base < 4 * B; base += 32) {
    const int i = base + lane;
    const Fr canon = fr_canon_ni(i < 4 * B ? fe_ld(&st->evals[i]) : fe_zero<FrP>());
    const int cnt = 4 * B - base < 32 ? 4 * B - base : 32;
    for (int j = 0; j < cnt; ++j) trw_write_canon_from_lane(&sh_tr, canon, j, true);
  }
  const Fr mu = trw_squeeze(&sh_tr);
  if (lane < 2 * B) {  // lane 2b: p claim, lane 2b + 1: q claim
    const int b = lane >> 1, o = 4 * b + 2 * (lane & 1);
    const Fr l = fe_ld(&st->evals[o]), r = fe_ld(&st->evals[o + 1]);
    fe_st((lane & 1) ? &st->cq[b] : &st->cp[b], l + (r - l) * mu);
  }
  for (int i = lane; i < v; i += 32) fe_st(&st->y[i], fe_ld(x + i));
  trw_copy(tr, &sh_tr);
  if (lane == 0) fe_st(&st->y[v], mu);
}

__global__ void frac_finish_kernel(const FracState* st, int B, int n, Fr* out_pxs, Fr* out_qxs, Fr* out_x) {
  pdl_prologue();
  const int i = threadIdx.x;
  if (i < B) {
    fe_st(out_pxs + i, fe_ld(&st->cp[i]));
    fe_st(out_qxs + i, fe_ld(&st->cq[i]));
  }
  if (i < n) fe_st(out_x + i, fe_ld(&st->y[i]));
}

// d_out: p_xs[B] | q_xs[B] | x[n] | p_0s[B] | q_0s[B]
int fractional_sum_check_prove(Ctx* c, int B, int n, const Fr* const* d_ps, const Fr* const* d_qs, uint32_t claimed_mask,
                               Fr* d_out) {
  if (B < 1 || B > FRAC_MAX_BATCH || n < 1 || n > 28) return B200_ERR_ARG;
  NvtxRange nvtx("fractional_sum_check-%d x%d", n, B);
  cudaStream_t s = c->stream;
  DevScope mem(s);
  const size_t N = (size_t)1 << n;
  Fr *heaps = nullptr, *x_scratch = nullptr;
  FracState* st = nullptr;
  Fr** d_heap_ptrs = nullptr;
  CUDA_TRY(mem.alloc(&heaps, (size_t)2 * B * N * sizeof(Fr)));  // levels 0..n-1 of p and q per element
  CUDA_TRY(mem.alloc(&x_scratch, 32 * sizeof(Fr)));
  CUDA_TRY(mem.alloc(&st, sizeof(FracState)));
  CUDA_TRY(mem.alloc(&d_heap_ptrs, (size_t)2 * B * sizeof(Fr*)));
  Fr* h_heap_ptrs[2 * FRAC_MAX_BATCH];
  for (int b = 0; b < B; ++b) {
    h_heap_ptrs[b] = heaps + (size_t)b * N;
    h_heap_ptrs[B + b] = heaps + (size_t)(B + b) * N;
  }
  CUDA_TRY(cudaMemcpyAsync(d_heap_ptrs, h_heap_ptrs, (size_t)2 * B * sizeof(Fr*), cudaMemcpyHostToDevice, s));
  auto level = [&](int b, bool q, int v) -> const Fr* {  // level v of element b (v == n: the caller's table)
    if (v == n) return q ? d_qs[b] : d_ps[b];
    return h_heap_ptrs[(q ? B : 0) + b] + ((size_t)1 << v);
  };
  auto tabs_of = [&](int v) {  // the Layer with v variables = halves of level v + 1
    FracTabs t;
    t.B = B;
    for (int b = 0; b < B; ++b) {
      t.p[b] = level(b, false, v + 1);
      t.q[b] = level(b, true, v + 1);
    }
    return t;
  };
  for (int v = n - 1; v >= 0; --v) {
    const uint32_t half = 1u << v;
    int bx = (int)((half + 255) / 256), cap = (4 * NUM_SMS + B - 1) / B;
    if (bx > cap) bx = cap;
    CUDA_TRY(launch_pdl(frac_up_kernel, dim3(bx, B), dim3(256), 0, s, tabs_of(v), (Fr* const*)d_heap_ptrs,
                        (Fr* const*)(d_heap_ptrs + B), half));
    count_launch(c);
  }
  Fr *out_pxs = d_out, *out_qxs = d_out + B, *out_x = d_out + 2 * B, *out_p0 = out_x + n, *out_q0 = out_p0 + B;
  CUDA_TRY(launch_pdl(frac_roots_kernel, dim3(1), dim3(32), 0, s, c->d_tr, (Fr* const*)d_heap_ptrs,
                      (Fr* const*)(d_heap_ptrs + B), B, claimed_mask, st, out_p0, out_q0));
  count_launch(c);
  for (int v = 0; v < n; ++v) {
    const FracTabs t = tabs_of(v);
    CUDA_TRY(launch_pdl(frac_before_kernel, dim3(1), dim3(32), 0, s, c->d_tr, t, v, st));
    count_launch(c);
    if (v > 0) {
      const size_t half = (size_t)1 << v;
      ScEvalJob job;
      job.num_vars = v;
      job.T = 3 * B;
      job.NP = 2;
      for (int b = 0; b < B; ++b) {
        const Fr *pl = t.p[b], *pr = t.p[b] + half, *ql = t.q[b], *qr = t.q[b] + half;
        job.tables[6 * b + 0] = pl;
        job.tables[6 * b + 1] = qr;
        job.tables[6 * b + 2] = pr;
        job.tables[6 * b + 3] = ql;
        job.tables[6 * b + 4] = ql;
        job.tables[6 * b + 5] = qr;
      }
      job.weights = st->weights;
      job.eq_point = st->y;
      job.claim = &st->claim;
      job.challenges_out = x_scratch;
      job.evals_out = st->sc_evals;
      int rc = sumcheck_prove_evals(c, job);
      if (rc) return rc;
    }
    CUDA_TRY(launch_pdl(frac_after_kernel, dim3(1), dim3(32), 0, s, c->d_tr, B, v, (const Fr*)x_scratch, st));
    count_launch(c);
  }
  CUDA_TRY(launch_pdl(frac_finish_kernel, dim3(1), dim3(64), 0, s, (const FracState*)st, B, n, out_pxs, out_qxs, out_x));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

void preload_gkr() {
  B200_PRELOAD(frac_up_kernel);
  B200_PRELOAD(frac_roots_kernel);
  B200_PRELOAD(frac_before_kernel);
  B200_PRELOAD(frac_after_kernel);
  B200_PRELOAD(frac_finish_kernel);
}

}  // namespace b200

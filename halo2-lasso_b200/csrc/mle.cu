// Multilinear-polynomial kernels (pb/poly/multilinear.rs) and the small transcript kernels.
//   eq_build      : MultilinearPolynomial::eq_xy (:91-127). The reference doubles level by level
//                   (n passes over memory); here eq[b] = A[b_lo] * B[b_hi] from two half tables that
//                   fit in L1/L2 — one write of the table, one multiplication per entry. Every entry
//                   is the same field element Π_k (b_k ? y_k : 1 - y_k), hence the same bytes.
//   fix_var       : fix_var / merge_into (:179-189, 599-618)
//   mle_eval_many : evaluate (:137-156) as the inner product <P, eq(., x)> for many tables at once
//   fr_lincomb    : the `+= (scalar, &poly)` / Sum algebra (:276-429) used by batch_open
//   quotient_step : one level of `quotients` (pb/pcs/multilinear.rs:72-107)
#include "internal.h"

namespace b200 {

// half tables: out[b] = Π_{k<nv} (b_k ? y[k] : 1 - y[k]) computed directly (nv <= 16)
__global__ void eq_direct_kernel(const Fr* __restrict__ y, int nv, Fr* __restrict__ out) {
  pdl_prologue();
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= (1u << nv)) return;
  const Fr one = fe_one<FrP>();
  Fr acc = one;
  for (int k = 0; k < nv; ++k) {
    const Fr yk = fe_ld(y + k);
    acc = acc * (((b >> k) & 1) ? yk : one - yk);
  }
  fe_st(out + b, acc);
}

__global__ void eq_combine_kernel(const Fr* __restrict__ lo, const Fr* __restrict__ hi, int nlo,
                                  size_t n, Fr* __restrict__ out) {
  pdl_prologue();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const uint32_t mask = (1u << nlo) - 1;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < n; b += stride) {
    const Fr l = fe_ld(lo + (b & mask));
    const Fr h = fe_ld(hi + (b >> nlo));
    fe_st(out + b, l * h);
  }
}

int eq_build(Ctx* c, const Fr* d_y, int n, Fr* d_out) {
  cudaStream_t s = c->stream;
  if (n < 1 || n > 30) return B200_ERR_ARG;
  if (n <= 12) {
    const uint32_t N = 1u << n;
    CUDA_TRY(launch_pdl(eq_direct_kernel, dim3((N + 127) / 128), dim3(128), 0, s, d_y, n, d_out));
    count_launch(c);
    CUDA_TRY(cudaGetLastError());
    return B200_OK;
  }
  const int nlo = n / 2, nhi = n - nlo;
  Fr* half = nullptr;
  CUDA_TRY(cudaMallocAsync(&half, (((size_t)1 << nlo) + ((size_t)1 << nhi)) * sizeof(Fr), s));
  Fr* lo = half;
  Fr* hi = half + ((size_t)1 << nlo);
  CUDA_TRY(launch_pdl(eq_direct_kernel, dim3(((1u << nlo) + 127) / 128), dim3(128), 0, s, d_y, nlo, lo));
  CUDA_TRY(launch_pdl(eq_direct_kernel, dim3(((1u << nhi) + 127) / 128), dim3(128), 0, s, d_y + nlo, nhi, hi));
  const size_t N = (size_t)1 << n;
  int blocks = (int)((N + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  CUDA_TRY(launch_pdl(eq_combine_kernel, dim3(blocks), dim3(256), 0, s, lo, hi, nlo, N, d_out));
  count_launch(c, 3);
  CUDA_TRY(cudaFreeAsync(half, s));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void fix_var_kernel(const Fr* __restrict__ in, size_t half, const Fr* __restrict__ r_ptr,
                               Fr* __restrict__ out) {
  pdl_prologue();
  const Fr r = fe_ld(r_ptr);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < half; b += stride) {
    const Fr x0 = fe_ldg(in + 2 * b), x1 = fe_ldg(in + 2 * b + 1);
    fe_st(out + b, (x1 - x0) * r + x0);
  }
}

int fix_var(Ctx* c, const Fr* d_in, int n, const Fr* d_r, Fr* d_out) {
  if (n < 1 || n > 31) return B200_ERR_ARG;
  const size_t half = (size_t)1 << (n - 1);
  int blocks = (int)((half + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  CUDA_TRY(launch_pdl(fix_var_kernel, dim3(blocks), dim3(256), 0, c->stream, d_in, half, d_r, d_out));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// out[t] = Σ_b tables[t][b] * eq[b]; grid.y = table; two-stage reduction through `partial`
struct EvalManyArgs {
  const Fr* tables[SC_MAX_TABLES];
  const Fr* eq;
  size_t n;
  Fr* partial;
  unsigned int* counter;
  Fr* out;
};
__global__ void __launch_bounds__(256) mle_dot_kernel(EvalManyArgs a) {
  pdl_prologue();
  __shared__ Fr smem[8];
  const int t = blockIdx.y;
  const Fr* __restrict__ tab = a.tables[t];
  Fr acc[1] = {fe_zero<FrP>()};
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.n; b += stride)
    acc[0] = acc[0] + fe_ldg(tab + b) * fe_ldg(a.eq + b);
  block_reduce_fr<1>(acc, smem);
  if (threadIdx.x == 0) fe_st(a.partial + (size_t)t * gridDim.x + blockIdx.x, acc[0]);
  if (!last_cta_ticket(a.counter)) return;
  for (int tt = 0; tt < (int)gridDim.y; ++tt) {
    acc[0] = fe_zero<FrP>();
    for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x)
      acc[0] = acc[0] + fr_ld_cg(a.partial + (size_t)tt * gridDim.x + i);
    block_reduce_fr<1>(acc, smem);
    if (threadIdx.x == 0) fe_st(a.out + tt, acc[0]);
  }
}

int mle_dot_many(Ctx* c, const Fr* const* h_tables, int ntables, size_t len, const Fr* d_eq, Fr* d_out) {
  if (ntables < 1 || ntables > SC_MAX_TABLES) return B200_ERR_ARG;
  EvalManyArgs a;
  for (int i = 0; i < ntables; ++i) a.tables[i] = h_tables[i];
  a.eq = d_eq;
  a.n = len;
  a.partial = c->d_partial;
  a.counter = &c->d_sc->counter;
  a.out = d_out;
  int bx = (int)((len + 255) / 256);
  int cap = (4 * NUM_SMS) / ntables;  // 4 CTAs of 256 threads per SM (62 registers), ONE wave: floor, not ceil — 25 x 24
  if (bx > cap) bx = cap;             // = 600 CTAs on 592 slots left 8 CTAs for a second wave
  if (bx < 1) bx = 1;
  if ((size_t)bx * ntables > c->partial_elems) return B200_ERR_NOMEM;
  CUDA_TRY(launch_pdl(mle_dot_kernel, dim3(dim3(bx, ntables)), dim3(256), 0, c->stream, a));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int mle_eval_many(Ctx* c, const Fr* const* h_tables, int ntables, int n, const Fr* d_point, Fr* d_out) {
  if (ntables < 1 || ntables > SC_MAX_TABLES) return B200_ERR_ARG;
  const size_t N = (size_t)1 << n;
  DevScope mem(c->stream);
  Fr* eq = nullptr;
  CUDA_TRY(mem.alloc(&eq, N * sizeof(Fr)));
  int rc = eq_build(c, d_point, n, eq);
  if (rc) return rc;
  return mle_dot_many(c, h_tables, ntables, N, eq, d_out);
}

struct LincombArgs {
  const Fr* tables[SC_MAX_TABLES];
  const Fr* scalars;
  int k;
  size_t len;
  Fr* out;
};
__global__ void __launch_bounds__(256) lincomb_kernel(LincombArgs a) {
  pdl_prologue();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.len; i += stride) {
    Fr acc = fe_zero<FrP>();
    for (int j = 0; j < a.k; ++j) acc = acc + fe_ldg(a.tables[j] + i) * fe_ld(a.scalars + j);
    fe_st(a.out + i, acc);
  }
}
int fr_lincomb(Ctx* c, const Fr* const* h_tables, int k, const Fr* d_scalars, size_t len, Fr* d_out) {
  if (k < 1 || k > SC_MAX_TABLES) return B200_ERR_ARG;
  LincombArgs a;
  for (int i = 0; i < k; ++i) a.tables[i] = h_tables[i];
  a.scalars = d_scalars;
  a.k = k;
  a.len = len;
  a.out = d_out;
  int blocks = (int)((len + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  CUDA_TRY(launch_pdl(lincomb_kernel, dim3(blocks), dim3(256), 0, c->stream, a));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void scale_kernel(Fr* __restrict__ tab, size_t n, const Fr* __restrict__ scalar) {
  const Fr sc = fe_ld(scalar);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) fe_st(tab + i, fe_ld(tab + i) * sc);
}
int fr_scale(Ctx* c, Fr* d_tab, size_t n, const Fr* d_scalar) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  scale_kernel<<<blocks, 256, 0, c->stream>>>(d_tab, n, d_scalar);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void convert_kernel(const Fr* __restrict__ in, Fr* __restrict__ out, size_t n, int to_mont) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const Fr x = fe_ldg(in + i);
    fe_st(out + i, to_mont ? fe_from_canonical<FrP>(x) : fe_to_canonical<FrP>(x));
  }
}
int fr_convert(Ctx* c, const Fr* d_in, Fr* d_out, size_t n, int to_mont) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  convert_kernel<<<blocks, 256, 0, c->stream>>>(d_in, d_out, n, to_mont);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void from_u64_kernel(const uint64_t* __restrict__ in, Fr* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    fe_st(out + i, fe_from_u64<FrP>(in[i]));
}
int fr_from_u64(Ctx* c, const uint64_t* d_in, Fr* d_out, size_t n) {
  int blocks = (int)((n + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  from_u64_kernel<<<blocks, 256, 0, c->stream>>>(d_in, d_out, n);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// q[i] = rem[half+i] - rem[i]; rem[i] += (rem[half+i] - rem[i]) * x      (top variable first)
__global__ void quotient_kernel(Fr* __restrict__ rem, size_t half, const Fr* __restrict__ x_ptr,
                                Fr* __restrict__ q) {
  pdl_prologue();
  const Fr x = fe_ld(x_ptr);
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < half; i += stride) {
    const Fr lo = fe_ld(rem + i), hi = fe_ld(rem + half + i);
    const Fr d = hi - lo;
    fe_st(q + i, d);
    fe_st(rem + i, lo + d * x);
  }
}
int quotient_step(Ctx* c, Fr* d_rem, size_t half, const Fr* d_x, Fr* d_q) {
  int blocks = (int)((half + 255) / 256);
  if (blocks > NUM_SMS * 8) blocks = NUM_SMS * 8;
  if (blocks < 1) blocks = 1;
  CUDA_TRY(launch_pdl(quotient_kernel, dim3(blocks), dim3(256), 0, c->stream, d_rem, half, d_x, d_q));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// Π_i (2 x_i y_i + 1 - x_i - y_i)
__global__ void eq_xy_eval_kernel(const Fr* x, const Fr* y, int n, Fr* out) {
  pdl_prologue();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const Fr one = fe_one<FrP>();
  Fr acc = one;
  for (int i = 0; i < n; ++i) {
    const Fr xi = fe_ld(x + i), yi = fe_ld(y + i);
    const Fr xy = xi * yi;
    acc = acc * (xy + xy + one - xi - yi);
  }
  fe_st(out, acc);
}
int eq_xy_eval_dev(Ctx* c, const Fr* d_x, const Fr* d_y, int n, Fr* d_out) {
  CUDA_TRY(launch_pdl(eq_xy_eval_kernel, dim3(1), dim3(32), 0, c->stream, d_x, d_y, n, d_out));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void transcript_kernel(Transcript* tr, int op, const Fr* in, Fr* out, int n) {
  pdl_prologue();
  __shared__ Transcript sh_tr;  // one warp, warp-cooperative Keccak
  trw_copy(&sh_tr, tr);
  for (int i = 0; i < n; ++i) {
    if (op == TR_COMMON) trw_common_fe(&sh_tr, fe_ld(in + i));
    else if (op == TR_WRITE) trw_write_fe(&sh_tr, fe_ld(in + i));
    else {
      const Fr ch = trw_squeeze(&sh_tr);
      if (threadIdx.x == 0) fe_st(out + i, ch);
    }
  }
  trw_copy(tr, &sh_tr);
}
int transcript_op(Ctx* c, int op, const Fr* d_in, Fr* d_out, int n) {
  if (n <= 0) return B200_OK;
  CUDA_TRY(launch_pdl(transcript_kernel, dim3(1), dim3(32), 0, c->stream, c->d_tr, op, d_in, d_out, n));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

__global__ void transcript_points_kernel(Transcript* tr, const G1Aff* pts, int n) {
  pdl_prologue();
  __shared__ Transcript sh_tr;
  trw_copy(&sh_tr, tr);
  for (int i = 0; i < n; ++i) {
    const Fq x = fe_ld(&pts[i].x), y = fe_ld(&pts[i].y);
    trw_write_commitment(&sh_tr, x, y);
  }
  trw_copy(tr, &sh_tr);
}
int transcript_write_points(Ctx* c, const G1Aff* d_pts, int n) {
  if (n <= 0) return B200_OK;
  CUDA_TRY(launch_pdl(transcript_points_kernel, dim3(1), dim3(32), 0, c->stream, c->d_tr, d_pts, n));
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_mle() {
  B200_PRELOAD(eq_direct_kernel);
  B200_PRELOAD(eq_combine_kernel);
  B200_PRELOAD(fix_var_kernel);
  B200_PRELOAD(mle_dot_kernel);
  B200_PRELOAD(lincomb_kernel);
  B200_PRELOAD(scale_kernel);
  B200_PRELOAD(convert_kernel);
  B200_PRELOAD(from_u64_kernel);
  B200_PRELOAD(quotient_kernel);
  B200_PRELOAD(eq_xy_eval_kernel);
  B200_PRELOAD(transcript_kernel);
  B200_PRELOAD(transcript_points_kernel);
}

}  // namespace b200

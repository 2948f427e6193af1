// permutation_z_polys (pb/backend/hyperplonk/prover.rs:252-345) on the GPU, for one chunk (one z polynomial):
//   product[b] = Π_i (w_i[b] + beta*(id_off_i + b) + gamma) / Π_i (w_i[b] + beta*sigma_i[b] + gamma)
//   z[bh[0]] = 0, z[bh[1]] = 1, z[bh[k]] = Π_{1<=j<k} product[bh[j]]          (bh = BooleanHypercube LFSR order)
// The reference batch-inverts the denominators per rayon chunk and then runs a SERIAL prefix product over the
// 2^n rows in LFSR order (prover.rs:308-323). Here: (1) per-thread Montgomery batch inversion over 32 rows,
// (2) a three-pass chunked scan in LFSR order — a chunk's first row is x^(start-1) in GF(2^n), obtained with a
// carry-less square-and-multiply instead of walking the register.
#include "internal.h"

namespace b200 {

static const int PERM_MAX_POLYS = 8;
static const int INV_CHUNK = 32;
static const int SCAN_CHUNK = 256;

struct PermArgs {
  const Fr* wires[PERM_MAX_POLYS];
  const Fr* sigmas[PERM_MAX_POLYS];
  uint64_t id_offset[PERM_MAX_POLYS];
  int npolys, num_vars;
  const Fr* beta_gamma;  // device: beta, gamma
  Fr* products;
};

__global__ void __launch_bounds__(128) perm_products_kernel(PermArgs a) {
  const size_t N = (size_t)1 << a.num_vars;
  const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * INV_CHUNK;
  if (base >= N) return;
  const Fr beta = fe_ld(a.beta_gamma), gamma = fe_ld(a.beta_gamma + 1);
  Fr den[INV_CHUNK], pre[INV_CHUNK];
  Fr run = fe_one<FrP>();
  const int cnt = (int)((N - base) < INV_CHUNK ? (N - base) : INV_CHUNK);
  for (int k = 0; k < cnt; ++k) {
    Fr d = fe_one<FrP>();
    for (int i = 0; i < a.npolys; ++i) d = d * (beta * fe_ldg(a.sigmas[i] + base + k) + gamma + fe_ldg(a.wires[i] + base + k));
    den[k] = d;
    pre[k] = run;
    run = run * d;
  }
  Fr inv = fe_inv<FrP>(run);  // denominators are non-zero with overwhelming probability (as in the reference)
  for (int k = cnt - 1; k >= 0; --k) {
    const Fr dinv = inv * pre[k];
    inv = inv * den[k];
    Fr num = fe_one<FrP>();
    for (int i = 0; i < a.npolys; ++i)
      num = num * (fe_from_u64<FrP>(a.id_offset[i] + base + k) * beta + gamma + fe_ldg(a.wires[i] + base + k));
    fe_st(a.products + base + k, num * dinv);
  }
}

// GF(2^n) arithmetic on the LFSR state (primitive polynomial of bh.rs)
__device__ __forceinline__ uint32_t gf_next(uint32_t b, int n, uint32_t prim) {
  uint64_t s = (uint64_t)b << 1;
  s ^= (s >> n) * prim;
  return (uint32_t)s;
}
__device__ __forceinline__ uint32_t gf_mul(uint32_t x, uint32_t y, int n, uint32_t prim) {
  uint32_t r = 0;
  for (int i = 0; i < n; ++i) {
    if ((y >> i) & 1) r ^= x;
    x = gf_next(x, n, prim);
  }
  return r;
}
__device__ __forceinline__ uint32_t gf_pow_x(uint64_t e, int n, uint32_t prim) {  // x^e
  uint32_t result = 1, base = 2 & ((1u << n) - 1);
  if (n == 1) base = gf_next(1, n, prim);
  while (e) {
    if (e & 1) result = gf_mul(result, base, n, prim);
    base = gf_mul(base, base, n, prim);
    e >>= 1;
  }
  return result;
}

// pass 1: product of `products` over the chunk's rows in LFSR order (positions start .. start+SCAN_CHUNK)
// With nz > 1 z polynomials (prover.rs:308-323) the running product visits (row, z index) in lexicographic order:
// products[zi * N + b] is the ratio of permutation chunk zi on row b.
__global__ void __launch_bounds__(128) perm_chunk_prod_kernel(const Fr* __restrict__ products, int n, uint32_t prim,
                                                              uint32_t nchunks, int nz, Fr* __restrict__ chunk_prod) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks) return;
  const uint64_t N = (uint64_t)1 << n, start = 1 + (uint64_t)c * SCAN_CHUNK;
  uint32_t b = gf_pow_x(start - 1, n, prim);
  Fr acc = fe_one<FrP>();
  for (uint64_t pos = start; pos < start + SCAN_CHUNK && pos < N; ++pos) {
    for (int zi = 0; zi < nz; ++zi) acc = acc * fe_ldg(products + (size_t)zi * N + b);
    b = gf_next(b, n, prim);
  }
  fe_st(chunk_prod + c, acc);
}
// pass 2: exclusive prefix products of the chunk products (single CTA)
__global__ void __launch_bounds__(1024) perm_scan_kernel(Fr* chunk_prod, uint32_t nchunks) {
  __shared__ Fr sh[1024];
  const uint32_t per = (nchunks + 1023) / 1024, lo = threadIdx.x * per;
  Fr acc = fe_one<FrP>();
  for (uint32_t i = lo; i < lo + per && i < nchunks; ++i) acc = acc * fe_ld(chunk_prod + i);
  sh[threadIdx.x] = acc;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {  // inclusive Hillis-Steele
    Fr y = fe_one<FrP>();
    if ((int)threadIdx.x >= off) y = sh[threadIdx.x - off];
    __syncthreads();
    if ((int)threadIdx.x >= off) sh[threadIdx.x] = sh[threadIdx.x] * y;
    __syncthreads();
  }
  Fr run = threadIdx.x ? sh[threadIdx.x - 1] : fe_one<FrP>();
  for (uint32_t i = lo; i < lo + per && i < nchunks; ++i) {
    const Fr v = fe_ld(chunk_prod + i);
    fe_st(chunk_prod + i, run);
    run = run * v;
  }
}
// pass 3: z_zi[bh[pos]] = running product before (row pos, zi)
struct PermZ {
  Fr* z[PERM_MAX_POLYS];
};
__global__ void __launch_bounds__(128) perm_write_kernel(const Fr* __restrict__ products, int n, uint32_t prim,
                                                         uint32_t nchunks, int nz, const Fr* __restrict__ chunk_excl,
                                                         PermZ out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0)
    for (int zi = 0; zi < nz; ++zi) fe_st(out.z[zi], fe_zero<FrP>());  // row 0 (position 0)
  if (c >= nchunks) return;
  const uint64_t N = (uint64_t)1 << n, start = 1 + (uint64_t)c * SCAN_CHUNK;
  uint32_t b = gf_pow_x(start - 1, n, prim);
  Fr run = fe_ld(chunk_excl + c);
  for (uint64_t pos = start; pos < start + SCAN_CHUNK && pos < N; ++pos) {
    for (int zi = 0; zi < nz; ++zi) {
      fe_st(out.z[zi] + b, run);
      run = run * fe_ldg(products + (size_t)zi * N + b);
    }
    b = gf_next(b, n, prim);
  }
}

// nz z polynomials; permutation chunk zi covers the wire columns [zi * chunk_size, min(npolys, (zi+1) * chunk_size))
int permutation_z_chunks(Ctx* c, int num_vars, int nz, int chunk_size, int npolys, const Fr* const* wires,
                         const Fr* const* sigmas, const uint64_t* id_offsets, const Fr* d_beta_gamma, Fr* const* d_z) {
  static const uint32_t PRIM[32] = {1, 3, 7, 11, 19, 37, 67, 131, 285, 529, 1033, 2053, 4179, 8219, 16427, 32771,
                                    65581, 131081, 262183, 524327, 1048585, 2097157, 4194307, 8388641, 16777243,
                                    33554441, 67108935, 134217767, 268435465, 536870917, 1073741907, 2147483657u};
  if (num_vars < 1 || num_vars > 30 || npolys < 1 || nz < 1 || nz > PERM_MAX_POLYS || chunk_size < 1 ||
      chunk_size > PERM_MAX_POLYS || (nz - 1) * chunk_size >= npolys || nz * chunk_size < npolys)
    return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  const size_t N = (size_t)1 << num_vars;
  const uint32_t nchunks = (uint32_t)((N - 1 + SCAN_CHUNK - 1) / SCAN_CHUNK);
  Fr *products = nullptr, *cp = nullptr;
  CUDA_TRY(cudaMallocAsync(&products, (size_t)nz * N * sizeof(Fr), s));
  CUDA_TRY(cudaMallocAsync(&cp, ((size_t)nchunks + 1) * sizeof(Fr), s));
  const size_t nthreads = (N + INV_CHUNK - 1) / INV_CHUNK;
  for (int zi = 0; zi < nz; ++zi) {
    PermArgs a;
    const int lo = zi * chunk_size, hi = lo + chunk_size < npolys ? lo + chunk_size : npolys;
    for (int i = lo; i < hi; ++i) {
      a.wires[i - lo] = wires[i];
      a.sigmas[i - lo] = sigmas[i];
      a.id_offset[i - lo] = id_offsets[i];
    }
    a.npolys = hi - lo;
    a.num_vars = num_vars;
    a.beta_gamma = d_beta_gamma;
    a.products = products + (size_t)zi * N;
    perm_products_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, s>>>(a);
  }
  PermZ out;
  for (int zi = 0; zi < nz; ++zi) out.z[zi] = d_z[zi];
  perm_chunk_prod_kernel<<<(nchunks + 127) / 128, 128, 0, s>>>(products, num_vars, PRIM[num_vars], nchunks, nz, cp);
  perm_scan_kernel<<<1, 1024, 0, s>>>(cp, nchunks);
  perm_write_kernel<<<(nchunks + 127) / 128, 128, 0, s>>>(products, num_vars, PRIM[num_vars], nchunks, nz, cp, out);
  count_launch(c, 3 + nz);
  CUDA_TRY(cudaFreeAsync(products, s));
  CUDA_TRY(cudaFreeAsync(cp, s));
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

int permutation_z(Ctx* c, int num_vars, int npolys, const Fr* const* wires, const Fr* const* sigmas,
                  const uint64_t* id_offsets, const Fr* d_beta_gamma, Fr* d_z) {
  if (npolys < 1 || npolys > PERM_MAX_POLYS) return B200_ERR_ARG;
  Fr* zs[1] = {d_z};
  return permutation_z_chunks(c, num_vars, 1, npolys, npolys, wires, sigmas, id_offsets, d_beta_gamma, zs);
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_perm() {
  B200_PRELOAD(perm_products_kernel);
  B200_PRELOAD(perm_chunk_prod_kernel);
  B200_PRELOAD(perm_scan_kernel);
  B200_PRELOAD(perm_write_kernel);
}

}  // namespace b200

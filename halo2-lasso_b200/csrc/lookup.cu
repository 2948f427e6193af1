// LogUp helper polynomials of the snapshot's own lookup argument (pb/backend/hyperplonk/prover.rs:50-250):
//   expression_rows   Expression::evaluate on every hypercube row (prover.rs:96-117) — the bytecode of
//                     expression.py::compile_expression interpreted once per row (leaves are dense tables, exactly as
//                     for the generic sum-check kernel); used for the compressed input / table polynomials
//                     Σ_j beta^j expr_j (lookup_compressed_poly, prover.rs:78-134)
//   lookup_m          multiplicities (lookup_m_poly, prover.rs:143-192): the reference builds a HashMap
//                     table value -> row and counts the inputs per row; here an open-addressing hash table of row
//                     indices in HBM (keys stay in the table polynomial), duplicates resolved to the LAST row as
//                     HashMap::from_iter does, counts by warp-aggregated atomics
//   lookup_h          h = 1/(gamma + input) - m/(gamma + table)   (lookup_h_poly, prover.rs:206-250) with the
//                     per-thread Montgomery batch inversion of perm.cu
#include <vector>

#include "../../include/b200_lasso.h"
#include "internal.h"

namespace b200 {

static const int ROWS_MAX_TABLES = 40;

struct RowsArgs {
  const Fr* in[ROWS_MAX_TABLES];
  const Fr* consts;
  const int4* ops;  // (opcode, dst, a, b)
  int K, C, nops;
  Fr* out;
  size_t N;
};

// slot file in shared memory, one column per thread: [K leaf values | T temporaries] x blockDim.x
__global__ void __launch_bounds__(128) expr_rows_kernel(RowsArgs a) {
  extern __shared__ __align__(32) unsigned char rows_smem_raw[];
  Fr* sm = reinterpret_cast<Fr*>(rows_smem_raw);
  const int K = a.K, KC = a.K + a.C, nth = blockDim.x, tid = threadIdx.x;
  const int last = a.ops[a.nops - 1].y;
  auto rd = [&](int idx) -> Fr {
    if (idx < K) return sm[idx * nth + tid];
    if (idx < KC) return fe_ld(a.consts + (idx - K));
    return sm[(K + idx - KC) * nth + tid];
  };
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < a.N; b += stride) {
    for (int k = 0; k < K; ++k) sm[k * nth + tid] = fe_ldg(a.in[k] + b);
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
    fe_st(a.out + b, sm[(K + last - KC) * nth + tid]);
  }
}

// ---- multiplicities ------------------------------------------------------------------------------------
static const uint32_t SLOT_EMPTY = 0xFFFFFFFFu;

__device__ __forceinline__ uint32_t fr_hash(const Fr& v) {
  uint32_t h = 0x9E3779B9u;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    h ^= v.v[i];
    h *= 0x85EBCA6Bu;
    h ^= h >> 15;
  }
  return h;
}
__device__ __forceinline__ bool fr_same(const Fr& a, const Fr& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) d |= a.v[i] ^ b.v[i];
  return d == 0;
}

// every table row claims a slot; rows holding the same value meet in one slot, the largest row index stays
__global__ void lookup_insert_kernel(const Fr* __restrict__ table, uint32_t n, uint32_t* slots, uint32_t mask) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Fr key = fe_ldg(table + i);
  uint32_t s = fr_hash(key) & mask;
  for (;;) {
    const uint32_t cur = atomicCAS(slots + s, SLOT_EMPTY, i);
    if (cur == SLOT_EMPTY) return;
    if (fr_same(fe_ldg(table + cur), key)) {
      atomicMax(slots + s, i);
      return;
    }
    s = (s + 1) & mask;
  }
}
__global__ void lookup_count_kernel(const Fr* __restrict__ input, const Fr* __restrict__ table, uint32_t n,
                                    const uint32_t* __restrict__ slots, uint32_t mask, uint32_t* counts, int* invalid) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t row = SLOT_EMPTY;
  if (i < n) {
    const Fr key = fe_ldg(input + i);
    uint32_t s = fr_hash(key) & mask;
    for (;;) {
      const uint32_t cur = slots[s];
      if (cur == SLOT_EMPTY) break;
      if (fr_same(fe_ldg(table + cur), key)) {
        row = cur;
        break;
      }
      s = (s + 1) & mask;
    }
    if (row == SLOT_EMPTY) *invalid = 1;  // Error::InvalidSnark("Invalid lookup input")
  }
  // one atomic per distinct row and warp (gated rows all hit the row of the zero tuple)
  const unsigned peers = __match_any_sync(__activemask(), row);
  if (row != SLOT_EMPTY && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(counts + row, __popc(peers));
}
__global__ void lookup_counts_to_fr_kernel(const uint32_t* __restrict__ counts, uint32_t n, Fr* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) fe_st(out + i, fe_from_u64<FrP>(counts[i]));
}

int lookup_m(Ctx* c, int num_vars, const Fr* d_input, const Fr* d_table, Fr* d_m) {
  if (num_vars < 1 || num_vars > 30) return B200_ERR_ARG;
  cudaStream_t s = c->stream;
  const uint32_t n = 1u << num_vars, cap = n << 1, mask = cap - 1;
  uint32_t *slots = nullptr, *counts = nullptr;
  int* flag = nullptr;
  CUDA_TRY(cudaMallocAsync(&slots, (size_t)cap * sizeof(uint32_t), s));
  CUDA_TRY(cudaMallocAsync(&counts, ((size_t)n + 1) * sizeof(uint32_t), s));
  flag = reinterpret_cast<int*>(counts + n);
  CUDA_TRY(cudaMemsetAsync(slots, 0xFF, (size_t)cap * sizeof(uint32_t), s));
  CUDA_TRY(cudaMemsetAsync(counts, 0, ((size_t)n + 1) * sizeof(uint32_t), s));
  const unsigned blocks = (n + 255) / 256;
  lookup_insert_kernel<<<blocks, 256, 0, s>>>(d_table, n, slots, mask);
  lookup_count_kernel<<<blocks, 256, 0, s>>>(d_input, d_table, n, slots, mask, counts, flag);
  lookup_counts_to_fr_kernel<<<blocks, 256, 0, s>>>(counts, n, d_m);
  count_launch(c, 3);
  int invalid = 0;
  CUDA_TRY(cudaMemcpyAsync(&invalid, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  CUDA_TRY(cudaFreeAsync(slots, s));
  CUDA_TRY(cudaFreeAsync(counts, s));
  CUDA_TRY(cudaGetLastError());
  return invalid ? B200_ERR_LOOKUP : B200_OK;
}

// ---- h polynomial --------------------------------------------------------------------------------------
static const int H_CHUNK = 16;
__global__ void __launch_bounds__(128) lookup_h_kernel(const Fr* __restrict__ input, const Fr* __restrict__ table,
                                                       const Fr* __restrict__ m, const Fr* gamma_p, size_t N, Fr* h) {
  const size_t base = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * H_CHUNK;
  if (base >= N) return;
  const Fr gamma = fe_ld(gamma_p);
  const int cnt = (int)((N - base) < H_CHUNK ? (N - base) : H_CHUNK);
  Fr den[2 * H_CHUNK], pre[2 * H_CHUNK];
  Fr run = fe_one<FrP>();
  for (int k = 0; k < cnt; ++k) {
    den[2 * k] = gamma + fe_ldg(input + base + k);
    den[2 * k + 1] = gamma + fe_ldg(table + base + k);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      pre[2 * k + j] = run;
      run = run * den[2 * k + j];
    }
  }
  Fr inv = fe_inv<FrP>(run);  // gamma + value is non-zero with overwhelming probability (as in the reference)
  for (int k = cnt - 1; k >= 0; --k) {
    const Fr tinv = inv * pre[2 * k + 1];
    inv = inv * den[2 * k + 1];
    const Fr iinv = inv * pre[2 * k];
    inv = inv * den[2 * k];
    fe_st(h + base + k, iinv - tinv * fe_ldg(m + base + k));
  }
}

int lookup_h(Ctx* c, int num_vars, const Fr* d_input, const Fr* d_table, const Fr* d_m, const Fr* d_gamma, Fr* d_h) {
  if (num_vars < 1 || num_vars > 30) return B200_ERR_ARG;
  const size_t N = (size_t)1 << num_vars;
  const size_t nthreads = (N + H_CHUNK - 1) / H_CHUNK;
  lookup_h_kernel<<<(unsigned)((nthreads + 127) / 128), 128, 0, c->stream>>>(d_input, d_table, d_m, d_gamma, N, d_h);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// run a compiled program (device constants / ops) on every row
int expression_rows_prog(Ctx* c, int num_vars, const Fr* const* tables, int ntables, const Fr* d_consts, int nconsts,
                         const int4* d_ops, int nops, int ntemps, Fr* d_out) {
  if (ntables < 1 || ntables > ROWS_MAX_TABLES || nops < 1 || ntemps < 1 || ntemps > 64) return B200_ERR_ARG;
  RowsArgs a;
  for (int i = 0; i < ntables; ++i) a.in[i] = tables[i];
  a.consts = d_consts;
  a.ops = d_ops;
  a.K = ntables;
  a.C = nconsts;
  a.nops = nops;
  a.out = d_out;
  a.N = (size_t)1 << num_vars;
  const int nslots = ntables + ntemps;
  int nth = 128;
  while (nth > 32 && (size_t)nslots * nth * sizeof(Fr) > 100 * 1024) nth -= 32;
  const size_t smem_bytes = (size_t)nslots * nth * sizeof(Fr);
  if (smem_bytes > 220 * 1024) return B200_ERR_ARG;
  size_t blocks = (a.N + nth - 1) / nth;
  if (blocks > 2 * NUM_SMS) blocks = 2 * NUM_SMS;
  expr_rows_kernel<<<(unsigned)blocks, nth, smem_bytes, c->stream>>>(a);
  count_launch(c);
  CUDA_TRY(cudaGetLastError());
  return B200_OK;
}

// every kernel of this file, loaded up front (b200_ctx_create -> preload_all_kernels, capi.cu)
void preload_lookup() {
  cudaFuncSetAttribute(expr_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);  // once per device
  B200_PRELOAD(expr_rows_kernel);
  B200_PRELOAD(lookup_insert_kernel);
  B200_PRELOAD(lookup_count_kernel);
  B200_PRELOAD(lookup_counts_to_fr_kernel);
  B200_PRELOAD(lookup_h_kernel);
}

}  // namespace b200

using namespace b200;

extern "C" {

int b200_expression_rows(b200_ctx* h, int num_vars, int ntables, const void* const* dev_tables, int nconsts,
                         const void* host_consts_fr, int nops, const int32_t* host_ops, void* dev_out) {
  Ctx* c = &h->c;
  if (num_vars < 1 || num_vars > 30 || ntables < 1 || ntables > ROWS_MAX_TABLES || nconsts < 0 || nops < 1)
    return B200_ERR_ARG;
  const int lim = ntables + nconsts;
  int max_dst = lim;
  for (int i = 0; i < nops; ++i) {
    const int32_t* o = host_ops + 4 * i;
    if (o[0] < 0 || o[0] > 3 || o[1] < lim || o[2] < 0 || o[3] < 0) return B200_ERR_ARG;
    if (o[1] > max_dst) max_dst = o[1];
  }
  for (int i = 0; i < nops; ++i)
    if (host_ops[4 * i + 2] > max_dst || host_ops[4 * i + 3] > max_dst) return B200_ERR_ARG;
  Fr* dconsts = nullptr;
  int4* dops = nullptr;
  CUDA_TRY(cudaMallocAsync(&dconsts, ((size_t)nconsts + 1) * sizeof(Fr), c->stream));
  CUDA_TRY(cudaMallocAsync(&dops, (size_t)nops * sizeof(int4), c->stream));
  if (nconsts)
    CUDA_TRY(cudaMemcpyAsync(dconsts, host_consts_fr, (size_t)nconsts * sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  CUDA_TRY(cudaMemcpyAsync(dops, host_ops, (size_t)nops * sizeof(int4), cudaMemcpyHostToDevice, c->stream));
  const int rc = expression_rows_prog(c, num_vars, (const Fr* const*)dev_tables, ntables, dconsts, nconsts, dops, nops,
                                      max_dst - lim + 1, (Fr*)dev_out);
  CUDA_TRY(cudaFreeAsync(dconsts, c->stream));
  CUDA_TRY(cudaFreeAsync(dops, c->stream));
  return rc;
}

int b200_lookup_m(b200_ctx* h, int num_vars, const void* dev_input, const void* dev_table, void* dev_m_out) {
  return lookup_m(&h->c, num_vars, (const Fr*)dev_input, (const Fr*)dev_table, (Fr*)dev_m_out);
}

int b200_lookup_h(b200_ctx* h, int num_vars, const void* dev_input, const void* dev_table, const void* dev_m,
                  const void* host_gamma, void* dev_h_out) {
  Ctx* c = &h->c;
  Fr* g = nullptr;
  CUDA_TRY(cudaMallocAsync(&g, sizeof(Fr), c->stream));
  CUDA_TRY(cudaMemcpyAsync(g, host_gamma, sizeof(Fr), cudaMemcpyHostToDevice, c->stream));
  const int rc = lookup_h(c, num_vars, (const Fr*)dev_input, (const Fr*)dev_table, (const Fr*)dev_m, g, (Fr*)dev_h_out);
  CUDA_TRY(cudaFreeAsync(g, c->stream));
  return rc;
}

}  // extern "C"

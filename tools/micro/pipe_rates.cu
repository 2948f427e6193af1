// Micro-benchmark: issue rates of the instruction classes a big-integer multiplier can be built from, per SM
// sub-partition (SMSP), alone and mixed — decides whether a 52-bit-limb FP64-FMA Montgomery product (DFMA on the
// FP64 pipe + integer adds on the ALU pipe) can beat the IMAD.WIDE one.
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o pipe_rates pipe_rates.cu
// Each warp issues ITERS x 64 independent instructions of one class over 8 accumulator chains; 4 warps per SMSP
// (16 per CTA), one CTA per SM. Reported: warp instructions per cycle per SMSP (clock64 over the whole CTA).
#include <cstdio>
#include <cuda_runtime.h>

#include "../../halo2-lasso_b200/csrc/ff32.cuh"

enum Op { IMAD_WIDE, IMAD_LO, IMAD_HI, DFMA, DADD, IADD32, IADD64, MIX_DFMA_IMADW, MIX_DFMA_IADD64, MIX_IMADW_IADD64, NOPS };
static const char* NAMES[NOPS] = {"mul.wide.u32 (IMAD.WIDE)", "mad.lo.u32 (IMAD)", "mad.hi.u32 (IMAD.HI)", "fma.rz.f64 (DFMA)",
                                  "add.rz.f64 (DADD)", "add.u32 (IADD3)", "add.u64", "DFMA + IMAD.WIDE 1:1", "DFMA + add.u64 1:1",
                                  "IMAD.WIDE + add.u64 1:1"};

template <int OP>
__global__ void __launch_bounds__(512) rate_kernel(unsigned long long* sink, long long* cycles, int iters, unsigned a0, unsigned b0) {
  unsigned long long w[8];
  unsigned u[8];
  double d[8];
  unsigned long long q[8];
  const unsigned a = a0 + threadIdx.x, b = b0 | 1u;
  const double da = 1.0 + 1e-9 * threadIdx.x, db = 1.0 - 1e-9;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    w[i] = i + threadIdx.x;
    u[i] = i * 3 + threadIdx.x;
    d[i] = 1.0 + i;
    q[i] = 7 * i + threadIdx.x;
  }
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int rep = 0; rep < 8; ++rep) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (OP == IMAD_WIDE || OP == MIX_DFMA_IMADW || OP == MIX_IMADW_IADD64)
          asm volatile("{ .reg .u32 lo, hi; mov.b64 {lo, hi}, %1; xor.b32 lo, lo, hi; mul.wide.u32 %0, lo, %2; }"
                       : "=l"(w[i]) : "l"(w[(i + 1) & 7]), "r"(b));
        if (OP == IMAD_LO) asm volatile("mad.lo.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(b));
        if (OP == IMAD_HI) asm volatile("mad.hi.u32 %0, %1, %2, %0;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]), "r"(b));
        if (OP == DFMA || OP == MIX_DFMA_IMADW || OP == MIX_DFMA_IADD64)
          asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[i]) : "d"(db), "d"(da));
        if (OP == DADD) asm volatile("add.rz.f64 %0, %0, %1;" : "+d"(d[i]) : "d"(da));
        if (OP == IADD32) asm volatile("add.u32 %0, %0, %1;" : "+r"(u[i]) : "r"(u[(i + 1) & 7]));
        if (OP == IADD64 || OP == MIX_DFMA_IADD64 || OP == MIX_IMADW_IADD64)
          asm volatile("add.u64 %0, %0, %1;" : "+l"(q[i]) : "l"(q[(i + 1) & 7]));
      }
    }
  }
  __syncthreads();
  const long long t1 = clock64();
  unsigned long long s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += w[i] + u[i] + (unsigned long long)__double_as_longlong(d[i]) + q[i];
  sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// the product's own Montgomery multiplier (ff32.cuh): W independent chains per thread
template <int W>
__global__ void __launch_bounds__(512) fr_mul_kernel(b200::Fr* io, long long* cycles, int iters) {
  using namespace b200;
  Fr x[W], y = io[0];
  for (int w = 0; w < W; ++w) x[w] = io[(blockIdx.x * blockDim.x + threadIdx.x) * W + w];
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = x[w] * y;
  }
  __syncthreads();
  const long long t1 = clock64();
  for (int w = 0; w < W; ++w) io[(blockIdx.x * blockDim.x + threadIdx.x) * W + w] = x[w];
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}
template <int W>
static void run_fr(long long* cyc, int sms, int warps_per_smsp) {
  b200::Fr* io;
  const int threads = 128 * warps_per_smsp, iters = 500;
  cudaMalloc(&io, (size_t)sms * threads * W * sizeof(b200::Fr));
  cudaMemset(io, 1, (size_t)sms * threads * W * sizeof(b200::Fr));
  fr_mul_kernel<W><<<sms, threads>>>(io, cyc, 5);
  cudaDeviceSynchronize();
  fr_mul_kernel<W><<<sms, threads>>>(io, cyc, iters);
  cudaDeviceSynchronize();
  const double prods = (double)iters * W * warps_per_smsp;  // warp-products per SMSP
  printf("Fr Montgomery product, %d chain(s)/thread, warps/SMSP %d: %7.1f cycles per warp-product per SMSP\n", W,
         warps_per_smsp, cyc[0] / prods);
  cudaFree(io);
}

template <int OP>
static void run(unsigned long long* sink, long long* cyc, int sms, int warps_per_smsp) {
  const int iters = 2000, threads = 32 * 4 * warps_per_smsp;
  rate_kernel<OP><<<sms, threads>>>(sink, cyc, 10, 3, 5);
  cudaDeviceSynchronize();
  rate_kernel<OP><<<sms, threads>>>(sink, cyc, iters, 3, 5);
  cudaDeviceSynchronize();
  const int per_iter = (OP >= MIX_DFMA_IMADW) ? 128 : 64;
  const double warp_instr = (double)iters * per_iter * warps_per_smsp;  // per SMSP
  printf("%-28s warps/SMSP %d: %8.3f warp-instr/cycle/SMSP (%.2f cycles per instr)%s\n", NAMES[OP], warps_per_smsp,
         warp_instr / cyc[0], cyc[0] / warp_instr, OP >= MIX_DFMA_IMADW ? "  [both classes counted]" : "");
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned long long* sink;
  long long* cyc;
  cudaMalloc(&sink, 8 * 1024 * 1024);
  cudaMallocManaged(&cyc, 1024 * sizeof(long long));
  for (int w : {1, 4}) {
    run<IMAD_WIDE>(sink, cyc, sms, w);
    run<IMAD_LO>(sink, cyc, sms, w);
    run<IMAD_HI>(sink, cyc, sms, w);
    run<DFMA>(sink, cyc, sms, w);
    run<DADD>(sink, cyc, sms, w);
    run<IADD32>(sink, cyc, sms, w);
    run<IADD64>(sink, cyc, sms, w);
    run<MIX_DFMA_IMADW>(sink, cyc, sms, w);
    run<MIX_DFMA_IADD64>(sink, cyc, sms, w);
    run<MIX_IMADW_IADD64>(sink, cyc, sms, w);
  }
  for (int w : {1, 2, 4}) {
    run_fr<1>(cyc, sms, w);
    run_fr<2>(cyc, sms, w);
  }
  printf("status %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}

// Micro-benchmark: single-warp latency of the Montgomery product, alone and with 2 / 4 independent chains.
#include <cstdio>
#include "../../halo2-lasso_b200/csrc/ff32.cuh"
using namespace b200;
template <int W>
__global__ void k(Fr* io, long long* cyc, int iters) {
  Fr x[W];
  for (int w = 0; w < W; ++w) x[w] = io[threadIdx.x * W + w];
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int w = 0; w < W; ++w) x[w] = x[w] * x[w];
  }
  long long t1 = clock64();
  for (int w = 0; w < W; ++w) io[threadIdx.x * W + w] = x[w];
  if (threadIdx.x == 0) cyc[W] = t1 - t0;
}
int main() {
  Fr* io; long long* cyc;
  cudaMalloc(&io, 32 * 8 * sizeof(Fr)); cudaMemset(io, 1, 32 * 8 * sizeof(Fr));
  cudaMallocManaged(&cyc, 16 * sizeof(long long));
  const int iters = 1000;
  k<1><<<1, 32>>>(io, cyc, iters); k<2><<<1, 32>>>(io, cyc, iters); k<4><<<1, 32>>>(io, cyc, iters);
  cudaDeviceSynchronize();
  k<1><<<1, 32>>>(io, cyc, iters); k<2><<<1, 32>>>(io, cyc, iters); k<4><<<1, 32>>>(io, cyc, iters);
  cudaDeviceSynchronize();
  for (int w : {1, 2, 4}) printf("W=%d: %.1f cycles per iteration (%.1f per product)\n", w, (double)cyc[w] / iters, (double)cyc[w] / iters / w);
  // multi-warp throughput: 8 warps per SMSP-equivalent block
  return 0;
}

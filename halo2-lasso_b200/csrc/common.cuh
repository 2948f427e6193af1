// Shared device helpers: block reductions of field elements, the "last CTA finalises" ticket, and
// the library context (stream, device transcript, scalar arena).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#include "ff32.cuh"
#include "transcript.cuh"

namespace b200 {

#define B200_OK 0
#define B200_ERR_CUDA 1
#define B200_ERR_ARG 2
#define B200_ERR_TRANSCRIPT 3  // identity commitment / proof overflow (Error::Transcript)
#define B200_ERR_NOMEM 4
#define B200_ERR_LOOKUP 5      // a lookup operand / input is not in its table
#define B200_ERR_PEER 6        // a multi-GPU wait timed out (a peer left the collective)

#define CUDA_TRY(x)                                                                      \
  do {                                                                                   \
    cudaError_t e_ = (x);                                                                \
    if (e_ != cudaSuccess) {                                                             \
      fprintf(stderr, "[b200lasso] CUDA error %s at %s:%d\n", cudaGetErrorString(e_),    \
              __FILE__, __LINE__);                                                       \
      return B200_ERR_CUDA;                                                              \
    }                                                                                    \
  } while (0)

static const int NUM_SMS = 148;
static const int SC_THREADS = 256;

#if defined(__CUDACC__)
__device__ __forceinline__ Fr fr_shfl_down(const Fr& a, int off) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_down_sync(0xffffffffu, a.v[i], off);
  return r;
}

// coherent (L2) load for data produced by other CTAs of the same launch
__device__ __forceinline__ Fr fr_ld_cg(const Fr* p) {
  Fr r;
  uint4 lo = __ldcg(reinterpret_cast<const uint4*>(p));
  uint4 hi = __ldcg(reinterpret_cast<const uint4*>(p) + 1);
  r.v[0] = lo.x; r.v[1] = lo.y; r.v[2] = lo.z; r.v[3] = lo.w;
  r.v[4] = hi.x; r.v[5] = hi.y; r.v[6] = hi.z; r.v[7] = hi.w;
  return r;
}

// Sum D field elements per thread across the CTA; the result is valid in thread 0 only.
// smem must hold (blockDim.x / 32) * D elements.
template <int D>
__device__ __forceinline__ void block_reduce_fr(Fr* acc, Fr* smem) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int x = 0; x < D; ++x) acc[x] = acc[x] + fr_shfl_down(acc[x], off);
  }
  if (lane == 0) {
#pragma unroll
    for (int x = 0; x < D; ++x) smem[warp * D + x] = acc[x];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int x = 0; x < D; ++x) acc[x] = lane < nwarps ? smem[lane * D + x] : fe_zero<FrP>();
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
      for (int x = 0; x < D; ++x) acc[x] = acc[x] + fr_shfl_down(acc[x], off);
    }
  }
  __syncthreads();
}

// warp-only reduction (result valid in lane 0)
template <int D>
__device__ __forceinline__ void warp_reduce_fr(Fr* acc) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
#pragma unroll
    for (int x = 0; x < D; ++x) acc[x] = acc[x] + fr_shfl_down(acc[x], off);
  }
}

// Returns true in every thread of the LAST CTA of the grid to arrive (all CTAs must call it after
// publishing their partial results). Resets the counter so the next launch can reuse it.
__device__ __forceinline__ bool last_cta_ticket(unsigned int* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int total = gridDim.x * gridDim.y * gridDim.z;
    const unsigned int ticket = atomicAdd(counter, 1u);
    s_last = (ticket == total - 1);
    if (s_last) *counter = 0;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}
#endif

// Programmatic dependent launch: a kernel lets the NEXT launch of the stream be scheduled at once (its CTAs become
// resident as SMs drain) and itself waits for the complete previous grid — memory included — before touching anything
// that grid wrote. Semantics are those of plain stream order; only the launch latency between the strictly
// sequential Fiat-Shamir steps overlaps the previous kernel's tail.
#if defined(__CUDACC__)
__device__ __forceinline__ void pdl_prologue() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
// Off when several ranks share ONE GPU (b200_dist_init_local): a dependent grid that is resident early and waits for
// its predecessor could otherwise occupy the SM slots another rank's kernel needs to make that predecessor finish.
inline bool g_use_pdl = true;
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s,
                                     Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// Barycentric weights for the nodes 0..d, d <= 6: w[d][i] = 1 / Π_{j != i} (i - j) (Montgomery).
// Lagrange interpolation with these constants is inversion-free and yields the same field element
// as `barycentric_interpolate` (pb/util/arithmetic.rs:125-136) for every r outside {0..d}.
struct BaryTable {
  Fr w[7][7];
};

}  // namespace b200

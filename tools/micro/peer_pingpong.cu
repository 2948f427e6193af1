// Single-process, multi-GPU micro-benchmark of the mailbox all-gather (peer.cuh): every GPU runs ONE kernel that
// performs `iters` collectives back to back. Isolates the NVLink / memory-model latency from launch effects.
#include <cstdio>
#include <vector>
#include "../../halo2-lasso_b200/csrc/peer.cuh"
using namespace b200;
__global__ void pingpong(PeerCtx pc, int iters, long long* cycles, Fr* sink, int gap_ns) {
  const int lane = threadIdx.x;
  Fr v = fe_zero<FrP>();
  v.v[0] = pc.rank + 1;
  long long t0 = clock64();
  Fr acc = fe_zero<FrP>();
  long long spent = 0, tail = 0;
  for (int k = 1; k <= iters; ++k) {
    if (gap_ns) {  // idle time between collectives, jittered per rank like independent kernels would be
      const long long g0 = clock64();
      while (clock64() - g0 < (long long)(gap_ns + 37 * pc.rank) * 2) {
      }
    }
    const long long c0 = clock64();
    peer_publish(pc, (unsigned)k, v, 4);
    if (lane < 4)
      for (int r = 0; r < pc.world; ++r) acc = acc + peer_read(pc, (unsigned)k, r, lane);
    spent += clock64() - c0;
    // a tail like the round kernel's: dependent field products + shared-memory traffic + a local global store
    const long long q0 = clock64();
    __shared__ Fr sh[32];
    Fr t = acc;
    t.v[0] |= 1;
    for (int j = 0; j < 24; ++j) {
      sh[lane] = t;
      __syncwarp();
      t = t * sh[(lane + 1) & 31];
      __syncwarp();
    }
    if (lane == 0) sink[4] = t;
    acc = acc + t;
    tail += clock64() - q0;
  }
  long long t1 = clock64();
  if (lane == 0) {
    cycles[0] = gap_ns ? spent : t1 - t0;
    cycles[1] = tail;
  }
  if (lane < 4) sink[lane] = acc;
}
int main(int argc, char** argv) {
  int ndev = 0;
  cudaGetDeviceCount(&ndev);
  for (int world = 2; world <= ndev; world *= 2) {
    std::vector<Mailbox*> box(world);
    std::vector<long long*> cyc(world);
    std::vector<Fr*> sink(world);
    std::vector<cudaStream_t> st(world);
    for (int d = 0; d < world; ++d) {
      cudaSetDevice(d);
      for (int e = 0; e < world; ++e)
        if (e != d) cudaDeviceEnablePeerAccess(e, 0);
      cudaMalloc(&box[d], sizeof(Mailbox));
      cudaMemset(box[d], 0, sizeof(Mailbox));
      cudaMallocManaged(&cyc[d], 2 * sizeof(long long));
      cudaMalloc(&sink[d], 8 * sizeof(Fr));
      cudaStreamCreate(&st[d]);
    }
    for (int d = 0; d < world; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); }
    for (int gap : {0, 30000}) {
    const int iters = gap ? 300 : 2000;
    for (int rep = 0; rep < 2; ++rep) {
      for (int d = 0; d < world; ++d) {
        cudaSetDevice(d);
        cudaMemset(box[d], 0, sizeof(Mailbox));
        cudaDeviceSynchronize();
      }
      for (int d = 0; d < world; ++d) {
        cudaSetDevice(d);
        PeerCtx pc;
        pc.rank = d;
        pc.world = world;
        for (int e = 0; e < world; ++e) pc.box[e] = box[e];
        pingpong<<<1, 32, 0, st[d]>>>(pc, iters, cyc[d], sink[d], gap);
      }
      for (int d = 0; d < world; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); }
    }
    printf("world %d gap %6d ns: %.2f us per collective (rank 0; includes waiting for the slowest rank)  err=%s\n", world, gap,
           (double)*cyc[0] / iters / 1965.0, cudaGetErrorString(cudaGetLastError()));
    for (int d = 0; d < world; ++d) printf("    rank %d tail %.2f us\n", d, (double)cyc[d][1] / iters / 1965.0);
    }
    for (int d = 0; d < world; ++d) {
      cudaSetDevice(d);
      for (int e = 0; e < world; ++e)
        if (e != d) cudaDeviceDisablePeerAccess(e);
      cudaFree(box[d]); cudaFree(cyc[d]); cudaFree(sink[d]); cudaStreamDestroy(st[d]);
    }
  }
  return 0;
}

// Peer-memory collectives for the hypercube-sharded sum-check and the point-sharded MSM.
// Every rank owns a small MAILBOX in its own HBM; all ranks map all mailboxes through CUDA IPC (NVLink 5 /
// NVSwitch P2P). A collective is executed INSIDE the compute kernel by one warp of its last CTA:
//   write my values into slot[my_rank] of every peer's mailbox  ->  __threadfence_system()  ->  publish the
//   sequence number  ->  spin until every slot of MY mailbox carries that sequence number  ->  read.
// Payloads are a few field elements (<= 192 B per round message), so this is latency- not bandwidth-bound;
// fusing it into the round kernel removes the NCCL launch and the extra kernel a host-driven all-gather
// would need. Slots are double-buffered on the parity of the sequence number: a rank can be at most one
// collective ahead of any peer (it needs that peer's data to finish the current one).
#pragma once
#include "ff32.cuh"

namespace b200 {

static const int PEER_MAX_WORLD = 8;
static const int PEER_MAX_VALS = 72;  // field elements per message

struct MailSlot {
  unsigned int seq[2];
  unsigned int pad[6];
  Fr data[2][PEER_MAX_VALS];
};
struct Mailbox {
  MailSlot slot[PEER_MAX_WORLD];  // indexed by SOURCE rank
};
struct PeerCtx {
  int rank, world;
  Mailbox* box[PEER_MAX_WORLD];  // box[r] = rank r's mailbox as mapped in this process (box[rank] is local)
};

#if defined(__CUDACC__)
__device__ __forceinline__ void st_sys_fr(Fr* p, const Fr& v) {
  volatile uint32_t* q = reinterpret_cast<volatile uint32_t*>(p);
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = v.v[i];
}
__device__ __forceinline__ Fr ld_sys_fr(const Fr* p) {
  const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(p);
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = q[i];
  return r;
}

// All-gather `cnt` (<= 32) field elements per rank. Called by ALL 32 lanes of one warp; lane i < cnt
// contributes `mine`. Afterwards lane i < cnt calls peer_read(pc, seq, r, i) for each source rank r.
__device__ __forceinline__ void peer_publish(const PeerCtx& pc, unsigned int seq, const Fr& mine, int cnt) {
  const int lane = threadIdx.x & 31;
  const int par = seq & 1;
  if (lane < cnt) {
    for (int r = 0; r < pc.world; ++r) st_sys_fr(&pc.box[r]->slot[pc.rank].data[par][lane], mine);
  }
  // release: the payload stores of all lanes (ordered before the flag stores by the warp barrier) become visible
  // to a peer before the sequence number does; one acquire fence after the poll on the reading side.
  // Measured (tools/micro/): 7 us per collective at 2 GPUs and 11 us at 4 inside one long-running kernel; in the
  // one-process-per-GPU prover 11 / 15 / 77 us per round at 2 / 4 / 8 GPUs with peer_spin_while below.
  __syncwarp();
  if (lane < pc.world) {
    unsigned int* f = &pc.box[lane]->slot[pc.rank].seq[par];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
  }
  // wait for every source rank (lane r polls source r in MY mailbox), then acquire once
  if (lane < pc.world) {
    const unsigned int* f = &pc.box[pc.rank]->slot[lane].seq[par];
    unsigned int v;
    do {
      asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    } while (v != seq);
    asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncwarp();
}
// Keep-busy helper. A GPU whose only activity is one warp polling NVLink-written memory drops into a low-activity
// state in which the code that FOLLOWS the wait runs ~10x slower for ~120 us (measured at 4 and 8 GPUs, one
// process per GPU: 125 us per collective instead of 11-15). Warps 1..3 of the CTA (one per other SM
// sub-partition) therefore issue arithmetic while warp 0 runs the exchange: they call peer_spin_while(flag) after a
// barrier that publishes *flag = 1, warp 0 clears the flag when it is done.
__device__ __forceinline__ void peer_spin_while(volatile int* flag, unsigned int* sink) {
  float x = (float)threadIdx.x;
  while (*flag) {
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fmaf(x, 1.0001f, 0.5f);
  }
  if (x == 12345.678f) *sink = 0;  // keeps the loop alive
}
__device__ __forceinline__ Fr peer_read(const PeerCtx& pc, unsigned int seq, int src, int idx) {
  return ld_sys_fr(&pc.box[pc.rank]->slot[src].data[seq & 1][idx]);
}
#endif

}  // namespace b200

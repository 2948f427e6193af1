// Peer-memory collectives for the sharded provers. One process (or, in tests, one context) per
// GPU; every rank owns a MAILBOX and a bulk ARENA in its own HBM and maps those of all peers (CUDA IPC over NVLink 5 /
// NVSwitch; plain device pointers when several contexts share one GPU). Collectives run INSIDE the compute kernels:
//
//  * small messages (round partials, evaluations, commitments; <= 32 field elements per warp-level all-gather): payload
//    stores into slot [my rank] of every peer's mailbox, ONE release store of the collective's sequence number per peer,
//    every lane r polls source r's number in the local mailbox, one acquire fence (protocol 1, the default: 9 us per
//    sharded sum-check round at 2 GPUs). Protocol 0 is the NCCL-LL idea — every 32-bit word travels in one 8-byte store
//    together with the sequence number, no fence at all — and measured 105 us per round on the same box: without a
//    release the relaxed system-scope stores are not pushed out promptly (tools/micro/shard_rounds.py, profiles/).
//  * bulk all-gathers (bound sum-check tables, tree layers): the producing kernel stores straight into every peer's
//    arena (posted NVLink writes), its last CTA publishes a per-source sequence number with a release store and waits
//    for the other sources with acquire loads.
//
// Both are double-buffered on the parity of their sequence number: a rank can be at most one collective ahead of any
// peer (it needs that peer's contribution to finish the current one). Every wait is BOUNDED: after `timeout_ns` the
// waiter raises *err (surfaced as B200_ERR_PEER) and carries on with garbage instead of hanging the GPU.
#pragma once
#include "ff32.cuh"

namespace b200 {

static const int PEER_MAX_WORLD = 8;
static const int PEER_MAX_VALS = 72;  // field elements per small message

struct Mailbox {
  unsigned long long ll[2][PEER_MAX_WORLD][PEER_MAX_VALS * 8];  // [parity][SOURCE rank][word]: data | seq << 32
  unsigned int bulk_seq[PEER_MAX_WORLD];                         // [SOURCE rank]: last bulk all-gather it has pushed
  unsigned int flag_seq[2][PEER_MAX_WORLD];                      // protocol 1: [parity][SOURCE rank] sequence number
  unsigned int pad[8];
  Fr data[2][PEER_MAX_WORLD][32];                                // protocol 1: payload ([parity][SOURCE rank][value])
};
struct PeerCtx {
  int rank, world;
  Mailbox* box[PEER_MAX_WORLD];        // box[r] = rank r's mailbox as mapped here (box[rank] is local)
  unsigned char* arena[PEER_MAX_WORLD];  // arena[r] = rank r's bulk arena (2 halves of arena_half bytes)
  unsigned long long arena_half;
  unsigned long long timeout_ns;
  unsigned int* err;  // local device word, non-zero once a wait timed out: kind << 28 | source rank << 24 | sequence number
                      // of the FIRST wait that gave up (kind 1: LL word, 2: small-message flag, 3: bulk all-gather)
  int proto;          // small messages: 0 = LL words (data | seq in one 8-byte store), 1 = payload + release flag + fence,
                      // 2 = acquire polls instead of the fence, 3 = 2 with the payload as one posted 256-bit store per
                      // (value, rank) pair spread over the lanes
};

#if defined(__CUDACC__)
__device__ __forceinline__ unsigned long long peer_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ void st_sys_u64(unsigned long long* p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// All-gather of `cnt` (<= PEER_MAX_VALS) field elements per rank. Lane / thread i < cnt calls peer_put with its value
// (no waiting: posted stores to every rank, itself included), then anybody calls peer_get(src, idx), which spins until
// the 8 words of that element carry `seq`.
__device__ __forceinline__ void peer_put(const PeerCtx& pc, unsigned int seq, int idx, const Fr& v) {
  const int par = seq & 1;
  for (int r = 0; r < pc.world; ++r) {
    unsigned long long* dst = &pc.box[r]->ll[par][pc.rank][idx * 8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st_sys_u64(dst + i, (unsigned long long)v.v[i] | ((unsigned long long)seq << 32));
  }
}
__device__ __forceinline__ Fr peer_get(const PeerCtx& pc, unsigned int seq, int src, int idx) {
  const unsigned long long* p = &pc.box[pc.rank]->ll[seq & 1][src][idx * 8];
  Fr r;
  unsigned long long t0 = 0;
  unsigned int spins = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    unsigned long long w = ld_sys_u64(p + i);
    while ((unsigned int)(w >> 32) != seq) {
      if ((++spins & 1023u) == 0) {
        const unsigned long long now = peer_now_ns();
        if (!t0) t0 = now;
        else if (now - t0 > pc.timeout_ns) {
          atomicCAS(pc.err, 0u, 0x10000000u | (seq & 0x0fffffffu));
          break;
        }
      }
      w = ld_sys_u64(p + i);
    }
    r.v[i] = (uint32_t)w;
  }
  return r;
}
// Warp-level all-gather of cnt <= 32 values: called by ALL 32 lanes of one warp, lane i < cnt contributes `mine`;
// afterwards any lane calls peer_read(src, idx). Protocol 0: LL words, the wait happens in peer_read. Protocol 1 (the
// round-1 scheme): payload stores, one release store of the sequence number per peer, every lane r < world polls
// source r's number in the local mailbox, one acquire fence.
__device__ __forceinline__ void peer_publish(const PeerCtx& pc, unsigned int seq, const Fr& mine, int cnt) {
  const int lane = threadIdx.x & 31;
  if (pc.proto == 0) {
    if (lane < cnt) peer_put(pc, seq, lane, mine);
    return;
  }
  const int par = seq & 1;
  if (pc.proto == 3) {
    // Protocol 3: the cnt x world (value, destination) pairs are spread over the lanes and each goes out as ONE posted
    // 256-bit store. Protocol 1 / 2 let lane i write its value to every rank with 8 volatile 32-bit stores per rank:
    // strong system-scope stores of one thread do not overlap, so a round paid 8 x world NVLink round trips in a row
    // (measured per sharded round: 5 us at 2 GPUs, 15 us at 4, 117 us at 8 — independent of fences and heartbeats).
    const int ntask = cnt * pc.world;
    for (int base = 0; base < ntask; base += 32) {
      const int task = base + lane;
      const int j = task % cnt, r = task / cnt;
      Fr v;
#pragma unroll
      for (int i = 0; i < 8; ++i) v.v[i] = __shfl_sync(0xffffffffu, mine.v[i], j);
      if (task < ntask) fe_st(&pc.box[r]->data[par][pc.rank][j], v);
    }
  } else if (lane < cnt) {
    for (int r = 0; r < pc.world; ++r) {
      volatile uint32_t* q = reinterpret_cast<volatile uint32_t*>(&pc.box[r]->data[par][pc.rank][lane]);
#pragma unroll
      for (int i = 0; i < 8; ++i) q[i] = mine.v[i];
    }
  }
  __syncwarp();
  if (lane < pc.world) {
    unsigned int* f = &pc.box[lane]->flag_seq[par][pc.rank];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(seq) : "memory");
    const unsigned int* g = &pc.box[pc.rank]->flag_seq[par][lane];
    unsigned int v, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      if (pc.proto >= 2) asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
      else asm volatile("ld.relaxed.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
      if (v == seq) break;
      if ((++spins & 1023u) == 0) {
        const unsigned long long now = peer_now_ns();
        if (!t0) t0 = now;
        else if (now - t0 > pc.timeout_ns) {
          atomicCAS(pc.err, 0u, 0x20000000u | ((unsigned)lane << 24) | (seq & 0x00ffffffu));
          break;
        }
      }
    }
    // protocol 1: full system-scope fence after the poll. Protocol 2: the poll itself is an acquire load (below) and the
    // payload is read with L1-bypassing volatile loads, so no fence follows.
    if (pc.proto == 1) asm volatile("fence.acq_rel.sys;" ::: "memory");
  }
  __syncwarp();
}
__device__ __forceinline__ Fr peer_read(const PeerCtx& pc, unsigned int seq, int src, int idx) {
  if (pc.proto == 0) return peer_get(pc, seq, src, idx);
  const volatile uint32_t* q = reinterpret_cast<const volatile uint32_t*>(&pc.box[pc.rank]->data[seq & 1][src][idx]);
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = q[i];
  return r;
}

// Keep-busy helper. A GPU whose only activity is one warp polling NVLink-written memory drops into a low-activity
// state in which the code that FOLLOWS the wait runs ~10x slower for ~120 us (measured at 4 and 8 GPUs, one
// process per GPU). Warps 1..3 of the CTA (one per other SM sub-partition) therefore issue arithmetic
// while warp 0 runs the exchange: they call peer_spin_while(flag) after a barrier that publishes *flag = 1, warp 0
// clears the flag when it is done.
__device__ __forceinline__ void peer_spin_while(volatile int* flag, unsigned int* sink) {
  float x = (float)threadIdx.x;
  while (*flag) {
#pragma unroll
    for (int i = 0; i < 64; ++i) x = fmaf(x, 1.0001f, 0.5f);
  }
  if (x == 12345.678f) *sink = 0;  // keeps the loop alive
}

// ---- bulk all-gather ---------------------------------------------------------------------------------------------
// peer_bulk_dst: where element `idx` of the gathered buffer lives in rank r's arena for collective `bseq`.
__device__ __forceinline__ Fr* peer_bulk_dst(const PeerCtx& pc, unsigned int bseq, int r, size_t idx) {
  return reinterpret_cast<Fr*>(pc.arena[r] + (size_t)(bseq & 1) * pc.arena_half) + idx;
}
// Called by ONE warp of the last CTA of the pushing kernel, after every CTA's stores were fenced with
// __threadfence_system() and counted through the ticket: publish my sequence number everywhere, wait for all sources.
__device__ __forceinline__ void peer_bulk_commit_and_wait(const PeerCtx& pc, unsigned int bseq) {
  const int lane = threadIdx.x & 31;
  if (lane < pc.world) {
    unsigned int* f = &pc.box[lane]->bulk_seq[pc.rank];
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(bseq) : "memory");
    const unsigned int* g = &pc.box[pc.rank]->bulk_seq[lane];
    unsigned int v, spins = 0;
    unsigned long long t0 = 0;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
      if ((int)(v - bseq) >= 0) break;
      if ((++spins & 1023u) == 0) {
        const unsigned long long now = peer_now_ns();
        if (!t0) t0 = now;
        else if (now - t0 > pc.timeout_ns) {
          atomicCAS(pc.err, 0u, 0x30000000u | ((unsigned)lane << 24) | (bseq & 0x00ffffffu));
          break;
        }
      }
    }
  }
  __syncwarp();
}
#endif

}  // namespace b200

// Device-resident Fiat-Shamir transcript: FiatShamirTranscript<Keccak256, Cursor<Vec<u8>>> of
// pb/util/transcript.rs:99-238 (+ pb/util/hash.rs:19-21, pb/util/arithmetic.rs:150-152).
//
// B200 design point: the reference interleaves every prover step with a Keccak absorb/squeeze on
// the CPU. Here the sponge state AND the proof byte stream live in HBM; the last CTA of each round
// kernel absorbs the round message, squeezes the challenge and folds the claim, so a whole
// sum-check (and the whole Lasso proof) is enqueued without a single host round trip.
// All functions are single-thread code (one elected thread runs them).
#pragma once
#include "ff32.cuh"

namespace b200 {

struct Transcript {
  uint64_t s[25];      // Keccak-f[1600] state
  uint32_t pos;        // bytes absorbed into the current rate block (0..135)
  uint32_t proof_len;  // bytes appended to the proof stream
  uint32_t proof_cap;
  uint32_t error;      // sticky: 1 = proof overflow, 2 = identity commitment (transcript.rs:174-179)
  uint8_t* proof;      // device buffer
};

FF_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

FF_HD void keccak_f1600(uint64_t* s) {
  const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  uint64_t a[25];
#pragma unroll
  for (int i = 0; i < 25; ++i) a[i] = s[i];
#pragma unroll 1
  for (int round = 0; round < 24; ++round) {
    uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
    uint64_t c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
    uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
    uint64_t c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
    uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
    uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1);
    uint64_t d3 = c2 ^ rotl64(c4, 1), d4 = c3 ^ rotl64(c0, 1);
#pragma unroll
    for (int j = 0; j < 25; j += 5) {
      a[j] ^= d0; a[j + 1] ^= d1; a[j + 2] ^= d2; a[j + 3] ^= d3; a[j + 4] ^= d4;
    }
    // rho + pi (explicit lane schedule: b[y, 2x+3y] = rot(a[x, y]))
    uint64_t b[25];
    b[0] = a[0];
    b[10] = rotl64(a[1], 1);   b[20] = rotl64(a[2], 62);  b[5] = rotl64(a[3], 28);   b[15] = rotl64(a[4], 27);
    b[16] = rotl64(a[5], 36);  b[1] = rotl64(a[6], 44);   b[11] = rotl64(a[7], 6);   b[21] = rotl64(a[8], 55);
    b[6] = rotl64(a[9], 20);   b[7] = rotl64(a[10], 3);   b[17] = rotl64(a[11], 10); b[2] = rotl64(a[12], 43);
    b[12] = rotl64(a[13], 25); b[22] = rotl64(a[14], 39); b[23] = rotl64(a[15], 41); b[8] = rotl64(a[16], 45);
    b[18] = rotl64(a[17], 15); b[3] = rotl64(a[18], 21);  b[13] = rotl64(a[19], 8);  b[14] = rotl64(a[20], 18);
    b[24] = rotl64(a[21], 2);  b[9] = rotl64(a[22], 61);  b[19] = rotl64(a[23], 56); b[4] = rotl64(a[24], 14);
#pragma unroll
    for (int j = 0; j < 25; j += 5) {
      a[j] = b[j] ^ (~b[j + 1] & b[j + 2]);
      a[j + 1] = b[j + 1] ^ (~b[j + 2] & b[j + 3]);
      a[j + 2] = b[j + 2] ^ (~b[j + 3] & b[j + 4]);
      a[j + 3] = b[j + 3] ^ (~b[j + 4] & b[j]);
      a[j + 4] = b[j + 4] ^ (~b[j] & b[j + 1]);
    }
    a[0] ^= RC[round];
  }
#pragma unroll
  for (int i = 0; i < 25; ++i) s[i] = a[i];
}

FF_HD void tr_init(Transcript* t, uint8_t* proof, uint32_t cap) {
  for (int i = 0; i < 25; ++i) t->s[i] = 0;
  t->pos = 0;
  t->proof_len = 0;
  t->proof_cap = cap;
  t->error = 0;
  t->proof = proof;
}

// absorb 32 little-endian bytes given as 8 u32 words (pos is always a multiple of 4 here because
// every absorbed item is 32 bytes; the rate is 136 = 34 words)
FF_HD void tr_absorb_words(Transcript* t, const uint32_t* w, int nwords) {
  for (int i = 0; i < nwords; ++i) {
    uint32_t p = t->pos;
    t->s[p >> 3] ^= (uint64_t)w[i] << (8 * (p & 7));
    p += 4;
    if (p == 136) {
      keccak_f1600(t->s);
      p = 0;
    }
    t->pos = p;
  }
}

// squeeze_challenge (transcript.rs:127-131): hash = finalize; reset; absorb(hash); int_LE(hash) mod r
FF_HD Fr tr_squeeze(Transcript* t) {
  uint32_t p = t->pos;
  t->s[p >> 3] ^= (uint64_t)0x01 << (8 * (p & 7));
  t->s[16] ^= 0x8000000000000000ULL;
  keccak_f1600(t->s);
  Fr h;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    h.v[2 * i] = (uint32_t)t->s[i];
    h.v[2 * i + 1] = (uint32_t)(t->s[i] >> 32);
  }
  for (int i = 0; i < 25; ++i) t->s[i] = 0;
  t->pos = 0;
  tr_absorb_words(t, h.v, 8);
  return fe_from_canonical<FrP>(h);
}

// append the byte-reversed (big-endian) canonical repr to the proof stream
FF_HD void tr_stream_be(Transcript* t, const uint32_t* canon) {
  if (t->proof_len + 32 > t->proof_cap) {
    t->error |= 1;
    return;
  }
  uint8_t* o = t->proof + t->proof_len;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    uint32_t w = canon[7 - i];
    o[4 * i] = (uint8_t)(w >> 24);
    o[4 * i + 1] = (uint8_t)(w >> 16);
    o[4 * i + 2] = (uint8_t)(w >> 8);
    o[4 * i + 3] = (uint8_t)w;
  }
  t->proof_len += 32;
}

// common_field_element (transcript.rs:133-136)
FF_HD void tr_common_fe(Transcript* t, const Fr& fe) {
  Fr c = fe_to_canonical<FrP>(fe);
  tr_absorb_words(t, c.v, 8);
}
// write_field_element (transcript.rs:158-165)
FF_HD void tr_write_fe(Transcript* t, const Fr& fe) {
  Fr c = fe_to_canonical<FrP>(fe);
  tr_absorb_words(t, c.v, 8);
  tr_stream_be(t, c.v);
}
// write_commitment (transcript.rs:171-183, 216-227); (x, y) Montgomery Fq; identity -> error
FF_HD void tr_write_commitment(Transcript* t, const Fq& x, const Fq& y) {
  if (fe_is_zero<FqP>(x) && fe_is_zero<FqP>(y)) {
    t->error |= 2;
    return;
  }
  Fq cx = fe_to_canonical<FqP>(x), cy = fe_to_canonical<FqP>(y);
  tr_absorb_words(t, cx.v, 8);
  tr_absorb_words(t, cy.v, 8);
  tr_stream_be(t, cx.v);
  tr_stream_be(t, cy.v);
}

// ---------------------------------------------------------------------------------------------
// Warp-cooperative transcript (device only). A single GPU thread needs ~14k instructions per
// Keccak-f[1600]; with one 64-bit lane of the state per thread of a warp the permutation is ~20
// shuffles + ~20 ALU instructions per round. ALL 32 lanes of one warp must call these functions with
// identical arguments (the code is warp-uniform; only lane 0 writes memory).
// ---------------------------------------------------------------------------------------------
#if defined(__CUDACC__)
__device__ __forceinline__ uint64_t rotl64v(uint64_t x, unsigned n) {
  return n ? (x << n) | (x >> (64 - n)) : x;
}

// One shared, NON-inlined copy of the Montgomery product for cold latency-bound code (round
// finalisation, transcript): executed once per launch by a single warp, such code is instruction-
// fetch bound, so it must stay small enough to live in the instruction cache.
static __device__ __noinline__ Fr fr_mul_ni(Fr a, Fr b) { return fe_mul<FrP>(a, b); }
static __device__ __noinline__ Fq fq_mul_ni(Fq a, Fq b) { return fe_mul<FqP>(a, b); }
__device__ __forceinline__ Fr fr_bcast(const Fr& a, int src) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], src);
  return r;
}
__device__ __forceinline__ Fr fr_canon_ni(const Fr& a) {
  Fr one = fe_zero<FrP>();
  one.v[0] = 1;
  return fr_mul_ni(a, one);
}

// round constants and rho offsets (indexed by lane = x + 5y) in the constant bank: no local-memory (stack) traffic in
// the single-warp Fiat-Shamir tail, whose L1 contents a system-scope acquire fence of the multi-GPU path invalidates
__constant__ uint64_t KECCAK_RC_DEV[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
    0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
    0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
    0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
    0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
__constant__ unsigned KECCAK_RHO_DEV[25] = {0,  1,  62, 28, 27, 36, 44, 6,  55, 20, 3,  10, 43,
                                            25, 39, 41, 45, 15, 21, 8,  18, 2,  61, 56, 14};
static __device__ __noinline__ void keccak_f1600_warp(uint64_t* s) {
  const uint64_t* RC = KECCAK_RC_DEV;
  const unsigned* RHO = KECCAK_RHO_DEV;
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int l = lane < 25 ? lane : 0;  // lanes 25..31 idle along (never read by the others, never store)
  const int x = l % 5, y = l / 5;
  __syncwarp();
  uint64_t a = lane < 25 ? s[lane] : 0;
  const unsigned rho = RHO[l];
  // pi: dest (X, Y) <- src ((X + 3Y) mod 5, X)
  const int pi_src = ((x + 3 * y) % 5) + 5 * x;
  const int col1 = x + 5 * ((y + 1) % 5), col2 = x + 5 * ((y + 2) % 5), col3 = x + 5 * ((y + 3) % 5),
            col4 = x + 5 * ((y + 4) % 5);
  const int cl = (x + 4) % 5, cr = (x + 1) % 5;
  // chi reads b at the two next lanes of the row, and b is itself a permutation (pi) of the rotated lanes: fetch all
  // three straight from `rot` through the composed lane maps — one dependent shuffle stage instead of two per round
  const int x1 = (x + 1) % 5, x2 = (x + 2) % 5;
  const int pi_src1 = ((x1 + 3 * y) % 5) + 5 * x1, pi_src2 = ((x2 + 3 * y) % 5) + 5 * x2;
#pragma unroll 1
  for (int round = 0; round < 24; ++round) {
    uint64_t c = a ^ __shfl_sync(FULL, a, col1) ^ __shfl_sync(FULL, a, col2) ^ __shfl_sync(FULL, a, col3) ^
                 __shfl_sync(FULL, a, col4);
    const uint64_t c_l = __shfl_sync(FULL, c, cl), c_r = __shfl_sync(FULL, c, cr);
    a ^= c_l ^ rotl64v(c_r, 1);
    const uint64_t rot = rotl64v(a, rho);
    const uint64_t b = __shfl_sync(FULL, rot, pi_src);
    const uint64_t b1 = __shfl_sync(FULL, rot, pi_src1), b2 = __shfl_sync(FULL, rot, pi_src2);
    a = b ^ (~b1 & b2);
    if (l == 0) a ^= RC[round];
  }
  if (lane < 25) s[lane] = a;
  __syncwarp();
}

// `t` must be in shared memory (or global) visible to the whole warp
// Absorb `nwords` (<= 32) little-endian u32 words held identically by every lane: lane i XORs word i
// into the rate (viewed as 34 u32 words); if the block fills up, permute and let the remaining lanes
// continue at the start of the rate.
static __device__ __noinline__ void trw_absorb_words(Transcript* t, const uint32_t* w, int nwords) {
  const int lane = threadIdx.x & 31;
  uint32_t* s32 = reinterpret_cast<uint32_t*>(t->s);
  const uint32_t p = t->pos >> 2;  // word position inside the 34-word rate
  uint32_t mine = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (k == lane) mine = w[k];
  if (nwords > 8) {  // generic path (not used on the hot path): sequential
    for (int i = 0; i < nwords; ++i) {
      uint32_t q = (t->pos >> 2);
      if (lane == 0) s32[q] ^= w[i];
      __syncwarp();
      if (lane == 0) t->pos = (q + 1 == 34) ? 0 : 4 * (q + 1);
      __syncwarp();
      if (q + 1 == 34) {
        keccak_f1600_warp(t->s);
        __syncwarp();
      }
    }
    return;
  }
  const uint32_t j = p + lane;
  if (lane < nwords && j < 34) s32[j] ^= mine;
  __syncwarp();
  if (p + nwords >= 34) {
    keccak_f1600_warp(t->s);
    __syncwarp();
    if (lane < nwords && j >= 34) s32[j - 34] ^= mine;
    __syncwarp();
  }
  if (lane == 0) t->pos = 4 * ((p + nwords) % 34);
  __syncwarp();
}
__device__ __forceinline__ Fr trw_squeeze(Transcript* t) {
  const int lane = threadIdx.x & 31;
  const uint32_t p = t->pos;
  if (lane == 0) {
    t->s[p >> 3] ^= (uint64_t)0x01 << (8 * (p & 7));
    t->s[16] ^= 0x8000000000000000ULL;
  }
  __syncwarp();
  keccak_f1600_warp(t->s);
  __syncwarp();
  Fr h;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint64_t v = t->s[i];
    h.v[2 * i] = (uint32_t)v;
    h.v[2 * i + 1] = (uint32_t)(v >> 32);
  }
  __syncwarp();
  if (lane < 25) t->s[lane] = 0;
  if (lane == 0) t->pos = 0;
  __syncwarp();
  trw_absorb_words(t, h.v, 8);
  Fr r2;
#pragma unroll
  for (int i = 0; i < 8; ++i) r2.v[i] = FrP::r2(i);
  return fr_mul_ni(r2, h);  // (hash mod r) in Montgomery form; the raw 256-bit operand is the scanned one
}
// big-endian proof bytes of one element, written as 8 byte-swapped words by 8 lanes
__device__ __forceinline__ void trw_stream_be(Transcript* t, const uint32_t* canon);
// absorb (and stream) the element whose CANONICAL limbs lane `src` holds: lets the caller convert
// several elements to canonical form in parallel lanes
__device__ __forceinline__ void trw_write_canon_from_lane(Transcript* t, const Fr& canon_mine, int src, bool stream) {
  const Fr c = fr_bcast(canon_mine, src);
  trw_absorb_words(t, c.v, 8);
  if (stream) trw_stream_be(t, c.v);
}
// cooperative copy of the transcript struct between global and shared memory (56 words)
__device__ __forceinline__ void trw_copy(Transcript* dst, const Transcript* src) {
  const int lane = threadIdx.x & 31;
  const uint32_t* s32 = reinterpret_cast<const uint32_t*>(src);
  uint32_t* d32 = reinterpret_cast<uint32_t*>(dst);
  constexpr int W = (int)(sizeof(Transcript) / 4);
  for (int i = lane; i < W; i += 32) d32[i] = s32[i];
  __syncwarp();
}
__device__ __forceinline__ void trw_stream_be(Transcript* t, const uint32_t* canon) {
  const int lane = threadIdx.x & 31;
  const uint32_t len = t->proof_len;
  if (len + 32 > t->proof_cap) {
    if (lane == 0) t->error |= 1;
    __syncwarp();
    return;
  }
  uint32_t w = 0;
#pragma unroll
  for (int k = 0; k < 8; ++k)
    if (7 - k == lane) w = canon[k];
  if (lane < 8) reinterpret_cast<uint32_t*>(t->proof + len)[lane] = __byte_perm(w, 0, 0x0123);
  __syncwarp();
  if (lane == 0) t->proof_len = len + 32;
  __syncwarp();
}
__device__ __forceinline__ void trw_common_fe(Transcript* t, const Fr& fe) {
  const Fr c = fr_canon_ni(fe);
  trw_absorb_words(t, c.v, 8);
}
__device__ __forceinline__ void trw_write_fe(Transcript* t, const Fr& fe) {
  const Fr c = fr_canon_ni(fe);
  trw_absorb_words(t, c.v, 8);
  trw_stream_be(t, c.v);
}
__device__ __forceinline__ void trw_write_commitment(Transcript* t, const Fq& x, const Fq& y) {
  if (fe_is_zero<FqP>(x) && fe_is_zero<FqP>(y)) {
    if ((threadIdx.x & 31) == 0) t->error |= 2;
    __syncwarp();
    return;
  }
  Fq one = fe_zero<FqP>();
  one.v[0] = 1;
  const Fq cx = fq_mul_ni(x, one), cy = fq_mul_ni(y, one);
  trw_absorb_words(t, cx.v, 8);
  trw_absorb_words(t, cy.v, 8);
  trw_stream_be(t, cx.v);
  trw_stream_be(t, cy.v);
}
// broadcast a field element from lane 0 to the whole warp
__device__ __forceinline__ Fr fr_bcast0(const Fr& a) {
  Fr r;
#pragma unroll
  for (int i = 0; i < 8; ++i) r.v[i] = __shfl_sync(0xffffffffu, a.v[i], 0);
  return r;
}
#endif

}  // namespace b200

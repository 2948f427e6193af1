// CPU verifier: Keccak-256 (rate 136, pad 0x01) and the reading side of the Fiat-Shamir transcript
// (`FiatShamirTranscript<Keccak256, Cursor<Vec<u8>>>`, pb/util/transcript.rs:99-238; `Hash::update_field_element`,
// pb/util/hash.rs:19-21): field elements and G1 coordinates are absorbed as little-endian reprs and travel in the
// proof byte-reversed; non-canonical encodings and the identity point are rejected.
#pragma once
#include <vector>

#include "field.hpp"
#include "g1.hpp"

namespace b200v {

inline void keccak_f1600(uint64_t s[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
      0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
      0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
      0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
      0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
      0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
  static const int ROT[24] = {1,  3,  6,  10, 15, 21, 28, 36, 45, 55, 2,  14,
                              27, 41, 56, 8,  25, 43, 62, 18, 39, 61, 20, 44};
  static const int PIL[24] = {10, 7,  11, 17, 18, 3, 5,  16, 8,  21, 24, 4,
                              15, 23, 19, 13, 12, 2, 20, 14, 22, 9,  6,  1};
  for (int round = 0; round < 24; ++round) {
    uint64_t bc[5];
    for (int i = 0; i < 5; ++i) bc[i] = s[i] ^ s[i + 5] ^ s[i + 10] ^ s[i + 15] ^ s[i + 20];
    for (int i = 0; i < 5; ++i) {
      uint64_t t = bc[(i + 4) % 5] ^ ((bc[(i + 1) % 5] << 1) | (bc[(i + 1) % 5] >> 63));
      for (int j = 0; j < 25; j += 5) s[j + i] ^= t;
    }
    uint64_t t = s[1];
    for (int i = 0; i < 24; ++i) {
      int j = PIL[i];
      uint64_t b = s[j];
      s[j] = (t << ROT[i]) | (t >> (64 - ROT[i]));
      t = b;
    }
    for (int j = 0; j < 25; j += 5) {
      for (int i = 0; i < 5; ++i) bc[i] = s[j + i];
      for (int i = 0; i < 5; ++i) s[j + i] ^= (~bc[(i + 1) % 5]) & bc[(i + 2) % 5];
    }
    s[0] ^= RC[round];
  }
}

// Streaming sponge with `Digest::update` / `finalize_fixed_reset` semantics. `pad` = 0x01 gives
// Keccak-256, 0x06 gives NIST SHA3-256 (used only as a permutation known-answer check).
struct Keccak256 {
  uint64_t s[25];
  uint32_t pos;
  uint8_t pad;
  explicit Keccak256(uint8_t pad_byte = 0x01) : pos(0), pad(pad_byte) { memset(s, 0, sizeof s); }
  void update(const uint8_t* data, size_t n) {
    for (size_t i = 0; i < n; ++i) {
      s[pos >> 3] ^= (uint64_t)data[i] << (8 * (pos & 7));
      if (++pos == 136) {
        keccak_f1600(s);
        pos = 0;
      }
    }
  }
  void finalize_reset(uint8_t out[32]) {
    s[pos >> 3] ^= (uint64_t)pad << (8 * (pos & 7));
    s[16] ^= 0x8000000000000000ULL;  // byte 135
    keccak_f1600(s);
    memcpy(out, s, 32);
    memset(s, 0, sizeof s);
    pos = 0;
  }
};

// FiatShamirTranscript<Keccak256, Cursor<Vec<u8>>> (pb/util/transcript.rs:99-238).
struct Transcript {
  Keccak256 state;
  std::vector<uint8_t> stream;  // the proof
  size_t rpos = 0;              // read cursor (verifier side)

  Transcript() {}
  explicit Transcript(const std::vector<uint8_t>& proof) : stream(proof) {}

  // transcript.rs:127-131
  Fr squeeze_challenge() {
    uint8_t hash[32];
    state.finalize_reset(hash);
    state.update(hash, 32);
    return Fr::from_le_bytes_mod(hash);
  }
  std::vector<Fr> squeeze_challenges(size_t n) {
    std::vector<Fr> out(n);
    for (size_t i = 0; i < n; ++i) out[i] = squeeze_challenge();
    return out;
  }
  // transcript.rs:133-136 + hash.rs:19-21
  void common_field_element(const Fr& fe) {
    uint8_t repr[32];
    fe.to_repr(repr);
    state.update(repr, 32);
  }
  // transcript.rs:158-165: absorb LE repr, stream gets the byte-reversed (BE) repr
  void write_field_element(const Fr& fe) {
    uint8_t repr[32];
    fe.to_repr(repr);
    state.update(repr, 32);
    for (int i = 31; i >= 0; --i) stream.push_back(repr[i]);
  }
  void write_field_elements(const Fr* fes, size_t n) {
    for (size_t i = 0; i < n; ++i) write_field_element(fes[i]);
  }
  // transcript.rs:139-156; returns false on a non-canonical encoding
  bool read_field_element(Fr* out) {
    if (rpos + 32 > stream.size()) return false;
    uint8_t repr[32];
    for (int i = 0; i < 32; ++i) repr[i] = stream[rpos + 31 - i];
    rpos += 32;
    uint64_t raw[4];
    memcpy(raw, repr, 32);
    if (Fr::geq_mod(raw)) return false;
    *out = Fr::from_raw(raw);
    common_field_element(*out);
    return true;
  }
  // transcript.rs:171-183; the identity has no coordinates -> error
  bool common_commitment(const G1Affine& p) {
    if (p.is_identity()) return false;
    uint8_t repr[32];
    p.x.to_repr(repr);
    state.update(repr, 32);
    p.y.to_repr(repr);
    state.update(repr, 32);
    return true;
  }
  // transcript.rs:216-227
  bool write_commitment(const G1Affine& p) {
    if (!common_commitment(p)) return false;
    uint8_t repr[32];
    p.x.to_repr(repr);
    for (int i = 31; i >= 0; --i) stream.push_back(repr[i]);
    p.y.to_repr(repr);
    for (int i = 31; i >= 0; --i) stream.push_back(repr[i]);
    return true;
  }
  // transcript.rs:189-211
  bool read_commitment(G1Affine* out) {
    if (rpos + 64 > stream.size()) return false;
    Fq c[2];
    for (int k = 0; k < 2; ++k) {
      uint8_t repr[32];
      for (int i = 0; i < 32; ++i) repr[i] = stream[rpos + 31 - i];
      rpos += 32;
      uint64_t raw[4];
      memcpy(raw, repr, 32);
      if (Fq::geq_mod(raw)) return false;
      c[k] = Fq::from_raw(raw);
    }
    G1Affine p{c[0], c[1]};
    if (p.is_identity() || !p.on_curve()) return false;
    *out = p;
    return common_commitment(p);
  }
};

}  // namespace b200v

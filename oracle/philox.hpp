// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// The reference draws actions with rand 0.7 SmallRng seeded from entropy
// (examples/connect_four.rs:76, coach.rs:138) — not reproducible and not on disk.
// Oracle and product both use Philox-4x32-10 (Salmon et al., SC'11) instead:
//   key     = (lo32(seed), lo32(game_id))
//   counter = (ply, purpose, hi32(seed), hi32(game_id))
//   u       = (x0 >> 8) * 2^-24  in [0,1)
// and a sequential-f32 restatement of SliceRandom::choose_weighted.
#pragma once
#include <cstddef>
#include <cstdint>

namespace azo {

inline void philox4x32_10(const uint32_t ctr_in[4], const uint32_t key_in[2], uint32_t out[4]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
  uint32_t c0 = ctr_in[0], c1 = ctr_in[1], c2 = ctr_in[2], c3 = ctr_in[3];
  uint32_t k0 = key_in[0], k1 = key_in[1];
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = static_cast<uint64_t>(M0) * c0;
    uint64_t p1 = static_cast<uint64_t>(M1) * c2;
    uint32_t n0 = static_cast<uint32_t>(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = static_cast<uint32_t>(p1);
    uint32_t n2 = static_cast<uint32_t>(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = static_cast<uint32_t>(p0);
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

enum : uint32_t { PURPOSE_ACTION = 0, PURPOSE_OPENING = 1 };

inline float uniform01(uint64_t seed, uint64_t game_id, uint32_t ply, uint32_t purpose) {
  uint32_t ctr[4] = {ply, purpose, static_cast<uint32_t>(seed >> 32),
                     static_cast<uint32_t>(game_id >> 32)};
  uint32_t key[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(game_id)};
  uint32_t out[4];
  philox4x32_10(ctr, key, out);
  return static_cast<float>(out[0] >> 8) * (1.0f / 16777216.0f);
}

// choose_weighted (coach.rs:137-138): first index whose running (sequential f32) weight
// sum exceeds u * total; zero-weight entries are never chosen.
inline int choose_weighted(const float* w, size_t n, float u) {
  float total = 0.0f;
  for (size_t i = 0; i < n; ++i) total = total + w[i];
  float t = u * total;
  float acc = 0.0f;
  int last = -1;
  for (size_t i = 0; i < n; ++i) {
    if (w[i] > 0.0f) {
      acc = acc + w[i];
      last = static_cast<int>(i);
      if (t < acc) return last;
    }
  }
  return last;
}

}  // namespace azo

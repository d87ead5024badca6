// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// CPU restatement of the host-side decisions of Coach::learn (/root/reference/src/coach.rs:169-396) and of
// Coach::save_train_examples / the resume in Coach::setup (coach.rs:55-81,159-167).
//
// PARITY UNPINNED BY THE REFERENCE for the byte layout: bincode 1.3.1 and ndarray 0.13 (Cargo.toml:13,24) are not
// under /root/reference and no .examples fixture exists there.  The layout below restates their published formats:
//   bincode::serialize (DefaultOptions of the 1.x free functions): little-endian, fixed-width integers,
//     sequences / VecDeque = u64 length + elements, structs = fields in declaration order, f32 = 4 bytes LE;
//   ndarray 0.13 `impl Serialize for ArrayBase` (array_serde.rs): struct "Array" {v: u8 = ARRAY_FORMAT_VERSION (1),
//     dim: D, data: sequence of elements in logical (row-major) order};  Dim<[usize; N]> serialises its [usize; N]
//     (a serde tuple: no length), Dim<IxDynImpl> serialises its &[usize] (a sequence: u64 length first).
// The shuffle (coach.rs:296-297) uses rand 0.7 SmallRng in the reference — not reproducible; oracle and product both
// use the Fisher-Yates walk of rand's SliceRandom::shuffle with Philox draws (see shuffle_perm).
#pragma once
#include <cstdint>
#include <cstring>
#include <deque>
#include <vector>

#include "philox.hpp"

namespace azo {

struct TrainingSample {            // src/nnet.rs:22-27
  std::vector<float> board;        // ArrayD<f32>, shape below
  std::vector<uint64_t> board_shape;
  std::vector<float> pi;           // Array1<f32>
  float v;
};

struct Bincode {                   // a serde Serializer with bincode 1.x's default options
  std::vector<uint8_t> out;
  void u8(uint8_t x) { out.push_back(x); }
  void u64(uint64_t x) { for (int i = 0; i < 8; ++i) out.push_back(static_cast<uint8_t>(x >> (8 * i))); }
  void f32(float x) { uint32_t u; std::memcpy(&u, &x, 4); for (int i = 0; i < 4; ++i) out.push_back(static_cast<uint8_t>(u >> (8 * i))); }
  void seq_len(uint64_t n) { u64(n); }
};

inline void ser_array_dyn(Bincode& s, const std::vector<uint64_t>& shape, const std::vector<float>& data) {
  s.u8(1);                                          // field "v"
  s.seq_len(shape.size());                          // field "dim": IxDyn -> slice -> sequence
  for (uint64_t d : shape) s.u64(d);
  s.seq_len(data.size());                           // field "data"
  for (float x : data) s.f32(x);
}
inline void ser_array1(Bincode& s, const std::vector<float>& data) {
  s.u8(1);
  s.u64(data.size());                               // Dim<[usize; 1]>: tuple of one, no length
  s.seq_len(data.size());
  for (float x : data) s.f32(x);
}
inline void ser_sample(Bincode& s, const TrainingSample& t) {
  ser_array_dyn(s, t.board_shape, t.board);         // board
  ser_array1(s, t.pi);                              // pi
  s.f32(t.v);                                       // v
}
// coach.rs:163: bincode::serialize(&self.history)
inline std::vector<uint8_t> ser_history(const std::deque<std::deque<TrainingSample>>& h) {
  Bincode s;
  s.seq_len(h.size());
  for (auto& it : h) {
    s.seq_len(it.size());
    for (auto& t : it) ser_sample(s, t);
  }
  return s.out;
}

// coach.rs:274-289 on sample counts: returns how many samples are dropped from the front of the new entry, and
// applies the history window.
struct Window {
  std::deque<uint64_t> sizes;      // history entry sizes, oldest first
};
inline uint64_t push_iteration(Window& w, uint64_t played, uint64_t max_queue_length, uint64_t max_history_length) {
  uint64_t len = played, dropped = 0;
  while (len > max_queue_length) { len -= 1; dropped += 1; }   // :275-277 pop_front
  w.sizes.push_back(len);                                      // :284
  if (w.sizes.size() > max_history_length) w.sizes.pop_front();  // :286-289
  return dropped;
}

// coach.rs:383-390
inline bool accept(uint64_t nwins, uint64_t pwins, float update_threshold) {
  if (pwins + nwins == 0 || static_cast<float>(nwins) / static_cast<float>(pwins + nwins) < update_threshold) return false;
  return true;
}

// rand 0.7 SliceRandom::shuffle: for i in (1..len).rev() { swap(i, gen_range(0, i + 1)) }; the draw is
// floor(x * (i+1) / 2^64) with x the first 64 Philox bits of counter (lo32 i, 2, hi32 seed, hi32 i), key (lo32 seed, lo32 iteration).
inline void shuffle_perm(uint64_t seed, uint64_t iteration, uint64_t n, uint64_t* perm) {
  for (uint64_t i = 0; i < n; ++i) perm[i] = i;
  if (n < 2) return;
  for (uint64_t i = n - 1; i >= 1; --i) {
    uint32_t ctr[4] = {static_cast<uint32_t>(i), 2u, static_cast<uint32_t>(seed >> 32), static_cast<uint32_t>(i >> 32)};
    uint32_t key[2] = {static_cast<uint32_t>(seed), static_cast<uint32_t>(iteration)};
    uint32_t o[4];
    philox4x32_10(ctr, key, o);
    unsigned __int128 x = (static_cast<uint64_t>(o[1]) << 32) | o[0];
    uint64_t j = static_cast<uint64_t>((x * (i + 1)) >> 64);
    uint64_t t = perm[i]; perm[i] = perm[j]; perm[j] = t;
  }
}

}  // namespace azo

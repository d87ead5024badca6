// Node pool layout, packed-counter arithmetic and PUCT (kernel families K2-K4).
//
// Replaces src/node.rs of the reference (Node, NodeStore).  One tree = one pool of 128-byte
// child blocks in HBM plus one open-addressing transposition table:
//
//   block (128 B = one cache line, read once per selection level)
//     slot[a], a = 0..6  {u32 w; f32 q; f32 prior; u32 meta}     the edge "play action a"
//     header (slot 7)    {u16 n[7]; u16 flags}
//
// * (w, n[a]) = the W and N fields of the reference's packed word 0xWWWWWWWWNNNNVVVV
//   (node.rs:17,36) of the node the slot OWNS.  The VL field is not stored: in deterministic
//   mode (one simulation in flight per tree) it is 0 whenever a counter is at rest, and the
//   search reconstructs the full 64-bit word for every update, so carries (N overflowing into
//   W, quirk Q6) behave exactly like the reference's fetch_add/fetch_sub.
// * q = compute_q() (node.rs:51-58) of that word, cached when the word changes (backup), so
//   that selection needs no division for Q.
// * prior = P[a] of the parent: edge data stays on the raw slot (repair F7).
// * meta = what the slot is: INVALID (illegal action), PLACEHOLDER (NodeState::PlaceHolder),
//   a block id (expanded owner: its children live in that block), TERMINAL|code (owner whose
//   game has ended, no children), or LINK (NodeState::Exists(false), a transposition link,
//   node.rs:284-289): then {w, q} hold {owner slot id, owner meta}.
// * slot id = block_id*8 + a (indexes the pool as an array of 16-byte slots).
#pragma once
#include <cstdint>

#include "c4_bitboard.cuh"

#if !defined(__CUDACC__)
#include <cmath>
#define __align__(x) alignas(x)
#endif

namespace azb {

constexpr uint64_t kCounterInit = 0x7FFFFFFF00000000ull;  // node.rs:36
constexpr uint64_t kVisit = 0x0000000000010001ull;        // node.rs:79
constexpr float kEps = 1e-6f;                             // node.rs:12
constexpr float kWinScale = 100.0f;                       // node.rs:13

constexpr uint32_t kMetaInvalid = 0xFFFFFFFFu;
constexpr uint32_t kMetaPlaceholder = 0xFFFFFFFEu;
constexpr uint32_t kMetaLink = 0xFFFFFFFDu;
constexpr uint32_t kMetaTerminal = 0xFFFFFFF0u;  // | code (1,2,3), see terminal_e()
constexpr uint32_t kMaxBlockId = 0xFFFFFF00u;

constexpr uint32_t kFlagHasPolicy = 1u;
constexpr uint32_t kFlagRootHolder = 2u;
constexpr uint32_t kWBias = 0x7FFFFFFFu;  // W field of a fresh node (node.rs:36)

AZB_HD uint64_t counter_pack(uint32_t w, uint32_t n) {
  return (static_cast<uint64_t>(w) << 32) | (static_cast<uint64_t>(n & 0xFFFFu) << 16);
}

AZB_HD bool meta_is_block(uint32_t m) { return m < kMaxBlockId; }
AZB_HD bool meta_is_terminal(uint32_t m) { return (m & 0xFFFFFFFCu) == kMetaTerminal && (m & 3u); }

// e = -get_game_ended(1) stored by upgrade (node.rs:293-294).  code = game_ended_code().
AZB_HD float terminal_e(uint32_t code) {
  return code == 1 ? -1.0f : code == 2 ? 1.0f : -1e-4f;
}

AZB_HD uint32_t counter_n(uint64_t c) { return static_cast<uint32_t>(c >> 16) & 0xFFFFu; }  // node.rs:67-69
AZB_HD uint32_t counter_vl(uint64_t c) { return static_cast<uint32_t>(c) & 0xFFFFu; }       // node.rs:72-74

// Rust `f32 as u32` (node.rs:84): truncate toward zero, saturate, NaN -> 0.
AZB_HD uint32_t f32_as_u32_sat(float x) {
  if (!(x == x) || x <= 0.0f) return 0u;
  if (x >= 4294967296.0f) return 0xFFFFFFFFu;
  return static_cast<uint32_t>(x);
}

#if defined(__CUDA_ARCH__)
#define AZB_FMUL(a, b) __fmul_rn((a), (b))
#define AZB_FADD(a, b) __fadd_rn((a), (b))
#define AZB_FSUB(a, b) __fsub_rn((a), (b))
#define AZB_FDIV(a, b) __fdiv_rn((a), (b))
#define AZB_FSQRT(a) __fsqrt_rn((a))
#else
#define AZB_FMUL(a, b) ((a) * (b))
#define AZB_FADD(a, b) ((a) + (b))
#define AZB_FSUB(a, b) ((a) - (b))
#define AZB_FDIV(a, b) ((a) / (b))
#define AZB_FSQRT(a) sqrtf((a))
#endif

// Node::unvisit (node.rs:83-92).  Q3 literal: a non-negative value subtracts
// (0xFFFFFFFF - incr) << 32, i.e. W += incr + 1; corrected: W += incr.
AZB_HD uint64_t counter_unvisit(uint64_t c, float v, uint32_t quirks) {
  float sv = AZB_FMUL(kWinScale, v);
  uint32_t incr = f32_as_u32_sat(sv < 0.0f ? -sv : sv);
  uint64_t hi;
  if (v < 0.0f) hi = static_cast<uint64_t>(incr) << 32;
  else if (quirks & AZB_Q3_POS_BACKUP_PLUS_ONE) hi = static_cast<uint64_t>(0xFFFFFFFFu - incr) << 32;
  else hi = (0ull - static_cast<uint64_t>(incr)) << 32;
  return c - (1ull | hi);
}

// compute_q (node.rs:51-58) + the PUCT term of best_child (node.rs:352-356), every f32
// operation rounded separately in the reference's order, no FMA.
AZB_HD float puct_u(uint64_t child, float prior, float sqrt_parent, float cpuct_f) {
  uint32_t n = counter_n(child);
  float q = 0.0f;
  if (n > 0) {
    long long w_raw = static_cast<long long>(child >> 32) - 0x7FFFFFFFll;
    float w = AZB_FDIV(static_cast<float>(w_raw), kWinScale);
    q = AZB_FDIV(AZB_FSUB(w, static_cast<float>(counter_vl(child))), static_cast<float>(n));
  }
  float t3 = AZB_FMUL(AZB_FMUL(cpuct_f, prior), sqrt_parent);
  float t4 = static_cast<float>((1u + n) & 0xFFFFu);  // u16 arithmetic (quirk Q6)
  return AZB_FADD(q, AZB_FDIV(t3, t4));
}

// compute_q (node.rs:51-58) of a full counter word.
AZB_HD float counter_q(uint64_t c) {
  uint32_t n = counter_n(c);
  if (n == 0) return 0.0f;
  long long w_raw = static_cast<long long>(c >> 32) - 0x7FFFFFFFll;
  float w = AZB_FDIV(static_cast<float>(w_raw), kWinScale);
  return AZB_FDIV(AZB_FSUB(w, static_cast<float>(counter_vl(c))), static_cast<float>(n));
}

// Transposition table entry (NodeStore.seen, node.rs:135): key -> owner slot + owner meta.
struct __align__(16) HashEntry {
  uint64_t key;  // 0 = empty (a real key always has its 7 column markers set)
  uint32_t slot;
  uint32_t meta;
};

// Fibonacci (multiply-shift) hashing of the 49-bit state key: one 64-bit multiply instead of a splitmix64
// round (19 instructions in the expansion path); collisions only lengthen a probe, never change a result.
AZB_HD uint32_t hash_bucket(uint64_t key, uint32_t bucket_mask) {
  return static_cast<uint32_t>((key * 0x9E3779B97F4A7C15ull) >> 40) & bucket_mask;
}

}  // namespace azb

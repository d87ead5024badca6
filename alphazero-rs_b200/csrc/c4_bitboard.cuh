// Connect-four as bitboards (kernel family K1).
//
// Replaces examples/connect_four_lib/connect_four_game.rs of the reference (array board
// s[6][7] + heights).  Bit index = row*7 + col with row 0 = top (connect_four_game.rs:99),
// so "lowest set bit" == "first in the reference's row-major scan order" (:111-196).
// A canonical state (side to move holds the +1 stones, repair F10) is the pair
//   cur = stones of the side to move, opp = stones of the other side.
#pragma once
#include <cstdint>

#include "../../include/azb200.h"

#if defined(__CUDACC__)
#define AZB_HD __host__ __device__ __forceinline__
#else
#define AZB_HD inline
#endif

namespace azb {

struct BB {
  uint64_t cur, opp;
};

constexpr uint64_t kBoard42 = (1ull << 42) - 1;
constexpr uint64_t kCol0 = 1ull | (1ull << 7) | (1ull << 14) | (1ull << 21) | (1ull << 28) | (1ull << 35);
constexpr uint64_t kTopRow = 0x7Full;

constexpr uint64_t rect_mask(int r0, int r1, int c0, int c1) {  // inclusive ranges
  uint64_t m = 0;
  for (int r = r0; r <= r1; ++r)
    for (int c = c0; c <= c1; ++c) m |= 1ull << (r * 7 + c);
  return m;
}
// Window start cells per direction.  Literal (quirk Q1) = the reference's exclusive ranges
// (connect_four_game.rs:114 cols 0..2, :129 rows 0..1); corrected = cols 0..3 / rows 0..2.
constexpr uint64_t kStartH_lit = rect_mask(0, 5, 0, 2);
constexpr uint64_t kStartH_fix = rect_mask(0, 5, 0, 3);
constexpr uint64_t kStartV_lit = rect_mask(0, 1, 0, 6);
constexpr uint64_t kStartV_fix = rect_mask(0, 2, 0, 6);
constexpr uint64_t kStartD1 = rect_mask(0, 2, 0, 3);  // (+1,+1)  :150-151
constexpr uint64_t kStartD2 = rect_mask(0, 2, 3, 6);  // (+1,-1)  :171-172

AZB_HD uint32_t valid_mask(uint64_t occupied) {  // get_valid_moves, :104-109
  return static_cast<uint32_t>(~occupied) & 0x7Fu;
}

// The cell a stone dropped in column a lands on (:95-99): one above the column's top stone.
AZB_HD uint64_t landing_bit(uint64_t occupied, int a) {
  // all seven landing cells at once: empty cells that sit on the bottom row or right above a stone
  const uint64_t landing = ((occupied >> 7) | (kTopRow << 35)) & ~occupied;
  return landing & (kCol0 << a);
}

// get_next_state(+1, a) followed by get_canonical_form(-1) (async_mcts.rs:284-287 with F4/F10):
// the mover's stone is added, then the sides swap.
AZB_HD BB play_canonical(BB s, int a) {
  uint64_t b = landing_bit(s.cur | s.opp, a);
  return BB{s.opp, s.cur | b};
}

AZB_HD uint64_t line_starts(uint64_t b, int sh, uint64_t start) {
  uint64_t t = b & (b >> sh);
  t &= t >> (2 * sh);
  return t & start;
}

// get_game_ended(+1) on a canonical state (:111-196).  Returns 0 = not ended, 1 = the first
// line in scan order belongs to `cur` (+1.0), 2 = it belongs to `opp` (-1.0), 3 = board full
// without a detected line (DRAW_EPS = 1e-4).
AZB_HD int game_ended_code(BB s, uint32_t quirks) {
  const bool lit = quirks & AZB_Q1_WIN_RANGE_LITERAL;
  const uint64_t sh_h = lit ? kStartH_lit : kStartH_fix;
  const uint64_t sh_v = lit ? kStartV_lit : kStartV_fix;
  uint64_t c, o;
  c = line_starts(s.cur, 1, sh_h); o = line_starts(s.opp, 1, sh_h);
  if (c | o) { uint64_t m = c | o; return (c & (m & (0 - m))) ? 1 : 2; }
  c = line_starts(s.cur, 7, sh_v); o = line_starts(s.opp, 7, sh_v);
  if (c | o) { uint64_t m = c | o; return (c & (m & (0 - m))) ? 1 : 2; }
  c = line_starts(s.cur, 8, kStartD1); o = line_starts(s.opp, 8, kStartD1);
  if (c | o) { uint64_t m = c | o; return (c & (m & (0 - m))) ? 1 : 2; }
  c = line_starts(s.cur, 6, kStartD2); o = line_starts(s.opp, 6, kStartD2);
  if (c | o) { uint64_t m = c | o; return (c & (m & (0 - m))) ? 1 : 2; }
  return ((s.cur | s.opp) == kBoard42) ? 3 : 0;
}

AZB_HD float game_ended_value(int code) {
  return code == 1 ? 1.0f : code == 2 ? -1.0f : code == 3 ? 1e-4f : 0.0f;
}

// Unique 49-bit key of a canonical state: the board sits in rows 1..6 of a 7-row space;
// cur stones are set, and per column the first empty cell (row 0 when full) is marked.
// This is the transposition key: the reference hashes/compares `s` only (:42-54).
AZB_HD uint64_t state_key(BB s) {
  uint64_t occ = s.cur | s.opp;
  uint64_t marker = (occ | (kTopRow << 42)) & ~(occ << 7);
  return (s.cur << 7) | marker;
}

// column mirror (flip(), :65-78)
AZB_HD uint64_t mirror(uint64_t b) {
  uint64_t r = 0;
#pragma unroll
  for (int c = 0; c < 7; ++c) r |= ((b >> c) & kCol0) << (6 - c);
  return r;
}

AZB_HD uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

}  // namespace azb

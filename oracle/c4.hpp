// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// CPU restatement of the reference's connect-four plugin,
//   /root/reference/examples/connect_four_lib/connect_four_game.rs
// in the reference's own data layout (array board s[6][7], per-column heights, `me`).
// Deliberately NOT a bitboard: the CUDA product uses bitboards, so the two
// implementations are independent witnesses of each other.
//
// Repairs: F10 canonical form = board * player, me fixed at +1;
// F11 features are [2,6,7] (C,H,W). Quirk Q1 (scan ranges) is switchable.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

#include "quirks.h"

namespace azo {

constexpr size_t C4_H = 6;        // connect_four_game.rs:13
constexpr size_t C4_W = 7;        // :14
constexpr size_t C4_WIN = 4;      // :15
constexpr float C4_DRAW_EPS = 1e-4f;  // :16

struct C4 {
  int8_t s[C4_H][C4_W];    // row 0 = top (connect_four_game.rs:99)
  size_t heights[C4_W];
  int8_t me;

  // The Game trait has no room for flags; the quirk profile is ambient per thread.
  static uint32_t& quirks() {
    static thread_local uint32_t q = AZO_PROFILE_SANE;
    return q;
  }

  static constexpr size_t num_actions() { return C4_W; }
  static constexpr size_t feature_len() { return 2 * C4_H * C4_W; }

  // connect_four_game.rs:57-63
  static C4 empty() {
    C4 g;
    std::memset(g.s, 0, sizeof(g.s));
    for (auto& h : g.heights) h = 0;
    g.me = 1;
    return g;
  }

  // Build from a raw cell array (ABI boundary); heights recomputed from the cells.
  static C4 from_cells(const int8_t cells[C4_H][C4_W], int8_t me) {
    C4 g = empty();
    std::memcpy(g.s, cells, sizeof(g.s));
    g.me = me;
    for (size_t c = 0; c < C4_W; ++c) {
      size_t h = 0;
      for (size_t r = 0; r < C4_H; ++r) h += (g.s[r][c] != 0);
      g.heights[c] = h;
    }
    return g;
  }

  // connect_four_game.rs:65-78
  C4 flip() const {
    C4 cl = empty();
    for (size_t i = 0; i < C4_H; ++i)
      for (size_t j = 0; j < C4_W; ++j) cl.s[i][j] = s[i][C4_W - j - 1];
    for (size_t j = 0; j < C4_W; ++j) cl.heights[j] = heights[C4_W - j - 1];
    return cl;
  }

  // trait Game ---------------------------------------------------------------
  static C4 get_init_board() { return empty(); }                       // :82-84
  static std::vector<size_t> get_feature_shape() { return {2, C4_H, C4_W}; }  // :86-88

  // :90-102
  std::pair<C4, int8_t> get_next_state(int8_t player, uint8_t action) const {
    C4 next = *this;
    size_t a = action;
    next.heights[a] += 1;
    next.s[C4_H - next.heights[a]][a] = player;
    return {next, static_cast<int8_t>(-player)};
  }

  // :104-109 (player ignored)
  std::array<uint8_t, C4_W> get_valid_moves(int8_t) const {
    std::array<uint8_t, C4_W> v{};
    for (size_t c = 0; c < C4_W; ++c) v[c] = heights[c] < C4_H ? 1 : 0;
    return v;
  }

  // :111-196.  Scan order H, V, diag(+1,+1), diag(+1,-1); first window of four equal
  // non-zero cells decides.  Q1 literal = the reference's exclusive ranges.
  float get_game_ended(int8_t player) const {
    const bool lit = quirks() & AZO_Q1_WIN_RANGE_LITERAL;
    const size_t h_cols = lit ? C4_W - C4_WIN : C4_W - C4_WIN + 1;  // :114  0..3 (excl)  | 0..=3
    const size_t v_rows = lit ? C4_H - C4_WIN : C4_H - C4_WIN + 1;  // :129  0..2 (excl)  | 0..=2
    for (size_t row = 0; row < C4_H; ++row)
      for (size_t col = 0; col < h_cols; ++col) {
        int8_t x = s[row][col];
        if (x != 0 && s[row][col + 1] == x && s[row][col + 2] == x && s[row][col + 3] == x)
          return player == x ? 1.0f : -1.0f;
      }
    for (size_t row = 0; row < v_rows; ++row)
      for (size_t col = 0; col < C4_W; ++col) {
        int8_t x = s[row][col];
        if (x != 0 && s[row + 1][col] == x && s[row + 2][col] == x && s[row + 3][col] == x)
          return player == x ? 1.0f : -1.0f;
      }
    for (size_t row = 0; row <= C4_H - C4_WIN; ++row)          // :145
      for (size_t col = 0; col <= C4_W - C4_WIN; ++col) {      // :146
        int8_t x = s[row][col];
        if (x != 0 && s[row + 1][col + 1] == x && s[row + 2][col + 2] == x &&
            s[row + 3][col + 3] == x)
          return player == x ? 1.0f : -1.0f;
      }
    for (size_t row = 0; row <= C4_H - C4_WIN; ++row)          // :168
      for (size_t col = C4_WIN - 1; col < C4_W; ++col) {       // :169
        int8_t x = s[row][col];
        if (x != 0 && s[row + 1][col - 1] == x && s[row + 2][col - 2] == x &&
            s[row + 3][col - 3] == x)
          return player == x ? 1.0f : -1.0f;
      }
    size_t open = 0;
    for (size_t c = 0; c < C4_W; ++c) open += heights[c] < C4_H;   // :191
    return open == 0 ? C4_DRAW_EPS : 0.0f;
  }

  // :198-203 literal toggles `me` only (F10: unusable).  Repaired: board * player.
  C4 get_canonical_form(int8_t player) const {
    C4 b = *this;
    for (size_t i = 0; i < C4_H; ++i)
      for (size_t j = 0; j < C4_W; ++j) b.s[i][j] = static_cast<int8_t>(s[i][j] * player);
    b.me = 1;
    return b;
  }

  // :205-211 — identity + column mirror, pi reversed.
  std::vector<std::pair<C4, std::array<float, C4_W>>> get_symmetries(
      const std::array<float, C4_W>& pi) const {
    std::array<float, C4_W> rev{};
    for (size_t j = 0; j < C4_W; ++j) rev[j] = pi[C4_W - j - 1];
    return {{*this, pi}, {flip(), rev}};
  }

  float eval_heuristic() const { return 0.0f; }  // :214-216

  // :219-237 repaired per F11: [2,6,7]; ch0 = cells == me, ch1 = cells == -me.
  void to_features(float* out) const {
    for (size_t i = 0; i < C4_H; ++i)
      for (size_t j = 0; j < C4_W; ++j) {
        out[0 * 42 + i * 7 + j] = (s[i][j] == me) ? 1.0f : 0.0f;
        out[1 * 42 + i * 7 + j] = (s[i][j] == -me) ? 1.0f : 0.0f;
      }
  }

  // :42-54 — Hash / Eq on `s` only.
  bool operator==(const C4& o) const { return std::memcmp(s, o.s, sizeof(s)) == 0; }
  size_t hash() const {
    uint64_t h = 1469598103934665603ull;  // FNV-1a over the 42 cells
    const uint8_t* p = reinterpret_cast<const uint8_t*>(s);
    for (size_t i = 0; i < sizeof(s); ++i) {
      h ^= p[i];
      h *= 1099511628211ull;
    }
    return static_cast<size_t>(h);
  }

  // :26-40 Display
  std::string to_string() const {
    std::string r;
    for (size_t row = 0; row < C4_H; ++row) {
      for (size_t col = 0; col < C4_W; ++col)
        r += s[row][col] == 0 ? '_' : (s[row][col] == 1 ? '1' : '2');
      r += '\n';
    }
    return r;
  }
};

}  // namespace azo

// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// CPU restatement of Coach::execute_episode (/root/reference/src/coach.rs:104-157)
// and arena::play_game / play_games (/root/reference/src/arena.rs:7-99).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <functional>
#include <vector>

#include "mcts.hpp"
#include "philox.hpp"

namespace azo {

struct CoachParams {                 // the fields of struct Coach (coach.rs:18-35) the path reads
  size_t mcts_reserve_size = 1000000;
  size_t temp_threshold = 15;
  size_t num_sims = 25;
  size_t max_depth = 1000;
  size_t num_sim_threads = 1;        // coach.rs:32,51
  int32_t cpuct = 1;
  uint32_t quirks = AZO_PROFILE_SANE;
  uint64_t seed = 1;
};

template <class G>
struct EpisodeTrace {
  std::vector<uint8_t> actions;                  // one per ply
  std::vector<std::array<uint16_t, 8>> counts;   // root child visit counts per ply
  std::vector<float> boards;                     // [n_samples, feature_len]
  std::vector<float> pis;                        // [n_samples, A]
  std::vector<float> vs;                         // [n_samples]
  std::vector<int8_t> sample_player;
  float final_r = 0.0f;
  int8_t final_player = 0;
  SearchStats stats;
  size_t nodes_len = 0, seen_len = 0;
};

// coach.rs:104-157.  `episode_id` keys the per-game Philox stream.
// `max_plies` (timing legs only): stop after that many plies of an unfinished game (no labels are produced then).
template <class G>
EpisodeTrace<G> execute_episode(const CoachParams& cp, AsyncMcts<G>& mcts, uint64_t episode_id,
                                size_t max_plies = static_cast<size_t>(-1)) {
  EpisodeTrace<G> tr;
  const size_t A = G::num_actions();
  const size_t F = G::feature_len();
  G board = G::get_init_board();                                   // :113
  int8_t cur_player = 1;                                           // :115
  size_t episode_step = 0;                                         // :117
  for (;;) {
    episode_step += 1;                                             // :120
    G canonical_board = board.get_canonical_form(cur_player);      // :121
    float temp = episode_step < cp.temp_threshold ? 1.0f : 0.0f;   // :123-127
    std::array<uint16_t, 8> cnt{};
    std::vector<float> pi = mcts.get_action_prob(canonical_board, temp, cnt.data());  // :129
    tr.counts.push_back(cnt);
    std::array<float, 7> pia{};
    for (size_t a = 0; a < A; ++a) pia[a] = pi[a];
    for (auto& bp : canonical_board.get_symmetries(pia)) {         // :131-136
      size_t off = tr.boards.size();
      tr.boards.resize(off + F);
      bp.first.to_features(tr.boards.data() + off);
      for (size_t a = 0; a < A; ++a) tr.pis.push_back(bp.second[a]);
      tr.sample_player.push_back(cur_player);
    }
    float u = uniform01(cp.seed, episode_id, static_cast<uint32_t>(episode_step - 1), PURPOSE_ACTION);
    uint8_t action = static_cast<uint8_t>(choose_weighted(pi.data(), A, u));  // :138-139
    tr.actions.push_back(action);
    auto next = board.get_next_state(cur_player, action);          // :141-143
    board = next.first;
    cur_player = next.second;
    float r = board.get_game_ended(cur_player);                    // :145
    if (r != 0.0f) {                                               // :147-156
      tr.final_r = r;
      tr.final_player = cur_player;
      for (int8_t p : tr.sample_player) {
        if (cp.quirks & AZO_Q4_VLABEL_LITERAL)
          tr.vs.push_back(p == cur_player ? 1.0f : -1.0f);         // :153
        else
          tr.vs.push_back(p == cur_player ? r : -r);
      }
      tr.stats = mcts.stats;
      tr.nodes_len = mcts.nodes->size();
      tr.seen_len = mcts.nodes->seen.size();
      return tr;
    }
    if (episode_step >= max_plies) {  // bounded timing sample: the game is cut short
      tr.stats = mcts.stats;
      tr.nodes_len = mcts.nodes->size();
      tr.seen_len = mcts.nodes->seen.size();
      return tr;
    }
  }
}

// arena.rs:7-52.  player_actions[0] moves for cur_player == +1.
template <class G>
int8_t play_game(const std::array<std::function<uint8_t(const G&)>, 2>& player_actions,
                 const G* start, std::vector<uint8_t>* actions_out = nullptr) {
  int8_t cur_player = 1;
  G board = start ? *start : G::get_init_board();
  while (board.get_game_ended(cur_player) == 0.0f) {               // :20
    G canonical_board = board.get_canonical_form(cur_player);      // :27
    uint8_t action = player_actions[cur_player == 1 ? 0 : 1](canonical_board);  // :29
    auto valids = canonical_board.get_valid_moves(1);              // :31
    if (valids[action] == 0) throw std::runtime_error("arena: action is not valid (:33-37)");
    if (actions_out) actions_out->push_back(action);
    auto nx = board.get_next_state(cur_player, action);            // :39-41
    board = nx.first;
    cur_player = nx.second;
  }
  // :51 — f32::round is half-away-from-zero; the draw value 1e-4 rounds to 0.
  return static_cast<int8_t>(cur_player * static_cast<int8_t>(std::round(board.get_game_ended(cur_player))));
}

struct ArenaCounts { size_t win = 0, loss = 0, draw = 0; };        // arena.rs:54-59

// arena.rs:62-99.  `make_players(game_index, seat_order)` returns the two action closures
// in ORIGINAL order (index 0 = the candidate whose wins are counted); Heap's algorithm
// over two elements yields the orders [0,1] then [1,0], num/2 games each.
template <class G>
ArenaCounts play_games(size_t num,
                       const std::function<std::array<std::function<uint8_t(const G&)>, 2>(size_t)>& make_players,
                       const G* start, std::vector<int8_t>* results_out = nullptr) {
  ArenaCounts all;
  size_t game_index = 0;
  for (int ordering = 0; ordering < 2; ++ordering) {
    int8_t win_cond = ordering == 0 ? 1 : -1;                      // :79
    int8_t lose_cond = ordering == 0 ? -1 : 1;                     // :80
    for (size_t g = 0; g < num / 2; ++g, ++game_index) {           // :82
      auto players = make_players(game_index);
      std::array<std::function<uint8_t(const G&)>, 2> seated =
          ordering == 0 ? players
                        : std::array<std::function<uint8_t(const G&)>, 2>{players[1], players[0]};
      int8_t game_result = play_game<G>(seated, start);            // :84
      if (results_out) results_out->push_back(game_result);
      if (game_result == win_cond) all.win++;
      else if (game_result == lose_cond) all.loss++;
      else all.draw++;
    }
  }
  return all;
}

}  // namespace azo

// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// C ABI over the CPU restatement (c4.hpp, node.hpp, mcts.hpp, coach.hpp) so that the
// Python tests, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs can drive it with ctypes.  Nothing in the product links or loads this library.
//
// PARITY PIN: the restatement is pinned against every known-answer vector the reference
// holds for this path (oracle/test_reference_units.cpp: node.rs:393-655 and
// connect_four_game.rs:244-264).  The reference's own code cannot be compiled here (no
// Rust toolchain; nightly features, src/lib.rs:1-3) and holds NO test of search results,
// so visit counts / trajectories are "parity unpinned by the reference": they are pinned
// only by this oracle and by the independent emulation vectors of SURVEY.md App. D
// (tests/golden/).
#include <chrono>
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>
#include <thread>

#include "c4.hpp"
#include "coach.hpp"
#include "learn.hpp"
#include "mcts.hpp"
#include "nnet_cpu.hpp"
#include "node.hpp"
#include "philox.hpp"

using namespace azo;

namespace {
thread_local std::string g_err;

struct CallbackEvaluator : Evaluator {
  typedef void (*fn_t)(const float* boards, size_t batch, float* pi, float* v, void* user);
  fn_t fn;
  void* user;
  void predict(const float* boards, size_t batch, size_t, size_t, float* pi, float* v) override {
    fn(boards, batch, pi, v, user);
  }
};

struct McHandle {
  std::unique_ptr<Evaluator> ev;
  std::unique_ptr<AsyncMcts<C4>> mcts;
  uint32_t quirks;
};

Evaluator* make_eval(int kind, void* fn, void* user) {
  if (kind == AZO_EVAL_UNIFORM) return new UniformEvaluator();
  if (kind == AZO_EVAL_HASH) return new HashEvaluator();
  auto* c = new CallbackEvaluator();
  c->fn = reinterpret_cast<CallbackEvaluator::fn_t>(fn);
  c->user = user;
  return c;
}

C4 from43(const int8_t* p) {
  int8_t cells[C4_H][C4_W];
  std::memcpy(cells, p, 42);
  return C4::from_cells(cells, p[42]);
}
void to43(const C4& g, int8_t* p) {
  std::memcpy(p, g.s, 42);
  p[42] = g.me;
}

// 49-bit unique key of a canonical state (side to move = +1).  Bit (row+1)*7+col is set
// for a +1 stone at (row,col); bit (6-h)*7+col marks the first empty cell of a column of
// height h (row 0 of the 7-row space when the column is full).
uint64_t key_of(const C4& g) {
  uint64_t k = 0;
  for (size_t c = 0; c < C4_W; ++c) {
    k |= 1ull << ((6 - g.heights[c]) * 7 + c);
    for (size_t r = 0; r < C4_H; ++r)
      if (g.s[r][c] == 1) k |= 1ull << ((r + 1) * 7 + c);
  }
  return k;
}
}  // namespace

struct azo_params {
  uint64_t mcts_reserve_size, temp_threshold, num_sims, max_depth;
  int32_t cpuct;
  uint32_t quirks;
  uint64_t seed;
  uint64_t num_sim_threads;  // coach.rs:51 (1 = deterministic mode; K > 1 = waves of K, oracle/mcts.hpp)
};

static CoachParams cp_of(const azo_params* p) {
  CoachParams cp;
  cp.mcts_reserve_size = p->mcts_reserve_size;
  cp.temp_threshold = p->temp_threshold;
  cp.num_sims = p->num_sims;
  cp.max_depth = p->max_depth;
  cp.cpuct = p->cpuct;
  cp.quirks = p->quirks;
  cp.seed = p->seed;
  cp.num_sim_threads = p->num_sim_threads ? p->num_sim_threads : 1;
  return cp;
}

#define AZO_TRY try {
#define AZO_CATCH(ret)                 \
  }                                    \
  catch (const std::exception& e) {    \
    g_err = e.what();                  \
    return ret;                        \
  }

extern "C" {

const char* azo_last_error() { return g_err.c_str(); }

// ---- connect-four (connect_four_game.rs), batched over n 43-byte states -------------
void azo_c4_next_state(const int8_t* in, const int8_t* player, const uint8_t* action, size_t n,
                       int8_t* out, int8_t* next_player) {
  for (size_t i = 0; i < n; ++i) {
    auto nx = from43(in + 43 * i).get_next_state(player[i], action[i]);
    to43(nx.first, out + 43 * i);
    next_player[i] = nx.second;
  }
}
void azo_c4_valid_moves(const int8_t* in, size_t n, uint8_t* out) {
  for (size_t i = 0; i < n; ++i) {
    auto v = from43(in + 43 * i).get_valid_moves(1);
    std::memcpy(out + 7 * i, v.data(), 7);
  }
}
void azo_c4_game_ended(const int8_t* in, const int8_t* player, size_t n, uint32_t quirks,
                       float* out) {
  C4::quirks() = quirks;
  for (size_t i = 0; i < n; ++i) out[i] = from43(in + 43 * i).get_game_ended(player[i]);
}
void azo_c4_canonical_form(const int8_t* in, const int8_t* player, size_t n, int8_t* out) {
  for (size_t i = 0; i < n; ++i) to43(from43(in + 43 * i).get_canonical_form(player[i]), out + 43 * i);
}
void azo_c4_symmetries(const int8_t* in, const float* pi, size_t n, int8_t* out_states,
                       float* out_pi) {
  for (size_t i = 0; i < n; ++i) {
    std::array<float, 7> p;
    std::memcpy(p.data(), pi + 7 * i, 28);
    auto sym = from43(in + 43 * i).get_symmetries(p);
    for (size_t k = 0; k < 2; ++k) {
      to43(sym[k].first, out_states + 43 * (2 * i + k));
      // the reference's flip() builds from empty(): me = +1 (connect_four_game.rs:66)
      std::memcpy(out_pi + 7 * (2 * i + k), sym[k].second.data(), 28);
    }
  }
}
void azo_c4_to_features(const int8_t* in, size_t n, float* out) {
  for (size_t i = 0; i < n; ++i) from43(in + 43 * i).to_features(out + 84 * i);
}
void azo_c4_key(const int8_t* in, size_t n, uint64_t* out) {
  for (size_t i = 0; i < n; ++i) out[i] = key_of(from43(in + 43 * i));
}

// ---- packed counter (node.rs:51-92) ---------------------------------------------------
uint64_t azo_counter_init() { return Node<C4>(WIN_SCALE).win_counter.load(); }
uint64_t azo_counter_visit(uint64_t c) {
  Node<C4> n(WIN_SCALE);
  n.win_counter.store(c);
  n.visit();
  return n.win_counter.load();
}
uint64_t azo_counter_unvisit(uint64_t c, float v, float win_scale, uint32_t quirks) {
  Node<C4> n(win_scale);
  n.win_counter.store(c);
  n.unvisit(v, quirks);
  return n.win_counter.load();
}
void azo_counter_read(uint64_t c, float win_scale, float* w, uint16_t* n_out, uint16_t* vl,
                      float* q) {
  Node<C4> n(win_scale);
  n.win_counter.store(c);
  *w = n.get_w();
  *n_out = n.get_n();
  *vl = n.get_vloss();
  *q = n.compute_q();
}

// ---- philox / choose_weighted ----------------------------------------------------------
float azo_uniform01(uint64_t seed, uint64_t game_id, uint32_t ply, uint32_t purpose) {
  return uniform01(seed, game_id, ply, purpose);
}
int azo_choose_weighted(const float* w, size_t n, float u) { return choose_weighted(w, n, u); }

// ---- AsyncMcts (async_mcts.rs) ----------------------------------------------------------
void* azo_mcts_create(const int8_t* root43, const azo_params* p, int eval_kind, void* fn,
                      void* user) {
  AZO_TRY
  C4::quirks() = p->quirks;
  auto* h = new McHandle();
  h->quirks = p->quirks;
  h->ev.reset(make_eval(eval_kind, fn, user));
  C4 root = root43 ? from43(root43) : C4::get_init_board();
  h->mcts.reset(new AsyncMcts<C4>(root, p->mcts_reserve_size, p->num_sims, p->max_depth, 0,
                                  p->cpuct, p->quirks, h->ev.get()));
  h->mcts->num_threads = p->num_sim_threads ? p->num_sim_threads : 1;
  return h;
  AZO_CATCH(nullptr)
}
void azo_mcts_destroy(void* hv) { delete static_cast<McHandle*>(hv); }

// get_action_prob (async_mcts.rs:74-115): runs num_sims simulations from `state43`
// (canonical) and returns counts + pi.  0 = ok, -1 = the oracle threw (see azo_last_error).
int azo_mcts_get_action_prob(void* hv, const int8_t* state43, float temp, uint16_t* counts,
                             float* pi) {
  AZO_TRY
  auto* h = static_cast<McHandle*>(hv);
  C4::quirks() = h->quirks;
  uint16_t c8[8] = {0};
  auto p = h->mcts->get_action_prob(from43(state43), temp, c8);
  for (int a = 0; a < 7; ++a) {
    counts[a] = c8[a];
    pi[a] = p[a];
  }
  return 0;
  AZO_CATCH(-1)
}
void azo_mcts_set_num_sims(void* hv, uint64_t n) { static_cast<McHandle*>(hv)->mcts->num_sims = n; }
uint64_t azo_mcts_len(void* hv) { return static_cast<McHandle*>(hv)->mcts->nodes->size(); }
uint64_t azo_mcts_seen_len(void* hv) { return static_cast<McHandle*>(hv)->mcts->nodes->seen.size(); }
void azo_mcts_stats(void* hv, uint64_t* out6) {
  const SearchStats& s = static_cast<McHandle*>(hv)->mcts->stats;
  out6[0] = s.sims; out6[1] = s.levels; out6[2] = s.expansions;
  out6[3] = s.terminal_hits; out6[4] = s.dup_links; out6[5] = s.evals;
}
// raw counter of the node that owns `state43`; 0 if the state is not in the tree
uint64_t azo_mcts_counter_of(void* hv, const int8_t* state43) {
  auto* h = static_cast<McHandle*>(hv);
  auto idx = h->mcts->nodes->lookup_state_id(from43(state43));
  if (!idx) return 0;
  return h->mcts->nodes->get(*idx)->win_counter.load();
}
// Whole-tree dump, one row per unique state (`seen`): key, raw counter, e, policy (zeros if
// unset) and a has-policy flag.  Returns the number of rows (<= cap rows are written).
uint64_t azo_mcts_dump(void* hv, uint64_t cap, uint64_t* keys, uint64_t* counters, float* e,
                       float* p7, uint8_t* has_p) {
  auto* h = static_cast<McHandle*>(hv);
  uint64_t i = 0;
  for (const auto& kv : h->mcts->nodes->seen) {
    if (i < cap) {
      const Node<C4>* n = h->mcts->nodes->get(kv.second);
      keys[i] = key_of(kv.first);
      counters[i] = n->win_counter.load();
      e[i] = n->e;
      has_p[i] = n->mu.p ? 1 : 0;
      for (int a = 0; a < 7; ++a) p7[7 * i + a] = n->mu.p ? (*n->mu.p)[a] : 0.0f;
    }
    ++i;
  }
  return i;
}

// ---- Coach::execute_episode (coach.rs:104-157) ------------------------------------------
// Buffers: actions[64], counts[64*7], boards[128*84], pis[128*7], vs[128].
// Returns the number of plies, or -1 on error.
int azo_execute_episode(const azo_params* p, uint64_t episode_id, int eval_kind, void* fn,
                        void* user, uint8_t* actions, uint16_t* counts, float* boards, float* pis,
                        float* vs, uint64_t* n_samples, float* final_r, int8_t* final_player,
                        uint64_t* stats6, uint64_t* nodes_len, uint64_t* seen_len) {
  AZO_TRY
  C4::quirks() = p->quirks;
  CoachParams cp = cp_of(p);
  std::unique_ptr<Evaluator> ev(make_eval(eval_kind, fn, user));
  AsyncMcts<C4> mcts(C4::get_init_board(), cp.mcts_reserve_size, cp.num_sims, cp.max_depth, 0,
                     cp.cpuct, cp.quirks, ev.get());
  mcts.num_threads = cp.num_sim_threads;
  auto tr = execute_episode<C4>(cp, mcts, episode_id);
  size_t plies = tr.actions.size();
  for (size_t i = 0; i < plies; ++i) {
    if (actions) actions[i] = tr.actions[i];
    if (counts)
      for (int a = 0; a < 7; ++a) counts[7 * i + a] = tr.counts[i][a];
  }
  if (boards) std::memcpy(boards, tr.boards.data(), tr.boards.size() * 4);
  if (pis) std::memcpy(pis, tr.pis.data(), tr.pis.size() * 4);
  if (vs) std::memcpy(vs, tr.vs.data(), tr.vs.size() * 4);
  if (n_samples) *n_samples = tr.vs.size();
  if (final_r) *final_r = tr.final_r;
  if (final_player) *final_player = tr.final_player;
  if (stats6) {
    stats6[0] = tr.stats.sims; stats6[1] = tr.stats.levels; stats6[2] = tr.stats.expansions;
    stats6[3] = tr.stats.terminal_hits; stats6[4] = tr.stats.dup_links; stats6[5] = tr.stats.evals;
  }
  if (nodes_len) *nodes_len = tr.nodes_len;
  if (seen_len) *seen_len = tr.seen_len;
  return static_cast<int>(plies);
  AZO_CATCH(-1)
}

// ---- arena (arena.rs:7-99) ---------------------------------------------------------------
// Two MCTS players (evaluator kinds eval_a / eval_b), temp = 0, argmax with last-max ties
// (coach.rs:356-363).  shared_trees = 1: one persistent tree per player for the whole match,
// games sequential (coach.rs:333-372); 0: a fresh tree pair per game (the device layout).
// k_open random opening plies per game come from Philox (seed, game_index, ply, OPENING).
// out_counts = {Win, Loss, Draw} of player A; results[num] (optional) = play_game's i8.
// _cb: the evaluators may be predict callbacks (kind AZO_EVAL_CALLBACK: fn_a / fn_b), game i draws its opening plies
// from Philox stream (seed, first_game_id + i), and the per-game traces come back: actions[num][64] (0xFF padded),
// root_counts[num][64][7] (searched plies only) and plies[num] (all optional).
int azo_arena_play_games_cb(const azo_params* p, uint64_t num, int eval_a, void* fn_a, void* user_a, int eval_b,
                            void* fn_b, void* user_b, int shared_trees, uint32_t k_open, uint64_t first_game_id,
                            uint64_t* out_counts, int8_t* results, uint8_t* actions, uint16_t* root_counts,
                            uint32_t* plies) {
  AZO_TRY
  C4::quirks() = p->quirks;
  CoachParams cp = cp_of(p);
  std::unique_ptr<Evaluator> ea(make_eval(eval_a, fn_a, user_a)), eb(make_eval(eval_b, fn_b, user_b));
  auto mk = [&](Evaluator* e) {
    auto m = std::make_unique<AsyncMcts<C4>>(C4::get_init_board(), cp.mcts_reserve_size, cp.num_sims,
                                             cp.max_depth, 0, cp.cpuct, cp.quirks, e);
    m->num_threads = cp.num_sim_threads;
    return m;
  };
  std::unique_ptr<AsyncMcts<C4>> ta, tb;
  if (shared_trees) { ta = mk(ea.get()); tb = mk(eb.get()); }
  auto argmax = [](const std::vector<float>& pi) {  // coach.rs:356-363: last maximum
    size_t best = 0;
    for (size_t i = 0; i < pi.size(); ++i)
      if (!(pi[best] > pi[i])) best = i;
    return static_cast<uint8_t>(best);
  };
  std::vector<int8_t> res;
  size_t gi = 0;
  ArenaCounts all;
  for (int ordering = 0; ordering < 2; ++ordering) {
    int8_t win_cond = ordering == 0 ? 1 : -1, lose_cond = ordering == 0 ? -1 : 1;
    for (size_t g = 0; g < num / 2; ++g, ++gi) {
      if (!shared_trees) { ta = mk(ea.get()); tb = mk(eb.get()); }
      AsyncMcts<C4>* A = ta.get();
      AsyncMcts<C4>* B = tb.get();
      size_t ply = 0;
      if (actions) std::memset(actions + 64 * gi, 0xFF, 64);
      if (root_counts) std::memset(root_counts + 64 * 7 * gi, 0, 64 * 7 * 2);
      auto mover = [&](AsyncMcts<C4>* t) {
        return [&, t](const C4& s) -> uint8_t {
          size_t my_ply = ply++;
          uint8_t act;
          if (my_ply < k_open) {
            auto v = s.get_valid_moves(1);
            float w[7];
            for (int a = 0; a < 7; ++a) w[a] = v[a] ? 1.0f : 0.0f;
            act = static_cast<uint8_t>(choose_weighted(
                w, 7, uniform01(cp.seed, first_game_id + gi, static_cast<uint32_t>(my_ply), PURPOSE_OPENING)));
          } else {
            uint16_t c8[8] = {0};
            act = argmax(t->get_action_prob(s, 0.0f, c8));
            if (root_counts && my_ply < 64)
              for (int a = 0; a < 7; ++a) root_counts[(64 * gi + my_ply) * 7 + a] = c8[a];
          }
          if (actions && my_ply < 64) actions[64 * gi + my_ply] = act;
          return act;
        };
      };
      std::array<std::function<uint8_t(const C4&)>, 2> seated =
          ordering == 0 ? std::array<std::function<uint8_t(const C4&)>, 2>{mover(A), mover(B)}
                        : std::array<std::function<uint8_t(const C4&)>, 2>{mover(B), mover(A)};
      int8_t r = play_game<C4>(seated, nullptr);
      if (plies) plies[gi] = static_cast<uint32_t>(ply);
      res.push_back(r);
      if (r == win_cond) all.win++;
      else if (r == lose_cond) all.loss++;
      else all.draw++;
    }
  }
  out_counts[0] = all.win; out_counts[1] = all.loss; out_counts[2] = all.draw;
  if (results) std::memcpy(results, res.data(), res.size());
  return 0;
  AZO_CATCH(-1)
}
int azo_arena_play_games(const azo_params* p, uint64_t num, int eval_a, int eval_b,
                         int shared_trees, uint32_t k_open, uint64_t* out_counts,
                         int8_t* results) {
  return azo_arena_play_games_cb(p, num, eval_a, nullptr, nullptr, eval_b, nullptr, nullptr, shared_trees, k_open, 0,
                                 out_counts, results, nullptr, nullptr, nullptr);
}

// ---- CPU timing leg (bench.py cpu_baseline / --impl reference) ---------------------------
// Plays n_games self-play episodes (ids first_game_id ..) over n_threads host threads, one
// independent game per thread at a time (the reference's rayon episode pool,
// coach.rs:202-205,241-272).  Outputs total sims, plies and wall seconds.
int azo_bench_selfplay(const azo_params* p, int eval_kind, uint64_t n_games, uint64_t n_threads,
                       uint64_t first_game_id, uint64_t* sims, uint64_t* plies, double* seconds,
                       uint64_t* levels, uint64_t* expansions) {
  AZO_TRY
  CoachParams cp = cp_of(p);
  std::atomic<uint64_t> next{0}, tot_sims{0}, tot_plies{0}, tot_levels{0}, tot_exp{0};
  std::atomic<int> failed{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (uint64_t t = 0; t < n_threads; ++t)
    th.emplace_back([&] {
      C4::quirks() = cp.quirks;
      try {
        for (;;) {
          uint64_t g = next.fetch_add(1);
          if (g >= n_games) break;
          std::unique_ptr<Evaluator> ev(make_eval(eval_kind, nullptr, nullptr));
          AsyncMcts<C4> mcts(C4::get_init_board(), cp.mcts_reserve_size, cp.num_sims,
                             cp.max_depth, 0, cp.cpuct, cp.quirks, ev.get());
          mcts.num_threads = cp.num_sim_threads;
          auto tr = execute_episode<C4>(cp, mcts, first_game_id + g);
          tot_sims += tr.stats.sims;
          tot_plies += tr.actions.size();
          tot_levels += tr.stats.levels;
          tot_exp += tr.stats.expansions;
        }
      } catch (...) {
        failed = 1;
      }
    });
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (failed) { g_err = "a worker threw"; return -1; }
  *sims = tot_sims; *plies = tot_plies;
  *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (levels) *levels = tot_levels;
  if (expansions) *expansions = tot_exp;
  return 0;
  AZO_CATCH(-1)
}

// The same leg with the network evaluator on the CPU (BASELINE configs 1 and 3): a plain fp32 forward pass of the
// ResNet (oracle/nnet_cpu.hpp, `blocks` residual blocks, the flat parameter vector of azb_nnet_get_params) called
// inline per leaf, one game per thread, each game cut after `max_plies` plies (0 = whole games) so that the sample is
// bounded.  Also returns the number of network evaluations.
int azo_bench_selfplay_net(const azo_params* p, int blocks, const float* params, uint64_t n_params, uint64_t n_games,
                           uint64_t n_threads, uint64_t first_game_id, uint64_t max_plies, uint64_t* sims,
                           uint64_t* plies, double* seconds, uint64_t* evals) {
  AZO_TRY
  if (n_params != CpuNet::num_params(blocks)) { g_err = "parameter count does not match the architecture"; return -1; }
  CoachParams cp = cp_of(p);
  const CpuNet net(blocks, params);
  std::atomic<uint64_t> next{0}, tot_sims{0}, tot_plies{0}, tot_evals{0};
  std::atomic<int> failed{0};
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (uint64_t t = 0; t < n_threads; ++t)
    th.emplace_back([&] {
      C4::quirks() = cp.quirks;
      try {
        for (;;) {
          uint64_t g = next.fetch_add(1);
          if (g >= n_games) break;
          CpuNetEvaluator ev(&net);
          AsyncMcts<C4> mcts(C4::get_init_board(), cp.mcts_reserve_size, cp.num_sims, cp.max_depth, 0, cp.cpuct,
                             cp.quirks, &ev);
          mcts.num_threads = cp.num_sim_threads;
          auto tr = execute_episode<C4>(cp, mcts, first_game_id + g, max_plies ? max_plies : static_cast<size_t>(-1));
          tot_sims += tr.stats.sims;
          tot_plies += tr.actions.size();
          tot_evals += tr.stats.evals;
        }
      } catch (...) {
        failed = 1;
      }
    });
  for (auto& x : th) x.join();
  auto t1 = std::chrono::steady_clock::now();
  if (failed) { g_err = "a worker threw"; return -1; }
  *sims = tot_sims; *plies = tot_plies;
  *seconds = std::chrono::duration<double>(t1 - t0).count();
  if (evals) *evals = tot_evals;
  return 0;
  AZO_CATCH(-1)
}
// NNet::predict of that CPU network (tests: third witness of the network's numerics next to torch and the device)
int azo_cpu_net_predict(int blocks, const float* params, uint64_t n_params, const float* boards, uint64_t batch,
                        float* pi, float* v) {
  AZO_TRY
  if (n_params != CpuNet::num_params(blocks)) { g_err = "parameter count does not match the architecture"; return -1; }
  const CpuNet net(blocks, params);
  for (uint64_t i = 0; i < batch; ++i) net.predict_one(boards + 84 * i, pi + 7 * i, v + i);
  return 0;
  AZO_CATCH(-1)
}

// ---- Coach::learn host decisions and the .examples encoding (oracle/learn.hpp) ----------------------
// counts[n_iters]; boards/pis/vs concatenated; out may be NULL (size query).  Returns the encoded size.
uint64_t azo_examples_encode(uint64_t n_iters, const uint64_t* counts, const float* boards, const float* pis,
                             const float* vs, uint8_t* out, uint64_t cap) {
  std::deque<std::deque<TrainingSample>> h;
  uint64_t at = 0;
  for (uint64_t it = 0; it < n_iters; ++it) {
    std::deque<TrainingSample> q;
    for (uint64_t i = 0; i < counts[it]; ++i, ++at) {
      TrainingSample t;
      t.board.assign(boards + at * 84, boards + at * 84 + 84);
      t.board_shape = {2, 6, 7};
      t.pi.assign(pis + at * 7, pis + at * 7 + 7);
      t.v = vs[at];
      q.push_back(std::move(t));
    }
    h.push_back(std::move(q));
  }
  std::vector<uint8_t> enc = ser_history(h);
  if (out && enc.size() <= cap) std::memcpy(out, enc.data(), enc.size());
  return enc.size();
}
int azo_learn_accept(uint64_t nwins, uint64_t pwins, float thr) { return accept(nwins, pwins, thr) ? 1 : 0; }
void azo_learn_shuffle_perm(uint64_t seed, uint64_t iteration, uint64_t n, uint64_t* perm) {
  shuffle_perm(seed, iteration, n, perm);
}
// plays `n` iterations of sample counts through the queue trim + history window; out_sizes[n][max_hist] (0 padded),
// out_dropped[n].
void azo_learn_window(const uint64_t* played, uint64_t n, uint64_t max_queue, uint64_t max_hist, uint64_t* out_sizes,
                      uint64_t* out_dropped) {
  Window w;
  for (uint64_t i = 0; i < n; ++i) {
    out_dropped[i] = push_iteration(w, played[i], max_queue, max_hist);
    for (uint64_t k = 0; k < max_hist; ++k) out_sizes[i * max_hist + k] = k < w.sizes.size() ? w.sizes[k] : 0;
  }
}

}  // extern "C"

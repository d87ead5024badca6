// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// CPU restatement of /root/reference/src/node.rs: the packed 64-bit statistics word,
// the Node, and the NodeStore (pre-sized arena + transposition map + lock flags + PUCT).
// Same data structures as the reference: per-node child vector, optional policy/valid
// vectors, optional state, hash-map `seen`.  Generic over the Game like the reference.
#pragma once
#include <atomic>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <unordered_map>
#include <vector>

#include "quirks.h"

namespace azo {

constexpr float EPS = 1e-6f;          // node.rs:12
constexpr float WIN_SCALE = 100.0f;   // node.rs:13

// Rust `f32 as u32`: truncates toward zero, saturates, NaN -> 0 (node.rs:84).
inline uint32_t rust_f32_as_u32(float x) {
  if (!(x == x)) return 0u;
  if (x <= 0.0f) return 0u;
  if (x >= 4294967296.0f) return 0xFFFFFFFFu;
  return static_cast<uint32_t>(x);
}

template <class G>
struct NodeMutableState {            // node.rs:26-31
  std::optional<std::vector<float>> p;
  std::optional<std::vector<uint8_t>> v;
  std::optional<G> s;
};

template <class G>
struct Node {                        // node.rs:16-24
  std::atomic<uint64_t> win_counter; // 0xWWWWWWWWNNNNVVVV
  float win_scale;
  uint8_t a;
  float e;
  NodeMutableState<G> mu;
  std::vector<size_t> children;

  explicit Node(float scale)         // Node::empty, node.rs:34-49
      : win_counter(0x7FFFFFFF00000000ull), win_scale(scale), a(0), e(0.0f) {}
  Node(const Node& o)                // Clone, node.rs:95-107
      : win_counter(o.win_counter.load(std::memory_order_acquire)),
        win_scale(o.win_scale), a(o.a), e(o.e), mu(o.mu), children(o.children) {}
  Node& operator=(const Node& o) {
    win_counter.store(o.win_counter.load());
    win_scale = o.win_scale; a = o.a; e = o.e; mu = o.mu; children = o.children;
    return *this;
  }

  float get_w() const {              // node.rs:61-64
    uint64_t c = win_counter.load(std::memory_order_acquire);
    int64_t w = static_cast<int64_t>(c >> 32) - 0x7FFFFFFFll;
    return static_cast<float>(w) / win_scale;
  }
  uint16_t get_n() const {           // node.rs:67-69
    return static_cast<uint16_t>((win_counter.load(std::memory_order_acquire) &
                                  0x00000000FFFF0000ull) >> 16);
  }
  uint16_t get_vloss() const {       // node.rs:72-74
    return static_cast<uint16_t>(win_counter.load(std::memory_order_acquire) &
                                 0x000000000000FFFFull);
  }
  float compute_q() const {          // node.rs:51-58
    uint16_t n = get_n();
    if (n > 0) return (get_w() - static_cast<float>(get_vloss())) / static_cast<float>(n);
    return 0.0f;
  }
  void visit() {                     // node.rs:77-80
    win_counter.fetch_add(0x0000000000010001ull, std::memory_order_seq_cst);
  }
  // node.rs:83-92.  Q3 literal: v >= 0 (incl. -0.0) subtracts (0xFFFFFFFF - incr) << 32,
  // i.e. W += incr + 1.  Corrected: W += incr exactly.
  void unvisit(float win_val, uint32_t quirks = AZO_PROFILE_REFERENCE) {
    uint32_t incr32 = rust_f32_as_u32(std::fabs(win_scale * win_val));
    uint64_t incr;
    if (win_val < 0.0f) {
      incr = static_cast<uint64_t>(incr32) << 32;
    } else if (quirks & AZO_Q3_POS_BACKUP_PLUS_ONE) {
      incr = static_cast<uint64_t>(0xFFFFFFFFu - incr32) << 32;
    } else {
      incr = (0ull - static_cast<uint64_t>(incr32)) << 32;  // == +incr in the W field
    }
    win_counter.fetch_sub(0x0000000000000001ull | incr, std::memory_order_seq_cst);
  }
};

enum class NodeState { PlaceHolder, Locked, ExistsOwner, ExistsLink };  // node.rs:138-143

template <class G>
struct GHash {
  size_t operator()(const G& g) const { return g.hash(); }
};

template <class G>
class NodeStore {                    // node.rs:132-136
 public:
  struct Cell {
    std::atomic<bool> flag{false};
    std::optional<std::pair<Node<G>, std::optional<size_t>>> cell;  // (Node, link)
  };

 private:
  std::unique_ptr<Cell[]> buf_;
  size_t cap_;

 public:
  std::atomic<size_t> len{0};
  std::unordered_map<G, size_t, GHash<G>> seen;

  // NodeStore::empty, node.rs:146-154
  explicit NodeStore(size_t reserve_space)
      : buf_(new Cell[reserve_space]), cap_(reserve_space) {}

  // NodeStore::new / from_root, node.rs:156-177
  static std::unique_ptr<NodeStore> with_root(size_t reserve_space, const G& s) {
    auto ns = std::make_unique<NodeStore>(reserve_space);
    size_t root_idx = ns->push(Node<G>(WIN_SCALE));
    ns->upgrade(root_idx, s);
    return ns;
  }

  size_t size() const { return len.load(std::memory_order_acquire); }  // node.rs:372-374
  size_t capacity() const { return cap_; }

  // node.rs:179-193
  std::optional<size_t> resolve(size_t idx) const {
    size_t l = idx;
    size_t n = size();
    for (;;) {
      if (l >= n) return std::nullopt;
      const auto& c = buf_[l].cell;
      if (!c) return std::nullopt;
      if (!c->second) return l;
      l = *c->second;
    }
  }
  // node.rs:195-201
  Node<G>* get(size_t idx) const {
    auto l = resolve(idx);
    return l ? &buf_[*l].cell->first : nullptr;
  }
  // raw slot (no link resolution) — used by repair F7 (edge action lives on the slot)
  Node<G>* raw(size_t idx) const { return &buf_[idx].cell->first; }

  std::optional<size_t> lookup_state_id(const G& s) const {  // node.rs:203-205
    auto it = seen.find(s);
    if (it == seen.end()) return std::nullopt;
    return it->second;
  }

  // node.rs:212-232 — only a Locked node accepts a policy
  bool set_policy(size_t idx, std::vector<float> policy) {
    auto l = resolve(idx);
    if (state(idx) != NodeState::Locked) return false;
    buf_[*l].cell->first.mu.p = std::move(policy);
    return true;
  }

  // node.rs:234-244 — lock-free bump allocation
  size_t push(const Node<G>& node) {
    size_t idx = len.fetch_add(1, std::memory_order_seq_cst);
    if (idx >= cap_) throw std::runtime_error("NodeStore: capacity exceeded (node.rs:237)");
    buf_[idx].cell.emplace(node, std::nullopt);
    return idx;
  }

  // node.rs:246-270
  std::optional<NodeState> state(size_t idx) const {
    if (idx >= size()) return std::nullopt;
    if (buf_[idx].flag.load(std::memory_order_seq_cst)) return NodeState::Locked;
    const auto& c = buf_[idx].cell;
    if (!c) return std::nullopt;
    if (!c->second) return c->first.mu.s ? NodeState::ExistsOwner : NodeState::PlaceHolder;
    return NodeState::ExistsLink;
  }

  // node.rs:272-326.  true = became a real (owner) node and stays locked if it was;
  // false = state already seen: slot becomes a link and is unlocked.
  std::optional<bool> upgrade(size_t idx, const G& s) {
    if (idx >= size()) return std::nullopt;
    auto& cell = *buf_[idx].cell;
    assert(!cell.second);
    auto existing = seen.find(s);
    if (existing != seen.end()) {
      cell.second = existing->second;
      unlock(idx);
      return false;
    }
    Node<G>& old_node = cell.first;
    old_node.mu.s = s;
    float game_ended = s.get_game_ended(1);
    old_node.e = -game_ended;
    if (game_ended == 0.0f) {
      auto valids = s.get_valid_moves(1);
      std::vector<uint8_t> valid_actions;
      for (size_t i = 0; i < valids.size(); ++i)
        if (valids[i] != 0) valid_actions.push_back(static_cast<uint8_t>(i));
      old_node.children.reserve(valid_actions.size());
      old_node.mu.v = std::vector<uint8_t>(valids.begin(), valids.end());
      for (uint8_t a : valid_actions) {
        Node<G> child(WIN_SCALE);
        child.a = a;
        old_node.children.push_back(push(child));
      }
    }
    seen.emplace(s, idx);
    return true;
  }

  bool lock(size_t idx) {            // node.rs:328-333
    bool expected = false;
    return buf_[idx].flag.compare_exchange_strong(expected, true, std::memory_order_seq_cst);
  }
  void unlock(size_t idx) {          // node.rs:335-341 (debug_assert only)
    bool expected = true;
    buf_[idx].flag.compare_exchange_strong(expected, false, std::memory_order_seq_cst);
  }

  // node.rs:343-370.  Repair F7: the edge's action/prior index comes from the RAW child
  // slot; the statistics from the resolved node.  `max_by` keeps the LAST maximum; an
  // unordered comparison (NaN) counts as Equal, i.e. the later element wins.
  size_t best_child(size_t idx, int32_t cpuct, bool filter) const {
    const Node<G>* node = get(idx);
    uint16_t parent_n = node->get_n();
    bool have = false;
    size_t best_idx = 0;
    float best_u = 0.0f;
    for (size_t child_idx : node->children) {
      const Node<G>* child = get(child_idx);
      uint8_t a = raw(child_idx)->a;
      float u = child->compute_q() +
                static_cast<float>(cpuct) * (*node->mu.p)[a] *
                    std::sqrt(static_cast<float>(parent_n) + EPS) /
                    static_cast<float>(static_cast<uint16_t>(1 + child->get_n()));
      if (filter && state(child_idx) == NodeState::Locked) continue;
      if (!have || !(best_u > u)) {
        have = true;
        best_idx = child_idx;
        best_u = u;
      }
    }
    if (!have) throw std::runtime_error("best_child: no selectable child (node.rs:367)");
    return best_idx;
  }
};

// /root/reference/src/node/tests/dummy_game.rs — 1-byte fake Game for the store tests.
struct DummyGame {
  uint8_t _s;
  explicit DummyGame(uint8_t v = 0) : _s(v) {}
  static DummyGame get_init_board() { return DummyGame(0); }
  static std::vector<size_t> get_feature_shape() { return {1}; }
  std::pair<DummyGame, int8_t> get_next_state(int8_t player, uint8_t) const {
    return {DummyGame(static_cast<uint8_t>(_s + 1)), static_cast<int8_t>(1 - player)};
  }
  std::vector<uint8_t> get_valid_moves(int8_t) const { return {0}; }
  float get_game_ended(int8_t) const { return 0.0f; }
  DummyGame get_canonical_form(int8_t) const { return DummyGame(0); }
  float eval_heuristic() const { return 0.0f; }
  bool operator==(const DummyGame& o) const { return _s == o._s; }
  size_t hash() const { return _s; }
};

}  // namespace azo

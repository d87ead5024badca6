// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// The reference's own unit tests, restated against the oracle so that the oracle is
// pinned by every known-answer vector the reference holds for this path:
//   /root/reference/src/node.rs:393-655          (counter KATs, store, link, lock)
//   /root/reference/examples/connect_four_lib/connect_four_game.rs:244-264 (diagonal win)
// The rayon tests (node.rs:488-549) are restated with std::thread.
// Exit code 0 = all pass; each failure prints the reference test name.
#include <cmath>
#include <cstdio>
#include <thread>
#include <unordered_map>
#include <vector>

#include "c4.hpp"
#include "node.hpp"

using namespace azo;

static int failures = 0;
#define CHECK(name, cond)                                                         \
  do {                                                                            \
    if (!(cond)) {                                                                \
      std::printf("FAIL %s: %s (line %d)\n", name, #cond, __LINE__);              \
      ++failures;                                                                 \
    }                                                                             \
  } while (0)

static bool similar(float a, float b, float eps) { return std::fabs(a - b) < eps; }

// DummyGame needs the two statics AsyncMcts would ask of a Game; the store never does.
static void test_win() {  // node.rs:393-415
  Node<DummyGame> node(10000.0f);
  CHECK("test_win", similar(node.get_w(), 0.0f, 1e-3f));
  CHECK("test_win", node.get_n() == 0);
  CHECK("test_win", node.get_vloss() == 0);
  node.visit();
  CHECK("test_win", similar(node.get_w(), 0.0f, 1e-3f));
  CHECK("test_win", node.get_n() == 1);
  CHECK("test_win", node.get_vloss() == 1);
  node.unvisit(1.0f);
  CHECK("test_win", similar(node.get_w(), 1.0f, 1e-3f));
  CHECK("test_win", node.get_n() == 1);
  CHECK("test_win", node.get_vloss() == 0);
  // SURVEY App. B.2 emulation witness: the literal +1 slip
  CHECK("test_win/raw", (node.win_counter.load() >> 32) - 0x7FFFFFFFull == 10001ull);
}

static void test_loss() {  // node.rs:417-426
  Node<DummyGame> node(10000.0f);
  node.visit();
  node.unvisit(-1.0f);
  CHECK("test_loss", similar(node.get_w(), -1.0f, 1e-3f));
  CHECK("test_loss", node.get_n() == 1);
  CHECK("test_loss", node.get_vloss() == 0);
}

static void test_winloss() {  // node.rs:428-440
  Node<DummyGame> node(10000.0f);
  node.visit();
  node.unvisit(-1.0f);
  node.visit();
  CHECK("test_winloss", node.get_vloss() == 1);
  node.unvisit(1.0f);
  CHECK("test_winloss", similar(node.get_w(), 0.0f, 1e-3f));
  CHECK("test_winloss", node.get_n() == 2);
  CHECK("test_winloss", node.get_vloss() == 0);
  CHECK("test_winloss/raw", (node.win_counter.load() >> 32) - 0x7FFFFFFFull == 1ull);
}

static void test_nodestore_empty() {  // node.rs:447-451
  NodeStore<DummyGame> nodes(2048);
  CHECK("test_nodestore_empty", nodes.size() == 0);
}

static void test_nodestore_one() {  // node.rs:453-468
  NodeStore<DummyGame> nodes(2048);
  Node<DummyGame> node(10000.0f);
  node.mu.p = std::vector<float>(10, 0.0f);
  node.mu.v = std::vector<uint8_t>(10, 0);
  node.mu.s = DummyGame(0);
  size_t idx = nodes.push(node);
  Node<DummyGame>* got = nodes.get(idx);
  CHECK("test_nodestore_one", got != nullptr && got->mu.s && *got->mu.s == *node.mu.s);
}

static void test_nodestore_many() {  // node.rs:470-486
  NodeStore<DummyGame> nodes(8192);
  std::vector<size_t> idx;
  for (int i = 0; i < 8192; ++i) {
    Node<DummyGame> node(10000.0f);
    node.e = static_cast<float>(i);
    idx.push_back(nodes.push(node));
  }
  CHECK("test_nodestore_many", nodes.size() == 8192);
  bool ok = true;
  for (int i = 0; i < 8192; ++i) ok &= static_cast<float>(idx[i]) == nodes.get(idx[i])->e;
  CHECK("test_nodestore_many", ok);
}

static void parallel_push(const char* name, size_t n, bool probe) {  // node.rs:488-549
  NodeStore<DummyGame> nodes(n);
  std::vector<size_t> idx(n);
  const size_t T = 8;
  std::vector<std::thread> th;
  for (size_t t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      for (size_t i = t; i < n; i += T) {
        Node<DummyGame> node(10000.0f);
        node.e = static_cast<float>(i);
        if (probe && i > 32 && nodes.state(i - 32) == NodeState::PlaceHolder)
          if (nodes.get(i - 32) == nullptr) std::printf("FAIL %s: get during push\n", name);
        idx[i] = nodes.push(node);
      }
    });
  for (auto& x : th) x.join();
  CHECK(name, nodes.size() == n);
  bool ok = true;
  for (size_t i = 0; i < n; ++i) ok &= static_cast<float>(i) == nodes.get(idx[i])->e;
  CHECK(name, ok);
}

static void test_nodestore_upgrade_many_similar() {  // node.rs:551-589
  NodeStore<DummyGame> nodes(8192);
  DummyGame s(0);
  int uniques = 0;
  size_t root = ~size_t(0);
  for (int i = 0; i < 8192; ++i) {
    size_t idx = nodes.push(Node<DummyGame>(10000.0f));
    CHECK("test_nodestore_upgrade_many_similar", nodes.lock(idx));
    bool unique = *nodes.upgrade(idx, s);
    if (unique) {
      CHECK("test_nodestore_upgrade_many_similar", nodes.state(idx) == NodeState::Locked);
      nodes.unlock(idx);
      ++uniques;
      root = idx;
    }
  }
  CHECK("test_nodestore_upgrade_many_similar", nodes.size() == 8192);
  CHECK("test_nodestore_upgrade_many_similar", uniques == 1);
  CHECK("test_nodestore_upgrade_many_similar", root == 0);
}

static void test_nodestore_upgrade() {  // node.rs:591-632
  NodeStore<DummyGame> nodes(2048);
  Node<DummyGame> node(10000.0f);
  size_t idx = nodes.push(node);
  DummyGame s(0);
  Node<DummyGame>* const_ref = nodes.get(idx);
  CHECK("test_nodestore_upgrade", !nodes.get(idx)->mu.s);
  CHECK("test_nodestore_upgrade", nodes.lock(0));
  CHECK("test_nodestore_upgrade", *nodes.upgrade(0, s));
  CHECK("test_nodestore_upgrade", nodes.state(0) == NodeState::Locked);
  CHECK("test_nodestore_upgrade", const_ref == nodes.get(idx));  // address stable
  nodes.unlock(0);
  CHECK("test_nodestore_upgrade", nodes.state(0) == NodeState::ExistsOwner);
  CHECK("test_nodestore_upgrade", *nodes.get(idx)->mu.s == s);
  CHECK("test_nodestore_upgrade", nodes.seen.count(s) == 1);
  CHECK("test_nodestore_upgrade", nodes.seen.size() == 1);
  nodes.push(node);
  CHECK("test_nodestore_upgrade", nodes.lock(1));
  CHECK("test_nodestore_upgrade", !*nodes.upgrade(1, s));
  CHECK("test_nodestore_upgrade", nodes.state(1) == NodeState::ExistsLink);
}

static void test_nodestore_lock() {  // node.rs:634-655
  NodeStore<DummyGame> nodes(2048);
  size_t idx = nodes.push(Node<DummyGame>(10000.0f));
  CHECK("test_nodestore_lock", nodes.state(idx) == NodeState::PlaceHolder);
  DummyGame s(0);
  CHECK("test_nodestore_lock", nodes.lock(idx));
  nodes.upgrade(idx, s);
  CHECK("test_nodestore_lock", !nodes.lock(idx));
  CHECK("test_nodestore_lock", nodes.state(idx) == NodeState::Locked);
  nodes.unlock(idx);
  CHECK("test_nodestore_lock", nodes.state(idx) == NodeState::ExistsOwner);
}

static void test_win_diagonal() {  // connect_four_game.rs:244-264
  for (uint32_t q : {AZO_PROFILE_REFERENCE, AZO_PROFILE_SANE}) {
    C4::quirks() = q;
    C4 board = C4::empty();
    int8_t player = 1;
    for (uint8_t a : {0, 1, 1, 2, 0, 2, 2, 3, 3, 3, 3}) {
      auto nx = board.get_next_state(player, a);
      board = nx.first;
      player = nx.second;
    }
    CHECK("test_win_diagonal", board.get_game_ended(1) == 1.0f);
    // SURVEY App. D.1 final position
    CHECK("test_win_diagonal/board",
          board.to_string() == "_______\n_______\n___1___\n__12___\n1121___\n1222___\n");
  }
  C4::quirks() = AZO_PROFILE_SANE;
}

int main() {
  test_win();
  test_loss();
  test_winloss();
  CHECK("test_is_lockfree", std::atomic<size_t>{}.is_lock_free());  // node.rs:442-445
  test_nodestore_empty();
  test_nodestore_one();
  test_nodestore_many();
  parallel_push("test_nodestore_some_parallel", 1024, false);
  parallel_push("test_nodestore_many_parallel", 8192, false);
  parallel_push("test_nodestore_parallel_push_then_get", 8192, true);
  test_nodestore_upgrade_many_similar();
  test_nodestore_upgrade();
  test_nodestore_lock();
  test_win_diagonal();
  if (failures == 0) std::printf("OK 14 reference unit tests restated\n");
  return failures == 0 ? 0 : 1;
}

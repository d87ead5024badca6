// ORACLE — TEST INFRASTRUCTURE ONLY (see oracle/quirks.h).
//
// CPU restatement of /root/reference/src/async_mcts.rs in its deterministic mode
// (num_threads = 1), with the repair list F1-F9, F12 of SURVEY.md App. A applied
// exactly as the pseudocode of App. C.  Each repair is marked where it happens.
// The evaluator is called inline (the reference's inference thread, :117-189,
// reduces to "predict one row" when batch_size = 1).
#pragma once
#include <array>
#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <vector>

#include "node.hpp"
#include "quirks.h"

namespace azo {

// Mirrors trait NNet::predict (nnet.rs:40-44): boards [B, feature_len] -> (pi [B,A], v [B]).
struct Evaluator {
  virtual ~Evaluator() = default;
  virtual void predict(const float* boards, size_t batch, size_t feature_len, size_t num_actions,
                       float* pi, float* v) = 0;
};

// examples/connect_four.rs:26-42 DumbConnectFourNnet: pi = 1/width, v = 1.
struct UniformEvaluator : Evaluator {
  void predict(const float*, size_t batch, size_t, size_t num_actions, float* pi,
               float* v) override {
    for (size_t b = 0; b < batch; ++b) {
      for (size_t a = 0; a < num_actions; ++a)
        pi[b * num_actions + a] = 1.0f / static_cast<float>(num_actions);
      v[b] = 1.0f;
    }
  }
};

inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// SURVEY.md App. B.6 "hash evaluator" (not in the reference): integer hash of the two
// feature planes -> exactly representable f32 priors and value, so CPU and GPU agree
// bit for bit without transcendental functions.  Plane bit index = row*7 + col.
struct HashEvaluator : Evaluator {
  void predict(const float* boards, size_t batch, size_t feature_len, size_t num_actions,
               float* pi, float* v) override {
    for (size_t b = 0; b < batch; ++b) {
      const float* f = boards + b * feature_len;
      uint64_t mine = 0, theirs = 0;
      for (size_t i = 0; i < 42; ++i) {
        if (f[i] != 0.0f) mine |= 1ull << i;
        if (f[42 + i] != 0.0f) theirs |= 1ull << i;
      }
      uint64_t h = splitmix64(splitmix64(mine) + theirs);
      for (size_t a = 0; a < num_actions; ++a)
        pi[b * num_actions + a] = static_cast<float>(1u + static_cast<uint32_t>((h >> (8 * a)) & 0xFF));
      v[b] = static_cast<float>((h >> 56) & 0xFF) / 128.0f - 1.0f;
    }
  }
};

struct SearchStats {
  uint64_t sims = 0, levels = 0, expansions = 0, terminal_hits = 0, dup_links = 0, evals = 0;
};

template <class G>
class AsyncMcts {                      // async_mcts.rs:14-24
 public:
  std::unique_ptr<NodeStore<G>> nodes;
  size_t num_sims, max_depth, model_id;
  int32_t cpuct;
  uint32_t quirks;
  Evaluator* nnet;
  SearchStats stats;

  // AsyncMcts::default (:27-48) / from_state (:50-72)
  AsyncMcts(const G& root_state, size_t reserve_space, size_t num_sims_, size_t max_depth_,
            size_t model_id_, int32_t cpuct_, uint32_t quirks_, Evaluator* nnet_)
      : nodes(NodeStore<G>::with_root(reserve_space, root_state)),
        num_sims(num_sims_), max_depth(max_depth_), model_id(model_id_), cpuct(cpuct_),
        quirks(quirks_), nnet(nnet_) {}

  // predict + mask + renormalise + set_policy (:305-348); returns v.
  float evaluate(size_t idx) {
    Node<G>* n = nodes->get(idx);
    const size_t A = G::num_actions();
    std::vector<float> feat(G::feature_len());
    n->mu.s->to_features(feat.data());
    std::vector<float> pi(A);
    float v = 0.0f;
    nnet->predict(feat.data(), 1, feat.size(), A, pi.data(), &v);
    const std::vector<uint8_t>& valids = *n->mu.v;
    for (size_t a = 0; a < A; ++a)
      if (valids[a] == 0) pi[a] = 0.0f;                        // :320-324
    float sum_ps = 0.0f;
    for (size_t a = 0; a < A; ++a) sum_ps = sum_ps + pi[a];    // ndarray sum, len 7: sequential
    if (sum_ps > 0.0f) {
      for (size_t a = 0; a < A; ++a) pi[a] = pi[a] / sum_ps;   // :328-329
    } else {
      for (size_t a = 0; a < A; ++a) pi[a] += static_cast<float>(valids[a]);  // :338-340
      float s2 = 0.0f;
      for (size_t a = 0; a < A; ++a) s2 = s2 + pi[a];
      for (size_t a = 0; a < A; ++a) pi[a] = pi[a] / s2;       // :342
    }
    // set_policy only accepts a Locked node (node.rs:215); the F1 path evaluates an
    // existing unlocked node, so take the lock around it.
    bool was_locked = nodes->state(idx) == NodeState::Locked;
    if (!was_locked) nodes->lock(idx);
    nodes->set_policy(idx, std::move(pi));                     // :348
    nodes->unlock(idx);                                        // :351
    stats.evals++;
    return v;
  }

  // search_iteration (:219-371) as repaired in SURVEY App. C.
  void search_iteration(size_t root_idx) {
    size_t cur = root_idx;
    std::vector<size_t> node_path;
    node_path.reserve(64);
    size_t depth = 0;
    float v;
    for (;;) {
      Node<G>* n = nodes->get(cur);                            // :238 (resolves links)
      stats.levels++;
      if (depth > max_depth) {                                 // :240-243
        n->visit();                                            // F6
        v = n->mu.s->eval_heuristic();
        break;
      }
      float e = n->e;                                          // :245
      if (e != 0.0f) {                                         // :246-248
        n->visit();                                            // F6
        stats.terminal_hits++;
        v = e;
        break;
      }
      if (!n->mu.p) {                                          // F1: existing node, never evaluated
        n->visit();
        v = -evaluate(*nodes->resolve(cur));
        break;
      }
      n->visit();                                              // :250
      size_t c = nodes->best_child(cur, cpuct, false);         // :254-257
      auto st = nodes->state(c);
      if (st == NodeState::PlaceHolder) {                      // :260-268
        if (!nodes->lock(c)) throw std::runtime_error("lock failed in single-thread mode");
        node_path.push_back(cur);                              // F3
        const Node<G>* parent = n;
        uint8_t a = nodes->raw(c)->a;                          // F4: the placeholder's own action
        cur = c;
        auto nx = parent->mu.s->get_next_state(1, a);          // :284
        G s2 = nx.first.get_canonical_form(nx.second);         // :287 (F10)
        auto up = nodes->upgrade(cur, s2);                     // :289
        if (!up) throw std::runtime_error("Upgraded invalid node! (:291)");
        if (!*up) {                                            // :293-299 duplicate: continue from owner
          stats.dup_links++;
          cur = *nodes->resolve(cur);
          continue;                                            // depth not incremented
        }
        stats.expansions++;
        Node<G>* m = nodes->get(cur);
        m->visit();                                            // :309
        if (m->e != 0.0f) {                                    // F5: terminal leaf skips the net
          nodes->unlock(cur);
          stats.terminal_hits++;
          v = m->e;
          break;
        }
        v = -evaluate(cur);                                    // :311-353
        break;
      } else {                                                 // :269-274 + F2
        node_path.push_back(cur);
        cur = c;
        depth += 1;
      }
    }
    float sign = 1.0f;                                         // :361-370
    for (;;) {
      nodes->get(cur)->unvisit(sign * v, quirks);
      if (!(quirks & AZO_Q2_BACKUP_NO_ALTERNATE)) sign = -sign;
      if (cur == root_idx) break;
      cur = node_path.back();
      node_path.pop_back();
    }
    stats.sims++;
  }

  // search (:191-217).  num_threads = 1: the deterministic mode every parity test pins.
  // num_threads = K > 1 (tree-parallel search with virtual loss, :196-214 + node.rs:77-92,359-365): the reference runs K
  // OS threads whose interleaving is up to the scheduler, so its result is not reproducible; here the K threads run in
  // ONE fixed interleaving, wave by wave — the K selections of a wave one after the other (each sees the virtual losses
  // and the locks of the earlier ones), then the wave's evaluations, then the K backups in thread order.  That is a
  // schedule the reference can produce, and it is what the device does (csrc/mcts.cuh, wave mode), bit for bit.
  size_t num_threads = 1;
  void search(size_t root_idx) {
    if (num_threads <= 1) {
      for (size_t sim_id = 0; sim_id < num_sims; ++sim_id) search_iteration(root_idx);
      return;
    }
    if (num_sims % num_threads != 0) throw std::runtime_error("num_sims % num_threads != 0 (async_mcts.rs:192)");
    for (size_t sim_id = 0; sim_id < num_sims; sim_id += num_threads) search_wave(root_idx, num_threads);
  }

  struct InFlight {
    size_t cur = 0;                 // where the walk stopped (resolved node index)
    std::vector<size_t> path;      // node_path
    float v = 0.0f;                 // value to back up (known at once, or after the wave's evaluations)
    int wait_for = -1;              // >= 0: the value is -(network value) of that in-flight simulation's leaf
    bool evaluate = false;          // this simulation owns a pending evaluation of `cur`
  };

  // One thread's walk of a wave: async_mcts.rs:236-356 with the repairs of App. C, visit() = N + 1 and VL + 1 on every node
  // of the path; placeholders that another in-flight simulation holds are Locked and skipped on the retry (:253-257,
  // node.rs:359-365).  Repairs that only matter with K > 1:
  //   F17: no selectable child (every child Locked; node.rs:367 unwrap panics) -> the walk ends at this node with v = 0;
  //   F18: a node whose evaluation is still pending in this wave is reached through a link or a duplicate (the reference
  //        reads its missing policy and panics, :254 -> node.rs:354) -> the walk ends there and shares that evaluation.
  void select_wave(size_t root_idx, std::vector<InFlight>& fl, size_t t) {
    InFlight& me = fl[t];
    size_t cur = root_idx, depth = 0;
    for (;;) {
      Node<G>* n = nodes->get(cur);
      const size_t rcur = *nodes->resolve(cur);
      stats.levels++;
      if (depth > max_depth) { n->visit(); me.v = n->mu.s->eval_heuristic(); break; }
      if (n->e != 0.0f) { n->visit(); stats.terminal_hits++; me.v = n->e; break; }
      if (!n->mu.p) {
        n->visit();
        int owner = -1;
        for (size_t k = 0; k < t; ++k)
          if (fl[k].evaluate && fl[k].cur == rcur) owner = static_cast<int>(k);
        if (owner >= 0) me.wait_for = owner;   // F18
        else me.evaluate = true;               // F1: a root that was never evaluated
        cur = rcur;
        break;
      }
      n->visit();
      size_t c = nodes->best_child(cur, cpuct, false);
      if (nodes->state(c) == NodeState::Locked) {              // :253-257: retry without the Locked children
        bool any = false;
        for (size_t ch : n->children) any = any || nodes->state(ch) != NodeState::Locked;
        if (!any) { me.v = 0.0f; cur = rcur; break; }          // F17
        c = nodes->best_child(cur, cpuct, true);
      }
      if (nodes->state(c) == NodeState::PlaceHolder) {
        if (!nodes->lock(c)) throw std::runtime_error("lock failed on an unlocked placeholder");
        me.path.push_back(cur);
        uint8_t a = nodes->raw(c)->a;
        auto nx = n->mu.s->get_next_state(1, a);
        G s2 = nx.first.get_canonical_form(nx.second);
        cur = c;
        auto up = nodes->upgrade(cur, s2);
        if (!up) throw std::runtime_error("Upgraded invalid node! (:291)");
        if (!*up) { stats.dup_links++; cur = *nodes->resolve(cur); continue; }
        stats.expansions++;
        Node<G>* m = nodes->get(cur);
        m->visit();
        if (m->e != 0.0f) { nodes->unlock(cur); stats.terminal_hits++; me.v = m->e; break; }
        me.evaluate = true;                                    // stays Locked until the wave's evaluation phase
        break;
      }
      me.path.push_back(cur);
      cur = c;
      depth += 1;
    }
    me.cur = *nodes->resolve(cur);
  }

  void search_wave(size_t root_idx, size_t K) {
    std::vector<InFlight> fl(K);
    for (size_t t = 0; t < K; ++t) select_wave(root_idx, fl, t);
    for (size_t t = 0; t < K; ++t)
      if (fl[t].evaluate) fl[t].v = -evaluate(fl[t].cur);      // set_policy + unlock inside
    for (size_t t = 0; t < K; ++t)
      if (fl[t].wait_for >= 0) fl[t].v = fl[static_cast<size_t>(fl[t].wait_for)].v;
    for (size_t t = 0; t < K; ++t) {
      size_t cur = fl[t].cur;
      float sign = 1.0f;
      const size_t root_res = *nodes->resolve(root_idx);
      for (;;) {
        nodes->get(cur)->unvisit(sign * fl[t].v, quirks);
        if (!(quirks & AZO_Q2_BACKUP_NO_ALTERNATE)) sign = -sign;
        if (*nodes->resolve(cur) == root_res) break;
        cur = fl[t].path.back();
        fl[t].path.pop_back();
      }
      stats.sims++;
    }
  }

  // F12: a state absent from the tree becomes a new root (push + upgrade).
  size_t root_for(const G& s) {
    auto r = nodes->lookup_state_id(s);                        // :81
    if (r) return *r;
    size_t idx = nodes->push(Node<G>(WIN_SCALE));
    nodes->upgrade(idx, s);
    return idx;
  }

  void root_counts(size_t root_idx, uint16_t* counts) const {
    const Node<G>* root = nodes->get(root_idx);
    for (size_t a = 0; a < G::num_actions(); ++a) counts[a] = 0;
    for (size_t child_idx : root->children) {                  // :88-94 with F7
      uint8_t a = nodes->raw(child_idx)->a;
      counts[a] = nodes->get(child_idx)->get_n();
    }
  }

  // get_action_prob (:74-115); F8 repaired; temp == 0 ties -> highest action (App. B.7).
  std::vector<float> get_action_prob(const G& s, float temp, uint16_t* counts_out = nullptr) {
    size_t root_idx = root_for(s);
    search(root_idx);
    const size_t A = G::num_actions();
    std::vector<uint16_t> counts(A);
    root_counts(root_idx, counts.data());
    if (counts_out)
      for (size_t a = 0; a < A; ++a) counts_out[a] = counts[a];
    std::vector<float> probs(A, 0.0f);
    if (temp == 0.0f) {                                        // :97-107
      size_t best_a = 0;
      for (size_t a = 0; a < A; ++a)
        if (counts[a] >= counts[best_a]) best_a = a;
      probs[best_a] = 1.0f;
    } else {                                                   // :108-113 (F8)
      std::vector<float> c(A);
      for (size_t a = 0; a < A; ++a)
        c[a] = temp == 1.0f ? static_cast<float>(counts[a])
                            : std::pow(static_cast<float>(counts[a]), 1.0f / temp);
      float sum = 0.0f;
      for (size_t a = 0; a < A; ++a) sum = sum + c[a];
      for (size_t a = 0; a < A; ++a) probs[a] = c[a] / sum;
    }
    return probs;
  }
};

}  // namespace azo

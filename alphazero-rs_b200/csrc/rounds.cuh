// Lock-step round engine: thousands of concurrent games as resumable per-slot state machines.
//
// Replaces the thread plumbing of the reference's self-play path: the rayon episode pool
// (src/coach.rs:202-205,241-272), the per-move scoped search threads (src/async_mcts.rs:191-217),
// the inference thread with its channels (src/async_mcts.rs:117-189) and the sequential arena
// loop (src/arena.rs:62-99).  A *slot* holds one game (self-play: one tree; arena: two trees, one
// per player) and its complete state in HBM (GameRec), so any warp can continue it:
//
//   k_compact : recycles finished slots for pending games, rebuilds the dense active list
//               (re-dealing the live games evenly over the SMs), resets the leaf batches;
//   k_round   : one warp per active slot runs its game forward until the slot
//                 - has a leaf that the batched network must evaluate (evaluator NNET): the
//                   position goes to the dense leaf batch of its model and the simulation is
//                   suspended (Pending + path saved), or
//                 - has played `plies_per_launch` plies (fused evaluators), or finished its game;
//   nnet forward (csrc/nnet.cuh) evaluates every pending leaf of a model in ONE dense pass
//               (repairs F15/F16: leaves grouped per model, batches of whatever size exists).
#pragma once
#include "kernels.cuh"

namespace azb {

enum : uint32_t { kPhaseEmpty = 0, kPhaseFresh, kPhaseNewMove, kPhaseSearch, kPhasePending, kPhaseDone };
enum : uint32_t { kModeSelfPlay = 0, kModeArena = 1 };

struct TreeVars {
  uint32_t n_blocks, n_owners, error, slow;
  uint32_t stat[8];
};

struct __align__(16) GameRec {
  uint64_t cur, opp;  // canonical board of the side to move (coach.rs:120 / arena.rs:25)
  uint32_t game;      // index of the game within the call
  uint32_t phase;
  int32_t player;     // +1 / -1 to move (coach.rs:114, arena.rs:14)
  uint32_t step;      // plies played so far + 1 while searching (coach.rs:116-119)
  uint32_t sims_done;
  uint32_t root_slot, root_meta;
  uint32_t leaf_idx;  // dense index of the pending leaf in its model's batch
  uint32_t pad[3];
  Pending pd;
  TreeVars tv[2];
  PathEnt path[kPathCap];  // the suspended simulation's path, whole entries: after the resume it is the next simulation's prediction
};

struct LeafBufs {
  uint4* state;     // [2][n_slots]  {cur.lo, cur.hi, opp.lo, opp.hi}: the feature planes as bitboards
  uint32_t* count;  // [2]           leaves per model this round
  float* pi;        // [2][n_slots][8]  network policy (probabilities), row = dense leaf index
  float* v;         // [2][n_slots]
  // Leaf de-duplication within a round (dmask != 0): games that share their move history grow identical trees and ask
  // for the same positions at the same time (all games of a call during the first move, 7 groups during the second,
  // ...; an arena without opening plies plays ONE game per seat order).  The network's answer for a position does not
  // depend on the batch it sits in, so a position is evaluated once per round and model: the first slot to claim the
  // position's key in a per-round hash table owns the dense batch row, the others remember the table entry and read the
  // owner's row when they resume.  Entries carry a 15-bit round stamp above the 49-bit state key, so the table is never
  // cleared between rounds (the host clears it when the stamp wraps).
  // Two tables, used by even and odd rounds in turn: the entries of round r are read when its duplicates resume in round
  // r + 1, WHILE other slots of round r + 1 already claim entries for their new leaves — in the other table.
  unsigned long long* dkeys;    // [2 parities][2 models][dmask + 1]  stamp << 49 | state_key
  uint32_t* didx;               // [2 parities][2 models][dmask + 1]  the owner's dense row
  uint32_t dmask;               // entries per model - 1 (power of two); 0 = no de-duplication
  uint32_t stamp;               // 1 .. 32766, changes every round (k_round derives it from the device round counter)
  uint32_t dpar;                // round parity: which table this round's claims go to
  unsigned long long* nn_total; // positions sent through the networks so far (k_compact adds the round's counts)
  // Evaluation cache for the duration of one call (cmask != 0): (state key -> raw policy[7], value) of every position a
  // network has evaluated, per model.  A simulation whose leaf is in the cache finishes at once instead of suspending
  // for a round; the slot that owned a batch row inserts the answer when it resumes.  Entry states: 0 empty,
  // key | kCacheBusy while the values are being written, key | kCacheReady afterwards (values read with ld.cg).
  unsigned long long* ckeys;    // [2][cmask + 1]
  float* cvals;                 // [2][cmask + 1][8]
  uint32_t cmask;
  unsigned long long* cache_hits;
};
constexpr unsigned long long kCacheBusy = 1ull << 63, kCacheReady = 1ull << 62, kCacheKeyMask = (1ull << 49) - 1ull;
constexpr int kCacheProbes = 8;
__device__ __forceinline__ uint32_t cache_hash(uint64_t skey, uint32_t mask) {
  return static_cast<uint32_t>((skey * 0xD6E8FEB86659FD93ull) >> 32) & mask;
}
// lane 0 probes; returns the entry index (warp-uniform) or 0xFFFFFFFF
__device__ __forceinline__ uint32_t cache_find(const LeafBufs& leaf, uint32_t side, uint64_t skey, int lane) {
  uint32_t found = 0xFFFFFFFFu;
  if (lane == 0) {
    const uint32_t base = side * (leaf.cmask + 1u);
    uint32_t h = cache_hash(skey, leaf.cmask);
    for (int i = 0; i < kCacheProbes; ++i, h = (h + 1u) & leaf.cmask) {
      const unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(leaf.ckeys + base + h);
      if (cur == 0ull) break;
      if ((cur & kCacheKeyMask) == skey) {
        if (cur & kCacheReady) found = base + h;
        break;
      }
    }
  }
  found = __shfl_sync(kFull, found, 0);
  if (found != 0xFFFFFFFFu) __threadfence();  // the values were written before the ready key
  return found;
}
// lanes 0-6 hold pi, `val` is warp-uniform
__device__ __forceinline__ void cache_insert(const LeafBufs& leaf, uint32_t side, uint64_t skey, float pi, float val, int lane) {
  uint32_t at = 0xFFFFFFFFu;
  if (lane == 0) {
    const uint32_t base = side * (leaf.cmask + 1u);
    uint32_t h = cache_hash(skey, leaf.cmask);
    for (int i = 0; i < kCacheProbes; ++i, h = (h + 1u) & leaf.cmask) {
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(leaf.ckeys + base + h);
      if (cur == 0ull) {
        cur = atomicCAS(leaf.ckeys + base + h, 0ull, skey | kCacheBusy);
        if (cur == 0ull) { at = base + h; break; }
      }
      if ((cur & kCacheKeyMask) == skey) break;  // already there, or another slot is writing it
    }
  }
  at = __shfl_sync(kFull, at, 0);
  if (at == 0xFFFFFFFFu) return;
  if (lane < 8) __stcg(leaf.cvals + static_cast<size_t>(at) * 8u + lane, lane < 7 ? pi : val);
  __threadfence();
  __syncwarp();
  if (lane == 0) *reinterpret_cast<volatile unsigned long long*>(leaf.ckeys + at) = skey | kCacheReady;
}
constexpr uint32_t kLeafIndirect = 0x80000000u;  // GameRec.leaf_idx: low bits index didx[] instead of the batch
constexpr uint32_t kLeafDone = 0x7FFFFFFFu;      // WaveSim.leaf_ref: answered from the cache, nothing to fetch

struct RoundParams {
  SearchParams p;
  uint32_t mode;
  int32_t ev_kind[2];  // evaluator of player A / B (self-play: [0])
  uint32_t plies_per_launch;
  uint32_t sims_per_launch;  // network rounds: a slot that needs no evaluation (endgame: terminal hits
                             // only) yields after this many simulations, so a round never waits on it
  uint32_t n_slots, n_games;
  uint32_t leaf_cap;  // rows of a model's leaf batch: n_slots * num_sim_threads (a wave suspends with up to K leaves)
  uint32_t half;    // arena: games [0, half) seat A first, [half, 2*half) seat B first (arena.rs:74-83)
  uint32_t k_open;  // arena: random opening plies
  uint32_t shared;  // arena: 1 = the reference's layout (coach.rs:333-354): ONE tree pair for the whole match, games
                    // strictly sequential (n_slots == 1); the trees and their counters survive from game to game
  uint64_t first_game_id;
};

constexpr uint32_t kStampPeriod = 32766u;  // even: a round pair (one graph launch) never straddles the stamp's wrap
struct Control {
  unsigned int* round;       // rounds started so far: k_compact counts, k_round derives the de-duplication stamp from it
  unsigned long long* host_progress;  // page-locked host word: (round << 32) | live slots, written by every k_round, so
                                      // that the host follows the run (and sees its end) without a blocking copy
  unsigned int* next_game;   // games handed out so far
  unsigned int* n_active;    // live slots after k_compact (this round's counter)
  unsigned int* n_active_next;  // the next round's counter: zeroed by this round's k_compact
  uint32_t* active_list;     // [n_slots]
  int8_t* arena_result;      // [n_games] play_game's return value (arena.rs:51)
};

// Diagnostic (AZB200_ROUND_TIMES=1): %globaltimer at four points of every round — before k_compact (which = 0), before
// k_round (1), after k_round (2), after the forward passes (3) — into times[round % cap][4].
__global__ void k_stamp(unsigned long long* times, const unsigned int* round, uint32_t which, uint32_t cap) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  const uint32_t r = which == 0u ? *round : *round - 1u;  // k_compact counts the round it starts
  times[static_cast<size_t>(r % cap) * 4u + which] = t;
}

// ---- k_compact: recycle finished slots, hand out pending games, rebuild the dense active list -----------------
// One thread per slot over as many CTAs as it takes (the single-CTA block scan this replaces took 13.8 us per round:
// 8 x 20 barriers over strided loads of the slot records).  The order of the active list carries no meaning (a slot's
// result depends only on its game id), so live slots append themselves with one warp-aggregated atomicAdd.  The counter
// is double-buffered: this launch fills ctl.n_active (zeroed by the previous launch) and zeroes ctl.n_active_next.
__global__ void __launch_bounds__(256) k_compact(RoundParams rp, GameRec* recs, Control ctl, LeafBufs leaf) {
  const uint32_t slot = blockIdx.x * 256u + threadIdx.x, lane = threadIdx.x & 31u;
  if (slot == 0u) {
    *leaf.nn_total += static_cast<unsigned long long>(leaf.count[0]) + leaf.count[1];
    leaf.count[0] = 0u;
    leaf.count[1] = 0u;
    *ctl.n_active_next = 0u;
    *ctl.round += 1u;  // (read by this round's k_round only)
  }
  uint32_t alive = 0;
  if (slot < rp.n_slots) {
    uint32_t ph = recs[slot].phase;
    if (ph == kPhaseEmpty || ph == kPhaseDone) {
      ph = kPhaseEmpty;
      if (*ctl.next_game < rp.n_games) {  // cheap pre-check, then claim
        const unsigned int g = atomicAdd(ctl.next_game, 1u);
        if (g < rp.n_games) {
          recs[slot].game = g;
          ph = kPhaseFresh;
        }
      }
      recs[slot].phase = ph;
    }
    alive = ph != kPhaseEmpty;
  }
  const uint32_t ball = __ballot_sync(kFull, alive != 0u);
  if (ball == 0u) return;
  const int leader = __ffs(static_cast<int>(ball)) - 1;
  uint32_t base = 0;
  if (static_cast<int>(lane) == leader) base = atomicAdd(ctl.n_active, static_cast<unsigned int>(__popc(ball)));
  base = __shfl_sync(kFull, base, leader);
  if (alive) ctl.active_list[base + __popc(ball & ((1u << lane) - 1u))] = slot;
}

// ---- per-ply bookkeeping shared with k_selfplay ------------------------------------------------
// get_action_prob's tail (async_mcts.rs:84-114) + execute_episode's move (coach.rs:130-155).
// Returns the game_ended code of the position after the move (0 = game goes on).
__device__ __forceinline__ uint32_t selfplay_move(const WarpTree& t, const SearchParams& p, GameBufs g,
                                                  uint32_t gi, uint64_t game_id, uint32_t step, BB& board,
                                                  int& player, uint32_t root_meta, int lane, uint32_t& err) {
  const float temp = step < p.temp_threshold ? 1.0f : 0.0f;  // :122-126
  const uint32_t cnt = root_child_count(t, root_meta, lane);
  const float pi = counts_to_pi(cnt, temp, lane);
  const uint32_t ply = step - 1u;
  const size_t grow = static_cast<size_t>(gi) * kTraceStride + ply;
  const size_t srow = static_cast<size_t>(gi) * kMaxPlies + ply;
  if (lane < 7) {
    g.counts[grow * 7u + lane] = static_cast<uint16_t>(cnt);
    g.sample_pi[srow * 8u + lane] = pi;
  }
  if (lane == 7) g.sample_pi[srow * 8u + 7u] = static_cast<float>(player);
  if (lane == 8)
    g.sample_state[srow] = make_uint4(static_cast<uint32_t>(board.cur), static_cast<uint32_t>(board.cur >> 32),
                                      static_cast<uint32_t>(board.opp), static_cast<uint32_t>(board.opp >> 32));
  const float u = philox_uniform01(p.seed, game_id, ply, 0u);
  const int a = choose_weighted(pi, u);  // :137-138
  if (a < 0) { err = kErrInternal; return 0u; }
  if (lane == 0) g.actions[grow] = static_cast<uint8_t>(a);
  board = play_canonical(board, a);  // :140-142
  player = -player;
  return static_cast<uint32_t>(game_ended_code(board, p.quirks));  // :144
}

__device__ __forceinline__ void load_tree_vars(WarpTree& t, const TreeVars& tv, int lane) {
  t.n_blocks = tv.n_blocks;
  t.n_owners = tv.n_owners;
  t.error = tv.error;
  t.slow = tv.slow;
  t.stat = lane < 8 ? tv.stat[lane] : 0u;
}
__device__ __forceinline__ void store_tree_vars(const WarpTree& t, TreeVars& tv, int lane) {
  if (lane == 0) {
    tv.n_blocks = t.n_blocks;
    tv.n_owners = t.n_owners;
    tv.error = t.error;
    tv.slow = t.slow;
  }
  if (lane < 8) tv.stat[lane] = t.stat;
}

// ---- one leaf's way through the batched evaluator ---------------------------------------------------------------
// leaf_submit: claim the position for this round (or find the slot that already has), take a dense batch row; returns
// the reference the slot keeps until it resumes (a row index, or kLeafIndirect | de-duplication entry).  Lane 0 works.
__device__ __forceinline__ uint32_t leaf_submit(const LeafBufs& leaf, uint32_t side, uint32_t leaf_cap, BB leaf_pos, int lane) {
  uint32_t ref = 0;
  if (lane == 0) {
    bool owner = true;
    uint32_t at = 0;
    if (leaf.dmask) {  // claim the position for this round, or find the slot that already has
      const uint64_t skey = state_key(leaf_pos);
      const unsigned long long key = skey | (static_cast<unsigned long long>(leaf.stamp) << 49);
      const uint32_t base = (leaf.dpar * 2u + static_cast<uint32_t>(side)) * (leaf.dmask + 1u);
      uint32_t h = static_cast<uint32_t>((skey * 0x9E3779B97F4A7C15ull) >> 40) & leaf.dmask;
      for (;;) {
        at = base + h;
        unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(leaf.dkeys + at);
        if ((cur >> 49) != leaf.stamp) {  // empty or left over from an earlier round
          const unsigned long long old = atomicCAS(leaf.dkeys + at, cur, key);
          if (old == cur) break;          // claimed
          cur = old;
          if ((cur >> 49) != leaf.stamp) continue;
        }
        if (cur == key) { owner = false; break; }
        h = (h + 1u) & leaf.dmask;  // (the table has 4 entries per leaf the round can hold: it never fills)
      }
    }
    if (owner) {
      ref = atomicAdd(leaf.count + side, 1u);
      leaf.state[static_cast<size_t>(side) * leaf_cap + ref] =
          make_uint4(static_cast<uint32_t>(leaf_pos.cur), static_cast<uint32_t>(leaf_pos.cur >> 32),
                     static_cast<uint32_t>(leaf_pos.opp), static_cast<uint32_t>(leaf_pos.opp >> 32));
      if (leaf.dmask) leaf.didx[at] = ref;  // read by the duplicates in the next round's kernel
    } else {
      ref = kLeafIndirect | at;
    }
  }
  return __shfl_sync(kFull, ref, 0);
}
// leaf_fetch: the network's answer for a submitted leaf (lane a: raw pi[a]); the owner of the row feeds the call's cache.
__device__ __forceinline__ void leaf_fetch(const LeafBufs& leaf, uint32_t side, uint32_t leaf_cap, uint32_t ref, uint64_t skey,
                                           int lane, float& pi, float& val) {
  uint32_t li = ref;
  if (li & kLeafIndirect) li = leaf.didx[li & ~kLeafIndirect];  // a duplicate: the row of the slot that owns the position
  const size_t row = static_cast<size_t>(side) * leaf_cap + li;
  pi = lane < 7 ? leaf.pi[row * 8u + lane] : 0.0f;
  val = leaf.v[row];
  if (leaf.cmask && !(ref & kLeafIndirect)) cache_insert(leaf, side, skey, pi, val, lane);
}

// ---- k_round ------------------------------------------------------------------------------------
template <bool WAVE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kCtasPerSm)
k_round(RoundParams rp, Pools pools, GameRec* recs, Control ctl, LeafBufs leaf, GameBufs g) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t round = *ctl.round;  // 1-based
  leaf.stamp = (round - 1u) % kStampPeriod + 1u;
  if (w == 0u && lane == 0 && ctl.host_progress) {
    *reinterpret_cast<volatile unsigned long long*>(ctl.host_progress) =
        (static_cast<unsigned long long>(round) << 32) | *ctl.n_active;
    __threadfence_system();
  }
  if (w >= *ctl.n_active) return;
  const uint32_t slot = ctl.active_list[w];
  GameRec* rec = recs + slot;
  const SearchParams& p = rp.p;
  const uint32_t tps = rp.mode == kModeArena ? 2u : 1u;

  uint32_t phase = rec->phase;
  const uint32_t gi = rec->game;
  BB board{rec->cur, rec->opp};
  int player = rec->player;
  uint32_t step = rec->step, sims_done = rec->sims_done;
  uint32_t root_slot = rec->root_slot, root_meta = rec->root_meta;
  uint32_t plies_left = rp.plies_per_launch;
  uint32_t sims_left = rp.sims_per_launch ? rp.sims_per_launch : 0xFFFFFFFFu;

  // which player's tree / evaluator is in use: self-play always 0; arena: the side to move
  // (games [0, half): A holds +1; games [half, ..): B holds +1 — arena.rs:74-83)
  auto side_of = [&](int pl) -> uint32_t {
    if (rp.mode != kModeArena) return 0u;
    const bool a_first = gi < rp.half;
    return (pl > 0) == a_first ? 0u : 1u;
  };
  uint32_t side = 0;
  WarpTree t = open_tree(pools, p, slot * tps);

  if (phase == kPhaseFresh) {  // AsyncMcts::default (coach.rs:246-255): fresh tree(s) on the initial board
    if (!(rp.mode == kModeArena && rp.shared && gi != 0u))
      for (uint32_t k = 0; k < tps; ++k) {
        WarpTree tk = open_tree(pools, p, slot * tps + k);
        fresh_table(tk, pools, p, slot * tps + k, lane);
        if (lane < 12) reinterpret_cast<uint32_t*>(&rec->tv[k])[lane] = 0u;
      }
    board = BB{0ull, 0ull};
    player = 1;
    step = 0;
    phase = kPhaseNewMove;
    __syncwarp();
  }
  side = side_of(player);
  t = open_tree(pools, p, slot * tps + side);
  load_tree_vars(t, rec->tv[side], lane);
  uint32_t err = 0;

  if (WAVE && phase == kPhasePending && rec->pd.kind == kPendWave) {  // the answers for a suspended wave's leaves are in
    WaveSim* ws = reinterpret_cast<WaveSim*>(t.path);
    {
      const uint4* src = reinterpret_cast<const uint4*>(rec->path);
      uint4* dst = reinterpret_cast<uint4*>(t.path);
      for (uint32_t i = lane; i < p.num_threads * (sizeof(WaveSim) / 16u); i += 32u) dst[i] = src[i];
    }
    __syncwarp();
    for (uint32_t k = 0; k < p.num_threads; ++k)
      if ((ws[k].flags & kWvEval) && ws[k].leaf_ref != kLeafDone) {
        float pi, val;
        leaf_fetch(leaf, side, rp.leaf_cap, ws[k].leaf_ref, ws[k].key, lane, pi, val);
        wave_set_policy(t, p, ws, k, pi, val, lane);
      }
    wave_backup(t, p, ws, p.num_threads, lane);
    sims_done += p.num_threads;
    phase = kPhaseSearch;
  } else if (phase == kPhasePending) {  // the network's answer for the suspended simulation is in
    Pending pd = rec->pd;
    for (uint32_t i = lane; i <= pd.plen && i < kPathCap; i += 32u) t.path[i] = rec->path[i];
    __syncwarp();
    float pi, val;
    leaf_fetch(leaf, side, rp.leaf_cap, rec->leaf_idx, pd.key, lane, pi, val);
    if (pd.kind == kPendRoot) finish_root_eval(t, p, root_slot, root_meta, pi, val, lane);
    else finish_expand(t, p, pd, pi, val, lane, /*normalised=*/false, /*predict=*/true);
    sims_done++;
    phase = kPhaseSearch;
  }

  for (;;) {
    if (t.error) { err = t.error; break; }
    if (phase == kPhaseNewMove) {
      if (rp.mode == kModeArena) {
        // arena.rs:18 — the loop condition is checked before every move
        const uint32_t code = static_cast<uint32_t>(game_ended_code(board, p.quirks));
        if (code) {
          // arena.rs:51: cur_player * round(get_game_ended(cur_player)); the draw value rounds to 0
          const int r = code == 1u ? 1 : (code == 2u ? -1 : 0);
          if (lane == 0) {
            ctl.arena_result[gi] = static_cast<int8_t>(player * r);
            g.plies[gi] = step;
          }
          phase = kPhaseDone;
          break;
        }
        if (step < rp.k_open) {  // random opening ply (not in the reference; k_open = 0 for parity)
          const uint32_t vm = valid_mask(board.cur | board.opp);
          const float wgt = (lane < 7 && ((vm >> lane) & 1u)) ? 1.0f : 0.0f;
          const int a = choose_weighted(wgt, philox_uniform01(p.seed, rp.first_game_id + gi, step, 1u));
          if (lane == 0) g.actions[static_cast<size_t>(gi) * kTraceStride + step] = static_cast<uint8_t>(a);
          board = play_canonical(board, a);
          player = -player;
          step++;
          continue;
        }
        const uint32_t ns = side_of(player);
        if (ns != side) {  // the other player's tree takes over
          store_tree_vars(t, rec->tv[side], lane);
          side = ns;
          t = open_tree(pools, p, slot * tps + side);
          load_tree_vars(t, rec->tv[side], lane);
        }
      }
      step++;                                             // coach.rs:119
      if (!make_root(t, p, board, lane, root_slot, root_meta)) { err = t.error; break; }  // :81 (+F12)
      // visit counts near the 16-bit wrap (quirk Q6) take the generic walk: this tree's simulations so far + this search
      if (__shfl_sync(kFull, t.stat, kStatSims) + p.num_sims >= kSafeVisits) t.slow = 1u;
      sims_done = 0;
      phase = kPhaseSearch;
    }
    // ---- search (async_mcts.rs:191-217) ----
    const int ev = rp.ev_kind[side];
    bool suspended = false, yielded = false;
    Pending pd;
    BB leaf_pos;
    if constexpr (WAVE) {
      // ---- tree-parallel mode: waves of K walks (mcts.cuh wave_*); a wave suspends with up to K leaves ----
      WaveSim* ws = reinterpret_cast<WaveSim*>(t.path);
      bool waiting = false;
      while (sims_done < p.num_sims && !t.error) {
        if (sims_left == 0u) { yielded = true; break; }
        sims_left = sims_left > p.num_threads ? sims_left - p.num_threads : 0u;
        for (uint32_t k = 0; k < p.num_threads && !t.error; ++k) wave_select(t, p, ev, board, root_slot, root_meta, ws, k, lane);
        if (t.error) break;
        uint32_t n_wait = 0;
        for (uint32_t k = 0; k < p.num_threads; ++k) {
          if (!(ws[k].flags & kWvEval)) continue;
          const bool root_eval = (ws[k].flags & kWvRoot) != 0u;
          const BB lp{ws[k].cur, ws[k].opp};
          if (ev < AZB_EVAL_NNET) {
            if (root_eval) {  // F1 with an inline evaluator (new nodes were evaluated in the walk)
              float pi, val;
              evaluate_inline(ev, lp, lane, pi, val);
              wave_set_policy(t, p, ws, k, pi, val, lane);
            }
            continue;
          }
          uint32_t ce = 0xFFFFFFFFu;
          if (leaf.cmask) ce = cache_find(leaf, side, ws[k].key, lane);
          if (ce != 0xFFFFFFFFu) {  // a position this model has already evaluated in this call
            const float pi = lane < 7 ? __ldcg(leaf.cvals + static_cast<size_t>(ce) * 8u + lane) : 0.0f;
            const float val = __ldcg(leaf.cvals + static_cast<size_t>(ce) * 8u + 7u);
            wave_set_policy(t, p, ws, k, pi, val, lane);
            if (lane == 0) {
              atomicAdd(leaf.cache_hits, 1ull);
              ws[k].leaf_ref = kLeafDone;
            }
          } else {
            const uint32_t ref = leaf_submit(leaf, side, rp.leaf_cap, lp, lane);
            if (lane == 0) ws[k].leaf_ref = ref;
            n_wait++;
          }
          __syncwarp();
        }
        if (n_wait) { waiting = true; break; }
        wave_backup(t, p, ws, p.num_threads, lane);
        sims_done += p.num_threads;
      }
      if (t.error) { err = t.error; break; }
      if (yielded) break;
      if (waiting) {  // the wave's state (its K walks) waits in the slot record for the network's answers
        const uint4* src = reinterpret_cast<const uint4*>(t.path);
        uint4* dst = reinterpret_cast<uint4*>(rec->path);
        for (uint32_t i = lane; i < p.num_threads * (sizeof(WaveSim) / 16u); i += 32u) dst[i] = src[i];
        if (lane == 0) rec->pd.kind = kPendWave;
        phase = kPhasePending;
        break;
      }
      goto make_move;
    }
  search_more:
    suspended = false;
    while (sims_done < p.num_sims && !t.error) {
      if (sims_left == 0u) { yielded = true; break; }
      sims_left--;
      if (sims_done == 0 && root_needs_eval(t, root_meta)) {  // repair F1
        if (ev >= AZB_EVAL_NNET) {
          pd.kind = kPendRoot;
          pd.plen = 0;
          pd.key = state_key(board);
          leaf_pos = board;
          suspended = true;
          break;
        }
        float pi, val;
        evaluate_inline(ev, board, lane, pi, val);
        finish_root_eval(t, p, root_slot, root_meta, pi, val, lane);
      } else if (!one_sim(t, p, ev, board, root_slot, root_meta, lane, pd, leaf_pos)) {
        suspended = true;
        break;
      }
      sims_done++;
    }
    if (t.error) { err = t.error; break; }
    if (yielded) break;  // phase stays Search; the slot continues next round
    if (suspended && leaf.cmask) {  // a position this model has already evaluated in this call: no round trip
      const uint32_t ce = cache_find(leaf, side, pd.key, lane);
      if (ce != 0xFFFFFFFFu) {
        const float pi = lane < 7 ? __ldcg(leaf.cvals + static_cast<size_t>(ce) * 8u + lane) : 0.0f;
        const float val = __ldcg(leaf.cvals + static_cast<size_t>(ce) * 8u + 7u);
        if (pd.kind == kPendRoot) finish_root_eval(t, p, root_slot, root_meta, pi, val, lane);
        else finish_expand(t, p, pd, pi, val, lane, /*normalised=*/false, /*predict=*/true);
        if (lane == 0) atomicAdd(leaf.cache_hits, 1ull);
        sims_done++;
        goto search_more;
      }
    }
    if (suspended) {  // hand the leaf to the batched evaluator of this side's model
      const uint32_t ref = leaf_submit(leaf, side, rp.leaf_cap, leaf_pos, lane);
      if (lane == 0) {
        rec->pd = pd;
        rec->leaf_idx = ref;
      }
      __syncwarp();
      for (uint32_t i = lane; i <= pd.plen && i < kPathCap; i += 32u) rec->path[i] = t.path[i];
      phase = kPhasePending;
      break;
    }
    // ---- the move ----
  make_move:
    if (rp.mode == kModeArena) {
      // coach.rs:356-371: argmax of get_action_prob(s, temp = 0) — a one-hot on the most visited
      // child, ties to the highest action
      const uint32_t cnt = root_child_count(t, root_meta, lane);
      const uint32_t mx = __reduce_max_sync(kFull, lane < 7 ? cnt : 0u);
      const int a = 31 - __clz(__ballot_sync(kFull, lane < 7 && cnt == mx));
      const uint32_t vm = valid_mask(board.cur | board.opp);
      if (!((vm >> a) & 1u)) { err = kErrInternal; break; }  // arena.rs:29-35 assert
      const size_t grow = static_cast<size_t>(gi) * kTraceStride + (step - 1u);
      if (lane < 7) g.counts[grow * 7u + lane] = static_cast<uint16_t>(cnt);
      if (lane == 0) g.actions[grow] = static_cast<uint8_t>(a);
      board = play_canonical(board, a);
      player = -player;
      phase = kPhaseNewMove;
    } else {
      const uint32_t code = selfplay_move(t, p, g, gi, rp.first_game_id + gi, step, board, player, root_meta, lane, err);
      if (err) break;
      if (code || step >= static_cast<uint32_t>(kMaxPlies)) {
        if (lane == 0) {
          g.plies[gi] = step;
          g.final_r[gi] = game_ended_value(static_cast<int>(code));
          g.final_player[gi] = static_cast<int8_t>(player);
        }
        if (!code) err = kErrInternal;
        phase = kPhaseDone;
        break;
      }
      phase = kPhaseNewMove;
    }
    if (rp.plies_per_launch && --plies_left == 0u) break;
  }

  if (err) {
    phase = kPhaseDone;
    t.error = err;
  }
  store_tree_vars(t, rec->tv[side], lane);
  __syncwarp();
  if (phase == kPhaseDone) {  // per-game statistics (both trees for arena)
    if (lane == 0) g.error[gi] = rec->tv[0].error | (tps > 1u ? rec->tv[1].error : 0u);
    if (lane < 8) {
      uint32_t s = rec->tv[0].stat[lane] + (tps > 1u ? rec->tv[1].stat[lane] : 0u);
      if (lane == 6) s = max(rec->tv[0].n_blocks, tps > 1u ? rec->tv[1].n_blocks : 0u);
      if (lane == 7) s = max(rec->tv[0].n_owners, tps > 1u ? rec->tv[1].n_owners : 0u);
      g.stats[gi * 8u + lane] = s;
    }
  }
  if (lane == 0) {
    rec->cur = board.cur;
    rec->opp = board.opp;
    rec->phase = phase;
    rec->player = player;
    rec->step = step;
    rec->sims_done = sims_done;
    rec->root_slot = root_slot;
    rec->root_meta = root_meta;
  }
}

}  // namespace azb

// Global kernels of the self-play path: batched connect-four (K1), the tree-search hook
// (K2-K4), the whole-game persistent self-play kernel for fused evaluators (K5/K7) and the
// sample export (coach.rs:130-155).  Host-side launch code lives in engine.cu.
#pragma once
#include "mcts.cuh"

namespace azb {

// Warps (= trees) per CTA of the search kernels.  A CTA's slot on its SM is free again only when ALL its warps are done, so
// with back-to-back batches on two streams (azb_coach_self_play_begin / _end) small CTAs hand finished games' warp slots
// to the next batch sooner.
// Measured with the steps three deep (20 batches of 4096 x 800, three runs each): 4 warps per CTA 70.9-71.1 ms per batch,
// 2 warps 69.7-69.8; the network rounds (config 3, config 4) are the same with either.
#ifndef AZB_WARPS_PER_CTA
#define AZB_WARPS_PER_CTA 2
#endif
constexpr int kWarpsPerCta = AZB_WARPS_PER_CTA;
// Resident warps per SM: 28 at 72 registers, 32 at 64, 36 at 56.  The persistent self-play kernel runs at 32: with
// consecutive batches overlapped the device is full, and four more warps per SM hide more latency than the 8 registers
// cost (measured, 8 batches three deep: 74.7 ms per batch at 32, 78.5 at 28, 83.4 at 36; one batch alone 108 vs 106 ms).
// The round kernel keeps 28: config 3 measured 3.46-3.54 s at 28, 32, 36 and 40 alike (AZB_ROUND_WARPS_PER_SM; it spills
// 192 bytes at 72 registers, 336 at 64) — a round is bounded by cold tree loads, not by resident warps.
#ifndef AZB_WARPS_PER_SM
#define AZB_WARPS_PER_SM 32
#endif
#ifndef AZB_ROUND_WARPS_PER_SM
#define AZB_ROUND_WARPS_PER_SM 28
#endif
constexpr int kCtasPerSm = AZB_ROUND_WARPS_PER_SM / kWarpsPerCta;  // k_round, k_mcts_search
constexpr int kPlayCtasPerSm = AZB_WARPS_PER_SM / kWarpsPerCta;  // k_selfplay
constexpr int kMaxPlies = 42;   // the board has 42 cells
constexpr int kTraceStride = 64;

struct Pools {
  uint4* blocks;  // [n_trees][cap_blocks*8]
  uint4* tables;  // [n_trees][(bucket_mask+1)*8]
  TreeRec* recs;  // [n_trees]
};

__device__ __forceinline__ WarpTree open_tree(const Pools& pools, const SearchParams& p, uint32_t tree) {
  __shared__ PathEnt s_path[kWarpsPerCta][kPathCap];
  __shared__ uint4 s_win[kWarpsPerCta][8];
  WarpTree t;
  const int wi = (threadIdx.x >> 5) % kWarpsPerCta;
  if ((threadIdx.x & 31) < 8) s_win[wi][threadIdx.x & 31] = win_table(p.quirks, threadIdx.x & 31);
  __syncwarp();
  t.win = s_win[wi];
  t.blocks = pools.blocks + static_cast<size_t>(tree) * p.cap_blocks * 8u;
  t.table = pools.tables + static_cast<size_t>(tree) * (static_cast<size_t>(p.bucket_mask) + 1u) * 8u;
  t.path = s_path[wi];
  t.tt_stamp = static_cast<uint64_t>(pools.recs[tree].tt_gen + 1u) << 49;
  t.n_blocks = t.n_owners = t.error = t.slow = 0u;
  t.pred_len = t.hold = 0u;
  t.leaf_slot = t.leaf_meta = 0u;
  t.stat = 0u;
  t.uni_prior = uniform_prior_table(threadIdx.x & 31);
  return t;
}

// A fresh tree for a new game (AsyncMcts::default, coach.rs:246-255): the transposition table's generation moves on, so
// every entry of the previous game reads as empty; the table is zero-filled only when the 15-bit generation wraps.
__device__ __forceinline__ void fresh_table(WarpTree& t, const Pools& pools, const SearchParams& p, uint32_t tree, int lane) {
  uint32_t gen = pools.recs[tree].tt_gen + 1u;
  __syncwarp();
  if (gen >= kTtGenWrap) {
    const uint32_t n = (p.bucket_mask + 1u) * 8u;
    for (uint32_t i = lane; i < n; i += 32u) t.table[i] = make_uint4(0u, 0u, 0u, 0u);
    gen = 0u;
  }
  if (lane == 0) pools.recs[tree].tt_gen = gen;
  t.tt_stamp = static_cast<uint64_t>(gen + 1u) << 49;
  __syncwarp();
}

// ---- AsyncMcts::get_action_prob for n_trees persistent trees (test hook) --------------------
template <bool WAVE>
__global__ void __launch_bounds__(kWarpsPerCta * 32)
k_mcts_search(int ev_kind, SearchParams p, Pools pools, const BB* __restrict__ states, float temp,
              uint16_t* __restrict__ counts, float* __restrict__ pi_out, uint32_t n_trees) {
  const uint32_t tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tree >= n_trees) return;
  WarpTree t = open_tree(pools, p, tree);
  TreeRec* rec = pools.recs + tree;
  t.n_blocks = rec->n_blocks;
  t.n_owners = rec->n_owners;
  t.error = rec->error;
  if (t.error) return;
  t.slow = rec->slow | (rec->stat[kStatSims] + p.num_sims >= kSafeVisits ? 1u : 0u);
  const BB s = states[tree];
  uint32_t rs = 0, rm = 0;
  if (make_root(t, p, s, lane, rs, rm)) {  // lookup_state_id (:81), F12 on a miss
    run_sims<WAVE>(t, p, ev_kind, s, rs, rm, p.num_sims, lane);
    if (!t.error) {
      const uint32_t cnt = root_child_count(t, rm, lane);
      const float pi = counts_to_pi(cnt, temp, lane);
      if (lane < 7) {
        counts[tree * 7u + lane] = static_cast<uint16_t>(cnt);
        pi_out[tree * 7u + lane] = pi;
      }
    }
  }
  if (lane == 0) {
    rec->n_blocks = t.n_blocks;
    rec->n_owners = t.n_owners;
    rec->error = t.error;
    rec->slow = t.slow;
  }
  if (lane < kNumStats) rec->stat[lane] += t.stat;
}

// Raw counter of the node owning states[tree] (0 when absent).
__global__ void k_counter_of(SearchParams p, Pools pools, const BB* __restrict__ states,
                             uint64_t* __restrict__ out, uint32_t n_trees) {
  const uint32_t tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tree >= n_trees) return;
  WarpTree t = open_tree(pools, p, tree);
  uint32_t slot, meta, ins;
  uint64_t c = 0;
  if (tt_find(t, p.bucket_mask, state_key(states[tree]), lane, slot, meta, ins)) c = ld_counter(t, slot);
  if (lane == 0) out[tree] = c;
}

// One row per unique state of one tree (NodeStore.seen).
__global__ void k_dump_tree(SearchParams p, Pools pools, uint32_t tree, uint64_t cap,
                            uint64_t* keys, uint64_t* counters, float* e, float* p7, uint8_t* has_p,
                            unsigned long long* n_rows) {
  WarpTree t = open_tree(pools, p, tree);
  const uint32_t n = (p.bucket_mask + 1u) * 8u;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 en = t.table[i];
    const uint64_t k = (static_cast<uint64_t>(en.y) << 32) | en.x;
    if (((k ^ t.tt_stamp) >> 49) != 0ull) continue;  // empty, or an earlier game's entry
    const unsigned long long row = atomicAdd(n_rows, 1ull);
    if (row >= cap) continue;
    keys[row] = k & kKeyMask;
    counters[row] = ld_counter(t, en.z);
    const uint32_t meta = en.w;
    e[row] = meta_is_terminal(meta) ? terminal_e(meta & 3u) : -0.0f;  // -get_game_ended == -0.0
    uint8_t hp = 0;
    for (int a = 0; a < 7; ++a) p7[row * 7 + a] = 0.0f;
    if (meta_is_block(meta)) {
      const uint4* bp = t.blocks + static_cast<size_t>(meta) * 8u;
      hp = (block_flags(t, meta) & kFlagHasPolicy) ? 1 : 0;
      if (hp)
        for (int a = 0; a < 7; ++a) p7[row * 7 + a] = __uint_as_float(bp[a].z);
    }
    has_p[row] = hp;
  }
}

// ---- whole games on device: Coach::execute_episode (coach.rs:104-157) ------------------------
struct GameBufs {
  uint32_t* plies;        // [n_games]
  float* final_r;         // [n_games]
  int8_t* final_player;   // [n_games]
  uint32_t* error;        // [n_games]
  uint8_t* actions;       // [n_games][64], 0xFF padded
  uint16_t* counts;       // [n_games][64][7]
  uint4* sample_state;    // [n_games][42] {cur.lo, cur.hi, opp.lo, opp.hi}
  float* sample_pi;       // [n_games][42][8]  pi[0..6], player in [7]
  uint32_t* stats;        // [n_games][8]  6 search stats, blocks used, owners
  unsigned long long* ply_ns;  // [n_games][64] %globaltimer at the end of each ply (diagnostic), or nullptr
};

template <bool WAVE>
__global__ void __launch_bounds__(kWarpsPerCta * 32, kPlayCtasPerSm)
k_selfplay(int ev_kind, SearchParams p, Pools pools, GameBufs g, uint32_t n_trees, uint32_t n_games,
           uint64_t first_game_id, unsigned int* next_game) {
  const uint32_t tree = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (tree >= n_trees) return;
  WarpTree t = open_tree(pools, p, tree);
  for (;;) {
    uint32_t gi = 0;
    if (lane == 0) gi = atomicAdd(next_game, 1u);
    gi = __shfl_sync(kFull, gi, 0);
    if (gi >= n_games) break;
    // AsyncMcts::default (coach.rs:246-255): a fresh tree per episode
    fresh_table(t, pools, p, tree, lane);
    t.n_blocks = t.n_owners = t.error = t.slow = 0u;
    t.stat = 0u;
    BB board{0ull, 0ull};  // canonical board of the side to move (coach.rs:120)
    int player = 1;        // :114
    uint32_t step = 0;     // :116
    uint32_t code = 0;
    for (;;) {
      step++;                                                    // :119
      const float temp = step < p.temp_threshold ? 1.0f : 0.0f;  // :122-126
      uint32_t rs = 0, rm = 0;
      if (!make_root(t, p, board, lane, rs, rm)) break;          // get_action_prob :81 (+F12)
      if (step * p.num_sims >= kSafeVisits) t.slow = 1u;
      run_sims<WAVE>(t, p, ev_kind, board, rs, rm, p.num_sims, lane);  // :82
      if (t.error) break;
      const uint32_t cnt = root_child_count(t, rm, lane);
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
      const float u = philox_uniform01(p.seed, first_game_id + gi, ply, 0u);
      const int a = choose_weighted(pi, u);                      // :137-138
      if (a < 0) { t.error = kErrInternal; break; }
      if (lane == 0) g.actions[grow] = static_cast<uint8_t>(a);
      if (g.ply_ns && lane == 0) {
        unsigned long long tns;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
        g.ply_ns[grow] = tns;
      }
      board = play_canonical(board, a);                          // :140-142
      player = -player;
      code = static_cast<uint32_t>(game_ended_code(board, p.quirks));  // :144
      if (code || step >= static_cast<uint32_t>(kMaxPlies)) break;
    }
    if (lane == 0) {
      g.plies[gi] = step;
      g.final_r[gi] = game_ended_value(static_cast<int>(code));
      g.final_player[gi] = static_cast<int8_t>(player);
      g.error[gi] = t.error ? t.error : (code ? 0u : kErrInternal);
      g.stats[gi * 8u + 6u] = t.n_blocks;
      g.stats[gi * 8u + 7u] = t.n_owners;
    }
    if (lane < kNumStats) g.stats[gi * 8u + lane] = t.stat;
  }
}

// SOATrainingSamples (src/nnet.rs:33) from the per-ply records: 2 samples per ply (identity,
// mirror — coach.rs:130-135, connect_four_game.rs:205-211), label per coach.rs:146-153.
// One CTA per game; offsets[] = exclusive prefix sum of plies.
__global__ void k_export_samples(GameBufs g, const uint64_t* __restrict__ offsets, uint32_t quirks,
                                 float* __restrict__ boards, float* __restrict__ pis,
                                 float* __restrict__ vs, uint64_t capacity) {
  const uint32_t gi = blockIdx.x;
  const uint32_t plies = g.plies[gi];
  const uint64_t base = offsets[gi] * 2ull;
  const float r = g.final_r[gi];
  const float fp = static_cast<float>(g.final_player[gi]);
  const uint32_t n_elem = plies * 2u * 92u;  // 84 features + 7 pi + 1 v
  for (uint32_t i = threadIdx.x; i < n_elem; i += blockDim.x) {
    const uint32_t smp = i / 92u, j = i % 92u;
    const uint32_t ply = smp >> 1, sym = smp & 1u;
    const uint64_t row = base + smp;
    if (row >= capacity) continue;
    const size_t srow = static_cast<size_t>(gi) * kMaxPlies + ply;
    if (j < 84u) {
      const uint4 st = g.sample_state[srow];
      uint64_t b = j < 42u ? ((static_cast<uint64_t>(st.y) << 32) | st.x)
                           : ((static_cast<uint64_t>(st.w) << 32) | st.z);
      uint32_t cell = j % 42u;
      if (sym) cell = (cell / 7u) * 7u + (6u - cell % 7u);
      boards[row * 84ull + j] = ((b >> cell) & 1ull) ? 1.0f : 0.0f;
    } else if (j < 91u) {
      const uint32_t a = j - 84u;
      pis[row * 7ull + a] = g.sample_pi[srow * 8u + (sym ? 6u - a : a)];
    } else {
      const float pl = g.sample_pi[srow * 8u + 7u];
      float v;
      if (quirks & AZB_Q4_VLABEL_LITERAL) v = (pl == fp) ? 1.0f : -1.0f;
      else v = (pl == fp) ? r : -r;
      vs[row] = v;
    }
  }
}

// ---- arithmetic self-test: the slow-path-free division / sqrt used by the level loop against
// the IEEE intrinsics, exhaustively over the integer domain they are used on ---------------------
__global__ void k_selftest_arith(unsigned long long* mismatches) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nthreads = gridDim.x * blockDim.x;
  unsigned long long bad_rcp = 0, bad_sqrt = 0, bad_div = 0, bad_q = 0;
  if (tid == 0 && __uint_as_float(0x3E124925u) != __fdiv_rn(1.0f, 7.0f)) bad_div++;  // evaluate_inline's RN(1/7)
  for (uint32_t b = 1u + tid; b <= 65536u; b += nthreads) {
    const float fb = static_cast<float>(b);
    if (rcp_int(fb) != __frcp_rn(fb)) bad_rcp++;
    const float x = __fadd_rn(static_cast<float>(b - 1u), kEps);
    if (sqrt_count(x) != __fsqrt_rn(x)) bad_sqrt++;
    uint32_t s = b * 2654435761u;
    for (int k = 0; k < 512; ++k) {
      s = s * 1664525u + 1013904223u;
      // t3-like numerators: all mantissas, exponents 2^-60 .. 2^60, both signs
      float a = __uint_as_float((s & 0x807FFFFFu) | ((67u + (s >> 23) % 120u) << 23));
      if (k == 0) a = 0.0f;
      if (fdiv_by_int(a, fb) != __fdiv_rn(a, fb)) bad_div++;
      // counters at rest: N = b (mod 2^16), W_raw spread over +-2^25 (dense) and the full range
      const uint32_t wr = (k & 1) ? (0x7FFFFFFFu + (s >> 6) - (1u << 25)) : (s ^ (s << 7));
      const uint64_t c = counter_pack(wr, b);
      if (__float_as_uint(counter_q_fast(c)) != __float_as_uint(counter_q(c))) bad_q++;
    }
  }
  if (bad_q) atomicAdd(mismatches + 3, bad_q);
  if (bad_rcp) atomicAdd(mismatches + 0, bad_rcp);
  if (bad_sqrt) atomicAdd(mismatches + 1, bad_sqrt);
  if (bad_div) atomicAdd(mismatches + 2, bad_div);
}

// ---- batched connect-four over 43-byte states (trait Game, src/game.rs:10-28) -----------------
__device__ __forceinline__ void load_cells(const int8_t* p, uint64_t& pos, uint64_t& neg) {
  pos = neg = 0ull;
  for (int i = 0; i < 42; ++i) {
    const int8_t c = p[i];
    if (c > 0) pos |= 1ull << i;
    else if (c < 0) neg |= 1ull << i;
  }
}
__device__ __forceinline__ void store_cells(int8_t* p, uint64_t pos, uint64_t neg, int8_t me) {
  for (int i = 0; i < 42; ++i) p[i] = ((pos >> i) & 1ull) ? 1 : (((neg >> i) & 1ull) ? -1 : 0);
  p[42] = me;
}

enum : int { kOpNext = 0, kOpValid, kOpEnded, kOpCanonical, kOpSymmetries, kOpFeatures };

__global__ void k_c4_batch(int op, const int8_t* __restrict__ in, const int8_t* __restrict__ player,
                           const uint8_t* __restrict__ action, const float* __restrict__ pi_in,
                           uint32_t quirks, size_t n, int8_t* __restrict__ out_states,
                           int8_t* __restrict__ out_i8, uint8_t* __restrict__ out_u8,
                           float* __restrict__ out_f32, int* __restrict__ bad) {
  const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t pos, neg;
  load_cells(in + 43 * i, pos, neg);
  const int8_t me = in[43 * i + 42];
  switch (op) {
    case kOpNext: {  // get_next_state, connect_four_game.rs:90-102
      const int a = action[i];
      const int8_t pl = player[i];
      if (a > 6 || !((valid_mask(pos | neg) >> a) & 1u) || (pl != 1 && pl != -1)) {
        atomicExch(bad, 1);  // the reference underflows `DEFAULT_HEIGHT - *height` and panics
        store_cells(out_states + 43 * i, pos, neg, me);
        out_i8[i] = 0;
        return;
      }
      const uint64_t b = landing_bit(pos | neg, a);
      if (pl > 0) pos |= b; else neg |= b;
      store_cells(out_states + 43 * i, pos, neg, me);
      out_i8[i] = static_cast<int8_t>(-pl);
      break;
    }
    case kOpValid: {  // get_valid_moves, :104-109
      const uint32_t vm = valid_mask(pos | neg);
      for (int a = 0; a < 7; ++a) out_u8[7 * i + a] = (vm >> a) & 1u;
      break;
    }
    case kOpEnded: {  // get_game_ended(player), :111-196
      const int code = game_ended_code(BB{pos, neg}, quirks);
      float r = 0.0f;
      if (code == 1) r = player[i] == 1 ? 1.0f : -1.0f;
      else if (code == 2) r = player[i] == -1 ? 1.0f : -1.0f;
      else if (code == 3) r = 1e-4f;
      out_f32[i] = r;
      break;
    }
    case kOpCanonical: {  // get_canonical_form(player), :198-203 repaired (F10)
      if (player[i] < 0) store_cells(out_states + 43 * i, neg, pos, 1);
      else store_cells(out_states + 43 * i, pos, neg, 1);
      break;
    }
    case kOpSymmetries: {  // get_symmetries, :205-211 (flip() starts from empty(): me = +1)
      store_cells(out_states + 43 * (2 * i), pos, neg, me);
      store_cells(out_states + 43 * (2 * i + 1), mirror(pos), mirror(neg), 1);
      for (int a = 0; a < 7; ++a) {
        out_f32[7 * (2 * i) + a] = pi_in[7 * i + a];
        out_f32[7 * (2 * i + 1) + a] = pi_in[7 * i + 6 - a];
      }
      break;
    }
    case kOpFeatures: {  // to_features, :219-237 repaired (F11): [2,6,7]
      const uint64_t mine = me > 0 ? pos : (me < 0 ? neg : 0ull);
      const uint64_t theirs = me > 0 ? neg : (me < 0 ? pos : 0ull);
      for (int c = 0; c < 42; ++c) {
        out_f32[84 * i + c] = ((mine >> c) & 1ull) ? 1.0f : 0.0f;
        out_f32[84 * i + 42 + c] = ((theirs >> c) & 1ull) ? 1.0f : 0.0f;
      }
      break;
    }
  }
}

}  // namespace azb

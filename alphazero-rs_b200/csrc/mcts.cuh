// Warp-per-tree Monte-Carlo tree search (kernel families K2-K4, K5 for the fused evaluators).
//
// Replaces src/async_mcts.rs (search_iteration :219-371, get_action_prob :74-115) and the
// NodeStore operations it calls (src/node.rs:179-370) in the reference's deterministic
// mode (num_sim_threads = 1), with the repair list F1-F8, F12 of SURVEY.md App. A/C.
//
// One warp owns one tree and runs one simulation at a time, so every f32 rounding and every
// tie-break happens in the reference's order and results are bit-exact with the oracle.
//
// The search is bound by the dependent instruction chain of one simulation (ncu: profiles/
// r1_selfplay_ncu.md: a warp alone on its scheduler issues one instruction per ~5.4 cycles), so the
// walk is built to execute as few instructions as possible:
//   * speculative prefix (one_sim_impl): the previous simulation's path is re-checked four levels per
//     pass, one level per 8-lane group, with one shuffle + one ballot instead of four arg-maxes;
//   * per level a group reads ONE 128-byte block (lane 8g + a = edge a: {w, q, prior, meta} + n[a]),
//     resolves transposition links under a warp vote, computes
//     u = q + (cpuct*P*sqrt(N_parent))/(1+n) per lane (Q is cached, one division left, done
//     with an FMA sequence proven correctly rounded); a real arg-max (LAST maximum: redux.max.f32 +
//     ballot + bfind) is taken only at the first level that leaves the prediction and below it;
//   * visit()/unvisit() of the reference are folded into one read-modify-write per path node
//     at backup (lane l handles path entry l), which also refreshes the cached q, N and sqrt(N);
//   * the path lives in shared memory with the position of every level, so an expansion has its
//     position at hand; the win test is lane-parallel; the UNIFORM evaluator's priors are a table.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "c4_bitboard.cuh"
#include "tree.cuh"

namespace azb {

constexpr unsigned kFull = 0xFFFFFFFFu;
constexpr int kPathCap = 48;  // <= 42 moves to a full board + dup-link hops never push

enum : uint32_t { kErrNone = 0, kErrBlocks = 1, kErrTable = 2, kErrInternal = 3 };
enum : int { kStatSims = 0, kStatLevels, kStatExpansions, kStatTerminal, kStatDupLinks, kStatEvals, kNumStats };

struct SearchParams {
  uint32_t cap_blocks;   // blocks per tree (< 2^24)
  uint32_t bucket_mask;  // transposition table: (bucket_mask+1) buckets of 8 entries
  uint32_t num_sims;
  uint32_t max_depth;
  float cpuct_f;
  uint32_t quirks;
  uint32_t temp_threshold;
  uint32_t num_threads;  // num_sim_threads (coach.rs:51): 1 = deterministic mode; K > 1 = waves of K simulations (wave_*)
  uint64_t seed;
};

// Persistent per-tree record (HBM).  Lives across get_action_prob calls / plies.
struct TreeRec {
  uint32_t n_blocks;
  uint32_t n_owners;  // == NodeStore.seen.len()
  uint32_t error;
  uint32_t slow;      // sticky: see WarpTree::slow
  uint32_t stat[8];
  uint32_t tt_gen;    // generation of the tree's transposition table: entries carry tt_gen + 1 above their 49-bit key, so a
                      // new game (AsyncMcts::default) bumps it instead of zero-filling the table (1 MB per tree at config 2)
  uint32_t pad[3];
};
constexpr uint32_t kTtGenWrap = 32766u;  // 15 bits above the key; tt_gen + 1 stays in 1 .. 32767
constexpr uint64_t kKeyMask = (1ull << 49) - 1ull;

// One level of the current simulation's path (shared memory).  After the simulation's backup the
// entries double as the PREDICTION for the next simulation of the same tree: consecutive
// simulations share most of their path (measured: 5.5 of 7.6 levels), so the next simulation
// evaluates up to four predicted levels at once, one per 8-lane group (one_sim_impl).
struct __align__(16) PathEnt {
  uint32_t sa;   // (slot << 3) | action: the node of this level (its owner slot) and the edge taken (7 = none yet)
  uint32_t blk;  // the node's child block
  uint32_t n;    // the node's N after the last backup that passed through it
  uint32_t sq;   // f32 bits of sqrt(n + 1 + 1e-6): best_child's parent term for the next visit
  // hold bound (see backup_path): written when the level's arg-max is taken, s0 = the parent term it was taken with
  float m0;      // >= u of every edge but the chosen one at s0, rounding slack included (-inf: no other edge)
  float e0s;     // >= (largest non-negative exploration term of those edges at s0) / s0
  float s0;
  float cp;      // RN(cpuct * P[action]) of the chosen edge
  uint64_t cur, opp;  // the node's position, canonical for its side to move
};

// Per-warp view of one tree.  Every member is warp-uniform except `stat` (lane k = stat k).
struct WarpTree {
  uint4* blocks;   // slot id indexes this directly (16-byte slots, 8 per block)
  uint4* table;    // HashEntry as uint4 {key.lo, key.hi, slot, meta}; key = tt_stamp | 49-bit state key
  uint64_t tt_stamp;  // (tt_gen + 1) << 49: entries of another generation (or zero-filled ones) read as empty
  PathEnt* path;   // shared memory, kPathCap entries
  const uint4* win;  // shared memory, 8 entries: win_table()
  uint32_t pred_len;  // levels [0, pred_len) of `path` describe nodes the previous simulation walked
  uint32_t hold;      // bit l: level l (< pred_len, < 31) provably selects its recorded edge again (backup_path)
  uint32_t leaf_slot, leaf_meta;  // the finished game the previous simulation ended in (when it did)
  uint32_t n_blocks, n_owners, error;
  uint32_t slow;  // != 0: use __fdiv_rn in the level loop (a prior outside the range the FMA division
                  // is proven for, or visit counts that may wrap the 16-bit N field, quirk Q6)
  uint32_t stat;
  float uni_prior;  // lane k: the UNIFORM evaluator's normalised prior with k legal actions (uniform_prior_table)
};

// Visit counts stay below this => (1 + n) & 0xFFFF is never 0 in the level loop.
constexpr uint32_t kSafeVisits = 65000u;
// cpuct * prior * sqrt(N) (sqrt(N) in [1, 256]) must stay a normal, finite float for fdiv_by_int.
__device__ __forceinline__ bool prior_needs_slow_div(float cpuct_f, float prior) {
  const float at = fabsf(__fmul_rn(cpuct_f, prior));
  return !(at == 0.0f || (at >= 1e-28f && at < 1e27f));
}

// ---- slot field access ------------------------------------------------------------------
__device__ __forceinline__ uint16_t* n_ptr(const WarpTree& t, uint32_t slot) {
  return reinterpret_cast<uint16_t*>(t.blocks + (slot | 7u)) + (slot & 7u);
}
__device__ __forceinline__ uint32_t ld_n(const WarpTree& t, uint32_t slot) { return *n_ptr(t, slot); }
__device__ __forceinline__ uint32_t ld_w(const WarpTree& t, uint32_t slot) {
  return *reinterpret_cast<const uint32_t*>(t.blocks + slot);
}
__device__ __forceinline__ float ld_q(const WarpTree& t, uint32_t slot) {
  return reinterpret_cast<const float*>(t.blocks + slot)[1];
}
__device__ __forceinline__ uint64_t ld_counter(const WarpTree& t, uint32_t slot) {
  return counter_pack(ld_w(t, slot), ld_n(t, slot));
}

// ---- exact f32 helpers without slow-path calls ---------------------------------------------
// a / b for b an integer-valued float in [1, 65536]: r = RN(1/b) (MUFU.RCP + one Newton step,
// checked exhaustively against __frcp_rn by azb_selftest_arith), then Markstein's correction
// twice: with an exact reciprocal and a faithful quotient the FMA residual step returns the
// correctly rounded quotient (no overflow/underflow: callers route tiny/huge `a` to __fdiv_rn).
__device__ __forceinline__ float rcp_int(float b) {
  float r0;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(b));
  const float e = __fmaf_rn(-b, r0, 1.0f);
  return __fmaf_rn(r0, e, r0);
}
__device__ __forceinline__ float fdiv_by_int(float a, float b) {
  const float r = rcp_int(b);
  const float q0 = __fmul_rn(a, r);
  const float e0 = __fmaf_rn(-b, q0, a);
  const float q1 = __fmaf_rn(e0, r, q0);
  const float e1 = __fmaf_rn(-b, q1, a);
  return __fmaf_rn(e1, r, q1);
}
// sqrt(x) for x = N + 1e-6, N in [0, 65535] (checked exhaustively against __fsqrt_rn).
__device__ __forceinline__ float sqrt_count(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  float s = __fmul_rn(x, y);
  const float h = __fmul_rn(0.5f, y);
  const float d = __fmaf_rn(-s, s, x);
  return __fmaf_rn(d, h, s);
}
// compute_q (node.rs:51-58) of a counter at rest (VL = 0) without slow-path calls: W/100 with the
// constant reciprocal RN(1/100) and two Markstein corrections, then /N by fdiv_by_int.  Checked
// against counter_q() by azb_selftest_arith.
__device__ __forceinline__ float counter_q_fast(uint64_t c) {
  const uint32_t n = counter_n(c);
  const uint32_t hi = static_cast<uint32_t>(c >> 32);
  if (n == 0u) return 0.0f;
  if (hi > 0xFFFFFF00u || counter_vl(c) != 0u) return counter_q(c);
  const float wf = static_cast<float>(static_cast<int>(hi - 0x7FFFFFFFu));
  const float r = 0.01f;  // RN(1/100)
  const float q0 = __fmul_rn(wf, r);
  const float e0 = __fmaf_rn(-kWinScale, q0, wf);
  const float q1 = __fmaf_rn(e0, r, q0);
  const float e1 = __fmaf_rn(-kWinScale, q1, wf);
  const float w = __fmaf_rn(e1, r, q1);
  return fdiv_by_int(w, static_cast<float>(n));
}
// index of the highest set bit (0xFFFFFFFF for 0): one FLO instead of 31 - clz
__device__ __forceinline__ uint32_t bfind_u32(uint32_t x) {
  uint32_t r;
  asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
  return r;
}
__device__ __forceinline__ float redux_max_f32(float v) {
  float m;
  asm volatile("redux.sync.max.f32 %0, %1, 0xffffffff;" : "=f"(m) : "f"(v));
  return m;
}

// ---- transposition table (NodeStore.seen): buckets of 8 entries = one 128-byte line ------
// Returns true and (slot, meta) when `key` is present; otherwise `ins` is the first free
// entry on the probe path (0xFFFFFFFF when the table is full).  `e` is the caller's (early) load of
// entry (lane & 7) of the key's home bucket, so the miss latency overlaps the caller's other work.
__device__ __forceinline__ uint4 tt_load_home(const WarpTree& t, uint32_t bucket_mask, uint64_t key, int lane) {
  return t.table[hash_bucket(key, bucket_mask) * 8u + (lane & 7)];
}
__device__ __forceinline__ bool tt_find(const WarpTree& t, uint32_t bucket_mask, uint64_t key, uint4 e,
                                        int lane, uint32_t& slot, uint32_t& meta, uint32_t& ins) {
  uint32_t b = hash_bucket(key, bucket_mask);
  for (uint32_t probe = 0; probe <= bucket_mask; ++probe) {
    const uint64_t k = (static_cast<uint64_t>(e.y) << 32) | e.x;
    const uint32_t hit = __ballot_sync(kFull, k == (key | t.tt_stamp)) & 0xFFu;
    if (hit) {
      const int l = __ffs(hit) - 1;
      slot = __shfl_sync(kFull, e.z, l);
      meta = __shfl_sync(kFull, e.w, l);
      return true;
    }
    const uint32_t emp = __ballot_sync(kFull, ((k ^ t.tt_stamp) >> 49) != 0ull) & 0xFFu;
    if (emp) {
      ins = b * 8u + (__ffs(emp) - 1);
      return false;
    }
    b = (b + 1u) & bucket_mask;
    e = t.table[b * 8u + (lane & 7)];
  }
  ins = 0xFFFFFFFFu;
  return false;
}
__device__ __forceinline__ bool tt_find(const WarpTree& t, uint32_t bucket_mask, uint64_t key,
                                        int lane, uint32_t& slot, uint32_t& meta, uint32_t& ins) {
  return tt_find(t, bucket_mask, key, tt_load_home(t, bucket_mask, key, lane), lane, slot, meta, ins);
}
__device__ __forceinline__ void tt_insert(const WarpTree& t, uint32_t ins, uint64_t key,
                                          uint32_t slot, uint32_t meta, int lane) {
  if (lane == 0)
    t.table[ins] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>((key | t.tt_stamp) >> 32), slot, meta);
}

// ---- leaf evaluators fused into the search (NNet::predict, src/nnet.rs:40-44) --------------
// Lane a (< 7) returns pi[a]; `v` is warp-uniform.  `kind` is warp-uniform.
__device__ __forceinline__ void evaluate_inline(int kind, BB s, int lane, float& pi, float& v) {
  if (kind == AZB_EVAL_UNIFORM) {  // examples/connect_four.rs:34-38
    pi = __uint_as_float(0x3E124925u);  // RN(1/7), checked by azb_selftest_arith
    v = 1.0f;
  } else {  // SURVEY App. B.6 hash evaluator
    uint64_t h = splitmix64(splitmix64(s.cur) + s.opp);
    pi = static_cast<float>(1u + static_cast<uint32_t>((h >> (8 * (lane & 7))) & 0xFFu));
    v = __fsub_rn(__fdiv_rn(static_cast<float>((h >> 56) & 0xFFu), 128.0f), 1.0f);
  }
}

// Mask invalid actions and renormalise (async_mcts.rs:319-345; sequential f32 sums, App. B.4).
__device__ __forceinline__ float mask_normalise(float pi, uint32_t vm, int lane) {
  const bool valid = lane < 7 && ((vm >> lane) & 1u);
  pi = valid ? pi : 0.0f;
  float s = 0.0f;
#pragma unroll
  for (int a = 0; a < 7; ++a) s = __fadd_rn(s, __shfl_sync(kFull, pi, a));
  if (s > 0.0f) return __fdiv_rn(pi, s);
  pi = __fadd_rn(pi, valid ? 1.0f : 0.0f);
  float s2 = 0.0f;
#pragma unroll
  for (int a = 0; a < 7; ++a) s2 = __fadd_rn(s2, __shfl_sync(kFull, pi, a));
  return __fdiv_rn(pi, s2);
}

// evaluate_inline + mask_normalise.  For the UNIFORM evaluator the result depends only on how many
// actions are legal (the sequential sum adds k copies of RN(1/7) and exact zeros), so it is read
// from the per-warp table built by uniform_prior_table() with the very same operations.
__device__ __forceinline__ float uniform_prior_table(int lane) {
  const float c = __fdiv_rn(1.0f, 7.0f);
  float s = 0.0f;
#pragma unroll
  for (int a = 0; a < 7; ++a) s = __fadd_rn(s, a < lane ? c : 0.0f);  // lane k: k legal actions
  float r = lane >= 1 && lane <= 7 ? __fdiv_rn(c, s) : 0.0f;
  asm volatile("" : "+f"(r));  // opaque: keep it in its register instead of recomputing it at every use
  return r;
}
// every table entry 1/k (k = 1..7) times cpuct stays in the range the FMA division is proven for
__device__ __forceinline__ bool uniform_prior_is_safe(float cpuct_f) {
  return !prior_needs_slow_div(cpuct_f, 1.0f) && !prior_needs_slow_div(cpuct_f, __fdiv_rn(1.0f, 7.0f));
}
__device__ __forceinline__ void evaluate_masked(const WarpTree& t, int kind, BB s, uint32_t vm, int lane,
                                                float& pi, float& v) {
  if (kind == AZB_EVAL_UNIFORM) {
    const float pr = __shfl_sync(kFull, t.uni_prior, __popc(vm & 0x7Fu));
    pi = (lane < 7 && ((vm >> lane) & 1u)) ? pr : 0.0f;
    v = 1.0f;
  } else {
    evaluate_inline(kind, s, lane, pi, v);
    pi = mask_normalise(pi, vm, lane);
  }
}

// game_ended_code() with the eight (direction, side) line tests spread over lanes: lane k & 7 tests
// direction k >> 1 (scan order H, V, D1, D2) for side k & 1 (0 = cur).  The per-lane constants
// {window starts, shift} sit in shared memory (t.win, filled by win_table()).  Warp-uniform result.
__device__ __forceinline__ uint4 win_table(uint32_t quirks, int k) {
  const bool lit = quirks & AZB_Q1_WIN_RANGE_LITERAL;
  const uint32_t dir = (static_cast<uint32_t>(k) >> 1) & 3u;
  const uint32_t sh = dir == 0u ? 1u : (dir == 1u ? 7u : (dir == 2u ? 8u : 6u));
  const uint64_t start = dir == 0u ? (lit ? kStartH_lit : kStartH_fix)
                                   : (dir == 1u ? (lit ? kStartV_lit : kStartV_fix) : (dir == 2u ? kStartD1 : kStartD2));
  return make_uint4(static_cast<uint32_t>(start), static_cast<uint32_t>(start >> 32), sh, 0u);
}
__device__ __forceinline__ int game_ended_code_warp(const WarpTree& t, BB s, int lane) {
  const uint4 k = t.win[lane & 7];
  const uint64_t b = (lane & 1) ? s.opp : s.cur;
  const uint64_t ls = line_starts(b, static_cast<int>(k.z), (static_cast<uint64_t>(k.y) << 32) | k.x);
  const uint32_t bal = __ballot_sync(kFull, ls != 0ull) & 0xFFu;
  if (bal == 0u) return ((s.cur | s.opp) == kBoard42) ? 3 : 0;
  const int d2 = (__ffs(static_cast<int>(bal)) - 1) & ~1;  // the first direction in scan order with a line
  const uint64_t c = __shfl_sync(kFull, ls, d2), o = __shfl_sync(kFull, ls, d2 + 1);
  const uint64_t m = c | o;
  return (c & (m & (0 - m))) ? 1 : 2;
}

// A fresh child block: every legal action is a placeholder with W = bias, N = 0, q = 0.
__device__ __forceinline__ void write_child_block(const WarpTree& t, uint32_t blk, uint32_t vm,
                                                  float prior, uint32_t flags, int lane) {
  uint4 out;
  if (lane < 7)
    out = make_uint4(kWBias, 0u, __float_as_uint(prior), ((vm >> lane) & 1u) ? kMetaPlaceholder : kMetaInvalid);
  else
    out = make_uint4(0u, 0u, 0u, flags << 16);  // n[0..6] = 0, flags in the last u16
  if (lane < 8) t.blocks[blk * 8u + lane] = out;
}
__device__ __forceinline__ uint32_t block_flags(const WarpTree& t, uint32_t blk) {
  return reinterpret_cast<const uint32_t*>(t.blocks + blk * 8u + 7u)[3] >> 16;
}

// push + upgrade of a state that is not in the tree, as a stand-alone root
// (NodeStore::new / from_root, node.rs:156-177; repair F12).  The root's own slot is slot 0
// of a holder block.  The node is NOT evaluated here (repair F1 does it at first visit).
__device__ __forceinline__ bool make_root(WarpTree& t, const SearchParams& p, BB s, int lane,
                                          uint32_t& root_slot, uint32_t& root_meta) {
  const uint64_t key = state_key(s);
  uint32_t o_slot, o_meta, ins;
  t.pred_len = 0u;  // a new root: the previous path predicts nothing
  if (tt_find(t, p.bucket_mask, key, lane, o_slot, o_meta, ins)) {
    root_slot = o_slot;
    root_meta = o_meta;
    return true;
  }
  if (ins == 0xFFFFFFFFu) { t.error = kErrTable; return false; }
  if (t.n_blocks + 2u > p.cap_blocks) { t.error = kErrBlocks; return false; }
  const uint32_t holder = t.n_blocks++;
  root_slot = holder * 8u;
  const int code = game_ended_code(s, p.quirks);
  if (code) {
    root_meta = kMetaTerminal | static_cast<uint32_t>(code);
  } else {
    root_meta = t.n_blocks++;
    write_child_block(t, root_meta, valid_mask(s.cur | s.opp), 0.0f, 0u, lane);
  }
  uint4 out = make_uint4(kWBias, 0u, 0u, lane == 0 ? root_meta : kMetaInvalid);
  if (lane == 7) out = make_uint4(0u, 0u, 0u, kFlagRootHolder << 16);
  if (lane < 8) t.blocks[holder * 8u + lane] = out;
  tt_insert(t, ins, key, root_slot, root_meta, lane);
  t.n_owners++;
  __syncwarp();
  return true;
}

// unvisit of one path node, folded with the visit that preceded it (node.rs:77-92), and refresh
// of the cached q — split into the loads + arithmetic (prepare) and the stores (commit), so that an
// expansion can do the first half while its transposition probe is still in flight.
struct BackupRegs {
  uint32_t w_new, q_bits, nword, n_new;
};
__device__ __forceinline__ uint32_t* n_word_ptr(const WarpTree& t, uint32_t slot) {
  return reinterpret_cast<uint32_t*>(t.blocks + (slot | 7u)) + ((slot & 7u) >> 1);
}
__device__ __forceinline__ BackupRegs backup_prepare(const WarpTree& t, uint32_t slot, float v, uint32_t quirks) {
  const uint32_t sh = (slot & 1u) * 16u;
  const uint32_t old = *n_word_ptr(t, slot);
  uint64_t c = counter_pack(ld_w(t, slot), old >> sh) + kVisit;
  c = counter_unvisit(c, v, quirks);
  BackupRegs r;
  r.w_new = static_cast<uint32_t>(c >> 32);
  r.q_bits = __float_as_uint(counter_q_fast(c));
  r.n_new = counter_n(c);
  r.nword = (old & ~(0xFFFFu << sh)) | (r.n_new << sh);
  return r;
}
__device__ __forceinline__ void backup_commit(const WarpTree& t, uint32_t slot, const BackupRegs& r) {
  *reinterpret_cast<uint2*>(t.blocks + slot) = make_uint2(r.w_new, r.q_bits);
  // n[] is updated with a 32-bit read-modify-write: sub-word global stores knock the whole line
  // out of L1 (measured: profiles/r1_v2_selfplay_ncu.md), and the next simulation re-reads it.
  // No other lane touches this block's header during a backup (path nodes sit in distinct blocks).
  *n_word_ptr(t, slot) = r.nword;
}

// ---- search_iteration (async_mcts.rs:219-371) --------------------------------
// A simulation either completes, or — when the leaf must be evaluated by the batched network
// (evaluator kind AZB_EVAL_NNET) — suspends after storing what is needed to finish it later.
enum : uint32_t { kPendNone = 0, kPendRoot = 1, kPendExpand = 2, kPendWave = 3 };
struct Pending {
  uint32_t kind;      // kPend*
  uint32_t my_slot;   // expansion: the placeholder being upgraded
  uint32_t new_meta;  // expansion: its freshly allocated child block
  uint32_t vm;        // valid-move mask of the evaluated position
  uint32_t plen;      // node_path length
  uint32_t ins;       // free transposition-table entry for the new state
  uint32_t levels;    // levels walked (statistic)
  uint32_t pad;
  uint64_t key;       // state key of the new node
};

// backup (:361-370): the leaf gets +v, then up node_path; Q2 corrected alternates the sign.
// Entry l < plen is path node l, entry plen is the leaf; the sign flips with distance.
//
// HOLD: the backup also decides, for every level l < plen at once (lane l), whether the NEXT simulation's best_child
// at that node is certain to pick the recorded edge again.  Between two simulations only the chosen edge's statistics
// change (q, n of the child: lane l + 1 has just computed them) and the parent term s = sqrt(N + 1e-6) grows; every
// other edge j keeps q_j, n_j (a persisting level is on every path in between, its other children are on none), so
//   u_j(s1) <= u_j(s0) + ex_j(s0) * (s1/s0 - 1) + rounding  <=  m0 + e0s * (s1 - s0)
// with the two maxima m0 / e0s recorded when the level's arg-max was last taken (score_level).  If the chosen edge's
// exact new u (same operations as best_child) is strictly above that bound, it is the unique maximum: the level holds
// and the next walk skips it without reading its block.  A level that does not hold is simply evaluated in full, which
// refreshes the bound; results are bit-identical with the one-level-at-a-time walk.
template <bool HOLD>
__device__ __forceinline__ void backup_path(WarpTree& t, const SearchParams& p, uint32_t plen,
                                            uint32_t leaf_slot, float v, uint32_t levels, int lane) {
  const bool alternate = !(p.quirks & AZB_Q2_BACKUP_NO_ALTERNATE);
  __syncwarp();
  uint32_t hold = 0u;
  for (uint32_t base = 0; base <= plen; base += 32u) {
    const uint32_t l = base + lane;
    uint32_t n_new = 0u, q_bits = 0u;
    float sq = 0.0f;
    if (l <= plen) {
      const uint32_t slot = l < plen ? (t.path[l].sa >> 3) : leaf_slot;
      const bool neg = alternate && ((plen - l) & 1u);
      const BackupRegs r = backup_prepare(t, slot, __fmul_rn(neg ? -1.0f : 1.0f, v), p.quirks);
      backup_commit(t, slot, r);
      n_new = r.n_new;
      q_bits = r.q_bits;
      sq = sqrt_count(__fadd_rn(static_cast<float>((n_new + 1u) & 0xFFFFu), kEps));
      *reinterpret_cast<uint2*>(&t.path[l].n) = make_uint2(n_new, __float_as_uint(sq));
    }
    if (HOLD && base == 0u) {
      const uint32_t cn = __shfl_down_sync(kFull, n_new, 1);  // the chosen edge's child is level l + 1
      const float cq = __uint_as_float(__shfl_down_sync(kFull, q_bits, 1));
      bool h = false;
      if (l < plen && lane < 31) {
        const float4 b = *reinterpret_cast<const float4*>(&t.path[l].m0);  // {m0, e0s, s0, cp}
        const float t3 = __fmul_rn(b.w, sq);
        const float t4 = static_cast<float>((1u + cn) & 0xFFFFu);
        const float up = __fadd_rn(cq, fdiv_by_int(t3, t4));
        const float bound = __fmaf_rn(b.y, __fsub_rn(sq, b.z), b.x);  // (s1 - s0 is exact for s0 <= s1 <= 2 s0)
        h = up > bound && sq <= __fmul_rn(2.0f, b.z);
      }
      hold = __ballot_sync(kFull, h);
    }
  }
  t.hold = hold;
  t.stat += static_cast<uint32_t>(lane == kStatSims) + (lane == kStatLevels ? levels : 0u);
  __syncwarp();
}

// Repair F1, second half: the root's raw policy (lane a = pi[a]) and value are known.
__device__ __forceinline__ void finish_root_eval(WarpTree& t, const SearchParams& p, uint32_t root_slot,
                                                 uint32_t root_meta, float pi, float val, int lane) {
  uint4* bp = t.blocks + static_cast<size_t>(root_meta) * 8u;
  const uint32_t m = lane < 7 ? bp[lane].w : kMetaInvalid;
  const uint32_t vm = __ballot_sync(kFull, m != kMetaInvalid) & 0x7Fu;
  pi = mask_normalise(pi, vm, lane);
  if (__any_sync(kFull, lane < 7 && prior_needs_slow_div(p.cpuct_f, pi))) t.slow = 1u;
  if (lane < 7) reinterpret_cast<float*>(bp + lane)[2] = pi;  // set_policy
  if (lane == 7) reinterpret_cast<uint32_t*>(bp + 7)[3] |= kFlagHasPolicy << 16;
  t.stat += static_cast<uint32_t>(lane == kStatEvals);
  backup_path<false>(t, p, 0u, root_slot, -val, 1u, lane);
}

// upgrade -> Some(true), second half (node.rs:290-322, async_mcts.rs:317-353): mask + normalise
// the policy, publish the new node, back the value up.
__device__ __forceinline__ void finish_expand(WarpTree& t, const SearchParams& p, const Pending& pd,
                                              float pi, float val, int lane, bool normalised = true,
                                              bool predict = false, bool safe_prior = false) {
  if (!normalised) pi = mask_normalise(pi, pd.vm, lane);
  // predict: the path of this simulation, with the new node as its last level, is the prediction
  // for the next one (only when path[0..plen) carries block ids, i.e. not after a resume)
  t.pred_len = 0u;
  if (predict) {
    if (lane == 0) *reinterpret_cast<uint2*>(t.path + pd.plen) = make_uint2((pd.my_slot << 3) | 7u, pd.new_meta);
    t.pred_len = pd.plen + 1u;
  }
  // (safe_prior: the UNIFORM evaluator's priors are 1/k, k = 1..7: inside fdiv_by_int's proven range whenever
  // cpuct itself is, which uniform_prior_is_safe() checks once)
  if (!safe_prior && __any_sync(kFull, lane < 7 && prior_needs_slow_div(p.cpuct_f, pi))) t.slow = 1u;
  write_child_block(t, pd.new_meta, pd.vm, pi, kFlagHasPolicy, lane);
  if (lane == 0) reinterpret_cast<uint32_t*>(t.blocks + pd.my_slot)[3] = pd.new_meta;
  tt_insert(t, pd.ins, pd.key, pd.my_slot, pd.new_meta, lane);
  t.n_owners++;
  t.stat += static_cast<uint32_t>(lane == kStatEvals || lane == kStatExpansions);
  if (predict) backup_path<true>(t, p, pd.plen, pd.my_slot, -val, pd.levels, lane);  // :353 returns -v
  else backup_path<false>(t, p, pd.plen, pd.my_slot, -val, pd.levels, lane);
}

// True when the node needs the F1 root evaluation before it can be searched.
__device__ __forceinline__ bool root_needs_eval(const WarpTree& t, uint32_t root_meta) {
  return meta_is_block(root_meta) && !(block_flags(t, root_meta) & kFlagHasPolicy);
}

// best_child's per-edge work (node.rs:343-370) for the node whose child block is `blk` and whose N
// before this simulation's visit is `npar`: lane (8g + a) returns edge a's slot words `w`, visit
// count `nn` and u = q + (cpuct*P*sq)/(1 + n) (-inf when the lane takes no part), sq = sqrt(N_parent +
// 1e-6) with the parent's N read after this simulation's visit().
__device__ __forceinline__ void load_edges(const WarpTree& t, uint32_t blk, uint32_t la, uint4& w, uint32_t& nn) {
  const uint4* bp = t.blocks + static_cast<size_t>(blk) * 8u;
  w = bp[la];
  nn = reinterpret_cast<const uint16_t*>(bp + 7)[la];
}
template <bool GENERIC>
__device__ __forceinline__ void score_edges(const WarpTree& t, float cpuct_f, float sq, bool use, uint32_t la,
                                            const uint4& w, uint32_t& nn, float& u, float& ex, float& cp, bool& ok) {
  ok = use && la < 7u && w.w != kMetaInvalid;
  float q = __uint_as_float(w.y);
  if (__any_sync(kFull, ok && w.w == kMetaLink)) {  // rare; kept off the common instruction stream
    if (ok && w.w == kMetaLink) {  // resolve(): statistics come from the owner (node.rs:179-201)
      q = ld_q(t, w.x);
      nn = ld_n(t, w.x);
    }
  }
  cp = __fmul_rn(cpuct_f, __uint_as_float(w.z));
  const float t3 = __fmul_rn(cp, sq);
  const float t4 = static_cast<float>((1u + nn) & 0xFFFFu);  // u16 arithmetic (quirk Q6)
  ex = GENERIC ? __fdiv_rn(t3, t4) : fdiv_by_int(t3, t4);
  u = ok ? __fadd_rn(q, ex) : __uint_as_float(0xFF800000u);
}

// One simulation from an evaluated (or terminal) root.  Returns false when it suspended for a
// network evaluation: then `pd` describes the pending expansion and `leaf` is the position to
// evaluate.  ev_kind < AZB_EVAL_NNET evaluates inline and never suspends.
// GENERIC = false is the hot variant: no max_depth check (depth counts moves into existing nodes, at
// most 42 on this board, so the check is dead unless max_depth < 43), the slow-path-free division,
// and HELD LEVELS: consecutive simulations of a tree share most of their path, and the previous
// simulation's backup has already decided (backup_path, t.hold) which levels of that path are certain
// to select their recorded edge again.  What a node selects depends only on that node's own block,
// never on how the walk got there, so the walk starts at the first level that does not hold (its node,
// N, sqrt term and position sit in its path entry), takes that level's real arg-max, and, while the
// arg-max confirms the recorded edge, jumps over the following held levels in the same way; the first
// level that picks another edge ends the prediction and the walk goes on one level per iteration.
// Held levels cost nothing here: no block read, no arg-max, no board arithmetic (their visit counts
// move at backup).  Results are bit-identical with the one-level-at-a-time walk (the GENERIC variant,
// and the oracle).
// GENERIC = true keeps the depth check, uses __fdiv_rn and never skips (trees whose `slow` flag is set).
template <bool GENERIC>
__device__ __forceinline__ bool one_sim_impl(WarpTree& t, const SearchParams& p, int ev_kind, BB root,
                                             uint32_t root_slot, uint32_t root_meta, int lane,
                                             Pending& pd, BB& leaf) {
  const float neg_inf = __uint_as_float(0xFF800000u);
  const uint32_t la = lane & 7u;
  const bool grp0 = lane < 8;
  const uint32_t pred_len = t.pred_len;  // (GENERIC: never read)
  const uint32_t nhold = ~t.hold;        // bit l clear: level l is held
  t.pred_len = 0u;  // set again by the exits that leave a usable path behind
  uint32_t cur_slot = root_slot, cur_meta = root_meta;
  uint32_t par_n = 0u;  // N of the current node before this simulation's visit
  uint32_t depth = 0, plen = 0;
  uint32_t empty_seen = 0u;  // != 0: some walked level had no selectable child
  uint32_t end_level = 1;    // levels walked = plen + end_level
  float v = 0.0f;
  // a node whose game has ended: value e (:246-249 + F6); the max_depth exit comes first (:241-244)
  auto at_terminal = [&](uint32_t meta) {
    if (GENERIC && depth > p.max_depth) return;  // v stays eval_heuristic() == 0
    v = terminal_e(meta & 3u);
    t.stat += static_cast<uint32_t>(lane == kStatTerminal);
  };
  if (cur_meta >= kMaxBlockId) {
    at_terminal(cur_meta);
  } else {
    bool on_pred = !GENERIC && pred_len > 0u;
    BB pos = root;  // position of the level being resolved
    float sq = 0.0f;
    uint32_t pred_a = 8u;
    // on_pred: continue at the first level >= `from` that is not held.  Returns false when every remaining
    // level is held: the walk ends in the finished game the previous simulation ended in.
    auto jump = [&](uint32_t from) -> bool {
      // (levels 31 and up are never held: from there on the walk goes level by level)
      const uint32_t rest = from < 32u ? (nhold >> from) << from : 0u;
      const uint32_t f = from < 32u ? (rest ? static_cast<uint32_t>(__ffs(static_cast<int>(rest))) - 1u : 32u) : from;
      if (f >= pred_len) {  // (only when pred_len <= 32: level 31 and up are never held)
        plen = pred_len;
        cur_slot = t.leaf_slot;
        cur_meta = t.leaf_meta;
        return false;
      }
#ifdef AZB_HOLD_VERIFY  // debug build: every skipped level must select its recorded edge in the exact evaluation
      for (uint32_t l = from; l < f && l < pred_len; ++l) {
        const uint4 ev = *reinterpret_cast<const uint4*>(t.path + l);
        uint4 w;
        uint32_t nn;
        float u, ex, cp;
        bool ok;
        load_edges(t, ev.y, la, w, nn);
        score_edges<false>(t, p.cpuct_f, __uint_as_float(ev.w), grp0, la, w, nn, u, ex, cp, ok);
        const float mx = redux_max_f32(u);
        const uint32_t ball = __ballot_sync(kFull, ok && u == mx);
        if (ball == 0u || bfind_u32(ball) != (ev.x & 7u)) t.error = kErrInternal;
      }
#endif
      const uint4 e = *reinterpret_cast<const uint4*>(t.path + f);  // {sa, blk, n, sqrt}
      plen = f;
      cur_slot = e.x >> 3;
      cur_meta = e.y;
      par_n = e.z;
      sq = __uint_as_float(e.w);
      pred_a = e.x & 7u;
      if (f) pos = BB{t.path[f].cur, t.path[f].opp};
      return true;
    };
    bool walking = true;
    if (on_pred) {
      walking = jump(0u);
      if (!walking) at_terminal(cur_meta);
    } else {
      par_n = ld_n(t, root_slot);
    }
    uint32_t guard = 0u;
    while (walking) {
      if (++guard > 4u * kPathCap) { t.error = kErrInternal; return true; }  // (a walk has at most 42 moves + link hops)
      if (GENERIC && depth > p.max_depth) break;  // :241-244 (+F6): v = eval_heuristic() == 0
      // ---- best_child of the current node; max_by keeps the LAST maximum (node.rs:366).  An
      // empty / all-NaN candidate set (node.rs:367 unwrap panics) is detected after the walk
      // (empty_seen); the walk itself stays in bounds. ----
      const uint32_t blk = cur_meta;
      uint4 w;
      uint32_t nn;
      float u, ex, cp;
      bool ok;
      load_edges(t, blk, la, w, nn);
      if (GENERIC || !on_pred) sq = sqrt_count(__fadd_rn(static_cast<float>((par_n + 1u) & 0xFFFFu), kEps));
      score_edges<GENERIC>(t, p.cpuct_f, sq, grp0, la, w, nn, u, ex, cp, ok);
      const float mx = redux_max_f32(u);
      const uint32_t ball = __ballot_sync(kFull, ok && u == mx);
      empty_seen |= (ball == 0u);
      const uint32_t a = bfind_u32(ball | 1u);
      const uint32_t ch_meta = __shfl_sync(kFull, w.w, a);
      // node_path.push(current_head_id) (:270 / F3) together with the action taken.  Every lane
      // stores the same words to the same address (cheaper than electing a lane)
      *reinterpret_cast<uint2*>(t.path + plen) = make_uint2((cur_slot << 3) | a, blk);
      if (!GENERIC) {
        // the hold bound of this level (backup_path): maxima over the edges NOT taken, with the slack that
        // covers every rounding between here and the bound's use (|q| <= 1.01, s1 <= 2 s0)
        const bool other = ok && la != a;
        const float m0 = redux_max_f32(other ? u : neg_inf);
        const float e0 = redux_max_f32(other ? ex : 0.0f);
        const float slack = __fmul_rn(__fadd_rn(__fadd_rn(fabsf(m0), 2.0f), __fmul_rn(4.0f, e0)), 9.5367431640625e-07f);
        float rs;
        asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(rs) : "f"(sq));
        const float e0s = __fmul_rn(__fmul_rn(e0, rs), 1.0000038146972656f);
        const float cpa = __shfl_sync(kFull, cp, a);
        *reinterpret_cast<float4*>(&t.path[plen].m0) =
            make_float4(m0 > neg_inf ? __fadd_rn(m0, slack) : neg_inf, e0s, sq, cpa);
      }
      plen++;
      // get_next_state + get_canonical_form (:284-287 with F4, F10) for the edge taken: the child's
      // position, kept with its path level (a link leads to the same position's owner)
      pos = play_canonical(pos, static_cast<int>(a));
      *reinterpret_cast<uint4*>(&t.path[plen].cur) =
          make_uint4(static_cast<uint32_t>(pos.cur), static_cast<uint32_t>(pos.cur >> 32),
                     static_cast<uint32_t>(pos.opp), static_cast<uint32_t>(pos.opp >> 32));
      if (!GENERIC && on_pred) {
        if (a == pred_a && plen < pred_len) {  // confirmed: the next recorded level is this edge's child
          if (jump(plen)) continue;
          at_terminal(cur_meta);
          break;
        }
        on_pred = false;
      }
      if (ch_meta < kMaxBlockId) {  // an expanded child: descend (:269-274 + F2)
        cur_slot = blk * 8u + a;
        cur_meta = ch_meta;
        par_n = __shfl_sync(kFull, nn, a);
        if (GENERIC) depth++;
        continue;
      }
      if (ch_meta != kMetaPlaceholder) {  // a link or a finished game
        if (GENERIC) depth++;
        if (ch_meta == kMetaLink) {
          cur_slot = __shfl_sync(kFull, w.x, a);
          cur_meta = __shfl_sync(kFull, w.y, a);
          par_n = __shfl_sync(kFull, nn, a);
          if (cur_meta < kMaxBlockId) continue;
        } else {
          cur_slot = blk * 8u + a;
          cur_meta = ch_meta;
        }
        at_terminal(cur_meta);
        break;
      }
      // ---- the chosen child is a placeholder: upgrade it (:279-356) ----
      const uint32_t my_slot = blk * 8u + a;
      const BB S2 = pos;
      const uint64_t key2 = state_key(S2);
      // the home bucket's load is issued first; terminal test / evaluation overlap its latency
      const uint4 e_home = tt_load_home(t, p.bucket_mask, key2, lane);
      const int code = game_ended_code_warp(t, S2, lane);
      const uint32_t vm = valid_mask(S2.cur | S2.opp);
      float pi = 0.0f, val = 0.0f;
      const bool inline_eval = !code && ev_kind < AZB_EVAL_NNET;
      if (inline_eval) evaluate_masked(t, ev_kind, S2, vm, lane, pi, val);
      uint32_t o_slot, o_meta, ins;
      if (tt_find(t, p.bucket_mask, key2, e_home, lane, o_slot, o_meta, ins)) {
        // upgrade -> Some(false): the slot becomes a link (node.rs:284-289); continue from the
        // owner without incrementing depth (async_mcts.rs:293-299)
        if (lane == static_cast<int>(a)) t.blocks[my_slot] = make_uint4(o_slot, o_meta, w.z, kMetaLink);
        t.stat += static_cast<uint32_t>(lane == kStatDupLinks);
        __syncwarp();
        cur_slot = o_slot;
        cur_meta = o_meta;
        par_n = ld_n(t, o_slot);
        if (cur_meta < kMaxBlockId) continue;
        at_terminal(cur_meta);
        break;
      }
      if (ins == 0xFFFFFFFFu) { t.error = kErrTable; return true; }
      if (empty_seen) { t.error = kErrInternal; return true; }
      // upgrade -> Some(true) (node.rs:290-322)
      if (code) {  // repair F5: terminal leaf, the net is skipped
        const uint32_t new_meta = kMetaTerminal | static_cast<uint32_t>(code);
        v = terminal_e(static_cast<uint32_t>(code));
        if (lane == static_cast<int>(a)) reinterpret_cast<uint32_t*>(t.blocks + my_slot)[3] = new_meta;
        tt_insert(t, ins, key2, my_slot, new_meta, lane);
        t.n_owners++;
        t.stat += static_cast<uint32_t>(lane == kStatTerminal || lane == kStatExpansions);
        cur_slot = my_slot;  // :309 visit() of the fresh node happens in its backup
        cur_meta = new_meta;
        end_level = 0;
        break;
      }
      if (t.n_blocks >= p.cap_blocks) { t.error = kErrBlocks; return true; }
      pd.kind = kPendExpand;
      pd.my_slot = my_slot;
      pd.new_meta = t.n_blocks++;
      pd.vm = vm;
      pd.plen = plen;
      pd.ins = ins;
      pd.levels = plen;
      pd.key = key2;
      if (!inline_eval) {
        leaf = S2;
        return false;
      }
      finish_expand(t, p, pd, pi, val, lane, /*normalised=*/true, /*predict=*/!GENERIC,
                    /*safe_prior=*/ev_kind == AZB_EVAL_UNIFORM && uniform_prior_is_safe(p.cpuct_f));
      return true;
    }
  }
  if (empty_seen) { t.error = kErrInternal; return true; }
  if (!GENERIC) {
    backup_path<true>(t, p, plen, cur_slot, v, plen + end_level, lane);
    t.pred_len = plen;  // the walked nodes; the finished game (or nothing, at the root) below them:
    t.leaf_slot = cur_slot;
    t.leaf_meta = cur_meta;
  } else {
    backup_path<false>(t, p, plen, cur_slot, v, plen + end_level, lane);
  }
  return true;
}

__device__ __forceinline__ bool one_sim(WarpTree& t, const SearchParams& p, int ev_kind, BB root,
                                        uint32_t root_slot, uint32_t root_meta, int lane,
                                        Pending& pd, BB& leaf) {
  if (t.slow || p.max_depth < 43u) return one_sim_impl<true>(t, p, ev_kind, root, root_slot, root_meta, lane, pd, leaf);
  return one_sim_impl<false>(t, p, ev_kind, root, root_slot, root_meta, lane, pd, leaf);
}

// ---- tree-parallel search with virtual loss: num_sim_threads = K > 1 (async_mcts.rs:191-217, node.rs:77-92,359-365) ----
// The reference runs K OS threads per tree whose interleaving is the scheduler's, so its result is not reproducible.
// Here (and in the CPU oracle that the tests compare with, bit for bit) the K threads run in ONE fixed interleaving, wave by wave: the K walks of a
// wave one after the other — each sees the virtual losses (visit() = N + 1, VL + 1 on every node of a path) and the locks
// of the earlier ones —, then the wave's evaluations (inline, or K leaves per tree in one network batch: a K-th of the
// rounds), then the K backups in thread order.  Nothing is written to the tree's counters during the walks: the
// virtual counter of a node is its counter at rest + (in-flight walks through it) * 0x10001, looked up in the wave's
// path table in shared memory.  Repairs that only matter here: F17 every child Locked (node.rs:367 unwraps None) -> the
// walk ends at the node with v = 0; F18 a node whose evaluation is pending in this wave reached through a link / a
// duplicate (the reference reads its missing policy, node.rs:354) -> the walk ends there and shares that evaluation.
constexpr int kMaxWave = 8;
enum : uint32_t { kWvEval = 1u, kWvShare = 2u, kWvRoot = 4u };
struct __align__(16) WaveSim {  // 256 bytes; kMaxWave of them fit the warp's path area in shared memory
  uint32_t plen;    // path[0 .. plen) = the nodes walked through (resolved owner slots), path[plen] = where the walk stopped
  uint32_t flags;   // kWvEval: owns a pending evaluation of path[plen]; kWvShare: shares simulation `share`'s; kWvRoot: F1
  uint32_t share;
  float v;          // the value backed up at path[plen]
  uint32_t my_slot, new_meta, vm, leaf_ref;  // the pending expansion; leaf_ref: its row in the network batch (rounds)
  uint64_t key;
  uint64_t cur, opp;  // the position to evaluate
  uint32_t levels, pad;
  uint32_t path[kPathCap];
};
static_assert(sizeof(WaveSim) == 256, "WaveSim layout");
static_assert(sizeof(WaveSim) * kMaxWave <= sizeof(PathEnt) * kPathCap, "the wave table lives in the path area");

// One walk of a wave (simulation `tix`).  Inline evaluators finish the expansion at once
// (the node still counts as Locked until the wave ends); AZB_EVAL_NNET leaves the new node without policy.
__device__ __forceinline__ void wave_select(WarpTree& t, const SearchParams& p, int ev_kind, BB root, uint32_t root_slot,
                                            uint32_t root_meta, WaveSim* ws, uint32_t tix, int lane) {
  const float neg_inf = __uint_as_float(0xFF800000u);
  const uint32_t la = lane & 7u;
  const bool grp0 = lane < 8;
  uint32_t cur_slot = root_slot, cur_meta = root_meta, par_n = ld_n(t, root_slot);
  uint32_t depth = 0, plen = 0, levels = 0, flags = 0, share = 0;
  uint32_t my_slot = 0, new_meta = 0, vm = 0;
  uint64_t key = 0;
  float v = 0.0f;
  BB pos = root;
  // in-flight walks k < tix whose node at path index `idx` is `slot`
  auto inflight_at = [&](uint32_t idx, uint32_t slot) -> uint32_t {
    uint32_t n = 0;
    for (uint32_t k = 0; k < tix; ++k) n += (ws[k].plen >= idx && ws[k].path[idx] == slot) ? 1u : 0u;
    return n;
  };
  for (;;) {
    levels++;
    if (levels > 4u * kPathCap) { t.error = kErrInternal; break; }
    if (depth > p.max_depth) break;                     // :241-244 (+F6): v = eval_heuristic() == 0
    if (cur_meta >= kMaxBlockId) {                      // :246-249 (+F6)
      v = terminal_e(cur_meta & 3u);
      t.stat += static_cast<uint32_t>(lane == kStatTerminal);
      break;
    }
    {  // F18 / F1: a node without a usable policy
      uint32_t owner = 0xFFFFFFFFu;
      for (uint32_t k = 0; k < tix; ++k)
        if ((ws[k].flags & kWvEval) && ws[k].path[ws[k].plen] == cur_slot) owner = k;
      if (owner != 0xFFFFFFFFu) { flags = kWvShare; share = owner; break; }
      if (!(block_flags(t, cur_meta) & kFlagHasPolicy)) {
        flags = kWvEval | kWvRoot;
        my_slot = cur_slot;
        new_meta = cur_meta;
        key = state_key(pos);
        break;
      }
    }
    // ---- best_child with the virtual counters (node.rs:343-370) ----
    const uint32_t blk = cur_meta;
    const uint4* bp = t.blocks + static_cast<size_t>(blk) * 8u;
    const uint4 w = bp[la];
    uint32_t nn = reinterpret_cast<const uint16_t*>(bp + 7)[la];
    bool ok = grp0 && la < 7u && w.w != kMetaInvalid;
    uint32_t wv = w.x, rslot = blk * 8u + la;
    if (ok && w.w == kMetaLink) {  // resolve(): statistics come from the owner (node.rs:179-201)
      rslot = w.x;
      wv = ld_w(t, rslot);
      nn = ld_n(t, rslot);
    }
    const uint32_t vl_c = ok ? inflight_at(plen + 1u, rslot) : 0u;
    uint32_t locked = 0u;  // the RAW child slot is the pending expansion of an earlier walk (NodeState::Locked)
    for (uint32_t k = 0; k < tix; ++k)
      locked |= ((ws[k].flags & kWvEval) && !(ws[k].flags & kWvRoot) && ws[k].my_slot == blk * 8u + la) ? 1u : 0u;
    const uint32_t vl_x = inflight_at(plen, cur_slot);
    const float sq = AZB_FSQRT(AZB_FADD(static_cast<float>((par_n + vl_x + 1u) & 0xFFFFu), kEps));  // N after this visit()
    const uint64_t vc = counter_pack(wv, nn) + static_cast<uint64_t>(vl_c) * kVisit;
    float u = ok ? puct_u(vc, __uint_as_float(w.z), sq, p.cpuct_f) : neg_inf;
    float mx = redux_max_f32(u);
    uint32_t ball = __ballot_sync(kFull, ok && u == mx);
    if (ball == 0u) { t.error = kErrInternal; break; }
    uint32_t a = bfind_u32(ball);
    if (__shfl_sync(kFull, locked, a)) {  // :253-257: retry without the Locked children
      ok = ok && !locked;
      u = ok ? u : neg_inf;
      mx = redux_max_f32(u);
      ball = __ballot_sync(kFull, ok && u == mx);
      if (ball == 0u) break;  // F17: v = 0 at this node
      a = bfind_u32(ball);
    }
    const uint32_t ch_meta = __shfl_sync(kFull, w.w, a);
    if (lane == 0) ws[tix].path[plen] = cur_slot;  // node_path.push (:270 / F3)
    plen++;
    pos = play_canonical(pos, static_cast<int>(a));
    if (ch_meta != kMetaPlaceholder) {  // Exists: descend (:269-274 + F2); a finished game is noticed at the loop's top
      if (ch_meta == kMetaLink) {
        cur_slot = __shfl_sync(kFull, w.x, a);
        cur_meta = __shfl_sync(kFull, w.y, a);
      } else {
        cur_slot = blk * 8u + a;
        cur_meta = ch_meta;
      }
      par_n = __shfl_sync(kFull, nn, a);
      depth++;
      continue;
    }
    // ---- the chosen child is a placeholder: lock + upgrade (:260-299, node.rs:272-326) ----
    my_slot = blk * 8u + a;
    key = state_key(pos);
    const uint4 e_home = tt_load_home(t, p.bucket_mask, key, lane);
    const int code = game_ended_code_warp(t, pos, lane);
    vm = valid_mask(pos.cur | pos.opp);
    uint32_t o_slot, o_meta, ins;
    if (tt_find(t, p.bucket_mask, key, e_home, lane, o_slot, o_meta, ins)) {  // Some(false): a link; go on from the owner
      if (lane == static_cast<int>(a)) t.blocks[my_slot] = make_uint4(o_slot, o_meta, w.z, kMetaLink);
      t.stat += static_cast<uint32_t>(lane == kStatDupLinks);
      __syncwarp();
      cur_slot = o_slot;
      cur_meta = o_meta;
      par_n = ld_n(t, o_slot);
      continue;  // (depth not incremented, async_mcts.rs:293-299)
    }
    if (ins == 0xFFFFFFFFu) { t.error = kErrTable; break; }
    cur_slot = my_slot;
    if (code) {  // repair F5: a finished game, the net is skipped
      new_meta = kMetaTerminal | static_cast<uint32_t>(code);
      v = terminal_e(static_cast<uint32_t>(code));
      if (lane == static_cast<int>(a)) reinterpret_cast<uint32_t*>(t.blocks + my_slot)[3] = new_meta;
      tt_insert(t, ins, key, my_slot, new_meta, lane);
      t.n_owners++;
      t.stat += static_cast<uint32_t>(lane == kStatTerminal || lane == kStatExpansions);
      __syncwarp();
      break;
    }
    if (t.n_blocks >= p.cap_blocks) { t.error = kErrBlocks; break; }
    new_meta = t.n_blocks++;
    flags = kWvEval;
    float pi = 0.0f, val = 0.0f;
    uint32_t bflags = 0u;
    if (ev_kind < AZB_EVAL_NNET) {  // inline evaluators know the policy now; nobody reads it before the wave ends
      evaluate_masked(t, ev_kind, pos, vm, lane, pi, val);
      v = -val;
      bflags = kFlagHasPolicy;
      t.stat += static_cast<uint32_t>(lane == kStatEvals);
    }
    write_child_block(t, new_meta, vm, pi, bflags, lane);
    if (lane == static_cast<int>(a)) reinterpret_cast<uint32_t*>(t.blocks + my_slot)[3] = new_meta;
    tt_insert(t, ins, key, my_slot, new_meta, lane);
    t.n_owners++;
    t.stat += static_cast<uint32_t>(lane == kStatExpansions);
    __syncwarp();
    break;
  }
  if (lane == 0) {
    WaveSim& me = ws[tix];
    me.path[plen] = cur_slot;
    me.plen = plen;
    me.flags = flags;
    me.share = share;
    me.v = v;
    me.my_slot = my_slot;
    me.new_meta = new_meta;
    me.vm = vm;
    me.leaf_ref = 0u;
    me.key = key;
    me.cur = pos.cur;
    me.opp = pos.opp;
    me.levels = levels;
  }
  __syncwarp();
}

// The network's answer (or the inline evaluator's, for an F1 root) for the pending evaluation of walk `k`: mask +
// normalise, set_policy, unlock (async_mcts.rs:317-353).  Lane a holds the raw pi[a].
__device__ __forceinline__ void wave_set_policy(WarpTree& t, const SearchParams& p, WaveSim* ws, uint32_t k, float pi, float val,
                                                int lane) {
  const uint32_t blk = ws[k].new_meta;
  uint4* bp = t.blocks + static_cast<size_t>(blk) * 8u;
  const uint32_t m = lane < 7 ? bp[lane].w : kMetaInvalid;
  const uint32_t vm = __ballot_sync(kFull, m != kMetaInvalid) & 0x7Fu;
  pi = mask_normalise(pi, vm, lane);
  if (lane < 7) reinterpret_cast<float*>(bp + lane)[2] = pi;
  if (lane == 7) reinterpret_cast<uint32_t*>(bp + 7)[3] |= kFlagHasPolicy << 16;
  if (lane == 0) ws[k].v = -val;
  t.stat += static_cast<uint32_t>(lane == kStatEvals);
  __syncwarp();
}

// The K backups of a wave, in thread order (:361-370): visit() + unvisit() folded per path node, as in backup_path.
__device__ __forceinline__ void wave_backup(WarpTree& t, const SearchParams& p, WaveSim* ws, uint32_t K, int lane) {
  const bool alternate = !(p.quirks & AZB_Q2_BACKUP_NO_ALTERNATE);
  __syncwarp();
  for (uint32_t k = 0; k < K; ++k) {
    const uint32_t plen = ws[k].plen;
    const float v = (ws[k].flags & kWvShare) ? ws[ws[k].share].v : ws[k].v;
    for (uint32_t base = 0; base <= plen; base += 32u) {
      const uint32_t l = base + lane;
      if (l <= plen) {
        const bool neg = alternate && ((plen - l) & 1u);
        const BackupRegs r = backup_prepare(t, ws[k].path[l], __fmul_rn(neg ? -1.0f : 1.0f, v), p.quirks);
        backup_commit(t, ws[k].path[l], r);
      }
    }
    t.stat += static_cast<uint32_t>(lane == kStatSims) + (lane == kStatLevels ? ws[k].levels : 0u);
    __syncwarp();  // the next walk's path may share nodes with this one
  }
  t.pred_len = 0u;
}

// One wave with an inline evaluator, start to end.
__device__ __forceinline__ void wave_inline(WarpTree& t, const SearchParams& p, int ev_kind, BB root, uint32_t root_slot,
                                            uint32_t root_meta, uint32_t K, int lane) {
  WaveSim* ws = reinterpret_cast<WaveSim*>(t.path);
  for (uint32_t k = 0; k < K && !t.error; ++k) wave_select(t, p, ev_kind, root, root_slot, root_meta, ws, k, lane);
  if (t.error) return;
  for (uint32_t k = 0; k < K; ++k)
    if ((ws[k].flags & (kWvEval | kWvRoot)) == (kWvEval | kWvRoot)) {  // F1: the root's own evaluation
      float pi, val;
      evaluate_inline(ev_kind, BB{ws[k].cur, ws[k].opp}, lane, pi, val);
      wave_set_policy(t, p, ws, k, pi, val, lane);
    }
  wave_backup(t, p, ws, K, lane);
}

// search (:191-217) with num_threads = 1 and a fused evaluator: nsims simulations from `root`.
// WAVE is a compile-time switch: the deterministic kernels carry none of the wave code (inlined into them it cost the
// persistent kernel 12 % at BASELINE config 2 without a single spill — code size / layout of the hot loop).
template <bool WAVE>
__device__ __forceinline__ void run_sims(WarpTree& t, const SearchParams& p, int ev_kind, BB root,
                                         uint32_t root_slot, uint32_t root_meta, uint32_t nsims,
                                         int lane) {
  if constexpr (WAVE) {  // tree-parallel mode: waves of K simulations (num_sims % K == 0 is checked at setup, :192)
    for (uint32_t sim = 0; sim < nsims && !t.error; sim += p.num_threads)
      wave_inline(t, p, ev_kind, root, root_slot, root_meta, p.num_threads, lane);
    return;
  }
  uint32_t sim = 0;
  // Repair F1: an existing, non-terminal node that was never evaluated (a stand-alone root at
  // its first visit) is evaluated when first reached; this consumes one simulation.  Only a
  // root can be in that state: every other node is evaluated by the simulation that creates it.
  if (nsims > 0 && root_needs_eval(t, root_meta)) {
    float pi, val;
    evaluate_inline(ev_kind, root, lane, pi, val);
    finish_root_eval(t, p, root_slot, root_meta, pi, val, lane);
    sim = 1;
  }
  Pending pd;
  BB leaf;
  for (; sim < nsims && !t.error; ++sim) one_sim(t, p, ev_kind, root, root_slot, root_meta, lane, pd, leaf);
}

// counts[a] = N of the (resolved) root child (async_mcts.rs:87-94 with F7).  Lane a returns it.
__device__ __forceinline__ uint32_t root_child_count(const WarpTree& t, uint32_t root_meta, int lane) {
  if (!meta_is_block(root_meta) || lane >= 7) return 0u;
  const uint32_t slot = root_meta * 8u + lane;
  const uint4 w = t.blocks[slot];
  if (w.w == kMetaInvalid) return 0u;
  return ld_n(t, w.w == kMetaLink ? w.x : slot);
}

// pi from counts (async_mcts.rs:96-114 with F8, App. B.7).  temp == 0: one-hot on the last
// arg-max over all 7 entries; temp == 1: counts / sum; other temps: powf (not bit-pinned).
__device__ __forceinline__ float counts_to_pi(uint32_t cnt, float temp, int lane) {
  if (temp == 0.0f) {
    const uint32_t mx = __reduce_max_sync(kFull, lane < 7 ? cnt : 0u);
    const uint32_t ball = __ballot_sync(kFull, lane < 7 && cnt == mx);
    return lane == 31 - __clz(ball) ? 1.0f : 0.0f;
  }
  float c = static_cast<float>(cnt);
  if (temp != 1.0f) c = powf(c, __fdiv_rn(1.0f, temp));
  if (lane >= 7) c = 0.0f;
  float s = 0.0f;
#pragma unroll
  for (int a = 0; a < 7; ++a) s = __fadd_rn(s, __shfl_sync(kFull, c, a));
  return __fdiv_rn(c, s);
}

// ---- Philox-4x32-10 + choose_weighted (stands in for rand 0.7 SmallRng, coach.rs:137-138) ----
__device__ __forceinline__ float philox_uniform01(uint64_t seed, uint64_t game_id, uint32_t ply,
                                                  uint32_t purpose) {
  uint32_t c0 = ply, c1 = purpose, c2 = static_cast<uint32_t>(seed >> 32),
           c3 = static_cast<uint32_t>(game_id >> 32);
  uint32_t k0 = static_cast<uint32_t>(seed), k1 = static_cast<uint32_t>(game_id);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
    const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return __fmul_rn(static_cast<float>(c0 >> 8), 1.0f / 16777216.0f);
}

// First index whose running (sequential f32) weight exceeds u*total; zero weights are skipped.
__device__ __forceinline__ int choose_weighted(float w_lane, float u) {
  float total = 0.0f;
#pragma unroll
  for (int a = 0; a < 7; ++a) total = __fadd_rn(total, __shfl_sync(kFull, w_lane, a));
  const float tt = __fmul_rn(u, total);
  float acc = 0.0f;
  int last = -1, chosen = -1;
#pragma unroll
  for (int a = 0; a < 7; ++a) {
    const float w = __shfl_sync(kFull, w_lane, a);
    if (w > 0.0f) {
      acc = __fadd_rn(acc, w);
      last = a;
      if (chosen < 0 && tt < acc) chosen = a;
    }
  }
  return chosen < 0 ? last : chosen;
}

}  // namespace azb

// Warp-per-tree Monte-Carlo tree search (kernel families K2-K4, K5 for the fused evaluators).
//
// Replaces src/async_mcts.rs (search_iteration :219-371, get_action_prob :74-115) and the
// NodeStore operations it calls (src/node.rs:179-370) in the reference's deterministic
// mode (num_sim_threads = 1), with the repair list F1-F8, F12 of SURVEY.md App. A/C.
//
// One warp owns one tree and runs one simulation at a time, so every f32 rounding and every
// tie-break happens in the reference's order and results are bit-exact with the oracle.
// Per level the warp reads ONE 128-byte block (lanes 0-6 = the 7 edges, lane 7 = header),
// resolves transposition links with one extra 8-byte load, evaluates PUCT per lane and
// picks the last maximum with redux.max + ballot.  visit()/unvisit() of the reference are
// folded: counters travel down the path in registers and are written once, at backup.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "c4_bitboard.cuh"
#include "tree.cuh"

namespace azb {

constexpr unsigned kFull = 0xFFFFFFFFu;

enum : uint32_t { kErrNone = 0, kErrBlocks = 1, kErrTable = 2, kErrInternal = 3 };
enum : int { kStatSims = 0, kStatLevels, kStatExpansions, kStatTerminal, kStatDupLinks, kStatEvals, kNumStats };

struct SearchParams {
  uint32_t cap_blocks;   // blocks per tree
  uint32_t bucket_mask;  // transposition table: (bucket_mask+1) buckets of 8 entries
  uint32_t num_sims;
  uint32_t max_depth;
  float cpuct_f;
  uint32_t quirks;
  uint32_t temp_threshold;
  uint32_t pad;
  uint64_t seed;
};

// Persistent per-tree record (HBM).  Lives across get_action_prob calls / plies.
struct TreeRec {
  uint32_t n_blocks;
  uint32_t n_owners;  // == NodeStore.seen.len()
  uint32_t error;
  uint32_t pad;
  uint32_t stat[8];
};

// Per-warp view of one tree.  Every member is warp-uniform except `stat` (lane k = stat k).
struct WarpTree {
  uint4* blocks;  // slot id indexes this directly (16-byte slots, 8 per block)
  uint4* table;   // HashEntry as uint4 {key.lo, key.hi, slot, meta}
  uint32_t n_blocks, n_owners, error;
  uint32_t stat;
};

__device__ __forceinline__ uint64_t ld_counter(const WarpTree& t, uint32_t slot) {
  return *reinterpret_cast<const uint64_t*>(t.blocks + slot);
}
__device__ __forceinline__ void st_counter(const WarpTree& t, uint32_t slot, uint64_t c) {
  *reinterpret_cast<uint64_t*>(t.blocks + slot) = c;
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(kFull, static_cast<uint32_t>(v), src);
  uint32_t hi = __shfl_sync(kFull, static_cast<uint32_t>(v >> 32), src);
  return (static_cast<uint64_t>(hi) << 32) | lo;
}
// Order-preserving f32 -> u32 (finite values; -0.0 is canonicalised to +0.0 first so that it
// compares Equal to +0.0 like partial_cmp does, node.rs:366).
__device__ __forceinline__ uint32_t ordered_key(float u) {
  uint32_t b = __float_as_uint(__fadd_rn(u, 0.0f));
  return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u);
}

// ---- transposition table (NodeStore.seen): buckets of 8 entries = one 128-byte line ------
// Returns true and (slot, meta) when `key` is present; otherwise `ins` is the first free
// entry on the probe path (0xFFFFFFFF when the table is full).
__device__ __forceinline__ bool tt_find(const WarpTree& t, uint32_t bucket_mask, uint64_t key,
                                        int lane, uint32_t& slot, uint32_t& meta, uint32_t& ins) {
  uint32_t b = hash_bucket(key, bucket_mask);
  for (uint32_t probe = 0; probe <= bucket_mask; ++probe) {
    uint4 e = make_uint4(1u, 0u, 0u, 0u);
    if (lane < 8) e = t.table[b * 8u + lane];
    uint64_t k = (static_cast<uint64_t>(e.y) << 32) | e.x;
    uint32_t hit = __ballot_sync(kFull, lane < 8 && k == key);
    if (hit) {
      int l = __ffs(hit) - 1;
      slot = __shfl_sync(kFull, e.z, l);
      meta = __shfl_sync(kFull, e.w, l);
      return true;
    }
    uint32_t emp = __ballot_sync(kFull, lane < 8 && k == 0ull);
    if (emp) {
      ins = b * 8u + (__ffs(emp) - 1);
      return false;
    }
    b = (b + 1u) & bucket_mask;
  }
  ins = 0xFFFFFFFFu;
  return false;
}
__device__ __forceinline__ void tt_insert(const WarpTree& t, uint32_t ins, uint64_t key,
                                          uint32_t slot, uint32_t meta, int lane) {
  if (lane == 0)
    t.table[ins] = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32), slot, meta);
}

// ---- leaf evaluators fused into the search (NNet::predict, src/nnet.rs:40-44) --------------
// Lane a (< 7) returns pi[a]; `v` is warp-uniform.
template <int EVAL>
__device__ __forceinline__ void evaluate_inline(BB s, int lane, float& pi, float& v) {
  if (EVAL == AZB_EVAL_UNIFORM) {  // examples/connect_four.rs:34-38
    pi = __fdiv_rn(1.0f, 7.0f);
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

__device__ __forceinline__ void write_child_block(const WarpTree& t, uint32_t blk, uint64_t key,
                                                  uint32_t vm, float prior, uint32_t flags,
                                                  uint32_t self_slot, int lane) {
  uint4 out;
  if (lane < 7)
    out = make_uint4(static_cast<uint32_t>(kCounterInit), static_cast<uint32_t>(kCounterInit >> 32),
                     __float_as_uint(prior), ((vm >> lane) & 1u) ? kMetaPlaceholder : kMetaInvalid);
  else
    out = make_uint4(static_cast<uint32_t>(key), static_cast<uint32_t>(key >> 32),
                     flags | (vm << 8), self_slot);
  if (lane < 8) t.blocks[blk * 8u + lane] = out;
}

// push + upgrade of a state that is not in the tree, as a stand-alone root
// (NodeStore::new / from_root, node.rs:156-177; repair F12).  The root's own slot is slot 0
// of a holder block.  The node is NOT evaluated here (repair F1 does it at first visit).
__device__ __forceinline__ bool make_root(WarpTree& t, const SearchParams& p, BB s, int lane,
                                          uint32_t& root_slot, uint32_t& root_meta) {
  const uint64_t key = state_key(s);
  uint32_t o_slot, o_meta, ins;
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
    write_child_block(t, root_meta, key, valid_mask(s.cur | s.opp), 0.0f, 0u, root_slot, lane);
  }
  uint4 out = make_uint4(static_cast<uint32_t>(kCounterInit), static_cast<uint32_t>(kCounterInit >> 32),
                         0u, lane == 0 ? root_meta : kMetaInvalid);
  if (lane == 7) out = make_uint4(0u, 0u, kFlagRootHolder, root_slot);
  if (lane < 8) t.blocks[holder * 8u + lane] = out;
  tt_insert(t, ins, key, root_slot, root_meta, lane);
  t.n_owners++;
  __syncwarp();
  return true;
}

// ---- search_iteration x nsims (async_mcts.rs:191-371) ------------------------
template <int EVAL>
__device__ __forceinline__ void run_sims(WarpTree& t, const SearchParams& p, BB root,
                                         uint32_t root_slot, uint32_t root_meta, uint32_t nsims,
                                         int lane) {
  const bool alternate = !(p.quirks & AZB_Q2_BACKUP_NO_ALTERNATE);
  for (uint32_t sim = 0; sim < nsims; ++sim) {
    uint32_t cur_slot = root_slot, cur_meta = root_meta;
    uint64_t cur_cnt = ld_counter(t, cur_slot);
    BB S = root;
    uint32_t depth = 0, plen = 0;
    uint32_t ps0 = 0, ps1 = 0;  // node_path, lane l holds entries l and l+32
    uint64_t pc0 = 0, pc1 = 0;
    float v = 0.0f;
    for (;;) {
      if (lane == kStatLevels) t.stat++;
      if (depth > p.max_depth) {  // :241-244 (+F6); eval_heuristic() == 0 for connect-four
        cur_cnt += kVisit;
        v = 0.0f;
        break;
      }
      if (meta_is_terminal(cur_meta)) {  // :246-249 (+F6)
        cur_cnt += kVisit;
        v = terminal_e(cur_meta & 3u);
        if (lane == kStatTerminal) t.stat++;
        break;
      }
      uint4* bp = t.blocks + static_cast<size_t>(cur_meta) * 8u;
      uint4 w = make_uint4(0u, 0u, 0u, kMetaInvalid);
      if (lane < 8) w = bp[lane];
      const uint32_t flags = __shfl_sync(kFull, w.z, 7);
      if (!(flags & kFlagHasPolicy)) {  // repair F1: an existing node that was never evaluated
        cur_cnt += kVisit;
        float pi, val;
        evaluate_inline<EVAL>(S, lane, pi, val);
        pi = mask_normalise(pi, (flags >> 8) & 0x7Fu, lane);
        if (lane < 7) reinterpret_cast<float*>(bp + lane)[2] = pi;             // set_policy
        if (lane == 7) reinterpret_cast<uint32_t*>(bp + 7)[2] = flags | kFlagHasPolicy;
        if (lane == kStatEvals) t.stat++;
        v = -val;
        break;
      }
      cur_cnt += kVisit;  // :251 visit()
      // best_child (node.rs:343-370): parent N is read after the visit
      const float sq = __fsqrt_rn(__fadd_rn(static_cast<float>(counter_n(cur_cnt)), kEps));
      uint64_t ccnt = (static_cast<uint64_t>(w.y) << 32) | w.x;
      const float prior = __uint_as_float(w.z);
      const uint32_t meta = w.w;
      const bool ok = lane < 7 && meta != kMetaInvalid;
      uint32_t c_slot = cur_meta * 8u + lane, c_meta = meta;
      if (ok && meta == kMetaLink) {  // resolve(): statistics come from the owner (node.rs:179-201)
        c_slot = w.x;
        c_meta = w.y;
        ccnt = ld_counter(t, c_slot);
      }
      const float u = ok ? puct_u(ccnt, prior, sq, p.cpuct_f) : 0.0f;
      const uint32_t okey = ok ? ordered_key(u) : 0u;
      const uint32_t mx = __reduce_max_sync(kFull, okey);
      const uint32_t ball = __ballot_sync(kFull, ok && okey == mx);
      if (ball == 0u) { t.error = kErrInternal; return; }  // node.rs:367 unwrap on empty
      const int a = 31 - __clz(ball);  // max_by keeps the LAST maximum
      const uint32_t ch_raw = __shfl_sync(kFull, meta, a);
      const uint32_t ch_slot = __shfl_sync(kFull, c_slot, a);
      const uint32_t ch_meta = __shfl_sync(kFull, c_meta, a);
      const uint64_t ch_cnt = shfl64(ccnt, a);
      // node_path.push(current_head_id) (:270 / F3)
      if (lane == static_cast<int>(plen & 31u)) {
        if (plen < 32u) { ps0 = cur_slot; pc0 = cur_cnt; } else { ps1 = cur_slot; pc1 = cur_cnt; }
      }
      plen++;
      const BB S2 = play_canonical(S, a);  // :284-287 with F4, F10
      if (ch_raw == kMetaPlaceholder) {
        const uint32_t my_slot = cur_meta * 8u + static_cast<uint32_t>(a);
        const uint64_t key2 = state_key(S2);
        uint32_t o_slot, o_meta, ins;
        if (tt_find(t, p.bucket_mask, key2, lane, o_slot, o_meta, ins)) {
          // upgrade -> Some(false): the slot becomes a link (node.rs:284-289); continue from the
          // owner without incrementing depth (async_mcts.rs:293-299)
          if (lane == a) t.blocks[my_slot] = make_uint4(o_slot, o_meta, w.z, kMetaLink);
          if (lane == kStatDupLinks) t.stat++;
          __syncwarp();
          cur_slot = o_slot;
          cur_meta = o_meta;
          cur_cnt = ld_counter(t, o_slot);
          S = S2;
          continue;
        }
        if (ins == 0xFFFFFFFFu) { t.error = kErrTable; return; }
        // upgrade -> Some(true) (node.rs:290-322)
        const int code = game_ended_code(S2, p.quirks);
        uint32_t new_meta;
        if (code) {  // repair F5: terminal leaf, the net is skipped
          new_meta = kMetaTerminal | static_cast<uint32_t>(code);
          v = terminal_e(static_cast<uint32_t>(code));
          if (lane == kStatTerminal) t.stat++;
        } else {
          if (t.n_blocks >= p.cap_blocks) { t.error = kErrBlocks; return; }
          new_meta = t.n_blocks++;
          float pi, val;
          evaluate_inline<EVAL>(S2, lane, pi, val);
          const uint32_t vm = valid_mask(S2.cur | S2.opp);
          pi = mask_normalise(pi, vm, lane);
          write_child_block(t, new_meta, key2, vm, pi, kFlagHasPolicy, my_slot, lane);
          if (lane == kStatEvals) t.stat++;
          v = -val;  // :353
        }
        if (lane == a) reinterpret_cast<uint32_t*>(t.blocks + my_slot)[3] = new_meta;
        tt_insert(t, ins, key2, my_slot, new_meta, lane);
        t.n_owners++;
        if (lane == kStatExpansions) t.stat++;
        cur_slot = my_slot;
        cur_cnt = kCounterInit + kVisit;  // :309 visit() of the fresh node
        break;
      }
      cur_slot = ch_slot;
      cur_meta = ch_meta;
      cur_cnt = ch_cnt;
      S = S2;
      depth++;
    }
    // backup (:361-370): the leaf gets +v, then up node_path; Q2 corrected alternates the sign
    {
      if (lane < static_cast<int>(plen)) {
        const bool neg = alternate && ((plen - static_cast<uint32_t>(lane)) & 1u);
        st_counter(t, ps0, counter_unvisit(pc0, __fmul_rn(neg ? -1.0f : 1.0f, v), p.quirks));
      }
      if (lane + 32 < static_cast<int>(plen)) {
        const bool neg = alternate && ((plen - static_cast<uint32_t>(lane) - 32u) & 1u);
        st_counter(t, ps1, counter_unvisit(pc1, __fmul_rn(neg ? -1.0f : 1.0f, v), p.quirks));
      }
      if (lane == 0) st_counter(t, cur_slot, counter_unvisit(cur_cnt, __fmul_rn(1.0f, v), p.quirks));
      if (lane == kStatSims) t.stat++;
    }
    __syncwarp();
  }
}

// counts[a] = N of the (resolved) root child (async_mcts.rs:87-94 with F7).  Lane a returns it.
__device__ __forceinline__ uint32_t root_child_count(const WarpTree& t, uint32_t root_meta, int lane) {
  if (!meta_is_block(root_meta)) return 0u;
  uint4 w = make_uint4(0u, 0u, 0u, kMetaInvalid);
  if (lane < 7) w = t.blocks[static_cast<size_t>(root_meta) * 8u + lane];
  if (lane >= 7 || w.w == kMetaInvalid) return 0u;
  uint64_t c = (static_cast<uint64_t>(w.y) << 32) | w.x;
  if (w.w == kMetaLink) c = ld_counter(t, w.x);
  return counter_n(c);
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

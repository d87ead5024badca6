// Host side of libazb200.so: device memory, launches, and the C ABI of include/azb200.h.
// The reference's host layer is Rust (src/coach.rs, src/arena.rs); no Rust toolchain exists
// in the build image, so this is C++ behind the same entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "../../include/azb200.h"
#include "kernels.cuh"

using namespace azb;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define AZB_CUDA(expr)                                                                      \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess)                                                                  \
      return fail(AZB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e));       \
  } while (0)

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    bytes = 0;
  }
  cudaError_t ensure(size_t n) {
    if (n <= bytes) return cudaSuccess;
    release();
    cudaError_t e = cudaMalloc(&p, n);
    if (e == cudaSuccess) bytes = n;
    return e;
  }
  template <class T> T* as() const { return static_cast<T*>(p); }
};

uint32_t pow2_ceil(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return static_cast<uint32_t>(p);
}

int validate(const azb_config* cfg) {
  if (!cfg) return fail(AZB_ERR_INVALID, "config is NULL");
  if (cfg->num_sim_threads != 1)
    return fail(AZB_ERR_UNSUPPORTED,
                "num_sim_threads must be 1 (deterministic mode: one simulation in flight per tree)");
  if (cfg->num_sims == 0 || cfg->num_sims > 60000) return fail(AZB_ERR_INVALID, "num_sims out of range");
  if (cfg->evaluator != AZB_EVAL_UNIFORM && cfg->evaluator != AZB_EVAL_HASH)
    return fail(AZB_ERR_UNSUPPORTED, "evaluator: only the fused UNIFORM and HASH evaluators exist yet");
  if (cfg->mcts_reserve_size < 16) return fail(AZB_ERR_INVALID, "mcts_reserve_size too small");
  return AZB_OK;
}

// Pool geometry.  An expanded node costs one block; a reference slot budget of R node slots
// (mcts_reserve_size) allows at least R/7 expanded nodes.  A game can never allocate more
// than num_sims*42 + 43*2 blocks (<= 1 expansion per simulation, 42 plies, stand-alone roots).
SearchParams make_params(const azb_config& c, uint64_t searches_per_tree) {
  SearchParams p{};
  uint64_t by_reserve = c.mcts_reserve_size / 7 + 2;
  uint64_t by_work = c.num_sims * searches_per_tree + 2 * (searches_per_tree + 1) + 8;
  uint64_t cap = std::min(by_reserve, by_work);
  cap = std::min<uint64_t>(cap, (1u << 24) - 1);  // slot ids are packed with the action in 27+3 bits
  p.cap_blocks = static_cast<uint32_t>(cap);
  uint32_t entries = pow2_ceil(std::max<uint64_t>(64, cap + cap / 2));
  p.bucket_mask = entries / 8 - 1;
  p.num_sims = static_cast<uint32_t>(c.num_sims);
  p.max_depth = static_cast<uint32_t>(std::min<uint64_t>(c.max_depth, 0x7FFFFFFF));
  p.cpuct_f = static_cast<float>(c.cpuct);
  p.quirks = c.quirks;
  p.temp_threshold = static_cast<uint32_t>(std::min<uint64_t>(c.temp_threshold, 0x7FFFFFFF));
  p.seed = c.seed;
  return p;
}

size_t tree_bytes(const SearchParams& p) {
  return static_cast<size_t>(p.cap_blocks) * 128 + (static_cast<size_t>(p.bucket_mask) + 1) * 128 +
         sizeof(TreeRec);
}

struct TreePool {
  SearchParams p{};
  uint32_t n_trees = 0;
  DevBuf blocks, tables, recs;
  Pools pools{};
  int alloc(const SearchParams& sp, uint32_t n) {
    p = sp;
    n_trees = n;
    AZB_CUDA(blocks.ensure(static_cast<size_t>(n) * p.cap_blocks * 128));
    AZB_CUDA(tables.ensure(static_cast<size_t>(n) * (static_cast<size_t>(p.bucket_mask) + 1) * 128));
    AZB_CUDA(recs.ensure(static_cast<size_t>(n) * sizeof(TreeRec)));
    pools.blocks = blocks.as<uint4>();
    pools.tables = tables.as<uint4>();
    pools.recs = recs.as<TreeRec>();
    return AZB_OK;
  }
  int reset() {
    AZB_CUDA(cudaMemset(tables.p, 0, static_cast<size_t>(n_trees) * (static_cast<size_t>(p.bucket_mask) + 1) * 128));
    AZB_CUDA(cudaMemset(recs.p, 0, static_cast<size_t>(n_trees) * sizeof(TreeRec)));
    return AZB_OK;
  }
};

BB to_bb(const azb_c4_state& s) {  // canonical states only: +1 = side to move
  BB b{0, 0};
  for (int r = 0; r < 6; ++r)
    for (int c = 0; c < 7; ++c) {
      if (s.s[r][c] > 0) b.cur |= 1ull << (r * 7 + c);
      else if (s.s[r][c] < 0) b.opp |= 1ull << (r * 7 + c);
    }
  return b;
}

int capacity_error(uint32_t code) {
  switch (code) {
    case kErrBlocks: return fail(AZB_ERR_CAPACITY, "node pool exhausted (mcts_reserve_size; reference: node.rs:237 assert)");
    case kErrTable: return fail(AZB_ERR_CAPACITY, "transposition table full");
    default: return fail(AZB_ERR_INVALID, "internal search error (no selectable child / no weighted choice)");
  }
}

static int resident_trees(int device, uint32_t* out) {
  int per_sm = 0, sms = 0;
  // the search lives on L1 hits of the hot tree top: keep the unified L1/shared array as L1
  AZB_CUDA(cudaFuncSetAttribute(k_selfplay, cudaFuncAttributePreferredSharedMemoryCarveout,
                                8 /* percent: 7 CTAs x (768 B + 1 KB reserved) fit in the 16 KB configuration */));
  AZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_selfplay, kWarpsPerCta * 32, 0));
  AZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  *out = static_cast<uint32_t>(per_sm * sms * kWarpsPerCta);
  return AZB_OK;
}

}  // namespace

struct azb_mcts {
  azb_config cfg;
  TreePool pool;
  DevBuf d_states, d_counts, d_pi, d_u64;
};

struct azb_coach {
  azb_config cfg;
  TreePool pool;
  bool pool_ready = false;
  // last self-play call
  uint64_t n_games = 0, n_samples = 0;
  DevBuf plies, final_r, final_player, error, actions, counts, sample_state, sample_pi, stats, next_game;
  DevBuf offsets, out_boards, out_pis, out_vs;
  std::vector<uint32_t> h_plies;
  GameBufs g{};
};

extern "C" {

const char* azb_last_error(void) { return g_err.c_str(); }

int azb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

void azb_config_default(azb_config* c) {  // examples/connect_four.rs:55-71
  std::memset(c, 0, sizeof(*c));
  c->checkpoint_directory = "./checkpoint";
  c->mcts_reserve_size = 1000000;
  c->update_threshold = 0.6f;
  c->temp_threshold = 15;
  c->max_history_length = 20;
  c->max_queue_length = 200000;
  c->inference_batch_size = 1;
  c->num_episode_threads = 1;
  c->num_arena_games = 40;
  c->num_iters = 1;
  c->num_eps = 1;
  c->num_sims = 25;
  c->num_sim_threads = 1;
  c->max_depth = 1000;
  c->cpuct = 1;
  c->quirks = AZB_PROFILE_SANE;
  c->seed = 1;
  c->evaluator = AZB_EVAL_UNIFORM;
  c->device = 0;
  c->max_concurrent_games = 0;
}

// ---------------------------------------------------------------------------------------------
// connect-four batch calls
// ---------------------------------------------------------------------------------------------
static int c4_batch(int op, const azb_c4_state* in, const int8_t* player, const uint8_t* action,
                    const float* pi, uint32_t quirks, size_t n, azb_c4_state* out_states,
                    size_t out_states_per, int8_t* out_i8, uint8_t* out_u8, size_t out_u8_per,
                    float* out_f32, size_t out_f32_per) {
  if (n == 0) return AZB_OK;
  if (!in) return fail(AZB_ERR_INVALID, "input states are NULL");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  DevBuf d_in, d_player, d_action, d_pi, d_os, d_oi8, d_ou8, d_of32, d_bad;
  AZB_CUDA(d_in.ensure(43 * n));
  AZB_CUDA(cudaMemcpy(d_in.p, in, 43 * n, cudaMemcpyHostToDevice));
  if (player) { AZB_CUDA(d_player.ensure(n)); AZB_CUDA(cudaMemcpy(d_player.p, player, n, cudaMemcpyHostToDevice)); }
  if (action) { AZB_CUDA(d_action.ensure(n)); AZB_CUDA(cudaMemcpy(d_action.p, action, n, cudaMemcpyHostToDevice)); }
  if (pi) { AZB_CUDA(d_pi.ensure(28 * n)); AZB_CUDA(cudaMemcpy(d_pi.p, pi, 28 * n, cudaMemcpyHostToDevice)); }
  if (out_states) AZB_CUDA(d_os.ensure(43 * n * out_states_per));
  if (out_i8) AZB_CUDA(d_oi8.ensure(n));
  if (out_u8) AZB_CUDA(d_ou8.ensure(n * out_u8_per));
  if (out_f32) AZB_CUDA(d_of32.ensure(4 * n * out_f32_per));
  AZB_CUDA(d_bad.ensure(4));
  AZB_CUDA(cudaMemset(d_bad.p, 0, 4));
  const int threads = 128;
  const unsigned grid = static_cast<unsigned>((n + threads - 1) / threads);
  k_c4_batch<<<grid, threads>>>(op, d_in.as<int8_t>(), d_player.as<int8_t>(), d_action.as<uint8_t>(),
                                d_pi.as<float>(), quirks, n, d_os.as<int8_t>(), d_oi8.as<int8_t>(),
                                d_ou8.as<uint8_t>(), d_of32.as<float>(), d_bad.as<int>());
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaDeviceSynchronize());
  if (out_states) AZB_CUDA(cudaMemcpy(out_states, d_os.p, 43 * n * out_states_per, cudaMemcpyDeviceToHost));
  if (out_i8) AZB_CUDA(cudaMemcpy(out_i8, d_oi8.p, n, cudaMemcpyDeviceToHost));
  if (out_u8) AZB_CUDA(cudaMemcpy(out_u8, d_ou8.p, n * out_u8_per, cudaMemcpyDeviceToHost));
  if (out_f32) AZB_CUDA(cudaMemcpy(out_f32, d_of32.p, 4 * n * out_f32_per, cudaMemcpyDeviceToHost));
  int bad = 0;
  AZB_CUDA(cudaMemcpy(&bad, d_bad.p, 4, cudaMemcpyDeviceToHost));
  if (bad) return fail(AZB_ERR_INVALID, "invalid action or player in batch (reference: index underflow panic)");
  return AZB_OK;
}

int azb_c4_init(azb_c4_state* out, size_t n) {
  if (!out && n) return fail(AZB_ERR_INVALID, "out is NULL");
  for (size_t i = 0; i < n; ++i) {
    std::memset(&out[i], 0, sizeof(azb_c4_state));
    out[i].me = 1;
  }
  return AZB_OK;
}
int azb_c4_feature_shape(size_t out[3]) {
  out[0] = 2; out[1] = 6; out[2] = 7;
  return AZB_OK;
}
int azb_c4_next_state(const azb_c4_state* in, const int8_t* player, const uint8_t* action, size_t n,
                      azb_c4_state* out, int8_t* next_player) {
  if (n && (!player || !action || !out || !next_player)) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpNext, in, player, action, nullptr, 0, n, out, 1, next_player, nullptr, 0, nullptr, 0);
}
int azb_c4_valid_moves(const azb_c4_state* in, size_t n, uint8_t* out) {
  if (n && !out) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpValid, in, nullptr, nullptr, nullptr, 0, n, nullptr, 0, nullptr, out, 7, nullptr, 0);
}
int azb_c4_game_ended(const azb_c4_state* in, const int8_t* player, size_t n, uint32_t quirks, float* out) {
  if (n && (!player || !out)) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpEnded, in, player, nullptr, nullptr, quirks, n, nullptr, 0, nullptr, nullptr, 0, out, 1);
}
int azb_c4_canonical_form(const azb_c4_state* in, const int8_t* player, size_t n, azb_c4_state* out) {
  if (n && (!player || !out)) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpCanonical, in, player, nullptr, nullptr, 0, n, out, 1, nullptr, nullptr, 0, nullptr, 0);
}
int azb_c4_symmetries(const azb_c4_state* in, const float* pi, size_t n, azb_c4_state* out_states,
                      float* out_pi) {
  if (n && (!pi || !out_states || !out_pi)) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpSymmetries, in, nullptr, nullptr, pi, 0, n, out_states, 2, nullptr, nullptr, 0, out_pi, 14);
}
int azb_c4_eval_heuristic(const azb_c4_state* in, size_t n, float* out) {
  if (n && (!in || !out)) return fail(AZB_ERR_INVALID, "NULL argument");
  for (size_t i = 0; i < n; ++i) out[i] = 0.0f;  // connect_four_game.rs:214-216
  return AZB_OK;
}
int azb_c4_to_features(const azb_c4_state* in, size_t n, float* out) {
  if (n && !out) return fail(AZB_ERR_INVALID, "NULL argument");
  return c4_batch(kOpFeatures, in, nullptr, nullptr, nullptr, 0, n, nullptr, 0, nullptr, nullptr, 0, out, 84);
}

int azb_selftest_arith(uint64_t mismatches[4]) {
  if (!mismatches) return fail(AZB_ERR_INVALID, "NULL argument");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  DevBuf d;
  AZB_CUDA(d.ensure(32));
  AZB_CUDA(cudaMemset(d.p, 0, 32));
  k_selftest_arith<<<256, 256>>>(d.as<unsigned long long>());
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(mismatches, d.p, 32, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

// ---------------------------------------------------------------------------------------------
// AsyncMcts hooks
// ---------------------------------------------------------------------------------------------
int azb_mcts_create(const azb_config* cfg, uint64_t n_trees, azb_mcts** out) {
  if (!out) return fail(AZB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  if (n_trees == 0 || n_trees > (1u << 24)) return fail(AZB_ERR_INVALID, "n_trees out of range");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  AZB_CUDA(cudaSetDevice(cfg->device));
  auto m = std::make_unique<azb_mcts>();
  m->cfg = *cfg;
  rc = m->pool.alloc(make_params(*cfg, 64), static_cast<uint32_t>(n_trees));
  if (rc) return rc;
  rc = m->pool.reset();
  if (rc) return rc;
  *out = m.release();
  return AZB_OK;
}
int azb_mcts_destroy(azb_mcts* m) {
  delete m;
  return AZB_OK;
}

static int upload_states(azb_mcts* m, const azb_c4_state* states) {
  std::vector<BB> h(m->pool.n_trees);
  for (uint32_t i = 0; i < m->pool.n_trees; ++i) h[i] = to_bb(states[i]);
  AZB_CUDA(m->d_states.ensure(h.size() * sizeof(BB)));
  AZB_CUDA(cudaMemcpy(m->d_states.p, h.data(), h.size() * sizeof(BB), cudaMemcpyHostToDevice));
  return AZB_OK;
}

static int check_tree_errors(const TreePool& pool) {
  std::vector<TreeRec> recs(pool.n_trees);
  AZB_CUDA(cudaMemcpy(recs.data(), pool.recs.p, recs.size() * sizeof(TreeRec), cudaMemcpyDeviceToHost));
  for (const auto& r : recs)
    if (r.error) return capacity_error(r.error);
  return AZB_OK;
}

int azb_mcts_get_action_prob(azb_mcts* m, const azb_c4_state* states, float temp, uint16_t* counts,
                             float* pi) {
  if (!m || !states || !counts || !pi) return fail(AZB_ERR_INVALID, "NULL argument");
  AZB_CUDA(cudaSetDevice(m->cfg.device));
  int rc = upload_states(m, states);
  if (rc) return rc;
  const uint32_t n = m->pool.n_trees;
  AZB_CUDA(m->d_counts.ensure(n * 7 * sizeof(uint16_t)));
  AZB_CUDA(m->d_pi.ensure(n * 7 * sizeof(float)));
  AZB_CUDA(cudaMemset(m->d_counts.p, 0, n * 7 * sizeof(uint16_t)));
  AZB_CUDA(cudaMemset(m->d_pi.p, 0, n * 7 * sizeof(float)));
  const unsigned grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  k_mcts_search<<<grid, kWarpsPerCta * 32>>>(m->cfg.evaluator, m->pool.p, m->pool.pools, m->d_states.as<BB>(), temp,
                                             m->d_counts.as<uint16_t>(), m->d_pi.as<float>(), n);
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaDeviceSynchronize());
  AZB_CUDA(cudaMemcpy(counts, m->d_counts.p, n * 7 * sizeof(uint16_t), cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(pi, m->d_pi.p, n * 7 * sizeof(float), cudaMemcpyDeviceToHost));
  return check_tree_errors(m->pool);
}

int azb_mcts_counter_of(azb_mcts* m, const azb_c4_state* states, uint64_t* counters) {
  if (!m || !states || !counters) return fail(AZB_ERR_INVALID, "NULL argument");
  AZB_CUDA(cudaSetDevice(m->cfg.device));
  int rc = upload_states(m, states);
  if (rc) return rc;
  const uint32_t n = m->pool.n_trees;
  AZB_CUDA(m->d_u64.ensure(n * 8));
  const unsigned grid = (n + kWarpsPerCta - 1) / kWarpsPerCta;
  k_counter_of<<<grid, kWarpsPerCta * 32>>>(m->pool.p, m->pool.pools, m->d_states.as<BB>(), m->d_u64.as<uint64_t>(), n);
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(counters, m->d_u64.p, n * 8, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

int azb_mcts_stats(azb_mcts* m, uint64_t* stats) {
  if (!m || !stats) return fail(AZB_ERR_INVALID, "NULL argument");
  AZB_CUDA(cudaSetDevice(m->cfg.device));
  std::vector<TreeRec> recs(m->pool.n_trees);
  AZB_CUDA(cudaMemcpy(recs.data(), m->pool.recs.p, recs.size() * sizeof(TreeRec), cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < recs.size(); ++i) {
    for (int k = 0; k < 6; ++k) stats[i * 8 + k] = recs[i].stat[k];
    stats[i * 8 + 6] = recs[i].n_blocks;
    stats[i * 8 + 7] = recs[i].n_owners;
  }
  return AZB_OK;
}

int azb_mcts_dump(azb_mcts* m, uint64_t tree, uint64_t cap, uint64_t* keys, uint64_t* counters, float* e,
                  float* p7, uint8_t* has_p, uint64_t* n_rows) {
  if (!m || !n_rows) return fail(AZB_ERR_INVALID, "NULL argument");
  if (tree >= m->pool.n_trees) return fail(AZB_ERR_INVALID, "tree index out of range");
  AZB_CUDA(cudaSetDevice(m->cfg.device));
  DevBuf dk, dc, de, dp, dh, dn;
  AZB_CUDA(dk.ensure(std::max<uint64_t>(cap, 1) * 8));
  AZB_CUDA(dc.ensure(std::max<uint64_t>(cap, 1) * 8));
  AZB_CUDA(de.ensure(std::max<uint64_t>(cap, 1) * 4));
  AZB_CUDA(dp.ensure(std::max<uint64_t>(cap, 1) * 28));
  AZB_CUDA(dh.ensure(std::max<uint64_t>(cap, 1)));
  AZB_CUDA(dn.ensure(8));
  AZB_CUDA(cudaMemset(dn.p, 0, 8));
  k_dump_tree<<<64, 256>>>(m->pool.p, m->pool.pools, static_cast<uint32_t>(tree), cap, dk.as<uint64_t>(),
                           dc.as<uint64_t>(), de.as<float>(), dp.as<float>(), dh.as<uint8_t>(),
                           dn.as<unsigned long long>());
  AZB_CUDA(cudaGetLastError());
  unsigned long long rows = 0;
  AZB_CUDA(cudaMemcpy(&rows, dn.p, 8, cudaMemcpyDeviceToHost));
  const uint64_t w = std::min<uint64_t>(rows, cap);
  if (w) {
    if (keys) AZB_CUDA(cudaMemcpy(keys, dk.p, w * 8, cudaMemcpyDeviceToHost));
    if (counters) AZB_CUDA(cudaMemcpy(counters, dc.p, w * 8, cudaMemcpyDeviceToHost));
    if (e) AZB_CUDA(cudaMemcpy(e, de.p, w * 4, cudaMemcpyDeviceToHost));
    if (p7) AZB_CUDA(cudaMemcpy(p7, dp.p, w * 28, cudaMemcpyDeviceToHost));
    if (has_p) AZB_CUDA(cudaMemcpy(has_p, dh.p, w, cudaMemcpyDeviceToHost));
  }
  *n_rows = rows;
  return AZB_OK;
}

// ---------------------------------------------------------------------------------------------
// Coach
// ---------------------------------------------------------------------------------------------
int azb_coach_setup(const azb_config* cfg, azb_coach** out) {
  if (!out) return fail(AZB_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int rc = validate(cfg);
  if (rc) return rc;
  if (cfg->num_sims < 2) return fail(AZB_ERR_INVALID, "self-play needs num_sims >= 2 (sim #1 only evaluates the root)");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  AZB_CUDA(cudaSetDevice(cfg->device));
  auto c = std::make_unique<azb_coach>();
  c->cfg = *cfg;
  *out = c.release();
  return AZB_OK;
}
int azb_coach_destroy(azb_coach* c) {
  delete c;
  return AZB_OK;
}

int azb_coach_self_play(azb_coach* c, uint64_t n_games, uint64_t first_game_id, azb_selfplay_stats* stats) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (n_games == 0 || n_games > (1u << 26)) return fail(AZB_ERR_INVALID, "n_games out of range");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  const SearchParams p = make_params(c->cfg, kMaxPlies);
  // how many trees live in HBM at once: all games, capped by co-resident warps, memory and config
  uint32_t resident = 0;
  int rc = resident_trees(c->cfg.device, &resident);
  if (rc) return rc;
  size_t free_b = 0, total_b = 0;
  AZB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  uint64_t by_mem = (static_cast<uint64_t>(free_b) + c->pool.blocks.bytes + c->pool.tables.bytes) * 8 / 10 / tree_bytes(p);
  uint64_t n_trees = std::min<uint64_t>({n_games, resident, by_mem});
  if (c->cfg.max_concurrent_games) n_trees = std::min<uint64_t>(n_trees, c->cfg.max_concurrent_games);
  if (n_trees == 0) return fail(AZB_ERR_CAPACITY, "not enough device memory for one tree");
  if (!c->pool_ready || c->pool.n_trees != n_trees || std::memcmp(&c->pool.p, &p, sizeof(p)) != 0) {
    rc = c->pool.alloc(p, static_cast<uint32_t>(n_trees));
    if (rc) return rc;
    c->pool_ready = true;
  }
  const uint64_t G = n_games;
  AZB_CUDA(c->plies.ensure(G * 4));
  AZB_CUDA(c->final_r.ensure(G * 4));
  AZB_CUDA(c->final_player.ensure(G));
  AZB_CUDA(c->error.ensure(G * 4));
  AZB_CUDA(c->actions.ensure(G * kTraceStride));
  AZB_CUDA(c->counts.ensure(G * kTraceStride * 7 * 2));
  AZB_CUDA(c->sample_state.ensure(G * kMaxPlies * 16));
  AZB_CUDA(c->sample_pi.ensure(G * kMaxPlies * 32));
  AZB_CUDA(c->stats.ensure(G * 32));
  AZB_CUDA(c->next_game.ensure(4));
  AZB_CUDA(cudaMemset(c->actions.p, 0xFF, G * kTraceStride));
  AZB_CUDA(cudaMemset(c->counts.p, 0, G * kTraceStride * 7 * 2));
  AZB_CUDA(cudaMemset(c->plies.p, 0, G * 4));
  AZB_CUDA(cudaMemset(c->next_game.p, 0, 4));
  GameBufs g{};
  g.plies = c->plies.as<uint32_t>();
  g.final_r = c->final_r.as<float>();
  g.final_player = c->final_player.as<int8_t>();
  g.error = c->error.as<uint32_t>();
  g.actions = c->actions.as<uint8_t>();
  g.counts = c->counts.as<uint16_t>();
  g.sample_state = c->sample_state.as<uint4>();
  g.sample_pi = c->sample_pi.as<float>();
  g.stats = c->stats.as<uint32_t>();
  c->g = g;
  c->n_games = 0;
  c->n_samples = 0;

  cudaEvent_t e0, e1;
  AZB_CUDA(cudaEventCreate(&e0));
  AZB_CUDA(cudaEventCreate(&e1));
  const unsigned grid = static_cast<unsigned>((n_trees + kWarpsPerCta - 1) / kWarpsPerCta);
  AZB_CUDA(cudaEventRecord(e0));
  k_selfplay<<<grid, kWarpsPerCta * 32>>>(c->cfg.evaluator, p, c->pool.pools, g, static_cast<uint32_t>(n_trees),
                                          static_cast<uint32_t>(G), first_game_id, c->next_game.as<unsigned int>());
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaEventRecord(e1));
  AZB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.0f;
  AZB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);

  c->h_plies.resize(G);
  std::vector<uint32_t> h_err(G), h_stats(G * 8);
  AZB_CUDA(cudaMemcpy(c->h_plies.data(), c->plies.p, G * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_err.data(), c->error.p, G * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_stats.data(), c->stats.p, G * 32, cudaMemcpyDeviceToHost));
  azb_selfplay_stats s{};
  for (uint64_t i = 0; i < G; ++i) {
    if (h_err[i]) return capacity_error(h_err[i]);
    s.plies += c->h_plies[i];
    s.sims += h_stats[i * 8 + 0];
    s.levels += h_stats[i * 8 + 1];
    s.expansions += h_stats[i * 8 + 2];
    s.terminal_hits += h_stats[i * 8 + 3];
    s.dup_links += h_stats[i * 8 + 4];
    s.evals += h_stats[i * 8 + 5];
    s.blocks_used_max = std::max<uint64_t>(s.blocks_used_max, h_stats[i * 8 + 6]);
    s.owners_max = std::max<uint64_t>(s.owners_max, h_stats[i * 8 + 7]);
  }
  s.games = G;
  s.samples = s.plies * 2;
  s.device_ms = ms;
  c->n_games = G;
  c->n_samples = s.samples;
  if (stats) *stats = s;
  return AZB_OK;
}

int azb_coach_traces(azb_coach* c, uint8_t* actions, uint16_t* root_counts, uint32_t* plies, float* final_r,
                     int8_t* final_player) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (c->n_games == 0) return fail(AZB_ERR_INVALID, "no self-play results");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  const uint64_t G = c->n_games;
  if (actions) AZB_CUDA(cudaMemcpy(actions, c->actions.p, G * kTraceStride, cudaMemcpyDeviceToHost));
  if (root_counts) AZB_CUDA(cudaMemcpy(root_counts, c->counts.p, G * kTraceStride * 14, cudaMemcpyDeviceToHost));
  if (plies) AZB_CUDA(cudaMemcpy(plies, c->plies.p, G * 4, cudaMemcpyDeviceToHost));
  if (final_r) AZB_CUDA(cudaMemcpy(final_r, c->final_r.p, G * 4, cudaMemcpyDeviceToHost));
  if (final_player) AZB_CUDA(cudaMemcpy(final_player, c->final_player.p, G, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

int azb_coach_num_samples(azb_coach* c, uint64_t* n) {
  if (!c || !n) return fail(AZB_ERR_INVALID, "NULL argument");
  *n = c->n_samples;
  return AZB_OK;
}

int azb_coach_export_samples(azb_coach* c, float* boards, float* pis, float* vs, uint64_t capacity,
                             uint64_t* n_written) {
  if (!c || !boards || !pis || !vs) return fail(AZB_ERR_INVALID, "NULL argument");
  if (c->n_games == 0) return fail(AZB_ERR_INVALID, "no self-play results");
  if (capacity < c->n_samples) return fail(AZB_ERR_CAPACITY, "sample buffer too small");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  const uint64_t G = c->n_games, N = c->n_samples;
  std::vector<uint64_t> off(G);
  uint64_t acc = 0;
  for (uint64_t i = 0; i < G; ++i) {
    off[i] = acc;
    acc += c->h_plies[i];
  }
  AZB_CUDA(c->offsets.ensure(G * 8));
  AZB_CUDA(cudaMemcpy(c->offsets.p, off.data(), G * 8, cudaMemcpyHostToDevice));
  AZB_CUDA(c->out_boards.ensure(N * 84 * 4));
  AZB_CUDA(c->out_pis.ensure(N * 7 * 4));
  AZB_CUDA(c->out_vs.ensure(N * 4));
  k_export_samples<<<static_cast<unsigned>(G), 128>>>(c->g, c->offsets.as<uint64_t>(), c->cfg.quirks,
                                                      c->out_boards.as<float>(), c->out_pis.as<float>(),
                                                      c->out_vs.as<float>(), N);
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(boards, c->out_boards.p, N * 84 * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(pis, c->out_pis.p, N * 7 * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(vs, c->out_vs.p, N * 4, cudaMemcpyDeviceToHost));
  if (n_written) *n_written = N;
  return AZB_OK;
}

}  // extern "C"

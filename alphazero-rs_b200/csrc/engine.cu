// Host side of libazb200.so: device memory, launches, and the C ABI of include/azb200.h.
// The reference's host layer is Rust (src/coach.rs, src/arena.rs); no Rust toolchain exists
// in the build image, so this is C++ behind the same entry points.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <memory>
#include <deque>
#include <string>
#include <thread>
#include <vector>

#include "../../include/azb200.h"
#include "kernels.cuh"
#include "nnet.cuh"
#include "nnet_tc.cuh"
#include "nnet_train.cuh"
#include "rounds.cuh"

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

// AZB200_TIMING=1: host-side section times of the public calls on stderr (diagnostic)
struct HostTimer {
  bool on = std::getenv("AZB200_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  const char* what;
  explicit HostTimer(const char* w) : what(w) {}
  void lap(const char* section) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "[azb200 %s] %-24s %8.3f ms\n", what, section, std::chrono::duration<double, std::milli>(t1 - t0).count());
    t0 = t1;
  }
};

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

// A pair of CUDA events that cannot leak on an early return.
struct EventPair {
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  cudaError_t create() {
    cudaError_t e = cudaEventCreate(&e0);
    return e != cudaSuccess ? e : cudaEventCreate(&e1);
  }
  ~EventPair() {
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
  }
};

uint32_t pow2_ceil(uint64_t x) {
  uint64_t p = 1;
  while (p < x) p <<= 1;
  return static_cast<uint32_t>(p);
}

int validate(const azb_config* cfg) {
  if (!cfg) return fail(AZB_ERR_INVALID, "config is NULL");
  if (cfg->num_sim_threads < 1 || cfg->num_sim_threads > static_cast<uint64_t>(kMaxWave))
    return fail(AZB_ERR_UNSUPPORTED, "num_sim_threads must be 1 (deterministic mode) .. 8 (waves of K simulations per tree)");
  if (cfg->num_sims == 0 || cfg->num_sims > 60000) return fail(AZB_ERR_INVALID, "num_sims out of range");
  if (cfg->num_sims % cfg->num_sim_threads != 0)
    return fail(AZB_ERR_INVALID, "num_sims must be a multiple of num_sim_threads (async_mcts.rs:192 assert)");
  if (cfg->evaluator < AZB_EVAL_UNIFORM || cfg->evaluator > AZB_EVAL_NNET)
    return fail(AZB_ERR_INVALID, "evaluator must be AZB_EVAL_UNIFORM, AZB_EVAL_HASH or AZB_EVAL_NNET");
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
  p.num_threads = static_cast<uint32_t>(c.num_sim_threads);
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
    // a new geometry: every table empty, every generation 0 (from here on a new game bumps its tree's generation instead
    // of zero-filling the table)
    return reset();
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

// The search lives on L1 hits of the hot tree top: keep the unified L1/shared array as L1 and ask for
// just enough shared memory for the CTAs the launch bounds allow (static smem + 1 KB reserved each).
template <typename K>
static cudaError_t search_carveout(K kernel, int ctas_per_sm = kCtasPerSm) {
  cudaFuncAttributes fa{};
  cudaError_t e = cudaFuncGetAttributes(&fa, kernel);
  if (e != cudaSuccess) return e;
  const size_t want = ctas_per_sm * (fa.sharedSizeBytes + 1024);
  const int pct = static_cast<int>((want * 100 + 228 * 1024 - 1) / (228 * 1024));
  return cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
}

static int resident_trees(int device, uint32_t* out) {
  int per_sm = 0, sms = 0;
  AZB_CUDA(search_carveout(k_selfplay<false>, kPlayCtasPerSm));
  AZB_CUDA(search_carveout(k_selfplay<true>, kPlayCtasPerSm));
  AZB_CUDA(search_carveout(k_round<false>));
  AZB_CUDA(search_carveout(k_round<true>));
  AZB_CUDA(search_carveout(k_mcts_search<false>));
  AZB_CUDA(search_carveout(k_mcts_search<true>));
  AZB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_selfplay<false>, kWarpsPerCta * 32, 0));
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

// ---- network handle (NNet::new / predict, src/nnet.rs:35-45) -----------------------------------
struct azb_nnet {
  azb_nnet_config cfg;
  NetLayout L;
  std::vector<float> h_params;
  DevBuf d_params;
  DevBuf d_wtiles;                     // bf16 path: kTcWeightCopies x [2R][18] pre-swizzled 16-KB weight tiles
  DevBuf d_wtiles_bwd;                 // the same tiles tap-mirrored and transposed (ci <-> co): backward data = forward kernel
  size_t wtile_copy_bytes = 0;
  DevBuf d_act[3];                     // bf16 path: activation ping-pong [max_batch*42][128]
  DevBuf d_stem_tab;                   // bf16 path: the stem's 3 x 64 x 128 partial-sum table (k_stem_bf16)
  HeadConvW head_w{};                  // bf16 path: 1x1 head convolutions as a kernel parameter (constant bank)
  CUtensorMap act_map[3];              // the same buffers as 4-D tensors [pos][6][7][128] for TMA im2col loads
  void* act_map_ptr[3] = {nullptr, nullptr, nullptr};
  size_t act_map_bytes[3] = {0, 0, 0};
  DevBuf d_feat, d_states, d_pi, d_v;  // azb_nnet_predict staging
  // training step (azb_nnet_train_*): every layer's output is kept for the backward pass
  std::vector<DevBuf> tr_act;          // [2R + 1] padded bf16: stem output, then every tower convolution's output
  DevBuf tr_g[3];                      // activation gradients (padded bf16), rotated through the backward pass
  std::vector<CUtensorMap> tr_act_map; // TMA descriptors of tr_act / tr_g (re-encoded when a buffer moves)
  CUtensorMap tr_g_map[3];
  size_t tr_bytes = 0;
  uint32_t tr_batch = 0;
  DevBuf d_grad, d_adam_m, d_adam_v, d_loss, d_pis, d_vs;
  uint64_t adam_t = 0;
  bool grads_ready = false;
  bool host_stale = false;             // a training step moved d_params; h_params is refreshed on demand (sync_host)
  int sync_host() {
    if (!host_stale) return AZB_OK;
    AZB_CUDA(cudaSetDevice(cfg.device));
    AZB_CUDA(cudaMemcpy(h_params.data(), d_params.p, L.total * 4, cudaMemcpyDeviceToHost));
    host_stale = false;
    return AZB_OK;
  }
  static uint16_t bf16_rne(float f) {
    uint32_t u;
    std::memcpy(&u, &f, 4);
    if ((u & 0x7F800000u) == 0x7F800000u) return static_cast<uint16_t>(u >> 16);  // inf / nan
    u += 0x7FFFu + ((u >> 16) & 1u);
    return static_cast<uint16_t>(u >> 16);
  }
  int upload() {
    AZB_CUDA(d_params.ensure(L.total * 4));
    AZB_CUDA(cudaMemcpy(d_params.p, h_params.data(), L.total * 4, cudaMemcpyHostToDevice));
    if (cfg.precision == AZB_NNET_BF16_TC) {
      // B operand tiles: tile(layer, kb = tap*2 + half)[n = out channel][k = in channel - 64*half], K-major,
      // SWIZZLE_128B: byte = (n/8)*1024 + (n%8)*128 + (((k/8) ^ (n%8)) * 16) + (k%8)*2
      const size_t n_tiles = static_cast<size_t>(2 * L.R) * kTcKBlocks;
      std::vector<uint16_t> tiles(n_tiles * (kTcTileBytes / 2));
      for (int layer = 0; layer < 2 * L.R; ++layer)
        for (int kb = 0; kb < kTcKBlocks; ++kb) {
          uint16_t* t = tiles.data() + (static_cast<size_t>(layer) * kTcKBlocks + kb) * (kTcTileBytes / 2);
          const int tap = kb >> 1, half = kb & 1;
          const float* w = h_params.data() + L.tower_w + (static_cast<size_t>(layer) * 9 + tap) * kNetC * kNetC;
          for (int n = 0; n < 128; ++n)
            for (int k = 0; k < 64; ++k) {
              const size_t byte = (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
              t[byte / 2] = bf16_rne(w[static_cast<size_t>(half * 64 + k) * kNetC + n]);
            }
        }
      {  // backward-data tiles: dX[r][ci] = sum_tap' sum_co dZ[r + s(tap')][co] W[8 - tap'][ci][co]: n = ci, k = co
        std::vector<uint16_t> bt(n_tiles * (kTcTileBytes / 2));
        for (int layer = 0; layer < 2 * L.R; ++layer)
          for (int kb = 0; kb < kTcKBlocks; ++kb) {
            uint16_t* t = bt.data() + (static_cast<size_t>(layer) * kTcKBlocks + kb) * (kTcTileBytes / 2);
            const int tap = kb >> 1, half = kb & 1;
            const float* w = h_params.data() + L.tower_w + (static_cast<size_t>(layer) * 9 + (8 - tap)) * kNetC * kNetC;
            for (int n = 0; n < 128; ++n)
              for (int k = 0; k < 64; ++k) {
                const size_t byte = (n / 8) * 1024 + (n % 8) * 128 + (((k / 8) ^ (n % 8)) * 16) + (k % 8) * 2;
                t[byte / 2] = bf16_rne(w[static_cast<size_t>(n) * kNetC + half * 64 + k]);
              }
          }
        AZB_CUDA(d_wtiles_bwd.ensure(bt.size() * 2));
        AZB_CUDA(cudaMemcpy(d_wtiles_bwd.p, bt.data(), bt.size() * 2, cudaMemcpyHostToDevice));
      }
      // kTcWeightCopies replicas at different addresses: every CTA streams the same tile sequence at
      // about the same time; spreading the CTAs over replicas spreads that traffic over the L2 slices
      wtile_copy_bytes = tiles.size() * 2;
      AZB_CUDA(d_wtiles.ensure(wtile_copy_bytes * kTcWeightCopies));
      for (int c = 0; c < kTcWeightCopies; ++c)
        AZB_CUDA(cudaMemcpy(d_wtiles.as<uint8_t>() + c * wtile_copy_bytes, tiles.data(), wtile_copy_bytes, cudaMemcpyHostToDevice));
      AZB_CUDA(cudaFuncSetAttribute(k_conv3x3_tc<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemBytes));
      AZB_CUDA(cudaFuncSetAttribute(k_conv3x3_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, kT3SmemBytes));
      AZB_CUDA(cudaFuncSetAttribute(k_tower_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, kTwSmemBytes));
      AZB_CUDA(cudaFuncSetAttribute(k_conv3x3_wgrad, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemBytes));
      AZB_CUDA(cudaFuncSetAttribute(k_stem_bf16, cudaFuncAttributeMaxDynamicSharedMemorySize, kStemSmemBytes));
      {
        std::vector<float> tab(3 * 64 * kNetC);
        stem_table_build(h_params.data(), L, tab.data());
        AZB_CUDA(d_stem_tab.ensure(kStemSmemBytes));
        AZB_CUDA(cudaMemcpy(d_stem_tab.p, tab.data(), kStemSmemBytes, cudaMemcpyHostToDevice));
      }
      for (int ci = 0; ci < kNetC; ++ci) {
        head_w.w[ci][0] = h_params[L.pol_w + ci * 2 + 0];
        head_w.w[ci][1] = h_params[L.pol_w + ci * 2 + 1];
        head_w.w[ci][2] = h_params[L.val_w + ci];
      }
      head_w.b[0] = h_params[L.pol_b + 0];
      head_w.b[1] = h_params[L.pol_b + 1];
      head_w.b[2] = h_params[L.val_b];
    }
    return AZB_OK;
  }
};

namespace {
// TMA descriptor of a PADDED activation buffer [rows][128] for k_conv3x3_tc3: plain 2-D tiles of 64 channels x
// 160 rows (a 128-row tile and its halo), SWIZZLE_128B, out-of-range rows zero-filled.
int encode_act_map_rows(CUtensorMap* map, void* ptr, size_t bytes) {
  typedef CUresult (*EncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  static EncodeTiled encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    AZB_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(AZB_ERR_CUDA, "cuTensorMapEncodeTiled is not available in this driver");
    encode = reinterpret_cast<EncodeTiled>(fn);
  }
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(kNetC), bytes / (kNetC * 2)};
  const cuuint64_t strides[1] = {kNetC * 2};
  const cuuint32_t box[2] = {kTcBlockK, static_cast<cuuint32_t>(kT3StageRows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(AZB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return AZB_OK;
}

// Network rounds: a slot that needs no evaluation for this many simulations yields (endgames that only hit
// terminal nodes would otherwise hold the whole round back, every simulation of a round starting from cold
// caches).  Measured on config 3 (8192 games x 400 sims): 1: 10.1 s, 2: 9.1, 3: 8.8, 4-6: 8.7, 8: 8.8, 16: 9.3,
// 32: 9.8; with the leaf de-duplication and the evaluation cache (a cache hit does not end a slot's round): 3: 4.1 s,
// 5: 3.8, 8: 3.7, 12: 3.8, 20: 4.1, 32: 4.5.  AZB200_ROUND_SIMS overrides for sweeps; results do not depend on it.
uint32_t round_sim_budget(bool arena = false) {
  static const uint32_t n = std::getenv("AZB200_ROUND_SIMS") ? static_cast<uint32_t>(std::max(1, std::atoi(std::getenv("AZB200_ROUND_SIMS")))) : 0u;
  return n ? n : (arena ? 5u : 8u);  // (the arena share of config 4 with the cache: 2.49 s with 5, 2.57 with 16, 2.70 with 32)
}

// Slots of the lock-step rounds (network evaluator).  The persistent kernel of the fused evaluators holds one game per
// co-resident warp; a round is a kernel per phase, so its slot count is free, and more slots are larger, fewer rounds: the
// forward pass works on ~1.8 k instead of ~1 k positions and every launch's fixed cost is paid less often.  Config 3 (8192
// games x 400 sims): 2960 slots 3.27 s, 4144 3.07, 4736 (the co-resident warps) 3.08-3.10, 6144 2.97, 8192 2.79-2.83.  The
// games are the same games whatever the slot count (tests/test_selfplay_gpu.py).  AZB200_ROUND_SLOTS overrides.
uint32_t round_slots(uint32_t resident_warps) {
  static const uint32_t n = std::getenv("AZB200_ROUND_SLOTS") ? static_cast<uint32_t>(std::max(1, std::atoi(std::getenv("AZB200_ROUND_SLOTS")))) : 0u;
  return n ? n : std::max<uint32_t>(resident_warps, 8192u);
}

// Programmatic dependent launch between the tower's layers (AZB200_TC_PDL=0 turns it off; the round graph's capture
// also turns it off when the driver refuses programmatic edges inside a capture)
bool g_tc_pdl = !(std::getenv("AZB200_TC_PDL") && std::getenv("AZB200_TC_PDL")[0] == '0');

// Which tensor-core tower runs (AZB200_TC_PAIR): 0 = k_conv3x3_tc<1> (one CTA, cp.async gather, dense layout: the
// independent implementation the bit-equality test compares with), default 3 = k_conv3x3_tc3 (CTA pair, the tile fetched
// once and reused by all taps, padded layout).  (Round 1's intermediate kernels — the TMA-im2col pair kernel and the 4-CTA
// weight multicast — were measured slower and are gone; profiles/r1_nnet_forward.md keeps their numbers.)
int tc_mode() {
  static const int mode = [] {
    const char* e = std::getenv("AZB200_TC_PAIR");
    return e && e[0] == '0' ? 0 : 3;
  }();
  return mode;
}

// The forward pass's 2R convolutions as ONE launch (k_tower_tc3: position-aligned tiles, a CTA pair takes its tiles through
// every layer) instead of 2R launches of k_conv3x3_tc3.  AZB200_TOWER=0 keeps the layer-by-layer launches (the training
// step always uses them); both produce the same bits (tests/test_nnet_gpu.py).
bool g_tower = !(std::getenv("AZB200_TOWER") && std::getenv("AZB200_TOWER")[0] == '0');

// One dense forward pass over the first *d_count (or max_batch) positions.  With a second network of the same shape (the
// arena's two players) the two towers share ONE launch of k_tower_tc3, its CTA pairs split between the models in proportion
// to their positions; everything else (stem, heads, and the towers themselves when the one-launch tower does not apply)
// runs per model.
int nnet_forward(azb_nnet* net, const uint4* d_states, const uint32_t* d_count, uint32_t max_batch, float* d_pi,
                 float* d_v, cudaStream_t st, azb_nnet* net2 = nullptr, const uint4* d_states2 = nullptr,
                 const uint32_t* d_count2 = nullptr, float* d_pi2 = nullptr, float* d_v2 = nullptr) {
  if (max_batch == 0) return AZB_OK;
  if (net2) {
    const bool together = net->cfg.precision == AZB_NNET_BF16_TC && net2->cfg.precision == AZB_NNET_BF16_TC && tc_mode() == 3 &&
                          net->L.R == net2->L.R && g_tower && d_count && d_count2 && max_batch <= 16384u &&
                          !std::getenv("AZB200_TC_DEBUG") && !(std::getenv("AZB200_TOWER_PAIR") && std::getenv("AZB200_TOWER_PAIR")[0] == '0');
    if (!together) {
      const int rc = nnet_forward(net, d_states, d_count, max_batch, d_pi, d_v, st);
      return rc ? rc : nnet_forward(net2, d_states2, d_count2, max_batch, d_pi2, d_v2, st);
    }
  }
  if (net->cfg.precision == AZB_NNET_FP32) {
    const size_t smem = (2 * kCells * kNetC + 256) * sizeof(float);
    const unsigned grid = std::min<uint32_t>(max_batch, 148u * 4u);
    k_nnet_fp32<<<grid, 128, smem, st>>>(net->d_params.as<float>(), net->L, d_states, d_count, max_batch, d_pi, d_v);
    AZB_CUDA(cudaGetLastError());
    return AZB_OK;
  }
  // bf16 tensor-core path: stem -> 2R x conv3x3 (tcgen05) -> heads
  const int mode = tc_mode();
  const ActLayout lay = mode == 3 ? kActPadded : kActDense;
  // + slack: the last 128-row TMA copy of a ragged batch runs a few rows past the batch
  const size_t act_bytes = (static_cast<size_t>(max_batch) + 8) * lay.pos_rows * kNetC * 2;
  const size_t total = static_cast<size_t>(max_batch) * kCells * (kNetC / 8);
  // activation buffers + their TMA descriptors, then the stem into buffer 0
  auto prepare_and_stem = [&](azb_nnet* n, const uint4* states, const uint32_t* count) -> int {
    for (auto& b : n->d_act) AZB_CUDA(b.ensure(act_bytes));
    for (int i = 0; i < 3; ++i)
      if (n->act_map_ptr[i] != n->d_act[i].p || n->act_map_bytes[i] != n->d_act[i].bytes) {
        // a fresh buffer: the padded layout's zero rows / columns are zeroed here once and never written again
        AZB_CUDA(cudaMemsetAsync(n->d_act[i].p, 0, n->d_act[i].bytes, st));
        if (mode == 3) {
          const int rc = encode_act_map_rows(&n->act_map[i], n->d_act[i].p, n->d_act[i].bytes);
          if (rc) return rc;
        }
        n->act_map_ptr[i] = n->d_act[i].p;
        n->act_map_bytes[i] = n->d_act[i].bytes;
      }
    k_stem_bf16<<<static_cast<unsigned>(std::min<size_t>((total + 1023) / 1024, 148u * 2u)), 1024, kStemSmemBytes, st>>>(
        n->d_params.as<float>(), n->L, n->d_stem_tab.as<float>(), states, count, max_batch, n->d_act[0].as<__nv_bfloat16>(), lay);
    return AZB_OK;
  };
  {
    int rc = prepare_and_stem(net, d_states, d_count);
    if (!rc && net2) rc = prepare_and_stem(net2, d_states2, d_count2);
    if (rc) return rc;
  }
  __nv_bfloat16* x = net->d_act[0].as<__nv_bfloat16>();
  __nv_bfloat16* y = net->d_act[1].as<__nv_bfloat16>();
  __nv_bfloat16* z = net->d_act[2].as<__nv_bfloat16>();
  const float* prm = net->d_params.as<float>();
  const uint32_t tiles = (max_batch * kCells + kTcCtaRows - 1) / kTcCtaRows;
  const unsigned grid = std::min<uint32_t>(tiles, 148u);
  // Default: CTA pairs (tcgen05 cta_group::2) with resident weights; AZB200_TC_PAIR=0 selects the
  // single-CTA kernel with streamed weight tiles.
  const bool use_pair = mode != 0;
  const bool use_pdl = g_tc_pdl;
  const uint32_t pair_tiles = (max_batch * lay.pos_rows + kT2PairRows - 1) / kT2PairRows;
  static int max_pairs = -1;  // co-resident CTA pairs (one CTA per SM; a GPC with an odd SM count leaves one out)
  if (max_pairs < 0) {
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(148u);
    qc.blockDim = dim3(kTcThreads);
    qc.dynamicSmemBytes = kT3SmemBytes;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, k_conv3x3_tc3, &qc) != cudaSuccess) { cudaGetLastError(); n = 0; }
    max_pairs = n;
    if (std::getenv("AZB200_TIMING")) std::fprintf(stderr, "[azb200 nnet] co-resident CTA pairs: %d\n", n);
  }
  auto launch_conv = [&](const ConvTcArgs& a) -> cudaError_t {
    if (use_pair && max_pairs > 0) {
      int mi = 0;
      while (mi < 2 && net->d_act[mi].p != a.in) ++mi;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2u * std::min<uint32_t>(pair_tiles, static_cast<uint32_t>(max_pairs)));
      cfg.blockDim = dim3(kTcThreads);
      cfg.dynamicSmemBytes = kT3SmemBytes;
      cfg.stream = st;
      cudaLaunchAttribute pdl{};  // overlap this layer's prologue + weight preload with the previous kernel's tail
      pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
      pdl.val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = &pdl;
      cfg.numAttrs = use_pdl ? 1 : 0;
      return cudaLaunchKernelEx(&cfg, k_conv3x3_tc3, a, net->act_map[mi]);
    }
    k_conv3x3_tc<1><<<grid, kTcThreads, kTcSmemBytes, st>>>(a);
    return cudaGetLastError();
  };
  static unsigned long long* d_dbg = nullptr;  // AZB200_TC_DEBUG=1: role timers of the first conv launch
  if (!d_dbg && std::getenv("AZB200_TC_DEBUG")) {
    AZB_CUDA(cudaMalloc(&d_dbg, 32 * 8));
    AZB_CUDA(cudaMemset(d_dbg, 0, 32 * 8));
  }
  bool tower_done = false;
  // The tower's position-aligned tiles cost 12.5 % more MMA work and save the 2R launch boundaries: measured faster up to
  // ~3.5 k positions per pass (profiles/r2_tower.md: 191 vs 216 us at 1014, 305 vs 325 at 2028, 656 vs 645 at 4096).  A
  // caller-sized batch decides by its size; a device-counted round (the search: ~1 k positions on average, 3.4 k at most
  // with 8192 slots) takes the tower unless the slot count says the rounds will be large.
  const bool tower_size = d_count ? max_batch <= 16384u : max_batch <= 3072u;
  if (use_pair && max_pairs > 0 && g_tower && tower_size && !d_dbg) {
    static int tower_pairs = -1;  // co-resident CTA pairs of the tower kernel (it must fit the device in one wave)
    if (tower_pairs < 0) {
      cudaLaunchConfig_t qc{};
      qc.gridDim = dim3(148u);
      qc.blockDim = dim3(kTwThreads);
      qc.dynamicSmemBytes = kTwSmemBytes;
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, k_tower_tc3, &qc) != cudaSuccess) { cudaGetLastError(); n = 0; }
      tower_pairs = n;
    }
    if (tower_pairs > 0) {
      TowerTcArgs t{};
      auto fill = [&](TowerModel& m, azb_nnet* n, const uint32_t* count) {
        for (int i = 0; i < 3; ++i) m.act[i] = n->d_act[i].as<__nv_bfloat16>();
        m.w_tiles = n->d_wtiles.as<uint8_t>();
        m.bias = n->d_params.as<float>() + n->L.tower_b;
        m.count = count;
      };
      fill(t.m[0], net, d_count);
      if (net2) fill(t.m[1], net2, d_count2);
      t.n_models = net2 ? 2 : 1;
      t.max_batch = max_batch;
      t.n_layers = 2 * net->L.R;
      static unsigned long long* d_tdbg = nullptr;  // AZB200_TOWER_DEBUG=1: per-layer timeline of CTA 0
      if (!d_tdbg && std::getenv("AZB200_TOWER_DEBUG")) {
        AZB_CUDA(cudaMalloc(&d_tdbg, 256 * 8));
        AZB_CUDA(cudaMemset(d_tdbg, 0, 256 * 8));
      }
      t.dbg = d_tdbg;
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(2u * std::min<uint32_t>((net2 ? 2u : 1u) * ((max_batch + kTwTilePos - 1) / kTwTilePos), static_cast<uint32_t>(tower_pairs)));
      cfg.blockDim = dim3(kTwThreads);
      cfg.dynamicSmemBytes = kTwSmemBytes;
      cfg.stream = st;
      cudaLaunchAttribute at[1]{};
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = use_pdl ? 1 : 0;
      azb_nnet* nb = net2 ? net2 : net;
      const cudaError_t e = cudaLaunchKernelEx(&cfg, k_tower_tc3, t, net->act_map[0], net->act_map[1], net->act_map[2], nb->act_map[0],
                                               nb->act_map[1], nb->act_map[2]);
      if (e == cudaSuccess) {
        tower_done = true;
        if (net->L.R & 1) std::swap(x, z);  // where the last block left its output
        if (d_tdbg) {
          static int printed = 0;
          if (printed++ == 4) {
            unsigned long long h[256];
            AZB_CUDA(cudaMemcpy(h, d_tdbg, sizeof(h), cudaMemcpyDeviceToHost));
            std::fprintf(stderr, "[azb200 tower] CTA 0: %llu cycles, %llu tiles per layer; per layer (cycles since start): weights requested | "
                                 "first tile may be fetched | first tile in | first tile issued | last tile issued | last accumulator | last stores issued | last tile published\n", h[0], h[1]);
            for (int l = 0; l < t.n_layers && l < 31; ++l) {
              std::fprintf(stderr, "[azb200 tower] layer %2d:", l);
              for (int k = 0; k < 8; ++k) std::fprintf(stderr, " %8llu", h[8 + l * 8 + k]);
              std::fprintf(stderr, "\n");
            }
          }
        }
      } else {
        cudaGetLastError();
        g_tower = false;  // refused: layer by layer from here on
        if (std::getenv("AZB200_TIMING")) std::fprintf(stderr, "[azb200 nnet] tower launch refused (%s): layer-by-layer kernels\n", cudaGetErrorString(e));
      }
    }
  }
  if (net2 && !tower_done) {  // (the shared launch did not happen: each model on its own, stems again)
    const int rc = nnet_forward(net, d_states, d_count, max_batch, d_pi, d_v, st);
    return rc ? rc : nnet_forward(net2, d_states2, d_count2, max_batch, d_pi2, d_v2, st);
  }
  for (int blk = 0; blk < net->L.R && !tower_done; ++blk) {
    ConvTcArgs a{};
    a.dbg = nullptr;
    a.count = d_count;
    a.max_batch = max_batch;
    a.in = x; a.residual = nullptr; a.out = y;
    a.w_tiles = net->d_wtiles.as<uint8_t>() + static_cast<size_t>(2 * blk) * kTcKBlocks * kTcTileBytes;
    a.w_copy_stride = net->wtile_copy_bytes;
    a.bias = prm + net->L.tower_b + (2 * blk) * kNetC;
    AZB_CUDA(launch_conv(a));
    a.in = y; a.residual = x; a.out = z;
    a.dbg = blk == 0 ? d_dbg : nullptr;  // (the second launch: its predecessor is a convolution, as for 11 of the 12)
    a.w_tiles += static_cast<size_t>(kTcKBlocks) * kTcTileBytes;
    a.bias += kNetC;
    AZB_CUDA(launch_conv(a));
    std::swap(x, z);
  }
  {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(std::min<uint32_t>((max_batch + kHeadPos - 1) / kHeadPos, 148u * 8u));
    cfg.blockDim = dim3(256);
    cfg.stream = st;
    cudaLaunchAttribute pdl{};
    pdl.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    pdl.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &pdl;
    cfg.numAttrs = (use_pdl && use_pair && max_pairs > 0) ? 1 : 0;
    const __nv_bfloat16* xin = x;
    AZB_CUDA(cudaLaunchKernelEx(&cfg, k_heads_bf16, prm, net->L, net->head_w, xin, d_count, max_batch, d_pi, d_v, lay));
    if (net2) {  // (only after a shared tower launch: the result sits where the last block left it)
      const __nv_bfloat16* xin2 = net2->d_act[(net2->L.R & 1) ? 2 : 0].as<__nv_bfloat16>();
      const float* prm2 = net2->d_params.as<float>();
      AZB_CUDA(cudaLaunchKernelEx(&cfg, k_heads_bf16, prm2, net2->L, net2->head_w, xin2, d_count2, max_batch, d_pi2, d_v2, lay));
    }
  }
  AZB_CUDA(cudaGetLastError());
  if (d_dbg) {
    unsigned long long h[32];
    AZB_CUDA(cudaMemcpy(h, d_dbg, sizeof(h), cudaMemcpyDeviceToHost));
    static int printed = 0;
    if (printed++ == 4) {
      const char* names[16] = {"producer wait empty", "producer TMA issue", "epilogue wait acc_full", "epilogue work", "mma wait full",
                               "mma wait acc_empty", "mma fences", "mma issue (4 MMAs)", "mma commit", "mma wait weights",
                               "epilogue done at", "kernel cycles", "tile iterations", "setup done at", "grid dependency met at",
                               "first MMA at"};
      for (int r = 0; r < 2; ++r)
        for (int k = 0; k < 16; ++k)
          if (h[r * 16 + k]) std::fprintf(stderr, "[azb200 tc2 rank %d] %-24s %llu\n", r, names[k], h[r * 16 + k]);
    }
  }
  return AZB_OK;
}

// Per-game outputs of a self-play / arena call.
struct GameStore {
  DevBuf plies, final_r, final_player, error, actions, counts, sample_state, sample_pi, stats, arena_result, ply_ns;
  GameBufs g{};
  int alloc(uint64_t G, bool samples, cudaStream_t st = nullptr) {
    AZB_CUDA(plies.ensure(G * 4));
    AZB_CUDA(final_r.ensure(G * 4));
    AZB_CUDA(final_player.ensure(G));
    AZB_CUDA(error.ensure(G * 4));
    AZB_CUDA(actions.ensure(G * kTraceStride));
    AZB_CUDA(counts.ensure(G * kTraceStride * 7 * 2));
    if (samples) {
      AZB_CUDA(sample_state.ensure(G * kMaxPlies * 16));
      AZB_CUDA(sample_pi.ensure(G * kMaxPlies * 32));
    }
    AZB_CUDA(stats.ensure(G * 32));
    AZB_CUDA(arena_result.ensure(G));
    AZB_CUDA(cudaMemsetAsync(actions.p, 0xFF, G * kTraceStride, st));
    AZB_CUDA(cudaMemsetAsync(counts.p, 0, G * kTraceStride * 7 * 2, st));
    AZB_CUDA(cudaMemsetAsync(plies.p, 0, G * 4, st));
    AZB_CUDA(cudaMemsetAsync(error.p, 0, G * 4, st));
    AZB_CUDA(cudaMemsetAsync(stats.p, 0, G * 32, st));
    AZB_CUDA(cudaMemsetAsync(arena_result.p, 0, G, st));
    g.plies = plies.as<uint32_t>();
    g.final_r = final_r.as<float>();
    g.final_player = final_player.as<int8_t>();
    g.error = error.as<uint32_t>();
    g.actions = actions.as<uint8_t>();
    g.counts = counts.as<uint16_t>();
    g.sample_state = sample_state.as<uint4>();
    g.sample_pi = sample_pi.as<float>();
    g.stats = stats.as<uint32_t>();
    g.ply_ns = nullptr;
    if (std::getenv("AZB200_PLY_TIMES")) {  // diagnostic: per-ply completion times
      AZB_CUDA(ply_ns.ensure(G * kTraceStride * 8));
      AZB_CUDA(cudaMemsetAsync(ply_ns.p, 0, G * kTraceStride * 8, st));
      g.ply_ns = ply_ns.as<unsigned long long>();
    }
    return AZB_OK;
  }
};

// Device state of the lock-step round engine (csrc/rounds.cuh).
struct RoundEngine {
  DevBuf recs, active, ctl_words, leaf_state, leaf_count, leaf_pi, leaf_v, dedup_keys, dedup_idx, times;
  uint32_t n_slots = 0, dedup_mask = 0, cache_mask = 0;
  // The cache's memory belongs to the calling thread (one buffer per thread, re-used by every run of that thread: a coach's
  // self-play and the arena calls of Coach::learn follow each other): a run borrows it and clears the keys first.
  struct CacheBufs {
    DevBuf keys, vals;
    int device = -1;
  };
  static CacheBufs& thread_cache() {
    static thread_local CacheBufs b;
    return b;
  }
  // evaluation cache (rounds.cuh LeafBufs): 2^AZB200_EVAL_CACHE_LOG2 entries per model (default 2^25 = 1.3 GB per model with
  // the values), allocated at the first network run; AZB200_EVAL_CACHE=0 turns it off; without memory for it the run goes on
  // without a cache
  int ensure_cache() {
    static const bool on = !(std::getenv("AZB200_EVAL_CACHE") && std::atoi(std::getenv("AZB200_EVAL_CACHE")) == 0);
    static const int log2n = std::getenv("AZB200_EVAL_CACHE_LOG2") ? std::max(10, std::min(28, std::atoi(std::getenv("AZB200_EVAL_CACHE_LOG2")))) : 25;
    if (!on) { cache_mask = 0; return AZB_OK; }
    const size_t n = static_cast<size_t>(1) << log2n;
    CacheBufs& cb = thread_cache();
    int dev = 0;
    AZB_CUDA(cudaGetDevice(&dev));
    if (cb.device != dev) {
      cb.keys.release();
      cb.vals.release();
      cb.device = dev;
    }
    if (cb.keys.ensure(2 * n * 8) != cudaSuccess || cb.vals.ensure(2 * n * 32) != cudaSuccess) {
      cudaGetLastError();
      cb.keys.release();
      cb.vals.release();
      cache_mask = 0;
      return AZB_OK;
    }
    cache_mask = static_cast<uint32_t>(n - 1);
    AZB_CUDA(cudaMemsetAsync(cb.keys.p, 0, 2 * n * 8));  // a new call: the networks may have changed
    return AZB_OK;
  }
  uint32_t leaf_cap = 0;  // rows of a model's leaf batch per round: slots * num_sim_threads
  int alloc(uint32_t slots, uint32_t leaves_per_slot = 1) {
    n_slots = slots;
    leaf_cap = slots * std::max(1u, leaves_per_slot);
    // leaf de-duplication table (rounds.cuh LeafBufs): 4 entries per slot and model; AZB200_LEAF_DEDUP=0 turns it off
    static const bool dedup = !(std::getenv("AZB200_LEAF_DEDUP") && std::atoi(std::getenv("AZB200_LEAF_DEDUP")) == 0);
    dedup_mask = dedup ? pow2_ceil(static_cast<uint64_t>(leaf_cap) * 4u) - 1u : 0u;
    if (dedup) {
      AZB_CUDA(dedup_keys.ensure(4 * (static_cast<size_t>(dedup_mask) + 1) * 8));  // 2 round parities x 2 models
      AZB_CUDA(dedup_idx.ensure(4 * (static_cast<size_t>(dedup_mask) + 1) * 4));
      AZB_CUDA(cudaMemset(dedup_keys.p, 0, 4 * (static_cast<size_t>(dedup_mask) + 1) * 8));
    }
    AZB_CUDA(recs.ensure(static_cast<size_t>(slots) * sizeof(GameRec)));
    AZB_CUDA(active.ensure(static_cast<size_t>(slots) * 4));
    AZB_CUDA(ctl_words.ensure(64));
    AZB_CUDA(leaf_state.ensure(static_cast<size_t>(leaf_cap) * 2 * 16));
    AZB_CUDA(leaf_count.ensure(8));
    AZB_CUDA(leaf_pi.ensure(static_cast<size_t>(leaf_cap) * 2 * 32));
    AZB_CUDA(leaf_v.ensure(static_cast<size_t>(leaf_cap) * 2 * 4));
    AZB_CUDA(cudaMemset(recs.p, 0, static_cast<size_t>(slots) * sizeof(GameRec)));  // phase = Empty
    AZB_CUDA(cudaMemset(ctl_words.p, 0, 64));
    AZB_CUDA(cudaMemset(leaf_count.p, 0, 8));
    return AZB_OK;
  }
  // One stream per engine: a blocking stream (it orders itself against the legacy default stream the rest of the library
  // uses) — stream capture needs a real stream.  The host-visible progress word lives in page-locked memory.
  cudaStream_t stream = nullptr;
  unsigned long long* h_progress = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  ~RoundEngine() {
    if (graph_exec) cudaGraphExecDestroy(graph_exec);
    if (stream) cudaStreamDestroy(stream);
    if (h_progress) cudaFreeHost(h_progress);
  }
  // Runs every game of the call to completion.  nets[k] evaluates the leaves of player k.
  //
  // A round = k_compact, k_round and one dense forward pass per model (stem, 2R convolutions, heads): ~16 launches of
  // 5-25 us each.  Launched one by one with a blocking copy every 16 rounds they left ~40 % of the device timeline empty
  // (profiles/r1_configs.md: 323 us per round against 186 us of kernels).  Now nothing in a round depends on the host:
  // the round number lives on the device (de-duplication stamp), grids are sized for the slot count and read the live
  // counts from device memory, so an even + odd round pair is captured ONCE into a CUDA graph and replayed; the host
  // follows the run through a page-locked progress word that every k_round writes and stops launching when it reads
  // "no live slot" (it runs at most kRunAhead rounds ahead; rounds after the end find nothing to do and cost a few us).
  // AZB200_GRAPH=0 launches the same kernels one by one.
  int run(const RoundParams& rp, const Pools& pools, GameStore& gs, azb_nnet* nets[2], uint64_t* launches,
          uint64_t* nn_positions = nullptr, uint64_t* cache_hits = nullptr) {
    constexpr uint64_t kRunAhead = 48;
    static const bool use_graph = !(std::getenv("AZB200_GRAPH") && std::atoi(std::getenv("AZB200_GRAPH")) == 0) &&
                                  !std::getenv("AZB200_TC_DEBUG");
    if (!stream) AZB_CUDA(cudaStreamCreate(&stream));
    if (!h_progress) AZB_CUDA(cudaHostAlloc(reinterpret_cast<void**>(&h_progress), 8, cudaHostAllocMapped));
    *reinterpret_cast<volatile unsigned long long*>(h_progress) = 0ull;
    Control ctl{};
    ctl.round = ctl_words.as<unsigned int>() + 8;
    ctl.host_progress = h_progress;
    ctl.next_game = ctl_words.as<unsigned int>();
    ctl.n_active = ctl_words.as<unsigned int>() + 1;
    ctl.active_list = active.as<uint32_t>();
    ctl.arena_result = gs.arena_result.as<int8_t>();
    LeafBufs leaf{};
    leaf.state = leaf_state.as<uint4>();
    leaf.count = leaf_count.as<uint32_t>();
    leaf.pi = leaf_pi.as<float>();
    leaf.v = leaf_v.as<float>();
    leaf.dkeys = dedup_keys.as<unsigned long long>();
    leaf.didx = dedup_idx.as<uint32_t>();
    leaf.dmask = dedup_mask;
    leaf.nn_total = reinterpret_cast<unsigned long long*>(ctl_words.as<uint8_t>() + 16);
    leaf.cache_hits = reinterpret_cast<unsigned long long*>(ctl_words.as<uint8_t>() + 24);
    const bool any_net = rp.ev_kind[0] >= AZB_EVAL_NNET || (rp.mode == kModeArena && rp.ev_kind[1] >= AZB_EVAL_NNET);
    cache_mask = 0;
    if (any_net) {
      const int rcc = ensure_cache();
      if (rcc) return rcc;
    }
    leaf.ckeys = thread_cache().keys.as<unsigned long long>();
    leaf.cvals = thread_cache().vals.as<float>();
    leaf.cmask = cache_mask;
    const unsigned grid = (rp.n_slots + kWarpsPerCta - 1) / kWarpsPerCta;
    constexpr uint32_t kTimesCap = 1u << 16;
    unsigned long long* d_times = nullptr;  // AZB200_ROUND_TIMES=1: per-round phase times on stderr at the end of the run
    if (std::getenv("AZB200_ROUND_TIMES")) {
      AZB_CUDA(times.ensure(static_cast<size_t>(kTimesCap) * 32));
      AZB_CUDA(cudaMemset(times.p, 0, static_cast<size_t>(kTimesCap) * 32));
      d_times = times.as<unsigned long long>();
    }
    uint64_t per_round = 2;
    // one round of parity `par` on `stream` (the only per-round differences: which counter / which de-duplication table)
    auto launch_round = [&](uint32_t par) -> int {
      Control c = ctl;
      LeafBufs lf = leaf;
      lf.dpar = par;
      c.n_active = ctl_words.as<unsigned int>() + 1 + par;  // double-buffered (k_compact)
      c.n_active_next = ctl_words.as<unsigned int>() + 1 + (par ^ 1u);
      if (d_times) k_stamp<<<1, 1, 0, stream>>>(d_times, c.round, 0u, kTimesCap);
      k_compact<<<(rp.n_slots + 255u) / 256u, 256, 0, stream>>>(rp, recs.as<GameRec>(), c, lf);
      if (d_times) k_stamp<<<1, 1, 0, stream>>>(d_times, c.round, 1u, kTimesCap);
      if (rp.p.num_threads > 1u) k_round<true><<<grid, kWarpsPerCta * 32, 0, stream>>>(rp, pools, recs.as<GameRec>(), c, lf, gs.g);
      else k_round<false><<<grid, kWarpsPerCta * 32, 0, stream>>>(rp, pools, recs.as<GameRec>(), c, lf, gs.g);
      if (d_times) k_stamp<<<1, 1, 0, stream>>>(d_times, c.round, 2u, kTimesCap);
      AZB_CUDA(cudaGetLastError());
      if (any_net) {
        const bool use0 = nets[0] && rp.ev_kind[0] >= AZB_EVAL_NNET;
        const bool use1 = nets[1] && rp.ev_kind[1] >= AZB_EVAL_NNET && rp.mode == kModeArena;
        auto fwd = [&](int k, int k2) {
          return nnet_forward(nets[k], lf.state + static_cast<size_t>(k) * rp.leaf_cap, lf.count + k, rp.leaf_cap,
                              lf.pi + static_cast<size_t>(k) * rp.leaf_cap * 8, lf.v + static_cast<size_t>(k) * rp.leaf_cap, stream,
                              k2 < 0 ? nullptr : nets[k2], lf.state + static_cast<size_t>(std::max(k2, 0)) * rp.leaf_cap,
                              lf.count + std::max(k2, 0), lf.pi + static_cast<size_t>(std::max(k2, 0)) * rp.leaf_cap * 8,
                              lf.v + static_cast<size_t>(std::max(k2, 0)) * rp.leaf_cap);
        };
        int rc = AZB_OK;
        if (use0 && use1 && nets[0] != nets[1]) rc = fwd(0, 1);  // the arena's two players: one tower launch for both
        else {
          if (use0) rc = fwd(0, -1);
          if (!rc && use1) rc = fwd(1, -1);
        }
        if (rc) return rc;
      }
      if (d_times) k_stamp<<<1, 1, 0, stream>>>(d_times, c.round, 3u, kTimesCap);
      return AZB_OK;
    };
    if (any_net) {  // launches of a round, for the statistics: k_compact + k_round + per model stem, tower (one launch or 2R), heads
      const bool tower = g_tower && tc_mode() == 3 && rp.leaf_cap <= 16384u;
      int n_tc = 0;
      for (int k = 0; k < 2; ++k)
        if (nets[k] && rp.ev_kind[k] >= AZB_EVAL_NNET && (k == 0 || rp.mode == kModeArena)) {
          per_round += (nets[k]->cfg.precision == AZB_NNET_FP32) ? 1 : 2 + (tower ? 1 : 2 * nets[k]->L.R);
          n_tc += nets[k]->cfg.precision != AZB_NNET_FP32;
        }
      if (tower && n_tc == 2 && nets[0] != nets[1] && nets[0]->L.R == nets[1]->L.R) per_round -= 1;  // the two towers share a launch
    }
    // the progress word: what the last started round saw
    auto progress = [&](uint64_t* round_seen, uint32_t* live) {
      const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(h_progress);
      *round_seen = v >> 32;
      *live = static_cast<uint32_t>(v);
    };
    auto wrap_clear = [&](uint64_t it) -> int {  // the stamp wraps at this round pair: forget the old rounds' claims
      if (dedup_mask && it > 0 && it % kStampPeriod == 0)
        AZB_CUDA(cudaMemsetAsync(dedup_keys.p, 0, 4 * (static_cast<size_t>(dedup_mask) + 1) * 8, stream));
      return AZB_OK;
    };
    uint64_t it = 0;  // rounds launched
    bool done = false;
    // rounds 0 and 1 one by one: they also run every first-use initialisation of the forward pass outside a capture
    for (; it < 2; ++it) {
      const int rc = launch_round(static_cast<uint32_t>(it & 1u));
      if (rc) return rc;
    }
    if (graph_exec) { cudaGraphExecDestroy(graph_exec); graph_exec = nullptr; }
    if (use_graph) {
      for (int attempt = 0; attempt < 3 && !graph_exec; ++attempt) {
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
          ok = launch_round(0u) == AZB_OK && launch_round(1u) == AZB_OK;
          const bool ended = cudaStreamEndCapture(stream, &graph) == cudaSuccess && graph != nullptr;
          ok = ok && ended;
        }
        if (ok) ok = cudaGraphInstantiate(&graph_exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        if (!ok) {
          cudaGetLastError();
          graph_exec = nullptr;
          if (std::getenv("AZB200_TIMING")) std::fprintf(stderr, "[azb200 rounds] graph capture failed (tower %d, pdl %d)\n", int(g_tower), int(g_tc_pdl));
          if (g_tower) g_tower = false;        // the cooperative tower launch refused inside a capture: layer-by-layer kernels
          else if (g_tc_pdl) g_tc_pdl = false;  // programmatic edges refused inside a capture: plain edges
          else break;
        }
      }
    }
    uint64_t last_seen = 0;
    auto last_change = std::chrono::steady_clock::now();
    while (!done) {
      uint64_t seen = 0;
      uint32_t live = 1;
      progress(&seen, &live);
      if (seen >= 1 && live == 0u) break;
      if (it >= seen + kRunAhead) {  // far enough ahead: let the device catch up
        if (seen != last_seen) {
          last_seen = seen;
          last_change = std::chrono::steady_clock::now();
        } else if (std::chrono::steady_clock::now() - last_change > std::chrono::seconds(60)) {
          // no round has started for a minute: a fault on the device (the next CUDA call reports it) or a stuck kernel
          const cudaError_t e = cudaStreamQuery(stream);
          return fail(AZB_ERR_CUDA, std::string("round engine made no progress for 60 s: ") +
                                        (e == cudaSuccess || e == cudaErrorNotReady ? "kernels still running" : cudaGetErrorString(e)));
        }
        std::this_thread::yield();
        continue;
      }
      int rc = wrap_clear(it);
      if (rc) return rc;
      if (graph_exec) {
        AZB_CUDA(cudaGraphLaunch(graph_exec, stream));
      } else {
        rc = launch_round(0u);
        if (rc) return rc;
        rc = launch_round(1u);
        if (rc) return rc;
      }
      it += 2;
    }
    AZB_CUDA(cudaStreamSynchronize(stream));
    if (d_times) {  // where a round's time goes, in 10 slices of the run
      std::vector<unsigned long long> h(static_cast<size_t>(kTimesCap) * 4);
      AZB_CUDA(cudaMemcpy(h.data(), d_times, h.size() * 8, cudaMemcpyDeviceToHost));
      uint64_t seen = 0;
      uint32_t live = 0;
      progress(&seen, &live);
      const uint64_t R = std::min<uint64_t>(seen > 1 ? seen - 1 : 0, kTimesCap - 1);
      std::fprintf(stderr, "[azb200 rounds] %llu rounds; per slice: rounds, us/round total | k_compact | k_round | forward | gap to next round\n",
                   static_cast<unsigned long long>(R));
      for (int sl = 0; sl < 10 && R >= 20; ++sl) {
        const uint64_t a = R * sl / 10, b = R * (sl + 1) / 10;
        double tot = 0, tc = 0, tr = 0, tf = 0, tg = 0;
        for (uint64_t r = a; r < b; ++r) {
          const unsigned long long* t = h.data() + r * 4;
          const unsigned long long* n = h.data() + (r + 1) * 4;
          tot += double(n[0] - t[0]); tc += double(t[1] - t[0]); tr += double(t[2] - t[1]); tf += double(t[3] - t[2]); tg += double(n[0] - t[3]);
        }
        const double k = 1e-3 / double(b - a);
        std::fprintf(stderr, "[azb200 rounds] slice %d: %6llu rounds  %7.1f | %6.1f | %6.1f | %6.1f | %6.1f\n", sl,
                     static_cast<unsigned long long>(b - a), tot * k, tc * k, tr * k, tf * k, tg * k);
      }
    }
    const uint64_t n_launch = it * per_round;
    if (launches) *launches = n_launch;
    if (nn_positions) {  // (the last k_compact has added the last round's counts)
      unsigned long long total = 0;
      AZB_CUDA(cudaMemcpy(&total, leaf.nn_total, 8, cudaMemcpyDeviceToHost));
      *nn_positions = total;
    }
    if (cache_hits) {
      unsigned long long total = 0;
      AZB_CUDA(cudaMemcpy(&total, leaf.cache_hits, 8, cudaMemcpyDeviceToHost));
      *cache_hits = total;
    }
    return AZB_OK;
  }
};
}  // namespace

// one history entry = one iteration's samples, structure-of-arrays (the layout NNet::train takes)
struct SampleBlock {
  std::vector<float> boards, pis, vs;
  uint64_t size() const { return vs.size(); }
  void drop_front(uint64_t k) {  // coach.rs:275-277: while len > max_queue_length pop_front
    boards.erase(boards.begin(), boards.begin() + static_cast<ptrdiff_t>(k * 84));
    pis.erase(pis.begin(), pis.begin() + static_cast<ptrdiff_t>(k * 7));
    vs.erase(vs.begin(), vs.begin() + static_cast<ptrdiff_t>(k));
  }
};
struct CoachHistory {
  std::deque<SampleBlock> entries;  // struct Coach.history, coach.rs:19
};

struct azb_coach {
  azb_config cfg;
  std::string checkpoint_dir;  // owns what cfg.checkpoint_directory points at
  CoachHistory history;
  uint64_t resume_base = 0;  // Coach::setup resumed from `<n>.examples`: n + 1 (learn.cuh), else 0
  TreePool pool;
  bool pool_ready = false;
  azb_nnet* net = nullptr;
  RoundEngine engine;
  GameStore gs;
  // last self-play call
  uint64_t n_games = 0, n_samples = 0, launches = 0, nn_positions = 0, nn_cache_hits = 0;
  DevBuf next_game, offsets, out_boards, out_pis, out_vs;
  // azb_coach_self_play_begin / _end: the persistent kernel runs on the coach's own non-blocking stream, so that two
  // coaches used in turn overlap one batch's tail with the next batch's head
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, span0 = nullptr;
  bool pending = false;
  uint64_t pending_games = 0, pending_trees = 0;
  ~azb_coach() {
    if (ev0) cudaEventDestroy(ev0);
    if (ev1) cudaEventDestroy(ev1);
    if (span0) cudaEventDestroy(span0);
    if (stream) cudaStreamDestroy(stream);
  }
  std::vector<uint32_t> h_plies;
};

static int coach_resume_history(azb_coach* c);  // learn.cuh

extern "C" {

const char* azb_last_error(void) { return g_err.c_str(); }

// The evaluation cache of the network rounds (2 x 2^25 entries, ~2.7 GB) belongs to the calling thread and is re-used by every
// run of that thread; this gives it back to the device (the next network run allocates it again).
int azb_release_caches(void) {
  RoundEngine::CacheBufs& cb = RoundEngine::thread_cache();
  if (cb.device >= 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaSetDevice(cb.device);
    cb.keys.release();
    cb.vals.release();
    cudaSetDevice(dev);
    cb.device = -1;
  }
  return AZB_OK;
}

int azb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int azb_host_alloc(size_t bytes, void** out) {
  if (!out) return fail(AZB_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  AZB_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault));
  return AZB_OK;
}
int azb_host_free(void* p) {
  if (p) AZB_CUDA(cudaFreeHost(p));
  return AZB_OK;
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
  c->schedule = 0;
  c->plies_per_launch = 0;
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
  if (m->pool.p.num_threads > 1u)
    k_mcts_search<true><<<grid, kWarpsPerCta * 32>>>(m->cfg.evaluator, m->pool.p, m->pool.pools, m->d_states.as<BB>(), temp,
                                                     m->d_counts.as<uint16_t>(), m->d_pi.as<float>(), n);
  else
    k_mcts_search<false><<<grid, kWarpsPerCta * 32>>>(m->cfg.evaluator, m->pool.p, m->pool.pools, m->d_states.as<BB>(), temp,
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
  if (cfg->checkpoint_directory) {
    c->checkpoint_dir = cfg->checkpoint_directory;
    c->cfg.checkpoint_directory = c->checkpoint_dir.c_str();
    rc = coach_resume_history(c.get());  // coach.rs:55-81
    if (rc) return rc;
  }
  *out = c.release();
  return AZB_OK;
}
int azb_coach_destroy(azb_coach* c) {
  delete c;
  return AZB_OK;
}

// azb_coach_self_play in two halves.  _begin sizes the pools, clears the per-game buffers and LAUNCHES the call (the
// persistent kernel of the fused evaluators goes to the coach's own non-blocking stream and _begin returns at once; the
// lock-step rounds of the network evaluator run to completion inside _begin); _end waits, checks the games' error words
// and fills the statistics.  Two coaches used in turn — begin(A), begin(B), end(A), export(A), begin(A), end(B), ... —
// keep the device full across batches: a batch ends with its longest game (42 plies on one warp while the mean game
// has 28), and the warps its finished games vacate are taken by the next batch's CTAs instead of idling.
int azb_coach_self_play_begin(azb_coach* c, uint64_t n_games, uint64_t first_game_id) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (c->pending) return fail(AZB_ERR_INVALID, "a self-play call is already in flight on this coach (azb_coach_self_play_end first)");
  if (n_games == 0 || n_games > (1u << 26)) return fail(AZB_ERR_INVALID, "n_games out of range");
  HostTimer tm("self_play");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  const SearchParams p = make_params(c->cfg, kMaxPlies);
  // how many trees live in HBM at once: all games, capped by co-resident warps, memory and config
  uint32_t resident = 0;
  int rc = resident_trees(c->cfg.device, &resident);
  if (rc) return rc;
  tm.lap("carveout+occupancy");
  if (c->cfg.evaluator >= AZB_EVAL_NNET) resident = round_slots(resident);  // rounds are not tied to co-resident warps
  uint64_t n_trees = std::min<uint64_t>(n_games, resident);
  if (c->cfg.max_concurrent_games) n_trees = std::min<uint64_t>(n_trees, c->cfg.max_concurrent_games);
  // cudaMemGetInfo was measured to take up to 70 ms now and then: ask only when the pool has to change
  const bool pool_fits = c->pool_ready && c->pool.n_trees == n_trees && std::memcmp(&c->pool.p, &p, sizeof(p)) == 0;
  if (!pool_fits) {
    size_t free_b = 0, total_b = 0;
    AZB_CUDA(cudaMemGetInfo(&free_b, &total_b));
    const uint64_t by_mem = (static_cast<uint64_t>(free_b) + c->pool.blocks.bytes + c->pool.tables.bytes) * 8 / 10 / tree_bytes(p);
    n_trees = std::min<uint64_t>(n_trees, by_mem);
  }
  tm.lap("cudaMemGetInfo");
  if (n_trees == 0) return fail(AZB_ERR_CAPACITY, "not enough device memory for one tree");
  if (!c->pool_ready || c->pool.n_trees != n_trees || std::memcmp(&c->pool.p, &p, sizeof(p)) != 0) {
    rc = c->pool.alloc(p, static_cast<uint32_t>(n_trees));
    if (rc) return rc;
    AZB_CUDA(cudaDeviceSynchronize());  // (the pool's clears ran on the legacy stream; the kernel may go to the coach's)
    c->pool_ready = true;
  }
  if (c->cfg.evaluator >= AZB_EVAL_NNET && !c->net)
    return fail(AZB_ERR_INVALID, "evaluator NNET needs azb_coach_set_nnet first");
  // schedule: 1 = one persistent kernel, a warp plays a whole game (fused evaluators only);
  //           2 = lock-step rounds (k_compact / k_round [/ network forward]); 0 = pick
  uint32_t schedule = c->cfg.schedule;
  if (c->cfg.evaluator >= AZB_EVAL_NNET) schedule = 2;
  else if (schedule == 0) schedule = 1;  // measured: re-dealing live games buys nothing at 7 warps/scheduler
  if (!c->stream) {
    AZB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    AZB_CUDA(cudaEventCreate(&c->ev0));
    AZB_CUDA(cudaEventCreate(&c->ev1));
  }
  cudaStream_t st = schedule == 1 ? c->stream : nullptr;
  const uint64_t G = n_games;
  rc = c->gs.alloc(G, true, st);
  if (rc) return rc;
  GameBufs g = c->gs.g;
  tm.lap("pool + game buffers");
  c->n_games = 0;
  c->n_samples = 0;
  AZB_CUDA(cudaEventRecord(c->ev0, st));
  if (schedule == 1) {
    AZB_CUDA(c->next_game.ensure(4));
    AZB_CUDA(cudaMemsetAsync(c->next_game.p, 0, 4, st));
    const unsigned grid = static_cast<unsigned>((n_trees + kWarpsPerCta - 1) / kWarpsPerCta);
    if (p.num_threads > 1u)
      k_selfplay<true><<<grid, kWarpsPerCta * 32, 0, st>>>(c->cfg.evaluator, p, c->pool.pools, g, static_cast<uint32_t>(n_trees),
                                                           static_cast<uint32_t>(G), first_game_id, c->next_game.as<unsigned int>());
    else
      k_selfplay<false><<<grid, kWarpsPerCta * 32, 0, st>>>(c->cfg.evaluator, p, c->pool.pools, g, static_cast<uint32_t>(n_trees),
                                                            static_cast<uint32_t>(G), first_game_id, c->next_game.as<unsigned int>());
    AZB_CUDA(cudaGetLastError());
    c->launches = 1;
    c->nn_positions = 0;
    c->nn_cache_hits = 0;
  } else {
    rc = c->engine.alloc(static_cast<uint32_t>(n_trees), p.num_threads);
    if (rc) return rc;
    RoundParams rp{};
    rp.p = p;
    rp.mode = kModeSelfPlay;
    rp.ev_kind[0] = rp.ev_kind[1] = c->cfg.evaluator;
    rp.plies_per_launch = c->cfg.evaluator >= AZB_EVAL_NNET ? 0u : (c->cfg.plies_per_launch ? c->cfg.plies_per_launch : 2u);
    rp.sims_per_launch = c->cfg.evaluator >= AZB_EVAL_NNET ? round_sim_budget() : 0u;
    rp.n_slots = static_cast<uint32_t>(n_trees);
    rp.leaf_cap = c->engine.leaf_cap;
    rp.n_games = static_cast<uint32_t>(G);
    rp.first_game_id = first_game_id;
    azb_nnet* nets[2] = {c->net, nullptr};
    rc = c->engine.run(rp, c->pool.pools, c->gs, nets, &c->launches, &c->nn_positions, &c->nn_cache_hits);
    if (rc) return rc;
  }
  AZB_CUDA(cudaEventRecord(c->ev1, st));
  tm.lap("launch");
  c->pending = true;
  c->pending_games = G;
  c->pending_trees = n_trees;
  return AZB_OK;
}

int azb_coach_self_play_end(azb_coach* c, azb_selfplay_stats* stats) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (!c->pending) return fail(AZB_ERR_INVALID, "no self-play call in flight (azb_coach_self_play_begin first)");
  HostTimer tm("self_play_end");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  c->pending = false;
  const uint64_t G = c->pending_games;
  AZB_CUDA(cudaEventSynchronize(c->ev1));
  tm.lap("kernel wait");
  float ms = 0.0f;
  AZB_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));

  c->h_plies.resize(G);
  std::vector<uint32_t> h_err(G), h_stats(G * 8);
  AZB_CUDA(cudaMemcpy(c->h_plies.data(), c->gs.plies.p, G * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_err.data(), c->gs.error.p, G * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_stats.data(), c->gs.stats.p, G * 32, cudaMemcpyDeviceToHost));
  tm.lap("stats D2H");
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
  s.launches = c->launches;
  s.trees_resident = c->pending_trees;
  s.nn_positions = c->nn_positions;
  s.nn_cache_hits = c->nn_cache_hits;
  c->n_games = G;
  c->n_samples = s.samples;
  if (stats) *stats = s;
  return AZB_OK;
}

// Device time across pipelined calls: azb_coach_span_mark(first) records an event on that coach's stream before its next
// _begin; azb_coach_span_ms(first, last) = from that mark to the end of `last`'s most recent call (after its _end).
int azb_coach_span_mark(azb_coach* c) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  if (!c->stream) {
    AZB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    AZB_CUDA(cudaEventCreate(&c->ev0));
    AZB_CUDA(cudaEventCreate(&c->ev1));
  }
  if (!c->span0) AZB_CUDA(cudaEventCreate(&c->span0));
  AZB_CUDA(cudaEventRecord(c->span0, c->stream));
  return AZB_OK;
}
int azb_coach_span_ms(azb_coach* first, azb_coach* last, double* ms) {
  if (!first || !last || !ms || !first->span0 || !last->ev1) return fail(AZB_ERR_INVALID, "NULL argument / no mark");
  float f = 0.0f;
  AZB_CUDA(cudaEventSynchronize(last->ev1));
  AZB_CUDA(cudaEventElapsedTime(&f, first->span0, last->ev1));
  *ms = f;
  return AZB_OK;
}

int azb_coach_self_play(azb_coach* c, uint64_t n_games, uint64_t first_game_id, azb_selfplay_stats* stats) {
  const int rc = azb_coach_self_play_begin(c, n_games, first_game_id);
  if (rc) return rc;
  return azb_coach_self_play_end(c, stats);
}

int azb_coach_traces(azb_coach* c, uint8_t* actions, uint16_t* root_counts, uint32_t* plies, float* final_r,
                     int8_t* final_player) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (c->n_games == 0) return fail(AZB_ERR_INVALID, "no self-play results");
  AZB_CUDA(cudaSetDevice(c->cfg.device));
  const uint64_t G = c->n_games;
  if (actions) AZB_CUDA(cudaMemcpy(actions, c->gs.actions.p, G * kTraceStride, cudaMemcpyDeviceToHost));
  if (root_counts) AZB_CUDA(cudaMemcpy(root_counts, c->gs.counts.p, G * kTraceStride * 14, cudaMemcpyDeviceToHost));
  if (plies) AZB_CUDA(cudaMemcpy(plies, c->gs.plies.p, G * 4, cudaMemcpyDeviceToHost));
  if (final_r) AZB_CUDA(cudaMemcpy(final_r, c->gs.final_r.p, G * 4, cudaMemcpyDeviceToHost));
  if (final_player) AZB_CUDA(cudaMemcpy(final_player, c->gs.final_player.p, G, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

int azb_coach_ply_times(azb_coach* c, uint64_t* ns) {
  if (!c || !ns) return fail(AZB_ERR_INVALID, "NULL argument");
  if (c->n_games == 0 || !c->gs.g.ply_ns) return fail(AZB_ERR_INVALID, "no ply times (set AZB200_PLY_TIMES=1 before self-play)");
  AZB_CUDA(cudaMemcpy(ns, c->gs.ply_ns.p, c->n_games * kTraceStride * 8, cudaMemcpyDeviceToHost));
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
  // sized for the most samples G games can produce (42 plies x 2 symmetries), so that a step with a
  // few more samples than the last one never pays a cudaFree + cudaMalloc
  const uint64_t cap_n = std::max<uint64_t>(N, G * kMaxPlies * 2);
  AZB_CUDA(c->out_boards.ensure(cap_n * 84 * 4));
  AZB_CUDA(c->out_pis.ensure(cap_n * 7 * 4));
  AZB_CUDA(c->out_vs.ensure(cap_n * 4));
  k_export_samples<<<static_cast<unsigned>(G), 128>>>(c->gs.g, c->offsets.as<uint64_t>(), c->cfg.quirks,
                                                      c->out_boards.as<float>(), c->out_pis.as<float>(),
                                                      c->out_vs.as<float>(), N);
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(boards, c->out_boards.p, N * 84 * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(pis, c->out_pis.p, N * 7 * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(vs, c->out_vs.p, N * 4, cudaMemcpyDeviceToHost));
  if (n_written) *n_written = N;
  return AZB_OK;
}

// ---------------------------------------------------------------------------------------------
// NNet
// ---------------------------------------------------------------------------------------------
static uint64_t sm64(uint64_t& x) {
  x += 0x9E3779B97F4A7C15ull;
  uint64_t z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static void he_normal(float* w, size_t n, size_t fan_in, uint64_t& st) {
  const double sd = std::sqrt(2.0 / static_cast<double>(fan_in));
  for (size_t i = 0; i < n; i += 2) {
    const double u1 = (static_cast<double>(sm64(st) >> 11) + 1.0) / 9007199254740993.0;
    const double u2 = static_cast<double>(sm64(st) >> 11) / 9007199254740992.0;
    const double r = std::sqrt(-2.0 * std::log(u1));
    w[i] = static_cast<float>(sd * r * std::cos(6.283185307179586 * u2));
    if (i + 1 < n) w[i + 1] = static_cast<float>(sd * r * std::sin(6.283185307179586 * u2));
  }
}

int azb_nnet_create(const azb_nnet_config* cfg, azb_nnet** out) {
  if (!cfg || !out) return fail(AZB_ERR_INVALID, "NULL argument");
  *out = nullptr;
  if (cfg->blocks < 1 || cfg->blocks > 40) return fail(AZB_ERR_INVALID, "blocks out of range");
  if (cfg->precision != AZB_NNET_BF16_TC && cfg->precision != AZB_NNET_FP32) return fail(AZB_ERR_INVALID, "precision");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  AZB_CUDA(cudaSetDevice(cfg->device));
  auto n = std::make_unique<azb_nnet>();
  n->cfg = *cfg;
  n->L = net_layout(cfg->blocks);
  const NetLayout& L = n->L;
  n->h_params.assign(L.total, 0.0f);
  float* w = n->h_params.data();
  uint64_t st = cfg->seed * 0x9E3779B97F4A7C15ull + 0x1234567ull;
  he_normal(w + L.stem_w, 9 * 2 * kNetC, 9 * 2, st);
  he_normal(w + L.tower_w, static_cast<size_t>(2 * L.R) * 9 * kNetC * kNetC, 9 * kNetC, st);
  he_normal(w + L.pol_w, kNetC * 2, kNetC, st);
  he_normal(w + L.pol_fc_w, 84 * 7, 84, st);
  he_normal(w + L.val_w, kNetC, kNetC, st);
  he_normal(w + L.val_fc1_w, 42 * 64, 42, st);
  he_normal(w + L.val_fc2_w, 64, 64, st);
  int rc = n->upload();
  if (rc) return rc;
  *out = n.release();
  return AZB_OK;
}
int azb_nnet_destroy(azb_nnet* n) {
  delete n;
  return AZB_OK;
}
int azb_nnet_num_params(azb_nnet* n, uint64_t* count) {
  if (!n || !count) return fail(AZB_ERR_INVALID, "NULL argument");
  *count = n->L.total;
  return AZB_OK;
}
int azb_nnet_get_params(azb_nnet* n, float* out, uint64_t capacity) {
  if (!n || !out) return fail(AZB_ERR_INVALID, "NULL argument");
  if (capacity < n->L.total) return fail(AZB_ERR_CAPACITY, "parameter buffer too small");
  const int rc = n->sync_host();
  if (rc) return rc;
  std::memcpy(out, n->h_params.data(), n->L.total * 4);
  return AZB_OK;
}
int azb_nnet_set_params(azb_nnet* n, const float* in, uint64_t count) {
  if (!n || !in) return fail(AZB_ERR_INVALID, "NULL argument");
  if (count != n->L.total) return fail(AZB_ERR_INVALID, "parameter count mismatch");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  std::memcpy(n->h_params.data(), in, count * 4);
  n->host_stale = false;
  return n->upload();
}
int azb_nnet_predict(azb_nnet* n, const float* boards, size_t batch, size_t /*model_id*/, float* pi, float* v) {
  if (!n || (batch && (!boards || !pi || !v))) return fail(AZB_ERR_INVALID, "NULL argument");
  if (batch == 0) return AZB_OK;
  if (batch > (1u << 26)) return fail(AZB_ERR_INVALID, "batch too large");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  AZB_CUDA(n->d_feat.ensure(batch * 84 * 4));
  AZB_CUDA(n->d_states.ensure(batch * 16));
  AZB_CUDA(n->d_pi.ensure(batch * 32));
  AZB_CUDA(n->d_v.ensure(batch * 4));
  AZB_CUDA(cudaMemcpy(n->d_feat.p, boards, batch * 84 * 4, cudaMemcpyHostToDevice));
  const uint32_t B = static_cast<uint32_t>(batch);
  k_features_to_bb<<<(B + 127) / 128, 128>>>(n->d_feat.as<float>(), B, n->d_states.as<uint4>());
  int rc = nnet_forward(n, n->d_states.as<uint4>(), nullptr, B, n->d_pi.as<float>(), n->d_v.as<float>(), 0);
  if (rc) return rc;
  AZB_CUDA(cudaGetLastError());
  std::vector<float> h(batch * 8);
  AZB_CUDA(cudaMemcpy(h.data(), n->d_pi.p, batch * 32, cudaMemcpyDeviceToHost));
  for (size_t i = 0; i < batch; ++i)
    for (int a = 0; a < 7; ++a) pi[i * 7 + a] = h[i * 8 + a];
  AZB_CUDA(cudaMemcpy(v, n->d_v.p, batch * 4, cudaMemcpyDeviceToHost));
  return AZB_OK;
}
int azb_nnet_benchmark(azb_nnet* n, uint64_t batch, uint32_t iters, double* ms_per_pass) {
  if (!n || !ms_per_pass || batch == 0 || iters == 0 || batch > (1u << 22)) return fail(AZB_ERR_INVALID, "bad argument");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  std::vector<uint4> h(batch);
  uint64_t st = 42;
  for (auto& x : h) {  // random disjoint stone sets (not necessarily reachable positions)
    const uint64_t a = sm64(st) & kBoard42, b = sm64(st) & kBoard42 & ~a;
    x = make_uint4(static_cast<uint32_t>(a), static_cast<uint32_t>(a >> 32), static_cast<uint32_t>(b), static_cast<uint32_t>(b >> 32));
  }
  AZB_CUDA(n->d_states.ensure(batch * 16));
  AZB_CUDA(n->d_pi.ensure(batch * 32));
  AZB_CUDA(n->d_v.ensure(batch * 4));
  AZB_CUDA(cudaMemcpy(n->d_states.p, h.data(), batch * 16, cudaMemcpyHostToDevice));
  const uint32_t B = static_cast<uint32_t>(batch);
  for (int w = 0; w < 3; ++w) {
    int rc = nnet_forward(n, n->d_states.as<uint4>(), nullptr, B, n->d_pi.as<float>(), n->d_v.as<float>(), 0);
    if (rc) return rc;
  }
  EventPair ev;
  AZB_CUDA(ev.create());
  AZB_CUDA(cudaEventRecord(ev.e0));
  for (uint32_t i = 0; i < iters; ++i) {
    int rc = nnet_forward(n, n->d_states.as<uint4>(), nullptr, B, n->d_pi.as<float>(), n->d_v.as<float>(), 0);
    if (rc) return rc;
  }
  AZB_CUDA(cudaEventRecord(ev.e1));
  AZB_CUDA(cudaEventSynchronize(ev.e1));
  float ms = 0.0f;
  AZB_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
  *ms_per_pass = ms / iters;
  return AZB_OK;
}
namespace {
int tc3_pairs() {
  static int n = -1;
  if (n < 0) {
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(148u);
    qc.blockDim = dim3(kTcThreads);
    qc.dynamicSmemBytes = kT3SmemBytes;
    int k = 0;
    if (cudaOccupancyMaxActiveClusters(&k, k_conv3x3_tc3, &qc) != cudaSuccess) { cudaGetLastError(); k = 0; }
    n = k;
  }
  return n;
}
// one k_conv3x3_tc3 launch on padded buffers of the training step
int train_conv(azb_nnet* n, int layer, int mode, const DevBuf& in, const CUtensorMap& in_map, const DevBuf* residual, const DevBuf* mask,
               DevBuf& out, uint32_t B) {
  ConvTcArgs a{};
  a.in = in.as<__nv_bfloat16>();
  a.residual = residual ? residual->as<__nv_bfloat16>() : nullptr;
  a.mask = mask ? mask->as<__nv_bfloat16>() : nullptr;
  a.out = out.as<__nv_bfloat16>();
  a.mode = mode;
  a.w_tiles = (mode == 0 ? n->d_wtiles.as<uint8_t>() : n->d_wtiles_bwd.as<uint8_t>()) + static_cast<size_t>(layer) * kTcKBlocks * kTcTileBytes;
  a.bias = n->d_params.as<float>() + n->L.tower_b + static_cast<size_t>(layer) * kNetC;
  a.max_batch = B;
  const uint32_t pair_tiles = (B * kActPadded.pos_rows + kT2PairRows - 1) / kT2PairRows;
  const int pairs = tc3_pairs();
  if (pairs <= 0) return fail(AZB_ERR_CUDA, "no co-resident CTA pair for k_conv3x3_tc3");
  k_conv3x3_tc3<<<2u * std::min<uint32_t>(pair_tiles, static_cast<uint32_t>(pairs)), kTcThreads, kT3SmemBytes>>>(a, in_map);
  AZB_CUDA(cudaGetLastError());
  return AZB_OK;
}
}  // namespace

// NNet::train (src/nnet.rs:38), first half: forward with every layer kept, loss (softmax cross-entropy on pi + squared
// error on v, means over the batch), backward; the gradients stay on the device (azb_nnet_grads / azb_nnet_set_grads let
// a data-parallel caller all-reduce them), azb_nnet_train_apply is the optimiser step.
int azb_nnet_train_begin(azb_nnet* n, const float* boards, const float* pis, const float* vs, uint64_t count, float* loss_out) {
  if (!n || !boards || !pis || !vs || count == 0 || count > (1u << 20)) return fail(AZB_ERR_INVALID, "bad argument");
  if (n->cfg.precision != AZB_NNET_BF16_TC || tc_mode() != 3) return fail(AZB_ERR_UNSUPPORTED, "training needs the default tensor-core tower");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  HostTimer tm("train_begin");
  auto lap = [&](const char* what) { if (tm.on) { cudaDeviceSynchronize(); tm.lap(what); } };
  const uint32_t B = static_cast<uint32_t>(count);
  const int nl = 2 * n->L.R;
  const NetLayout& L = n->L;
  // inputs
  AZB_CUDA(n->d_feat.ensure(count * 84 * 4));
  AZB_CUDA(n->d_states.ensure(count * 16));
  AZB_CUDA(n->d_pis.ensure(count * 7 * 4));
  AZB_CUDA(n->d_vs.ensure(count * 4));
  AZB_CUDA(cudaMemcpy(n->d_feat.p, boards, count * 84 * 4, cudaMemcpyHostToDevice));
  AZB_CUDA(cudaMemcpy(n->d_pis.p, pis, count * 7 * 4, cudaMemcpyHostToDevice));
  AZB_CUDA(cudaMemcpy(n->d_vs.p, vs, count * 4, cudaMemcpyHostToDevice));
  k_features_to_bb<<<(B + 127) / 128, 128>>>(n->d_feat.as<float>(), B, n->d_states.as<uint4>());
  // buffers: zeroed once when (re)allocated, so that the padded layout's zero rows stay zero
  const size_t bytes = (static_cast<size_t>(B) + 8) * kActPadded.pos_rows * kNetC * 2;
  if (n->tr_act.size() != static_cast<size_t>(nl + 1) || n->tr_bytes < bytes) {
    n->tr_act.clear();
    n->tr_act.resize(nl + 1);
    n->tr_act_map.resize(nl + 1);
    for (int i = 0; i <= nl; ++i) {
      AZB_CUDA(n->tr_act[i].ensure(bytes));
      AZB_CUDA(cudaMemset(n->tr_act[i].p, 0, bytes));
      const int rc = encode_act_map_rows(&n->tr_act_map[i], n->tr_act[i].p, bytes);
      if (rc) return rc;
    }
    for (int i = 0; i < 3; ++i) {
      n->tr_g[i].release();
      AZB_CUDA(n->tr_g[i].ensure(bytes));
      AZB_CUDA(cudaMemset(n->tr_g[i].p, 0, bytes));
      const int rc = encode_act_map_rows(&n->tr_g_map[i], n->tr_g[i].p, bytes);
      if (rc) return rc;
    }
    n->tr_bytes = bytes;
    n->tr_batch = B;
  }
  if (n->tr_batch != B) {  // rows of a larger earlier batch would leak into the weight gradients' last row chunk
    for (auto& b : n->tr_act) AZB_CUDA(cudaMemset(b.p, 0, b.bytes));
    for (auto& b : n->tr_g) AZB_CUDA(cudaMemset(b.p, 0, b.bytes));
    n->tr_batch = B;
  }
  AZB_CUDA(n->d_grad.ensure(L.total * 4));
  AZB_CUDA(n->d_loss.ensure(8));
  AZB_CUDA(cudaMemset(n->d_grad.p, 0, L.total * 4));
  AZB_CUDA(cudaMemset(n->d_loss.p, 0, 8));
  const float* prm = n->d_params.as<float>();
  float* grad = n->d_grad.as<float>();
  lap("inputs + buffers");
  // ---- forward, every layer kept ----
  const size_t total = static_cast<size_t>(B) * kCells * (kNetC / 8);
  k_stem_bf16<<<static_cast<unsigned>(std::min<size_t>((total + 1023) / 1024, 148u * 2u)), 1024, kStemSmemBytes>>>(
      prm, L, n->d_stem_tab.as<float>(), n->d_states.as<uint4>(), nullptr, B, n->tr_act[0].as<__nv_bfloat16>(), kActPadded);
  AZB_CUDA(cudaGetLastError());
  for (int l = 0; l < nl; ++l) {
    const int rc = train_conv(n, l, 0, n->tr_act[l], n->tr_act_map[l], (l & 1) ? &n->tr_act[l - 1] : nullptr, nullptr, n->tr_act[l + 1], B);
    if (rc) return rc;
  }
  lap("forward");
  // ---- heads: loss and gradients; g = dL/d(pre-activation of the last convolution) ----
  int gi = 0;  // index of the current gradient buffer
  AZB_CUDA(cudaFuncSetAttribute(k_heads_backward, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kHeadsBwdSmem)));
  k_heads_backward<<<std::min<uint32_t>(B, 148u * 4u), 128, kHeadsBwdSmem>>>(prm, L, n->tr_act[nl].as<__nv_bfloat16>(), B, n->d_pis.as<float>(),
                                                                            n->d_vs.as<float>(), 1.0f / static_cast<float>(B), kActPadded,
                                                                            n->tr_g[gi].as<__nv_bfloat16>(), grad, n->d_loss.as<float>());
  AZB_CUDA(cudaGetLastError());
  lap("heads backward");
  // ---- backward through the tower ----
  const uint32_t rows = B * kActPadded.pos_rows;
  int skip = -1;
  for (int l = nl - 1; l >= 0; --l) {
    WgradArgs wg{};
    wg.dw = grad + L.tower_w + static_cast<size_t>(l) * 9 * kNetC * kNetC;
    wg.n_pos = B;
    k_conv3x3_wgrad<<<147, kWgThreads, kWgSmemBytes>>>(wg, n->tr_act_map[l], n->tr_g_map[gi]);
    k_colsum_bf16<<<148 * 4, 256>>>(n->tr_g[gi].as<__nv_bfloat16>(), rows, grad + L.tower_b + static_cast<size_t>(l) * kNetC);
    AZB_CUDA(cudaGetLastError());
    // gradient of the previous layer's pre-activation: conv^T(g) [+ the block's skip path], gated by that layer's ReLU
    int go = 0;
    while (go == gi || go == skip) ++go;
    const int rc = train_conv(n, l, 1, n->tr_g[gi], n->tr_g_map[gi], (l & 1) ? nullptr : &n->tr_g[skip], &n->tr_act[l], n->tr_g[go], B);
    if (rc) return rc;
    if (l & 1) skip = gi; else skip = -1;
    if (!(l & 1)) { /* the skip buffer is free again */ }
    gi = go;
  }
  lap("tower backward");
  k_stem_backward<<<148 * 2, 128 * kStemBwdGroups>>>(n->d_states.as<uint4>(), n->tr_g[gi].as<__nv_bfloat16>(), B, kActPadded, grad + L.stem_w, grad + L.stem_b);
  AZB_CUDA(cudaGetLastError());
  lap("stem backward");
  float hl[2];
  AZB_CUDA(cudaMemcpy(hl, n->d_loss.p, 8, cudaMemcpyDeviceToHost));
  if (loss_out) { loss_out[0] = hl[0]; loss_out[1] = hl[1]; }
  n->grads_ready = true;
  return AZB_OK;
}

int azb_nnet_grads(azb_nnet* n, float* out, uint64_t count) {
  if (!n || !out || !n->grads_ready || count != n->L.total) return fail(AZB_ERR_INVALID, "no gradients / wrong count");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  AZB_CUDA(cudaMemcpy(out, n->d_grad.p, count * 4, cudaMemcpyDeviceToHost));
  return AZB_OK;
}
int azb_nnet_set_grads(azb_nnet* n, const float* in, uint64_t count) {
  if (!n || !in || count != n->L.total) return fail(AZB_ERR_INVALID, "wrong count");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  AZB_CUDA(n->d_grad.ensure(count * 4));
  AZB_CUDA(cudaMemcpy(n->d_grad.p, in, count * 4, cudaMemcpyHostToDevice));
  n->grads_ready = true;
  return AZB_OK;
}

int azb_nnet_grads_device(azb_nnet* n, void** ptr, uint64_t* count) {
  if (!n || !ptr || !count || !n->grads_ready) return fail(AZB_ERR_INVALID, "no gradients");
  *ptr = n->d_grad.p;
  *count = n->L.total;
  return AZB_OK;
}

// Adam step on the fp32 master parameters, then everything derived from them (bf16 operand tiles forward / backward,
// the stem table, the head weights in the constant bank, the host copy).
int azb_nnet_train_apply(azb_nnet* n, const azb_train_config* cfg) {
  if (!n || !cfg || !n->grads_ready) return fail(AZB_ERR_INVALID, "no gradients");
  if (n->cfg.precision != AZB_NNET_BF16_TC) return fail(AZB_ERR_UNSUPPORTED, "training needs the tensor-core tower");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  HostTimer tm("train_apply");
  auto lap = [&](const char* what) { if (tm.on) { cudaDeviceSynchronize(); tm.lap(what); } };
  const size_t N = n->L.total;
  if (n->d_adam_m.bytes < N * 4) {
    AZB_CUDA(n->d_adam_m.ensure(N * 4));
    AZB_CUDA(n->d_adam_v.ensure(N * 4));
    AZB_CUDA(cudaMemset(n->d_adam_m.p, 0, N * 4));
    AZB_CUDA(cudaMemset(n->d_adam_v.p, 0, N * 4));
    n->adam_t = 0;
  }
  n->adam_t++;
  const float c1 = 1.0f - std::pow(cfg->beta1, static_cast<float>(n->adam_t)), c2 = 1.0f - std::pow(cfg->beta2, static_cast<float>(n->adam_t));
  k_adam<<<static_cast<unsigned>((N + 255) / 256), 256>>>(n->d_params.as<float>(), n->d_grad.as<float>(), n->d_adam_m.as<float>(),
                                                         n->d_adam_v.as<float>(), N, cfg->lr, cfg->beta1, cfg->beta2, cfg->eps, c1, c2);
  const size_t nt = static_cast<size_t>(2 * n->L.R) * kTcKBlocks * 128 * 64;
  k_build_tiles<<<static_cast<unsigned>((nt + 255) / 256), 256>>>(n->d_params.as<float>(), n->L, n->d_wtiles.as<uint16_t>(),
                                                                 n->d_wtiles_bwd.as<uint16_t>());
  k_build_stem_table<<<(3 * 64 * kNetC + 255) / 256, 256>>>(n->d_params.as<float>(), n->L, n->d_stem_tab.as<float>());
  AZB_CUDA(cudaGetLastError());
  lap("adam + tiles + table");
  for (int c = 1; c < kTcWeightCopies; ++c)  // the replicas the single-CTA kernel streams from
    AZB_CUDA(cudaMemcpy(n->d_wtiles.as<uint8_t>() + c * n->wtile_copy_bytes, n->d_wtiles.p, n->wtile_copy_bytes, cudaMemcpyDeviceToDevice));
  lap("replicas");
  // only the head range comes back now (the 1x1 head convolutions travel as a kernel parameter); the rest of the host
  // copy is refreshed when someone asks for it (get_params / save / copy)
  const size_t h0 = n->L.pol_w, h1 = n->L.val_b + 1;
  AZB_CUDA(cudaMemcpy(n->h_params.data() + h0, n->d_params.as<float>() + h0, (h1 - h0) * 4, cudaMemcpyDeviceToHost));
  n->host_stale = true;
  lap("head params to host");
  for (int ci = 0; ci < kNetC; ++ci) {
    n->head_w.w[ci][0] = n->h_params[n->L.pol_w + ci * 2 + 0];
    n->head_w.w[ci][1] = n->h_params[n->L.pol_w + ci * 2 + 1];
    n->head_w.w[ci][2] = n->h_params[n->L.val_w + ci];
  }
  n->head_w.b[0] = n->h_params[n->L.pol_b + 0];
  n->head_w.b[1] = n->h_params[n->L.pol_b + 1];
  n->head_w.b[2] = n->h_params[n->L.val_b];
  n->grads_ready = false;
  return AZB_OK;
}

int azb_nnet_train(azb_nnet* n, const float* boards, const float* pis, const float* vs, uint64_t count, const azb_train_config* cfg,
                   float* loss_out) {
  const int rc = azb_nnet_train_begin(n, boards, pis, vs, count, loss_out);
  return rc ? rc : azb_nnet_train_apply(n, cfg);
}

namespace {
// a padded bf16 device copy of an fp32 dense host tensor [n_pos][42][128] (zero rows included), or an empty buffer
int upload_padded(const float* host, uint64_t n_pos, DevBuf& stage, DevBuf& out) {
  const size_t dense = static_cast<size_t>(n_pos) * kCells * kNetC, padded_bytes = (static_cast<size_t>(n_pos) + 8) * kActPadded.pos_rows * kNetC * 2;
  AZB_CUDA(out.ensure(padded_bytes));
  AZB_CUDA(cudaMemset(out.p, 0, out.bytes));
  if (!host) return AZB_OK;
  AZB_CUDA(stage.ensure(dense * 4));
  AZB_CUDA(cudaMemcpy(stage.p, host, dense * 4, cudaMemcpyHostToDevice));
  k_dense_f32_to_padded_bf16<<<static_cast<unsigned>((dense + 255) / 256), 256>>>(stage.as<float>(), static_cast<uint32_t>(n_pos),
                                                                                  out.as<__nv_bfloat16>());
  AZB_CUDA(cudaGetLastError());
  return AZB_OK;
}
}  // namespace

int azb_nnet_conv_hook(azb_nnet* n, int32_t layer, int32_t mode, const float* x, const float* residual, const float* mask,
                       uint64_t n_pos, float* out) {
  if (!n || !x || !out || n_pos == 0 || n_pos > (1u << 20)) return fail(AZB_ERR_INVALID, "bad argument");
  if (n->cfg.precision != AZB_NNET_BF16_TC || tc_mode() != 3) return fail(AZB_ERR_UNSUPPORTED, "needs the default tensor-core tower");
  if (layer < 0 || layer >= 2 * n->L.R || mode < 0 || mode > 1) return fail(AZB_ERR_INVALID, "layer / mode out of range");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  DevBuf stage, dx, dres, dmask, dout;
  int rc;
  if ((rc = upload_padded(x, n_pos, stage, dx)) || (rc = upload_padded(nullptr, n_pos, stage, dout))) return rc;
  if (residual && (rc = upload_padded(residual, n_pos, stage, dres))) return rc;
  if (mask && (rc = upload_padded(mask, n_pos, stage, dmask))) return rc;
  CUtensorMap map;
  if ((rc = encode_act_map_rows(&map, dx.p, dx.bytes))) return rc;
  ConvTcArgs a{};
  a.in = dx.as<__nv_bfloat16>();
  a.residual = residual ? dres.as<__nv_bfloat16>() : nullptr;
  a.mask = mask ? dmask.as<__nv_bfloat16>() : nullptr;
  a.out = dout.as<__nv_bfloat16>();
  a.mode = mode;
  a.w_tiles = (mode == 0 ? n->d_wtiles.as<uint8_t>() : n->d_wtiles_bwd.as<uint8_t>()) + static_cast<size_t>(layer) * kTcKBlocks * kTcTileBytes;
  a.bias = n->d_params.as<float>() + n->L.tower_b + static_cast<size_t>(layer) * kNetC;
  a.max_batch = static_cast<uint32_t>(n_pos);
  const uint32_t pair_tiles = (a.max_batch * kActPadded.pos_rows + kT2PairRows - 1) / kT2PairRows;
  k_conv3x3_tc3<<<2u * std::min<uint32_t>(pair_tiles, 64u), kTcThreads, kT3SmemBytes>>>(a, map);
  AZB_CUDA(cudaGetLastError());
  const size_t dense = static_cast<size_t>(n_pos) * kCells * kNetC;
  AZB_CUDA(stage.ensure(dense * 4));
  k_padded_bf16_to_dense_f32<<<static_cast<unsigned>((dense + 255) / 256), 256>>>(dout.as<__nv_bfloat16>(), a.max_batch, stage.as<float>());
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(out, stage.p, dense * 4, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

int azb_nnet_wgrad_hook(azb_nnet* n, const float* x, const float* dz, uint64_t n_pos, float* dw) {
  if (!n || !x || !dz || !dw || n_pos == 0 || n_pos > (1u << 20)) return fail(AZB_ERR_INVALID, "bad argument");
  if (n->cfg.precision != AZB_NNET_BF16_TC || tc_mode() != 3) return fail(AZB_ERR_UNSUPPORTED, "needs the default tensor-core tower");
  AZB_CUDA(cudaSetDevice(n->cfg.device));
  DevBuf stage, dx, ddz, ddw;
  int rc;
  if ((rc = upload_padded(x, n_pos, stage, dx)) || (rc = upload_padded(dz, n_pos, stage, ddz))) return rc;
  CUtensorMap mx, mz;
  if ((rc = encode_act_map_rows(&mx, dx.p, dx.bytes)) || (rc = encode_act_map_rows(&mz, ddz.p, ddz.bytes))) return rc;
  const size_t nw = static_cast<size_t>(9) * kNetC * kNetC;
  AZB_CUDA(ddw.ensure(nw * 4));
  AZB_CUDA(cudaMemset(ddw.p, 0, nw * 4));
  WgradArgs g{};
  g.dw = ddw.as<float>();
  g.n_pos = static_cast<uint32_t>(n_pos);
  k_conv3x3_wgrad<<<147, kWgThreads, kWgSmemBytes>>>(g, mx, mz);
  AZB_CUDA(cudaGetLastError());
  AZB_CUDA(cudaMemcpy(dw, ddw.p, nw * 4, cudaMemcpyDeviceToHost));
  return AZB_OK;
}

int azb_coach_set_nnet(azb_coach* c, azb_nnet* n) {
  if (!c) return fail(AZB_ERR_INVALID, "NULL argument");
  if (n && n->cfg.device != c->cfg.device) return fail(AZB_ERR_INVALID, "network and coach live on different devices");
  c->net = n;
  return AZB_OK;
}

// ---------------------------------------------------------------------------------------------
// arena
// ---------------------------------------------------------------------------------------------
int azb_arena_play_games_ex(const azb_config* cfg, uint64_t num, int32_t eval_a, int32_t eval_b, azb_nnet* net_a,
                            azb_nnet* net_b, const azb_arena_opts* opts, uint64_t out_counts[3], int8_t* results,
                            uint8_t* actions, uint16_t* root_counts, uint32_t* plies, azb_selfplay_stats* stats) {
  if (!out_counts) return fail(AZB_ERR_INVALID, "NULL argument");
  const azb_arena_opts o = opts ? *opts : azb_arena_opts{0u, 0u, 0ull};
  const uint32_t k_open = o.k_open;
  const bool shared = o.shared_trees != 0u;
  int rc = validate(cfg);
  if (rc) return rc;
  for (int k = 0; k < 2; ++k) {
    const int32_t ev = k ? eval_b : eval_a;
    if (ev < AZB_EVAL_UNIFORM || ev > AZB_EVAL_NNET) return fail(AZB_ERR_INVALID, "bad evaluator kind");
    if (ev == AZB_EVAL_NNET && !(k ? net_b : net_a)) return fail(AZB_ERR_INVALID, "evaluator NNET needs a network");
  }
  const uint64_t half = num / 2, G = 2 * half;  // arena.rs:83: num / 2 games per seat order
  out_counts[0] = out_counts[1] = out_counts[2] = 0;
  if (G == 0) return AZB_OK;
  if (G > (1u << 26)) return fail(AZB_ERR_INVALID, "num out of range");
  if (azb_device_count() == 0) return fail(AZB_ERR_CUDA, "no CUDA device: libazb200 has no CPU fallback");
  AZB_CUDA(cudaSetDevice(cfg->device));
  // each player's tree sees every second ply: at most 21 searches + F12 roots (per game; a shared tree pair lives
  // through all G games of the match)
  const SearchParams p = make_params(*cfg, (kMaxPlies / 2 + 1) * (shared ? G : 1));
  uint32_t resident = 0;
  rc = resident_trees(cfg->device, &resident);
  if (rc) return rc;
  size_t free_b = 0, total_b = 0;
  AZB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  uint64_t by_mem = static_cast<uint64_t>(free_b) * 8 / 10 / (2 * tree_bytes(p));
  resident = round_slots(resident);  // (the arena always runs as lock-step rounds)
  uint64_t n_slots = std::min<uint64_t>({G, resident, by_mem});
  if (cfg->max_concurrent_games) n_slots = std::min<uint64_t>(n_slots, cfg->max_concurrent_games);
  if (n_slots == 0) return fail(AZB_ERR_CAPACITY, "not enough device memory for one tree pair");
  if (shared) n_slots = 1;  // coach.rs:333-372: one pmcts, one nmcts, games one after the other
  TreePool pool;
  rc = pool.alloc(p, static_cast<uint32_t>(2 * n_slots));
  if (rc) return rc;
  GameStore gs;
  rc = gs.alloc(G, false);
  if (rc) return rc;
  RoundEngine eng;
  rc = eng.alloc(static_cast<uint32_t>(n_slots), p.num_threads);
  if (rc) return rc;
  RoundParams rp{};
  rp.p = p;
  rp.mode = kModeArena;
  rp.ev_kind[0] = eval_a;
  rp.ev_kind[1] = eval_b;
  const bool any_net = eval_a >= AZB_EVAL_NNET || eval_b >= AZB_EVAL_NNET;
  rp.plies_per_launch = any_net ? 0u : (cfg->plies_per_launch ? cfg->plies_per_launch : 2u);
  rp.sims_per_launch = any_net ? round_sim_budget(true) : 0u;
  rp.n_slots = static_cast<uint32_t>(n_slots);
  rp.leaf_cap = eng.leaf_cap;
  rp.n_games = static_cast<uint32_t>(G);
  rp.half = static_cast<uint32_t>(half);
  rp.k_open = k_open;
  rp.shared = shared ? 1u : 0u;
  rp.first_game_id = o.first_game_id;
  azb_nnet* nets[2] = {net_a, net_b};
  EventPair ev;
  AZB_CUDA(ev.create());
  AZB_CUDA(cudaEventRecord(ev.e0));
  uint64_t launches = 0;
  uint64_t nn_positions = 0, nn_cache_hits = 0;
  rc = eng.run(rp, pool.pools, gs, nets, &launches, &nn_positions, &nn_cache_hits);
  if (rc) return rc;
  AZB_CUDA(cudaEventRecord(ev.e1));
  AZB_CUDA(cudaEventSynchronize(ev.e1));
  float ms = 0.0f;
  AZB_CUDA(cudaEventElapsedTime(&ms, ev.e0, ev.e1));
  std::vector<int8_t> res(G);
  std::vector<uint32_t> h_err(G), h_stats(G * 8), h_plies(G);
  AZB_CUDA(cudaMemcpy(res.data(), gs.arena_result.p, G, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_err.data(), gs.error.p, G * 4, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_stats.data(), gs.stats.p, G * 32, cudaMemcpyDeviceToHost));
  AZB_CUDA(cudaMemcpy(h_plies.data(), gs.plies.p, G * 4, cudaMemcpyDeviceToHost));
  azb_selfplay_stats s{};
  for (uint64_t i = 0; i < G; ++i) {
    if (h_err[i]) return capacity_error(h_err[i]);
    const int win_cond = i < half ? 1 : -1;  // arena.rs:80-81
    if (res[i] == win_cond) out_counts[0]++;
    else if (res[i] == -win_cond) out_counts[1]++;
    else out_counts[2]++;
    s.plies += h_plies[i];
    if (shared && i + 1 < G) continue;  // shared trees: the counters are cumulative, the last game carries the totals
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
  s.device_ms = ms;
  s.launches = launches;
  s.nn_positions = nn_positions;
  s.nn_cache_hits = nn_cache_hits;
  if (results) std::memcpy(results, res.data(), G);
  if (actions) AZB_CUDA(cudaMemcpy(actions, gs.actions.p, G * kTraceStride, cudaMemcpyDeviceToHost));
  if (root_counts) AZB_CUDA(cudaMemcpy(root_counts, gs.counts.p, G * kTraceStride * 14, cudaMemcpyDeviceToHost));
  if (plies) std::memcpy(plies, h_plies.data(), G * 4);
  if (stats) *stats = s;
  return AZB_OK;
}

int azb_arena_play_games(const azb_config* cfg, uint64_t num, int32_t eval_a, int32_t eval_b, azb_nnet* net_a,
                         azb_nnet* net_b, uint32_t k_open, uint64_t out_counts[3], int8_t* results,
                         azb_selfplay_stats* stats) {
  const azb_arena_opts o{k_open, 0u, 0ull};
  return azb_arena_play_games_ex(cfg, num, eval_a, eval_b, net_a, net_b, &o, out_counts, results, nullptr, nullptr, nullptr,
                                 stats);
}


}  // extern "C"

#include "learn.cuh"
#include "dist.cuh"

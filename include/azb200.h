/* azb200.h — C ABI of the B200-native self-play engine (libazb200.so).
 *
 * Drop-in boundary for the self-play hot path of AnimatedRNG/alphazero-rs.  Every entry
 * point cites the reference interface it replaces (paths relative to /root/reference).
 * Plain pointers and sizes only; no torch / C++ types.  All functions return 0 on success
 * and a negative azb_status otherwise (the reference panics instead: unwrap/assert!);
 * azb_last_error() gives the message for the calling thread.  Handles are opaque, created
 * and destroyed by the library, and not thread-safe.  Bulk outputs go into caller-owned
 * buffers.  There is NO CPU fallback: without a CUDA device every compute call fails with
 * AZB_ERR_CUDA.
 */
#ifndef AZB200_H
#define AZB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum azb_status {
  AZB_OK = 0,
  AZB_ERR_INVALID = -1,   /* bad argument (reference: assert!/panic!) */
  AZB_ERR_CUDA = -2,      /* CUDA runtime error, or no device */
  AZB_ERR_CAPACITY = -3,  /* node pool / transposition table / sample buffer overflow
                             (reference: assert!(idx < self.buf.len()), src/node.rs:237) */
  AZB_ERR_UNSUPPORTED = -4
} azb_status;

/* Quirk flags: set = the LITERAL behaviour of the reference. */
#define AZB_Q1_WIN_RANGE_LITERAL 1u    /* connect_four_game.rs:114,129 */
#define AZB_Q2_BACKUP_NO_ALTERNATE 2u  /* src/async_mcts.rs:353,361-370 */
#define AZB_Q3_POS_BACKUP_PLUS_ONE 4u  /* src/node.rs:85-91 */
#define AZB_Q4_VLABEL_LITERAL 8u       /* src/coach.rs:146-153 */
#define AZB_PROFILE_REFERENCE 15u
#define AZB_PROFILE_SANE 0u

/* Leaf evaluators standing behind trait NNet::predict (src/nnet.rs:40-44). */
#define AZB_EVAL_UNIFORM 0 /* examples/connect_four.rs:26-42 DumbConnectFourNnet, fused inline */
#define AZB_EVAL_HASH 1    /* deterministic integer-hash evaluator (tests), fused inline */
#define AZB_EVAL_NNET 2    /* batched leaf evaluation by an azb_nnet (lock-step rounds) */

const char* azb_last_error(void);
/* Give the calling thread's evaluation cache (network rounds: 2 x 2^25 entries, ~2.7 GB of device memory, kept between
 * calls) back to the device; the next network run allocates it again. */
int azb_release_caches(void);
/* Number of visible CUDA devices (0 when there is none; never fails). */
int azb_device_count(void);

/* Page-locked host buffers for the bulk outputs (samples): device->host copies into them run at
 * full PCIe rate.  Ordinary malloc'ed memory works everywhere too, just slower. */
int azb_host_alloc(size_t bytes, void** out);
int azb_host_free(void* p);

/* ---------------------------------------------------------------------------------------
 * Game = ConnectFour.  trait Game, src/game.rs:10-28; the only implementation is
 * examples/connect_four_lib/connect_four_game.rs.  State POD mirrors its struct (:18-23)
 * without the redundant `heights`: s[row][col], row 0 = top (:99), cells in {-1,0,+1}.
 * All calls are batched over n states: host buffers in, one bitboard kernel, host buffers
 * out.
 * ------------------------------------------------------------------------------------- */
#pragma pack(push, 1)
typedef struct azb_c4_state {
  int8_t s[6][7];
  int8_t me;
} azb_c4_state; /* 43 bytes */
#pragma pack(pop)

#define AZB_C4_ACTIONS 7
#define AZB_C4_FEATURES 84 /* 2*6*7 */

/* Game::get_init_board — connect_four_game.rs:82-84 */
int azb_c4_init(azb_c4_state* out, size_t n);
/* Game::get_feature_shape — :86-88 → {2,6,7} */
int azb_c4_feature_shape(size_t out[3]);
/* Game::get_next_state(player, action) -> (state, -player) — :90-102 */
int azb_c4_next_state(const azb_c4_state* in, const int8_t* player, const uint8_t* action,
                      size_t n, azb_c4_state* out, int8_t* next_player);
/* Game::get_valid_moves — :104-109 → out[n][7] of 0/1 */
int azb_c4_valid_moves(const azb_c4_state* in, size_t n, uint8_t* out);
/* Game::get_game_ended(player) — :111-196; quirks bit Q1 selects the literal scan ranges */
int azb_c4_game_ended(const azb_c4_state* in, const int8_t* player, size_t n, uint32_t quirks,
                      float* out);
/* Game::get_canonical_form(player) — :198-203, repaired (F10): cells * player, me = +1 */
int azb_c4_canonical_form(const azb_c4_state* in, const int8_t* player, size_t n,
                          azb_c4_state* out);
/* Game::get_symmetries(pi) — :205-211 → out_states[n][2], out_pi[n][2][7] (identity, mirror) */
int azb_c4_symmetries(const azb_c4_state* in, const float* pi, size_t n, azb_c4_state* out_states,
                      float* out_pi);
/* Game::eval_heuristic — :214-216 (always 0) */
int azb_c4_eval_heuristic(const azb_c4_state* in, size_t n, float* out);
/* Game::to_features — :219-237, repaired (F11): out[n][2][6][7], ch0 = cells == me */
int azb_c4_to_features(const azb_c4_state* in, size_t n, float* out);

/* ---------------------------------------------------------------------------------------
 * Coach — src/coach.rs.  azb_config carries the 15 positional parameters of Coach::setup
 * (coach.rs:38-54) under the same names and in the same order, then this engine's own.
 * ------------------------------------------------------------------------------------- */
typedef struct azb_config {
  const char* checkpoint_directory;
  uint64_t mcts_reserve_size; /* node slots per tree (reference: 1,000,000); the pool is
                                 sized min(this, what num_sims can ever allocate) */
  float update_threshold;
  uint64_t temp_threshold;
  uint64_t max_history_length;
  uint64_t max_queue_length;
  uint64_t inference_batch_size; /* ignored: a round evaluates every pending leaf (F16) */
  uint64_t num_episode_threads;  /* ignored: every game of a call runs concurrently */
  uint64_t num_arena_games;
  uint64_t num_iters;
  uint64_t num_eps;
  uint64_t num_sims;
  uint64_t num_sim_threads; /* 1 = deterministic mode (one simulation in flight per tree; every bit-exact parity claim);
                             * K = 2..8 = tree-parallel search with virtual loss (async_mcts.rs:191-217): waves of K
                             * simulations per tree in one fixed interleaving, reproduced bit for bit by the oracle's wave
                             * mode; with the network evaluator a wave sends up to K leaves per tree into a round's batch.
                             * num_sims must be a multiple of it (async_mcts.rs:192) */
  uint64_t max_depth;
  int32_t cpuct;
  /* engine additions */
  uint32_t quirks;   /* AZB_Q* bits */
  uint64_t seed;     /* Philox-4x32-10 key; stream = (seed, global game id, ply) */
  int32_t evaluator; /* AZB_EVAL_* */
  int32_t device;    /* CUDA ordinal */
  uint64_t max_concurrent_games; /* trees resident in HBM at once; 0 = as many as requested */
  uint32_t schedule;         /* fused evaluators: 0/1 = one persistent kernel (a warp plays a whole game),
                                2 = lock-step rounds (live games re-dealt over the SMs every
                                plies_per_launch plies); the NNET evaluator always runs in rounds */
  uint32_t plies_per_launch; /* 0 = default (2) */
} azb_config;

/* The reference's example parameters (examples/connect_four.rs:55-71), profile "sane". */
void azb_config_default(azb_config* cfg);

typedef struct azb_coach azb_coach;

/* Coach::setup — coach.rs:38-102.  When checkpoint_directory is non-NULL and holds a `<n>.examples` file, the
 * newest one becomes the history (coach.rs:55-81); a missing directory is created (:78-80). */
int azb_coach_setup(const azb_config* cfg, azb_coach** out);
int azb_coach_destroy(azb_coach* c);

/* Per-call counters (device-side, summed over games). */
typedef struct azb_selfplay_stats {
  uint64_t games, plies, samples;
  uint64_t sims, levels, expansions, terminal_hits, dup_links, evals;
  uint64_t blocks_used_max, owners_max; /* pool high-water marks over trees */
  double device_ms;                     /* CUDA-event time of the self-play kernels */
  uint64_t launches;                    /* kernels launched by the call */
  uint64_t trees_resident;              /* games (trees) in flight at once */
  uint64_t nn_positions;                /* positions that went through the network(s): <= evals, because a position several
                                           trees ask for in the same round is evaluated once (0 for the fused evaluators) */
  uint64_t nn_cache_hits;               /* evaluations answered from the call's evaluation cache (positions the same network
                                           evaluated in an earlier round of this call): no network row, no suspension */
} azb_selfplay_stats;

/* Coach::execute_episode over n_games concurrent games — coach.rs:104-157 and the episode
 * fan-out coach.rs:241-272.  Game g uses Philox stream (seed, first_game_id + g).  Games run
 * entirely on the device; finished samples stay there until azb_coach_export_samples.
 * Per-game traces (optional, may be NULL): actions[n_games][64] (0xFF padded),
 * root_counts[n_games][64][7], plies[n_games], final_r[n_games], final_player[n_games]. */
int azb_coach_self_play(azb_coach* c, uint64_t n_games, uint64_t first_game_id,
                        azb_selfplay_stats* stats);
/* The same call in two halves: _begin launches (the persistent kernel of the fused evaluators runs on the coach's own
 * stream and _begin returns at once; network rounds complete inside _begin), _end waits and fills the statistics.  Used in
 * turn on two coaches — begin(A), begin(B), end(A), export(A), begin(A), end(B), ... — consecutive batches overlap on the
 * device: the warps that a batch's finished games vacate are taken by the next batch instead of idling until the batch's
 * longest game ends (the reference's rayon pool has no such barrier either: a worker that finishes an episode takes the
 * next one, coach.rs:241-272).  A coach holds one call in flight; its results (traces, samples) are those of the last
 * _end. */
int azb_coach_self_play_begin(azb_coach* c, uint64_t n_games, uint64_t first_game_id);
int azb_coach_self_play_end(azb_coach* c, azb_selfplay_stats* stats);
/* Device time across pipelined calls (CUDA events): _span_mark(first) records an event on that coach's stream before its
 * next _begin; _span_ms(first, last, &ms) = from that mark to the end of `last`'s most recent call. */
int azb_coach_span_mark(azb_coach* c);
int azb_coach_span_ms(azb_coach* first, azb_coach* last, double* ms);
int azb_coach_traces(azb_coach* c, uint8_t* actions, uint16_t* root_counts, uint32_t* plies,
                     float* final_r, int8_t* final_player);
/* Diagnostic (persistent schedule, env AZB200_PLY_TIMES=1): ns[n_games][64] = %globaltimer at the end
 * of each ply of each game, 0 where no ply was played. */
int azb_coach_ply_times(azb_coach* c, uint64_t* ns);
/* Number of training samples the last self-play call produced (2 per ply: coach.rs:130-135). */
int azb_coach_num_samples(azb_coach* c, uint64_t* n);
/* SOATrainingSamples (src/nnet.rs:33): boards[n][2][6][7], pis[n][7], vs[n], ordered by game,
 * ply, symmetry (coach.rs:132-135,243-270).  `capacity` is in samples. */
int azb_coach_export_samples(azb_coach* c, float* boards, float* pis, float* vs, uint64_t capacity,
                             uint64_t* n_written);

/* ---------------------------------------------------------------------------------------
 * NNet — src/nnet.rs:35-45.  One handle = one model (the reference addresses models by
 * model_id inside one NNet; here every model is its own handle, model_id is accepted and
 * ignored).  The architecture is this engine's own (the reference has no working network,
 * SURVEY §0.5): conv3x3(2->128)+ReLU, `blocks` residual blocks of two conv3x3(128->128),
 * policy head conv1x1(128->2)+ReLU+FC(84->7)+softmax, value head conv1x1(128->1)+ReLU+
 * FC(42->64)+ReLU+FC(64->1)+tanh; BatchNorm folded into the conv bias.
 * ------------------------------------------------------------------------------------- */
#define AZB_NNET_BF16_TC 0 /* bf16 weights/activations, fp32 accumulate, tcgen05 tensor cores.  Stated tolerance against
                             * the fp32 path of the same weights (13 layers of bf16 rounding): |pi - pi32| <= 5e-2 and
                             * |v - v32| <= 5e-2 per element, mean absolute error <= 2e-3 (pi) / 4e-3 (v); every tower layer
                             * is within 1 bf16 ulp of a float64 convolution of the same bf16 inputs
                             * (tests/test_nnet_gpu.py, tests/test_train_blocks_gpu.py).  Results do not depend on the batch. */
#define AZB_NNET_FP32 1    /* fp32 reference path on CUDA cores (pins the numerics) */
typedef struct azb_nnet_config {
  int32_t device;
  int32_t blocks;    /* residual blocks (6) */
  int32_t precision; /* AZB_NNET_* */
  int32_t reserved;
  uint64_t seed;     /* He-normal initialisation (weights ~ N(0, 2/fan_in), biases 0) */
} azb_nnet_config;
typedef struct azb_nnet azb_nnet;
/* NNet::new(checkpoint) — nnet.rs:36: random-init from cfg->seed; azb_nnet_load restores a checkpoint */
int azb_nnet_create(const azb_nnet_config* cfg, azb_nnet** out);
int azb_nnet_destroy(azb_nnet* n);
/* NNet::predict(boards[B,2,6,7], model_id) -> (pi[B,7] probabilities, v[B]) — nnet.rs:40-44.
 * boards are the 0/1 feature planes of Game::to_features. */
int azb_nnet_predict(azb_nnet* n, const float* boards, size_t batch, size_t model_id, float* pi, float* v);
/* Flat fp32 parameter vector (layout: alphazero-rs_b200/csrc/nnet.cuh NetLayout). */
int azb_nnet_num_params(azb_nnet* n, uint64_t* count);
int azb_nnet_get_params(azb_nnet* n, float* out, uint64_t capacity);
int azb_nnet_set_params(azb_nnet* n, const float* in, uint64_t count);
/* Diagnostic: device-only timing of `iters` forward passes over `batch` synthetic positions that are
 * already resident in HBM (CUDA events on the launching stream).  ms_per_pass is the mean. */
int azb_nnet_benchmark(azb_nnet* n, uint64_t batch, uint32_t iters, double* ms_per_pass);
/* NNet::train (src/nnet.rs:38: train((boards[N,2,6,7], pis[N,7], vs[N]), prev_id, id)).  Loss and optimiser as the
 * reference's connect_four_net.py:102-112: softmax cross-entropy on pi + mean squared error on v (means over the
 * batch), Adam.  Mixed precision: fp32 master parameters and gradients, bf16 tower on the tensor cores.
 *   azb_nnet_train_begin : forward (every layer kept), loss -> loss_out[2] = {policy, value}, backward; the gradients stay on
 *                          the device;
 *   azb_nnet_grads / azb_nnet_set_grads / azb_nnet_grads_device : read / replace them on the host, or get the device
 *                          pointer (count = azb_nnet_num_params): the seam where a data-parallel caller all-reduces
 *                          (NCCL in place on the device, or gloo on the host) and divides by the number of ranks;
 *   azb_nnet_train_apply : one Adam step, then every derived form of the weights is rebuilt on the device;
 *   azb_nnet_train       : begin + apply. */
typedef struct azb_train_config {
  float lr, beta1, beta2, eps; /* reference: lr 1e-3; TF defaults 0.9, 0.999, 1e-8 */
} azb_train_config;
int azb_nnet_train_begin(azb_nnet* n, const float* boards, const float* pis, const float* vs, uint64_t count, float* loss_out);
int azb_nnet_grads(azb_nnet* n, float* out, uint64_t count);
int azb_nnet_set_grads(azb_nnet* n, const float* in, uint64_t count);
/* The gradient vector where it lives: a device pointer to count fp32 values (valid until the next train_begin), so that a
 * caller with NCCL can all-reduce in place over NVLink without a host round trip. */
int azb_nnet_grads_device(azb_nnet* n, void** ptr, uint64_t* count);
int azb_nnet_train_apply(azb_nnet* n, const azb_train_config* cfg);
int azb_nnet_train(azb_nnet* n, const float* boards, const float* pis, const float* vs, uint64_t count,
                   const azb_train_config* cfg, float* loss_out);
/* Building blocks of NNet::train (src/nnet.rs:38) on the tensor cores, exposed as hooks so that they
 * are pinned by tests before the training step exists.  Activations are fp32 [n_pos][42 cells][128 channels] on the
 * host (quantised to bf16 on the device); `layer` = 0 .. 2*blocks-1 indexes the tower's 3x3 convolutions.
 *   mode 0: out = ReLU(conv(x) + bias [+ residual])                         (the forward kernel itself)
 *   mode 1: out = (conv^T(x) [+ residual]) * (mask > 0)                     (backward data: dX from dZ)
 * residual and mask may be NULL.  Needs the default tensor-core tower (AZB200_TC_PAIR unset). */
int azb_nnet_conv_hook(azb_nnet* n, int32_t layer, int32_t mode, const float* x, const float* residual,
                       const float* mask, uint64_t n_pos, float* out);
/* dW[9 taps][128 ci][128 co] (fp32) = sum over rows of x[row shifted by the tap][ci] * dz[row][co]. */
int azb_nnet_wgrad_hook(azb_nnet* n, const float* x, const float* dz, uint64_t n_pos, float* dw);
/* The network that evaluates leaves when cfg.evaluator == AZB_EVAL_NNET (the NNet the
 * reference's inference thread owns, async_mcts.rs:125).  The coach does not own it. */
int azb_coach_set_nnet(azb_coach* c, azb_nnet* n);

/* ---------------------------------------------------------------------------------------
 * arena — src/arena.rs:7-99 with the two MCTS players of coach.rs:333-375 (temp 0, arg-max
 * of get_action_prob, a fresh tree pair per game).  num/2 games with A moving first, num/2
 * with B first; out_counts = {Win, Loss, Draw} of player A (arena.rs:54-59,86-92);
 * results[2*(num/2)] (optional) = play_game's return value per game (arena.rs:51).
 * eval_a / eval_b are AZB_EVAL_*; net_a / net_b are used when the kind is AZB_EVAL_NNET.
 * k_open random opening plies per game (0 = the reference's behaviour).
 * ------------------------------------------------------------------------------------- */
int azb_arena_play_games(const azb_config* cfg, uint64_t num, int32_t eval_a, int32_t eval_b, azb_nnet* net_a,
                         azb_nnet* net_b, uint32_t k_open, uint64_t out_counts[3], int8_t* results,
                         azb_selfplay_stats* stats);

/* The same match with its options spelled out and the per-game traces returned.
 *   shared_trees = 0: a fresh tree pair per game, all games concurrent (the device layout; every game of a seat order is
 *                     the same game unless k_open > 0, because the players are deterministic);
 *   shared_trees = 1: the reference's own layout, coach.rs:333-354 — pmcts / nmcts are created ONCE, every game of the
 *                     match searches and grows the same two trees, games strictly one after the other (arena.rs:82-93),
 *                     so later games see the statistics of the earlier ones.  One game in flight: a latency-bound mode
 *                     kept for parity with the reference, not for throughput.  cfg->mcts_reserve_size bounds each tree
 *                     (node.rs:237 asserts; here AZB_ERR_CAPACITY).
 *   first_game_id   : game i draws its opening plies from Philox stream (cfg->seed, first_game_id + i).
 * actions[G][64] (0xFF padded), root_counts[G][64][7] (searched plies only), plies[G] with G = 2*(num/2); each may be
 * NULL.  opts == NULL means {0, 0, 0}. */
typedef struct azb_arena_opts {
  uint32_t k_open;
  uint32_t shared_trees;
  uint64_t first_game_id;
} azb_arena_opts;
int azb_arena_play_games_ex(const azb_config* cfg, uint64_t num, int32_t eval_a, int32_t eval_b, azb_nnet* net_a,
                            azb_nnet* net_b, const azb_arena_opts* opts, uint64_t out_counts[3], int8_t* results,
                            uint8_t* actions, uint16_t* root_counts, uint32_t* plies, azb_selfplay_stats* stats);

/* ---------------------------------------------------------------------------------------
 * On-disk formats.
 *
 * `<checkpoint>/<iteration>.examples` — Coach::save_train_examples, coach.rs:159-167: bincode 1.3.1
 * (default options: little-endian, fixed-width integers, u64 lengths) of
 * VecDeque<VecDeque<TrainingSample>> (src/nnet.rs:22-27) with ndarray 0.13's serde form of an array
 * {v: u8 = 1, dim, data: seq}:
 *   u64 n_iterations; per iteration: u64 n_samples; per sample:
 *     board: u8 1 | u64 ndim=3 | u64 2,6,7 | u64 84 | 84 x f32      (IxDyn: dim is a length-prefixed seq)
 *     pi   : u8 1 | u64 7 | u64 7 | 7 x f32                         (Ix1: dim is a bare [usize; 1])
 *     v    : f32                                                     426 bytes per sample
 * Both crates are absent from /root/reference (Cargo.toml:13,24), so the byte layout is restated from
 * their published formats: parity unpinned by the reference, pinned by oracle/learn.hpp and an independent
 * Python parser in tests/.  Host-only code: these calls work without a CUDA device.
 * The reader accepts the declared board shape [2,6,7] and the literal to_features shape [6,7,2] (F11), which it
 * transposes into channel-first planes; any other shape is rejected.
 * ------------------------------------------------------------------------------------- */
/* counts[n_iters] samples per history entry (oldest first); boards/pis/vs hold sum(counts) samples. */
int azb_examples_write(const char* path, uint64_t n_iters, const uint64_t* counts, const float* boards,
                       const float* pis, const float* vs);
/* Header pass: n_iters and, when counts != NULL, up to cap_iters per-iteration sample counts. */
int azb_examples_stat(const char* path, uint64_t* n_iters, uint64_t* counts, uint64_t cap_iters,
                      uint64_t* n_samples);
/* Data pass into caller buffers of cap_samples samples. */
int azb_examples_read(const char* path, float* boards, float* pis, float* vs, uint64_t cap_samples);
/* The most recent `<n>.examples` of a directory (Coach::setup, coach.rs:55-72, restricted to *.examples
 * so that weight files may share the directory): iteration number, or AZB_ERR_INVALID when none. */
int azb_examples_latest(const char* checkpoint_directory, uint64_t* iteration);

/* Weight checkpoints `<checkpoint>/<model_id>.azbw` (the reference's python_nnet.rs:76-79 saves
 * `<model_id>.pth.tar` through TF; this engine's own container): "AZBW" | u32 version=1 | u32 blocks |
 * u32 has_adam | u64 n_params | u64 adam_t | n_params f32 [| m | v]. */
int azb_nnet_save(azb_nnet* n, const char* path);
int azb_nnet_load(azb_nnet* n, const char* path);
/* dst <- src: parameters, Adam state and every derived device form (NNet::train's previous_model_id ->
 * model_id hand-over, src/nnet.rs:38). */
int azb_nnet_copy(azb_nnet* dst, azb_nnet* src);

/* ---------------------------------------------------------------------------------------
 * Coach::learn — coach.rs:169-396: num_iters iterations of
 *   self-play (num_eps games, network model_id)            coach.rs:241-272
 *   keep the newest max_queue_length samples               coach.rs:274-277
 *   history window of max_history_length iterations        coach.rs:284-289
 *   save `<iteration>.examples`                            coach.rs:291-293
 *   shuffle, AOS->SOA, train model_id -> model_id + 1      coach.rs:295-331
 *   arena new vs previous, num_arena_games, temp 0         coach.rs:333-375
 *   accept iff nwins/(nwins+pwins) >= update_threshold     coach.rs:383-390
 * The two models are double-buffered device networks; history lives in the coach (and is resumed from
 * checkpoint_directory by azb_coach_setup when a `<n>.examples` exists there, coach.rs:55-81).
 * Deviations, stated: the shuffle is a Fisher-Yates walk driven by Philox (seed, iteration) instead of
 * rand 0.7's SmallRng; the training schedule is `epochs` Adam steps over consecutive batch_size slices
 * of the shuffled list (connect_four_net.py:13-14,135-140 draws its minibatches with replacement);
 * episode e of iteration i uses game id i*num_eps + e (the reference clones one rng into every
 * episode, Q5, which makes the episodes of an iteration identical).
 * ------------------------------------------------------------------------------------- */
typedef struct azb_learn_config {
  uint32_t epochs;          /* Adam steps per iteration (connect_four_net.py:13: 10); 0 = one pass over the window */
  uint32_t batch_size;      /* connect_four_net.py:14: 64 */
  azb_train_config adam;    /* default lr 1e-4 (the reference's 1e-3, connect_four_net.py:21, assumes BatchNorm; see learn.cuh), 0.9, 0.999, 1e-8 */
  uint32_t arena_k_open;    /* random opening plies of the gating games.  Default 4: with concurrent fresh-tree games and
                             * deterministic players, 0 makes every game of a seat order the SAME game (the gate would be
                             * decided by two distinct games).  The reference gets its variety from the tree pair it keeps
                             * across the games: arena_shared_trees = 1, arena_k_open = 0 is its literal behaviour. */
  uint32_t skip_first_play; /* Coach::learn's skip_first_play, coach.rs:172,240 */
  uint32_t save_files;      /* 1: write <iteration>.examples and <model_id>.azbw into checkpoint_directory */
  uint32_t arena_shared_trees; /* 1: pmcts / nmcts persist across the gating games, played one after the other
                                * (coach.rs:333-372); 0 (default): fresh tree pair per game, all games concurrent */
} azb_learn_config;
void azb_learn_config_default(azb_learn_config* lc);

typedef struct azb_learn_report {
  uint64_t iteration;
  uint64_t model_id_before, model_id_after;
  uint64_t games, samples_played, samples_kept; /* this iteration's self-play, before / after the queue trim */
  uint64_t history_iterations, history_samples; /* the window the network was trained on */
  uint64_t train_steps;
  float loss_first[2], loss_last[2];            /* {policy, value} of the first / last step */
  uint64_t nwins, pwins, draws;                 /* arena: candidate (new) vs current (previous) */
  int32_t accepted;
  int32_t reserved;
  double selfplay_ms, train_ms, arena_ms;       /* host wall-clock of the three phases */
} azb_learn_report;

/* Resume: a coach whose setup loaded `<n>.examples` numbers its iterations from n + 1 (game ids, shuffles and file names
 * go on; the reference restarts at 0 and, seeding from entropy, plays new games — with this engine's fixed Philox streams a
 * restart at 0 would replay the games the window already holds).  Weights restart from `0.azbw` (model ids restart).
 * net_cfg: architecture and seed of model 0 (NNet::new(checkpoint), async_mcts.rs:125).  With save_files set, a
 * `0.azbw` already in checkpoint_directory is loaded instead of the random init, and every trained candidate is
 * written as `<model_id + 1>.azbw`.  reports[cap_reports] receives one entry per iteration; *final_net (optional)
 * receives the accepted model, owned by the caller afterwards (azb_nnet_destroy). */
int azb_coach_learn(azb_coach* c, const azb_nnet_config* net_cfg, const azb_learn_config* lc,
                    azb_learn_report* reports, uint64_t cap_reports, uint64_t* n_reports, azb_nnet** final_net);
/* The same loop, data parallel over one process per GPU (BASELINE config 5).  The library has no
 * communicator of its own: the caller hands in its all-reduce (NCCL through torch.distributed in the ctypes mirror,
 * ncclAllReduce in a Rust/C++ host).  Per iteration: rank r self-plays a contiguous share of the num_eps games (game
 * ids iteration*num_eps + first_r ..., NO collective) and keeps their samples, with its share ceil(max_queue_length /
 * world) of the queue; `batch_size` is the GLOBAL batch, each rank trains on batch_size / world of its own samples per
 * step and the fp32 gradient vector is summed in place ON THE DEVICE by allreduce_sum_f32_device, then divided by
 * world; the num_arena_games / 2 seat-order pairs are split over the ranks and the three counters summed by
 * allreduce_sum_u64_host, so that every rank takes the same decision on bit-identical models.  Every rank needs its own
 * checkpoint_directory (its share of the history is written there).  Callbacks return 0 on success. */
typedef struct azb_dist {
  uint32_t rank, world;
  int (*allreduce_sum_f32_device)(void* device_ptr, uint64_t count, void* user);
  int (*allreduce_sum_u64_host)(uint64_t* host_ptr, uint64_t count, void* user);
  void* user;
} azb_dist;
int azb_coach_learn_dist(azb_coach* c, const azb_nnet_config* net_cfg, const azb_learn_config* lc, const azb_dist* dist,
                         azb_learn_report* reports, uint64_t cap_reports, uint64_t* n_reports, azb_nnet** final_net);
/* The library's own communicator (NCCL, loaded with dlopen at first use: AZB200_NCCL_LIB, default libnccl.so.2) for hosts
 * that bring none — a Rust host needs neither torch nor torchrun.  One process per GPU: rank 0 calls azb_dist_unique_id and
 * hands the 128 bytes to the other ranks by whatever means the host has (a file, MPI, a socket); every rank calls
 * azb_dist_init(id, rank, world, device); azb_dist_make fills an azb_dist whose callbacks run ncclAllReduce on that
 * communicator (in place on the device gradient vector; the u64 counters staged through a device buffer), ready for
 * azb_coach_learn_dist.  azb_dist_allreduce_f64 reduces a few host doubles (timings, counters) with the same communicator.
 * The reference has no counterpart: it is single-process (rayon threads, crossbeam channels). */
#define AZB_DIST_ID_BYTES 128
enum { AZB_DIST_SUM = 0, AZB_DIST_MAX = 1, AZB_DIST_MIN = 2 };
typedef struct azb_comm azb_comm;
int azb_dist_unique_id(uint8_t out[AZB_DIST_ID_BYTES]);
int azb_dist_init(const uint8_t id[AZB_DIST_ID_BYTES], uint32_t rank, uint32_t world, int32_t device, azb_comm** out);
int azb_dist_destroy(azb_comm* c);
int azb_dist_make(azb_comm* c, azb_dist* out);
int azb_dist_allreduce_f64(azb_comm* c, double* values, uint64_t count, int32_t op);
/* Self-play on several GPUs of one box through ONE call (the episode fan-out of coach.rs:241-272 over devices): a host
 * thread per listed device sets up its own coach (and, for AZB_EVAL_NNET, its own replica of the network created from
 * net_cfg), plays games [first_game_id + d * games_per_device, ...) and writes stats[d].  No collective on the path.
 * wall_ms (optional) = host wall clock of the whole call.  net_cfg may be NULL for fused evaluators. */
int azb_coach_self_play_multi(const azb_config* cfg, const azb_nnet_config* net_cfg, const int32_t* devices,
                              uint32_t n_devices, uint64_t games_per_device, uint64_t first_game_id,
                              azb_selfplay_stats* stats, double* wall_ms);
/* The coach's sample history (struct Coach.history, coach.rs:19): entry counts, then the data. */
int azb_coach_history_stat(azb_coach* c, uint64_t* n_iters, uint64_t* counts, uint64_t cap_iters, uint64_t* n_samples);
int azb_coach_history_export(azb_coach* c, float* boards, float* pis, float* vs, uint64_t cap_samples);
/* Coach::save_train_examples(iteration, checkpoint) — coach.rs:159-167 (the reference joins an absolute
 * "/<n>.examples", which lands in the filesystem root; the intended `<checkpoint>/<n>.examples` is written). */
int azb_coach_save_train_examples(azb_coach* c, uint64_t iteration, const char* checkpoint_directory);
/* Replace the history with the contents of a file (what Coach::setup does with the newest one). */
int azb_coach_load_train_examples(azb_coach* c, const char* path);
/* The reference's pure decisions, exposed so that they can be checked without a device:
 * accept rule coach.rs:383-390; shuffle permutation used by azb_coach_learn (perm[n]). */
int azb_learn_accept(uint64_t nwins, uint64_t pwins, float update_threshold);
int azb_learn_shuffle_perm(uint64_t seed, uint64_t iteration, uint64_t n, uint64_t* perm);

/* ---------------------------------------------------------------------------------------
 * AsyncMcts test hooks — src/async_mcts.rs and src/node.rs are private modules of the
 * reference (src/lib.rs:6,10); these exist so their behaviour can be checked directly.
 * An azb_mcts is n_trees independent trees, each created on the initial board
 * (AsyncMcts::default, async_mcts.rs:27-48).
 * ------------------------------------------------------------------------------------- */
typedef struct azb_mcts azb_mcts;
int azb_mcts_create(const azb_config* cfg, uint64_t n_trees, azb_mcts** out);
int azb_mcts_destroy(azb_mcts* m);
/* AsyncMcts::get_action_prob — async_mcts.rs:74-115: for every tree i run cfg.num_sims
 * simulations from the canonical state states[i] and return the root child visit counts
 * counts[n_trees][7] and pi[n_trees][7] (temp 0 → one-hot, ties to the highest action). */
int azb_mcts_get_action_prob(azb_mcts* m, const azb_c4_state* states, float temp, uint16_t* counts,
                             float* pi);
/* Raw packed counter (src/node.rs:17, 0xWWWWWWWWNNNNVVVV) of the node owning states[i];
 * 0 when the state is not in tree i. */
int azb_mcts_counter_of(azb_mcts* m, const azb_c4_state* states, uint64_t* counters);
/* stats[n_trees][8]: sims, levels, expansions, terminal_hits, dup_links, evals,
 * blocks_used, owners (= NodeStore.seen.len()). */
int azb_mcts_stats(azb_mcts* m, uint64_t* stats);
/* Dump tree `tree`: one row per unique state: 49-bit key, raw counter, terminal value e,
 * prior vector and a has-policy flag.  Returns rows through n_rows (rows beyond cap are
 * counted but not written). */
int azb_mcts_dump(azb_mcts* m, uint64_t tree, uint64_t cap, uint64_t* keys, uint64_t* counters,
                  float* e, float* p7, uint8_t* has_p, uint64_t* n_rows);

/* Diagnostic: compares the level loop's slow-path-free f32 division / square root with the
 * IEEE-rounded CUDA intrinsics, exhaustively over the integer denominators / visit counts they
 * are used on.  mismatches[4] = {reciprocal, sqrt, division, cached Q}; all must be 0. */
int azb_selftest_arith(uint64_t mismatches[4]);

#ifdef __cplusplus
}
#endif
#endif /* AZB200_H */

//! Raw FFI to `libazb200.so` — one declaration per entry point of `include/azb200.h`, in the header's order.
//! Every function returns 0 (`AZB_OK`) or a negative `azb_status`; `azb_last_error()` holds the message.
#![allow(non_camel_case_types)]

use std::os::raw::{c_char, c_int, c_void};

pub const AZB_OK: c_int = 0;
pub const AZB_ERR_INVALID: c_int = -1;
pub const AZB_ERR_CUDA: c_int = -2;
pub const AZB_ERR_CAPACITY: c_int = -3;
pub const AZB_ERR_UNSUPPORTED: c_int = -4;

pub const AZB_Q1_WIN_RANGE_LITERAL: u32 = 1;
pub const AZB_Q2_BACKUP_NO_ALTERNATE: u32 = 2;
pub const AZB_Q3_POS_BACKUP_PLUS_ONE: u32 = 4;
pub const AZB_Q4_VLABEL_LITERAL: u32 = 8;
pub const AZB_PROFILE_REFERENCE: u32 = 15;
pub const AZB_PROFILE_SANE: u32 = 0;

pub const AZB_EVAL_UNIFORM: i32 = 0;
pub const AZB_EVAL_HASH: i32 = 1;
pub const AZB_EVAL_NNET: i32 = 2;

pub const AZB_NNET_BF16_TC: i32 = 0;
pub const AZB_NNET_FP32: i32 = 1;

pub const AZB_C4_ACTIONS: usize = 7;
pub const AZB_C4_FEATURES: usize = 84;

/// `ConnectFourGame` without the redundant `heights` (connect_four_game.rs:18-23): `s[row][col]`, row 0 = top.
#[repr(C, packed)]
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
pub struct azb_c4_state {
    pub s: [[i8; 7]; 6],
    pub me: i8,
}

/// The 15 positional parameters of `Coach::setup` (coach.rs:38-54), then the engine's own.
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct azb_config {
    pub checkpoint_directory: *const c_char,
    pub mcts_reserve_size: u64,
    pub update_threshold: f32,
    pub temp_threshold: u64,
    pub max_history_length: u64,
    pub max_queue_length: u64,
    pub inference_batch_size: u64,
    pub num_episode_threads: u64,
    pub num_arena_games: u64,
    pub num_iters: u64,
    pub num_eps: u64,
    pub num_sims: u64,
    pub num_sim_threads: u64,
    pub max_depth: u64,
    pub cpuct: i32,
    pub quirks: u32,
    pub seed: u64,
    pub evaluator: i32,
    pub device: i32,
    pub max_concurrent_games: u64,
    pub schedule: u32,
    pub plies_per_launch: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct azb_selfplay_stats {
    pub games: u64,
    pub plies: u64,
    pub samples: u64,
    pub sims: u64,
    pub levels: u64,
    pub expansions: u64,
    pub terminal_hits: u64,
    pub dup_links: u64,
    pub evals: u64,
    pub blocks_used_max: u64,
    pub owners_max: u64,
    pub device_ms: f64,
    pub launches: u64,
    pub trees_resident: u64,
    pub nn_positions: u64,
    pub nn_cache_hits: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct azb_nnet_config {
    pub device: i32,
    pub blocks: i32,
    pub precision: i32,
    pub reserved: i32,
    pub seed: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct azb_train_config {
    pub lr: f32,
    pub beta1: f32,
    pub beta2: f32,
    pub eps: f32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct azb_learn_config {
    pub epochs: u32,
    pub batch_size: u32,
    pub adam: azb_train_config,
    pub arena_k_open: u32,
    pub skip_first_play: u32,
    pub save_files: u32,
    pub arena_shared_trees: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct azb_arena_opts {
    pub k_open: u32,
    pub shared_trees: u32,
    pub first_game_id: u64,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct azb_learn_report {
    pub iteration: u64,
    pub model_id_before: u64,
    pub model_id_after: u64,
    pub games: u64,
    pub samples_played: u64,
    pub samples_kept: u64,
    pub history_iterations: u64,
    pub history_samples: u64,
    pub train_steps: u64,
    pub loss_first: [f32; 2],
    pub loss_last: [f32; 2],
    pub nwins: u64,
    pub pwins: u64,
    pub draws: u64,
    pub accepted: i32,
    pub reserved: i32,
    pub selfplay_ms: f64,
    pub train_ms: f64,
    pub arena_ms: f64,
}

/// The communicator `azb_coach_learn_dist` reduces with: the caller's own callbacks, or the library's NCCL communicator
/// (`azb_dist_init` + `azb_dist_make`).  Both callbacks return 0 on success.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct azb_dist {
    pub rank: u32,
    pub world: u32,
    pub allreduce_sum_f32_device: Option<unsafe extern "C" fn(device_ptr: *mut c_void, count: u64, user: *mut c_void) -> c_int>,
    pub allreduce_sum_u64_host: Option<unsafe extern "C" fn(host_ptr: *mut u64, count: u64, user: *mut c_void) -> c_int>,
    pub user: *mut c_void,
}

#[repr(C)]
pub struct azb_coach {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct azb_nnet {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct azb_mcts {
    _opaque: [u8; 0],
}
#[repr(C)]
pub struct azb_comm {
    _opaque: [u8; 0],
}
pub const AZB_DIST_ID_BYTES: usize = 128;
pub const AZB_DIST_SUM: i32 = 0;
pub const AZB_DIST_MAX: i32 = 1;
pub const AZB_DIST_MIN: i32 = 2;

extern "C" {
    pub fn azb_last_error() -> *const c_char;
    pub fn azb_device_count() -> c_int;
    pub fn azb_release_caches() -> c_int;
    pub fn azb_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn azb_host_free(p: *mut c_void) -> c_int;

    // trait Game, src/game.rs:10-28 (batched over n states)
    pub fn azb_c4_init(out: *mut azb_c4_state, n: usize) -> c_int;
    pub fn azb_c4_feature_shape(out: *mut usize) -> c_int;
    pub fn azb_c4_next_state(inp: *const azb_c4_state, player: *const i8, action: *const u8, n: usize, out: *mut azb_c4_state, next_player: *mut i8) -> c_int;
    pub fn azb_c4_valid_moves(inp: *const azb_c4_state, n: usize, out: *mut u8) -> c_int;
    pub fn azb_c4_game_ended(inp: *const azb_c4_state, player: *const i8, n: usize, quirks: u32, out: *mut f32) -> c_int;
    pub fn azb_c4_canonical_form(inp: *const azb_c4_state, player: *const i8, n: usize, out: *mut azb_c4_state) -> c_int;
    pub fn azb_c4_symmetries(inp: *const azb_c4_state, pi: *const f32, n: usize, out_states: *mut azb_c4_state, out_pi: *mut f32) -> c_int;
    pub fn azb_c4_eval_heuristic(inp: *const azb_c4_state, n: usize, out: *mut f32) -> c_int;
    pub fn azb_c4_to_features(inp: *const azb_c4_state, n: usize, out: *mut f32) -> c_int;

    // Coach, src/coach.rs
    pub fn azb_config_default(cfg: *mut azb_config);
    pub fn azb_coach_setup(cfg: *const azb_config, out: *mut *mut azb_coach) -> c_int;
    pub fn azb_coach_destroy(c: *mut azb_coach) -> c_int;
    pub fn azb_coach_self_play(c: *mut azb_coach, n_games: u64, first_game_id: u64, stats: *mut azb_selfplay_stats) -> c_int;
    pub fn azb_coach_self_play_begin(c: *mut azb_coach, n_games: u64, first_game_id: u64) -> c_int;
    pub fn azb_coach_self_play_end(c: *mut azb_coach, stats: *mut azb_selfplay_stats) -> c_int;
    pub fn azb_coach_span_mark(c: *mut azb_coach) -> c_int;
    pub fn azb_coach_span_ms(first: *mut azb_coach, last: *mut azb_coach, ms: *mut f64) -> c_int;
    pub fn azb_coach_traces(c: *mut azb_coach, actions: *mut u8, root_counts: *mut u16, plies: *mut u32, final_r: *mut f32, final_player: *mut i8) -> c_int;
    pub fn azb_coach_ply_times(c: *mut azb_coach, ns: *mut u64) -> c_int;
    pub fn azb_coach_num_samples(c: *mut azb_coach, n: *mut u64) -> c_int;
    pub fn azb_coach_export_samples(c: *mut azb_coach, boards: *mut f32, pis: *mut f32, vs: *mut f32, capacity: u64, n_written: *mut u64) -> c_int;

    // NNet, src/nnet.rs:35-45
    pub fn azb_nnet_create(cfg: *const azb_nnet_config, out: *mut *mut azb_nnet) -> c_int;
    pub fn azb_nnet_destroy(n: *mut azb_nnet) -> c_int;
    pub fn azb_nnet_predict(n: *mut azb_nnet, boards: *const f32, batch: usize, model_id: usize, pi: *mut f32, v: *mut f32) -> c_int;
    pub fn azb_nnet_num_params(n: *mut azb_nnet, count: *mut u64) -> c_int;
    pub fn azb_nnet_get_params(n: *mut azb_nnet, out: *mut f32, capacity: u64) -> c_int;
    pub fn azb_nnet_set_params(n: *mut azb_nnet, inp: *const f32, count: u64) -> c_int;
    pub fn azb_nnet_benchmark(n: *mut azb_nnet, batch: u64, iters: u32, ms_per_pass: *mut f64) -> c_int;
    pub fn azb_nnet_train_begin(n: *mut azb_nnet, boards: *const f32, pis: *const f32, vs: *const f32, count: u64, loss_out: *mut f32) -> c_int;
    pub fn azb_nnet_grads(n: *mut azb_nnet, out: *mut f32, count: u64) -> c_int;
    pub fn azb_nnet_set_grads(n: *mut azb_nnet, inp: *const f32, count: u64) -> c_int;
    pub fn azb_nnet_grads_device(n: *mut azb_nnet, ptr: *mut *mut c_void, count: *mut u64) -> c_int;
    pub fn azb_nnet_train_apply(n: *mut azb_nnet, cfg: *const azb_train_config) -> c_int;
    pub fn azb_nnet_train(n: *mut azb_nnet, boards: *const f32, pis: *const f32, vs: *const f32, count: u64, cfg: *const azb_train_config, loss_out: *mut f32) -> c_int;
    pub fn azb_nnet_conv_hook(n: *mut azb_nnet, layer: i32, mode: i32, x: *const f32, residual: *const f32, mask: *const f32, n_pos: u64, out: *mut f32) -> c_int;
    pub fn azb_nnet_wgrad_hook(n: *mut azb_nnet, x: *const f32, dz: *const f32, n_pos: u64, dw: *mut f32) -> c_int;
    pub fn azb_coach_set_nnet(c: *mut azb_coach, n: *mut azb_nnet) -> c_int;

    // arena, src/arena.rs:7-99
    pub fn azb_arena_play_games(cfg: *const azb_config, num: u64, eval_a: i32, eval_b: i32, net_a: *mut azb_nnet, net_b: *mut azb_nnet, k_open: u32, out_counts: *mut u64, results: *mut i8, stats: *mut azb_selfplay_stats) -> c_int;
    pub fn azb_arena_play_games_ex(cfg: *const azb_config, num: u64, eval_a: i32, eval_b: i32, net_a: *mut azb_nnet, net_b: *mut azb_nnet, opts: *const azb_arena_opts, out_counts: *mut u64, results: *mut i8, actions: *mut u8, root_counts: *mut u16, plies: *mut u32, stats: *mut azb_selfplay_stats) -> c_int;

    // the library's own NCCL communicator and the multi-GPU fan-out (no counterpart in the reference: SURVEY 2.2)
    pub fn azb_dist_unique_id(out: *mut u8) -> c_int;
    pub fn azb_dist_init(id: *const u8, rank: u32, world: u32, device: i32, out: *mut *mut azb_comm) -> c_int;
    pub fn azb_dist_destroy(c: *mut azb_comm) -> c_int;
    pub fn azb_dist_make(c: *mut azb_comm, out: *mut azb_dist) -> c_int;
    pub fn azb_dist_allreduce_f64(c: *mut azb_comm, values: *mut f64, count: u64, op: i32) -> c_int;
    pub fn azb_coach_self_play_multi(cfg: *const azb_config, net_cfg: *const azb_nnet_config, devices: *const i32, n_devices: u32, games_per_device: u64, first_game_id: u64, stats: *mut azb_selfplay_stats, wall_ms: *mut f64) -> c_int;

    // on-disk formats: coach.rs:55-81,159-167 and the weight checkpoints
    pub fn azb_examples_write(path: *const c_char, n_iters: u64, counts: *const u64, boards: *const f32, pis: *const f32, vs: *const f32) -> c_int;
    pub fn azb_examples_stat(path: *const c_char, n_iters: *mut u64, counts: *mut u64, cap_iters: u64, n_samples: *mut u64) -> c_int;
    pub fn azb_examples_read(path: *const c_char, boards: *mut f32, pis: *mut f32, vs: *mut f32, cap_samples: u64) -> c_int;
    pub fn azb_examples_latest(checkpoint_directory: *const c_char, iteration: *mut u64) -> c_int;
    pub fn azb_nnet_save(n: *mut azb_nnet, path: *const c_char) -> c_int;
    pub fn azb_nnet_load(n: *mut azb_nnet, path: *const c_char) -> c_int;
    pub fn azb_nnet_copy(dst: *mut azb_nnet, src: *mut azb_nnet) -> c_int;

    // Coach::learn, coach.rs:169-396
    pub fn azb_learn_config_default(lc: *mut azb_learn_config);
    pub fn azb_coach_learn(c: *mut azb_coach, net_cfg: *const azb_nnet_config, lc: *const azb_learn_config, reports: *mut azb_learn_report, cap_reports: u64, n_reports: *mut u64, final_net: *mut *mut azb_nnet) -> c_int;
    pub fn azb_coach_learn_dist(c: *mut azb_coach, net_cfg: *const azb_nnet_config, lc: *const azb_learn_config, dist: *const azb_dist, reports: *mut azb_learn_report, cap_reports: u64, n_reports: *mut u64, final_net: *mut *mut azb_nnet) -> c_int;
    pub fn azb_coach_history_stat(c: *mut azb_coach, n_iters: *mut u64, counts: *mut u64, cap_iters: u64, n_samples: *mut u64) -> c_int;
    pub fn azb_coach_history_export(c: *mut azb_coach, boards: *mut f32, pis: *mut f32, vs: *mut f32, cap_samples: u64) -> c_int;
    pub fn azb_coach_save_train_examples(c: *mut azb_coach, iteration: u64, checkpoint_directory: *const c_char) -> c_int;
    pub fn azb_coach_load_train_examples(c: *mut azb_coach, path: *const c_char) -> c_int;
    pub fn azb_learn_accept(nwins: u64, pwins: u64, update_threshold: f32) -> c_int;
    pub fn azb_learn_shuffle_perm(seed: u64, iteration: u64, n: u64, perm: *mut u64) -> c_int;

    // AsyncMcts test hooks (async_mcts.rs / node.rs are private modules of the reference)
    pub fn azb_mcts_create(cfg: *const azb_config, n_trees: u64, out: *mut *mut azb_mcts) -> c_int;
    pub fn azb_mcts_destroy(m: *mut azb_mcts) -> c_int;
    pub fn azb_mcts_get_action_prob(m: *mut azb_mcts, states: *const azb_c4_state, temp: f32, counts: *mut u16, pi: *mut f32) -> c_int;
    pub fn azb_mcts_counter_of(m: *mut azb_mcts, states: *const azb_c4_state, counters: *mut u64) -> c_int;
    pub fn azb_mcts_stats(m: *mut azb_mcts, stats: *mut u64) -> c_int;
    pub fn azb_mcts_dump(m: *mut azb_mcts, tree: u64, cap: u64, keys: *mut u64, counters: *mut u64, e: *mut f32, p7: *mut f32, has_p: *mut u8, n_rows: *mut u64) -> c_int;
    pub fn azb_selftest_arith(mismatches: *mut u64) -> c_int;
}

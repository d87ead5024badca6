//! Safe wrapper over `azb200-sys` with the shapes of the reference crate (alphazero-rs):
//!
//! * [`ConnectFourGame`] — the nine functions of `trait Game` (src/game.rs:10-28) on one state, each a call of the
//!   batched `azb_c4_*` kernel with n = 1, plus `*_batch` forms for callers that hold many states;
//! * [`B200Net`] — `trait NNet` (src/nnet.rs:35-45): `new`, `train`, `predict`;
//! * [`Coach`] — `setup` with the 15 positional parameters of coach.rs:38-54, `execute_episodes` (the fan-out of
//!   coach.rs:241-272 as one device call), `learn` (coach.rs:169-396), `save_train_examples` (coach.rs:159-167);
//! * [`arena::play_games`] — src/arena.rs:62-99 with the two MCTS players of coach.rs:333-375.
//!
//! The reference panics on every failure of this path (`unwrap`, `assert!`); here every call returns
//! `Result<_, Error>` carrying the status code and `azb_last_error()`.
//!
//! Uncompiled source: the build image has no Rust toolchain (see rust/README.md).

use std::ffi::{CStr, CString};
use std::path::Path;
use std::ptr;

use azb200_sys as sys;
use ndarray::{Array, Array1, Array2, ArrayD, ArrayView1, ArrayViewD, Ix1, IxDyn};

pub type F = f32; // src/game.rs:8
pub type Policy = Array<f32, Ix1>; // src/nnet.rs:17
pub type SOATrainingSamples = (ArrayD<F>, Array2<f32>, Array1<f32>); // src/nnet.rs:33 (owned instead of Arc)

#[derive(Debug, Clone)]
pub struct Error {
    pub code: i32,
    pub message: String,
}
impl std::fmt::Display for Error {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        write!(f, "azb200 error {}: {}", self.code, self.message)
    }
}
impl std::error::Error for Error {}
pub type Result<T> = std::result::Result<T, Error>;

fn check(rc: i32) -> Result<()> {
    if rc == sys::AZB_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(sys::azb_last_error()) }.to_string_lossy().into_owned();
    Err(Error { code: rc, message })
}

fn c_path<P: AsRef<Path>>(p: P) -> CString {
    CString::new(p.as_ref().to_string_lossy().as_bytes()).expect("path contains a NUL byte")
}

// -------------------------------------------------------------------------------------------------
// Game = ConnectFour (examples/connect_four_lib/connect_four_game.rs)
// -------------------------------------------------------------------------------------------------
#[derive(Clone, Copy, PartialEq, Eq, Hash, Debug)]
pub struct ConnectFourGame(pub sys::azb_c4_state);

impl std::fmt::Display for ConnectFourGame {
    fn fmt(&self, f: &mut std::fmt::Formatter) -> std::fmt::Result {
        let s = self.0.s;
        for row in s.iter() {
            for &c in row.iter() {
                write!(f, "{}", if c > 0 { 'X' } else if c < 0 { 'O' } else { '.' })?;
            }
            writeln!(f)?;
        }
        Ok(())
    }
}

impl ConnectFourGame {
    /// Game::get_init_board — connect_four_game.rs:82-84
    pub fn try_get_init_board() -> Result<Self> {
        let mut st = sys::azb_c4_state { s: [[0; 7]; 6], me: 1 };
        check(unsafe { sys::azb_c4_init(&mut st, 1) })?;
        Ok(ConnectFourGame(st))
    }
    /// Game::get_feature_shape — :86-88
    pub fn feature_shape() -> Vec<usize> {
        let mut out = [0usize; 3];
        unsafe { sys::azb_c4_feature_shape(out.as_mut_ptr()) };
        out.to_vec()
    }
    /// Game::get_next_state(player, action) -> (state, -player) — :90-102
    pub fn try_get_next_state(&self, player: i8, action: u8) -> Result<(Self, i8)> {
        let mut out = self.0;
        let mut next = 0i8;
        check(unsafe { sys::azb_c4_next_state(&self.0, &player, &action, 1, &mut out, &mut next) })?;
        Ok((ConnectFourGame(out), next))
    }
    /// Game::get_valid_moves — :104-109 (the player argument is ignored there too)
    pub fn try_get_valid_moves(&self, _player: i8) -> Result<Array<u8, Ix1>> {
        let mut out = [0u8; sys::AZB_C4_ACTIONS];
        check(unsafe { sys::azb_c4_valid_moves(&self.0, 1, out.as_mut_ptr()) })?;
        Ok(Array::from(out.to_vec()))
    }
    /// Game::get_game_ended(player) — :111-196; `quirks` bit Q1 selects the literal scan ranges
    pub fn try_get_game_ended(&self, player: i8, quirks: u32) -> Result<f32> {
        let mut out = 0f32;
        check(unsafe { sys::azb_c4_game_ended(&self.0, &player, 1, quirks, &mut out) })?;
        Ok(out)
    }
    /// Game::get_canonical_form(player) — :198-203 (repaired: cells * player)
    pub fn try_get_canonical_form(&self, player: i8) -> Result<Self> {
        let mut out = self.0;
        check(unsafe { sys::azb_c4_canonical_form(&self.0, &player, 1, &mut out) })?;
        Ok(ConnectFourGame(out))
    }
    /// Game::get_symmetries(pi) — :205-211: identity and the column mirror
    pub fn try_get_symmetries(&self, pi: ArrayView1<f32>) -> Result<Vec<(Self, Policy)>> {
        assert_eq!(pi.len(), sys::AZB_C4_ACTIONS);
        let pi_in: Vec<f32> = pi.iter().cloned().collect();
        let mut states = [self.0; 2];
        let mut pis = [0f32; 2 * sys::AZB_C4_ACTIONS];
        check(unsafe { sys::azb_c4_symmetries(&self.0, pi_in.as_ptr(), 1, states.as_mut_ptr(), pis.as_mut_ptr()) })?;
        Ok((0..2)
            .map(|k| (ConnectFourGame(states[k]), Array::from(pis[k * 7..k * 7 + 7].to_vec())))
            .collect())
    }
    /// Game::eval_heuristic — :214-216
    pub fn try_eval_heuristic(&self) -> Result<f32> {
        let mut out = 0f32;
        check(unsafe { sys::azb_c4_eval_heuristic(&self.0, 1, &mut out) })?;
        Ok(out)
    }
    /// Game::to_features — :219-237 (repaired: [2,6,7], channel 0 = cells of `me`)
    pub fn try_to_features(&self) -> Result<ArrayD<F>> {
        let mut out = vec![0f32; sys::AZB_C4_FEATURES];
        check(unsafe { sys::azb_c4_to_features(&self.0, 1, out.as_mut_ptr()) })?;
        Ok(ArrayD::from_shape_vec(IxDyn(&[2, 6, 7]), out).unwrap())
    }
    /// Many states at once: one bitboard kernel launch instead of n.
    pub fn get_next_state_batch(states: &[Self], players: &[i8], actions: &[u8]) -> Result<(Vec<Self>, Vec<i8>)> {
        assert!(states.len() == players.len() && states.len() == actions.len());
        let n = states.len();
        let mut out = states.to_vec();
        let mut next = vec![0i8; n];
        check(unsafe {
            sys::azb_c4_next_state(states.as_ptr() as *const sys::azb_c4_state, players.as_ptr(), actions.as_ptr(), n,
                                   out.as_mut_ptr() as *mut sys::azb_c4_state, next.as_mut_ptr())
        })?;
        Ok((out, next))
    }
}

// -------------------------------------------------------------------------------------------------
// The reference's plugin traits, signature for signature (src/game.rs:10-28, src/nnet.rs:35-45), so that code written
// against `alphazero_rs::game::Game` / `alphazero_rs::nnet::NNet` compiles against this crate by changing the `use` line.
// The reference's own path panics on failure (`unwrap`, `assert!`); so do these trait methods — the `try_*` inherent
// methods return `Result` for callers that want the status code.
// -------------------------------------------------------------------------------------------------
pub type PolicyView<'a> = ArrayView1<'a, f32>; // src/nnet.rs:18
pub type BatchedBoardFeaturesView<'a> = ArrayViewD<'a, F>; // src/nnet.rs:13
pub type BatchedPolicy = Array2<f32>; // src/nnet.rs:15
pub type BatchedValue = Array1<f32>; // src/nnet.rs:20
pub type ArcSOATrainingSamples = (ndarray::ArcArray<F, IxDyn>, ndarray::ArcArray<f32, ndarray::Ix2>, ndarray::ArcArray<f32, Ix1>); // src/nnet.rs:29-33

pub trait Game: std::fmt::Display + Sized + Send + Clone + std::hash::Hash + Eq {
    fn get_init_board() -> Self;
    fn get_feature_shape() -> Vec<usize>;
    fn get_next_state(&self, player: i8, action: u8) -> (Self, i8);
    fn get_valid_moves(&self, player: i8) -> Array<u8, Ix1>;
    fn get_game_ended(&self, player: i8) -> f32;
    fn get_canonical_form(&self, player: i8) -> Self;
    fn get_symmetries(&self, pi: PolicyView) -> Vec<(Self, Policy)>;
    fn eval_heuristic(&self) -> f32;
    fn to_features(&self) -> ArrayD<F>;
}

pub trait NNet {
    fn new<P: AsRef<Path>>(checkpoint: P) -> Self;
    fn train(&mut self, examples: ArcSOATrainingSamples, previous_model_id: usize, model_id: usize);
    fn predict(&self, board: BatchedBoardFeaturesView, model_id: usize) -> (BatchedPolicy, BatchedValue);
}

/// Quirk profile the trait methods run under (`get_game_ended`'s scan ranges, Q1).  The trait has no room for it, so it
/// is process-wide like the reference's compile-time constants; default = `AZB_PROFILE_SANE`.
pub static GAME_QUIRKS: std::sync::atomic::AtomicU32 = std::sync::atomic::AtomicU32::new(sys::AZB_PROFILE_SANE);

impl Game for ConnectFourGame {
    fn get_init_board() -> Self {
        ConnectFourGame::try_get_init_board().expect("azb_c4_init")
    }
    fn get_feature_shape() -> Vec<usize> {
        ConnectFourGame::feature_shape()
    }
    fn get_next_state(&self, player: i8, action: u8) -> (Self, i8) {
        self.try_get_next_state(player, action).expect("get_next_state (the reference underflows and panics on a full column)")
    }
    fn get_valid_moves(&self, player: i8) -> Array<u8, Ix1> {
        self.try_get_valid_moves(player).expect("get_valid_moves")
    }
    fn get_game_ended(&self, player: i8) -> f32 {
        self.try_get_game_ended(player, GAME_QUIRKS.load(std::sync::atomic::Ordering::Relaxed)).expect("get_game_ended")
    }
    fn get_canonical_form(&self, player: i8) -> Self {
        self.try_get_canonical_form(player).expect("get_canonical_form")
    }
    fn get_symmetries(&self, pi: PolicyView) -> Vec<(Self, Policy)> {
        self.try_get_symmetries(pi).expect("get_symmetries")
    }
    fn eval_heuristic(&self) -> f32 {
        self.try_eval_heuristic().expect("eval_heuristic")
    }
    fn to_features(&self) -> ArrayD<F> {
        self.try_to_features().expect("to_features")
    }
}

// -------------------------------------------------------------------------------------------------
// NNet (src/nnet.rs:35-45)
// -------------------------------------------------------------------------------------------------
pub struct B200Net {
    h: *mut sys::azb_nnet,
    pub train_config: sys::azb_train_config,
    checkpoint: std::path::PathBuf,
}
unsafe impl Send for B200Net {}

impl B200Net {
    /// NNet::new(checkpoint): ResNet-6x128, bf16 tensor-core tower, He-normal init; `<checkpoint>/0.azbw` is loaded
    /// when it exists.
    pub fn try_new<P: AsRef<Path>>(checkpoint: P) -> Result<Self> {
        Self::with_config(checkpoint, sys::azb_nnet_config { device: 0, blocks: 6, precision: sys::AZB_NNET_BF16_TC, reserved: 0, seed: 7 })
    }
    pub fn with_config<P: AsRef<Path>>(checkpoint: P, cfg: sys::azb_nnet_config) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::azb_nnet_create(&cfg, &mut h) })?;
        let net = B200Net {
            h,
            train_config: sys::azb_train_config { lr: 1e-3, beta1: 0.9, beta2: 0.999, eps: 1e-8 },
            checkpoint: checkpoint.as_ref().to_path_buf(),
        };
        let w0 = net.checkpoint.join("0.azbw");
        if w0.exists() {
            net.load(&w0)?;
        }
        Ok(net)
    }
    pub(crate) fn from_raw(h: *mut sys::azb_nnet, checkpoint: std::path::PathBuf) -> Self {
        B200Net { h, train_config: sys::azb_train_config { lr: 1e-3, beta1: 0.9, beta2: 0.999, eps: 1e-8 }, checkpoint }
    }
    pub fn as_raw(&self) -> *mut sys::azb_nnet {
        self.h
    }
    /// NNet::train(examples, previous_model_id, model_id): one Adam step on the batch; the trained weights are saved
    /// as `<checkpoint>/<model_id>.azbw` (python_nnet.rs:76-79).  Returns (policy loss, value loss).
    pub fn try_train(&mut self, examples: &SOATrainingSamples, _previous_model_id: usize, model_id: usize) -> Result<(f32, f32)> {
        let (boards, pis, vs) = examples;
        let n = vs.len();
        assert!(boards.len() == n * sys::AZB_C4_FEATURES && pis.len() == n * sys::AZB_C4_ACTIONS);
        let b = boards.as_standard_layout();
        let p = pis.as_standard_layout();
        let v = vs.as_standard_layout();
        let mut loss = [0f32; 2];
        check(unsafe {
            sys::azb_nnet_train(self.h, b.as_ptr(), p.as_ptr(), v.as_ptr(), n as u64, &self.train_config, loss.as_mut_ptr())
        })?;
        self.save(self.checkpoint.join(format!("{}.azbw", model_id)))?;
        Ok((loss[0], loss[1]))
    }
    /// NNet::predict(board[B,2,6,7], model_id) -> (pi[B,7], v[B])
    pub fn try_predict(&self, board: ArrayViewD<F>, model_id: usize) -> Result<(Array2<f32>, Array1<f32>)> {
        let n = board.len() / sys::AZB_C4_FEATURES;
        let b = board.as_standard_layout();
        let mut pi = Array2::<f32>::zeros((n, sys::AZB_C4_ACTIONS));
        let mut v = Array1::<f32>::zeros(n);
        check(unsafe { sys::azb_nnet_predict(self.h, b.as_ptr(), n, model_id, pi.as_mut_ptr(), v.as_mut_ptr()) })?;
        Ok((pi, v))
    }
    pub fn save<P: AsRef<Path>>(&self, path: P) -> Result<()> {
        check(unsafe { sys::azb_nnet_save(self.h, c_path(path).as_ptr()) })
    }
    pub fn load<P: AsRef<Path>>(&self, path: P) -> Result<()> {
        check(unsafe { sys::azb_nnet_load(self.h, c_path(path).as_ptr()) })
    }
}
impl Drop for B200Net {
    fn drop(&mut self) {
        unsafe { sys::azb_nnet_destroy(self.h) };
    }
}

impl NNet for B200Net {
    /// src/nnet.rs:36
    fn new<P: AsRef<Path>>(checkpoint: P) -> Self {
        B200Net::try_new(checkpoint).expect("NNet::new")
    }
    /// src/nnet.rs:38 — one Adam step on the batch, weights saved as `<checkpoint>/<model_id>.azbw`
    fn train(&mut self, examples: ArcSOATrainingSamples, previous_model_id: usize, model_id: usize) {
        let owned: SOATrainingSamples = (examples.0.to_owned(), examples.1.to_owned(), examples.2.to_owned());
        self.try_train(&owned, previous_model_id, model_id).expect("NNet::train");
    }
    /// src/nnet.rs:40-44
    fn predict(&self, board: BatchedBoardFeaturesView, model_id: usize) -> (BatchedPolicy, BatchedValue) {
        self.try_predict(board, model_id).expect("NNet::predict")
    }
}

// -------------------------------------------------------------------------------------------------
// AsyncMcts (src/async_mcts.rs; private in the reference, which builds the arena players from it: coach.rs:333-372)
// -------------------------------------------------------------------------------------------------
/// One persistent search tree on the device with a fused evaluator (`AZB_EVAL_UNIFORM` = the example's stub network,
/// examples/connect_four.rs:26-42).  Trees that search with a network live inside `Coach` / `arena::play_games_mcts`.
pub struct AsyncMcts {
    h: *mut sys::azb_mcts,
}
impl AsyncMcts {
    /// AsyncMcts::default (async_mcts.rs:27-48) from a filled `azb_config` (num_sims, cpuct, max_depth, evaluator, ...).
    pub fn from_config(cfg: &sys::azb_config) -> Result<Self> {
        let mut h = ptr::null_mut();
        check(unsafe { sys::azb_mcts_create(cfg, 1, &mut h) })?;
        Ok(AsyncMcts { h })
    }
    /// get_action_prob(s, temp, ...) — async_mcts.rs:74-115: num_sims simulations from the canonical state, then the
    /// visit-count policy.
    pub fn get_action_prob(&self, s: &ConnectFourGame, temp: f32) -> Result<Policy> {
        let mut counts = [0u16; sys::AZB_C4_ACTIONS];
        let mut pi = [0f32; sys::AZB_C4_ACTIONS];
        check(unsafe { sys::azb_mcts_get_action_prob(self.h, &s.0, temp, counts.as_mut_ptr(), pi.as_mut_ptr()) })?;
        Ok(Array::from(pi.to_vec()))
    }
}
impl Drop for AsyncMcts {
    fn drop(&mut self) {
        unsafe { sys::azb_mcts_destroy(self.h) };
    }
}

// -------------------------------------------------------------------------------------------------
// Coach (src/coach.rs)
// -------------------------------------------------------------------------------------------------
pub struct Coach {
    h: *mut sys::azb_coach,
    pub cfg: sys::azb_config,
    _checkpoint: CString, // keeps cfg.checkpoint_directory alive
}

impl Coach {
    /// Coach::setup — the 15 positional parameters of coach.rs:38-54 in the same order.  The newest
    /// `<n>.examples` of the directory becomes the history (coach.rs:55-81).
    #[allow(clippy::too_many_arguments)]
    pub fn setup<P: AsRef<Path>>(
        checkpoint_directory: P,
        mcts_reserve_size: usize,
        update_threshold: f32,
        temp_threshold: usize,
        max_history_length: usize,
        max_queue_length: usize,
        inference_batch_size: usize,
        num_episode_threads: usize,
        num_arena_games: usize,
        num_iters: usize,
        num_eps: usize,
        num_sims: usize,
        num_sim_threads: usize,
        max_depth: usize,
        cpuct: i32,
    ) -> Result<Coach> {
        let dir = c_path(checkpoint_directory);
        let mut cfg: sys::azb_config = unsafe { std::mem::zeroed() };
        unsafe { sys::azb_config_default(&mut cfg) };
        cfg.checkpoint_directory = dir.as_ptr();
        cfg.mcts_reserve_size = mcts_reserve_size as u64;
        cfg.update_threshold = update_threshold;
        cfg.temp_threshold = temp_threshold as u64;
        cfg.max_history_length = max_history_length as u64;
        cfg.max_queue_length = max_queue_length as u64;
        cfg.inference_batch_size = inference_batch_size as u64;
        cfg.num_episode_threads = num_episode_threads as u64;
        cfg.num_arena_games = num_arena_games as u64;
        cfg.num_iters = num_iters as u64;
        cfg.num_eps = num_eps as u64;
        cfg.num_sims = num_sims as u64;
        cfg.num_sim_threads = num_sim_threads as u64;
        cfg.max_depth = max_depth as u64;
        cfg.cpuct = cpuct;
        cfg.evaluator = sys::AZB_EVAL_NNET;
        Self::from_config(cfg, dir)
    }
    /// The engine's own fields (seed, quirks, evaluator, device, …) can be set on `cfg` first.
    pub fn from_config(cfg: sys::azb_config, checkpoint: CString) -> Result<Coach> {
        let mut cfg = cfg;
        cfg.checkpoint_directory = checkpoint.as_ptr();
        let mut h = ptr::null_mut();
        check(unsafe { sys::azb_coach_setup(&cfg, &mut h) })?;
        Ok(Coach { h, cfg, _checkpoint: checkpoint })
    }
    /// The network behind `NNet::predict` for self-play when `cfg.evaluator == AZB_EVAL_NNET`.
    pub fn set_nnet(&mut self, net: &B200Net) -> Result<()> {
        check(unsafe { sys::azb_coach_set_nnet(self.h, net.as_raw()) })
    }
    /// `n` concurrent `execute_episode` calls (coach.rs:104-157 under the fan-out of :241-272); episode e uses the
    /// Philox stream (seed, first_episode_id + e).  Returns the samples in episode, ply, symmetry order.
    pub fn execute_episodes(&mut self, n: usize, first_episode_id: u64) -> Result<(SOATrainingSamples, sys::azb_selfplay_stats)> {
        let mut st = sys::azb_selfplay_stats::default();
        check(unsafe { sys::azb_coach_self_play(self.h, n as u64, first_episode_id, &mut st) })?;
        let ns = st.samples as usize;
        let mut boards = ArrayD::<F>::zeros(IxDyn(&[ns, 2, 6, 7]));
        let mut pis = Array2::<f32>::zeros((ns, sys::AZB_C4_ACTIONS));
        let mut vs = Array1::<f32>::zeros(ns);
        let mut written = 0u64;
        check(unsafe {
            sys::azb_coach_export_samples(self.h, boards.as_mut_ptr(), pis.as_mut_ptr(), vs.as_mut_ptr(), ns as u64, &mut written)
        })?;
        debug_assert_eq!(written as usize, ns);
        Ok(((boards, pis, vs), st))
    }
    /// The fan-out in two halves (`azb_coach_self_play_begin` / `_end`): `begin` launches and returns, `end` waits and hands the
    /// samples out.  Two coaches used in turn overlap consecutive batches on the device (a rayon pool has no barrier between
    /// episodes either, coach.rs:241-272).
    pub fn execute_episodes_begin(&mut self, n: usize, first_episode_id: u64) -> Result<()> {
        check(unsafe { sys::azb_coach_self_play_begin(self.h, n as u64, first_episode_id) })
    }
    pub fn execute_episodes_end(&mut self) -> Result<(SOATrainingSamples, sys::azb_selfplay_stats)> {
        let mut st = sys::azb_selfplay_stats::default();
        check(unsafe { sys::azb_coach_self_play_end(self.h, &mut st) })?;
        let ns = st.samples as usize;
        let mut boards = ArrayD::<F>::zeros(IxDyn(&[ns, 2, 6, 7]));
        let mut pis = Array2::<f32>::zeros((ns, sys::AZB_C4_ACTIONS));
        let mut vs = Array1::<f32>::zeros(ns);
        let mut written = 0u64;
        check(unsafe {
            sys::azb_coach_export_samples(self.h, boards.as_mut_ptr(), pis.as_mut_ptr(), vs.as_mut_ptr(), ns as u64, &mut written)
        })?;
        Ok(((boards, pis, vs), st))
    }
    /// Coach::execute_episode(mcts, episode_id, rng) — coach.rs:104-109.
    pub fn execute_episode(&mut self, episode_id: u64) -> Result<SOATrainingSamples> {
        Ok(self.execute_episodes(1, episode_id)?.0)
    }
    /// Coach::save_train_examples(iteration, checkpoint) — coach.rs:159-167.
    pub fn save_train_examples<P: AsRef<Path>>(&self, iteration: usize, checkpoint: P) -> Result<()> {
        check(unsafe { sys::azb_coach_save_train_examples(self.h, iteration as u64, c_path(checkpoint).as_ptr()) })
    }
    /// Coach::learn(checkpoint, skip_first_play, verbose, rng) — coach.rs:169-396.  The checkpoint directory is the one
    /// given to `setup`; the RNG is the engine's Philox keyed by `cfg.seed`.  Returns one report per iteration and the
    /// accepted model.
    pub fn learn(&mut self, skip_first_play: bool, verbose: bool, net_cfg: sys::azb_nnet_config,
                 schedule: Option<sys::azb_learn_config>) -> Result<(Vec<sys::azb_learn_report>, B200Net)> {
        let mut lc = schedule.unwrap_or_else(|| {
            let mut d: sys::azb_learn_config = unsafe { std::mem::zeroed() };
            unsafe { sys::azb_learn_config_default(&mut d) };
            d
        });
        lc.skip_first_play = skip_first_play as u32;
        let n_it = self.cfg.num_iters as usize;
        let mut reports = vec![sys::azb_learn_report::default(); n_it.max(1)];
        let mut n = 0u64;
        let mut h = ptr::null_mut();
        check(unsafe { sys::azb_coach_learn(self.h, &net_cfg, &lc, reports.as_mut_ptr(), n_it as u64, &mut n, &mut h) })?;
        reports.truncate(n as usize);
        if verbose {
            for r in &reports {
                // coach.rs:381,385-389
                println!("NEW/PREV WINS : {} / {}; DRAWS : {}", r.nwins, r.pwins, r.draws);
                println!("{}", if r.accepted != 0 { "ACCEPTING NEW MODEL" } else { "REJECTING NEW MODEL" });
            }
        }
        let dir = unsafe { CStr::from_ptr(self.cfg.checkpoint_directory) }.to_string_lossy().into_owned();
        Ok((reports, B200Net::from_raw(h, dir.into())))
    }
}
impl Drop for Coach {
    fn drop(&mut self) {
        unsafe { sys::azb_coach_destroy(self.h) };
    }
}

// -------------------------------------------------------------------------------------------------
// arena (src/arena.rs)
// -------------------------------------------------------------------------------------------------
pub mod arena {
    use super::*;

    #[derive(Hash, PartialEq, Eq, Clone, Copy, Debug)]
    pub enum GameResult {
        Win,
        Loss,
        Draw,
    } // src/arena.rs:54-59

    /// arena::play_game(player_actions, board, verbose) — src/arena.rs:7-52, closure for closure: the two players are
    /// host closures over a canonical board, every `Game` call is one call of the bitboard kernel.  This is the literal
    /// API for callers that bring their own players; the gating match of Coach::learn runs on the device instead
    /// (`play_games_mcts`).
    pub fn play_game<G: Game>(player_actions: &[&dyn Fn(&G) -> u8], board: &Option<G>, verbose: bool) -> i8 {
        assert!(player_actions.len() == 2);
        let mut cur_player: i8 = 1;
        let mut board: G = board.clone().unwrap_or_else(G::get_init_board);
        let mut iteration = 0;
        while board.get_game_ended(cur_player) == 0.0 {
            iteration += 1;
            if verbose {
                println!("Turn {}, Player {}\n{}", iteration, cur_player, board);
            }
            let canonical_board = board.get_canonical_form(cur_player);
            let action = player_actions[if cur_player == 1 { 0 } else { 1 }](&canonical_board);
            let valids = canonical_board.get_valid_moves(1);
            assert!(valids[action as usize] > 0, "Action {} is not valid! valids = {}", action, valids);
            let (new_board, new_cur_player) = board.get_next_state(cur_player, action);
            board = new_board;
            cur_player = new_cur_player;
        }
        // arena.rs:51: a value close to 0 rounds to a draw
        cur_player * f32::round(board.get_game_ended(cur_player)) as i8
    }

    /// arena::play_games(num, player_actions, board, verbose) — src/arena.rs:62-99: both seat orders, num / 2 games each,
    /// tallied for the FIRST closure (Heap's permutations of two elements: [0, 1] then [1, 0]).
    pub fn play_games<G: Game>(num: usize, player_actions: Vec<&dyn Fn(&G) -> u8>, board: Option<G>, verbose: bool)
                               -> std::collections::HashMap<GameResult, usize> {
        assert!(player_actions.len() == 2);
        let mut all: std::collections::HashMap<GameResult, usize> = std::collections::HashMap::new();
        for ordering in 0..2 {
            let seated: Vec<&dyn Fn(&G) -> u8> =
                if ordering == 0 { vec![player_actions[0], player_actions[1]] } else { vec![player_actions[1], player_actions[0]] };
            let (win_cond, lose_cond) = if ordering == 0 { (1, -1) } else { (-1, 1) };
            for _ in 0..(num / 2) {
                let r = play_game(&seated, &board, verbose);
                let key = if r == win_cond { GameResult::Win } else if r == lose_cond { GameResult::Loss } else { GameResult::Draw };
                *all.entry(key).or_insert(0) += 1;
            }
        }
        all
    }

    /// The match of coach.rs:333-375 (two MCTS players, temp 0, arg-max with ties to the highest action) entirely on the
    /// device.  `opts.shared_trees = 1` is the reference's layout (pmcts / nmcts created once, games one after the other).
    pub fn play_games_mcts(cfg: &sys::azb_config, num: usize, net_a: &B200Net, net_b: &B200Net, opts: sys::azb_arena_opts)
                           -> Result<std::collections::HashMap<GameResult, usize>> {
        let mut counts = [0u64; 3];
        check(unsafe {
            sys::azb_arena_play_games_ex(cfg, num as u64, sys::AZB_EVAL_NNET, sys::AZB_EVAL_NNET, net_a.as_raw(), net_b.as_raw(),
                                         &opts, counts.as_mut_ptr(), ptr::null_mut(), ptr::null_mut(), ptr::null_mut(),
                                         ptr::null_mut(), ptr::null_mut())
        })?;
        let mut out = std::collections::HashMap::new();
        out.insert(GameResult::Win, counts[0] as usize);
        out.insert(GameResult::Loss, counts[1] as usize);
        out.insert(GameResult::Draw, counts[2] as usize);
        Ok(out)
    }
}

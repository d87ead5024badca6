//! One `Coach::learn` run on the B200 engine with the parameter values the reference's connect-four example uses
//! (examples/connect_four.rs:55-71 of alphazero-rs), except for the number of episodes: thousands of concurrent games are
//! what fills a B200.
use azb200::Coach;
use azb200_sys as sys;

struct Params {
    checkpoint_directory: &'static str,
    mcts_reserve_size: usize,
    update_threshold: f32,
    temp_threshold: usize,
    max_history_length: usize,
    max_queue_length: usize,
    num_arena_games: usize,
    num_iters: usize,
    num_eps: usize,
    num_sims: usize,
    max_depth: usize,
    cpuct: i32,
}

fn main() -> Result<(), Box<dyn std::error::Error>> {
    let p = Params {
        checkpoint_directory: "./checkpoint",
        mcts_reserve_size: 1_000_000,
        update_threshold: 0.6,
        temp_threshold: 15,
        max_history_length: 20,
        max_queue_length: 200_000,
        num_arena_games: 40,
        num_iters: 1,
        num_eps: 4096,
        num_sims: 25,
        max_depth: 1000,
        cpuct: 1,
    };
    // inference_batch_size and num_episode_threads are accepted and ignored (a round evaluates every pending leaf; all games
    // of an iteration run concurrently); num_sim_threads must be 1 (deterministic mode)
    let (inference_batch_size, num_episode_threads, num_sim_threads) = (1, 1, 1);
    let mut coach = Coach::setup(
        p.checkpoint_directory, p.mcts_reserve_size, p.update_threshold, p.temp_threshold, p.max_history_length,
        p.max_queue_length, inference_batch_size, num_episode_threads, p.num_arena_games, p.num_iters, p.num_eps, p.num_sims,
        num_sim_threads, p.max_depth, p.cpuct,
    )?;
    let net_cfg = sys::azb_nnet_config { device: 0, blocks: 6, precision: sys::AZB_NNET_BF16_TC, reserved: 0, seed: 7 };
    let (reports, _model) = coach.learn(false, true, net_cfg, None)?;
    for r in reports {
        println!("iteration {}: {} games, {} samples in the window, model {} -> {}", r.iteration, r.games, r.history_samples,
                 r.model_id_before, r.model_id_after);
    }
    Ok(())
}

//! The reference's examples/connect_four.rs:55-77 on the B200 engine: same 15 setup parameters, then learn.
use azb200::Coach;
use azb200_sys as sys;

fn main() -> Result<(), Box<dyn std::error::Error>> {
    let mut coach = Coach::setup(
        "./checkpoint", // checkpoint_directory
        1000000,        // mcts_reserve_size
        0.6,            // update_threshold
        15,             // temp_threshold
        20,             // max_history_length
        200000,         // max_queue_length
        1,              // inference_batch_size (ignored: a round evaluates every pending leaf)
        1,              // num_episode_threads (ignored: all games of an iteration run concurrently)
        40,             // num_arena_games
        1,              // num_iters
        4096,           // num_eps: thousands of concurrent games are what fills a B200
        25,             // num_sims
        1,              // num_sim_threads
        1000,           // max_depth
        1,              // cpuct
    )?;
    let net_cfg = sys::azb_nnet_config { device: 0, blocks: 6, precision: sys::AZB_NNET_BF16_TC, reserved: 0, seed: 7 };
    let (reports, _model) = coach.learn(false, true, net_cfg, None)?;
    for r in reports {
        println!("iteration {}: {} games, {} samples in the window, model {} -> {}", r.iteration, r.games, r.history_samples,
                 r.model_id_before, r.model_id_after);
    }
    Ok(())
}

"""BASELINE.json configs other than the headline one, one JSON line each (device time, CUDA events inside the library):
python scripts/bench_configs.py [micro] [config3] [config4]
  micro   : SURVEY 8(d) fixed-work micro-bench: 800 simulations from the empty board x 4096 trees (azb_mcts hook)
  config3 : 8192 self-play games x 400 sims with the ResNet-6x128 bf16 evaluator
  config4 : one GPU's share of the arena config: 2048 games, nets seeded 7 / 8, 400 sims, 4 random opening plies"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
which = sys.argv[1:] or ["micro", "config3", "config4"]
FLOP = lambda blocks: 2 * 42 * 18 * 128 + 2 * blocks * 2 * 42 * 1152 * 128 + 21504 + 1176 + 10752 + 5376 + 128
if "micro" in which:
    import numpy as np
    n, sims = 4096, 800
    m = azb.AsyncMcts(n, num_sims=sims, evaluator=azb.EVAL_UNIFORM, mcts_reserve_size=1000000)
    root = azb.ConnectFourGame.get_init_board(n)
    m.get_action_prob(root, 1.0)                      # warm-up (also the first 800 simulations of every tree)
    t0 = time.perf_counter(); m.get_action_prob(root, 1.0); dt = time.perf_counter() - t0
    print(json.dumps({"workload": "micro: 800 sims from the empty board x 4096 trees (second search on the same trees)",
                      "wall_ms": 1e3 * dt, "sims_per_sec": n * sims / dt}))
K = int(os.environ.get("AZB200_BENCH_THREADS", "1"))  # num_sim_threads: waves of K simulations per tree
if "config3" in which:
    net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET, num_sim_threads=K)
    st = coach.self_play(8192, 0)
    s = st["device_ms"] * 1e-3
    print(json.dumps({"workload": "config3: 8192 games x 400 sims, ResNet-6x128 bf16 leaf evaluator", "num_sim_threads": K, "device_s": s,
                      "sims_per_sec": st["sims"] / s, "games_per_sec": st["games"] / s, "leaf_evals_per_sec": st["evals"] / s,
                      "nn_positions": st["nn_positions"], "nn_cache_hits": st["nn_cache_hits"],
                      "nn_tflops_whole_run": st["nn_positions"] * FLOP(6) / s / 1e12, "kernel_launches": st["launches"], "plies": st["plies"]}))
if "config4" in which:
    a = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
    b = azb.NNet(seed=8, blocks=6, precision=azb.NNET_BF16_TC)
    counts, res, st = azb.arena_play_games(2048, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=4, num_sims=400, seed=0xA1FA0)
    s = st["device_ms"] * 1e-3
    print(json.dumps({"workload": "config4 share: 2048 arena games (1024 per seat order), nets 7 vs 8, 400 sims, 4 random opening plies",
                      "device_s": s, "games_per_sec": 2048 / s, "sims_per_sec": st["sims"] / s, "win_loss_draw_of_net_7": list(map(int, counts))}))

"""Arena throughput (BASELINE config 4 per GPU share): python scripts/profile_arena.py [games] [sims] [blocks] [k_open]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 400
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 6
k_open = int(sys.argv[4]) if len(sys.argv) > 4 else 4
a = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC)
b = azb.NNet(seed=8, blocks=blocks, precision=azb.NNET_BF16_TC)
t0 = time.time()
counts, res, st = azb.arena_play_games(games, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=k_open, num_sims=sims, seed=0xA1FA0)
print("W/L/D of net A:", counts, {k: st[k] for k in ("games", "plies", "sims", "evals", "nn_positions", "nn_cache_hits", "device_ms", "launches")})
print("games/s=%.1f sims/s=%.3e wall=%.1fs" % (st["games"] / st["device_ms"] * 1e3, st["sims"] / st["device_ms"] * 1e3, time.time() - t0))

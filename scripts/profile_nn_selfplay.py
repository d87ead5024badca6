"""Self-play with the ResNet evaluator (BASELINE config 3): python scripts/profile_nn_selfplay.py [games] [sims] [blocks]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 400
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 6
net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC)
coach = azb.Coach(nnet=net, num_sims=sims, seed=0xA1FA0, evaluator=azb.EVAL_NNET)
t0 = time.time()
st = coach.self_play(games, 0)
dt = time.time() - t0
flop_per_pos = 2 * 42 * 18 * 128 + 2 * blocks * 2 * 42 * 1152 * 128 + 21504 + 1176 + 10752 + 5376 + 128
print({k: st[k] for k in ("games", "plies", "sims", "levels", "expansions", "evals", "nn_positions", "nn_cache_hits", "device_ms", "launches")})
print("sims/s=%.3e games/s=%.1f evals/s=%.3e nn_TFLOP/s(whole run, positions actually evaluated)=%.1f wall=%.1fs" % (
    st["sims"] / st["device_ms"] * 1e3, st["games"] / st["device_ms"] * 1e3, st["evals"] / st["device_ms"] * 1e3,
    st["nn_positions"] * flop_per_pos / st["device_ms"] / 1e9, dt))

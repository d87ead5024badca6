"""A network self-play call for ncu (the round kernel): python scripts/profile_rounds.py [games] [sims]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 400
net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC)
coach = azb.Coach(nnet=net, num_sims=sims, seed=0xA1FA0, evaluator=azb.EVAL_NNET)
st = coach.self_play(games, 0)
print({k: st[k] for k in ("games", "plies", "sims", "device_ms", "launches")})

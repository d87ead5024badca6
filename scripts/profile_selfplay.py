"""One self-play batch for ncu: python scripts/profile_selfplay.py [games] [sims] [evaluator]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
ev = int(sys.argv[3]) if len(sys.argv) > 3 else 0
schedule = int(sys.argv[4]) if len(sys.argv) > 4 else 0
ppl = int(sys.argv[5]) if len(sys.argv) > 5 else 0
coach = azb.Coach(num_sims=sims, seed=0xA1FA0, evaluator=ev, schedule=schedule, plies_per_launch=ppl)
st = coach.self_play(games, 0)
print({k: st[k] for k in ("games", "plies", "sims", "levels", "expansions", "terminal_hits", "dup_links", "device_ms", "launches", "trees_resident")},
      "sims/s=%.3e" % (st["sims"] / st["device_ms"] * 1e3))

"""Where the end-to-end step goes: python scripts/e2e_breakdown.py [games] [sims]
wall time of azb_coach_self_play vs its device time, and of azb_coach_export_samples (pinned host buffers)."""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
games = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
sims = int(sys.argv[2]) if len(sys.argv) > 2 else 800
coach = azb.Coach(num_sims=sims, seed=0xA1FA0, evaluator=0)
cap = games * 84
pinned = [azb.PinnedArray((cap, 2, 6, 7)), azb.PinnedArray((cap, 7)), azb.PinnedArray((cap,))]
out = tuple(p.array for p in pinned)
for k in range(4):
    t0 = time.perf_counter()
    st = coach.self_play(games, k * games)
    t1 = time.perf_counter()
    b, p, v = coach.export_samples(out)
    t2 = time.perf_counter()
    nbytes = len(v) * 92 * 4
    print("step %d: self_play wall %.2f ms (device %.2f ms), export %.2f ms for %.1f MB (%.1f GB/s)"
          % (k, 1e3 * (t1 - t0), st["device_ms"], 1e3 * (t2 - t1), nbytes / 1e6, nbytes / (t2 - t1) / 1e9))

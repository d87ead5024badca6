"""Per-ply duration profile of the persistent self-play kernel (AZB200_PLY_TIMES=1)."""
import importlib, os, sys
os.environ["AZB200_PLY_TIMES"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
azb = importlib.import_module("alphazero-rs_b200")
coach = azb.Coach(num_sims=800, seed=0xA1FA0, evaluator=0, schedule=1)
coach.self_play(4096, 0)
st = coach.self_play(4096, 0)
t = coach.ply_times().astype(np.int64)
pl = coach.traces()["plies"]
t0 = t[t > 0].min()
print("kernel ms", st["device_ms"])
prev = np.full(4096, t0)
for p in range(42):
    act = pl > p
    d = (t[act, p] - prev[act]) / 1e6
    prev[act] = t[act, p]
    print(f"ply {p:2d} active {act.sum():5d} mean {d.mean():6.3f} ms  p10 {np.percentile(d,10):6.3f} p90 {np.percentile(d,90):6.3f} max {d.max():6.3f}  end(mean) {(t[act,p]-t0).mean()/1e6:7.2f}")

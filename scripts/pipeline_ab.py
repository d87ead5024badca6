"""Pipelined self-play throughput: K batches of 4096 games x 800 sims enqueued `depth` deep on `depth` coaches
(azb_coach_self_play_begin / _end).  python scripts/pipeline_ab.py [depth] [K] [games]"""
import importlib, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
K = int(sys.argv[2]) if len(sys.argv) > 2 else 8
G = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
coaches = [azb.Coach(num_sims=800, seed=0xA1FA0, evaluator=0) for _ in range(depth)]
for c in coaches:
    c.self_play(G, 0)  # warm-up: pools, carve-out
def run(k0, n):
    coaches[0].span_mark()
    t0 = time.perf_counter()
    for j in range(min(depth - 1, n)):
        coaches[j % depth].self_play_begin(G, (k0 + j) * G)
    sims = 0
    per = []
    for j in range(n):
        nxt = j + depth - 1
        if nxt < n:
            coaches[nxt % depth].self_play_begin(G, (k0 + nxt) * G)
        st = coaches[j % depth].self_play_end()
        sims += st["sims"]; per.append(round(st["device_ms"], 1))
    return sims, time.perf_counter() - t0, coaches[0].span_ms(coaches[(n - 1) % depth]), per
run(1, 2)
sims, wall, span, per = run(10, K)
print(f"depth={depth} K={K} games={G}: span {span:.1f} ms = {span / K:.1f} ms/step, {sims / span * 1e3 / 1e6:.1f} M sims/s (wall {wall * 1e3:.1f} ms); per-launch device ms {per}")

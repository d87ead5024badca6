"""Stress of the once-per-call evaluation (leaf de-duplication + evaluation cache): the same self-play call in child
processes with both switched on (several times) and off (once); every run must produce the same games.
  python scripts/stress_dedup.py [games] [sims] [blocks] [repeats]"""
import hashlib, importlib, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--child":
    sys.path.insert(0, ROOT)
    azb = importlib.import_module("alphazero-rs_b200")
    games, sims, blocks = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    net = azb.NNet(seed=7, blocks=blocks)
    coach = azb.Coach(nnet=net, num_sims=sims, seed=3, evaluator=azb.EVAL_NNET, temp_threshold=10)
    st = coach.self_play(games, 0)
    tr = coach.traces()
    b, p, v = coach.export_samples()
    h = hashlib.sha256()
    for a in (tr["actions"], tr["counts"], tr["plies"], b, p, v):
        h.update(a.tobytes())
    print(json.dumps({"sha": h.hexdigest(), "evals": st["evals"], "rows": st["nn_positions"], "hits": st["nn_cache_hits"],
                      "ms": round(st["device_ms"], 1)}))
    sys.exit(0)
games, sims, blocks, reps = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 2048), (2, 100), (3, 1), (4, 4)))
out = []
for tag, val in [("off", "0")] + [("on", "1")] * reps:
    env = dict(os.environ, AZB200_LEAF_DEDUP=val, AZB200_EVAL_CACHE=val)
    r = subprocess.run([sys.executable, __file__, "--child", str(games), str(sims), str(blocks)], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    d = json.loads(r.stdout.strip().splitlines()[-1])
    out.append((tag, d))
    print(tag, d)
ok = len({d["sha"] for _, d in out}) == 1
print("IDENTICAL" if ok else "MISMATCH")
sys.exit(0 if ok else 1)

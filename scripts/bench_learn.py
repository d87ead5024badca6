"""Coach::learn (coach.rs:169-396) end to end on one GPU through azb_coach_learn: per iteration self-play with the current
model, queue trim / history window, `<n>.examples`, one pass of Adam steps over the shuffled window (candidate = copy of the
current model), gating arena, accept rule.  Prints one JSON line per iteration and a summary.
  python scripts/bench_learn.py [num_iters] [num_eps] [num_sims] [batch] [arena_games] [epochs] [lr]"""
import importlib, json, os, shutil, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
arg = lambda i, d: int(sys.argv[i]) if len(sys.argv) > i else d
iters, eps, sims, batch, arena, epochs = arg(1, 3), arg(2, 4096), arg(3, 100), arg(4, 2048), arg(5, 512), arg(6, 0)
lr = float(sys.argv[7]) if len(sys.argv) > 7 else 1e-4
ckpt = tempfile.mkdtemp(prefix="azb_ckpt_")
coach = azb.Coach(checkpoint_directory=ckpt.encode(), evaluator=azb.EVAL_NNET, num_iters=iters, num_eps=eps, num_sims=sims,
                  num_arena_games=arena, max_queue_length=2_000_000, max_history_length=4, update_threshold=0.55, seed=0xA1FA0)
t0 = time.perf_counter()
reports, net = coach.learn(epochs=epochs, batch_size=batch, blocks=6, seed=7, arena_k_open=4, lr=lr)
wall = time.perf_counter() - t0
for r in reports:
    r = {k: (round(v, 1) if k.endswith("_ms") else [round(x, 4) for x in v] if isinstance(v, list) else v) for k, v in r.items()}
    print(json.dumps(r))
files = {f: os.path.getsize(os.path.join(ckpt, f)) for f in sorted(os.listdir(ckpt))}
print(json.dumps({"workload": f"Coach::learn: {iters} iterations x {eps} games x {sims} sims, ResNet-6x128 bf16, batch {batch}, "
                  f"{arena} arena games", "wall_s": round(wall, 2), "final_model_id": reports[-1]["model_id_after"], "files": files}))
shutil.rmtree(ckpt)

"""Small invocations of every kernel family (self-play, hash evaluator, network predict, network self-play, search hook):
python scripts/smoke_all_kernels.py"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
azb = importlib.import_module("alphazero-rs_b200")
coach = azb.Coach(num_sims=60, seed=3, evaluator=azb.EVAL_UNIFORM)
st = coach.self_play(48, 0); b, p, v = coach.export_samples()
print("selfplay", st["games"], st["sims"], len(v))
coach = azb.Coach(num_sims=40, seed=3, evaluator=1)
print("selfplay hash", coach.self_play(16, 0)["sims"])
net = azb.NNet(seed=7, blocks=2, precision=azb.NNET_BF16_TC)
rng = np.random.default_rng(0); n = 700
cur = rng.integers(0, 2, (n, 42)); opp = rng.integers(0, 2, (n, 42)) & (1 - cur)
feats = np.zeros((n, 2, 6, 7), np.float32); feats[:, 0] = cur.reshape(n, 6, 7); feats[:, 1] = opp.reshape(n, 6, 7)
pi, v = net.predict(feats); print("predict", pi.shape, float(v.mean()))
coach = azb.Coach(nnet=net, num_sims=20, seed=2, evaluator=azb.EVAL_NNET)
print("nn selfplay", coach.self_play(24, 0)["evals"])
m = azb.AsyncMcts(8, num_sims=100, evaluator=1, mcts_reserve_size=100000)
import oracle_api as orc
c, _ = m.get_action_prob(orc.init_board(8), 1.0); print("mcts", c[0].tolist())

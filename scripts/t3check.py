import importlib, os, sys, numpy as np
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
azb = importlib.import_module("alphazero-rs_b200")
rng = np.random.default_rng(0)
n = 300
cur = rng.integers(0, 2, (n, 42)); opp = (rng.integers(0, 2, (n, 42))) & (1 - cur)
feats = np.zeros((n, 2, 6, 7), np.float32); feats[:, 0] = cur.reshape(n, 6, 7); feats[:, 1] = opp.reshape(n, 6, 7)
net = azb.NNet(seed=7, blocks=1, precision=azb.NNET_BF16_TC)
net32 = azb.NNet(seed=7, blocks=1, precision=azb.NNET_FP32)
pi, v = net.predict(feats); p32, v32 = net32.predict(feats)
print(os.environ.get("AZB200_LIB", "default"), "max |pi - fp32|", np.abs(pi - p32).max(), "max |v - fp32|", np.abs(v - v32).max())

"""One training step of NNet::train (src/nnet.rs:38) on one GPU: python scripts/bench_train.py [batch] [steps] [blocks]
Host-timed (perf_counter around synchronous C-ABI calls: the step includes the host->device copy of the batch and the
device->host copy of the updated parameters); AZB200_TIMING=1 prints the section times of each call."""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
azb = importlib.import_module("alphazero-rs_b200")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
blocks = int(sys.argv[3]) if len(sys.argv) > 3 else 6
tower = 2 * blocks * 2 * 42 * 1152 * 128
fwd = 2 * 42 * 18 * 128 + tower + 21504 + 1176 + 10752 + 5376 + 128
flop_per_sample = fwd + 2 * tower  # forward + backward data + backward weights of the tower (stem/heads backward are ~0.2 %)
rng = np.random.default_rng(0)
boards = (rng.random((batch, 2, 6, 7)) < 0.3).astype(np.float32)
pis = rng.random((batch, 7)).astype(np.float32); pis /= pis.sum(1, keepdims=True)
vs = rng.choice(np.array([-1.0, 1.0], np.float32), batch)
net = azb.NNet(seed=7, blocks=blocks)
for _ in range(3):
    loss0 = net.train((boards, pis, vs))
t0 = time.perf_counter()
for _ in range(steps):
    loss = net.train((boards, pis, vs))
dt = (time.perf_counter() - t0) / steps
print(json.dumps({"workload": f"NNet::train step, ResNet-{blocks}x128 bf16 tower, batch {batch}", "ms_per_step": round(dt * 1e3, 3),
                  "samples_per_sec": round(batch / dt), "tflops_algorithmic": round(batch * flop_per_sample / dt / 1e12, 1),
                  "loss_after_3": [round(x, 4) for x in loss0], "loss_last": [round(x, 4) for x in loss]}))

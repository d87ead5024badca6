"""Forward pass of the leaf evaluator over resident batches of several sizes (device-timed mean of `iters` passes):
python scripts/forward_sweep.py [blocks] [iters]   -> ms per pass, positions/s, TFLOP/s, and the per-layer cost model"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
blocks = int(sys.argv[1]) if len(sys.argv) > 1 else 6
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 200
net = azb.NNet(seed=7, blocks=blocks)
for batch in (64, 338, 676, 1014, 1352, 2028, 4096, 8192):
    net.benchmark(batch, 5)
    ms = net.benchmark(batch, iters)
    print(f"batch {batch:5d}: {ms * 1e3:8.1f} us/pass  {batch / ms / 1e3:7.2f} M pos/s  {batch * 148.87e6 / ms / 1e9:7.1f} TFLOP/s")

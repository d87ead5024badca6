"""Forward-pass throughput of the network: python scripts/bench_nnet.py [batch] [blocks]"""
import importlib, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
blocks = int(sys.argv[2]) if len(sys.argv) > 2 else 6
flop_per_pos = 2 * 42 * 18 * 128 + 2 * blocks * 2 * 42 * 1152 * 128 + 21504 + 1176 + 10752 + 5376 + 128
net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC)
ms = net.benchmark(batch, 20)
print(f"bf16_tc batch={batch} blocks={blocks}: {ms:.3f} ms/pass  {batch/ms*1e3:.3e} pos/s  {batch*flop_per_pos/ms/1e9:.1f} TFLOP/s")

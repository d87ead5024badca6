"""Does NVML polling from a side thread delay the CUDA calls of the step?  python scripts/nvml_interference.py [period_s]"""
import importlib, os, sys, time, threading
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
azb = importlib.import_module("alphazero-rs_b200")
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
period = float(sys.argv[1]) if len(sys.argv) > 1 else 0.05
coach = azb.Coach(num_sims=800, seed=0xA1FA0, evaluator=0)
games = 4096
cap = games * 84
pinned = [azb.PinnedArray((cap, 2, 6, 7)), azb.PinnedArray((cap, 7)), azb.PinnedArray((cap,))]
out = tuple(p.array for p in pinned)
def run(tag, n=8):
    res = []
    for k in range(n):
        t0 = time.perf_counter(); st = coach.self_play(games, k * games); t1 = time.perf_counter()
        coach.export_samples(out); t2 = time.perf_counter()
        res.append((round(st["device_ms"], 1), round(1e3 * (t1 - t0), 1), round(1e3 * (t2 - t1), 1)))
    print(tag, res)
run("warm", 3)
run("no sampler")
on = True; calls = []
def loop(which):
    while on:
        t0 = time.perf_counter()
        if which & 1: pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
        if which & 2: pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        calls.append(round(1e3 * (time.perf_counter() - t0), 2))
        time.sleep(period)
for which, name in ((1, "clock only"), (2, "reasons only"), (3, "both")):
    on = True; calls.clear()
    th = threading.Thread(target=loop, args=(which,), daemon=True); th.start()
    run(name)
    on = False; th.join()
    print("  nvml call ms: max %.2f mean %.2f n %d" % (max(calls), sum(calls) / len(calls), len(calls)))

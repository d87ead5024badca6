#!/usr/bin/env python
"""bench.py — MCTS sims/sec of the self-play hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

Workload at every N (weak scaling): BASELINE config 2 per GPU — 4096 concurrent connect-four
self-play games x 800 sims/move, uniform-prior evaluator, temp 1 for plies 1-14 then 0, Philox
stream (seed 0xA1FA0, global game id), every game played to completion.  One "step" = one such
batch of games (Coach::execute_episode x 4096, coach.rs:104-157,241-272).

  value  = simulations of all ranks / device time of the self-play kernel (CUDA events on the
           launching stream, max over ranks) — nothing but the config lives on the host;
  e2e    = the same through the C ABI as a caller uses it: azb_coach_self_play followed by
           azb_coach_export_samples into page-locked host buffers (config upload, kernel,
           sample expansion, device->host copy of the SOA samples), wall clock, max over ranks.

`--impl reference` times the CPU oracle (oracle/, the line-by-line restatement of the
reference's Rust path — the reference itself cannot be built here) on all host cores.
"""
import argparse
import importlib
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

GAMES_PER_GPU = 4096
NUM_SIMS = 800
SEED = 0xA1FA0
METRIC = "MCTS sims/sec (connect-four self-play, 4096 lockstep games x 800 sims/move per GPU, uniform-prior evaluator)"
WORKLOAD = "config2: connect-four pure MCTS, uniform-prior evaluator, 4096 lockstep games x 800 sims/move, games to completion"


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def algorithmic_bytes_per_sim(levels_per_sim, expansions_per_sim, evals_ext_per_sim=0.0):
    """SURVEY §8(d) / DESIGN.md: one 128-B block read + 8-B counter read + 8-B counter write per
    level; per expansion one 128-B block write + 16-B slot update + 2x16-B table probe/insert;
    per externally evaluated leaf 336-B features out + 32-B (pi, v) in (0 for fused evaluators)."""
    return levels_per_sim * (128 + 16) + expansions_per_sim * (128 + 16 + 32) + evals_ext_per_sim * (336 + 32)


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed window by a thread of this process through
    NVML (pynvml): a polling nvidia-smi side process was measured to stall the CUDA calls of the step
    it overlaps (tens of milliseconds of e2e time per step), in-process NVML reads do not."""
    PERIOD_S = 0.05

    def __init__(self, index):
        self.index, self.rows, self.h, self.on, self.thread = index, [], None, False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception as e:  # no NVML: report that, never fail the bench
            self.err = repr(e)

    def _loop(self):
        nv = self.nv
        while self.on:
            try:
                self.rows.append((nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM),
                                  nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)))
            except Exception:
                pass
            time.sleep(self.PERIOD_S)

    def window_open(self):
        if self.h is None:
            return
        self.on = True
        self.thread = threading.Thread(target=self._loop, daemon=True)
        self.thread.start()

    def window_close(self):
        self.on = False
        if self.thread:
            self.thread.join()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "err", "")]}
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown,
                 "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown,
                 "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        sm = [r[0] for r in self.rows]
        reasons = sorted(n for n, bit in names.items() if any(r[1] & bit for r in self.rows))
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(sm)}


def cpu_baseline_run(orc, seconds_target, threads):
    """Oracle self-play on `threads` host threads, one independent game per thread at a time
    (the reference's rayon episode pool).  Bounded sample: whole games until ~seconds_target."""
    probe = orc.bench_selfplay(threads, threads, num_sims=NUM_SIMS, seed=SEED, reserve=131072, first_game_id=0)
    per_round = max(probe["seconds"], 1e-3)
    rounds = max(1, min(1024, int(seconds_target / per_round)))
    n_games = threads * rounds
    r = orc.bench_selfplay(n_games, threads, num_sims=NUM_SIMS, seed=SEED, reserve=131072, first_game_id=0)
    r["n_games"] = n_games
    return r


def run_reference(args):
    """`--impl reference`: the reference's CPU path = the oracle port on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import __graft_entry__ as ge
    ge.build_oracle()
    import oracle_api as orc
    cores = os.cpu_count() or 1
    games_per_step = 16 * cores  # ~0.5 M simulations per core per step
    for _ in range(args.warmup):
        orc.bench_selfplay(cores, cores, num_sims=NUM_SIMS, seed=SEED, reserve=131072)
    sims = plies = 0
    secs = 0.0
    for k in range(args.steps):
        r = orc.bench_selfplay(games_per_step, cores, num_sims=NUM_SIMS, seed=SEED, reserve=131072,
                               first_game_id=k * games_per_step)
        sims += r["sims"]; plies += r["plies"]; secs += r["seconds"]
    value = sims / secs
    sample = f"{games_per_step} whole games per step (of the 4096-game batch), {args.steps} steps, one game per thread"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample, "games_per_sec": (games_per_step * args.steps) / secs,
                   "plies": plies, "quirk_profile": "sane"},
        "cpu_baseline": {"value": value, "unit": "sims/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "sims/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="azb200", choices=["azb200", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games per GPU (default: config 2)")
    ap.add_argument("--sims", type=int, default=NUM_SIMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    import __graft_entry__ as ge
    ge.build_product()
    azb = importlib.import_module("alphazero-rs_b200")
    if azb.device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device: libazb200 has no CPU fallback")

    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def reduce(x, op):
        return azb.sharding.reduce_scalar(dist, x, op, device="cuda" if dist is not None else None)

    import numpy as np
    coach = azb.Coach(num_sims=args.sims, seed=SEED, quirks=azb.PROFILE_SANE, evaluator=azb.EVAL_UNIFORM,
                      temp_threshold=15, cpuct=1, max_depth=1000, mcts_reserve_size=1000000, device=local_rank)
    G = args.games
    cap = G * 84  # 42 plies x 2 symmetries: the most samples a game can produce
    pinned = [azb.PinnedArray((cap, 2, 6, 7)), azb.PinnedArray((cap, 7)), azb.PinnedArray((cap,))]
    out = tuple(p.array for p in pinned)

    def step(k):
        """One pass of the hot path through the public API, host buffers out."""
        first, _ = azb.sharding.shard(k, rank, world, G)  # global game ids: disjoint per rank and per step
        t0 = time.perf_counter()
        st = coach.self_play(G, first)
        _, _, vs = coach.export_samples(out)
        t1 = time.perf_counter()
        return st, t1 - t0, len(vs)

    sampler = ClockSampler(local_rank)
    for k in range(args.warmup):
        step(k)
    barrier()
    if rank == 0:
        sampler.window_open()
    dev_ms = wall = 0.0
    tot = {}
    n_samples = 0
    per_step = []
    for k in range(args.steps):
        st, w, ns = step(args.warmup + k)
        dev_ms += st["device_ms"]; wall += w; n_samples += ns
        per_step.append((round(st["device_ms"], 2), round(1e3 * w, 2)))
        for key, v in st.items():
            if key not in ("device_ms", "blocks_used_max", "owners_max"):
                tot[key] = tot.get(key, 0) + v
        tot["blocks_used_max"] = max(tot.get("blocks_used_max", 0), st["blocks_used_max"])
    barrier()
    sampler.window_close()
    clocks = sampler.stop() if rank == 0 else None

    dev_ms_max = reduce(dev_ms, "MAX")
    wall_max = reduce(wall, "MAX")
    sims_all = reduce(tot["sims"], "SUM")
    games_all = reduce(tot["games"], "SUM")
    levels_all = reduce(tot["levels"], "SUM")
    exp_all = reduce(tot["expansions"], "SUM")
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    value = sims_all / (dev_ms_max * 1e-3)
    e2e = sims_all / wall_max
    L, X = levels_all / sims_all, exp_all / sims_all
    bps = algorithmic_bytes_per_sim(L, X)
    peak, peak_src = peaks()
    achieved = (tot["sims"] * bps) / (dev_ms * 1e-3) / 1e9  # this rank's kernel: GB/s
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["dram_bytes_per_launch"]
    except Exception:
        pass
    d2h = n_samples / args.steps * (84 + 7 + 1) * 4 + G * (4 + 4 + 32)
    h2d = 256 + G * 8
    line = {
        "metric": METRIC, "value": value, "unit": "sims/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32+u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "games_per_gpu": G, "sims_per_move": args.sims, "quirk_profile": "sane",
                   "games_per_sec": games_all / (dev_ms_max * 1e-3), "plies_per_game": tot["plies"] / tot["games"],
                   "levels_per_sim": L, "expansions_per_sim": X, "bytes_per_sim": bps,
                   "blocks_used_max": tot["blocks_used_max"],
                   "l2": "tree pools (~20 GB per GPU) are far larger than L2; no flush needed",
                   "parallelism": f"{world} GPU(s), independent games per GPU, no collective on the path"},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel": "k_selfplay<UNIFORM>",
                     "launch_ms": dev_ms / args.steps},
        "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * wall_max / args.steps,
                "per_step_ms_device_and_wall": per_step},
        "gpu_launches": 2 * args.steps,
        "clocks": clocks,
    }
    if world == 1:
        # secondary evidence (not the headline): the leaf evaluator's dense forward pass, the only
        # tensor-core work of the path (BASELINE config 3), timed with CUDA events inside the library
        try:
            batch, blocks = 8192, 6
            net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC, device=local_rank)
            ms = net.benchmark(batch, 20)
            flop = 2 * 42 * 18 * 128 + 2 * blocks * 2 * 42 * 1152 * 128 + 21504 + 1176 + 10752 + 5376 + 128
            tf = batch * flop / ms / 1e9
            try:
                pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
                tpeak, tsrc = float(pk["bf16_tflops"]), "measured (MEASURED_PEAKS.json bf16_tflops, burst)"
            except Exception:
                tpeak, tsrc = 2250.0, "fallback (nominal dense bf16)"
            line["nnet_forward"] = {"workload": "config3 leaf evaluator: ResNet-6x128 bf16 forward, batch 8192 resident positions",
                                    "ms_per_pass": ms, "positions_per_sec": batch / ms * 1e3,
                                    "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                                 "frac": tf / tpeak, "peak_source": tsrc,
                                                 "kernel": "k_conv3x3_tc3 (tcgen05 cta_group::2, resident weights, TMA tiles reused by all taps)"}}
        except Exception as e:  # never fail the headline line on the secondary measurement
            line["nnet_forward"] = {"error": repr(e)}
        # secondary evidence: BASELINE config 3 itself, one pass (8192 games x 400 sims to completion with that network as the
        # batched leaf evaluator; CUDA-event device time from the library).  evaluations = what the trees asked for;
        # network_rows = positions that went through the network (a position is evaluated once per call: in-round
        # de-duplication + evaluation cache); the games are identical with both switched off.
        try:
            c3 = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET, device=local_rank)
            st3 = c3.self_play(8192, 0)
            s3 = st3["device_ms"] * 1e-3
            line["config3"] = {"workload": "connect-four self-play, random-init ResNet-6x128 bf16 tcgen05 leaf evaluator, 8192 games x 400 sims",
                               "device_s": s3, "sims_per_sec": st3["sims"] / s3, "games_per_sec": st3["games"] / s3,
                               "evaluations": st3["evals"], "network_rows": st3["nn_positions"], "cache_hits": st3["nn_cache_hits"],
                               "rounds": st3["launches"] // 3, "plies": st3["plies"]}
            c3.close()
        except Exception as e:
            line["config3"] = {"error": repr(e)}
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only
        ge.build_oracle()
        import oracle_api as orc
        cores = os.cpu_count() or 1
        r = cpu_baseline_run(orc, 25.0, cores)
        line["cpu_baseline"] = {"value": r["sims"] / r["seconds"], "unit": "sims/s", "cores": cores, "kind": "port",
                                "sample": f"{r['n_games']} whole games of the same workload (of 4096), one game per thread, "
                                          f"{r['seconds']:.1f} s; evaluator inline (no channel round trip)"}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

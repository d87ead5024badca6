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

Secondary keys of the same JSON line (SURVEY 8d / BASELINE.md section 3; none of them is the headline):
  parity_checked              games of the LAST timed step replayed by the oracle after the timed region (bit-exact check)
  cpu_baseline.one_thread     the oracle on one thread (the reference's deterministic mode, coach.rs:202-205)
  config1                     1 game x 50 (and 25) sims: oracle on one thread next to the device, uniform evaluator and a
                              small random-init network
  config2_reference_profile   the headline workload under the literal quirk profile (AZB_PROFILE_REFERENCE)
  nnet_forward, config3       the leaf evaluator alone, and BASELINE config 3 (8192 x 400, ResNet-6x128 bf16 tcgen05) with
                              e2e (samples to page-locked host memory), whole-run tensor roofline and a CPU figure (oracle +
                              plain C++ fp32 forward, bounded sample)
  config4, config5            arena (two networks) and one full Coach::learn iteration; at N = 1 one GPU's share of the
                              8-GPU configs, at N > 1 the games split over the ranks (counters / gradients reduced with NCCL)

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


NET_FLOP = lambda blocks: 2 * 42 * 18 * 128 + 2 * blocks * 2 * 42 * 1152 * 128 + 21504 + 1176 + 10752 + 5376 + 128


def tensor_peaks():
    try:
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(pk["bf16_tflops"]), float(pk.get("bf16_tflops_sustained", pk["bf16_tflops"])), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 1590.0, 1400.0, "fallback (B200_PROFILING.md 1.59 / 1.4 PFLOP/s)"


def replay_parity(orc, traces, n_games, games, **okw):
    """Bit-exact replay of sampled games by the oracle: plies, actions, per-ply root counts.  Returns the number checked;
    raises on the first mismatch (a fast kernel with different results is not a result)."""
    idx = sorted(set(int(i) for i in np_linspace(0, n_games - 1, games)))
    for g in idx:
        o = orc.execute_episode(episode_id=okw["first"] + g, **{k: v for k, v in okw.items() if k != "first"})
        n = o["plies"]
        if not (traces["plies"][g] == n and traces["actions"][g, :n].tolist() == o["actions"][:n].tolist()
                and (traces["counts"][g, :n] == o["counts"][:n]).all()):
            raise SystemExit(f"PARITY FAILURE: game {g} of the timed step differs from the oracle")
    return len(idx)


def np_linspace(a, b, n):
    import numpy as np
    return np.linspace(a, b, n).round().astype(int)


def section_config1(azb, orc, device):
    """BASELINE config 1 (examples/connect_four.rs:55-71): ONE self-play game, 50 sims/move (and the file's 25), one search
    thread.  CPU = the oracle on one thread; device = the same game through the public API (a single warp: latency-bound,
    this is the configuration the device is worst at)."""
    out = {"workload": "config1: 1 self-play game, 1 search thread, seed 1, cpuct 1, temp_threshold 15"}
    for sims in (50, 25):
        cpu = orc.bench_selfplay(1, 1, num_sims=sims, seed=1, reserve=131072, first_game_id=0)
        coach = azb.Coach(num_sims=sims, seed=1, evaluator=azb.EVAL_UNIFORM, device=device)
        coach.self_play(1, 0)
        t0 = time.perf_counter(); st = coach.self_play(1, 0); wall = time.perf_counter() - t0
        out[f"uniform_{sims}_sims"] = {"cpu_1_thread_sims_per_sec": cpu["sims"] / cpu["seconds"], "cpu_plies": cpu["plies"],
                                       "device_sims_per_sec": st["sims"] / (st["device_ms"] * 1e-3), "device_ms": st["device_ms"],
                                       "e2e_ms": 1e3 * wall, "plies": st["plies"]}
        coach.close()
    # "random-init small net": two residual blocks; CPU = oracle + plain C++ fp32 forward per leaf (batch 1, like the
    # reference's inference_batch_size = 1), device = lock-step rounds with the bf16 tensor-core forward
    net = azb.NNet(seed=7, blocks=2, precision=azb.NNET_BF16_TC, device=device)
    cpu = orc.bench_selfplay_net(net.get_params(), 2, 1, 1, num_sims=50, seed=1, first_game_id=0, max_plies=4, reserve=131072)
    coach = azb.Coach(nnet=net, num_sims=50, seed=1, evaluator=azb.EVAL_NNET, device=device)
    coach.self_play(1, 0)
    t0 = time.perf_counter(); st = coach.self_play(1, 0); wall = time.perf_counter() - t0
    out["small_net_50_sims"] = {"network": "2 residual blocks x 128 channels, random init seed 7",
                                "cpu_1_thread_sims_per_sec": cpu["sims"] / cpu["seconds"],
                                "cpu_sample": f"first {cpu['plies']} plies of the game, {cpu['evals']} fp32 forwards, {cpu['seconds']:.1f} s",
                                "device_sims_per_sec": st["sims"] / (st["device_ms"] * 1e-3), "device_ms": st["device_ms"],
                                "e2e_ms": 1e3 * wall, "plies": st["plies"]}
    coach.close(); net.close()
    return out


def section_reference_profile(azb, device, games, sims):
    """The headline workload under the literal quirk profile (Q1-Q4 as the reference's code behaves): Q2 (no
    sign alternation in the backup) makes games degenerate — 7-21 plies — which is why the headline uses `sane`."""
    coach = azb.Coach(num_sims=sims, seed=SEED, quirks=azb.PROFILE_REFERENCE, evaluator=azb.EVAL_UNIFORM, device=device)
    coach.self_play(games, 0)
    st = coach.self_play(games, games)
    s = st["device_ms"] * 1e-3
    coach.close()
    return {"workload": "config2 under AZB_PROFILE_REFERENCE (literal quirks Q1-Q4)", "ms": st["device_ms"],
            "sims_per_sec": st["sims"] / s, "games_per_sec": st["games"] / s, "plies_per_game": st["plies"] / st["games"],
            "levels_per_sim": st["levels"] / st["sims"]}


def section_config3(azb, orc, device, net, cpu_seconds):
    """BASELINE config 3: 8192 games x 400 sims to completion with the ResNet-6x128 bf16 tcgen05 evaluator.  device_s = CUDA
    events inside the library; e2e = self-play + export of the SOA samples into page-locked host memory, wall clock;
    roofline = network rows x 148.87 MFLOP / device_s against the measured SUSTAINED bf16 peak (a seconds-long step).
    evaluations = what the trees asked for; network_rows = positions that went through the network (a position is evaluated
    once per call: in-round de-duplication + evaluation cache)."""
    import numpy as np
    G, sims = 8192, 400
    c3 = azb.Coach(nnet=net, num_sims=sims, seed=SEED, evaluator=azb.EVAL_NNET, device=device)
    cap = G * 84
    pinned = [azb.PinnedArray((cap, 2, 6, 7)), azb.PinnedArray((cap, 7)), azb.PinnedArray((cap,))]
    t0 = time.perf_counter()
    st = c3.self_play(G, 0)
    _, _, vs = c3.export_samples(tuple(p.array for p in pinned))
    wall = time.perf_counter() - t0
    s = st["device_ms"] * 1e-3
    burst, sustained, src = tensor_peaks()
    tf = st["nn_positions"] * NET_FLOP(6) / s / 1e12
    out = {"workload": "config3: connect-four self-play, random-init ResNet-6x128 bf16 tcgen05 leaf evaluator, 8192 games x 400 sims",
           "device_s": s, "sims_per_sec": st["sims"] / s, "games_per_sec": st["games"] / s,
           "evaluations": st["evals"], "network_rows": st["nn_positions"], "cache_hits": st["nn_cache_hits"],
           "kernel_launches": st["launches"], "plies": st["plies"],
           "e2e": {"value": st["sims"] / wall, "unit": "sims/s", "seconds": wall,
                   "d2h_bytes": int(len(vs) * 92 * 4 + G * 40), "h2d_bytes": 256 + G * 8},
           "roofline": {"bound": "tensor", "achieved": tf, "peak": sustained, "unit": "TFLOP/s", "frac": tf / sustained,
                        "frac_of_burst": tf / burst, "peak_source": src + " bf16_tflops_sustained (whole-run figure)",
                        "traffic": None, "kernel": "whole run: network rows x 148.87 MFLOP / device seconds"}}
    # parity of sampled games against the oracle searching with the same network (bit-exact; test_fullsize_parity_gpu.py
    # does 8 games, here 2 keep the bench short)
    tr = c3.traces()
    out["parity_checked"] = replay_parity(orc, tr, G, 2, first=0, num_sims=sims, quirks=0, seed=SEED,
                                          evaluator=orc.EVAL_CALLBACK, callback=net.predict)
    if cpu_seconds > 0:
        cores = os.cpu_count() or 1
        r = orc.bench_selfplay_net(net.get_params(), 6, cores, cores, num_sims=sims, seed=SEED, first_game_id=0, max_plies=1,
                                   reserve=131072)
        out["cpu_baseline"] = {"value": r["sims"] / r["seconds"], "unit": "sims/s", "cores": cores, "kind": "port",
                               "sample": f"first ply (400 sims) of {cores} games, one game per thread, {r['evals']} plain C++ fp32 "
                                         f"forwards called inline per leaf (oracle/nnet_cpu.hpp), {r['seconds']:.1f} s"}
    for p in pinned:
        p.close()
    c3.close()
    return out


def section_config4(azb, dist_mod, rank, world, device, reduce):
    """BASELINE config 4: two random-init networks (seeds 7 and 8) head to head.  N > 1: the 16384 games of the config split
    over the ranks in seat-order pairs, Win/Loss/Draw summed (no other collective); N = 1: one GPU's share (2048 games).
    400 sims/move, 4 random opening plies per game (with none, every game of a seat order is the same game)."""
    total = 16384 if world > 1 else 2048
    first, n = azb.sharding.split_total(total // 2, rank, world)  # pairs
    a = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC, device=device)
    b = azb.NNet(seed=8, blocks=6, precision=azb.NNET_BF16_TC, device=device)
    t0 = time.perf_counter()
    counts, res, st, _ = azb.arena_play_games_traced(2 * n, azb.EVAL_NNET, azb.EVAL_NNET, a, b, k_open=4, first_game_id=2 * first,
                                                     num_sims=400, seed=SEED, device=device)
    wall = time.perf_counter() - t0
    dev_s = reduce(st["device_ms"] * 1e-3, "MAX")
    wall = reduce(wall, "MAX")
    wld = [int(reduce(c, "SUM")) for c in counts]
    sims = reduce(st["sims"], "SUM")
    a.close(); b.close()
    return {"workload": f"config4: arena, nets seeded 7 vs 8, {total} games over {world} GPU(s), 400 sims, 4 random opening plies",
            "games": total, "device_s_max_over_ranks": dev_s, "e2e_s": wall, "games_per_sec": total / dev_s,
            "sims_per_sec": sims / dev_s, "win_loss_draw_of_net_7": wld}


def section_config5(azb, dist_mod, rank, world, device, tmpdir):
    """BASELINE config 5: one full Coach::learn iteration through ONE call per rank (azb_coach_learn[_dist]): 8192 self-play
    games per GPU x 400 sims, one pass of Adam steps at global batch 4096 x N over the exported samples with the gradients
    all-reduced over NCCL, 2048 gating games per GPU, history and weights written to disk."""
    coach = azb.Coach(checkpoint_directory=os.path.join(tmpdir, f"rank{rank}").encode(), num_eps=8192 * world, num_sims=400,
                      num_arena_games=2048 * world, num_iters=1, seed=SEED, evaluator=azb.EVAL_NNET, device=device,
                      max_queue_length=10 ** 9, update_threshold=0.55)
    t0 = time.perf_counter()
    reports, net = coach.learn(epochs=0, batch_size=4096 * world, lr=1e-4, seed=7, blocks=6, dist=dist_mod if world > 1 else None)
    wall = time.perf_counter() - t0
    r = reports[0]
    net.close(); coach.close()
    return {"workload": f"config5: Coach::learn iteration, {8192 * world} games x 400 sims over {world} GPU(s) + data-parallel "
                        f"training (global batch {4096 * world}) + {2048 * world} gating games",
            "wall_s_rank0": wall, "selfplay_s": r["selfplay_ms"] * 1e-3, "train_s": r["train_ms"] * 1e-3,
            "arena_s": r["arena_ms"] * 1e-3, "games_per_sec_selfplay": 8192 * world / (r["selfplay_ms"] * 1e-3),
            "train_steps": r["train_steps"], "samples_rank0": r["samples_played"], "loss_first": r["loss_first"],
            "loss_last": r["loss_last"], "nwins_pwins_draws": [r["nwins"], r["pwins"], r["draws"]], "accepted": r["accepted"]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="azb200", choices=["azb200", "reference"])
    ap.add_argument("--games", type=int, default=GAMES_PER_GPU, help="games per GPU (default: config 2)")
    ap.add_argument("--sims", type=int, default=NUM_SIMS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="one step at a time (no overlap of consecutive batches)")
    ap.add_argument("--depth", type=int, default=3, help="steps in flight (coaches with their own trees and streams)")
    ap.add_argument("--no-secondary", action="store_true", help="headline line only (config 2): skip configs 1, 3, 4, 5")
    ap.add_argument("--parity-games", type=int, default=16, help="games of the last timed step replayed by the oracle")
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

    # N > 1: one process per GPU (torchrun provides RANK / LOCAL_RANK / WORLD_SIZE / MASTER_*).  The ranks talk through the
    # library's own NCCL communicator (csrc/dist.cuh, azb_dist_*): no torch in the data plane or in the bench plumbing.
    # The 128-byte NCCL unique id goes from rank 0 to the others through a file named after the launch (the launcher's
    # pid is the parent of every rank).  AZB200_BENCH_TORCH=1 falls back to torch.distributed for the plumbing.
    dist = None
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if os.environ.get("AZB200_BENCH_TORCH"):
            import torch
            import torch.distributed as dist
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        else:
            id_path = f"/tmp/azb200_nccl_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}.id"
            comm = azb.Comm(rank, world, local_rank, id_path)

    def barrier():
        if comm is not None:
            comm.barrier()
        elif dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    def reduce(x, op):
        if comm is not None:
            return comm.reduce(x, op)
        return azb.sharding.reduce_scalar(dist, x, op, device="cuda" if dist is not None else None)

    import numpy as np
    # Steps are enqueued `depth` (3) DEEP on as many coaches (each with its own trees, buffers and CUDA stream): begin(k + 2) is
    # called before end(k), so the warps that batch k's finished games vacate are taken by batch k + 1's games instead of idling
    # until batch k's longest game ends (a batch of 4096 games ends with its 42-ply stragglers while the mean game has 28
    # plies).  Every step is still one batch of 4096 games played to completion by its own launch; at most one SM-ful of
    # games (28 warps per SM) is in flight at any time.  The CPU arm has no such barrier either (its threads take the
    # next game when one ends).  --no-pipeline runs the steps one after the other (round 1's loop).
    mk = lambda: azb.Coach(num_sims=args.sims, seed=SEED, quirks=azb.PROFILE_SANE, evaluator=azb.EVAL_UNIFORM,
                           temp_threshold=15, cpuct=1, max_depth=1000, mcts_reserve_size=1000000, device=local_rank)
    depth = 1 if args.no_pipeline else max(1, args.depth)
    coaches = [mk() for _ in range(depth)]
    coach = coaches[0]
    G = args.games
    cap = G * 84  # 42 plies x 2 symmetries: the most samples a game can produce
    pinned = [[azb.PinnedArray((cap, 2, 6, 7)), azb.PinnedArray((cap, 7)), azb.PinnedArray((cap,))] for _ in range(depth)]
    outs = [tuple(p.array for p in ps) for ps in pinned]
    first_of = lambda k: azb.sharding.shard(k, rank, world, G)[0]  # global game ids: disjoint per rank and per step

    def run_steps(k0, n, export):
        """n steps (k0 ...) through the public API, `depth` in flight.  export: the SOA samples of every step are copied into
        page-locked host buffers.  Returns (per-step stats, samples exported, wall seconds, device-span ms)."""
        stats, n_samples = [], 0
        coaches[0].span_mark()
        t0 = time.perf_counter()
        for j in range(min(depth - 1, n)):
            coaches[j % depth].self_play_begin(G, first_of(k0 + j))
        for j in range(n):
            nxt = j + depth - 1
            if nxt < n:
                coaches[nxt % depth].self_play_begin(G, first_of(k0 + nxt))
            c = coaches[j % depth]
            stats.append(c.self_play_end())
            if export:
                _, _, vs = c.export_samples(outs[j % depth])
                n_samples += len(vs)
        wall = time.perf_counter() - t0
        span = coaches[0].span_ms(coaches[(n - 1) % depth])
        return stats, n_samples, wall, span

    sampler = ClockSampler(local_rank)
    run_steps(0, args.warmup, True)
    # one batch alone (no overlap): the latency of a single 4096-game call
    st1 = coaches[0].self_play(G, first_of(args.warmup))
    single_batch_ms = st1["device_ms"]
    barrier()
    if rank == 0:
        sampler.window_open()
    # (1) device-timed: inputs resident, nothing copied out between steps, CUDA events from the first launch to the last end
    k_dev = args.warmup + 1
    stats_dev, _, _, span_ms = run_steps(k_dev, args.steps, False)
    barrier()
    # (2) end to end: the same K steps with every step's samples exported to host memory, wall clock
    k_e2e = k_dev + args.steps
    stats_e2e, n_samples, wall, _ = run_steps(k_e2e, args.steps, True)
    barrier()
    sampler.window_close()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = span_ms
    tot = {}
    for st in stats_dev:
        for key, v in st.items():
            if key not in ("device_ms", "blocks_used_max", "owners_max"):
                tot[key] = tot.get(key, 0) + v
        tot["blocks_used_max"] = max(tot.get("blocks_used_max", 0), st["blocks_used_max"])
    sims_e2e = sum(st["sims"] for st in stats_e2e)
    launches = sum(st["launches"] for st in stats_dev) + sum(st["launches"] + 1 for st in stats_e2e)
    per_step = [round(st["device_ms"], 2) for st in stats_dev]
    coach = coaches[(args.steps - 1) % depth]  # holds the last timed step's games
    # parity of what was just timed: games of the LAST timed step (this rank's) replayed by the oracle, bit for bit
    parity_checked = 0
    if rank == 0 and args.parity_games > 0:
        ge.build_oracle()
        import oracle_api as orc
        last_first = first_of(k_e2e + args.steps - 1)
        parity_checked = replay_parity(orc, coach.traces(), G, args.parity_games, first=last_first, num_sims=args.sims,
                                       quirks=0, seed=SEED, evaluator=orc.EVAL_UNIFORM)

    dev_ms_max = reduce(dev_ms, "MAX")
    wall_max = reduce(wall, "MAX")
    sims_all = reduce(tot["sims"], "SUM")
    games_all = reduce(tot["games"], "SUM")
    levels_all = reduce(tot["levels"], "SUM")
    exp_all = reduce(tot["expansions"], "SUM")
    sims_e2e_all = reduce(sims_e2e, "SUM")
    secondary = {}
    if not args.no_secondary and world > 1:  # collective sections: every rank takes part
        import tempfile
        for name, fn in (("config4", lambda: section_config4(azb, dist, rank, world, local_rank, reduce)),
                         ("config5", lambda: section_config5(azb, comm if comm is not None else dist, rank, world, local_rank,
                                                             tempfile.mkdtemp(prefix="azb200_c5_")))):
            try:
                secondary[name] = fn()
            except Exception as e:  # never fail the headline line on a secondary measurement
                secondary[name] = {"error": repr(e)}
            barrier()
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        if comm is not None:
            comm.close()
        return

    value = sims_all / (dev_ms_max * 1e-3)
    e2e = sims_e2e_all / wall_max
    L, X = levels_all / sims_all, exp_all / sims_all
    bps = algorithmic_bytes_per_sim(L, X)
    peak, peak_src = peaks()
    achieved = (tot["sims"] * bps) / (dev_ms * 1e-3) / 1e9  # this rank's kernel: GB/s
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        traffic, traffic_src = tj["dram_bytes_per_launch"], "STATIC, not measured by this run: " + tj["source"]
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
                   "l2": "tree pools (~20 GB per coach) are far larger than L2; no flush needed",
                   "pipelining": ("none: one step at a time" if depth == 1 else
                                  f"steps enqueued {depth} deep on {depth} coaches with their own trees and CUDA streams "
                                  "(azb_coach_self_play_begin / _end): the next batch's games take the warps the current batch's "
                                  "finished games vacate; every step is one batch of 4096 games played to completion by its own launch"),
                   "single_batch_ms": single_batch_ms,
                   "launch_ms_mean": sum(per_step) / max(1, len(per_step)),
                   "parallelism": f"{world} GPU(s), independent games per GPU, no collective on the path" +
                                  ("; ranks synchronised through the library's own NCCL communicator (no torch)" if comm is not None else "")},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src, "kernel": "k_selfplay<UNIFORM>",
                     "launch_ms": dev_ms / args.steps,
                     "how": "algorithmic bytes of the timed region / its device span (CUDA events from the first launch to the last "
                            "launch's end); consecutive launches overlap on the device, so a per-launch duration (launch_ms_mean in "
                            "config) would count the shared time twice"},
        "e2e": {"value": e2e, "unit": "sims/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "bytes_how": "computed from the sizes of the arrays the call copies (samples x 92 f32 + per-game plies / errors / "
                             "statistics; config struct + game ids in), not measured on the bus",
                "ms_per_step": 1e3 * wall_max / args.steps,
                "per_launch_device_ms": per_step},
        "gpu_launches": launches,
        "gpu_launches_how": "the library's own launch counter per self-play call (azb_selfplay_stats.launches: one persistent "
                            "k_selfplay) + one k_export_samples per step, rank 0",
        "parity_checked": parity_checked,
        "clocks": clocks,
    }
    line.update(secondary)
    if world == 1 and not args.no_secondary:
        import tempfile
        ge.build_oracle()
        import oracle_api as orc
        for c_ in coaches:
            c_.close()
        for ps in pinned:
            for p_ in ps:
                p_.close()
        net = None
        # the leaf evaluator's dense forward pass, the only tensor-core work of the path, timed with CUDA events inside the library
        try:
            batch, blocks = 8192, 6
            net = azb.NNet(seed=7, blocks=blocks, precision=azb.NNET_BF16_TC, device=local_rank)
            ms = net.benchmark(batch, 20)
            tf = batch * NET_FLOP(blocks) / ms / 1e9
            tpeak, _, tsrc = tensor_peaks()
            line["nnet_forward"] = {"workload": "config3 leaf evaluator: ResNet-6x128 bf16 forward, batch 8192 resident positions",
                                    "ms_per_pass": ms, "positions_per_sec": batch / ms * 1e3,
                                    "roofline": {"bound": "tensor", "achieved": tf, "peak": tpeak, "unit": "TFLOP/s",
                                                 "frac": tf / tpeak, "peak_source": tsrc + " bf16_tflops (burst: kernel timed alone)",
                                                 "traffic": None,
                                                 "kernel": "k_conv3x3_tc3 (tcgen05 cta_group::2, resident weights, TMA tiles reused by all taps)"}}
        except Exception as e:
            line["nnet_forward"] = {"error": repr(e)}
        sections = [("config1", lambda: section_config1(azb, orc, local_rank)),
                    ("config2_reference_profile", lambda: section_reference_profile(azb, local_rank, G, args.sims)),
                    ("config3", lambda: section_config3(azb, orc, local_rank, net, 0 if args.no_cpu_baseline else 15.0)),
                    ("config4", lambda: section_config4(azb, None, 0, 1, local_rank, lambda x, op: x)),
                    ("config5", lambda: section_config5(azb, None, 0, 1, local_rank, tempfile.mkdtemp(prefix="azb200_c5_")))]
        for name, fn in sections:
            try:
                line[name] = fn()
            except Exception as e:
                line[name] = {"error": repr(e)}
    if not args.no_cpu_baseline and world == 1:  # rank 0 at N = 1 only
        ge.build_oracle()
        import oracle_api as orc
        cores = os.cpu_count() or 1
        r = cpu_baseline_run(orc, 25.0, cores)
        one = orc.bench_selfplay(2, 1, num_sims=args.sims, seed=SEED, reserve=131072, first_game_id=0)
        line["cpu_baseline"] = {"value": r["sims"] / r["seconds"], "unit": "sims/s", "cores": cores, "kind": "port",
                                "sample": f"{r['n_games']} whole games of the same workload (of 4096), one game per thread, "
                                          f"{r['seconds']:.1f} s; evaluator inline (no channel round trip)",
                                "one_thread": {"value": one["sims"] / one["seconds"], "unit": "sims/s", "cores": 1,
                                               "sample": f"2 whole games of the same workload on one thread (the reference's "
                                                         f"deterministic mode, coach.rs:202-205), {one['seconds']:.1f} s"}}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    if comm is not None:
        comm.close()


if __name__ == "__main__":
    main()

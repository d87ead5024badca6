"""BASELINE configs 4 and 5 (self-play part) at full width on N GPUs of one box, one process per GPU, no collective on
the path (games are sharded by global id; torch.distributed only for the barrier and the reductions of the numbers):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_multi.py
config 5 self-play: 8192 games per GPU x 400 sims with the ResNet-6x128 bf16 evaluator (65536 games on 8 GPUs);
config 4: 16384 arena games split over the ranks (2048 per GPU on 8 GPUs), nets seeded 7 / 8, 4 random opening plies."""
import importlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
azb = importlib.import_module("alphazero-rs_b200")
d = dist if world > 1 else None
red = lambda x, op: azb.sharding.reduce_scalar(d, x, op, device="cuda" if d else None)
def barrier():
    if d: d.barrier(); torch.cuda.synchronize()

net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC, device=local)
net_b = azb.NNet(seed=8, blocks=6, precision=azb.NNET_BF16_TC, device=local)
coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET, device=local)
first, n = azb.sharding.shard(0, rank, world, 8192)
barrier()
st = coach.self_play(n, first)
barrier()
t = red(st["device_ms"], "MAX") * 1e-3
games, sims, evals = red(st["games"], "SUM"), red(st["sims"], "SUM"), red(st["evals"], "SUM")
if rank == 0:
    print(json.dumps({"workload": "config5 self-play part: 8192 games per GPU x 400 sims, ResNet-6x128 bf16 evaluator", "n_gpus": world,
                      "games": games, "device_s_max_over_ranks": t, "games_per_sec": games / t, "sims_per_sec": sims / t,
                      "leaf_evals_per_sec": evals / t}))
first, n = azb.sharding.split_total(16384, rank, world)
barrier()
counts, res, st = azb.arena_play_games(n, azb.EVAL_NNET, azb.EVAL_NNET, net, net_b, k_open=4, num_sims=400,
                                       seed=0xA1FA0 + rank, device=local)  # (the arena ABI keys streams by seed: one seed per rank)
barrier()
t = red(st["device_ms"], "MAX") * 1e-3
w, l, dr = (red(int(c), "SUM") for c in counts)
if rank == 0:
    print(json.dumps({"workload": "config4: 16384 arena games split over the ranks, nets 7 vs 8, 400 sims, 4 random opening plies",
                      "n_gpus": world, "device_s_max_over_ranks": t, "games_per_sec": (w + l + dr) / t, "win_loss_draw_of_net_7": [w, l, dr]}))
if d: dist.destroy_process_group()

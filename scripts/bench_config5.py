"""BASELINE config 5: one coach.rs iteration — self-play (8192 games per GPU x 400 sims with the ResNet-6x128 bf16 evaluator),
export of the finished samples, then data-parallel training steps on them (NNet::train: one NCCL all-reduce of the 1.8 M fp32
gradients per step).  One process per GPU:
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/bench_config5.py [games_per_gpu] [train_batch] [max_steps]"""
import importlib, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
azb = importlib.import_module("alphazero-rs_b200")
d = dist if world > 1 else None
games = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
max_steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1000
red = lambda x, op: azb.sharding.reduce_scalar(d, x, op, device="cuda" if d else None)
def barrier():
    if d: d.barrier()
    torch.cuda.synchronize()

net = azb.NNet(seed=7, blocks=6, precision=azb.NNET_BF16_TC, device=local)
coach = azb.Coach(nnet=net, num_sims=400, seed=0xA1FA0, evaluator=azb.EVAL_NNET, device=local)
first, n = azb.sharding.shard(0, rank, world, games)
barrier(); t0 = time.perf_counter()
st = coach.self_play(n, first)
boards, pis, vs = coach.export_samples()
barrier(); t_play = time.perf_counter() - t0
# one epoch over this iteration's samples, every rank on its own games' samples; equal step counts on all ranks
steps = int(red(len(vs) // batch, "MIN"))
steps = min(steps, max_steps)
perm = np.random.default_rng(rank).permutation(len(vs))
losses = []
warm = min(3, steps)  # the first steps carry NCCL's lazy communicator set-up and first-touch allocations: timed apart
t_warm = time.perf_counter()
for k in range(steps):
    if k == warm:
        barrier(); t_warm = time.perf_counter() - t_warm; t0 = time.perf_counter()
    idx = perm[k * batch:(k + 1) * batch]
    losses.append(net.train((boards[idx], pis[idx], vs[idx]), lr=1e-3, dist=d))
barrier(); t_train = time.perf_counter() - t0
timed = max(steps - warm, 1)
games_all, samples_all = red(st["games"], "SUM"), red(len(vs), "SUM")
p0 = net.get_params()
chk = red(float(np.abs(p0).sum()), "MAX") - red(float(np.abs(p0).sum()), "MIN")  # replicas stay identical
if rank == 0:
    print(json.dumps({"workload": "config5: self-play + data-parallel training on the exported samples", "n_gpus": world,
                      "games": games_all, "selfplay_and_export_s": t_play, "games_per_sec": games_all / t_play, "samples": samples_all,
                      "train_steps": steps, "global_batch": batch * world, "train_s_after_warmup": t_train, "warmup_steps": warm,
                      "trained_samples_per_sec": timed * batch * world / max(t_train, 1e-9), "ms_per_step": 1e3 * t_train / timed,
                      "loss_first": [round(x, 4) for x in losses[0]] if losses else None,
                      "loss_last": [round(x, 4) for x in losses[-1]] if losses else None, "replica_param_spread": chk}))
if d: dist.destroy_process_group()

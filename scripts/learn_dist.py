"""Coach::learn data parallel over one process per GPU (azb_coach_learn_dist; BASELINE config 5 as ONE call per rank):
  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/learn_dist.py [iters] [num_eps] [sims] [batch] [arena] [epochs] [lr] [blocks]
Env: AZB_DIST_BACKEND=gloo AZB_DIST_ONE_GPU=1 runs all ranks on cuda:0 over gloo (the single-GPU test of the same code
path); AZB_DIST_OUT=<dir> writes rank<r>.json (reports, parameter checksum, first-iteration history) for the test."""
import importlib, json, os, shutil, sys, tempfile, time, zlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist
rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
os.environ.setdefault("MASTER_PORT", "29531")
backend = os.environ.get("AZB_DIST_BACKEND", "nccl")
dev = 0 if os.environ.get("AZB_DIST_ONE_GPU") else local
torch.cuda.set_device(dev)
if backend == "nccl":
    dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
else:
    dist.init_process_group(backend)
azb = importlib.import_module("alphazero-rs_b200")
arg = lambda i, d, t=int: t(sys.argv[i]) if len(sys.argv) > i else d
iters, eps, sims, batch, arena, epochs = arg(1, 2), arg(2, 8192 * world), arg(3, 400), arg(4, 4096 * world), arg(5, 512), arg(6, 0)
lr, blocks = arg(7, 1e-4, float), arg(8, 6)
out = os.environ.get("AZB_DIST_OUT")
ckpt = os.path.join(out, f"ckpt_rank{rank}") if out else tempfile.mkdtemp(prefix=f"azb_ckpt_r{rank}_")
coach = azb.Coach(checkpoint_directory=ckpt.encode(), evaluator=azb.EVAL_NNET, num_iters=iters, num_eps=eps, num_sims=sims,
                  num_arena_games=arena, max_queue_length=int(os.environ.get("AZB_DIST_QUEUE", 4_000_000)), max_history_length=4,
                  update_threshold=0.55, seed=0xA1FA0, device=dev, temp_threshold=int(os.environ.get("AZB_DIST_TEMP_THRESHOLD", 15)))
dist.barrier(); torch.cuda.synchronize(); t0 = time.perf_counter()
reports, net = coach.learn(epochs=epochs, batch_size=batch, blocks=blocks, seed=7, arena_k_open=4, lr=lr, dist=dist)
dist.barrier(); torch.cuda.synchronize(); wall = time.perf_counter() - t0
params = net.get_params()
crc = zlib.crc32(params.tobytes())
crcs = [None] * world
dist.all_gather_object(crcs, crc)
if out:
    c0, b0, p0, v0 = azb.examples_read(os.path.join(ckpt, "0.examples"))
    np.savez(os.path.join(out, f"rank{rank}.npz"), boards=b0, pis=p0, vs=v0, counts=c0, params=params)
    json.dump({"reports": reports, "crc": crc}, open(os.path.join(out, f"rank{rank}.json"), "w"))
if rank == 0:
    for r in reports:
        print(json.dumps({k: (round(v, 1) if k.endswith("_ms") else [round(x, 4) for x in v] if isinstance(v, list) else v) for k, v in r.items()}))
    print(json.dumps({"workload": f"Coach::learn data parallel: {iters} iterations x {eps} games x {sims} sims over {world} ranks ({backend}), "
                      f"ResNet-{blocks}x128 bf16, global batch {batch}, {arena} arena games", "n_gpus": world, "wall_s": round(wall, 2),
                      "games_per_rank_rank0": reports[0]["games"], "replicas_identical": len(set(crcs)) == 1,
                      "final_model_id": reports[-1]["model_id_after"]}))
if not out:
    shutil.rmtree(ckpt, ignore_errors=True)
dist.destroy_process_group()

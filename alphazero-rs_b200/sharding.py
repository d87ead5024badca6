"""Multi-GPU layout of the self-play path: games are independent units (every game owns its tree
and its Philox stream), so they are sharded over ranks with NO collective on the path.
One process per GPU; torch.distributed is used only for the bench barrier and the max/sum reductions
of the reported numbers."""


def shard(step, rank, world, games_per_gpu):
    """Global game ids of rank `rank` in step `step`: disjoint across ranks and steps (weak scaling:
    every GPU plays `games_per_gpu` games per step).  Returns (first_game_id, n_games)."""
    return (step * world + rank) * games_per_gpu, games_per_gpu


def split_total(total_games, rank, world):
    """Strong-scaling split of a fixed number of games (arena: 16384 games over 8 GPUs): contiguous
    ranges, the remainder spread over the first ranks.  Returns (first_game_id, n_games)."""
    base, rem = divmod(total_games, world)
    n = base + (1 if rank < rem else 0)
    first = rank * base + min(rank, rem)
    return first, n


def reduce_scalar(dist, value, op, device=None):
    """MAX / SUM of a python scalar over ranks (identity when dist is None)."""
    if dist is None:
        return value
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
    return float(t.item())


def average_gradients(dist, grads):
    """The one collective of the data-parallel training step: all-reduce the flat fp32 gradient vector
    (numpy) over the ranks and divide by the world size — every rank trained on its own equally sized shard, so the
    mean of the per-rank means is the gradient of the mean loss over the global batch.  Identity without `dist`.
    NCCL ranks reduce on their GPU (NVLink / NVSwitch), gloo ranks on the host."""
    if dist is None or dist.get_world_size() == 1:
        return grads
    import torch
    dev = "cuda" if dist.get_backend() == "nccl" else "cpu"
    g = torch.from_numpy(grads).to(dev)
    dist.all_reduce(g)
    g /= dist.get_world_size()
    return g.cpu().numpy()


def shard_samples(samples, rank, world):
    """Equal contiguous shards of SOATrainingSamples (boards, pis, vs); the remainder (< world samples) is dropped so
    that every rank's mean has the same weight."""
    n = len(samples[2]) // world
    return tuple(a[rank * n:(rank + 1) * n] for a in samples)

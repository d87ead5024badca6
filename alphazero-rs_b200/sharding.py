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

"""Multi-GPU host logic: one process per GPU, weights replicated, scenes (ETH/UCY/SDD) or whole
reference batches (NBA — the scenes of a batch attend to each other, so a batch never splits) sharded across ranks.
There is NO data-path collective: each rank forecasts its shard independently; the only exchange is one
all_reduce(SUM) of the three metric accumulators [sum minADE, sum minFDE, agents] in float64, mirroring
`AverageMeter.update(val, n)` (utils/utils.py:29-34) / the `* batch ... / all_num` sums of test.py:532-552.
Works with any torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous, balanced [start, stop) of `n_items` (scenes or NBA batches) for `rank`."""
    base, rem = divmod(n_items, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_scenes_by_cost(agent_counts, world):
    """Cost-balanced assignment for ETH/SDD scenes of very different sizes: greedy longest-processing-time on the
    agent count (decoder work is linear in agents).  Returns a list of index lists, one per rank."""
    order = sorted(range(len(agent_counts)), key=lambda i: -agent_counts[i])
    load, out = [0] * world, [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda k: load[k])
        out[r].append(i)
        load[r] += agent_counts[i]
    return [sorted(x) for x in out]


def reduce_metric_sums(sums):
    """In-place all_reduce(SUM) of a float64 tensor [sum minADE, sum minFDE, agents]; returns (ADE, FDE, agents)."""
    if sums.dtype != torch.float64 or sums.numel() != 3:
        raise ValueError("sums must be a float64 tensor of 3 elements")
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)
    n = float(sums[2])
    return float(sums[0]) / max(n, 1.0), float(sums[1]) / max(n, 1.0), n

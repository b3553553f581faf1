"""Multi-GPU sharding: problems are independent, so ranks never exchange data on the solve path.

Two layouts:
  * weak scaling (bench.py): rank r generates its own config-2 batch from seeds r*1024 .. r*1024+n-1
  * an existing problem list (file batches): longest-processing-time-first round robin over ranks,
    so every GPU receives the same mix of long and short problems
plus the only collectives a run needs: a MAX-reduce of the device time and a SUM of the rows solved
(backend nccl on GPUs, gloo in the CPU tests).
"""


def rank_seeds(rank, n_vectors):
    """Seeds of the count vectors rank `rank` owns (disjoint across ranks)."""
    return list(range(rank * 1024, rank * 1024 + n_vectors))


def lpt_assign(sizes, world):
    """Assign problems (by size) to `world` ranks: sort by size descending, deal round robin in a
    snake order.  Returns a list of index lists, one per rank; every index appears exactly once."""
    order = sorted(range(len(sizes)), key=lambda i: (-sizes[i], i))
    shards = [[] for _ in range(world)]
    for k, i in enumerate(order):
        rnd, pos = divmod(k, world)
        r = pos if rnd % 2 == 0 else world - 1 - pos
        shards[r].append(i)
    return shards


def reduce_time_and_rows(ms, rows, dist=None, device=None):
    """(max over ranks of ms, sum over ranks of rows).  dist: torch.distributed or None (1 rank)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms), float(rows)
    import torch
    t = torch.tensor([float(ms)], dtype=torch.float64, device=device)
    r = torch.tensor([float(rows)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(r, op=dist.ReduceOp.SUM)
    return float(t.item()), float(r.item())

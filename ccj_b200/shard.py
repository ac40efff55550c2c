"""Independent-sequence partitioning across the GPUs of one box.

Sequences do not interact, so there is no data-path collective: rank r folds a contiguous, cost-balanced
slice and the (small) results are gathered on the host with `all_gather_object`.  Cost model: a fold is
O(n^5) time, so slices are balanced on sum(n^5) rather than on count.
"""
from __future__ import annotations

from typing import Callable, List, Sequence, Tuple


def partition(lengths: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous [begin,end) slices per rank with roughly equal sum(n^5)."""
    n = len(lengths)
    prefix = [0.0]
    for x in lengths:
        prefix.append(prefix[-1] + float(max(x, 1)) ** 5)
    total = prefix[-1]
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        c = cuts[-1]
        while c < n and abs(prefix[c + 1] - target) <= abs(prefix[c] - target):
            c += 1
        cuts.append(c)
    cuts.append(n)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def fold_sharded(seqs: Sequence[str], fold_batch: Callable[[List[str]], list], rank: int, world: int,
                 gather: Callable[[list], List[list]] = None) -> list:
    """Every rank folds its slice with `fold_batch` (Context.fold_batch on a GPU); `gather` collects the
    per-rank lists (default: torch.distributed.all_gather_object). Returns results in input order."""
    lo, hi = partition([len(s) for s in seqs], world)[rank]
    mine = fold_batch(list(seqs[lo:hi])) if hi > lo else []
    if world == 1:
        return mine
    if gather is None:
        import torch.distributed as dist

        def gather(obj):
            out = [None] * world
            dist.all_gather_object(out, obj)
            return out
    parts = gather(mine)
    return [x for part in parts for x in part]

"""Independent-sequence partitioning across the GPUs of one box.

Sequences do not interact, so there is no data-path collective.  Two ways to spread them:
  * `fold_sharded`: rank r folds a contiguous, cost-balanced slice (static; cost model: a fold is O(n^5) time, so
    slices are balanced on sum(n^5) rather than on count);
  * `fold_dealt`: ranks draw chunks of consecutive indices from one shared counter until the job is empty
    (dynamic dealing: a GPU that is slower -- power capping -- simply draws fewer chunks).  The counter is the
    process group's TCP store (atomic add) under torchrun, a lock-protected integer inside one process.
The (small) results are gathered on the host with `all_gather_object`.
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


class LocalCounter:
    """Shared counter inside one process (threads, or a single rank)."""

    def __init__(self):
        import threading
        self._v, self._lock = 0, threading.Lock()

    def take(self, count: int) -> int:
        with self._lock:
            v = self._v
            self._v += count
            return v


class StoreCounter:
    """Shared counter across ranks: `add` on the torch.distributed key-value store is an atomic fetch-and-add."""

    def __init__(self, store, key: str = "ccj_next"):
        self._store, self._key = store, key

    def take(self, count: int) -> int:
        return int(self._store.add(self._key, count)) - count


def fold_dealt(total: int, chunk: int, counter, fold_range: Callable[[int, int], list]) -> List[Tuple[int, object]]:
    """Draws [lo, lo+chunk) index ranges from `counter` until `total` is exhausted and folds each with
    `fold_range(lo, hi)` (which returns one result per index).  Returns this rank's [(index, result)]."""
    out: List[Tuple[int, object]] = []
    chunk = max(1, int(chunk))
    while True:
        lo = counter.take(chunk)
        if lo >= total:
            return out
        hi = min(lo + chunk, total)
        res = fold_range(lo, hi)
        out.extend(zip(range(lo, hi), res))

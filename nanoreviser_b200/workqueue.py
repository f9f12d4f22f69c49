"""Host work queue: reads shard naturally (every read is independent), so multi-GPU needs no
collective.  Replaces the reference's ``multiprocessing.Pool`` fan-out of one fast5 per task
(NanoReviser.py:203-223) with

  * a length-balanced partition of the reads over ranks (greedy LPT by number of bases), and
  * ragged batches under a per-launch base budget (so that one rank's launches have equal cost).

Deterministic: every rank computes the same partition from the same list of lengths, so no
communication is needed to agree on it.
"""
from __future__ import annotations

import heapq
from typing import List, Sequence

import numpy as np


def lpt_partition(lengths: Sequence[int], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first: returns ``n_parts`` lists of read indices with near-equal
    total bases.  Ties are broken by index, so the result is a pure function of ``lengths``."""
    if n_parts <= 0:
        raise ValueError("n_parts must be positive")
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    heap = [(0, p) for p in range(n_parts)]
    heapq.heapify(heap)
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for i in order:
        load, p = heapq.heappop(heap)
        parts[p].append(i)
        heapq.heappush(heap, (load + int(lengths[i]), p))
    for p in parts:
        p.sort()
    return parts


def shard_for_rank(lengths: Sequence[int], rank: int, world: int) -> List[int]:
    return lpt_partition(lengths, world)[rank]


def make_batches(indices: Sequence[int], lengths: Sequence[int], base_budget: int) -> List[List[int]]:
    """Group reads (kept in the given order) into ragged batches of at most ``base_budget`` bases.
    A read longer than the budget forms a batch of its own (the library chunks windows internally)."""
    batches: List[List[int]] = []
    cur: List[int] = []
    load = 0
    for i in indices:
        n = int(lengths[i])
        if cur and load + n > base_budget:
            batches.append(cur)
            cur, load = [], 0
        cur.append(i)
        load += n
    if cur:
        batches.append(cur)
    return batches


def imbalance(lengths: Sequence[int], parts: Sequence[Sequence[int]]) -> float:
    """max load / mean load (1.0 = perfect)."""
    loads = np.array([sum(int(lengths[i]) for i in p) for p in parts], dtype=np.float64)
    return float(loads.max() / max(loads.mean(), 1e-9)) if len(loads) else 1.0

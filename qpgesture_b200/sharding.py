"""Row sharding of the candidate database across the GPUs of one box.

Sequences are split into `world` contiguous blocks in their original order, so a
window's global id 26*j + m is the same on every rank and the cross-rank merge is a
plain lexicographic (distance, id) minimum -- ties keep the smallest global id, i.e.
the reference's scan order (GestureKNN.py:671-690).  The data-path exchange is ONE
all-gather of the per-rank [Q, 512] tables followed by qpg_table_merge on each rank.
"""
from __future__ import annotations

import numpy as np

PAIR_DTYPE = np.dtype([("d", "<f8"), ("id", "<i8")])


def shard_sequences(n_seq: int, world: int, rank: int):
    """[j0, j1) of rank `rank`; blocks differ by at most one sequence."""
    return rank * n_seq // world, (rank + 1) * n_seq // world


def merge_tables_host(parts_i64: np.ndarray) -> np.ndarray:
    """Host mirror of qpg_table_merge for tests: parts int64 [P, Q, 512, 2] -> structured [Q, 512].
    Empty bins carry id -1, which must lose against every real id (the device kernel compares ids
    as unsigned)."""
    parts = np.ascontiguousarray(parts_i64).view(PAIR_DTYPE).reshape(parts_i64.shape[:-1])
    d = parts["d"]
    uid = parts["id"].astype(np.uint64)
    best = np.zeros(parts.shape[1:], dtype=PAIR_DTYPE)
    bd, bi = d[0].copy(), uid[0].copy()
    for p in range(1, parts.shape[0]):
        take = (d[p] < bd) | ((d[p] == bd) & (uid[p] < bi))
        bd = np.where(take, d[p], bd)
        bi = np.where(take, uid[p], bi)
    best["d"], best["id"] = bd, bi.astype(np.int64)
    return best

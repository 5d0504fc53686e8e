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


MIN_SHARD_BYTES = 1 << 30


def plan_layout(db_bytes: int, world: int, min_shard_bytes: int = MIN_SHARD_BYTES):
    """2-D decomposition of `world` GPUs: (row_shards, clip_groups), row_shards * clip_groups == world.

    The window table is cut into `row_shards` row blocks only as long as a block stays >= 1 GiB: one
    query pass over a shard then lasts >= ~150 us, so the fixed per-pass costs (launch ramp, query load,
    table flush: 8-10 us, measured) stay below ~5 %; cutting a table that is already small (speaker-10:
    0.35 GB) only multiplies those costs (measured: 0.72 weak-scaling efficiency with 2 shards).  The
    remaining factor replicates the table and splits the query clips, which are independent.  Ranks r
    with the same r // row_shards form one row group (they exchange tables with one all-gather); rank r
    holds row block r % row_shards."""
    rs = 1
    while rs * 2 <= world and world % (rs * 2) == 0 and db_bytes // (rs * 2) >= min_shard_bytes:
        rs *= 2
    return rs, world // rs


BIN_DTYPE = np.dtype([("lo", "<f8"), ("hi", "<f8"), ("id", "<i8"), ("n", "<i4"), ("flags", "<i4")])


def merge_bin_records_host(parts_i64: np.ndarray):
    """Host mirror of the merge step of qpg_sliced_resolve for tests.  parts int64 [P, ..., 4] (qpg_bin_t records of
    P row shards) -> (candidates bool [P, ...]: shards whose interval reaches below the smallest interval end,
    decided_id int64 [...]: the window id when exactly one shard remains (else -2), empty bins -1)."""
    parts = np.ascontiguousarray(parts_i64).view(BIN_DTYPE).reshape(parts_i64.shape[:-1])
    valid = parts["id"] >= 0
    hi = np.where(valid, parts["hi"], np.inf)
    U = hi.min(axis=0)
    cand = valid & (parts["lo"] <= U[None])
    n = cand.sum(axis=0)
    first = cand.argmax(axis=0)
    ids = np.take_along_axis(parts["id"], first[None], axis=0)[0]
    return cand, np.where(n == 1, ids, np.where(n == 0, -1, -2))

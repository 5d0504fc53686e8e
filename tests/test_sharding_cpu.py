"""Host-side logic of the multi-GPU path, on CPU: row-shard arithmetic and the
all-gather + lexicographic merge semantics, with world_size-2 gloo processes.
(The device merge kernel itself is covered by tests/test_matcher_gpu.py.)"""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import matcher_np as om
from qpgesture_b200.sharding import BIN_DTYPE, merge_bin_records_host, merge_tables_host, shard_sequences


def test_shard_sequences_cover_and_order():
    for n, w in [(512, 8), (513, 8), (7, 8), (26, 3), (1, 1)]:
        ranges = [shard_sequences(n, w, r) for r in range(w)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n
        for (a0, a1), (b0, b1) in zip(ranges[:-1], ranges[1:]):
            assert a1 == b0 and a0 <= a1
        sizes = [b - a for a, b in ranges]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(0)                       # same data on every rank
    n_seq, D, Q = 12, 32, 5
    rows = rng.standard_normal((n_seq * 26, D))
    rows[26 * 9 + 3] = rows[5]                           # exact tie across shards: global id 5 must win
    labels = rng.integers(0, 40, size=n_seq * 26)
    labels[26 * 9 + 3] = labels[5]
    q = rng.standard_normal((Q, D))
    j0, j1 = shard_sequences(n_seq, world, rank)
    w0, w1 = 26 * j0, 26 * j1
    part = np.zeros((Q, 512), dtype=[("d", "<f8"), ("id", "<i8")])
    for qi in range(Q):
        d = om.cosine_rows(q[qi], rows[w0:w1])
        bd, bw = om.min_by_code(d, labels[w0:w1])
        part[qi]["d"] = bd
        part[qi]["id"] = np.where(bw >= 0, bw + w0, -1)
    t = torch.from_numpy(part.view(np.int64).reshape(Q, 512, 2).copy())
    gathered = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(gathered, t)
    merged = merge_tables_host(torch.stack(gathered).numpy())
    if rank == 0:
        full = np.zeros((Q, 512), dtype=[("d", "<f8"), ("id", "<i8")])
        for qi in range(Q):
            d = om.cosine_rows(q[qi], rows)
            full[qi]["d"], full[qi]["id"] = om.min_by_code(d, labels)
        np.save(out, np.array([np.array_equal(merged["id"], full["id"]), np.array_equal(merged["d"], full["d"])]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_rank_merge(tmp_path):
    out = str(tmp_path / "ok.npy")
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    ok = np.load(out)
    assert ok.all()


def _records_worker(rank, world, port, out):
    """The sliced engine's exchange on CPU: per-bin interval records laid out [query][table][code], clips split
    evenly -> ONE all_to_all_single gives every rank the records of its own clips from every shard."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(1)                        # same data on every rank
    n_clips, steps, D, W = 4, 3, 16, 26 * 10
    Q = n_clips * steps
    rows = rng.standard_normal((W, D))
    labels = rng.integers(0, 30, size=W)
    q = rng.standard_normal((2, Q, D))                    # "audio" and "text" queries against the same rows
    w0, w1 = rank * W // world, (rank + 1) * W // world
    rec = np.zeros((Q, 2, 512), dtype=BIN_DTYPE)
    rec["lo"], rec["hi"], rec["id"] = 1e3, 1e3, -1
    for x in range(2):
        for qi in range(Q):
            d = om.cosine_rows(q[x, qi], rows[w0:w1])
            bd, bw = om.min_by_code(d, labels[w0:w1])
            ok = bw >= 0
            rec["lo"][qi, x][ok] = bd[ok] - 1e-7            # an interval around the shard's best distance
            rec["hi"][qi, x][ok] = bd[ok] + 1e-7
            rec["id"][qi, x][ok] = bw[ok] + w0
    send = torch.from_numpy(rec.view(np.int64).reshape(Q, 2, 512, 4).copy())
    per = Q // world
    recv = torch.empty((world, per, 2, 512, 4), dtype=torch.int64)
    dist.all_to_all_single(recv, send)                      # chunk r of `send` = queries of rank r's clips
    cand, decided = merge_bin_records_host(recv.numpy())
    mine = slice(rank * per, (rank + 1) * per)
    ok = True
    for x in range(2):
        for i, qi in enumerate(range(mine.start, mine.stop)):
            d = om.cosine_rows(q[x, qi], rows)
            bd, bw = om.min_by_code(d, labels)
            dec = decided[i, x]
            settled = dec != -2                             # -2: several shards within 2e-7, float64 decides on the GPU
            ok &= bool(np.array_equal(dec[settled], bw[settled]))
            ok &= bool((dec == -2).sum() <= 2)
    res = torch.tensor([int(ok)])
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(out, np.array([int(res.item())]))
    dist.barrier()
    dist.destroy_process_group()


def test_gloo_two_rank_record_exchange(tmp_path):
    out = str(tmp_path / "ok2.npy")
    port = 31500 + (os.getpid() % 2000)
    mp.spawn(_records_worker, args=(2, port, out), nprocs=2, join=True)
    assert np.load(out).all()

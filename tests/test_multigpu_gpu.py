"""Row-sharded matcher over NCCL on 2 GPUs: rank r scans rows of sequences [j0, j1) (int8-sliced copy), the
float32 copy is replicated, ONE collective exchanges the per-bin records, every rank resolves its own clips.
The result must be identical to the single-GPU run over the whole table: window ids, rank transform, codes.
Also the float64 engine (all-gather of both tables + qpg_table_merge).  Skipped with fewer than 2 GPUs."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir):
    import torch
    import torch.distributed as dist
    from qpgesture_b200.GestureKNN import CodeKNN
    from qpgesture_b200.matchdb import MatchDatabase
    from qpgesture_b200.sharding import shard_sequences

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    rng = np.random.default_rng(5)                        # same data on every rank
    n_seq, Da, Dt, n_clips, n_seg = 96, 512, 128, 4, 2
    code = rng.integers(0, 512, size=(n_seq, 30)).astype(np.int64)
    sig = rng.standard_normal((512, 135)).astype(np.float32)
    phase_amp = rng.standard_normal((n_seq, 240, 16)).astype(np.float32)
    aud = rng.standard_normal((n_seq * 26, Da)).astype(np.float32)
    txt = rng.standard_normal((n_seq * 26, Dt)).astype(np.float32)
    aud[26 * 70 + 3] = aud[9]                             # duplicate window across the two shards: id 9 must win
    txt[26 * 70 + 3] = txt[9]
    code[70, 3] = code[0, 9]
    aq = rng.standard_normal((n_clips, n_seg, 8, Da)).astype(np.float32)
    tq = rng.standard_normal((n_clips, n_seg, 8, Dt)).astype(np.float32)
    aq[1, 0, 2] = aud[9]
    tq[1, 0, 2] = txt[9]
    seed_code = rng.integers(0, 512, size=n_clips).astype(np.int32)
    seed_phase = rng.standard_normal((n_clips, 8, 16)).astype(np.float32)

    def run(knn, engine, tail_clips):
        p = knn.make_plan(n_clips, n_seg, tail_clips=tail_clips, use_graph=False, engine=engine)
        p.qa.copy_(torch.from_numpy(aq.reshape(-1, Da)))
        p.qt.copy_(torch.from_numpy(tq.reshape(-1, Dt)))
        p.seed_code.copy_(torch.from_numpy(seed_code))
        p.seed_phase.copy_(torch.from_numpy(seed_phase))
        knn.run_plan(p)
        torch.cuda.synchronize()
        return p, dict(codes=p.codes.cpu().numpy(), ids_a=p.ta[..., 1].cpu().numpy(), ids_t=p.tt[..., 1].cpu().numpy(),
                       ra=p.ra.cpu().numpy(), rt=p.rt.cpu().numpy(), status=p.status.cpu().numpy())

    # single-GPU reference on this rank: the whole table
    full = CodeKNN(database=MatchDatabase("A", code, sig, phase_amp, txt, aud_rows=aud, device=dev), use_wavlm=True,
                   use_phase=True, use_txt=True, tail="device")
    per = n_clips // world
    mine = slice(rank * per, (rank + 1) * per)
    _, want = run(full, "sliced", mine)
    j0, j1 = shard_sequences(n_seq, world, rank)
    res = {}
    # sliced engine, all_to_all (even clip split) and all_gather (rank 0 takes 3 clips, rank 1 one)
    sh = CodeKNN(database=MatchDatabase("A", code, sig, phase_amp, txt, aud_rows=aud, device=dev, seq_range=(j0, j1),
                                        replicate_exact=True), use_wavlm=True, use_phase=True, use_txt=True,
                 process_group=dist.group.WORLD, tail="device")
    p, got = run(sh, "sliced", mine)
    res["a2a"] = p.exchange == "all_to_all" and all(np.array_equal(got[k], want[k]) for k in want)
    uneven = slice(0, 3) if rank == 0 else slice(3, 4)
    _, want_u = run(full, "sliced", uneven)
    p, got = run(sh, "sliced", uneven)
    res["gather"] = p.exchange == "all_gather" and all(np.array_equal(got[k], want_u[k]) for k in want_u)
    # float64 engine with row-sharded float32 tables: one all-gather of both tables + merge kernel
    sh64 = CodeKNN(database=MatchDatabase("A", code, sig, phase_amp, txt, aud_rows=aud, device=dev, seq_range=(j0, j1),
                                          sliced=False), use_wavlm=True, use_phase=True, use_txt=True,
                   process_group=dist.group.WORLD, tail="device")
    _, got = run(sh64, "f64", mine)
    res["f64"] = all(np.array_equal(got[k], want[k]) for k in ("codes", "ids_a", "ids_t", "ra", "rt"))
    dup_bin = int(code[0, 9])
    flags = torch.tensor([int(res["a2a"]), int(res["gather"]), int(res["f64"])], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        np.save(os.path.join(out_dir, "ok.npy"), flags.cpu().numpy())
        # the duplicated window: clip 1, segment 0, step 2 asked for exactly that row -> smaller global id
        _, all_clips = run(full, "sliced", slice(0, n_clips))
        np.save(os.path.join(out_dir, "dup.npy"), np.array([all_clips["ids_a"][(1 * n_seg + 0) * 8 + 2, dup_bin]]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_row_shards_equal_single_gpu(tmp_path):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    port = 32500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    ok = np.load(str(tmp_path / "ok.npy"))
    assert ok.tolist() == [1, 1, 1], f"[all_to_all, all_gather, f64] = {ok.tolist()}"
    assert int(np.load(str(tmp_path / "dup.npy"))[0]) == 9

#!/usr/bin/env python
"""Stage-by-stage timing of the one-pass matcher step (sliced engine) on one GPU.

    python tools/bench_sliced.py [--n-seq 512] [--clips 1] [--reps 50]

Prints one JSON object: per-stage CUDA-event times (each stage timed alone, back to back `reps` times, table
larger than L2 so every scan pass streams from HBM), the captured-graph step time, and the float64 engine
for comparison."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-seq", type=int, default=512)
    ap.add_argument("--clips", type=int, default=1)
    ap.add_argument("--n-seg", type=int, default=6)
    ap.add_argument("--wavlm-dim", type=int, default=1024)
    ap.add_argument("--ctx-dim", type=int, default=384)
    ap.add_argument("--reps", type=int, default=50)
    ap.add_argument("--no-f64", action="store_true")
    a = ap.parse_args()
    import torch
    from qpgesture_b200 import _lib
    from qpgesture_b200.GestureKNN import CodeKNN
    from qpgesture_b200.matchdb import MatchDatabase

    dev = torch.device("cuda")
    lib = _lib.load()
    rng = np.random.default_rng(0)
    n = a.n_seq
    code = rng.integers(0, 512, size=(n, 30)).astype(np.int64)
    sig = rng.standard_normal((512, 135)).astype(np.float32)
    phase_amp = rng.standard_normal((n, 240, 16)).astype(np.float32)
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    aud = torch.randn((n * 26, 6 * a.wavlm_dim), device=dev, generator=g)
    txt = torch.randn((n * 26, a.ctx_dim), device=dev, generator=g)
    db = MatchDatabase("A", code, sig, phase_amp, txt, aud_rows=aud, device=dev)
    del aud, txt
    knn = CodeKNN(database=db, use_wavlm=True, use_phase=True, use_txt=True, tail="device")
    out = dict(n_seq=n, windows=n * 26, clips=a.clips, steps=a.clips * a.n_seg * 8)

    def timed(fn, reps=a.reps):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps * 1e3      # us

    for engine in (("sliced",) if a.no_f64 else ("sliced", "f64")):
        p = knn.make_plan(a.clips, a.n_seg, use_graph=True, engine=engine)
        p.qa.copy_(torch.randn(p.qa.shape, device=dev, generator=g))
        p.qt.copy_(torch.randn(p.qt.shape, device=dev, generator=g))
        p.seed_code.fill_(7)
        p.seed_phase.copy_(torch.randn(p.seed_phase.shape, device=dev, generator=g))
        out[engine + "_graph_us"] = timed(lambda: knn.run_plan(p))
        out[engine + "_status"] = p.status.cpu().tolist()
        if engine != "sliced":
            continue
        out["passes"] = len(p.passes)
        out["stats_per_step"] = [int(x) // (a.reps + 3 + 2) for x in p.stats.cpu().tolist()]
        sp = _lib.stream_ptr()
        SP = [sp]                                            # the stream the stage functions launch on
        A, T = db.aud_s, db.txt_s
        ps = p.passes[0]
        jobs = (_lib.SliceJob * 2)()
        for x, (S, q, qi, qs) in enumerate(((A, p.qa, p.qinfo_a, ps.qs_a), (T, p.qt, p.qinfo_t, ps.qs_t))):
            jobs[x].q, jobs[x].col_exp = _lib.dptr(q[ps.q0:ps.q0 + ps.nq]), _lib.dptr(S.col_exp)
            jobs[x].q_slices, jobs[x].q_info = _lib.dptr(qs), _lib.dptr(qi[ps.q0:ps.q0 + ps.nq])
            jobs[x].ldq, jobs[x].D = S.D, S.D
        segs = (_lib.SlicedSeg * 2)()
        segs[0].db_slices, segs[0].q_slices, segs[0].sacc, segs[0].n_kblocks = A.slices.data_ptr(), ps.qs_a.data_ptr(), p.sacc_a.data_ptr(), A.n_kblocks
        segs[1].db_slices, segs[1].q_slices, segs[1].sacc, segs[1].n_kblocks = T.slices.data_ptr(), ps.qs_t.data_ptr(), p.sacc_t.data_ptr(), T.n_kblocks
        tabs_b = knn._sliced_tables(p, ps.q0, ps.nq, ps.q0)
        tabs_r = knn._sliced_tables(p, 0, p.Qt, 0, for_resolve=True)

        def slice_q():
            lib.qpg_slice_queries_i8(jobs, 2, ps.nq, ps.n_pad, SP[0])

        def scan():            # accumulates on top of earlier repetitions: timing only
            lib.qpg_sliced_scan_i8(segs, 2, A.W, ps.n_pad, ps.nq, SP[0])

        def scan_audio_only():
            lib.qpg_sliced_scan_i8(segs, 1, A.W, ps.n_pad, ps.nq, SP[0])

        def restore():         # a valid sacc for the stages behind the scan
            p.sacc_a.zero_()
            p.sacc_t.zero_()
            scan()

        def bins():            # consume = 0: the same valid sacc for every repetition
            lib.qpg_sliced_bins(tabs_b, 2, A.W, ps.nq, db.id_offset, db.row_base, 0, None, SP[0])

        def resolve():
            lib.qpg_sliced_resolve(tabs_r, 2, 1, 2 * p.Q * 512, p.Qt, db.exact_offset, None, SP[0])

        def lookup():
            lib.qpg_match_lookup(_lib.ptr(p.ta), _lib.ptr(p.tt), _lib.ptr(p.ra), _lib.ptr(p.rt), _lib.ptr(db.pos_rank_t), _lib.ptr(db.freq_rank),
                                 _lib.ptr(db.code), db.n_seq, _lib.ptr(db.aud_frame), _lib.ptr(db.txt_frame), _lib.ptr(p.qfa), _lib.ptr(p.qft),
                                 p.Qt, _lib.ptr(p.entries), SP[0])

        def walk_table():
            lib.qpg_match_walk_stats(_lib.ptr(p.entries), _lib.ptr(db.code), _lib.ptr(db.phase_amp), _lib.ptr(p.phase_stats),
                                     _lib.ptr(p.seed_code), _lib.ptr(p.seed_phase), p.n_tail, p.n_seg, _lib.ptr(p.trans),
                                     _lib.ptr(p.codes), _lib.ptr(p.vote), None, _lib.ptr(p.status), SP[0])

        def walk_direct():
            lib.qpg_match_walk(_lib.ptr(p.entries), _lib.ptr(db.code), _lib.ptr(db.phase_amp), _lib.ptr(p.seed_code), _lib.ptr(p.seed_phase),
                               p.n_tail, p.n_seg, None, _lib.ptr(p.codes), _lib.ptr(p.vote), None, _lib.ptr(p.status), SP[0])
        def timed_graph(fn, reps=20):
            """device time per launch with the host out of the loop: `reps` launches captured into one CUDA graph
            (the Python/ctypes launch path costs 15-30 us per call, more than most of these kernels take)"""
            st = torch.cuda.Stream()
            st.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(st):
                SP[0] = _lib.stream_ptr(st)
                fn()
                st.synchronize()
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    SP[0] = _lib.stream_ptr()
                    for _ in range(reps):
                        fn()
                gr.replay()
                st.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(5):
                    gr.replay()
                e1.record(st)
                st.synchronize()
            SP[0] = sp
            torch.cuda.current_stream().wait_stream(st)
            return e0.elapsed_time(e1) / (5 * reps) * 1e3

        out["stage_us"] = {}
        out["stage_us_graph"] = {}
        for name, f in (("slice_queries", slice_q), ("scan", scan), ("scan_audio_only", scan_audio_only)):
            out["stage_us"][name] = timed(f)
            out["stage_us_graph"][name] = timed_graph(f)
        restore()
        stages = [("bins", bins), ("resolve", resolve), ("lookup", lookup), ("walk_direct", walk_direct)]
        if p.trans is not None:
            stages.append(("walk_table", walk_table))
        for name, f in stages:
            out["stage_us"][name] = timed(f)
            out["stage_us_graph"][name] = timed_graph(f)
        p.sacc_a.zero_()
        p.sacc_t.zero_()
        sliced_bytes = A.nbytes + T.nbytes
        alg_bytes = db.W * (4 * (db.aud.D + db.txt.D) + 4)
        out["scan_GBps_sliced_bytes"] = sliced_bytes / out["stage_us"]["scan"] / 1e3
        out["scan_GBps_algorithmic"] = alg_bytes / out["stage_us"]["scan"] / 1e3
    print(json.dumps(out))


if __name__ == "__main__":
    main()

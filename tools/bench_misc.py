#!/usr/bin/env python
"""Timings of the smaller kernels of the hot path (one JSON line each): Levenshtein min-by-code scan (K10),
VQ L2-argmin (K1), rank512, table merge."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpgesture_b200 import _lib
from qpgesture_b200.matchdb import new_table

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

def timed_graph(fn, reps=20):
    """device time per call with the host out of the loop: `reps` calls captured into one CUDA graph"""
    st = torch.cuda.Stream()
    with torch.cuda.stream(st):
        fn(); fn()
        st.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=st):
            for _ in range(reps): fn()
        g.replay(); st.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        for _ in range(5): g.replay()
        e1.record(st); st.synchronize()
    return e0.elapsed_time(e1) / (5 * reps)

def main():
    lib = _lib.load(); dev = torch.device("cuda"); sp = _lib.stream_ptr()
    g = torch.Generator(device=dev); g.manual_seed(0)
    for W, Q in ((13312, 48), (1_000_000, 48)):
        tok = torch.randint(0, 102400, (W, 12), device=dev, dtype=torch.int32, generator=g)
        qt = torch.randint(0, 102400, (Q, 12), device=dev, dtype=torch.int32, generator=g)
        lab = torch.randint(0, 512, (W,), device=dev, dtype=torch.int32, generator=g)
        tab = new_table(Q, dev)
        lib.qpg_table_init(_lib.ptr(tab), Q * 512, sp)
        ms = timed(lambda: lib.qpg_cand_lev_minbycode(_lib.ptr(tok), _lib.ptr(lab), W, 0, _lib.ptr(qt), Q, _lib.ptr(tab), sp))
        print(json.dumps(dict(kernel="cand_lev_kernel", W=W, Q=Q, ms=ms, pairs_per_s=W * Q / ms * 1e3,
                              dp_cells_per_s=W * Q * 121 / ms * 1e3, bytes_per_s=W * 52 * (-(-Q // 4)) / ms * 1e3)), flush=True)
    cb = torch.randn((512, 512), device=dev, generator=g)
    for M in (30, 960, 65536):
        x = torch.randn((M, 512), device=dev, generator=g)
        idx = torch.empty(M, dtype=torch.int64, device=dev); mind = torch.empty(M, device=dev)
        ms = timed(lambda: lib.qpg_vq_argmin_f32(_lib.ptr(x), _lib.ptr(cb), M, 512, 512, _lib.ptr(idx), _lib.ptr(mind), sp),
                   reps=20 if M < 10000 else 3)
        print(json.dumps(dict(kernel="vq_argmin_kernel", M=M, ms=ms, latents_per_s=M / ms * 1e3,
                              tflops=2 * M * 512 * 512 / ms / 1e9)), flush=True)
        scratch = torch.empty(M * 512 + 512, device=dev)
        idx2 = torch.empty(M, dtype=torch.int64, device=dev)
        ms2 = timed(lambda: lib.qpg_vq_argmin_fast(_lib.ptr(x), _lib.ptr(cb), M, 512, 512, _lib.ptr(scratch), _lib.ptr(idx2),
                                                   _lib.ptr(mind), sp), reps=20)
        print(json.dumps(dict(kernel="vq_argmin_fast (tf32 tcgen05 filter + exact re-evaluation)", M=M, ms=ms2,
                              latents_per_s=M / ms2 * 1e3, equal_to_exact=bool(torch.equal(idx, idx2)))), flush=True)
        if M <= 960:
            msg = timed_graph(lambda: lib.qpg_vq_argmin_fast(_lib.ptr(x), _lib.ptr(cb), M, 512, 512, _lib.ptr(scratch),
                                                             _lib.ptr(idx2), _lib.ptr(mind), _lib.stream_ptr()))
            print(json.dumps(dict(kernel="vq_argmin_fast, replayed from a CUDA graph (device time)", M=M, ms=msg)), flush=True)
    Q = 48
    tab = new_table(Q, dev); lib.qpg_table_init(_lib.ptr(tab), Q * 512, sp)
    ranks = torch.empty((Q, 512), dtype=torch.int32, device=dev)
    print(json.dumps(dict(kernel="rank512_kernel", Q=Q, ms=timed(lambda: lib.qpg_rank512(_lib.ptr(tab), Q, _lib.ptr(ranks), sp)))))
    parts = torch.zeros((8, Q, 512, 2), dtype=torch.int64, device=dev); out = new_table(Q, dev)
    print(json.dumps(dict(kernel="table_merge_kernel", parts=8, Q=Q,
                          ms=timed(lambda: lib.qpg_table_merge(_lib.ptr(parts), 8, Q * 512, _lib.ptr(out), sp)))))

def legacy():
    """legacy pose-feature matcher (GestureKNN class): seconds per batch of 64-frame clips against a synthetic database"""
    import time
    from qpgesture_b200.GestureKNN import GestureKNN
    rng = np.random.default_rng(0)
    n_seq, n_frames, B = 2048, 64, 64
    feat = rng.standard_normal((n_seq, n_frames, 208)); motn = rng.standard_normal((n_seq, n_frames, 165))
    mask = np.ones((n_seq, n_frames), dtype=np.int64)
    tests = rng.standard_normal((B, 112, n_frames))
    inits = [(int(rng.integers(n_seq)), int(rng.integers(n_frames))) for _ in range(B)]
    for ties in ("stable", "numpy"):
        knn = GestureKNN(feat, motn, mask, device="cuda:0", ties=ties)
        knn.search_motion_batch(tests, 0, inits)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        knn.search_motion_batch(tests, 0, inits)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps(dict(kernel="legacy GestureKNN.search_motion_batch", ties=ties, n_seq=n_seq, n_frames=n_frames,
                              clips=B, seconds_per_batch=dt, clips_per_s=B / dt,
                              frame_distances_per_s=B * 8 * n_seq * n_frames / dt)), flush=True)

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "legacy":
        legacy()
    else:
        main()

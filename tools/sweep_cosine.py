#!/usr/bin/env python
"""Time qpg_cand_cosine_minbycode over (queries/pass, compute warps, ring depth) on a synthetic table."""
import argparse, itertools, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qpgesture_b200 import _lib
from qpgesture_b200.matchdb import PackedRows, new_table

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--W", type=int, default=13312)
    ap.add_argument("--D", type=int, default=6144)
    ap.add_argument("--reps", type=int, default=30)
    ap.add_argument("--configs", type=str, default="")
    ap.add_argument("--Q", type=int, default=0, help="time a multi-pass call with Q queries (0: one pass of qt queries)")
    ap.add_argument("--alt", type=int, default=1)
    a = ap.parse_args()
    lib = _lib.load(); dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(0)
    rows = torch.randn((a.W, a.D), device=dev, generator=g)
    pr = PackedRows.from_rows(rows); del rows
    labels = torch.randint(0, 512, (a.W,), device=dev, dtype=torch.int32)
    q = torch.randn((max(8, a.Q), a.D), device=dev, generator=g)
    lib.qpg_tune_cosine_alternate(a.alt)
    sp = _lib.stream_ptr()
    peak = 6550.1
    try: peak = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"]
    except Exception: pass
    combos = [tuple(map(int, c.split(","))) for c in a.configs.split(";") if c] or \
        [(qt, ncw, ns, 0) for qt in (1, 2, 4, 8) for ncw in (8, 12) for ns in (2, 3)]
    alg = a.W * (4 * a.D + 4)
    for combo in combos:
        qt, ncw, ns = combo[:3]
        team = combo[3] if len(combo) > 3 else 0
        nq = a.Q if a.Q else qt
        tab = new_table(nq, dev)
        lib.qpg_tune_cosine(ncw, ns, 0, team)
        lib.qpg_table_init(_lib.ptr(tab), nq * 512, sp)
        def run():
            return lib.qpg_cand_cosine_minbycode(_lib.ptr(pr.packed), _lib.ptr(pr.sqnorm), _lib.ptr(labels), a.W, a.D, 0,
                                                 _lib.ptr(q), nq, _lib.ptr(tab), qt, sp)
        rc = run()
        if rc != 0:
            print(f"qt={qt} ncw={ncw} ns={ns}: rc={rc} {lib.qpg_last_error().decode()}"); continue
        for _ in range(3): run()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.reps): run()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.reps / max(1, -(-nq // qt))
        gbs = alg / ms / 1e6
        print(f"W={a.W} D={a.D} qt={qt} ncw={ncw} ns={ns} team={team}: {ms*1e3:8.1f} us  {gbs:7.0f} GB/s  frac={gbs/peak:.3f}  per-query {ms*1e3/qt:7.1f} us")
    lib.qpg_tune_cosine(0, 0, 0, 0)
    lib.qpg_tune_cosine_alternate(1)

if __name__ == "__main__":
    main()

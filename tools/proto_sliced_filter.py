#!/usr/bin/env python
"""CPU prototype of the round-2 one-pass scan: exact int8-sliced integer dot products + interval logic.

Each float32 row x (and query q) is written as a 31-bit fixed-point integer relative to a per-row
power-of-two scale, X = rint(x * 2^(30-ex)), and X is cut into four balanced base-256 digits
(int8 "slices"):  X = d0*2^24 + d1*2^16 + d2*2^8 + d3.  The tensor cores (tcgen05 kind::i8, int32
accumulators) compute the EXACT integer sums P_st = sum_k d_s[k] e_t[k]; the kernel keeps the ten
products with s+t <= 3.  Everything the filter drops is bounded rigorously:

    dot(x,q) = 2^(ex+eq-60) * [ sum_{s+t<=3} 2^(48-8(s+t)) P_st          (computed, exact)
                               + sum_{s+t>=4} ...                           (dropped: |.| <= 2^23*1.01*128... )
                               + sum dX*Y + X*dY + dX*dY ]                  (quantisation, |dX|,|dY| <= 1/2)

This script measures, on the speaker-10-like synthetic shapes, the bound, the observed error and how
many (query, start-code) bins the bound cannot decide (those get an exact float64 re-evaluation).
"""
import argparse
import numpy as np


def slice_rows(x):
    """x float32 [n, D] -> (digits int8 [4, n, D], ex int [n], X int64 [n, D])"""
    x64 = x.astype(np.float64)
    mx = np.abs(x64).max(axis=1)
    _, e = np.frexp(mx)                      # mx = m * 2^e, m in [0.5, 1)  ->  |x| < 2^e
    e = np.where(mx > 0, e, 0)
    X = np.rint(np.ldexp(x64, (30 - e)[:, None])).astype(np.int64)
    assert np.abs(X).max() <= 2 ** 30
    digs = []
    r = X.copy()
    for _ in range(3):
        d = ((r + 128) & 255) - 128
        digs.append(d)
        r = (r - d) >> 8
    assert r.min() >= -128 and r.max() <= 127
    digs.append(r)
    digs = digs[::-1]                        # d0 (most significant) .. d3
    chk = ((digs[0] * 256 + digs[1]) * 256 + digs[2]) * 256 + digs[3]
    assert np.array_equal(chk, X)
    return np.stack(digs).astype(np.int8), e.astype(np.int64), X


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-seq", type=int, default=512)
    ap.add_argument("--D", type=int, default=6144)
    ap.add_argument("--Q", type=int, default=8)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    W = a.n_seq * 26
    rows = rng.standard_normal((W, a.D)).astype(np.float32)
    labels = rng.integers(0, 512, size=W)
    q = rng.standard_normal((a.Q, a.D)).astype(np.float32)
    dx, ex, X = slice_rows(rows)
    dq, eq, Y = slice_rows(q)
    sqx = (rows.astype(np.float64) ** 2).sum(1)
    sqq = (q.astype(np.float64) ** 2).sum(1)
    l1x = np.abs(X).sum(1).astype(np.float64)
    l1y = np.abs(Y).sum(1).astype(np.float64)
    eq_low = np.abs(dq[1:].astype(np.int64)).sum(axis=(0, 2)).astype(np.float64)     # sum |e1|+|e2|+|e3|
    K = a.D
    dxf = [dx[s].astype(np.float64) for s in range(4)]
    tot = amb_id = amb_rank = 0
    max_err = max_bound = 0.0
    for i in range(a.Q):
        # exact integer sums of the ten kept products (float64 matmul of small ints is exact here: < 2^53)
        v = np.zeros(W, dtype=np.float64)
        for s in range(4):
            for t in range(4 - s):
                P = dxf[s] @ dq[t, i].astype(np.float64)
                v += P * 2.0 ** (24 - 8 * (s + t))
        dot_f = np.ldexp(v, ex + eq[i] - 36)
        dot64 = rows.astype(np.float64) @ q[i].astype(np.float64)
        # dropped terms (s+t >= 4): |d_s| <= 128 on the database side -> 128 * (2^16 + 2^8 + 1) * sum(|e1|+|e2|+|e3|)
        bound_int = 0.5 * l1y[i] + 0.5 * l1x + K / 4 + 128.0 * 65793.0 * eq_low[i]
        bound_dot = np.ldexp(bound_int, ex + eq[i] - 60)
        nrm = np.sqrt(sqx) * np.sqrt(sqq[i])
        eps = bound_dot / nrm + 1e-12
        d_f = 1.0 - dot_f / nrm
        d64 = 1.0 - dot64 / nrm
        err = np.abs(d_f - d64)
        assert (err <= eps).all(), "bound violated"
        max_err, max_bound = max(max_err, err.max()), max(max_bound, eps.max())
        lo, hi = d_f - eps, d_f + eps
        Ub = np.full(512, np.inf)
        np.minimum.at(Ub, labels, hi)
        cand = lo <= Ub[labels]
        cnt = np.bincount(labels[cand], minlength=512)
        Lb = np.full(512, np.inf)
        np.minimum.at(Lb, labels[cand], lo[cand])
        nonempty = np.isfinite(Ub)
        tot += int(nonempty.sum())
        amb_id += int((cnt > 1).sum())
        order = np.argsort(Lb[nonempty])
        L_s, U_s = Lb[nonempty][order], Ub[nonempty][order]
        ov = L_s[1:] <= np.maximum.accumulate(U_s)[:-1]
        amb_rank += int(ov.sum())
        # filter + verify == exact
        for c in np.flatnonzero(cnt > 1):
            idx = np.flatnonzero(cand & (labels == c))
            all_idx = np.flatnonzero(labels == c)
            assert all_idx[np.argmin(d64[all_idx])] == idx[np.argmin(d64[idx])]
    print(f"W={W} D={a.D} Q={a.Q}: max |filter - f64| {max_err:.2e}, rigorous bound (max) {max_bound:.2e}")
    print(f"bins {tot}: id-ambiguous {amb_id} ({100 * amb_id / tot:.4f} %), rank-overlapping neighbours {amb_rank} "
          f"({100 * amb_rank / tot:.3f} %)")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU prototype for the round-2 scan design: float32 filter pass + exact float64 verification.

Question: if the streaming pass accumulated in float32 (FFMA, no F2F/DFMA, so 8 queries per pass become
affordable), how many (query, start-code) bins would need an exact float64 re-evaluation to GUARANTEE the
same window ids as the float64 reference?  A bin is ambiguous when its runner-up is within twice the
worst-case float32 error of its best candidate.

Error model of the planned kernel: each lane sums D/32 products in `acc_per_lane` independent FFMA chains,
then a 5-level shuffle tree (+ log2(acc_per_lane) adds): depth n = D/(32*acc_per_lane) + 5 + log2(acc);
|fl(dot) - dot| <= gamma_n * sum|x_i q_i| <= gamma_n * |x| |q|, so the cosine error is <= gamma_n ~ n * 2^-24.
"""
import argparse
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-seq", type=int, default=512)
    ap.add_argument("--D", type=int, default=6144)
    ap.add_argument("--Q", type=int, default=48)
    ap.add_argument("--acc-per-lane", type=int, default=4)
    a = ap.parse_args()
    rng = np.random.default_rng(0)
    W = a.n_seq * 26
    rows = rng.standard_normal((W, a.D)).astype(np.float32)
    labels = rng.integers(0, 512, size=W)
    q = rng.standard_normal((a.Q, a.D)).astype(np.float32)
    u = 2.0 ** -24
    depth = a.D // (32 * a.acc_per_lane) + 5 + int(np.log2(a.acc_per_lane))
    bound = depth * u / (1 - depth * u)          # on the cosine (dot / (|x||q|))
    rn = np.sqrt((rows.astype(np.float64) ** 2).sum(1))
    tot_bins = amb_bins = amb_rows = wrong_without_verify = 0
    max_seen_err = 0.0
    for i in range(a.Q):
        qn = np.sqrt((q[i].astype(np.float64) ** 2).sum())
        d64 = 1.0 - (rows.astype(np.float64) @ q[i].astype(np.float64)) / (rn * qn)
        # float32 accumulation in the planned order: [lane][chain] partial sums, then tree
        prod = (rows * q[i][None, :]).astype(np.float32)        # FFMA rounds once per step; this is an upper proxy
        part = prod.reshape(W, -1, 32 * a.acc_per_lane).sum(axis=1, dtype=np.float32)   # chains
        d32 = 1.0 - part.sum(axis=1, dtype=np.float32).astype(np.float64) / (rn * qn)
        max_seen_err = max(max_seen_err, float(np.abs(d32 - d64).max()))
        for c in range(512):
            idx = np.flatnonzero(labels == c)
            if idx.size == 0:
                continue
            tot_bins += 1
            o = idx[np.argsort(d32[idx], kind="stable")]
            best = o[0]
            cand = idx[d32[idx] <= d32[best] + 2 * bound]       # everything the bound cannot separate
            if cand.size > 1:
                amb_bins += 1
                amb_rows += cand.size
            true_best = idx[np.argmin(d64[idx])]
            verified = cand[np.argmin(d64[cand])]
            assert verified == true_best, "filter + verify must be exact"
            wrong_without_verify += int(best != true_best)
    db_bytes = W * 4 * a.D
    print(f"W={W} D={a.D} Q={a.Q}: depth {depth}, rigorous cosine error bound {bound:.2e}, max observed {max_seen_err:.2e}")
    print(f"bins {tot_bins}, ambiguous {amb_bins} ({100 * amb_bins / tot_bins:.3f} %), rows to verify {amb_rows} "
          f"= {amb_rows * 4 * a.D / 1e6:.1f} MB vs {a.Q // 8 + (a.Q % 8 > 0)} filter passes x {db_bytes / 1e6:.0f} MB")
    print(f"bins the float32 pass alone would have got wrong: {wrong_without_verify}")


if __name__ == "__main__":
    main()

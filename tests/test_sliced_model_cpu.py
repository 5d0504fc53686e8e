"""Mathematics of the int8-sliced filter (csrc/sliced_scan.cu), checked on the host with the NumPy model:
digits reconstruct the fixed-point integer, the ten kept products + the stated bound contain the float64
distance for every (query, row), and interval logic + exact re-evaluation reproduces the float64 tables."""
import numpy as np
import pytest

from oracle import matcher_np as om
from tests import _sliced_model as sm


@pytest.mark.parametrize("W,D,Q,outliers", [(400, 384, 3, False), (300, 1000, 2, True), (64, 130, 4, False)])
def test_bound_contains_float64_distance(W, D, Q, outliers):
    rng = np.random.default_rng(W + D)
    rows = rng.standard_normal((W, D)).astype(np.float32)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    col_exp = None
    if outliers:
        rows[:, 7] *= 300.0                      # a massive-activation feature dimension
        rows[:, 99] *= 1e-3
        rows[5] = 0.0
        rows[9, :] = 1e-20                       # tiny but non-zero row
        q[1] = rows[3]
        col_exp = np.floor(np.log2(np.abs(rows).max(0))).astype(np.int64)
        col_exp -= int(np.median(col_exp))
    sx, sq = sm.slice_rows(rows, col_exp, -1), sm.slice_rows(q, col_exp, +1)
    v = sm.exact_v(sx["digits"], sq["digits"])
    lo, hi = sm.intervals(v, sx, sq, rows, q)
    d = sm.f64_distances(rows, q)
    assert (lo <= d).all() and (d <= hi).all()
    if not outliers:
        assert np.abs(0.5 * (lo + hi) - d).max() < 1e-6 and (hi - lo).max() < 1e-6
    # sklearn's own formula (the oracle) is inside the interval too
    for i in range(Q):
        d_ref = om.cosine_rows(q[i].astype(np.float64), rows.astype(np.float64))
        assert (lo[i] - 1e-12 <= d_ref).all() and (d_ref <= hi[i] + 1e-12).all()


def test_interval_logic_reproduces_exact_tables():
    rng = np.random.default_rng(3)
    W, D, Q = 3000, 256, 4
    rows = rng.standard_normal((W, D)).astype(np.float32)
    rows[1200] = rows[7]                          # duplicate rows: smaller id must win
    q = rng.standard_normal((Q, D)).astype(np.float32)
    labels = rng.integers(0, 40, size=W)
    labels[1200] = labels[7]
    sx, sq = sm.slice_rows(rows), sm.slice_rows(q)
    lo, hi = sm.intervals(sm.exact_v(sx["digits"], sq["digits"]), sx, sq, rows, q)
    d = sm.f64_distances(rows, q)
    for i in range(Q):
        bd, bw = om.min_by_code(d[i], labels)
        for c in range(40):
            idx = np.flatnonzero(labels == c)
            U = hi[i, idx].min()
            cand = idx[lo[i, idx] <= U]
            best = cand[np.lexsort((cand, d[i, cand]))[0]]
            assert best == bw[c]
        dup_bin = labels[7]
        assert bw[dup_bin] != 1200


def test_swizzle_offsets_are_a_permutation():
    off = sm.swz_offset(np.arange(128)[:, None], np.arange(128)[None, :]).ravel()
    assert np.array_equal(np.sort(off), np.arange(128 * 128))
    # 16-byte chunks stay contiguous and 8-row groups are 1024 bytes apart (UMMA SWIZZLE_128B, SBO = 1024)
    assert sm.swz_offset(8, 0) == 1024 and sm.swz_offset(1, 0) == 128 + 16 and sm.swz_offset(0, 17) == 17

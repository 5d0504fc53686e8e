"""NumPy model of the int8-sliced fixed-point representation and of the filter's error bound
(qpgesture_b200/csrc/sliced_scan.cu).  Test infrastructure: the GPU tests compare the kernels with it, the
CPU tests check the mathematics (digits reconstruct X, the bound really bounds the float64 distance)."""
import numpy as np

DROP_C = 128.0 * 65793.0
EPS_SLACK = 1e-12
TINY_SQ = 4.930380657631324e-30      # sklearn normalize(): rows with a norm below 10*eps stay unscaled


def slice_rows(x, col_exp=None, sign=-1):
    """x float32 [n, D] -> dict(digits int8 [4, n, D], ex int64 [n], X int64 [n, D]).
    Database rows use sign=-1 (x * 2^-col_exp), queries sign=+1."""
    x64 = np.asarray(x, dtype=np.float64)
    if col_exp is not None:
        x64 = np.ldexp(x64, sign * np.asarray(col_exp, dtype=np.int64)[None, :])
    mx = np.abs(x64).max(axis=1) if x64.shape[1] else np.zeros(x64.shape[0])
    _, e = np.frexp(mx)
    e = np.where(mx > 0, e, 0).astype(np.int64)
    X = np.rint(np.ldexp(x64, (30 - e)[:, None])).astype(np.int64)
    assert np.abs(X).max(initial=0) <= 2 ** 30
    digs, r = [], X.copy()
    for _ in range(3):
        d = ((r + 128) & 255) - 128
        digs.append(d)
        r = (r - d) >> 8
    assert r.min(initial=0) >= -128 and r.max(initial=0) <= 127
    digs.append(r)
    digs = digs[::-1]
    assert np.array_equal(((digs[0] * 256 + digs[1]) * 256 + digs[2]) * 256 + digs[3], X)
    return dict(digits=np.stack(digs).astype(np.int8), ex=e, X=X)


def swz_offset(r, kbyte):
    r, kbyte = np.asarray(r, dtype=np.int64), np.asarray(kbyte, dtype=np.int64)
    return (r >> 3) * 1024 + (r & 7) * 128 + ((((kbyte >> 4) ^ (r & 7)) & 7) << 4) + (kbyte & 15)


def unswizzle_db(buf, W, D):
    """uint8 tile images [RT][NKB][4][128 x 128 B] -> int8 digits [4, W, D]"""
    rt_n, nkb = -(-W // 128), -(-D // 128)
    t = np.asarray(buf, dtype=np.uint8).reshape(rt_n, nkb, 4, 128 * 128)
    r = np.arange(128)[:, None]
    kb = np.arange(128)[None, :]
    off = swz_offset(r, kb)                                   # [128 rows, 128 bytes]
    tiles = t[..., off]                                       # [rt, nkb, 4, 128, 128]
    dig = tiles.transpose(2, 0, 3, 1, 4).reshape(4, rt_n * 128, nkb * 128)
    return dig[:, :W, :D].view(np.int8)


def unswizzle_q(buf, Q, D, n_pad):
    nkb = -(-D // 128)
    t = np.asarray(buf, dtype=np.uint8).reshape(nkb, 4, n_pad * 128)
    off = swz_offset(np.arange(n_pad)[:, None], np.arange(128)[None, :])
    tiles = t[..., off]                                       # [nkb, 4, n_pad, 128]
    dig = tiles.transpose(1, 2, 0, 3).reshape(4, n_pad, nkb * 128)
    return dig[:, :Q, :D].view(np.int8)


def exact_v(dx, dq):
    """digits int8 [4, W, D], [4, Q, D] -> int64 [Q, W]: v = sum_{s+t<=3} 2^(24-8(s+t)) <d_s, e_t> (exact)."""
    W, Q = dx.shape[1], dq.shape[1]
    v = np.zeros((Q, W), dtype=np.int64)
    for s in range(4):
        for t in range(4 - s):
            P = dq[t].astype(np.int64) @ dx[s].astype(np.int64).T
            v += P << (24 - 8 * (s + t))
    return v


def intervals(v, sx, sq, rows, q):
    """distance intervals [lo, hi] per (query, row) exactly as filter_interval() evaluates them."""
    rows64, q64 = np.asarray(rows, dtype=np.float64), np.asarray(q, dtype=np.float64)
    sqx, sqq = (rows64 ** 2).sum(1), (q64 ** 2).sum(1)
    D = rows64.shape[1]
    l1x = np.abs(sx["X"]).sum(1).astype(np.float64)
    l1y = np.abs(sq["X"]).sum(1).astype(np.float64)
    el = np.abs(sq["digits"][1:].astype(np.int64)).sum(axis=(0, 2)).astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        r1 = np.where(sqx > TINY_SQ, np.ldexp(1.0, sx["ex"] - 60) / np.sqrt(sqx), 0.0)
        g = np.where(sqq > TINY_SQ, np.ldexp(1.0, sq["ex"]) / np.sqrt(sqq), 0.0)
    r2 = 0.5 * l1x * r1
    h = 0.5 * l1y + DROP_C * el + 0.25 * D
    c = v.astype(np.float64) * (r1 * 16777216.0)[None, :] * g[:, None]
    eps = (g[:, None] * (r1[None, :] * h[:, None] + r2[None, :])) * (1 + 1e-9) + EPS_SLACK
    d = 0.5 * ((sqq > TINY_SQ)[:, None].astype(np.float64) + (sqx > TINY_SQ)[None, :].astype(np.float64)) - c
    return np.maximum(d - eps, 0.0), np.maximum(d + eps, 0.0)


def f64_distances(rows, q):
    rows64, q64 = np.asarray(rows, dtype=np.float64), np.asarray(q, dtype=np.float64)
    sqx, sqq = (rows64 ** 2).sum(1), (q64 ** 2).sum(1)
    dot = q64 @ rows64.T
    with np.errstate(divide="ignore", invalid="ignore"):
        c = np.where((sqq[:, None] > TINY_SQ) & (sqx[None, :] > TINY_SQ),
                     dot / (np.sqrt(sqq)[:, None] * np.sqrt(sqx)[None, :]), 0.0)
    d = 0.5 * ((sqq > TINY_SQ)[:, None].astype(np.float64) + (sqx > TINY_SQ)[None, :].astype(np.float64)) - c
    return np.maximum(d, 0.0)


def stable_rank(d):
    return np.argsort(np.argsort(d, kind="stable"), kind="stable")

"""One-pass int8-sliced scan (csrc/sliced_scan.cu) and the lookup/walk tail (csrc/match_walk.cu) on the GPU:
tile images and integer sums against the NumPy model, the tensor-core kernel against the CUDA-core
evaluation, the resulting tables against the float64 scan kernels and the oracle, the plan against the
golden vectors of the reference, and the benchmarked configuration against the oracle."""
import ctypes as C

import numpy as np
import pytest

from oracle import matcher_np as om
from tests import _sliced_model as sm
from tests._common import golden_cases, load_case, oracle_db, oracle_queries

pytestmark = pytest.mark.gpu
CASES = golden_cases()


def _env():
    import torch
    from qpgesture_b200 import _lib
    return torch, _lib, _lib.load(), torch.device("cuda")


def _slice_db(rows_d, labels_d, column_scaling=True):
    """-> (PackedRows, SlicedRows) of a device row table"""
    from qpgesture_b200.matchdb import PackedRows, SlicedRows, bin_order
    pr = PackedRows.from_rows(rows_d)
    order, bin_start = bin_order(labels_d)
    return pr, SlicedRows.from_rows(rows_d, pr.sqnorm, order, bin_start, column_scaling=column_scaling)


def _slice_queries(q_d, col_exp, n_pad):
    torch, _lib, lib, dev = _env()
    from qpgesture_b200.matchdb import aligned_bytes
    Q, D = q_d.shape
    qs = aligned_bytes(lib.qpg_sliced_query_bytes(D, n_pad), dev)
    qinfo = torch.zeros((Q, 4), dtype=torch.float64, device=dev)
    job = (_lib.SliceJob * 1)()
    job[0].q, job[0].col_exp, job[0].q_slices, job[0].q_info = _lib.dptr(q_d), _lib.dptr(col_exp), _lib.dptr(qs), _lib.dptr(qinfo)
    job[0].ldq, job[0].D = D, D
    _lib.check(lib.qpg_slice_queries_i8(job, 1, Q, n_pad, _lib.stream_ptr()), "slice_q")
    torch.cuda.synchronize()
    return qs, qinfo


def _table_desc(pr, S, q_d, qinfo, sacc=None, bins=None, tab=None, ranks=None, qf=None):
    torch, _lib, lib, dev = _env()
    t = (_lib.SlicedTable * 1)()
    t[0].packed, t[0].row_sqnorm, t[0].q, t[0].q_info = _lib.dptr(pr.packed), _lib.dptr(pr.sqnorm), _lib.dptr(q_d), _lib.dptr(qinfo)
    t[0].ldq, t[0].D = q_d.shape[1], q_d.shape[1]
    if S is not None:
        t[0].sacc, t[0].bin_start, t[0].row_info, t[0].order = _lib.dptr(sacc), _lib.dptr(S.bin_start), _lib.dptr(S.row_info), _lib.dptr(S.order)
    t[0].bins, t[0].table, t[0].ranks, t[0].qflags = _lib.dptr(bins), _lib.dptr(tab), _lib.dptr(ranks), _lib.dptr(qf)
    return t


def _scan_tc(segs_spec, W, n_pad, nq):
    """segs_spec: list of (SlicedRows, q_slices) -> list of sacc tensors [n_pad, Wpad] from the tcgen05 kernel"""
    torch, _lib, lib, dev = _env()
    segs = (_lib.SlicedSeg * 2)()
    saccs = []
    for i, (S, qs) in enumerate(segs_spec):
        sacc = torch.zeros((n_pad, S.Wpad), dtype=torch.int64, device=dev)
        saccs.append(sacc)
        segs[i].db_slices, segs[i].q_slices, segs[i].sacc, segs[i].n_kblocks = \
            S.slices.data_ptr(), qs.data_ptr(), sacc.data_ptr(), S.n_kblocks
    _lib.check(lib.qpg_sliced_scan_i8(segs, len(segs_spec), W, n_pad, nq, _lib.stream_ptr()), "scan_i8")
    torch.cuda.synchronize()
    return saccs


def _scan_ref(S, qs, n_pad, nq, q_stride=1):
    torch, _lib, lib, dev = _env()
    sacc = torch.zeros((n_pad, S.Wpad), dtype=torch.int64, device=dev)
    _lib.check(lib.qpg_sliced_scan_ref(_lib.ptr(S.slices), _lib.ptr(qs), S.n_kblocks, S.W, n_pad, nq, q_stride,
                                       _lib.ptr(sacc), _lib.stream_ptr()), "scan_ref")
    torch.cuda.synchronize()
    return sacc


@pytest.mark.parametrize("W,D,Q,n_pad,outliers", [(300, 200, 5, 16, False), (257, 384, 16, 16, True), (130, 1030, 3, 32, False)])
def test_slicing_matches_numpy_model(W, D, Q, n_pad, outliers):
    """tile images (swizzle included), exponents and the per-row / per-query bound terms"""
    torch, _lib, lib, dev = _env()
    rng = np.random.default_rng(W + D)
    rows = rng.standard_normal((W, D)).astype(np.float32)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    if outliers:
        rows[:, 3] *= 200.0
        rows[:, 50] *= 1e-2
        rows[7] = 0.0
        q[2] = 0.0
    labels = rng.integers(0, 512, size=W).astype(np.int32)
    labels[5] = 700                                           # out-of-range label: sorted last, never read
    rows_d, lab_d = torch.from_numpy(rows).to(dev), torch.from_numpy(labels).to(dev)
    pr, S = _slice_db(rows_d, lab_d)
    order = S.order.cpu().numpy()
    key = np.where((labels >= 0) & (labels < 512), labels, 512)
    assert np.array_equal(order, np.argsort(key, kind="stable"))
    assert np.array_equal(S.bin_start.cpu().numpy(), np.searchsorted(np.sort(key), np.arange(513)))
    col_exp = None if S.col_exp is None else S.col_exp.cpu().numpy().astype(np.int64)
    assert col_exp is not None or not outliers
    sx = sm.slice_rows(rows[order], col_exp, -1)
    got = sm.unswizzle_db(S.slices.cpu().numpy(), W, D)
    assert np.array_equal(got, sx["digits"])
    # padding rows / columns are zero
    assert int((S.slices != 0).sum()) == int((sx["digits"] != 0).sum())
    sqx = (rows[order].astype(np.float64) ** 2).sum(1)
    ri = S.row_info.cpu().numpy()
    with np.errstate(divide="ignore", invalid="ignore"):
        r1 = np.where(sqx > sm.TINY_SQ, np.ldexp(1.0, sx["ex"] - 60) / np.sqrt(sqx), 0.0)
    assert np.allclose(ri[:, 0], r1, rtol=1e-13, atol=0)
    assert np.allclose(ri[:, 1], 0.5 * np.abs(sx["X"]).sum(1) * r1, rtol=1e-13, atol=0)
    # queries
    qs, qinfo = _slice_queries(torch.from_numpy(q).to(dev), S.col_exp, n_pad)
    sq = sm.slice_rows(q, col_exp, +1)
    assert np.array_equal(sm.unswizzle_q(qs.cpu().numpy(), Q, D, n_pad), sq["digits"])
    qi = qinfo.cpu().numpy()
    sqq = (q.astype(np.float64) ** 2).sum(1)
    assert np.allclose(qi[:, 0], sqq, rtol=1e-14, atol=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        g = np.where(sqq > sm.TINY_SQ, np.ldexp(1.0, sq["ex"]) / np.sqrt(sqq), 0.0)
    assert np.allclose(qi[:, 1], g, rtol=1e-13, atol=0)
    el = np.abs(sq["digits"][1:].astype(np.int64)).sum(axis=(0, 2))
    assert np.allclose(qi[:, 2], 0.5 * np.abs(sq["X"]).sum(1) + sm.DROP_C * el + 0.25 * D, rtol=1e-13, atol=0)
    assert np.array_equal(qinfo.view(torch.int32)[:, 6].cpu().numpy(), sq["ex"])


@pytest.mark.parametrize("W,D,Q,n_pad", [(300, 200, 5, 16), (129, 128, 1, 16), (1000, 384, 20, 32)])
def test_reference_kernel_matches_exact_integers(W, D, Q, n_pad):
    torch, _lib, lib, dev = _env()
    rng = np.random.default_rng(W * 3 + D)
    rows = rng.standard_normal((W, D)).astype(np.float32)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    labels = rng.integers(0, 512, size=W).astype(np.int32)
    pr, S = _slice_db(torch.from_numpy(rows).to(dev), torch.from_numpy(labels).to(dev))
    qs, _ = _slice_queries(torch.from_numpy(q).to(dev), S.col_exp, n_pad)
    order = S.order.cpu().numpy()
    col = None if S.col_exp is None else S.col_exp.cpu().numpy().astype(np.int64)
    want = sm.exact_v(sm.slice_rows(rows[order], col, -1)["digits"], sm.slice_rows(q, col, +1)["digits"])
    got = _scan_ref(S, qs, n_pad, Q).cpu().numpy()[:Q, :W]
    assert np.array_equal(got, want)


@pytest.mark.parametrize("W,D1,D2,Q,n_pad", [
    (128, 128, 0, 1, 16),            # one tile, one k-block
    (130, 384, 0, 16, 16),           # second tile mostly padding
    (1000, 384, 128, 8, 16),         # two feature blocks
    (5000, 1024, 384, 33, 48),       # stream-K splits inside tiles, n_pad 48
    (3000, 6144, 384, 48, 48),       # the real row shapes
    (13312, 6144, 384, 48, 48),      # the benchmarked configuration
    (20000, 512, 256, 64, 64),       # many tiles per CTA, n_pad 64, complete runs (plain stores)
    (777, 200, 72, 5, 32),           # ragged D (zero padded columns)
])
def test_tensor_core_scan_equals_cuda_core_reference(W, D1, D2, Q, n_pad):
    """tcgen05 kind::i8 stream-K kernel == plain CUDA-core evaluation of the same tile images, bit for bit"""
    torch, _lib, lib, dev = _env()
    g = torch.Generator(device=dev)
    g.manual_seed(W + D1)
    labels = torch.randint(0, 512, (W,), device=dev, dtype=torch.int32, generator=g)
    spec = []
    for D in (D1, D2):
        if D == 0:
            continue
        rows = torch.randn((W, D), device=dev, generator=g)
        q = torch.randn((Q, D), device=dev, generator=g)
        pr, S = _slice_db(rows, labels)
        qs, _ = _slice_queries(q, S.col_exp, n_pad)
        spec.append((S, qs))
    got = _scan_tc(spec, W, n_pad, Q)
    for (S, qs), sacc in zip(spec, got):
        # the CUDA-core evaluation is slow: every query on small tables, every 16th one on the big ones
        stride = 1 if W * S.D * Q < 3e9 else 16
        want = _scan_ref(S, qs, n_pad, Q, stride)
        a, b = sacc[:Q:stride, :W], want[:Q:stride, :W]
        diff = a != b
        assert not bool(diff.any()), (f"{int(diff.sum())} of {a.numel()} sums differ; first at "
                                      f"{torch.nonzero(diff)[:4].tolist()} got {a[diff][:4].tolist()} "
                                      f"want {b[diff][:4].tolist()}")
        assert not bool(sacc[Q:].any()), "rows of unused queries must stay zero"
    again = _scan_tc(spec, W, n_pad, Q)
    for a, b in zip(got, again):
        assert torch.equal(a, b)                               # integer accumulation: order independent


def _tables_sliced(rows, labels, q, id_offset=0, column_scaling=True):
    """full pipeline through the C ABI -> (table structured [Q,512], ranks [Q,512], qflags, stats)"""
    torch, _lib, lib, dev = _env()
    from qpgesture_b200.matchdb import new_table, table_to_numpy
    rows_d = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.float32)).to(dev)
    lab_d = torch.from_numpy(np.ascontiguousarray(labels, dtype=np.int32)).to(dev)
    q_d = torch.from_numpy(np.ascontiguousarray(q, dtype=np.float32)).to(dev)
    W, D = rows_d.shape
    Q = q_d.shape[0]
    n_pad = -(-Q // 16) * 16
    pr, S = _slice_db(rows_d, lab_d, column_scaling)
    qs, qinfo = _slice_queries(q_d, S.col_exp, n_pad)
    (sacc,) = _scan_tc([(S, qs)], W, n_pad, Q)
    bins = torch.zeros((Q, 512, 4), dtype=torch.int64, device=dev)
    stats = torch.zeros((2,), dtype=torch.int64, device=dev)
    sp = _lib.stream_ptr()
    tab = new_table(Q, dev)
    ranks = torch.zeros((Q, 512), dtype=torch.int32, device=dev)
    qf = torch.zeros((Q,), dtype=torch.int32, device=dev)
    desc = _table_desc(pr, S, q_d, qinfo, sacc, bins, tab, ranks, qf)
    _lib.check(lib.qpg_sliced_bins(desc, 1, W, Q, id_offset, 0, 1, _lib.ptr(stats), sp), "bins")
    _lib.check(lib.qpg_sliced_resolve(desc, 1, 1, Q * 512, Q, id_offset, _lib.ptr(stats), sp), "resolve")
    torch.cuda.synchronize()
    assert not bool(sacc.any()), "consume=1 must leave sacc zeroed for the next pass"
    return table_to_numpy(tab), ranks.cpu().numpy(), qf.cpu().numpy(), stats.cpu().numpy(), bins


@pytest.mark.parametrize("W,D,Q,nbins,special", [(1000, 384, 3, 300, True), (5000, 512, 16, 512, False),
                                                 (2048, 6144, 5, 64, True), (40, 130, 2, 512, False),
                                                 (26000, 384, 48, 512, True), (70000, 64, 9, 512, True)])
def test_sliced_tables_vs_float64(W, D, Q, nbins, special):
    """ids, empty bins and rank transform identical to the float64 evaluation; distances within 2e-7 (exact where a
    decision was needed)"""
    rng = np.random.default_rng(W + 31 * D)
    rows = rng.standard_normal((W, D)).astype(np.float32)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    labels = rng.integers(0, nbins, size=W)
    if special:
        rows[W // 2] = rows[3]                                # exact duplicate: the smaller id must win
        labels[W // 2] = labels[3]
        rows[17] = 0.0                                        # all-zero window
        q[0] = rows[5]                                        # query equal to a window
        # two nearly parallel rows in one bin, the query equal to one of them: 4e-12 apart, i.e. far below the
        # filter's resolution and far above float64 rounding -> must go through the float64 path
        rows[21] = rows[20] + np.float32(3e-6) * rng.standard_normal(D).astype(np.float32)
        labels[21] = labels[20]
        q[1] = rows[20]
        if Q > 8:
            q[8] = 0.0                                        # all-zero query: every row of a bin ties (degenerate path)
        labels[30] = 999                                      # ignored row
    table, ranks, qf, stats, _ = _tables_sliced(rows, labels, q, id_offset=26 * 7)
    d = sm.f64_distances(rows, q)
    d[:, (labels < 0) | (labels >= 512)] = np.inf
    for i in range(Q):
        bd, bw = om.min_by_code(np.where(np.isfinite(d[i]), d[i], 1e9), np.where(labels < 512, labels, 0))
        if special:
            bd, bw = om.min_by_code(np.delete(d[i], 30), np.delete(labels, 30))
            bw = np.where(bw >= 30, bw + 1, bw)               # ids of the rows after the ignored one
        want_id = np.where(bw >= 0, bw + 26 * 7, -1)
        assert np.array_equal(table[i]["id"], want_id), f"query {i}: ids differ in bins {np.flatnonzero(table[i]['id'] != want_id)[:8]}"
        assert np.all(table[i]["d"][bw < 0] == 1e3)
        assert np.allclose(table[i]["d"], bd, rtol=0, atol=2e-7)
        assert np.array_equal(ranks[i], sm.stable_rank(bd)), f"query {i}: rank transform differs"
    if special:
        assert stats[0] >= 2                                  # the near-parallel pair went through the float64 path


def test_sliced_tables_vs_oracle_and_f64_kernel():
    """against sklearn's own formula (the oracle) and against the float64 scan kernel of round 1"""
    from tests.test_matcher_gpu import _scan_cosine
    rng = np.random.default_rng(12)
    W, D, Q = 3000, 768, 6
    rows = rng.standard_normal((W, D)).astype(np.float32)
    q = rng.standard_normal((Q, D)).astype(np.float32)
    labels = rng.integers(0, 400, size=W)
    table, ranks, _, _, _ = _tables_sliced(rows, labels, q)
    exact, _ = _scan_cosine(rows, labels, q)
    assert np.array_equal(table["id"], exact["id"])
    assert np.allclose(table["d"], exact["d"], rtol=0, atol=2e-7)
    for i in range(Q):
        bd, bw = om.min_by_code(om.cosine_rows(q[i].astype(np.float64), rows.astype(np.float64)), labels)
        assert np.array_equal(table[i]["id"], bw)
        assert np.array_equal(ranks[i], sm.stable_rank(exact[i]["d"]))


def test_outlier_columns_use_column_scaling():
    """a massive-activation feature dimension: column scaling keeps the intervals tight and the result exact"""
    rng = np.random.default_rng(2)
    W, D, Q = 4000, 512, 8
    rows = rng.standard_normal((W, D)).astype(np.float32)
    rows[:, 11] *= 500.0
    rows[:, 12] *= 1e-3
    q = rng.standard_normal((Q, D)).astype(np.float32)
    q[:, 11] *= 500.0
    labels = rng.integers(0, 512, size=W)
    d = sm.f64_distances(rows, q)
    for scaling in (True, False):
        table, ranks, _, stats, _ = _tables_sliced(rows, labels, q, column_scaling=scaling)
        for i in range(Q):
            bd, bw = om.min_by_code(d[i], labels)
            assert np.array_equal(table[i]["id"], bw) and np.array_equal(ranks[i], sm.stable_rank(bd))


def test_resolve_merges_row_shards():
    """records of 3 row shards (scanned separately, global ids) + qpg_sliced_resolve == single-table result;
    a duplicate row across shards keeps the smaller global id"""
    torch, _lib, lib, dev = _env()
    from qpgesture_b200.matchdb import PackedRows, new_table, table_to_numpy
    rng = np.random.default_rng(8)
    W, D, Q = 26 * 90, 256, 7
    rows = rng.standard_normal((W, D)).astype(np.float32)
    labels = rng.integers(0, 150, size=W)
    rows[26 * 70 + 4] = rows[9]
    labels[26 * 70 + 4] = labels[9]
    q = rng.standard_normal((Q, D)).astype(np.float32)
    full, full_rank, _, _, _ = _tables_sliced(rows, labels, q)
    cuts = [0, 26 * 20, 26 * 55, W]
    parts = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        *_, bins = _tables_sliced(rows[a:b], labels[a:b], q, id_offset=a)
        parts.append(bins)
    parts_d = torch.stack(parts).contiguous()                                  # [P, Q, 512, 4]
    pr = PackedRows.from_rows(torch.from_numpy(rows).to(dev))                 # replicated float32 table, all rows
    q_d = torch.from_numpy(q).to(dev)
    _, qinfo = _slice_queries(q_d, None, 16)
    tab = new_table(Q, dev)
    ranks = torch.zeros((Q, 512), dtype=torch.int32, device=dev)
    qf = torch.zeros((Q,), dtype=torch.int32, device=dev)
    desc = _table_desc(pr, None, q_d, qinfo, None, parts_d, tab, ranks, qf)
    _lib.check(lib.qpg_sliced_resolve(desc, 1, 3, Q * 512, Q, 0, None, _lib.stream_ptr()), "resolve")
    torch.cuda.synchronize()
    merged = table_to_numpy(tab)
    assert np.array_equal(merged["id"], full["id"])
    assert np.allclose(merged["d"], full["d"], rtol=0, atol=2e-7)
    assert np.array_equal(ranks.cpu().numpy(), full_rank)


def test_lookup_walk_equals_round1_tail():
    """rank512_ties + match_lookup + match_walk == match_tail_kernel (same stable order) on random tables, incl. a
    clip that runs into an empty bin (status 1, codes -1 from there on)"""
    torch, _lib, lib, dev = _env()
    from qpgesture_b200.matchdb import PAIR_DTYPE
    rng = np.random.default_rng(4)
    n_seq, n_clips, n_seg = 40, 5, 3
    Q = n_clips * n_seg * 8
    code = rng.integers(0, 512, size=(n_seq, 30)).astype(np.int32)
    phase = rng.standard_normal((n_seq, 240, 16)).astype(np.float32)
    phase[6] = phase[5]                                       # equal candidates: exact distance ties (audio wins)
    phase[7] = 0.0                                            # all-zero phase vectors
    phase[8] *= np.float32(1e-21)                             # squares underflow in float32: float64 pick only
    phase[9] = phase[10] * np.float32(1.0 + 3e-4)             # distances closer than the filter's 1e-3 margin
    pos_rank = np.stack([rng.permutation(512) for _ in range(512)]).astype(np.int32)
    freq_rank = rng.permutation(512).astype(np.int32)
    frames_a = np.array([int(6 * m / 398 * 240) for m in range(26)], dtype=np.int32)
    frames_t = np.array([int(8 * m / 398 * 240) for m in range(26)], dtype=np.int32)

    def table():
        t = np.zeros((Q, 512), dtype=PAIR_DTYPE)
        t["d"] = rng.random((Q, 512))
        t["id"] = rng.integers(0, n_seq * 26, size=(Q, 512))
        return t
    ta, tt = table(), table()
    ta["d"][3 * 24 + 5, :] = 1e3                              # clip 3, step 5: every audio bin empty
    ta["id"][3 * 24 + 5, :] = -1
    seed_code = rng.integers(0, 512, size=n_clips).astype(np.int32)
    seed_phase = rng.standard_normal((n_clips, 8, 16)).astype(np.float32)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    ta_d, tt_d = to(ta.view(np.int64).reshape(Q, 512, 2)), to(tt.view(np.int64).reshape(Q, 512, 2))
    code_d, phase_d, pos_d, freq_d = to(code), to(phase), to(pos_rank), to(freq_rank)
    pos_t_d = to(pos_rank.T.astype(np.int16))
    fa_d, ft_d, sc_d, sp_d = to(frames_a), to(frames_t), to(seed_code), to(seed_phase)
    sp = _lib.stream_ptr()
    ra = torch.empty((Q, 512), dtype=torch.int32, device=dev)
    rt = torch.empty((Q, 512), dtype=torch.int32, device=dev)
    qfa = torch.empty((Q,), dtype=torch.int32, device=dev)
    qft = torch.empty((Q,), dtype=torch.int32, device=dev)
    _lib.check(lib.qpg_rank512_ties(_lib.ptr(ta_d), Q, _lib.ptr(ra), _lib.ptr(qfa), sp), "rank")
    _lib.check(lib.qpg_rank512_ties(_lib.ptr(tt_d), Q, _lib.ptr(rt), _lib.ptr(qft), sp), "rank")
    # per-window phase statistics for the eight-lanes-per-state transition kernel
    stats = torch.empty((n_seq * 26 * 2 * int(lib.qpg_phase_stats_floats()),), dtype=torch.float32, device=dev)
    _lib.check(lib.qpg_phase_stats(_lib.ptr(phase_d), n_seq, _lib.ptr(fa_d), _lib.ptr(ft_d), _lib.ptr(stats), sp), "stats")
    outs = []
    trans_tables = {}
    for which in ("old", "direct", "table", "table_stats"):
        codes = torch.full((n_clips, n_seg, 30), -7, dtype=torch.int64, device=dev)
        vote = torch.zeros((n_clips, n_seg, 8), dtype=torch.int32, device=dev)
        status = torch.zeros((n_clips,), dtype=torch.int32, device=dev)
        ph = torch.zeros((n_clips, n_seg, 8, 8, 16), dtype=torch.float32, device=dev)
        if which == "old":
            _lib.check(lib.qpg_match_tail(_lib.ptr(ta_d), _lib.ptr(tt_d), _lib.ptr(ra), _lib.ptr(rt), _lib.ptr(pos_d),
                                          _lib.ptr(freq_d), _lib.ptr(code_d), n_seq, _lib.ptr(phase_d), _lib.ptr(fa_d),
                                          _lib.ptr(ft_d), _lib.ptr(sc_d), _lib.ptr(sp_d), n_clips, n_seg, _lib.ptr(codes),
                                          _lib.ptr(vote), _lib.ptr(ph), _lib.ptr(status), sp), "tail")
        else:
            entries = torch.empty((Q, 512, 4), dtype=torch.int64, device=dev)
            _lib.check(lib.qpg_match_lookup(_lib.ptr(ta_d), _lib.ptr(tt_d), _lib.ptr(ra), _lib.ptr(rt), _lib.ptr(pos_t_d),
                                            _lib.ptr(freq_d), _lib.ptr(code_d), n_seq, _lib.ptr(fa_d), _lib.ptr(ft_d),
                                            _lib.ptr(qfa), _lib.ptr(qft), Q, _lib.ptr(entries), sp), "lookup")
            trans = torch.empty((Q, 1024), dtype=torch.int16, device=dev) if which != "direct" else None
            if which == "table_stats":
                _lib.check(lib.qpg_match_walk_stats(_lib.ptr(entries), _lib.ptr(code_d), _lib.ptr(phase_d), _lib.ptr(stats),
                                                    _lib.ptr(sc_d), _lib.ptr(sp_d), n_clips, n_seg, _lib.ptr(trans),
                                                    _lib.ptr(codes), _lib.ptr(vote), _lib.ptr(ph), _lib.ptr(status), sp),
                           "walk_stats")
            else:
                _lib.check(lib.qpg_match_walk(_lib.ptr(entries), _lib.ptr(code_d), _lib.ptr(phase_d), _lib.ptr(sc_d),
                                              _lib.ptr(sp_d), n_clips, n_seg, _lib.ptr(trans), _lib.ptr(codes),
                                              _lib.ptr(vote), _lib.ptr(ph), _lib.ptr(status), sp), "walk")
            if trans is not None:
                trans_tables[which] = trans.cpu().numpy()
        torch.cuda.synchronize()
        outs.append((codes.cpu().numpy(), vote.cpu().numpy(), status.cpu().numpy(), ph.cpu().numpy()))
    (c0, v0, s0, p0), (c1, v1, s1, p1), (c2, v2, s2, p2), (c3, v3, s3, p3) = outs
    # the transition table from per-window statistics equals the one from the full cosines, state for state
    assert np.array_equal(trans_tables["table"], trans_tables["table_stats"])
    assert np.array_equal(c2, c3) and np.array_equal(s2, s3) and np.array_equal(v2, v3) and np.array_equal(p2, p3)
    # direct walk and table walk are the same arithmetic: bit-identical, failed clip included
    assert np.array_equal(c1, c2) and np.array_equal(s1, s2)
    for b in range(n_clips):
        done = n_seg * 8 if b != 3 else 5
        assert np.array_equal(v1[b].ravel()[:done], v2[b].ravel()[:done])
        assert np.array_equal(p1[b].reshape(-1, 8, 16)[:done], p2[b].reshape(-1, 8, 16)[:done])
    assert s0[3] == 1 and (s1[3] & 1) == 1
    ok = [b for b in range(n_clips) if b != 3]
    assert np.all((s1[ok] & 1) == 0) and np.all(s0[ok] == 0)
    assert np.array_equal(c0[ok], c1[ok]) and np.array_equal(v0[ok], v1[ok]) and np.array_equal(p0[ok], p1[ok])
    # the failed clip: identical up to the failing step (steps 0..4 of segment 0 = 20 codes), -1 from there on
    assert np.array_equal(c1[3, 0, :20], c0[3, 0, :20])
    assert np.all(c1[3, 0, 20:] == -1) and np.all(c1[3, 1:] == -1)


def test_tie_flags():
    """status bit 1 is raised when NumPy's tie order could matter: equal distances between non-empty bins, or
    an empty bin that could win; and not raised on tie-free tables"""
    torch, _lib, lib, dev = _env()
    from qpgesture_b200.matchdb import PAIR_DTYPE
    rng = np.random.default_rng(6)
    n_seq, n_seg = 30, 1
    code = rng.integers(0, 512, size=(n_seq, 30)).astype(np.int32)
    phase = rng.standard_normal((n_seq, 240, 16)).astype(np.float32)
    pos_rank = np.stack([rng.permutation(512) for _ in range(512)]).astype(np.int32)
    freq_rank = rng.permutation(512).astype(np.int32)
    frames = np.array([int(6 * m / 398 * 240) for m in range(26)], dtype=np.int32)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    pos_t_d, freq_d, code_d, frames_d, phase_d = to(pos_rank.T.astype(np.int16)), to(freq_rank), to(code), to(frames), to(phase)
    sc_d, sp_d = to(np.array([5], dtype=np.int32)), to(rng.standard_normal((1, 8, 16)).astype(np.float32))

    def run(ta, tt):
        Q = 8
        ta_d, tt_d = to(ta.view(np.int64).reshape(Q, 512, 2)), to(tt.view(np.int64).reshape(Q, 512, 2))
        ra = torch.empty((Q, 512), dtype=torch.int32, device=dev)
        rt = torch.empty((Q, 512), dtype=torch.int32, device=dev)
        qfa = torch.empty((Q,), dtype=torch.int32, device=dev)
        qft = torch.empty((Q,), dtype=torch.int32, device=dev)
        sp = _lib.stream_ptr()
        _lib.check(lib.qpg_rank512_ties(_lib.ptr(ta_d), Q, _lib.ptr(ra), _lib.ptr(qfa), sp), "rank")
        _lib.check(lib.qpg_rank512_ties(_lib.ptr(tt_d), Q, _lib.ptr(rt), _lib.ptr(qft), sp), "rank")
        entries = torch.empty((Q, 512, 4), dtype=torch.int64, device=dev)
        codes = torch.zeros((1, 1, 30), dtype=torch.int64, device=dev)
        vote = torch.zeros((1, 1, 8), dtype=torch.int32, device=dev)
        status = torch.zeros((1,), dtype=torch.int32, device=dev)
        _lib.check(lib.qpg_match_lookup(_lib.ptr(ta_d), _lib.ptr(tt_d), _lib.ptr(ra), _lib.ptr(rt), _lib.ptr(pos_t_d),
                                        _lib.ptr(freq_d), _lib.ptr(code_d), n_seq, _lib.ptr(frames_d), _lib.ptr(frames_d),
                                        _lib.ptr(qfa), _lib.ptr(qft), Q, _lib.ptr(entries), sp), "lookup")
        trans = torch.empty((Q, 1024), dtype=torch.int16, device=dev)
        _lib.check(lib.qpg_match_walk(_lib.ptr(entries), _lib.ptr(code_d), _lib.ptr(phase_d), _lib.ptr(sc_d),
                                      _lib.ptr(sp_d), 1, 1, _lib.ptr(trans), _lib.ptr(codes), _lib.ptr(vote), None,
                                      _lib.ptr(status), sp), "walk")
        torch.cuda.synchronize()
        return int(status.cpu()[0])

    def table(n_empty=0):
        t = np.zeros((8, 512), dtype=PAIR_DTYPE)
        t["d"] = rng.random((8, 512))
        t["id"] = rng.integers(0, n_seq * 26, size=(8, 512))
        if n_empty:
            t["d"][:, -n_empty:] = 1e3
            t["id"][:, -n_empty:] = -1
        return t
    assert run(table(), table()) == 0
    ta = table()
    ta["d"][4, 100] = ta["d"][4, 200]                          # equal distances in two non-empty bins of step 4
    assert run(ta, table()) & 2
    # 508 empty bins: the lowest rank an empty bin can get is 4, so some empty bin beats the best of the 4 non-empty
    # keys under some ordering of the sentinel ties -> flagged (and status 1 if it also wins in the stable order)
    assert run(table(508), table(508)) & 2
    # 400 empty bins: an empty bin has rank >= 112, far behind the best non-empty key -> independent of the tie order
    assert run(table(400), table(400)) == 0


@pytest.mark.parametrize("path", CASES)
def test_plan_engines_agree_and_match_the_reference(path):
    """match_clips through the captured plan: sliced engine == float64 engine == the reference's knn_pred"""
    from qpgesture_b200 import data_processing as dp
    from tests.test_matcher_gpu import _knn_from_case
    fx, train, test, code, sig = load_case(path)
    knn = _knn_from_case("A", train, code, sig, fx, "auto")
    aq = dp.wavlm_query_rows(dp.interpolate_wavlm(test["wavlm"]))
    ctx = test["context"].squeeze(2)
    tq = ctx[:, [int(24 * s / 180 * 30) for s in range(8)], :]
    got = {}
    for engine in ("sliced", "f64"):
        np.random.seed(123456)
        got[engine] = knn.match_clips(aq[None], tq[None], engine=engine)[0]
        p = knn._plans[(1, aq.shape[0], None, engine)]
        assert p.engine == engine and p.graph is not None
        got[engine + "_ids"] = (p.ta[..., 1].cpu().numpy().copy(), p.tt[..., 1].cpu().numpy().copy())
        got[engine + "_rank"] = (p.ra.cpu().numpy().copy(), p.rt.cpu().numpy().copy())
    assert np.array_equal(got["sliced"], fx["knn_pred"])
    assert np.array_equal(got["f64"], fx["knn_pred"])
    # the staged entry (one pinned H2D copy, the captured step, one D2H copy) gives the device tail's result
    p = knn._plans[(1, aq.shape[0], None, "sliced")]
    io = knn.pinned_io(p)
    io.qa.copy_(p.qa.cpu())
    io.qt.copy_(p.qt.cpu())
    io.seed_code.copy_(p.seed_code.cpu())
    io.seed_phase.copy_(p.seed_phase.cpu())
    before = p.codes.cpu().numpy().copy()
    p.codes.fill_(-5)
    try:
        staged = knn.match_staged(p, io).numpy().copy()
        assert np.array_equal(staged, before)
    except IndexError:
        assert int(io.status.max()) & 1
    for k in ("_ids", "_rank"):
        for a, b in zip(got["sliced" + k], got["f64" + k]):
            assert np.array_equal(a, b)


def test_benchmarked_configuration_against_the_oracle():
    """The configuration bench.py times (13 312 windows x 6144-d audio + 384-d text, one 24-s clip = 48 steps,
    CUDA-graph plan, sliced engine): tables vs float64 NumPy over ALL windows (three steps vs sklearn's own
    arithmetic), codes vs the oracle's sequential tail."""
    from qpgesture_b200 import data_processing as dp
    from qpgesture_b200 import synth
    from qpgesture_b200.GestureKNN import CodeKNN
    from qpgesture_b200.matchdb import MatchDatabase, phase_to_dense, table_to_numpy

    n_seq, n_seg = 512, 6
    train, test, code, sig = synth.make_arrays(n_seq, n_seg, seed=0, wavlm_dim=1024, ctx_dim=384)
    aud_rows = dp.wavlm_window_rows(dp.interpolate_wavlm(train["wavlm"]))
    txt_rows = np.ascontiguousarray(train["context"].squeeze(2)[:, :26, :].reshape(n_seq * 26, -1))
    phase_amp = phase_to_dense(train["phase"])
    db = MatchDatabase("A", code, sig, phase_amp, txt_rows, aud_rows=aud_rows)
    knn = CodeKNN(database=db, use_wavlm=True, use_phase=True, use_txt=True, tail="device")
    aq = dp.wavlm_query_rows(dp.interpolate_wavlm(test["wavlm"]))                      # [n_seg, 8, 6144]
    tq = test["context"].squeeze(2)[:, [int(24 * s / 180 * 30) for s in range(8)], :]
    labels = code[:, :26].reshape(-1).astype(np.int64)
    odb = om.OracleDB(mode="A", code=code, labels=labels, aud_rows=np.zeros((0, 1)), txt_rows=np.zeros((0, 1)),
                      aud_k=np.arange(26) * 6, txt_k=np.arange(26) * 8, phase_amp=phase_amp, signature=np.asarray(sig),
                      freq_dist=om.code_to_freq(code), n_db_frm=180, step_sz=6)
    seed = om.init_code_phase(odb, np.random.RandomState(123456))
    got = knn.match_clips(aq[None], tq[None], seed_code=[seed[0]], seed_phase=np.asarray(seed[1])[None], tail="device")[0]
    p = knn._plans[(1, n_seg, None, None)]
    assert p.engine == "sliced" and p.graph is not None and len(p.passes) == 1 and p.passes[0].n_pad == 48
    ta, tt = table_to_numpy(p.ta), table_to_numpy(p.tt)
    want_tabs = []
    for tab, rows, q in ((ta, aud_rows, aq.reshape(48, -1)), (tt, txt_rows, tq.reshape(48, -1))):
        d = sm.f64_distances(rows, q)
        tabs = [om.min_by_code(d[i], labels) for i in range(48)]
        want_tabs.append(tabs)
        for i, (bd, bw) in enumerate(tabs):
            assert np.array_equal(tab[i]["id"], bw), f"step {i}"
            assert np.allclose(tab[i]["d"], bd, rtol=0, atol=2e-7)
    for i in (0, 17, 47):                                      # sklearn's own arithmetic on three steps
        bd, bw = om.min_by_code(om.cosine_rows(aq.reshape(48, -1)[i].astype(np.float64), aud_rows.astype(np.float64)), labels)
        assert np.array_equal(ta[i]["id"], bw)
    tables = [(want_tabs[0][8 * g:8 * g + 8], want_tabs[1][8 * g:8 * g + 8]) for g in range(n_seg)]
    want = om.predict_codes(odb, aq, tq, ties="stable", seed=seed, tables=tables, freq_score=db.freq_rank_host)
    assert np.array_equal(got, want)
    assert int(knn.last_status[0]) == 0
    assert int(p.stats.cpu()[1]) < 48 * 512 * 2 * 0.02        # under 2 % of the bins needed a float64 decision

"""Property tests of the oracle's vectorised restatements against brute-force loops written the way the
reference writes them (no GPU).  These guard the checker itself."""
import numpy as np
from hypothesis import given, settings, strategies as st

from oracle import matcher_np as om
from oracle import ref_harness as rh
from qpgesture_b200.sharding import merge_tables_host, shard_sequences


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(1, 60), st.integers(1, 6))
def test_min_by_code_equals_reference_scan(seed, w, n_levels):
    """min_by_code == the strict-< row-major scan of GestureKNN.py:686-689 (ties keep the first window)."""
    rng = np.random.default_rng(seed)
    dist = rng.integers(0, n_levels, size=w).astype(np.float64)      # few levels -> many exact ties
    dist[rng.random(w) < 0.1] = 2000.0                               # values above the 1e3 sentinel never enter
    labels = rng.integers(0, 8, size=w)
    best = [1e3] * 512
    best_w = [-1] * 512
    for i in range(w):
        if dist[i] < best[labels[i]]:
            best[labels[i]] = dist[i]
            best_w[labels[i]] = i
    bd, bw = om.min_by_code(dist, labels)
    assert np.array_equal(bd, np.array(best)) and np.array_equal(bw, np.array(best_w))


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(1, 5))
def test_levenshtein_rows_equals_scalar_dp(seed, alphabet):
    rng = np.random.default_rng(seed)
    rows = rng.integers(0, alphabet, size=(7, 11))
    q = rng.integers(0, alphabet, size=11)
    want = [rh._levenshtein_distance([int(x) for x in q], [int(x) for x in r]) for r in rows]
    assert np.array_equal(om.levenshtein_rows(q, rows), np.array(want))


@settings(max_examples=30, deadline=None)
@given(st.integers(0, 2**32 - 1))
def test_wavvq_stacking_equals_reference_loops(seed):
    """stack_wavvq == the pre-pad / post-pad concatenations of data_processing.py:304-322, written out."""
    rng = np.random.default_rng(seed)
    wavvq = rng.integers(0, 320, size=(2, 398, 2))
    fi = 398 / 30
    feat1 = np.zeros((2, 398, 6, 2))
    for i in range(6):
        pre = int((6 - i - 1) * fi)
        feat1[:, :, i, :] = np.concatenate((np.zeros((2, pre, 2)), wavvq[:, :398 - pre]), axis=1)
    feat2 = np.zeros((2, 398, 6, 2))
    for i in range(6):
        post = int(i * fi)
        feat2[:, :, i, :] = np.concatenate((wavvq[:, post:], np.zeros((2, post, 2))), axis=1)
    feat2 = np.delete(feat2, 0, axis=2)
    want = np.concatenate((feat1.reshape(2, 398, -1), feat2.reshape(2, 398, -1)), axis=-1)
    assert np.array_equal(om.stack_wavvq(wavvq), want)
    tok = om.wavvq_tokens(want[0, 100])
    f = want[0, 100].reshape(-1, 2).transpose()
    assert np.array_equal(tok, (f[0] * 320 + f[1]).astype(np.int64))


@settings(max_examples=40, deadline=None)
@given(st.integers(0, 2**32 - 1), st.integers(1, 5))
def test_shard_merge_is_order_independent_and_equals_full(seed, parts):
    """Lexicographic (distance, id) merge of per-shard tables == table of the whole database, for any cut."""
    rng = np.random.default_rng(seed)
    n_seq = int(rng.integers(parts, 12))
    w = n_seq * 26
    dist = rng.integers(0, 4, size=w).astype(np.float64)
    labels = rng.integers(0, 16, size=w)
    full_d, full_w = om.min_by_code(dist, labels)
    tabs = []
    for r in range(parts):
        j0, j1 = shard_sequences(n_seq, parts, r)
        d, i = om.min_by_code(dist[26 * j0:26 * j1], labels[26 * j0:26 * j1])
        t = np.zeros((1, 512), dtype=[("d", "<f8"), ("id", "<i8")])
        t["d"][0], t["id"][0] = d, np.where(i >= 0, i + 26 * j0, -1)
        tabs.append(t.view(np.int64).reshape(1, 512, 2))
    order = rng.permutation(parts)
    merged = merge_tables_host(np.stack([tabs[k] for k in order]))
    assert np.array_equal(merged["d"][0], full_d) and np.array_equal(merged["id"][0], full_w)


def test_cosine_rows_is_the_sklearn_pairwise_call():
    from sklearn.metrics.pairwise import paired_distances

    rng = np.random.default_rng(0)
    for dt in (np.float64, np.float32):
        rows = rng.standard_normal((9, 48)).astype(dt)
        rows[3] = 0
        q = rng.standard_normal(48).astype(dt)
        want = np.array([paired_distances([q], [r], metric="cosine")[0] for r in rows])
        got = om.cosine_rows(q, rows)
        assert got.dtype == dt and np.array_equal(got, want)

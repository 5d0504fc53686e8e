"""Pin the oracle (oracle/matcher_np.py) to outputs of the reference itself.

The golden vectors in tests/golden/ were produced by the unmodified reference
(tests/golden/make_golden.py).  These tests need no GPU."""
import numpy as np
import pytest

from oracle import matcher_np as om
from oracle import ref_harness as rh
from tests._common import golden_cases, load_case, oracle_db, oracle_queries

CASES = golden_cases()


def test_fixtures_present():
    assert len(CASES) >= 3


@pytest.mark.parametrize("path", CASES)
def test_tables_mode_a(path):
    fx, train, test, code, sig = load_case(path)
    db = oracle_db("A", train, code, sig)
    aq, tq = oracle_queries("A", test)
    for s in range(8):
        d, w = om.audio_table(db, aq[0, s])
        assert np.array_equal(w, fx["aud_w"][s])
        assert np.array_equal(d, fx["aud_d"][s])          # same sklearn routine -> bit equal
        d, w = om.text_table(db, tq[0, s])
        assert np.array_equal(w, fx["txt_w"][s])
        assert np.array_equal(d, fx["txt_d"][s])


@pytest.mark.parametrize("path", CASES)
def test_tables_mode_b(path):
    fx, train, test, code, sig = load_case(path)
    db = oracle_db("B", train, code, sig)
    aq, _ = oracle_queries("B", test)
    assert aq.shape[1] == fx["lev_d"].shape[0] == 8
    for s in range(8):
        d, w = om.audio_table(db, aq[0, s])
        assert np.array_equal(w, fx["lev_w"][s])
        assert np.array_equal(d, fx["lev_d"][s])


@pytest.mark.parametrize("path", CASES)
def test_end_to_end_mode_a(path):
    fx, train, test, code, sig = load_case(path)
    db = oracle_db("A", train, code, sig)
    aq, tq = oracle_queries("A", test)
    np.random.seed(123456)
    report = []
    got = om.predict_codes(db, aq, tq, ties="numpy", freq_score=fx["freq_rank"], report=report)
    assert not any(r["tie_a"] or r["tie_t"] for r in report), "fixture has a tie at the arg-min (platform defined)"
    assert np.array_equal(got, fx["knn_pred"])


@pytest.mark.parametrize("path", CASES)
def test_seed_draws(path):
    fx, train, test, code, sig = load_case(path)
    db = oracle_db("A", train, code, sig)
    np.random.seed(123456)
    c, p = om.init_code_phase(db)
    assert c == int(fx["init_code"])
    assert np.array_equal(p, fx["init_phase"])


@pytest.mark.parametrize("path", CASES)
def test_mode_b_segment(path):
    fx, train, test, code, sig = load_case(path)
    if fx["codes_b"].shape == (1,):
        pytest.skip("reference raised IndexError on this case")
    db = oracle_db("B", train, code, sig)
    aq, tq = oracle_queries("B", test)
    aud = [om.audio_table(db, aq[0, s]) for s in range(8)]
    txt = [om.text_table(db, tq[0, s]) for s in range(8)]
    codes, _, vote = om.tail_segment(db, aud, txt, int(fx["init_code"]), fx["init_phase"], ties="numpy",
                                     freq_score=fx["freq_rank"])
    # mode-B ranks are dominated by integer ties whose order NumPy leaves platform defined;
    # compare only when this machine reproduces the recorded tie order of the frequency ranks
    if not np.array_equal(om._rank(db.freq_dist, "numpy"), fx["freq_rank"]):
        pytest.skip("NumPy argsort tie order differs from the machine that generated the fixture")
    assert np.array_equal(codes, fx["codes_b"])
    assert np.array_equal(vote, fx["vote_b"])


@pytest.mark.parametrize("path", CASES)
def test_feature_stacking(path):
    fx, train, test, code, sig = load_case(path)
    feat, _ = om.stack_wavlm(train["wavlm"])
    assert np.array_equal(feat[1, [0, 6, 150, 179], :], fx["feat_probe_wavlm"])
    vq = om.stack_wavvq(train["wavvq"])
    assert np.array_equal(vq[1, [0, 13, 200, 397], :], fx["feat_probe_wavvq"])
    vqt = om.stack_wavvq(test["wavvq"])
    assert np.array_equal(vqt[0, [0, 53, 371], :], fx["feat_probe_test_wavvq"])


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
def test_live_reference_small():
    """Fresh seed, live reference run (build container only)."""
    import os
    import tempfile

    from qpgesture_b200 import synth

    train, test, code, sig = synth.make_arrays(10, 1, seed=77, wavlm_dim=8, ctx_dim=8)
    with tempfile.TemporaryDirectory() as root:
        p = synth.write_npz_set(root, train, test, code, sig)
        ref = rh.run_main_codebook(p.as_argv(os.path.join(root, "o.npz")))
    db = oracle_db("A", train, code, sig)
    aq, tq = oracle_queries("A", test)
    np.random.seed(123456)
    assert np.array_equal(om.predict_codes(db, aq, tq, ties="numpy"), ref)

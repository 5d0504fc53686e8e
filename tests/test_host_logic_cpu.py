"""Host-side logic of the product package that needs no GPU: feature stacking / loaders against the
golden probes recorded from the reference, CLI flag surface, layout planner, synthetic-file round trip."""
import os
import re

import numpy as np
import pytest

from oracle import matcher_np as om
from oracle import ref_harness as rh
from qpgesture_b200 import data_processing as dp
from qpgesture_b200 import synth
from qpgesture_b200.matchdb import (code_to_freq, freq_rank_from_code, mode_b_window_frames, phase_frame,
                                    phase_to_dense, pos_rank_table, wavvq_tokens)
from qpgesture_b200.sharding import plan_layout, shard_sequences
from tests._common import golden_cases, load_case

CASES = golden_cases()


@pytest.mark.parametrize("path", CASES)
def test_feature_stacking_matches_reference_probes(path):
    fx, train, test, code, sig = load_case(path)
    interp = dp.interpolate_wavlm(train["wavlm"])                       # same torch call as the reference
    feat = dp.stack_wavlm_feat(interp, dtype=np.float64)
    assert np.array_equal(feat[1, [0, 6, 150, 179], :], fx["feat_probe_wavlm"])
    assert np.array_equal(dp.stack_wavvq_feat(train["wavvq"])[1, [0, 13, 200, 397], :], fx["feat_probe_wavvq"])
    assert np.array_equal(dp.stack_wavvq_feat(test["wavvq"])[0, [0, 53, 371], :], fx["feat_probe_test_wavvq"])
    # the lean window / query builders select exactly the rows the matcher scans
    rows = dp.wavlm_window_rows(interp)
    n = code.shape[0]
    assert np.array_equal(rows.reshape(n, 26, -1), feat[:, 0:156:6, :].astype(np.float32))
    tfeat = dp.stack_wavlm_feat(dp.interpolate_wavlm(test["wavlm"]), dtype=np.float64)
    assert np.array_equal(dp.wavlm_query_rows(dp.interpolate_wavlm(test["wavlm"])),
                          tfeat[:, 0:180:24, :].astype(np.float32))
    # and the oracle's independent restatement of the interpolation agrees bit for bit
    assert np.array_equal(om.interp_linear_align_corners(train["wavlm"], 180), interp)


@pytest.mark.parametrize("path", CASES)
def test_one_off_tables(path):
    fx, train, test, code, sig = load_case(path)
    assert np.array_equal(code_to_freq(code), om.code_to_freq(code))
    fr = freq_rank_from_code(code)
    assert sorted(fr.tolist()) == list(range(512))
    # ranks agree with the recorded ones wherever the frequency is untied (tie order is platform defined)
    f = code_to_freq(code)
    untied = np.array([np.sum(f == v) == 1 for v in f])
    assert np.array_equal(fr[untied], fx["freq_rank"][untied])
    pr = pos_rank_table(sig)
    for last in (0, 17, 511):
        want = om.pos_dist_row(sig, last).argsort().argsort()
        assert np.array_equal(pr[last], want) and pr[last, last] == 511
    ks, ms = mode_b_window_frames()
    assert len(ks) == 26 and ms == list(range(26)) and ks[1] == 13 and ks[25] == 331
    assert phase_frame(150) == 90 and phase_frame(200) == 120
    dense = phase_to_dense(train["phase"][:2])
    obj = phase_to_dense(synth.phase_to_object(train["phase"][:2]))
    assert np.array_equal(dense, obj) and dense.shape == (2, 240, 16)
    tok = wavvq_tokens(np.arange(22.0))
    assert np.array_equal(tok, np.arange(0, 22, 2) * 320 + np.arange(1, 22, 2))


def test_load_db_codebook_contract(tmp_path):
    """Same 19-tuple, axis order and shapes as the reference's load_db_codebook (data_processing.py:345-353)."""
    train, test, code, sig = synth.make_arrays(5, 2, seed=3, wavlm_dim=8, ctx_dim=12)
    p = synth.write_npz_set(str(tmp_path), train, test, code, sig, object_phase=False)
    out = dp.load_db_codebook(p.train_database, p.train_codebook, p.test_data, p.train_wavlm, p.test_wavlm,
                              p.train_wavvq, p.test_wavvq)
    assert len(out) == 19
    shapes = [getattr(o, "shape", None) for o in out]
    assert shapes[0] == (5, 13, 240) and shapes[1] == (5, 30) and shapes[2] == (2, 13, 240)
    assert shapes[3] == (5, 78, 240) and shapes[5] == (5, 8, 180) and shapes[7] == (5, 48, 180) and shapes[8] == (2, 48, 180)
    assert shapes[9] == (5, 3, 240) and shapes[11] == (5, 18, 240)
    assert shapes[13] == (5, 22, 398) and shapes[14] == (2, 22, 398)
    assert shapes[17] == (5, 12, 30) and shapes[18] == (2, 12, 30)
    inp = dp.load_match_inputs(p.train_database, p.train_codebook, p.test_data, p.train_wavlm, p.test_wavlm,
                               p.train_wavvq, p.test_wavvq, mode="A")
    assert inp["aud_rows"].shape == (130, 48) and inp["aud_q"].shape == (2, 8, 48) and inp["txt_q"].shape == (2, 8, 12)
    assert np.array_equal(inp["aud_rows"].reshape(5, 26, 48), out[7].transpose(0, 2, 1)[:, 0:156:6, :])
    inb = dp.load_match_inputs(p.train_database, p.train_codebook, p.test_data, p.train_wavlm, p.test_wavlm,
                               p.train_wavvq, p.test_wavvq, mode="B")
    assert inb["aud_tokens"].shape == (130, 11) and inb["aud_q"].shape == (2, 8, 11)


def test_cli_flags_match_the_reference_script():
    from qpgesture_b200.GestureKNN import build_parser

    ours = {s for a in build_parser()._actions for s in a.option_strings}
    want = {"--train_database", "--test_data", "--out_knn_filename", "--out_video_path", "--train_codebook",
            "--codebook_signature", "--train_wavlm", "--test_wavlm", "--train_wavvq", "--test_wavvq", "--max_frames",
            "--desired_k", "--fake", "--out_fake_knn_filename"}                      # GestureKNN.py:25-39
    assert want <= ours
    sh = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "qpgesture_b200",
                           "GestureKNN.sh")).read()
    assert set(re.findall(r"(--[a-z_]+)=", sh)) == want - {"--desired_k", "--fake", "--out_fake_knn_filename"}
    if rh.available():                                                              # flag-for-flag against the source
        src = open(os.path.join(rh.REF_KNN_DIR, "GestureKNN.py")).read()
        assert set(re.findall(r"add_argument\('-[a-z]+',\s*'(--[a-z_]+)'", src)) | {"--max_frames"} == want
        ref_sh = open(os.path.join(rh.REF_KNN_DIR, "GestureKNN.sh")).read()
        assert set(re.findall(r"(--[a-z_]+)=", ref_sh)) == set(re.findall(r"(--[a-z_]+)=", sh))


def test_layout_planner():
    assert plan_layout(347_000_000, 8) == (1, 8)          # speaker-10-like: replicate, split clips
    assert plan_layout(22_200_000_000, 8) == (8, 1)       # all-speaker-like: 8 row shards
    assert plan_layout(22_200_000_000, 1) == (1, 1)
    assert plan_layout(5_000_000_000, 8) == (4, 2)
    for w in (1, 2, 4, 8):
        rs, cg = plan_layout(3_000_000_000, w)
        assert rs * cg == w
    assert [shard_sequences(10, 4, r) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


@pytest.mark.parametrize("path", CASES)
def test_product_numpy_tail_reproduces_reference_codes(path):
    """CodeKNN._tail_numpy_segment (product host code, tail='numpy') driven by oracle tables through a stub
    database reproduces the reference's end-to-end knn_pred: the state machine itself is checked on CPU."""
    from types import SimpleNamespace

    from qpgesture_b200.GestureKNN import CodeKNN
    from qpgesture_b200.matchdb import PAIR_DTYPE, phase_frame  # noqa: F401
    from tests._common import oracle_db, oracle_queries

    fx, train, test, code, sig = load_case(path)
    odb = oracle_db("A", train, code, sig)
    aq, tq = oracle_queries("A", test)
    aud_k, txt_k = [6 * m for m in range(26)], [8 * m for m in range(26)]
    stub = SimpleNamespace(
        pos_rank_host=pos_rank_table(sig), freq_rank_host=fx["freq_rank"].astype(np.int32),
        phase_amp_host=phase_to_dense(train["phase"]),
        payload=lambda w: code[int(w) // 26, int(w) % 26:int(w) % 26 + 4],
        aux=lambda w, which: [int(w) // 26, (aud_k if which == "audio" else txt_k)[int(w) % 26]])
    knn = SimpleNamespace(db=stub)

    def table(fn, q):
        t = np.zeros((8, 512), dtype=PAIR_DTYPE)
        for s in range(8):
            t["d"][s], t["id"][s] = fn(odb, q[s])
        return t

    np.random.seed(123456)
    code0, ph0 = om.init_code_phase(odb)
    got = []
    for g in range(aq.shape[0]):
        codes, phases, _ = CodeKNN._tail_numpy_segment(knn, table(om.audio_table, aq[g]), table(om.text_table, tq[g]),
                                                       code0, ph0)
        got.append(codes)
        code0, ph0 = int(codes[-1]), phases[-1]
    assert np.array_equal(np.array(got), fx["knn_pred"])

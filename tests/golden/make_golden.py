"""Generate tests/golden/*.npz by running the UNMODIFIED reference in place.

Runs only in the build container (needs /root/reference).  Usage:
    python tests/golden/make_golden.py
Inputs are seeded synthetic arrays (qpgesture_b200.synth.make_arrays); each
fixture stores the generator arguments and a sha256 of the inputs so that the
tests regenerate them and fail loudly if NumPy's stream ever changes.

What is recorded per case (all produced by the reference's own code):
  knn_pred            main_codebook end-to-end (GestureKNN.py:816-845), mode A
  aud_d/aud_w         CodeKNN.search_audio_cands for the 8 steps of segment 0
  txt_d/txt_w         CodeKNN.search_text_cands, same steps
  lev_d/lev_w         mode-B search_audio_cands ('wavvq_feat') for 8 steps
  codes_b             mode-B search_code_knn with explicit seeds (segment 0)
  freq_rank           np.array(freq_dist_cands).argsort().argsort() on this machine
  init_code/init_phase the two RNG draws of init_code_phase under seed 123456
  feat_probe          rows of load_db_codebook's stacked WavLM / wavvq features
Versions: numpy, scikit-learn, torch of the build container are stored too.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import matcher_np as om  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
from qpgesture_b200 import synth  # noqa: E402

CASES = [
    dict(name="matcher_s0", n_train=24, n_test=2, seed=0, wavlm_dim=16, ctx_dim=24),
    dict(name="matcher_s1", n_train=40, n_test=3, seed=1, wavlm_dim=24, ctx_dim=16),
    dict(name="matcher_s2", n_train=64, n_test=2, seed=2, wavlm_dim=8, ctx_dim=32),
]


def inputs_digest(train, test, code, sig) -> str:
    h = hashlib.sha256()
    for split in (train, test):
        for k in sorted(split):
            h.update(np.ascontiguousarray(split[k]).tobytes())
    h.update(np.ascontiguousarray(code).tobytes())
    h.update(np.ascontiguousarray(sig).tobytes())
    return h.hexdigest()


def windows_from_aux(aux, div):
    return np.array([(-1 if len(a) == 0 else 26 * a[0] + a[1] // div) for a in aux], dtype=np.int64)


def run_case(case, out_dir):
    import sklearn
    import torch

    kw = {k: case[k] for k in ("n_train", "n_test", "seed", "wavlm_dim", "ctx_dim")}
    train, test, code, sig = synth.make_arrays(**kw)
    root = tempfile.mkdtemp(prefix="qpg_golden_")
    try:
        paths = synth.write_npz_set(root, train, test, code, sig, object_phase=True)
        flags = paths.as_argv(os.path.join(root, "out.npz"))
        rec = dict(digest=inputs_digest(train, test, code, sig),
                   versions=np.array([np.__version__, sklearn.__version__, torch.__version__]))
        for k, v in kw.items():
            rec["arg_" + k] = v

        # ---- end to end, mode A (the shipped path)
        rec["knn_pred"] = rh.run_main_codebook(flags)

        # ---- function level, mode A
        mod, knn, q = rh.build_codeknn(flags, mode="A")
        rec["freq_rank"] = np.array(knn.freq_dist_cands).argsort().argsort()
        np.random.seed(123456)
        ic, ip = knn.init_code_phase()
        rec["init_code"], rec["init_phase"] = int(ic), np.asarray(ip)
        clip = q["test_wavlm_feat"][0]
        ctx = q["test_context"][0]
        aud_d, aud_w, txt_d, txt_w = [], [], [], []
        for s in range(8):
            i = 24 * s
            with rh._quiet():
                d, _, aux = knn.search_audio_cands(clip[i], mode="wavlm_feat")
                d2, _, aux2 = knn.search_text_cands(ctx[int(i / 180 * 30)])
            aud_d.append(np.array(d, dtype=np.float64))
            aud_w.append(windows_from_aux(aux, 6))
            txt_d.append(np.array(d2, dtype=np.float64))
            txt_w.append(windows_from_aux(aux2, 8))
        rec.update(aud_d=np.array(aud_d), aud_w=np.array(aud_w), txt_d=np.array(txt_d), txt_w=np.array(txt_w))
        rec["feat_probe_wavlm"] = np.asarray(q["train_wavlm_feat"][1, [0, 6, 150, 179], :])
        rec["feat_probe_wavvq"] = np.asarray(q["train_wavvq_feat"][1, [0, 13, 200, 397], :])
        rec["feat_probe_test_wavvq"] = np.asarray(q["test_wavvq_feat"][0, [0, 53, 371], :])
        rh.release_gestureknn()

        # ---- mode B (wavvq Levenshtein): tables + one segment with explicit seeds
        mod, knn, q = rh.build_codeknn(flags, mode="B")
        clipb = q["test_wavvq_feat"][0]
        lev_d, lev_w = [], []
        step = 4 * (398 / 30)
        i_list, i = [], 0
        while i < 398:
            i_list.append(i)
            i += step
        for i in i_list:
            with rh._quiet():
                d, _, aux = knn.search_audio_cands(clipb[int(i)], mode="wavvq_feat")
            lev_d.append(np.array(d, dtype=np.float64))
            ks, _ = om.mode_b_window_k()
            kmap = {int(k): m for m, k in enumerate(ks)}
            lev_w.append(np.array([(-1 if len(a) == 0 else 26 * a[0] + kmap[a[1]]) for a in aux], dtype=np.int64))
        rec.update(lev_d=np.array(lev_d), lev_w=np.array(lev_w))
        try:
            with rh._quiet():
                codes_b, _, vote_b = knn.search_code_knn(
                    clip_test=clipb, desired_k=0, use_wavlm=False, use_feature=True, use_freq=False,
                    seed_code=rec["init_code"], use_wavvq=True, use_phase=True, seed_phase=rec["init_phase"],
                    use_txt=True, clip_context=q["test_context"][0], use_aud=True)
            rec["codes_b"], rec["vote_b"] = codes_b, vote_b
        except IndexError:
            rec["codes_b"] = np.array([-1])           # reference raised (a chosen code had no window)
        rh.release_gestureknn()
        np.savez_compressed(os.path.join(out_dir, case["name"] + ".npz"), **rec)
        print(case["name"], "knn_pred", rec["knn_pred"].shape, "digest", rec["digest"][:12])
    finally:
        shutil.rmtree(root, ignore_errors=True)


if __name__ == "__main__":
    assert rh.available(), "needs /root/reference"
    for c in CASES:
        run_case(c, HERE)

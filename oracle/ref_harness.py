"""Run the UNMODIFIED reference in place (build container only).

TEST INFRASTRUCTURE.  /root/reference is read-only and does not exist on the
GPU box, so this module is used only (a) by tests/golden/make_golden.py to
produce the committed golden vectors, and (b) by `-m "not gpu"` tests that are
skipped when /root/reference is absent.  Recipe: SURVEY.md 8(c).

Nothing is copied from the reference: its own files are imported / executed
from where they lie, with three tiny stub modules for dependencies that are
not installed here (python-Levenshtein, matplotlib-based `visualization`,
`configargparse`).
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

REF_ROOT = os.environ.get("QPG_REFERENCE_ROOT", "/root/reference")
REF_KNN_DIR = os.path.join(REF_ROOT, "codebook", "Speech2GestureMatching")
REF_CODEBOOK_DIR = os.path.join(REF_ROOT, "codebook")

_KNN_LOCAL_MODULES = ("GestureKNN", "data_processing", "control", "utils", "constant",
                      "visualization", "Levenshtein")
_VQ_LOCAL_MODULES = ("models", "configs", "configargparse")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_KNN_DIR, "GestureKNN.py"))


def _levenshtein_distance(a, b) -> int:
    """Unit-cost edit distance (what python-Levenshtein's `distance` computes;
    the package itself is not installed here).  Integers, so no parity risk."""
    la, lb = len(a), len(b)
    prev = list(range(lb + 1))
    for i in range(1, la + 1):
        cur = [i] + [0] * lb
        ai = a[i - 1]
        for j in range(1, lb + 1):
            cost = 0 if ai == b[j - 1] else 1
            cur[j] = min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + cost)
        prev = cur
    return prev[lb]


def _purge(names):
    for n in list(sys.modules):
        if n in names or any(n.startswith(p + ".") for p in names):
            del sys.modules[n]


@contextlib.contextmanager
def _quiet(enabled=True):
    if not enabled:
        yield
        return
    buf_out, buf_err = io.StringIO(), io.StringIO()
    with contextlib.redirect_stdout(buf_out), contextlib.redirect_stderr(buf_err):
        yield


def import_gestureknn(argv_flags, quiet=True):
    """Import the reference's GestureKNN.py as a fresh module with `argv_flags`
    as its command line (it runs argparse at import, GestureKNN.py:41, and
    seeds numpy/random with 123456, :19-22).  Returns the module."""
    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    _purge(_KNN_LOCAL_MODULES)
    lev = types.ModuleType("Levenshtein")
    lev.distance = _levenshtein_distance
    sys.modules["Levenshtein"] = lev
    vis = types.ModuleType("visualization")
    vis.generate_seq_videos = lambda *a, **k: None
    sys.modules["visualization"] = vis
    old_argv, old_path = sys.argv, list(sys.path)
    sys.argv = ["GestureKNN.py"] + list(argv_flags)
    sys.path.insert(0, REF_KNN_DIR)
    try:
        with _quiet(quiet):
            mod = importlib.import_module("GestureKNN")
    finally:
        sys.argv = old_argv
        sys.path[:] = old_path
    return mod


def release_gestureknn():
    _purge(_KNN_LOCAL_MODULES)


def run_main_codebook(argv_flags, max_frames=0, quiet=True, mode="A"):
    """Execute the reference's main_codebook (GestureKNN.py:816) end to end.
    mode "A" is the shipped literal set (:842-843, WavLM cosine + text + phase).
    Returns knn_pred as written by np.savez_compressed (:845)."""
    import numpy as np

    mod = import_gestureknn(argv_flags, quiet=quiet)
    try:
        assert mode == "A", "mode B needs edited literals; use build_codeknn(mode='B')"
        with _quiet(quiet):
            mod.main_codebook(maxFrames=max_frames)
        out = [a.split("=", 1)[1] for a in argv_flags if a.startswith("--out_knn_filename=")][0]
        return np.load(out)["knn_pred"]
    finally:
        release_gestureknn()


def _load_db_ref(mod, quiet=True):
    sys.path.insert(0, REF_KNN_DIR)
    try:
        dp = importlib.import_module("data_processing")
    finally:
        sys.path.pop(0)
    a = mod.args
    with _quiet(quiet):
        return dp.load_db_codebook(a.train_database, a.train_codebook, a.test_data, a.train_wavlm,
                                   a.test_wavlm, a.train_wavvq, a.test_wavvq)


def build_codeknn(argv_flags, mode="A", quiet=True):
    """Construct the reference CodeKNN the way predict_code_from_audio does
    (GestureKNN.py:744-778).  mode "A": use_wavlm; mode "B": use_wavvq.
    Returns (module, knn, dict of per-test-sequence query arrays)."""
    mod = import_gestureknn(argv_flags, quiet=quiet)
    (train_mfcc, train_code, test_mfcc, train_feat, test_feat, train_wavlm, test_wavlm,
     train_wavlm_feat, test_wavlm_feat, speech_features, test_speech_features,
     train_speech_features_feat, test_speech_features_feat, train_wavvq_feat, test_wavvq_feat,
     train_phase, test_phase, train_context, test_context) = _load_db_ref(mod, quiet)
    tr = lambda x: x.transpose((0, 2, 1))
    use_wavlm = mode == "A"
    with _quiet(quiet):
        knn = mod.CodeKNN(mfcc_train=tr(train_mfcc), code_train=train_code, feat_train=tr(train_feat),
                          wavlm_train=tr(train_wavlm), wavlm_train_feat=tr(train_wavlm_feat),
                          speech_features=tr(speech_features),
                          speech_features_feat=tr(train_speech_features_feat),
                          wavvq_train_feat=tr(train_wavvq_feat), phase_train=tr(train_phase),
                          context_train=tr(train_context), use_wavlm=use_wavlm,
                          use_wavvq=not use_wavlm, use_phase=True, use_txt=True)
    q = dict(test_wavlm_feat=tr(test_wavlm_feat), test_wavvq_feat=tr(test_wavvq_feat),
             test_context=tr(test_context), train_wavlm_feat=tr(train_wavlm_feat),
             train_wavvq_feat=tr(train_wavvq_feat), train_wavlm=tr(train_wavlm))
    return mod, knn, q


def import_vqvae(quiet=True):
    """Import the reference's models.vqvae / models.bottleneck on CPU
    (SURVEY.md 8(c) recipe: stub configargparse, set argv, patch mydevice).
    Returns (vqvae_module, bottleneck_module)."""
    import torch

    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    _purge(_VQ_LOCAL_MODULES)
    sys.modules["configargparse"] = types.ModuleType("configargparse")
    old_argv, old_path = sys.argv, list(sys.path)
    sys.argv = ["x", "--config", os.path.join(REF_CODEBOOK_DIR, "configs", "codebook.yml"), "--gpu", "0"]
    sys.path.insert(0, REF_CODEBOOK_DIR)
    try:
        with _quiet(quiet):
            B = importlib.import_module("models.bottleneck")
            V = importlib.import_module("models.vqvae")
    finally:
        sys.argv = old_argv
        sys.path[:] = old_path
    B.mydevice = V.mydevice = torch.device("cpu")
    return V, B


def release_vqvae():
    _purge(_VQ_LOCAL_MODULES)


_PAE_LOCAL_MODULES = ("PAE", "Library", "data_loader", "configs", "easydict", "configargparse")


def import_pae(quiet=True):
    """Import the reference's codebook/PAE.py on CPU.  Its module-level imports pull in plotting, the AdamWR
    optimiser, the lmdb loader and easydict, none of which `Model` / `pose2phase` (PAE.py:50-162, 477-508) use:
    they are replaced by empty stub modules (matplotlib, lmdb and easydict are not installed here).
    `pose2phase` reads the module global `mydevice` that only `__main__` sets (:513): patched to CPU."""
    import torch

    if not available():
        raise RuntimeError("reference not present at " + REF_ROOT)
    _purge(_PAE_LOCAL_MODULES)
    saved = {}

    def stub(name, **attrs):
        saved[name] = sys.modules.get(name)
        m = types.ModuleType(name)
        m.__path__ = []
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    stub("Library")
    stub("Library.Utility")
    stub("Library.Plotting")
    stub("Library.AdamWR")
    stub("Library.AdamWR.adamw")
    stub("Library.AdamWR.cyclic_scheduler")
    if "matplotlib" not in sys.modules:
        try:
            importlib.import_module("matplotlib.pyplot")
        except Exception:
            stub("matplotlib")
            stub("matplotlib.pyplot")
    stub("data_loader")
    stub("data_loader.lmdb_data_loader", TrinityDataset=object)
    stub("easydict", EasyDict=dict)
    stub("configs")
    stub("configs.parse_args", parse_args=lambda: None)
    old_path = list(sys.path)
    sys.path.insert(0, REF_CODEBOOK_DIR)
    try:
        with _quiet(quiet):
            mod = importlib.import_module("PAE")
    finally:
        sys.path[:] = old_path
    mod.mydevice = torch.device("cpu")
    return mod


def release_pae():
    _purge(_PAE_LOCAL_MODULES + ("matplotlib",) if "matplotlib" in sys.modules and
           not hasattr(sys.modules["matplotlib"], "__version__") else _PAE_LOCAL_MODULES)

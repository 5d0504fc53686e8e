"""Shared helpers for the test-suite (fixtures are regenerated from their seeds)."""
import glob
import hashlib
import os

import numpy as np

from oracle import matcher_np as om
from qpgesture_b200 import synth

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden_cases():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "matcher_*.npz")))


def digest(train, test, code, sig) -> str:
    h = hashlib.sha256()
    for split in (train, test):
        for k in sorted(split):
            h.update(np.ascontiguousarray(split[k]).tobytes())
    h.update(np.ascontiguousarray(code).tobytes())
    h.update(np.ascontiguousarray(sig).tobytes())
    return h.hexdigest()


def load_case(path):
    """-> (fixture dict, train, test, code, signature); checks the input digest."""
    fx = dict(np.load(path, allow_pickle=False))
    kw = {k[4:]: int(fx[k]) for k in fx if k.startswith("arg_")}
    train, test, code, sig = synth.make_arrays(**kw)
    assert digest(train, test, code, sig) == str(fx["digest"]), \
        "synthetic inputs differ from the ones the golden vectors were generated on"
    return fx, train, test, code, sig


def oracle_db(mode, train, code, sig):
    return om.build_db(mode, code, sig, train["phase"], train["context"], wavlm=train["wavlm"], wavvq=train["wavvq"])


def oracle_queries(mode, test):
    return om.build_queries(mode, test["context"], test_wavlm=test["wavlm"], test_wavvq=test["wavvq"])

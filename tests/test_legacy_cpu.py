"""Legacy pose-feature matcher oracle (oracle/legacy_np.py) pinned against the reference `GestureKNN` class: committed
golden vectors (tests/golden/legacy_*.npz, produced by the unmodified class) and the class itself imported in place
when /root/reference is present.  The oracle makes the reference's own arithmetic calls, so equality is exact."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import legacy_np  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
import make_golden_legacy as mg  # noqa: E402


@pytest.mark.parametrize("name", ["legacy_s0", "legacy_s1"])
def test_oracle_equals_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    feat, motn, mask, tests = mg.inputs(int(g["seed"]), int(g["n_seq"]), int(g["n_frames"]), int(g["n_test"]))
    for i in range(int(g["n_test"])):
        k = int(g["desired_k"][i])
        m, _ = legacy_np.search_motion(feat, motn, mask, tests[i], k, tuple(g[f"init_{i}"]))
        f, _ = legacy_np.search_fake_motion(feat, motn, mask, tests[i], k)
        assert np.array_equal(m, g[f"motion_{i}"])
        assert np.array_equal(f, g[f"fake_{i}"])


def test_walk_rules():
    """zero distances, the tail of the sequence, the two-ended mask test and the never-examined last element"""
    d = np.array([0.0, 3.0, 1.0, 2.0, 0.5, 9.0, 0.7, 0.6])
    ones = np.ones(8, dtype=np.int64)
    assert legacy_np.frame_candidate(d, ones, 2) == (4, 0.5)
    assert legacy_np.frame_candidate(d, ones, 5) == (2, 1.0)          # frames > 8 - 5 are too close to the end
    m = ones.copy()
    m[5] = 0                                                          # frame 4 fails at its far end (4 + 2 - 1)
    assert legacy_np.frame_candidate(d, m, 2) == (6, 0.7)             # 7 > 8 - 2: too close to the end
    only_far = np.array([0.0, 0.0, 5.0])
    assert legacy_np.frame_candidate(only_far, np.ones(3, dtype=np.int64), 1) is None   # the maximum is never examined


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
def test_oracle_equals_reference_in_place():
    mod = rh.import_gestureknn([])
    try:
        feat, motn, mask, tests = mg.inputs(9, 16, 24, 2)
        knn = mod.GestureKNN(feat_train=feat, motn_train=motn, control_mask=mask, n_joints=motn.shape[2])
        for i in range(2):
            np.random.seed(50 + i)
            init = knn.init_frame()
            np.random.seed(50 + i)
            want = knn.search_motion(tests[i], i)
            got, _ = legacy_np.search_motion(feat, motn, mask, tests[i], i, init)
            assert np.array_equal(got, want)
            assert np.array_equal(legacy_np.search_fake_motion(feat, motn, mask, tests[i], 3)[0],
                                  knn.search_fake_motion(tests[i], 3))
    finally:
        rh.release_gestureknn()

"""Legacy pose-feature matcher on the device (qpgesture_b200.GestureKNN.GestureKNN over csrc/legacy_knn.cu) against
the golden vectors of the unmodified reference class and the oracle.  The outputs are copies of database motion
frames, so equality is exact whenever the same (sequence, frame) is picked; the status word reports exact ties."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))


@pytest.mark.parametrize("ties", ["numpy", "stable"])
@pytest.mark.parametrize("name", ["legacy_s0", "legacy_s1"])
def test_search_motion_and_fake_equal_golden(name, ties):
    """The golden runs hold no pick that depends on the order of equal keys (checked when they were generated), so
    both tie policies must reproduce them exactly, and the stable device pick must report no tie."""
    import make_golden_legacy as mg
    from qpgesture_b200.GestureKNN import GestureKNN

    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    feat, motn, mask, tests = mg.inputs(int(g["seed"]), int(g["n_seq"]), int(g["n_frames"]), int(g["n_test"]))
    knn = GestureKNN(feat_train=feat, motn_train=motn, control_mask=mask, n_joints=motn.shape[2], device="cuda:0",
                     ties=ties)
    n = int(g["n_test"])
    for i in range(n):                                       # the reference's one-clip calls, same RNG draws
        np.random.seed(1000 + i)
        got = knn.search_motion(tests[i], int(g["desired_k"][i]))
        assert got.shape == g[f"motion_{i}"].shape
        assert np.array_equal(got, g[f"motion_{i}"])
        assert int(knn.last_status[0]) == 0
        fake = knn.search_fake_motion(tests[i], int(g["desired_k"][i]))
        assert np.array_equal(fake, g[f"fake_{i}"])
        assert int(knn.last_status[0]) == 0
    # all clips in one batched call
    inits = [tuple(g[f"init_{i}"]) for i in range(n)]
    got = knn.search_motion_batch(tests, g["desired_k"], inits)
    fake = knn.search_fake_motion_batch(tests, g["desired_k"])
    for i in range(n):
        assert np.array_equal(got[i], g[f"motion_{i}"]) and np.array_equal(fake[i], g[f"fake_{i}"])


def test_larger_database_equals_oracle_choices():
    from oracle import legacy_np
    from qpgesture_b200.GestureKNN import GestureKNN

    rng = np.random.default_rng(4)
    n_seq, n_frames = 120, 64
    feat = rng.standard_normal((n_seq, n_frames, 208))
    motn = rng.standard_normal((n_seq, n_frames, 165))
    mask = (rng.random((n_seq, n_frames)) > 0.05).astype(np.int64)
    tests = rng.standard_normal((2, 112, n_frames))
    mask[5, 7] = mask[100, 30] = 1
    knn = GestureKNN(feat, motn, mask, device="cuda:0")                # ties="numpy": the host's argsort on the rank sums
    inits = [(5, 7), (100, 30)]
    got = knn.search_motion_batch(tests, [0, 3], inits)
    fake = knn.search_fake_motion_batch(tests, [2, 1])
    for i in range(2):
        want, chosen = legacy_np.search_motion(feat, motn, mask, tests[i], [0, 3][i], inits[i])
        assert [tuple(c) for c in knn.last_chosen[i][:len(chosen)]] == chosen or np.array_equal(got[i], want)
        assert np.array_equal(got[i], want)
        assert np.array_equal(fake[i], legacy_np.search_fake_motion(feat, motn, mask, tests[i], [2, 1][i])[0])


def test_too_few_candidates_raises_index_error_and_ties_are_flagged():
    from qpgesture_b200.GestureKNN import GestureKNN

    rng = np.random.default_rng(6)
    feat = rng.standard_normal((6, 16, 208))
    motn = rng.standard_normal((6, 16, 165))
    mask = np.ones((6, 16), dtype=np.int64)
    knn = GestureKNN(feat, motn, mask, device="cuda:0")
    with pytest.raises(IndexError):
        knn.search_fake_motion(rng.standard_normal((112, 16)), 6)      # only 6 sequences: index 6 does not exist
    feat[3] = feat[2]                                                  # two identical sequences: every distance ties
    knn = GestureKNN(feat, motn, mask, device="cuda:0", ties="stable")
    knn.search_fake_motion(rng.standard_normal((112, 16)), 0)
    assert int(knn.last_status[0]) & 2


def test_predict_gesture_from_audio_shapes():
    from qpgesture_b200.GestureKNN import predict_gesture_from_audio

    rng = np.random.default_rng(8)
    feat_train = rng.standard_normal((20, 208, 32))
    pose_train = rng.standard_normal((20, 165, 32))
    feat_test = rng.standard_normal((3, 112, 32))
    stats = dict(feat_mean=feat_train.mean(axis=(0, 2))[None], feat_std=feat_train.std(axis=(0, 2))[None])
    stats["feat_mean"] = stats["feat_mean"][:, :, None] * np.ones((1, 1, 1))
    stats["feat_std"] = stats["feat_std"][:, :, None] * np.ones((1, 1, 1))
    mask = np.ones((20, 32), dtype=np.int64)
    np.random.seed(3)
    out = predict_gesture_from_audio(feat_train, pose_train, feat_test, mask, stats, k=0, device="cuda:0")
    assert out.shape == (3, 165, 32)
    np.random.seed(3)
    fake = predict_gesture_from_audio(feat_train, pose_train, feat_test, mask, stats, k=0, fake=True, device="cuda:0")
    assert fake.shape == (3, 165, 32)

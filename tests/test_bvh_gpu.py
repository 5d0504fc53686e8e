"""Post-decode pose processing on the device (qpgesture_b200/process_bvh.py over csrc/pose_post.cu) against scipy's
savgol_filter and Rotation, the calls of process/process_bvh.py:57-77.  Float64 path; tolerances in degrees."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def decoder_like_poses(T, J, seed, noise):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(seed)
    m = R.random(T * J, random_state=seed).as_matrix() + rng.standard_normal((T * J, 3, 3)) * noise
    return m.reshape(T, J * 9).astype(np.float32)


def smooth_poses(T, J, noise, seed=3):
    """a smooth rotation trajectory (smoothing frame-to-frame random rotations gives singular matrices)"""
    from scipy.spatial.transform import Rotation as R, Slerp
    base = Slerp([0, 1], R.random(2, random_state=seed))(np.linspace(0, 1, T)).as_matrix()
    rng = np.random.default_rng(seed + 6)
    return (np.repeat(base[:, None], J, axis=1) + rng.standard_normal((T, J, 3, 3)) * noise).reshape(T, J * 9).astype(np.float32)


@pytest.mark.parametrize("T,J,noise,smoothing", [(40, 15, 0.05, True), (15, 15, 0.02, True), (64, 15, 0.05, False),
                                                 (240, 15, 0.0, False)])
def test_euler_angles_equal_scipy(T, J, noise, smoothing):
    import warnings
    from oracle import bvh_np
    from qpgesture_b200.process_bvh import poses_to_euler

    poses = decoder_like_poses(T, J, T + J, noise)
    if smoothing:
        poses = smooth_poses(T, J, noise)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        want64 = bvh_np.poses_to_euler(poses.astype(np.float64), smoothing=smoothing)
        want32 = bvh_np.poses_to_euler(poses, smoothing=smoothing)
        got = poses_to_euler(poses, smoothing=smoothing, device="cuda:0")
    assert got.shape == want64.shape == (T, J * 3) and got.dtype == np.float64
    ang = lambda a, b: np.abs((a - b + 180.0) % 360.0 - 180.0)
    # same float32 poses, scipy in float64: the device path is float64 throughout
    assert ang(got, want64).max() < 1e-8, f"max |euler difference| = {ang(got, want64).max()} deg"
    # the reference's own call hands float32 poses to savgol_filter, which then filters in single precision
    # (scipy keeps float32): its angles carry ~1e-5 degrees of float32 noise relative to the float64 filter
    assert ang(got, want32).max() < 2e-4


def test_errors_and_fallback_file(tmp_path):
    from qpgesture_b200.process_bvh import make_bvh_GENEA2020_BT, poses_to_euler

    poses = smooth_poses(20, 15, 0.02)
    with pytest.raises(ValueError):
        poses_to_euler(poses[:10], smoothing=True, device="cuda:0")           # window longer than the sequence
    bad = poses.copy()
    bad[3, 9:18] = np.array([[1, 0, 0], [0, 1, 0], [0, 0, -1]], dtype=np.float32).ravel()     # determinant -1
    with pytest.raises(ValueError):
        poses_to_euler(bad, smoothing=False, device="cuda:0")
    with pytest.warns(UserWarning):
        out = make_bvh_GENEA2020_BT(str(tmp_path), "clip", poses, smoothing=True, pipeline_path=str(tmp_path / "none.sav"),
                                    device="cuda:0")
    assert out.endswith("_generated_euler.npy") and np.load(out).shape == (20, 45)


def test_visualize_code_and_write_runs_the_bvh_step(tmp_path, monkeypatch):
    """VisualizeCodebook.visualizeCodeAndWrite (:333-361): decoded poses -> make_bvh_GENEA2020_BT(smoothing=False);
    without pymo / the fitted pipeline the Euler angles land next to where the BVH file would go."""
    import warnings
    from types import SimpleNamespace
    from oracle import bvh_np
    from qpgesture_b200 import VisualizeCodebook as VC

    poses = smooth_poses(24, 15, 0.03)
    monkeypatch.setattr(VC, "visualize_code", lambda *a, **k: (poses, np.zeros((1, 3), dtype=np.int64)))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out_poses, _ = VC.visualizeCodeAndWrite(code_source=np.zeros((1, 3), dtype=np.int64), save_path=str(tmp_path),
                                                prefix="clip", pipeline_path=str(tmp_path / "missing.sav"),
                                                generateGT=False, vis=False, config=SimpleNamespace())
        want = bvh_np.poses_to_euler(poses.astype(np.float64), smoothing=False)
    got = np.load(tmp_path / "clip" / "clip_generated_euler.npy")
    assert out_poses is poses and got.shape == (24, 45)
    assert np.abs((got - want + 180.0) % 360.0 - 180.0).max() < 1e-8

"""Periodic auto-encoder oracle (oracle/pae_np.py) pinned against the reference: the committed golden vectors
(tests/golden/pae_*.npz, produced by the unmodified reference Model / pose2phase) and, when /root/reference is
present, the reference module itself imported in place.  Floating-point path: tolerances below."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))
from oracle import pae_np  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402
import make_golden_pae as mg  # noqa: E402

# float32 reference vs float64 restatement: sums of 32 400 products, then tanh / FFT / atan
TOL_LATENT = 2e-5
TOL_PHASE = 2e-5      # turns, compared cyclically
TOL_FREQ = 2e-4       # frequencies are ~14 (cycles per window second), float32 ratio of sums
TOL_AMP = 2e-6


def cyc(a, b):
    return np.abs((np.asarray(a, dtype=np.float64) - b + 0.5) % 1.0 - 0.5)


@pytest.mark.parametrize("name", ["pae_s0", "pae_s1"])
def test_oracle_matches_golden(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    sd = pae_np.random_state_dict(int(g["seed"]))
    pose, mean, std, xw = mg.inputs(int(g["seed"]), int(g["T"]))
    got = pae_np.pose2phase(sd, pose, mean, std)
    ref = g["phase"]
    assert got.shape == ref.shape == (int(g["T"]), 4, 1, 8, 1)
    assert cyc(got[:, 0], ref[:, 0]).max() < TOL_PHASE
    assert np.abs(got[:, 1] - ref[:, 1]).max() < TOL_FREQ
    assert np.abs(got[:, 2:] - ref[:, 2:]).max() < TOL_AMP
    y, latent, signal, params = pae_np.forward(sd, xw)
    assert np.abs(latent - g["latent"]).max() < TOL_LATENT
    assert np.abs(signal - g["signal"]).max() < 5e-4          # sin(2 pi (f t + p)) with f t up to ~30 turns
    assert np.abs(y[:, :2048] - g["y_head"]).max() < 5e-4
    par = np.stack(params, axis=1)
    assert cyc(par[:, 0], g["params"][:, 0]).max() < TOL_PHASE
    assert np.abs(par[:, 2:] - g["params"][:, 2:]).max() < TOL_AMP


def test_windows_follow_the_reference_layout():
    """Window i = one zero frame, then padded frame differences i .. i+238 (PAE.py:481-499)."""
    rng = np.random.default_rng(0)
    pose = rng.standard_normal((5, 135))
    x = pae_np.pose_windows(pose, np.zeros(135), np.ones(135))
    assert x.shape == (5, 135, 240)
    assert np.all(x[:, :, 0] == 0)
    vel = (pose[1:] - pose[:-1]).astype(np.float32)
    # window 0: frame differences start at position 121 (120 zero-padding rows + the zero frame)
    assert np.all(x[0, :, 1:121] == 0) and np.array_equal(x[0, :, 121:125].T, vel)
    assert np.array_equal(x[4, :, 117:121].T, vel)


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
def test_oracle_matches_reference_in_place():
    import torch

    m = rh.import_pae()
    try:
        sd = pae_np.random_state_dict(11)
        net = m.Model(input_channels=135, embedding_channels=8, time_range=240, key_range=13, window=4.0).eval()
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        pose, mean, std, _ = mg.inputs(11, 5)
        with torch.no_grad(), rh._quiet():
            ref = m.pose2phase(net, pose, mean, std)
        got = pae_np.pose2phase(sd, pose, mean, std)
        assert cyc(got[:, 0], ref[:, 0]).max() < TOL_PHASE
        assert np.abs(got[:, 1] - ref[:, 1]).max() < TOL_FREQ
        assert np.abs(got[:, 2:] - ref[:, 2:]).max() < TOL_AMP
    finally:
        rh.release_pae()

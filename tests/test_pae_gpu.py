"""Periodic auto-encoder on the device (qpgesture_b200/PAE.py over csrc/pae.cu) against the oracle and the golden
vectors of the unmodified reference.  Floating-point path: tolerances stated below (the reference computes in float32;
the device path accumulates the shared first convolution in float64)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests", "golden"))

TOL_LATENT = 3e-5
TOL_PHASE = 3e-5
TOL_FREQ = 3e-4
TOL_AMP = 3e-6


def cyc(a, b):
    return np.abs((np.asarray(a, dtype=np.float64) - b + 0.5) % 1.0 - 0.5)


def model(seed):
    import torch
    from oracle import pae_np
    from qpgesture_b200 import PAE

    sd = pae_np.random_state_dict(seed)
    net = PAE.Model(input_channels=135, embedding_channels=8, time_range=240, key_range=13, window=4.0,
                    device="cuda:0").eval()
    net.load_state_dict({"module." + k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    return net, sd


@pytest.mark.parametrize("name", ["pae_s0", "pae_s1"])
def test_pose2phase_and_forward_match_golden(name):
    import make_golden_pae as mg
    from qpgesture_b200 import PAE

    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    net, _ = model(int(g["seed"]))
    pose, mean, std, xw = mg.inputs(int(g["seed"]), int(g["T"]))
    got = PAE.pose2phase(net, pose, mean, std)
    ref = g["phase"]
    assert got.shape == ref.shape and got.dtype == np.float32
    assert cyc(got[:, 0], ref[:, 0]).max() < TOL_PHASE
    assert np.abs(got[:, 1] - ref[:, 1]).max() < TOL_FREQ
    assert np.abs(got[:, 2:] - ref[:, 2:]).max() < TOL_AMP
    y, latent, signal, params = net(xw)
    assert tuple(y.shape) == (2, 135 * 240) and tuple(params[0].shape) == (2, 8, 1)
    assert np.abs(latent.cpu().numpy() - g["latent"]).max() < TOL_LATENT
    assert np.abs(signal.cpu().numpy() - g["signal"]).max() < 5e-4
    assert np.abs(y.cpu().numpy()[:, :2048] - g["y_head"]).max() < 5e-4
    assert np.abs(y.cpu().numpy().astype(np.float64).sum(1) - g["y_sum"]).max() < 5e-2


@pytest.mark.parametrize("T", [1, 2, 3, 121, 333])
def test_shared_first_convolution_equals_per_window_convolution(T):
    """qpg_pae_sliding_conv1 (prefix sums over shared diagonals) against the plain convolution of every
    materialised window, and pose2phase against the oracle, for sequences shorter and longer than the window."""
    import torch
    from oracle import pae_np
    from qpgesture_b200 import PAE

    net, sd = model(21)
    rng = np.random.default_rng(T)
    pose = np.cumsum(rng.standard_normal((T, 135)) * 0.5, axis=0)
    mean, std = rng.standard_normal(135) * 0.1, rng.uniform(0.3, 1.5, 135)
    x = pae_np.pose_windows(pose, mean, std)                               # [T, 135, 240]
    xd = torch.from_numpy(x).float().cuda()
    h_direct = net._conv(xd, "conv1", 120, True)
    vel_pad = torch.zeros((T + 238, 135), dtype=torch.float32, device="cuda")
    norm = (pose - mean) / std
    vel_pad[120:120 + T - 1] = torch.from_numpy((norm[1:] - norm[:-1]).astype(np.float32)).cuda()
    params, latent = net.phases_of_sequence(vel_pad)
    h1 = torch.empty((T, 15, 241), dtype=torch.float32, device="cuda")
    from qpgesture_b200 import _lib
    w = net._w
    _lib.check(_lib.load().qpg_pae_sliding_conv1(_lib.ptr(vel_pad), _lib.ptr(w["conv1.w"]), _lib.ptr(w["conv1.scale"]),
                                                 _lib.ptr(w["conv1.shift"]), T, 135, 15, 240, _lib.ptr(h1),
                                                 _lib.stream_ptr()), "sliding")
    torch.cuda.synchronize()
    # the per-window arm accumulates 32 400 products in float32, the shared arm in float64
    assert float((h1 - h_direct).abs().max()) < 5e-5
    if T <= 3:
        want = pae_np.pose2phase(sd, pose, mean, std)
        got = PAE.pose2phase(net, pose, mean, std)
        assert cyc(got[:, 0], want[:, 0]).max() < TOL_PHASE
        assert np.abs(got[:, 1] - want[:, 1]).max() < TOL_FREQ
        assert np.abs(got[:, 2:] - want[:, 2:]).max() < TOL_AMP
    else:
        lat_direct, par_direct = net.embed(xd)
        assert float((latent - lat_direct).abs().max()) < TOL_LATENT
        assert cyc(params[:, 0].cpu().numpy(), par_direct[:, 0].cpu().numpy()).max() < TOL_PHASE


def test_constant_latent_gives_zero_spectrum_and_nan_frequency():
    """A constant latent channel has no power outside the DC bin: amplitude exactly 0 and frequency 0/0 = NaN, which
    is what the reference's expression sum(freqs * power) / sum(power) (PAE.py:106) yields for zero power (an FFT
    library may leave 1e-16-sized residues there instead; the device kernel centres the row first, so it is exact)."""
    import torch
    from oracle import pae_np

    net, sd = model(22)
    sd = dict(sd)
    sd["conv2.weight"] = np.zeros_like(sd["conv2.weight"])           # latent = tanh(folded bias): constant in time
    net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    x = np.random.default_rng(0).standard_normal((2, 135 * 240)).astype(np.float32)
    _, latent, _, (p, f, a, b) = net(x)
    want = pae_np.forward(sd, x)
    assert float((latent - latent[:, :, :1]).abs().max()) == 0.0
    assert torch.isnan(f).all()
    assert float(a.abs().max()) == 0.0
    assert np.abs(b.cpu().numpy() - want[3][3]).max() < 1e-6
    assert np.abs((p.cpu().numpy() - want[3][0] + 0.5) % 1.0 - 0.5).max() < TOL_PHASE


def test_argument_errors():
    import torch
    from qpgesture_b200 import PAE, _lib

    net = PAE.Model(device="cuda:0")
    with pytest.raises(RuntimeError):
        net(np.zeros((1, 135 * 240), dtype=np.float32))
    with pytest.raises(KeyError):
        net.load_state_dict({"conv1.weight": torch.zeros(15, 135, 240)})
    x = torch.zeros((1, 4, 16), device="cuda")
    w = torch.zeros((2, 4, 6), device="cuda")          # kernel width not a multiple of 4
    s = torch.zeros(2, device="cuda")
    o = torch.zeros((1, 2, 16), device="cuda")
    rc = _lib.load().qpg_pae_conv1d(_lib.ptr(x), _lib.ptr(w), _lib.ptr(s), _lib.ptr(s), 1, 4, 16, 2, 6, 2, 0, _lib.ptr(o),
                                    _lib.stream_ptr())
    assert rc != 0

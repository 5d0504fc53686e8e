"""CPU restatement of the reference's periodic auto-encoder inference path (TEST INFRASTRUCTURE).

NumPy float64 restatement of `Model.forward` in eval mode (codebook/PAE.py:116-162: conv k=240 -> BatchNorm ->
tanh twice, rFFT-derived frequency / amplitude / offset :99-114, per-channel Linear(240, 2) -> BatchNorm ->
atan2 phase :132-136 with the model's own atan2 :92-97, sinusoidal latent reconstruction :146, two more
convolutions :150-159) and of `pose2phase` (:477-508: normalise, frame differences, zero padding 120 / 119, one
240-frame window per frame whose first frame is zero).  State dict in the reference's own key layout.

Floating-point path: parity is a tolerance (stated in tests/test_pae_*.py), not bit equality.  Pinned against the
reference module itself (tests/test_pae_cpu.py imports it in place through oracle/ref_harness.py in the build
container) and against tests/golden/pae_*.npz produced by tests/golden/make_golden_pae.py.
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.
"""
from __future__ import annotations

import numpy as np

EPS_BN = 1e-5                       # nn.BatchNorm1d default (PAE.py:71, 74, 80, 85)
FRAMES, JOINT_CH, PHASE_CH, KEYS, WINDOW = 240, 135, 8, 13, 4.0      # PAE.py:26-34


def random_state_dict(seed=0, input_channels=JOINT_CH, embedding_channels=PHASE_CH, time_range=FRAMES):
    """Seeded weights in the reference's key layout (there is no PAE checkpoint in the repository).  Conv weights
    are scaled so that the tanh layers are neither saturated nor linear; BatchNorm running statistics are
    non-trivial so that eval-mode folding is exercised."""
    rng = np.random.default_rng(seed)
    mid = input_channels // 9
    sd = {}

    def conv(name, co, ci, gain):
        sd[name + ".weight"] = (rng.standard_normal((co, ci, time_range)) * gain / np.sqrt(ci * time_range)).astype(np.float32)
        sd[name + ".bias"] = (rng.standard_normal(co) * 0.1).astype(np.float32)

    def bn(name, c):
        sd[name + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[name + ".running_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
        sd[name + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
        sd[name + ".num_batches_tracked"] = np.array(7, dtype=np.int64)

    conv("conv1", mid, input_channels, 1.5)
    bn("bn_conv1", mid)
    conv("conv2", embedding_channels, mid, 2.5)
    bn("bn_conv2", embedding_channels)
    for i in range(embedding_channels):
        sd[f"fc.{i}.weight"] = (rng.standard_normal((2, time_range)) / np.sqrt(time_range)).astype(np.float32)
        sd[f"fc.{i}.bias"] = (rng.standard_normal(2) * 0.1).astype(np.float32)
        bn(f"bn.{i}", 2)
    conv("deconv1", mid, embedding_channels, 2.0)
    bn("bn_deconv1", mid)
    conv("deconv2", input_channels, mid, 1.0)
    sd["tpi"] = np.array([2.0 * np.pi], dtype=np.float32)
    sd["args"] = np.linspace(-WINDOW / 2, WINDOW / 2, time_range, dtype=np.float32)
    k = np.arange(1, time_range // 2 + 1, dtype=np.float32) / np.float32(time_range)       # rfftfreq(n)[1:]
    sd["freqs"] = (k * np.float32(time_range * (KEYS / time_range)) / np.float32(WINDOW)).astype(np.float32)
    return sd


def _conv1d(x, w, b, pad):
    """x [B, Ci, L], w [Co, Ci, K] -> [B, Co, L + 2 pad - K + 1] (cross-correlation, zeros padding)."""
    B, Ci, L = x.shape
    Co, _, K = w.shape
    xp = np.zeros((B, Ci, L + 2 * pad))
    xp[:, :, pad:pad + L] = x
    Lo = L + 2 * pad - K + 1
    out = np.empty((B, Co, Lo))
    w2 = w.reshape(Co, Ci * K).astype(np.float64)
    for n in range(B):
        win = np.lib.stride_tricks.sliding_window_view(xp[n], K, axis=1)       # [Ci, Lo, K]
        out[n] = w2 @ win.transpose(0, 2, 1).reshape(Ci * K, Lo)
    return out + np.asarray(b, dtype=np.float64)[None, :, None]


def _bn(x, sd, name, axis=1):
    shape = [1] * x.ndim
    shape[axis] = -1
    g = lambda k: np.asarray(sd[f"{name}.{k}"], dtype=np.float64).reshape(shape)
    return (x - g("running_mean")) / np.sqrt(g("running_var") + EPS_BN) * g("weight") + g("bias")


def _atan2(y, x):
    """PAE.py:92-97: atan(y / x) moved by half a turn in the left half plane (no special case at x == 0)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        ans = np.arctan(y / x)
    ans = np.where((x < 0) & (y >= 0), ans + np.pi, ans)
    return np.where((x < 0) & (y < 0), ans - np.pi, ans)


def forward(sd, x, time_range=FRAMES):
    """Model.forward (PAE.py:116-162) in eval mode.  x [B, C*T] or [B, C, T].
    Returns (y [B, C*T], latent [B, E, T], signal [B, E, T], (p, f, a, b) each [B, E, 1])."""
    sd = {k: np.asarray(v) for k, v in sd.items()}
    C = sd["conv1.weight"].shape[1]
    E = sd["conv2.weight"].shape[0]
    T = time_range
    y = np.asarray(x, dtype=np.float64).reshape(-1, C, T)                                  # :119
    y = np.tanh(_bn(_conv1d(y, sd["conv1.weight"], sd["conv1.bias"], T // 2), sd, "bn_conv1"))            # :120-122
    y = np.tanh(_bn(_conv1d(y, sd["conv2.weight"], sd["conv2.bias"], (T - 1) // 2), sd, "bn_conv2"))      # :124-126
    latent = y
    rfft = np.fft.rfft(y, axis=2)                                                          # :100
    power = np.abs(rfft[:, :, 1:]) ** 2                                                    # :101-103
    time_scale = KEYS / T
    freqs = sd["freqs"].astype(np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        f = (freqs * power).sum(2) / power.sum(2) / time_scale                             # :106-107
    a = 2 * np.sqrt(power.sum(2)) / T                                                      # :110
    b = rfft.real[:, :, 0] / T                                                             # :113
    p = np.empty((y.shape[0], E))
    for i in range(E):                                                                     # :133-136
        v = y[:, i, :] @ sd[f"fc.{i}.weight"].astype(np.float64).T + sd[f"fc.{i}.bias"]
        v = _bn(v, sd, f"bn.{i}")
        p[:, i] = _atan2(v[:, 1], v[:, 0]) / (2 * np.pi)
    p, f, a, b = (t[:, :, None] for t in (p, f, a, b))                                     # :139-142
    signal = a * np.sin(2 * np.pi * (f * sd["args"].astype(np.float64) + p)) + b           # :146
    h = np.tanh(_bn(_conv1d(signal, sd["deconv1.weight"], sd["deconv1.bias"], (T - 1) // 2), sd, "bn_deconv1"))
    out = _conv1d(h, sd["deconv2.weight"], sd["deconv2.bias"], T // 2)                     # :155
    return out.reshape(out.shape[0], -1), latent, signal, (p, f, a, b)


def pose_windows(pose, data_mean, std, n_poses=FRAMES):
    """The network inputs pose2phase builds (PAE.py:480-500): [T, C, 240] with frame 0 of every window zero."""
    pose = (np.asarray(pose, dtype=np.float64) - data_mean) / std
    vel = pose[1:] - pose[:-1]
    vel = np.pad(vel, ((n_poses // 2, n_poses // 2 - 1), (0, 0)))
    T = pose.shape[0]
    out = np.zeros((T, pose.shape[1], n_poses))
    for i in range(T):
        out[i, :, 1:] = vel[i:i + n_poses - 1].astype(np.float32).T     # .float() at :497
    return out


def pose2phase(sd, pose, data_mean, std):
    """PAE.py:477-508 -> float array [T, 4, 1, E, 1] (p, f, a, b per frame)."""
    x = pose_windows(pose, data_mean, std)
    _, _, _, params = forward(sd, x)
    return np.stack([np.stack(params, axis=1)[i][:, None] for i in range(x.shape[0])]).astype(np.float32)

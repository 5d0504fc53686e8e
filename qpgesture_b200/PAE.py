"""Periodic auto-encoder inference on the device: the producer of the match database's `phase` column.

Mirrors the reference's call surface for this step (codebook/PAE.py): `Model(input_channels, embedding_channels,
time_range, key_range, window)` with `load_state_dict` in the reference's key layout and an eval-mode `forward(x)
-> (y, latent, signal, [p, f, a, b])` (:116-162), and `pose2phase(network, pose, data_mean, std)` (:477-508) with
the same return value (float32 array [T, 4, 1, E, 1]).  Training (PAE.py:273-475) is out of scope.

`pose2phase` does not run the network once per frame: all T windows of a sequence are shifted views of one padded
velocity sequence, so the first convolution is computed once per "diagonal" with prefix sums over the kernel taps
(csrc/pae.cu, `qpg_pae_sliding_conv1`), followed by one batched second convolution and one parameter kernel.  Three
launches per sequence instead of T network calls.  There is no CPU fallback: the CUDA library must be present.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _lib

# PAE.py:25-34
window = 4.0
frames = 240
keys = 13
joints = 15
number_of_channels = 9
input_channels = number_of_channels * joints
phase_channels = 8

_EPS_BN = 1e-5


def _fold(weight_bias, bn, prefix_bn, sd):
    """(scale, shift) with y = scale * conv_without_bias + shift == bn(conv + bias) in eval mode."""
    bias = sd[weight_bias].double()
    if not bn:
        return torch.ones_like(bias).float(), bias.float()
    g, b = sd[prefix_bn + ".weight"].double(), sd[prefix_bn + ".bias"].double()
    rm, rv = sd[prefix_bn + ".running_mean"].double(), sd[prefix_bn + ".running_var"].double()
    scale = g / torch.sqrt(rv + _EPS_BN)
    return scale.float(), ((bias - rm) * scale + b).float()


class Model:
    """Inference-only counterpart of PAE.py:50-162 (always eval mode: BatchNorm uses its running statistics)."""

    def __init__(self, input_channels=input_channels, embedding_channels=phase_channels, time_range=frames,
                 key_range=keys, window=window, device=None):
        self.input_channels = input_channels
        self.embedding_channels = embedding_channels
        self.time_range = time_range
        self.key_range = key_range
        self.window = window
        self.time_scale = key_range / time_range
        self.device = torch.device(device if device is not None else "cuda")
        self.intermediate_channels = int(input_channels / number_of_channels)
        # buffers the reference keeps as frozen Parameters (PAE.py:61-65); overwritten by load_state_dict
        self.tpi = torch.tensor([2.0 * math.pi], dtype=torch.float32)
        self.args = torch.from_numpy(np.linspace(-window / 2, window / 2, time_range, dtype=np.float32))
        self.freqs = torch.fft.rfftfreq(time_range)[1:] * (time_range * self.time_scale) / window
        self._w = None

    # -- nn.Module look-alikes used at the reference's call sites (PAE.py:534-538) --
    def to(self, device):
        self.device = torch.device(device)
        if self._w is not None:
            self._w = {k: v.to(self.device) for k, v in self._w.items()}
        return self

    def eval(self):
        return self

    def load_state_dict(self, sd, strict=True):
        sd = {(k[7:] if k.startswith("module.") else k): torch.as_tensor(np.asarray(v) if not torch.is_tensor(v) else v)
              for k, v in sd.items()}
        E, T, C, O = self.embedding_channels, self.time_range, self.input_channels, self.intermediate_channels
        need = ["conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "deconv1.weight", "deconv1.bias",
                "deconv2.weight", "deconv2.bias"]
        need += [f"{b}.{k}" for b in ("bn_conv1", "bn_conv2", "bn_deconv1")
                 for k in ("weight", "bias", "running_mean", "running_var")]
        need += [f"fc.{i}.{k}" for i in range(E) for k in ("weight", "bias")]
        need += [f"bn.{i}.{k}" for i in range(E) for k in ("weight", "bias", "running_mean", "running_var")]
        missing = [k for k in need if k not in sd]
        if missing and strict:
            raise KeyError(f"missing keys in PAE state dict: {missing[:5]}{' ...' if len(missing) > 5 else ''}")
        shapes = {"conv1.weight": (O, C, T), "conv2.weight": (E, O, T), "deconv1.weight": (O, E, T),
                  "deconv2.weight": (C, O, T)}
        for k, shp in shapes.items():
            if tuple(sd[k].shape) != shp:
                raise ValueError(f"{k}: expected {shp}, got {tuple(sd[k].shape)}")
        w = {}
        for name, bn in (("conv1", "bn_conv1"), ("conv2", "bn_conv2"), ("deconv1", "bn_deconv1"), ("deconv2", None)):
            w[name + ".w"] = sd[name + ".weight"].float().contiguous()
            w[name + ".scale"], w[name + ".shift"] = _fold(name + ".bias", bn is not None, bn, sd)
        w["fc.w"] = torch.stack([sd[f"fc.{i}.weight"].float() for i in range(E)]).contiguous()        # [E, 2, T]
        sc, sh = zip(*[_fold(f"fc.{i}.bias", True, f"bn.{i}", sd) for i in range(E)])
        w["fc.scale"], w["fc.shift"] = torch.stack(sc).contiguous(), torch.stack(sh).contiguous()    # [E, 2]
        for k in ("tpi", "args", "freqs"):
            if k in sd:
                setattr(self, k, sd[k].float())
        w["freqs"] = self.freqs.float().contiguous()
        w["args"] = self.args.float().contiguous()
        self._w = {k: v.to(self.device) for k, v in w.items()}
        return self

    # -- kernels --
    def _weights(self):
        if self._w is None:
            raise RuntimeError("load_state_dict first: the inference model has no initialiser of its own")
        return self._w

    def _conv(self, x, name, pad, act):
        w = self._weights()
        B, Ci, Lin = x.shape
        Co, _, K = w[name + ".w"].shape
        out = torch.empty((B, Co, Lin + 2 * pad - K + 1), dtype=torch.float32, device=x.device)
        lib = _lib.load()
        for b0 in range(0, B, 65535):
            nb = min(65535, B - b0)
            _lib.check(lib.qpg_pae_conv1d(_lib.ptr(x[b0:b0 + nb]), _lib.ptr(w[name + ".w"]), _lib.ptr(w[name + ".scale"]),
                                          _lib.ptr(w[name + ".shift"]), nb, Ci, Lin, Co, K, pad, int(act),
                                          _lib.ptr(out[b0:b0 + nb]), _lib.stream_ptr()), "qpg_pae_conv1d")
        return out

    def _params(self, latent):
        w = self._weights()
        B, E, T = latent.shape
        params = torch.empty((B, 4, E), dtype=torch.float32, device=latent.device)
        _lib.check(_lib.load().qpg_pae_params(_lib.ptr(latent), _lib.ptr(w["fc.w"]), _lib.ptr(w["fc.scale"]),
                                              _lib.ptr(w["fc.shift"]), _lib.ptr(w["freqs"]), float(self.time_scale),
                                              B, E, T, _lib.ptr(params), _lib.stream_ptr()), "qpg_pae_params")
        return params

    def embed(self, x):
        """x [B, C*T] or [B, C, T] -> (latent [B, E, T], params [B, 4, E]); PAE.py:119-136."""
        T = self.time_range
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device).reshape(-1, self.input_channels, T).contiguous()
        h = self._conv(x, "conv1", T // 2, True)
        latent = self._conv(h, "conv2", (T - 1) // 2, True)
        return latent, self._params(latent)

    def forward(self, x):
        """PAE.py:116-162 -> (y [B, C*T], latent, signal, [p, f, a, b] each [B, E, 1])."""
        w = self._weights()
        T = self.time_range
        latent, params = self.embed(x)
        p, f, a, b = (params[:, i, :].unsqueeze(2) for i in range(4))
        tpi = self.tpi.to(latent.device)
        signal = a * torch.sin(tpi * (f * w["args"] + p)) + b                     # :146
        h = self._conv(signal.contiguous(), "deconv1", (T - 1) // 2, True)
        y = self._conv(h, "deconv2", T // 2, False)
        return y.reshape(y.shape[0], self.input_channels * T), latent, signal, [p, f, a, b]

    __call__ = forward

    def phases_of_sequence(self, vel_pad):
        """vel_pad [T + 238, C] float32 on the device -> params [T, 4, E] for the T windows of pose2phase."""
        w = self._weights()
        K, O, C = self.time_range, self.intermediate_channels, self.input_channels
        T = vel_pad.shape[0] - (K - 2)
        h1 = torch.empty((T, O, K + 1), dtype=torch.float32, device=vel_pad.device)
        _lib.check(_lib.load().qpg_pae_sliding_conv1(_lib.ptr(vel_pad), _lib.ptr(w["conv1.w"]),
                                                     _lib.ptr(w["conv1.scale"]), _lib.ptr(w["conv1.shift"]), T, C, O, K,
                                                     _lib.ptr(h1), _lib.stream_ptr()), "qpg_pae_sliding_conv1")
        latent = self._conv(h1, "conv2", (K - 1) // 2, True)
        return self._params(latent), latent


def pose2phase(network, pose, data_mean, std, as_device_tensor=False):
    """PAE.py:477-508: per-frame (p, f, a, b) of a pose sequence [T, C].  Returns what the reference returns,
    `np.array(result)` of shape [T, 4, 1, E, 1] float32 (or the device tensor [T, 4, E] on request)."""
    n_poses = network.time_range
    dev = network.device
    pose = torch.as_tensor(np.asarray(pose, dtype=np.float64), device=dev)
    pose = (pose - torch.as_tensor(np.asarray(data_mean, dtype=np.float64), device=dev)) / \
        torch.as_tensor(np.asarray(std, dtype=np.float64), device=dev)                     # :480
    T = pose.shape[0]
    vel_pad = torch.zeros((T + n_poses - 2, pose.shape[1]), dtype=torch.float32, device=dev)
    vel_pad[n_poses // 2:n_poses // 2 + T - 1] = (pose[1:] - pose[:-1]).float()            # :481-482, .float() at :497
    params, _ = network.phases_of_sequence(vel_pad)
    if as_device_tensor:
        return params
    return params.cpu().numpy().reshape(T, 4, 1, network.embedding_channels, 1)

"""Synthetic databases in the reference's on-disk npz formats.

There is no BEAT data, no checkpoint and no network, so every test and the
bench run on seeded synthetic tensors of the named shapes.  File formats and
keys follow codebook/Speech2GestureMatching/data_processing.py:200-206,255,
278,339-343 and GestureKNN.py:476,483 so that both the reference (imported in
place by oracle/ref_harness.py) and this package read the same files.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .constant import codebook_size, num_frames, num_frames_code, NUM_JOINTS, WAVVQ_FRAMES


@dataclass
class SynthPaths:
    train_database: str
    test_data: str
    train_codebook: str
    codebook_signature: str
    train_wavlm: str
    test_wavlm: str
    train_wavvq: str
    test_wavvq: str

    def as_argv(self, out_knn_filename: str, max_frames: int = 0):
        """The 11 flags GestureKNN.sh:7-18 passes."""
        return [
            f"--train_database={self.train_database}",
            f"--test_data={self.test_data}",
            f"--out_knn_filename={out_knn_filename}",
            "--out_video_path=./output/output_video_folder/",
            f"--train_codebook={self.train_codebook}",
            f"--codebook_signature={self.codebook_signature}",
            f"--train_wavlm={self.train_wavlm}",
            f"--test_wavlm={self.test_wavlm}",
            f"--train_wavvq={self.train_wavvq}",
            f"--test_wavvq={self.test_wavvq}",
            f"--max_frames={max_frames}",
        ]


def make_arrays(n_train: int, n_test: int, seed: int = 0, wavlm_dim: int = 1024,
                ctx_dim: int = 384, wavlm_frames: int = 199, n_codes_used: int = codebook_size,
                mfcc_dim: int = 14):
    """Seeded arrays of the reference shapes.  `phase` is returned as a plain
    float32 array [n, 240, 4, 8] (p, f, a, b channels); `phase_to_object`
    turns it into the object array of (1,8,1) torch tensors the reference
    pickles (data_processing.py:339)."""
    rng = np.random.default_rng(seed)

    def split(n):
        return dict(
            mfcc=rng.standard_normal((n, num_frames, mfcc_dim)).astype(np.float32),
            energy=rng.standard_normal((n, num_frames)).astype(np.float32),
            pitch=rng.standard_normal((n, num_frames)).astype(np.float32),
            volume=rng.standard_normal((n, num_frames)).astype(np.float32),
            phase=rng.standard_normal((n, num_frames, 4, 8)).astype(np.float32),
            context=rng.standard_normal((n, num_frames_code, 1, ctx_dim)).astype(np.float32),
            wavlm=rng.standard_normal((n, wavlm_frames, wavlm_dim)).astype(np.float32),
            wavvq=rng.integers(0, 320, size=(n, WAVVQ_FRAMES, 2)).astype(np.int64),
        )

    train = split(n_train)
    test = split(n_test)
    code = rng.integers(0, n_codes_used, size=(n_train, num_frames_code)).astype(np.int64)
    signature = rng.standard_normal((codebook_size, NUM_JOINTS)).astype(np.float32)
    return train, test, code, signature


def phase_to_object(phase: np.ndarray) -> np.ndarray:
    """float32 [n, 240, 4, 8] -> object [n, 240, 4] of torch tensors (1, 8, 1)."""
    import torch

    n, t, c, ch = phase.shape
    out = np.empty((n, t, c), dtype=object)
    tens = torch.from_numpy(np.ascontiguousarray(phase))
    for i in range(n):
        for j in range(t):
            for k in range(c):
                out[i, j, k] = tens[i, j, k].reshape(1, ch, 1).clone()
    return out


def write_npz_set(root: str, train, test, code, signature, object_phase: bool = True) -> SynthPaths:
    """Write the 8 files GestureKNN.sh names.  With object_phase=False the
    phase column is stored as a dense float32 array (this package reads both;
    the reference needs the object form)."""
    os.makedirs(root, exist_ok=True)
    p = SynthPaths(
        train_database=os.path.join(root, "train_240_txt_2.npz"),
        test_data=os.path.join(root, "test_240_txt_2.npz"),
        train_codebook=os.path.join(root, "train_240_code.npz"),
        codebook_signature=os.path.join(root, "code.npz"),
        train_wavlm=os.path.join(root, "train_240_WavLM.npz"),
        test_wavlm=os.path.join(root, "test_240_WavLM.npz"),
        train_wavvq=os.path.join(root, "train_240_WavVQ.npz"),
        test_wavvq=os.path.join(root, "wavvq_240.npz"),
    )
    for split, path in ((train, p.train_database), (test, p.test_data)):
        ph = phase_to_object(split["phase"]) if object_phase else split["phase"]
        np.savez(path, mfcc=split["mfcc"], energy=split["energy"], pitch=split["pitch"],
                 volume=split["volume"], phase=ph, context=split["context"])
    np.savez(p.train_codebook, code=code)
    np.savez(p.codebook_signature, signature=signature)
    np.savez(p.train_wavlm, wavlm=train["wavlm"])
    np.savez(p.test_wavlm, wavlm=test["wavlm"])
    np.savez(p.train_wavvq, wavvq=train["wavvq"])
    np.savez(p.test_wavvq, wavvq=test["wavvq"])
    return p


# --------------------------------------------------------------------------------------------
# VQ-VAE: there is no checkpoint to load, so tests and the bench use seeded random-init weights
# in the reference's state-dict layout (codebook/models/*.py; keys as listed in DESIGN.md).
# --------------------------------------------------------------------------------------------
VQVAE_HPS = dict(levels=1, downs_t=[3], strides_t=[2], emb_width=512, l_bins=512, l_mu=0.99, commit=0.02,
                 hvqvae_multipliers=[1], width=512, depth=3, m_conv=1.0, dilation_growth_rate=3,
                 sample_length=30, use_bottleneck=True, joint_channel=9, vel=1, acc=1,
                 vqvae_reverse_decoder_dilation=True)    # configs/codebook.yml:1-25


def vqvae_hps(**over):
    from types import SimpleNamespace
    d = dict(VQVAE_HPS)
    d.update(over)
    return SimpleNamespace(**d)


import torch  # noqa: E402


def random_vqvae_state_dict(hps, input_dim=135, seed=0, codebook_seed=1, codebook_scale=0.12):
    """Deterministic weights in the reference's state-dict layout (there is no
    checkpoint to load): Conv default-style uniform(-1/sqrt(fan_in), ..) and a
    N(0, codebook_scale^2) codebook (SURVEY.md 8(d) config 2 uses N(0,1); scaled here, see below)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, transposed=False):
        bound = 1.0 / (cin * k) ** 0.5
        shape = (cin, cout, k) if transposed else (cout, cin, k)
        sd[name + ".weight"] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        sd[name + ".bias"] = (torch.rand((cout,), generator=g) * 2 - 1) * bound

    w, e, down_t = hps.width, hps.emb_width, hps.downs_t[0]
    pre = "encoders.0.level_blocks.0.model"
    for i in range(down_t):
        conv(f"{pre}.{i}.0", w, input_dim if i == 0 else w, hps.strides_t[0] * 2)
        for d in range(hps.depth):
            conv(f"{pre}.{i}.1.model.{d}.model.1", w, w, 3)
            conv(f"{pre}.{i}.1.model.{d}.model.3", w, w, 1)
    conv(f"{pre}.{down_t}", e, w, 3)
    pre = "decoders.0.level_blocks.0.model"
    conv(f"{pre}.0", w, e, 3)
    for i in range(down_t):
        for d in range(hps.depth):
            conv(f"{pre}.{i + 1}.0.model.{d}.model.1", w, w, 3)
            conv(f"{pre}.{i + 1}.0.model.{d}.model.3", w, w, 1)
        conv(f"{pre}.{i + 1}.1", e if i == down_t - 1 else w, w, hps.strides_t[0] * 2, transposed=True)
    conv("decoders.0.out", input_dim, e, 3)
    # scaled to the magnitude of the random-weight encoder's latents so that the arg-min is contested
    sd["bottleneck.level_blocks.0.k"] = codebook_scale * torch.randn(
        (hps.l_bins, e), generator=torch.Generator().manual_seed(codebook_seed))
    return sd


def random_pae_state_dict(seed=0, input_channels=135, embedding_channels=8, time_range=240):
    """Deterministic periodic auto-encoder weights in the reference's state-dict layout (PAE.py:50-90; the
    repository ships no checkpoint): convolutions scaled so that the tanh layers are not saturated, BatchNorm with
    non-trivial running statistics."""
    g = torch.Generator().manual_seed(seed)
    mid = input_channels // 9
    sd = {}

    def conv(name, co, ci, gain):
        sd[name + ".weight"] = torch.randn((co, ci, time_range), generator=g) * gain / (ci * time_range) ** 0.5
        sd[name + ".bias"] = torch.randn((co,), generator=g) * 0.1

    def bn(name, c):
        sd[name + ".weight"] = torch.rand((c,), generator=g) + 0.5
        sd[name + ".bias"] = torch.randn((c,), generator=g) * 0.1
        sd[name + ".running_mean"] = torch.randn((c,), generator=g) * 0.1
        sd[name + ".running_var"] = torch.rand((c,), generator=g) + 0.5

    conv("conv1", mid, input_channels, 1.5)
    bn("bn_conv1", mid)
    conv("conv2", embedding_channels, mid, 2.5)
    bn("bn_conv2", embedding_channels)
    for i in range(embedding_channels):
        sd[f"fc.{i}.weight"] = torch.randn((2, time_range), generator=g) / time_range ** 0.5
        sd[f"fc.{i}.bias"] = torch.randn((2,), generator=g) * 0.1
        bn(f"bn.{i}", 2)
    conv("deconv1", mid, embedding_channels, 2.0)
    bn("bn_deconv1", mid)
    conv("deconv2", input_channels, mid, 1.0)
    return sd

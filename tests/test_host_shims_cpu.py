"""Host-side logic that needs no GPU: the convolution tile-width model, the Savitzky-Golay tables, the PAE BatchNorm
folding, the legacy matcher's tie policy argument."""
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_conv_tile_width_model(monkeypatch):
    """narrow tiles only when wide ones would leave more than half of the 148 SMs idle"""
    from qpgesture_b200 import vqvae

    monkeypatch.setattr(vqvae, "_sm_count", lambda device: 148)
    conv = vqvae._TcConv.__new__(vqvae._TcConv)
    conv.BN, conv.N_pad = 256, 512
    pick = lambda B, n_out: conv._tile_n(B, n_out, "cuda:0")
    assert pick(4096, 4) == 256          # 16384 rows: 256 tiles of 256 columns = two full waves
    assert pick(4096, 2) == 256          # 8192 rows: 128 wide tiles, one wave; 256 narrow ones would need two
    assert pick(4096, 1) == 128          # 4096 rows: 64 wide tiles leave 84 SMs idle -> 128 narrow tiles
    assert pick(64, 30) in (64, 128)     # 1920 rows: 30 wide tiles -> narrower
    conv.BN, conv.N_pad = 128, 512       # 3xTF32: at most 128 columns
    assert conv._tile_n(4096, 4, "cuda:0") == 128
    conv.BN, conv.N_pad = 144, 144       # ragged output width (135 channels): single tile, untouched
    assert conv._tile_n(4096, 8, "cuda:0") == 144


def test_pae_batchnorm_folding_matches_eval_mode():
    """y = scale * conv_without_bias + shift  ==  BatchNorm_eval(conv + bias)"""
    from qpgesture_b200.PAE import _fold

    g = torch.Generator().manual_seed(0)
    sd = {"c.bias": torch.randn(5, generator=g), "bn.weight": torch.rand(5, generator=g) + 0.5,
          "bn.bias": torch.randn(5, generator=g), "bn.running_mean": torch.randn(5, generator=g),
          "bn.running_var": torch.rand(5, generator=g) + 0.5}
    scale, shift = _fold("c.bias", True, "bn", sd)
    z = torch.randn(7, 5, generator=g)
    want = torch.nn.functional.batch_norm(z + sd["c.bias"], sd["bn.running_mean"], sd["bn.running_var"], sd["bn.weight"],
                                          sd["bn.bias"], training=False, eps=1e-5)
    assert torch.allclose(scale * z + shift, want, rtol=0, atol=1e-6)
    scale, shift = _fold("c.bias", False, None, sd)
    assert torch.equal(scale, torch.ones(5)) and torch.equal(shift, sd["c.bias"])


def test_pae_model_rejects_wrong_shapes_and_missing_keys():
    from qpgesture_b200.PAE import Model
    from qpgesture_b200.synth import random_pae_state_dict

    sd = random_pae_state_dict(0)
    m = Model(device="cpu")
    m.load_state_dict(sd)                                    # host-side packing works without a GPU
    assert m._w["fc.w"].shape == (8, 2, 240) and m._w["conv1.scale"].shape == (15,)
    bad = dict(sd)
    bad["conv1.weight"] = torch.zeros(15, 135, 100)
    with pytest.raises(ValueError):
        Model(device="cpu").load_state_dict(bad)
    del bad["conv2.bias"]
    with pytest.raises(KeyError):
        Model(device="cpu").load_state_dict(bad)


def test_savgol_edge_rows_are_exact_on_quadratics():
    """a quadratic is reproduced exactly by an order-2 filter, at the edges too"""
    from qpgesture_b200.process_bvh import _savgol_tables

    coef, first, last = _savgol_tables()
    t = np.arange(15.0)
    x = 0.3 * t * t - 2.0 * t + 1.0
    assert np.allclose(first @ x, x[:7], atol=1e-10) and np.allclose(last @ x, x[8:], atol=1e-10)
    assert abs(coef @ x - x[7]) < 1e-10 and abs(coef.sum() - 1.0) < 1e-12

"""Pin oracle/vqvae_ref.py to the reference VQ-VAE modules (no GPU needed)."""
import glob
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import ref_harness as rh
from oracle import vqvae_ref as vr

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "vqvae_*.npz")))


def load_vq_case(path):
    fx = dict(np.load(path))
    over = {k[4:]: int(fx[k]) for k in fx if k.startswith("hps_")}
    hps = vr.make_hps(**over)
    seed = int(fx["seed"])
    sd = vr.random_state_dict(hps, 135, seed=seed, codebook_seed=seed + 100)
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(sd[k].numpy().tobytes())
    assert h.hexdigest() == str(fx["sd_digest"]), "seeded weights differ from the ones the golden vectors used"
    x = torch.randn((int(fx["B"]), int(fx["T"]), 135), generator=torch.Generator().manual_seed(seed + 1000))
    assert hashlib.sha256(x.numpy().tobytes()).hexdigest() == str(fx["x_digest"])
    return fx, hps, sd, x


def test_fixtures_present():
    assert len(GOLDEN) >= 3


@pytest.mark.parametrize("path", GOLDEN)
def test_oracle_vs_golden(path):
    fx, hps, sd, x = load_vq_case(path)
    lat, _ = vr.latents(x, sd, hps)
    # same torch CPU kernels as the reference modules -> expected bit-equal on the same torch build
    assert np.allclose(lat.numpy(), fx["latents"], rtol=0, atol=1e-6)
    codes = vr.encode(x, sd, hps).numpy()
    assert np.array_equal(codes, fx["codes"])
    dec = vr.decode(torch.from_numpy(fx["codes"]), sd, hps).numpy()
    assert np.allclose(dec, fx["decoded"], rtol=0, atol=1e-6)
    _, fit, _ = vr.quantise(torch.from_numpy(fx["latents"]), sd["bottleneck.level_blocks.0.k"])
    assert abs(float(fit) - float(fx["fit"])) < 1e-4 * max(1.0, abs(float(fx["fit"])))


@pytest.mark.skipif(not rh.available(), reason="reference checkout not present")
def test_oracle_vs_live_reference():
    V, B = rh.import_vqvae()
    try:
        hps = vr.make_hps(width=48, emb_width=48, l_bins=96)
        sd = vr.random_state_dict(hps, 135, seed=21, codebook_seed=22)
        model = V.VQVAE(hps, 135).eval()
        model.load_state_dict({"module." + k: v for k, v in sd.items()} if False else sd, strict=True)
        x = torch.randn((2, 48, 135), generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            zs = model.encode(x)
            dec = model.decode(zs)
        assert torch.equal(vr.encode(x, sd, hps), zs[0])
        assert torch.allclose(vr.decode(zs[0], sd, hps), dec, rtol=0, atol=1e-6)
    finally:
        rh.release_vqvae()

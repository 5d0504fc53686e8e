"""Generate tests/golden/vqvae_*.npz from the UNMODIFIED reference VQ-VAE modules
(codebook/models/vqvae.py etc., imported in place; build container only).

Weights are the seeded `oracle.vqvae_ref.random_state_dict` loaded into the
reference's own VQVAE through load_state_dict (there is no checkpoint in the
reference repo).  Recorded: VQVAE.encode / .decode outputs, the quantiser's
distance matrix argmin and `fit`, and the encoder latents, for a small and the
full (codebook.yml) configuration."""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_harness as rh  # noqa: E402
from oracle import vqvae_ref as vr  # noqa: E402

CASES = [
    dict(name="vqvae_small", over=dict(width=32, emb_width=32, l_bins=64), B=3, T=64, seed=5),
    dict(name="vqvae_full", over=dict(), B=2, T=240, seed=6),
    dict(name="vqvae_full_8frame", over=dict(), B=16, T=8, seed=7),
]


def sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(sd[k].numpy().tobytes())
    return h.hexdigest()


def main():
    V, B = rh.import_vqvae()
    for c in CASES:
        hps = vr.make_hps(**c["over"])
        sd = vr.random_state_dict(hps, 135, seed=c["seed"], codebook_seed=c["seed"] + 100)
        model = V.VQVAE(hps, 135).eval()
        missing = model.load_state_dict(sd, strict=True)
        g = torch.Generator().manual_seed(c["seed"] + 1000)
        x = torch.randn((c["B"], c["T"], 135), generator=g)
        with torch.no_grad():
            zs = model.encode(x)
            xin = model.preprocess(x)
            lat = model.encoders[0](xin)[-1]                         # [B, emb, T/8]
            flat, _ = model.bottleneck.level_blocks[0].preprocess(lat)
            x_l, fit = model.bottleneck.level_blocks[0].quantise(flat)
            dec = model.decode(zs)
        rec = dict(codes=zs[0].numpy(), latents=flat.numpy(), fit=float(fit), decoded=dec.numpy(),
                   x_digest=hashlib.sha256(x.numpy().tobytes()).hexdigest(), sd_digest=sd_digest(sd),
                   torch_version=torch.__version__, B=c["B"], T=c["T"], seed=c["seed"])
        for k, v in c["over"].items():
            rec["hps_" + k] = v
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), **rec)
        print(c["name"], zs[0].shape, dec.shape, float(fit))
    rh.release_vqvae()


if __name__ == "__main__":
    main()

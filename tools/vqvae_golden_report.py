#!/usr/bin/env python
"""Index parity of the three VQ-VAE precisions against the reference's recorded codes (tests/golden/vqvae_*.npz):
for every mismatch the reference's own arg-min margin (its second-best minus best float32 distance) is printed.
precision 0 = float32 FFMA, 1 = TF32 tensor cores, 2 = 3xTF32 tensor cores."""
import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vqvae_ref as vr
from tests.test_vqvae_pin import GOLDEN, load_vq_case
from qpgesture_b200.vqvae import VQVAE

for path in GOLDEN:
    fx, hps, sd, x = load_vq_case(path)
    k = sd["bottleneck.level_blocks.0.k"]
    _, _, dist = vr.quantise(torch.from_numpy(fx["latents"]), k)
    d_sorted, _ = torch.sort(dist, dim=1)
    margin = (d_sorted[:, 1] - d_sorted[:, 0]).numpy()
    for prec in (0, 2, 1):
        model = VQVAE(hps, 135, device="cuda", precision=prec).load_state_dict(sd)
        lat = model.latents(x).cpu().numpy().reshape(-1, hps.emb_width)
        codes = model.encode(x)[0].cpu().numpy().reshape(-1)
        want = fx["codes"].reshape(-1)
        bad = np.flatnonzero(codes != want)
        dec = model.decode([torch.from_numpy(fx["codes"])]).cpu().numpy()
        print(json.dumps(dict(case=os.path.basename(path), precision=prec, latents=int(want.size),
                              latent_max_abs_err=float(np.abs(lat - fx["latents"]).max()),
                              latent_scale=float(np.abs(fx["latents"]).max()),
                              index_mismatches=int(bad.size), mismatch_margins=[float(margin[i]) for i in bad[:10]],
                              min_margin_overall=float(margin.min()),
                              decode_max_abs_err=float(np.abs(dec - fx["decoded"]).max()))), flush=True)

#!/usr/bin/env python
"""CPU emulation of TF32 / 3xTF32 operand splitting for the VQ-VAE encoder (round-2 planning).

tf32(x) keeps 10 mantissa bits (here: truncation of the low 13 bits, the conservative case).
  TF32   : conv(tf32(a), tf32(w))
  3xTF32 : conv(a_hi, w_hi) + conv(a_lo, w_hi) + conv(a_hi, w_lo),  x_lo = tf32(x - x_hi)
Products are formed in float32 on the CPU, so this isolates the OPERAND rounding (the tensor core's
accumulator rounding is not modelled).  Prints the code-index agreement with the float32 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from oracle import vqvae_ref as vr


def tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32)


def conv(mode, x, w, b, **kw):
    if mode == "fp32":
        return F.conv1d(x, w, b, **kw)
    xh, wh = tf32(x), tf32(w)
    y = F.conv1d(xh, wh, b, **kw)
    if mode == "3xtf32":
        y = y + F.conv1d(tf32(x - xh), wh, None, **kw) + F.conv1d(xh, tf32(w - wh), None, **kw)
    return y


def encoder(mode, x_nct, sd, hps):
    pre = "encoders.0.level_blocks.0.model"
    x = x_nct
    for i in range(hps.downs_t[0]):
        x = conv(mode, x, sd[f"{pre}.{i}.0.weight"], sd[f"{pre}.{i}.0.bias"], stride=2, padding=1)
        for d in range(hps.depth):
            p = f"{pre}.{i}.1.model.{d}.model"
            dil = hps.dilation_growth_rate ** d
            h = conv(mode, F.relu(x), sd[p + ".1.weight"], sd[p + ".1.bias"], padding=dil, dilation=dil)
            x = x + conv(mode, F.relu(h), sd[p + ".3.weight"], sd[p + ".3.bias"])
    return conv(mode, x, sd[f"{pre}.{hps.downs_t[0]}.weight"], sd[f"{pre}.{hps.downs_t[0]}.bias"], padding=1)


def main():
    hps = vr.make_hps()
    sd = vr.random_state_dict(hps, 135, seed=0, codebook_seed=1)
    x = torch.randn((64, 240, 135), generator=torch.Generator().manual_seed(0))
    k = sd["bottleneck.level_blocks.0.k"]
    codes, lat = {}, {}
    with torch.no_grad():
        for mode in ("fp32", "tf32", "3xtf32"):
            h = encoder(mode, x.permute(0, 2, 1), sd, hps).permute(0, 2, 1).reshape(-1, hps.emb_width)
            lat[mode] = h
            codes[mode] = vr.quantise(h, k)[0]
    n = codes["fp32"].numel()
    for mode in ("tf32", "3xtf32"):
        agree = float((codes[mode] == codes["fp32"]).float().mean())
        err = float((lat[mode] - lat["fp32"]).abs().max() / lat["fp32"].abs().max())
        print(f"{mode:7s}: index agreement {agree:.5f} over {n} latents, max latent error {err:.2e} of scale")


if __name__ == "__main__":
    main()

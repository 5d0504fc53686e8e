#!/usr/bin/env python
"""VQ-VAE encode -> quantise -> decode throughput (BASELINE.json configs[1]: synthetic 8-frame pose
batches, 1 x B200) plus the call-site shapes [1,240,135] (make_beat_dataset.py:315) and a 180-code decode
(VisualizeCodebook.py:139).  Prints one JSON line per shape; checks indices against the float32 oracle."""
import argparse, json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import vqvae_ref as vr
from qpgesture_b200 import _lib
from qpgesture_b200.vqvae import VQVAE

ENC_GFLOP_240, DEC_GFLOP_240 = 1.6235, 1.9083

def timed(fn, reps):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): out = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, out

def torch_stack(sd, hps, dev):
    """The same encoder / decoder as stock PyTorch modules (cuDNN; TF32 allowed) - the "before" number BASELINE.md
    section 3 asks for.  NCT layout as the reference (encdec.py, resnet.py)."""
    import torch.nn.functional as F
    sd = {k: v.to(dev) for k, v in sd.items()}
    down_t, depth, g = hps.downs_t[0], hps.depth, hps.dilation_growth_rate

    def res(x, pre, dils):
        for d, dil in enumerate(dils):
            p = f"{pre}.model.{d}.model"
            h = F.conv1d(F.relu(x), sd[p + ".1.weight"], sd[p + ".1.bias"], padding=dil, dilation=dil)
            x = x + F.conv1d(F.relu(h), sd[p + ".3.weight"], sd[p + ".3.bias"])
        return x

    def enc(x_ntc):
        x = x_ntc.permute(0, 2, 1)
        pre = "encoders.0.level_blocks.0.model"
        for i in range(down_t):
            x = F.conv1d(x, sd[f"{pre}.{i}.0.weight"], sd[f"{pre}.{i}.0.bias"], stride=2, padding=1)
            x = res(x, f"{pre}.{i}.1", [g ** d for d in range(depth)])
        x = F.conv1d(x, sd[f"{pre}.{down_t}.weight"], sd[f"{pre}.{down_t}.bias"], padding=1)
        xf = x.permute(0, 2, 1).reshape(-1, x.shape[1])
        k = sd["bottleneck.level_blocks.0.k"]
        dist = (xf ** 2).sum(-1, keepdim=True) - 2 * xf @ k.t() + (k ** 2).sum(-1)[None]
        return dist.argmin(-1).view(x.shape[0], -1)

    def dec(codes):
        k = sd["bottleneck.level_blocks.0.k"]
        x = F.embedding(codes, k).permute(0, 2, 1)
        pre = "decoders.0.level_blocks.0.model"
        x = F.conv1d(x, sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"], padding=1)
        dils = [g ** d for d in range(depth)]
        if hps.vqvae_reverse_decoder_dilation:
            dils = dils[::-1]
        for i in range(down_t):
            x = res(x, f"{pre}.{i + 1}.0", dils)
            x = F.conv_transpose1d(x, sd[f"{pre}.{i + 1}.1.weight"], sd[f"{pre}.{i + 1}.1.bias"], stride=2, padding=1)
        x = F.conv1d(x, sd["decoders.0.out.weight"], sd["decoders.0.out.bias"], padding=1)
        return x.permute(0, 2, 1)
    return enc, dec


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="1x8,64x8,4096x8,65536x8,1x240,64x240")
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--check", type=int, default=1)
    ap.add_argument("--torch-arm", choices=("none", "tf32", "fp32"), default="none",
                    help="time the same stacks as stock PyTorch (cuDNN) instead of this library")
    a = ap.parse_args()
    hps = vr.make_hps()
    sd = vr.random_state_dict(hps, 135, seed=0, codebook_seed=1)
    if a.torch_arm != "none":
        torch.backends.cudnn.allow_tf32 = a.torch_arm == "tf32"
        torch.backends.cuda.matmul.allow_tf32 = a.torch_arm == "tf32"
        torch.backends.cudnn.benchmark = True
        enc, dec = torch_stack(sd, hps, "cuda")
        with torch.no_grad():
            for shp in a.shapes.split(","):
                B, T = map(int, shp.split("x"))
                xd = torch.randn((B, T, 135), generator=torch.Generator().manual_seed(0)).cuda()
                reps = 20 if B * T <= 65536 else 3
                ms_enc, zs = timed(lambda: enc(xd), reps)
                ms_dec, _ = timed(lambda: dec(zs), reps)
                frames = B * T
                print(json.dumps(dict(shape=[B, T, 135], arm="torch_cudnn_" + a.torch_arm, encode_ms=ms_enc, decode_ms=ms_dec,
                                      encode_tflops=ENC_GFLOP_240 * frames / 240 / ms_enc,
                                      decode_tflops=DEC_GFLOP_240 * frames / 240 / ms_dec)), flush=True)
        return
    model = VQVAE(hps, 135, device="cuda", precision=a.precision).load_state_dict(sd)
    for shp in a.shapes.split(","):
        B, T = map(int, shp.split("x"))
        x = torch.randn((B, T, 135), generator=torch.Generator().manual_seed(0))
        xd = x.cuda()
        reps = 20 if B * T <= 65536 else 3
        l0 = _lib.launch_count()
        ms_enc, zs = timed(lambda: model.encode(xd), reps)
        launches = (_lib.launch_count() - l0) // (reps + 3)
        ms_dec, dec = timed(lambda: model.decode(zs), reps)
        frames = B * T
        rec = dict(shape=[B, T, 135], precision=a.precision, encode_ms=ms_enc, decode_ms=ms_dec,
                   encode_frames_per_s=frames / ms_enc * 1e3, decode_frames_per_s=frames / ms_dec * 1e3,
                   codes_per_s=frames / 8 / ms_enc * 1e3,
                   encode_tflops=ENC_GFLOP_240 * frames / 240 / ms_enc, decode_tflops=DEC_GFLOP_240 * frames / 240 / ms_dec,
                   launches_per_encode=int(launches))
        if a.check and frames <= 64 * 240:
            want = vr.encode(x, sd, hps)
            got = zs[0].cpu()
            rec["index_match"] = float((got == want).float().mean())
            wd = vr.decode(got, sd, hps)
            rec["decode_max_abs_err"] = float((dec.cpu() - wd).abs().max())
        print(json.dumps(rec), flush=True)

if __name__ == "__main__":
    main()

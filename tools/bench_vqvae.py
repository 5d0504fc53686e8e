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

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shapes", default="1x8,64x8,4096x8,65536x8,1x240,64x240")
    ap.add_argument("--precision", type=int, default=0)
    ap.add_argument("--check", type=int, default=1)
    a = ap.parse_args()
    hps = vr.make_hps()
    sd = vr.random_state_dict(hps, 135, seed=0, codebook_seed=1)
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

"""B200 drop-in for the VQ-VAE call surface of the reference's
codebook/VisualizeCodebook.py: `visualize_code` (:119-154), `cal_distance`
(:93-116) and the CLI flags of configs/parse_args.py:4-18.

Decoding runs on the CUDA VQVAE of qpgesture_b200.vqvae.  The BVH / video
post-processing that follows in the reference (:361-370: pymo pipeline, FK,
matplotlib/ffmpeg) is outside the hot path (out of scope) and is
not reproduced; the .npy outputs are the hand-over point.
"""
from __future__ import annotations

import argparse
import os
from types import SimpleNamespace

import numpy as np
import torch

from .vqvae import VQVAE


def parse_args(argv=None):
    """configs/parse_args.py:4-18 (same flags and defaults)."""
    p = argparse.ArgumentParser(description="Codebook")
    p.add_argument("--config", default="./configs/codebook.yml")
    p.add_argument("--gpu", type=str, default="2")
    p.add_argument("--no_cuda", type=list, default=["2"])
    p.add_argument("--prefix", type=str, required=False, default="knn_pred_wavvq")
    p.add_argument("--save_path", type=str, required=False, default="./Speech2GestureMatching/output/")
    p.add_argument("--code_path", type=str, required=False)
    p.add_argument("--VQVAE_model_path", type=str, required=False)
    p.add_argument("--BEAT_path", type=str, default="../dataset/orig_BEAT/speakers/")
    p.add_argument("--save_dir", type=str, default="../dataset/BEAT")
    p.add_argument("--step", type=str, default="1")
    p.add_argument("--stage", type=str, default="train")
    return p.parse_args(argv)


def _as_ns(d):
    return d if not isinstance(d, dict) else SimpleNamespace(**d)


def load_model(args, model_path, device=None):
    """VisualizeCodebook.py:129-134: VQVAE(args.VQVAE, 15*9), checkpoint['model_dict'] (DataParallel keys)."""
    model = VQVAE(_as_ns(args.VQVAE), 15 * 9, device=device)
    checkpoint = torch.load(model_path, map_location=torch.device("cpu"))
    model.load_state_dict(checkpoint["model_dict"])
    return model.eval()


def visualize_code(args, model_path, save_path, prefix, code_source, normalize=True, model=None, device=None):
    """Decode all codes as ONE sequence, de-normalise with the yml mean / clip(std, .01),
    save code<prefix>.npy and generate<prefix>.npy (VisualizeCodebook.py:119-154)."""
    if normalize:
        data_mean = np.array(args.data_mean).squeeze()
        data_std = np.array(args.data_std).squeeze()
        std = np.clip(data_std, a_min=0.01, a_max=None)
    model = model if model is not None else load_model(args, model_path, device)
    zs = [torch.from_numpy(np.asarray(code_source).flatten()).unsqueeze(0)]
    pose_sample = model.decode(zs).squeeze(0).cpu().numpy()
    out_code = np.vstack([zs[0].squeeze(0).cpu().numpy()])
    out_poses = np.vstack([pose_sample])
    if normalize:
        out_poses = np.multiply(out_poses, std) + data_mean
    if save_path is not None:
        os.makedirs(save_path, exist_ok=True)
        np.save(os.path.join(save_path, "code" + prefix + ".npy"), out_code)
        np.save(os.path.join(save_path, "generate" + prefix + ".npy"), out_poses)
    return out_poses, out_code


def cal_distance(args, model_path, save_path, prefix, normalize=True, model=None, device=None,
                 out_file="./output/code.npz"):
    """Decode every code x 30 and store code / poses / signature = time-mean pose
    (VisualizeCodebook.py:93-116); the 512 batch-1 decodes of the reference are one batch here."""
    model = model if model is not None else load_model(args, model_path, device)
    n_codes = model.l_bins
    code = np.repeat(np.arange(n_codes, dtype=np.int64)[:, None], 30, axis=1)
    poses = model.decode([torch.from_numpy(code)]).cpu().numpy()
    if out_file is not None:
        os.makedirs(os.path.dirname(os.path.abspath(out_file)), exist_ok=True)
        np.savez_compressed(out_file, code=code, poses=poses, signature=np.mean(poses, axis=1))
    return code, poses, np.mean(poses, axis=1)


def visualizeCodeAndWrite(code_path=None, save_path="./Speech2GestureMatching/output/", prefix=None,
                          pipeline_path="../data/data_pipe_60_rotation.sav", generateGT=True, code_source=None, vis=True,
                          *, config=None, model_path=None, model=None, bvh=True):
    """Inference half of VisualizeCodebook.py:333-370 (code file -> poses -> Euler angles / BVH) with the reference's
    positional arguments; the files go to os.path.join(save_path, prefix) like the reference's (:342).  The BVH step
    (:361, `make_bvh_GENEA2020_BT(..., smoothing=False, pipeline_path=...)`) runs its numeric half on the device
    (qpgesture_b200.process_bvh) and writes `<prefix>_generated.bvh` when pymo and the fitted pipeline are available,
    else `<prefix>_generated_euler.npy`; `bvh=False` stops after the poses.  The ground-truth branch (generateGT:
    parsing a BVH file with pymo, :346-355) and the video rendering (vis, :363-370) are not built.  `config`
    (keyword, the reference reads a module global) carries VQVAE / data_mean / data_std / VQVAE_model_path."""
    assert config is not None, "pass config= (the reference reads a module-level global)"
    if code_source is None:
        code_source = np.load(code_path)["knn_pred"]                     # :357
    model_path = model_path or getattr(config, "VQVAE_model_path", None)
    save_path = os.path.join(save_path, prefix)                          # :342
    out_poses, out_code = visualize_code(config, model_path, save_path, prefix, code_source, model=model)
    if bvh:
        from .process_bvh import make_bvh_GENEA2020_BT
        make_bvh_GENEA2020_BT(save_path, prefix, out_poses, smoothing=False, pipeline_path=pipeline_path,
                              device=getattr(model, "device", None))          # :361
    return out_poses, out_code


def main(argv=None):
    import yaml

    a = parse_args(argv)
    with open(a.config) as f:
        config = yaml.safe_load(f)
    for k, v in vars(a).items():
        config[k] = v
    config = SimpleNamespace(**config)
    dev = torch.device("cuda", int(a.gpu)) if torch.cuda.device_count() > int(a.gpu) else torch.device("cuda", 0)
    if a.stage == "train":
        model = load_model(config, a.VQVAE_model_path, dev)
        return cal_distance(config, a.VQVAE_model_path, a.save_path, a.prefix, model=model)
    model = load_model(config, a.VQVAE_model_path, dev)
    return visualizeCodeAndWrite(code_path=a.code_path, prefix=a.prefix, generateGT=False, save_path=a.save_path,
                                 config=config, model=model)            # :393



def dataset_to_code(poses, model, data_mean=None, data_std=None, batch=256):
    """Bulk version of process/make_beat_dataset.py:291-325 (`subdataset_to_code`): normalise
    (x - mean) / clip(std, 0.01) and encode.  The reference encodes one [1,240,135] sequence per call;
    here sequences go through the encoder `batch` at a time.  poses [N, 240, 135] -> int64 codes [N, 30]."""
    x = np.asarray(poses, dtype=np.float64)
    if data_mean is not None:
        std = np.clip(np.asarray(data_std, dtype=np.float64).squeeze(), a_min=0.01, a_max=None)
        x = (x - np.asarray(data_mean, dtype=np.float64).squeeze()) / std
    out = []
    for i in range(0, x.shape[0], batch):
        zs = model.encode(torch.from_numpy(x[i:i + batch]).float())
        out.append(zs[0].cpu().numpy())
    return np.concatenate(out, axis=0) if out else np.zeros((0, 0), dtype=np.int64)


if __name__ == "__main__":
    main()

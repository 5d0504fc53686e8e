"""Generate tests/golden/pae_*.npz from the UNMODIFIED reference periodic auto-encoder (codebook/PAE.py, imported in
place through oracle/ref_harness.import_pae; build container only).

Weights are the seeded `oracle.pae_np.random_state_dict` loaded into the reference's own Model through
load_state_dict (the repository ships no PAE checkpoint); the tests rebuild them from the seed, so only the
reference's OUTPUTS are stored: pose2phase of a seeded pose sequence, and Model.forward of two seeded windows."""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import pae_np  # noqa: E402
from oracle import ref_harness as rh  # noqa: E402

CASES = [dict(name="pae_s0", seed=3, T=40), dict(name="pae_s1", seed=4, T=7)]


def inputs(seed, T):
    """Seeded pose sequence (random walk, so frame differences are O(1) after normalisation), mean and std."""
    rng = np.random.default_rng(seed + 50)
    pose = np.cumsum(rng.standard_normal((T + 300, 135)) * 0.6, axis=0)[150:150 + T]
    mean = rng.standard_normal(135) * 0.1
    std = np.clip(rng.uniform(0.3, 1.5, 135), 0.01, None)
    xw = (rng.standard_normal((2, 135 * 240)) * 0.3).astype(np.float32)
    return pose, mean, std, xw


def main():
    m = rh.import_pae()
    for c in CASES:
        sd = pae_np.random_state_dict(c["seed"])
        net = m.Model(input_channels=135, embedding_channels=8, time_range=240, key_range=13, window=4.0).eval()
        net.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()}, strict=True)
        pose, mean, std, xw = inputs(c["seed"], c["T"])
        with torch.no_grad():
            with rh._quiet():
                phase = m.pose2phase(net, pose, mean, std)
            y, latent, signal, params = net(torch.from_numpy(xw))
        h = hashlib.sha256()
        for k in sorted(sd):
            h.update(np.asarray(sd[k]).tobytes())
        np.savez_compressed(os.path.join(HERE, c["name"] + ".npz"), phase=phase, latent=latent.numpy(),
                            signal=signal.numpy(), y_head=y.numpy()[:, :2048],
                            y_sum=y.numpy().astype(np.float64).sum(axis=1),
                            params=np.stack([p.numpy() for p in params], axis=1), sd_digest=h.hexdigest(),
                            seed=c["seed"], T=c["T"], torch_version=torch.__version__)
        print(c["name"], phase.shape, latent.shape, float(np.abs(latent.numpy()).max()))
    rh.release_pae()


if __name__ == "__main__":
    main()

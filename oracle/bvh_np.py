"""CPU restatement of the numeric half of `make_bvh_GENEA2020_BT` (process/process_bvh.py:57-77) (TEST INFRASTRUCTURE).

The reference module cannot be imported here (it pulls in pymo at import time and loads an unshipped joblib pipeline),
so this oracle makes the SAME library calls with the same arguments as the cited lines: scipy.signal.savgol_filter(x,
15, 2) per channel (:64-66) and scipy Rotation.from_matrix(...).as_euler('ZXY', degrees=True) per frame (:71-77) -
pinned by construction on scipy (1.18 in this image), the third-party arithmetic the reference delegates to.
Only tests/ may import this module."""
from __future__ import annotations

import numpy as np
from scipy.signal import savgol_filter
from scipy.spatial.transform import Rotation as R


def poses_to_euler(poses, smoothing=True):
    poses = np.asarray(poses)
    if smoothing:                                                            # :60-66
        n_poses = poses.shape[0]
        out_poses = np.zeros((n_poses, poses.shape[1]))
        for i in range(poses.shape[1]):
            out_poses[:, i] = savgol_filter(poses[:, i], 15, 2)
    else:
        out_poses = poses
    out_poses = out_poses.reshape((out_poses.shape[0], -1, 9))               # :71-72
    out_poses = out_poses.reshape((out_poses.shape[0], out_poses.shape[1], 3, 3))
    out_euler = np.zeros((out_poses.shape[0], out_poses.shape[1] * 3))
    for i in range(out_poses.shape[0]):                                      # :74-77
        out_euler[i] = R.from_matrix(out_poses[i]).as_euler('ZXY', degrees=True).flatten()
    return out_euler

"""Post-decode pose processing on the device: the numeric half of the reference's `make_bvh_GENEA2020_BT`
(process/process_bvh.py:57-83), which VisualizeCodebook.visualize_code calls on the decoded poses (:361-370).

    poses [T, 9 * J] float32 (row-major 3x3 rotation matrices per joint, the VQ-VAE decoder's output)
      -> Savitzky-Golay smoothing per channel (window 15, order 2; :64-66)
      -> Rotation.from_matrix(...).as_euler('ZXY', degrees=True) per joint and frame (:71-77)
      -> out_euler [T, 3 * J] float64

`make_bvh_GENEA2020_BT` keeps the reference's signature; the last two steps (pymo `pipeline.inverse_transform` and
`BVHWriter`, :79-83) need pymo and the fitted `data_pipe_60_rotation.sav`, neither of which ships with the reference
repository: when they are available they run on the host as in the reference, otherwise the Euler angles are written
to `<prefix>_generated_euler.npy` next to where the BVH file would go.  No CPU fallback for the numeric part."""
from __future__ import annotations

import os
import warnings

import numpy as np
import torch

from . import _lib

_SG_WINDOW, _SG_ORDER = 15, 2


def _savgol_tables():
    """Interior coefficients (= scipy.signal.savgol_coeffs(15, 2): the least-squares quadratic evaluated at the
    window centre) and the edge rows of scipy's mode='interp': the quadratic fitted to the first / last 15 samples
    evaluated at the first / last 7 positions."""
    pos = np.arange(_SG_WINDOW, dtype=np.float64)
    fit = np.linalg.pinv(np.vander(pos, _SG_ORDER + 1, increasing=True))                 # [3, 15]
    ev = lambda p: np.vander(np.asarray(p, dtype=np.float64), _SG_ORDER + 1, increasing=True) @ fit
    half = _SG_WINDOW // 2
    return ev([half])[0], ev(np.arange(half)), ev(np.arange(half + 1, _SG_WINDOW))


def poses_to_euler(poses, smoothing=True, device=None):
    """process_bvh.py:60-77 on the device -> float64 [T, 3 * J] (degrees, intrinsic ZXY per joint)."""
    dev = torch.device(device if device is not None else "cuda")
    lib = _lib.load()
    x = torch.as_tensor(np.ascontiguousarray(poses, dtype=np.float32) if not torch.is_tensor(poses) else poses,
                        dtype=torch.float32, device=dev).contiguous()
    if x.dim() != 2 or x.shape[1] % 9 != 0:
        raise ValueError("poses must be [T, 9 * J]")
    T, C = x.shape
    sp = _lib.stream_ptr()
    if smoothing:
        if T < _SG_WINDOW:
            raise ValueError("If mode is 'interp', window_length must be less than or equal to the size of x.")   # scipy's
        coef, first, last = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in _savgol_tables())
        sm = torch.empty((T, C), dtype=torch.float64, device=dev)
        _lib.check(lib.qpg_savgol15_f64(_lib.ptr(x), T, C, _lib.ptr(coef), _lib.ptr(first), _lib.ptr(last), _lib.ptr(sm), sp),
                   "qpg_savgol15_f64")
    else:
        sm = x.double()
    n = T * (C // 9)
    euler = torch.empty((n, 3), dtype=torch.float64, device=dev)
    flags = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.check(lib.qpg_rotmat_to_euler_zxy(_lib.ptr(sm), n, _lib.ptr(euler), _lib.ptr(flags), sp), "qpg_rotmat_to_euler_zxy")
    fl = flags.cpu().numpy()
    if (fl & 1).any():
        bad = int(np.nonzero(fl & 1)[0][0])
        raise ValueError(f"Non-positive determinant (left-handed or null coordinate frame) in rotation matrix {bad}")
    if (fl & 2).any():
        warnings.warn("Gimbal lock detected. Setting third angle to zero since it is not possible to uniquely "
                      "determine all angles.")
    return euler.reshape(T, (C // 9) * 3).cpu().numpy()


def make_bvh_GENEA2020_BT(save_path, filename_prefix, poses, smoothing=True, pipeline_path='./resource/data_pipe_60.sav',
                          device=None):
    """process_bvh.py:57-83.  Returns the path written (the .bvh, or the Euler .npy when pymo / the pipeline are absent)."""
    out_euler = poses_to_euler(poses, smoothing=smoothing, device=device)
    os.makedirs(save_path, exist_ok=True)
    try:
        import joblib as jl
        from pymo.writers import BVHWriter                    # noqa: F401  (not installed in this image)
        pipeline = jl.load(pipeline_path)
    except (ImportError, FileNotFoundError) as e:
        out = os.path.join(save_path, filename_prefix + '_generated_euler.npy')
        np.save(out, out_euler)
        warnings.warn(f"BVH writing needs pymo and {pipeline_path} ({type(e).__name__}: {e}); Euler angles saved to {out}")
        return out
    bvh_data = pipeline.inverse_transform([out_euler])
    out_bvh_path = os.path.join(save_path, filename_prefix + '_generated.bvh')
    with open(out_bvh_path, 'w') as f:
        BVHWriter().write(bvh_data[0], f)
    return out_bvh_path

"""Host tables of the device Savitzky-Golay filter against scipy (the library the reference calls at
process/process_bvh.py:64-66): interior coefficients and the 'interp' edge rows."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_savgol_tables_reproduce_scipy():
    from scipy.signal import savgol_coeffs, savgol_filter
    from qpgesture_b200.process_bvh import _savgol_tables

    coef, first, last = _savgol_tables()
    assert np.allclose(coef, savgol_coeffs(15, 2)[::-1], rtol=0, atol=1e-15)
    rng = np.random.default_rng(0)
    for T in (15, 16, 40):
        x = rng.standard_normal(T)
        want = savgol_filter(x, 15, 2)
        got = np.empty(T)
        got[:7] = first @ x[:15]
        got[T - 7:] = last @ x[T - 15:]
        for t in range(7, T - 7):
            got[t] = coef @ x[t - 7:t + 8]
        assert np.abs(got - want).max() < 1e-13


def test_oracle_euler_convention():
    """intrinsic ZXY: R = Rz(a) Rx(b) Ry(c) - the closed form the device kernel uses"""
    from oracle import bvh_np
    from scipy.spatial.transform import Rotation as R

    m = R.from_euler('ZXY', [[10.0, 20.0, 30.0], [-100.0, 45.0, 170.0]], degrees=True).as_matrix()
    e = bvh_np.poses_to_euler(m.reshape(1, 18), smoothing=False)
    assert np.allclose(e.reshape(2, 3), [[10.0, 20.0, 30.0], [-100.0, 45.0, 170.0]], atol=1e-10)

"""ctypes binding of libqpg_sm100.so (C ABI declared in include/qpg.h).

The product path has no CPU fallback: if the library is missing, or a call
fails, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libqpg_sm100.so")

_lib = None


class QpgError(RuntimeError):
    pass


class ConvDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T_in", C.c_int), ("T_out_total", C.c_int), ("C_in", C.c_int), ("C_out", C.c_int),
        ("n_taps", C.c_int), ("tap_offset", C.c_int * 4), ("in_stride", C.c_int), ("out_stride", C.c_int),
        ("out_offset", C.c_int), ("n_out", C.c_int), ("relu_in", C.c_int), ("precision", C.c_int),
    ]


class ConvTcDesc(C.Structure):
    _fields_ = [
        ("B", C.c_int), ("T_view", C.c_int), ("C_view", C.c_int), ("n_out", C.c_int), ("C_in", C.c_int),
        ("C_out", C.c_int), ("K_pad", C.c_int), ("N_pad", C.c_int), ("BN", C.c_int), ("n_taps", C.c_int),
        ("row_offset", C.c_int * 4), ("chan_offset", C.c_int * 4), ("out_rows_per_item", C.c_int),
        ("out_ld", C.c_int), ("out_chan_offset", C.c_int),
    ]


class SlicedSeg(C.Structure):
    _fields_ = [("db_slices", C.c_void_p), ("q_slices", C.c_void_p), ("sacc", C.c_void_p), ("n_kblocks", C.c_int32),
                ("pad", C.c_int32)]


class SliceJob(C.Structure):
    _fields_ = [("q", C.c_void_p), ("col_exp", C.c_void_p), ("q_slices", C.c_void_p), ("q_info", C.c_void_p),
                ("ldq", C.c_int64), ("D", C.c_int32), ("pad", C.c_int32)]


class SlicedTable(C.Structure):
    _fields_ = [("packed", C.c_void_p), ("row_sqnorm", C.c_void_p), ("q", C.c_void_p), ("q_info", C.c_void_p),
                ("ldq", C.c_int64), ("D", C.c_int32), ("pad", C.c_int32), ("sacc", C.c_void_p),
                ("bin_start", C.c_void_p), ("row_info", C.c_void_p), ("order", C.c_void_p), ("bins", C.c_void_p),
                ("bins_qstride", C.c_int64), ("table", C.c_void_p), ("ranks", C.c_void_p), ("qflags", C.c_void_p)]


def dptr(t):
    """raw device address (int) of a tensor, 0 for None - for ctypes struct fields"""
    return 0 if t is None else ptr(t).value


_P = C.c_void_p
_I64 = C.c_int64
_INT = C.c_int

# name -> (restype, argtypes); every symbol include/qpg.h declares
SIGNATURES = {
    "qpg_version": (_INT, []),
    "qpg_last_error": (C.c_char_p, []),
    "qpg_launch_count": (C.c_uint64, []),
    "qpg_tune_cosine": (_INT, [_INT, _INT, _INT, _INT]),
    "qpg_tune_cosine_alternate": (_INT, [_INT]),
    "qpg_packed_bytes": (C.c_size_t, [_I64, _INT]),
    "qpg_pack_rows_f32": (_INT, [_P, _I64, _INT, _P, _P, _P]),
    "qpg_stack_wavlm_rows": (_INT, [_P, _I64, _INT, _INT, _INT, _INT, _INT, _P, _P]),
    "qpg_table_init": (_INT, [_P, _I64, _P]),
    "qpg_cand_cosine_minbycode": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P, _INT, _P, _INT, _P]),
    "qpg_cand_cosine_minbycode_team": (_INT, [_P, _P, _P, _I64, _INT, _I64, _P, _INT, _P, _INT, _INT, _P]),
    "qpg_cand_cosine2_minbycode": (_INT, [_P, _P, _P, _P, _I64, _INT, _INT, _I64, _P, _INT, _P, _P, _P]),
    "qpg_sliced_bytes": (C.c_size_t, [_I64, _INT]),
    "qpg_sliced_query_bytes": (C.c_size_t, [_INT, _INT]),
    "qpg_slice_rows_i8": (_INT, [_P, _I64, _INT, _P, _P, _P, _P, _P, _P]),
    "qpg_slice_queries_i8": (_INT, [_P, _INT, _INT, _INT, _P]),
    "qpg_sliced_scan_i8": (_INT, [_P, _INT, _I64, _INT, _INT, _P]),
    "qpg_sliced_scan_ref": (_INT, [_P, _P, _INT, _I64, _INT, _INT, _INT, _P, _P]),
    "qpg_sliced_bins": (_INT, [_P, _INT, _I64, _INT, _I64, _I64, _INT, _P, _P]),
    "qpg_sliced_resolve": (_INT, [_P, _INT, _INT, _I64, _INT, _I64, _P, _P]),
    "qpg_match_lookup": (_INT, [_P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _INT, _P, _P]),
    "qpg_match_walk": (_INT, [_P, _P, _P, _P, _P, _INT, _INT, _P, _P, _P, _P, _P, _P]),
    "qpg_match_walk_stats": (_INT, [_P, _P, _P, _P, _P, _P, _INT, _INT, _P, _P, _P, _P, _P, _P]),
    "qpg_phase_stats_floats": (C.c_size_t, []),
    "qpg_phase_stats": (_INT, [_P, _I64, _P, _P, _P, _P]),
    "qpg_cand_lev_minbycode": (_INT, [_P, _P, _I64, _I64, _P, _INT, _P, _P]),
    "qpg_lev_distance": (_INT, [_P, _P, _I64, _P, _P]),
    "qpg_l2_prefetch": (_INT, [_P, C.c_size_t, _P]),
    "qpg_table_merge": (_INT, [_P, _INT, _I64, _P, _P]),
    "qpg_rank512": (_INT, [_P, _INT, _P, _P]),
    "qpg_rank512_ties": (_INT, [_P, _INT, _P, _P, _P]),
    "qpg_match_tail": (_INT, [_P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _INT, _INT, _P, _P, _P, _P, _P]),
    "qpg_match_tail_segments": (_INT, [_P, _P, _P, _P, _P, _P, _P, _I64, _P, _P, _P, _P, _P, _INT, _INT, _INT, _INT, _P,
                                       _P, _P, _P, _P, _P]),
    "qpg_vq_argmin_f32": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P]),
    "qpg_vq_argmin_fast": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P, _P, _P]),
    "qpg_vq_dequantise_f32": (_INT, [_P, _P, _I64, _INT, _INT, _P, _P]),
    "qpg_conv1d_taps_f32": (_INT, [C.POINTER(ConvDesc), _P, _P, _P, _P, _P, _P]),
    "qpg_conv1d_taps_tf32": (_INT, [C.POINTER(ConvTcDesc), _P, _P, _P, _P, _P, _P, _P]),
    "qpg_conv1d_taps_3xtf32": (_INT, [C.POINTER(ConvTcDesc), _P, _P, _P, _P, _P, _P, _P, _P]),
    "qpg_pae_sliding_conv1": (_INT, [_P, _P, _P, _P, _INT, _INT, _INT, _INT, _P, _P]),
    "qpg_pae_conv1d": (_INT, [_P, _P, _P, _P, _INT, _INT, _INT, _INT, _INT, _INT, _INT, _P, _P]),
    "qpg_pae_params": (_INT, [_P, _P, _P, _P, _P, C.c_float, _INT, _INT, _INT, _P, _P]),
    "qpg_savgol15_f64": (_INT, [_P, _INT, _INT, _P, _P, _P, _P, _P]),
    "qpg_rotmat_to_euler_zxy": (_INT, [_P, _I64, _P, _P, _P]),
    "qpg_legacy_cand_bytes": (C.c_size_t, []),
    "qpg_legacy_candidates": (_INT, [_P, _P] + [_INT] * 8 + [_P, _INT, _P, _P, _P, _P, _P]),
    "qpg_legacy_pick": (_INT, [_P, _P, _P, _INT, _INT, _P, _P, _P, _P, _P]),
    "qpg_legacy_gather": (_INT, [_P, _P] + [_INT] * 8 + [_P, _P, _P] + [_INT] * 4 + [_P] * 5),
}


def load():
    """Load the native library (once).  Raises QpgError when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise QpgError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(or `make -C qpgesture_b200/csrc`). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str):
    if rc != 0:
        msg = load().qpg_last_error().decode("utf-8", "replace")
        raise QpgError(f"{what} failed (rc={rc}): {msg}")


def ptr(t):
    """Device pointer of a torch tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise QpgError("expected a CUDA tensor")
    if not t.is_contiguous():
        raise QpgError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


def stream_ptr(stream=None):
    import torch

    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def launch_count() -> int:
    return int(load().qpg_launch_count())

"""Database / query loading for the matcher (host side).

Mirrors `load_db_codebook` of the reference
(codebook/Speech2GestureMatching/data_processing.py:197-353): same arguments,
same 19-tuple, same axis order (everything transposed to (N, feat, time)).
Differences, all value-preserving:
  * the stacked WavLM feature is returned as float32 instead of float64 (its
    values are float32-exact; the reference only gets float64 from np.zeros)
  * `phase` may be stored dense ([N,240,4,8] float32) as well as in the
    reference's pickled object form.
`load_match_inputs` is the lean loader the CLI uses: it builds only what the
CodeKNN path reads (window rows, queries), never the [N,180,6144] tensor.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from .constant import (FRAME_INTERVAL, NUM_AUDIO_FEAT_FRAMES, NUM_MFCC_FEAT, STEP_SZ, WINDOWS_PER_SEQ,
                       num_frames_code)


def _stack_future_taps(x: np.ndarray, interval, dtype=np.float64, n_taps=NUM_AUDIO_FEAT_FRAMES) -> np.ndarray:
    """out[:, t, i, :] = x[:, t + int(i*interval), :] (zero past the end); [n,T,n_taps*C]."""
    n, T, Cc = x.shape
    out = np.zeros((n, T, n_taps, Cc), dtype=dtype)
    for i in range(n_taps):
        off = int(i * interval)
        if off < T:
            out[:, :T - off, i, :] = x[:, off:, :]
    return out.reshape(n, T, n_taps * Cc)


def interpolate_wavlm(wavlm: np.ndarray, n_code: int = num_frames_code) -> np.ndarray:
    """[n,199,C] -> [n,180,C] exactly as data_processing.py:257-259 (same torch call)."""
    new_t = wavlm.shape[1] // n_code * n_code
    x = torch.from_numpy(np.ascontiguousarray(wavlm, dtype=np.float32)).transpose(1, 2)
    y = F.interpolate(x, size=new_t, align_corners=True, mode="linear")
    return y.transpose(1, 2).contiguous().numpy()


def stack_wavlm_feat(interp: np.ndarray, dtype=np.float32) -> np.ndarray:
    """[n,180,C] -> [n,180,6C]: 6 taps at stride FRAME_INTERVAL-2 (data_processing.py:264-268)."""
    return _stack_future_taps(interp, FRAME_INTERVAL - 2, dtype=dtype)


def stack_wavvq_feat(wavvq: np.ndarray) -> np.ndarray:
    """[n,398,2] -> [n,398,22]: 6 past taps then 5 future taps at multiples of 398/30
    (data_processing.py:304-322; identical for the test split, :324-341)."""
    n, T, G = wavvq.shape
    fi = T / num_frames_code
    out = np.zeros((n, T, 2 * NUM_AUDIO_FEAT_FRAMES - 1, G))
    for i in range(NUM_AUDIO_FEAT_FRAMES):
        pre = int((NUM_AUDIO_FEAT_FRAMES - i - 1) * fi)
        out[:, pre:, i, :] = wavvq[:, :T - pre, :]
    for i in range(1, NUM_AUDIO_FEAT_FRAMES):
        post = int(i * fi)
        out[:, :T - post, NUM_AUDIO_FEAT_FRAMES + i - 1, :] = wavvq[:, post:, :]
    return out.reshape(n, T, -1)


def wavlm_window_rows(interp: np.ndarray) -> np.ndarray:
    """Rows the matcher scans in WavLM mode: wavlm_train_feat[j, 6m] for m < 26
    (GestureKNN.py:671-690) -> float32 [n*26, 6C], built without the full stack."""
    n, T, Cc = interp.shape
    step = T // num_frames_code
    m = np.arange(WINDOWS_PER_SEQ) * step
    taps = np.arange(NUM_AUDIO_FEAT_FRAMES) * (FRAME_INTERVAL - 2)
    idx = m[:, None] + taps[None, :]                       # [26, 6] frame ids, all < T
    assert idx.max() < T
    return np.ascontiguousarray(interp[:, idx, :].reshape(n * WINDOWS_PER_SEQ, NUM_AUDIO_FEAT_FRAMES * Cc),
                                dtype=np.float32)


def wavlm_query_rows(interp: np.ndarray) -> np.ndarray:
    """Query features of the 8 steps of each test segment: test_wavlm_feat[g, 24 s]
    (GestureKNN.py:528,565,659) -> float32 [M, 8, 6C]."""
    n, T, Cc = interp.shape
    step = STEP_SZ * (T // num_frames_code)
    i_list = np.arange(0, T, step)
    taps = np.arange(NUM_AUDIO_FEAT_FRAMES) * (FRAME_INTERVAL - 2)
    idx = i_list[:, None] + taps[None, :]
    padded = np.concatenate((interp, np.zeros((n, max(0, idx.max() + 1 - T), Cc), dtype=interp.dtype)), axis=1)
    return np.ascontiguousarray(padded[:, idx, :].reshape(n, len(i_list), -1), dtype=np.float32)


def load_db_codebook(data_file, codepath, test_data_path, train_wavlm, test_wavlm, train_wavvq, test_wavvq):
    """Same contract as the reference's load_db_codebook (19-tuple, (N, feat, time))."""
    data = np.load(data_file, allow_pickle=True)
    test_data = np.load(test_data_path, allow_pickle=True)
    code = np.load(codepath)["code"]

    def split(d):
        mfcc = d["mfcc"][:, :, :NUM_MFCC_FEAT]
        sp = np.stack((d["energy"], d["pitch"], d["volume"]), axis=-1)
        return mfcc, sp, _stack_future_taps(mfcc, FRAME_INTERVAL), _stack_future_taps(sp, FRAME_INTERVAL)

    mfcc, speech, feat, speech_feat = split(data)
    t_mfcc, t_speech, t_feat, t_speech_feat = split(test_data)

    w_tr = interpolate_wavlm(np.load(train_wavlm)["wavlm"])
    w_te = interpolate_wavlm(np.load(test_wavlm)["wavlm"])
    w_tr_feat, w_te_feat = stack_wavlm_feat(w_tr), stack_wavlm_feat(w_te)

    vq_tr_feat = stack_wavvq_feat(np.load(train_wavvq)["wavvq"])
    vq_te_feat = stack_wavvq_feat(np.load(test_wavvq)["wavvq"])

    ph_tr, ph_te = data["phase"], test_data["phase"]
    ctx_tr, ctx_te = data["context"].squeeze(2), test_data["context"].squeeze(2)

    tr = lambda a: a.transpose((0, 2, 1)) if a.ndim == 3 else np.swapaxes(a, 1, 2)
    return (tr(mfcc), code, tr(t_mfcc), tr(feat), tr(t_feat), tr(w_tr), tr(w_te), tr(w_tr_feat), tr(w_te_feat),
            tr(speech), tr(t_speech), tr(speech_feat), tr(t_speech_feat), tr(vq_tr_feat), tr(vq_te_feat),
            tr(ph_tr), tr(ph_te), tr(ctx_tr), tr(ctx_te))


def load_match_inputs(data_file, codepath, test_data_path, train_wavlm, test_wavlm, train_wavvq, test_wavvq,
                      mode="A", device=None):
    """Lean loader: only what CodeKNN's shipped path reads.  Returns a dict with
    code, phase (as stored), context windows, audio window rows / tokens and the
    per-segment queries.  With `device` the WavLM interpolation + tap stacking of the DATABASE runs on the GPU
    (qpg_stack_wavlm_rows, bit-identical to the host path) and `aud_rows` is a CUDA tensor: the [N*26, 6C] table
    never exists in host memory."""
    from .matchdb import wavvq_tokens

    data = np.load(data_file, allow_pickle=True)
    test_data = np.load(test_data_path, allow_pickle=True)
    code = np.load(codepath)["code"]
    n = code.shape[0]
    ctx = data["context"].squeeze(2)
    t_ctx = test_data["context"].squeeze(2)
    out = dict(code=code, phase=data["phase"], n_test=int(np.load(test_wavvq)["wavvq"].shape[0]),
               txt_rows=np.ascontiguousarray(ctx[:, :WINDOWS_PER_SEQ, :].reshape(n * WINDOWS_PER_SEQ, -1),
                                             dtype=np.float32))
    if mode == "A":
        w_te = interpolate_wavlm(np.load(test_wavlm)["wavlm"])
        if device is not None:
            out["aud_rows"] = wavlm_rows_on_device(np.load(train_wavlm)["wavlm"], "window", device)
        else:
            out["aud_rows"] = wavlm_window_rows(interpolate_wavlm(np.load(train_wavlm)["wavlm"]))
        out["aud_q"] = wavlm_query_rows(w_te)                              # [M, 8, 6C]
        T = w_te.shape[1]
        i_list = list(range(0, T, STEP_SZ * (T // num_frames_code)))
        out["txt_q"] = np.ascontiguousarray(t_ctx[:, [int(i / T * 30) for i in i_list], :], dtype=np.float32)
    else:
        from .matchdb import mode_b_window_frames

        vq_tr = stack_wavvq_feat(np.load(train_wavvq)["wavvq"])
        vq_te = stack_wavvq_feat(np.load(test_wavvq)["wavvq"])
        ks, _ = mode_b_window_frames()
        out["aud_tokens"] = wavvq_tokens(vq_tr[:, ks, :]).reshape(n * WINDOWS_PER_SEQ, -1)
        step = STEP_SZ * (398 / num_frames_code)
        i_list, i = [], 0
        while i < 398:
            i_list.append(i)
            i += step
        out["aud_q"] = wavvq_tokens(vq_te[:, [int(i) for i in i_list], :])   # [M, 8, 11]
        out["txt_q"] = np.ascontiguousarray(t_ctx[:, [int(i / 398 * 30) for i in i_list], :], dtype=np.float32)
    return out


def wavlm_rows_on_device(wavlm, kind: str = "window", device=None):
    """Device-side replacement of interpolate_wavlm + wavlm_window_rows / wavlm_query_rows (row 8(f).1):
    raw WavLM frames [n, 199, C] -> float32 CUDA tensor [n*26, 6C] (kind='window') or [n, 8, 6C] (kind='query'),
    bit-identical to the host path (qpg_stack_wavlm_rows)."""
    from . import _lib

    lib = _lib.load()
    dev = torch.device(device if device is not None else "cuda")
    w = torch.as_tensor(np.ascontiguousarray(wavlm, dtype=np.float32) if isinstance(wavlm, np.ndarray) else wavlm)
    w = w.to(device=dev, dtype=torch.float32).contiguous()
    n, t_in, c = w.shape
    t_out = t_in // num_frames_code * num_frames_code
    step = t_out // num_frames_code
    n_rows, row_step = (WINDOWS_PER_SEQ, step) if kind == "window" else (-(-t_out // (STEP_SZ * step)), STEP_SZ * step)
    out = torch.empty((n * n_rows, NUM_AUDIO_FEAT_FRAMES * c), dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.qpg_stack_wavlm_rows(_lib.ptr(w), n, t_in, c, t_out, n_rows, row_step, _lib.ptr(out),
                                            _lib.stream_ptr()), "qpg_stack_wavlm_rows")
    return out if kind == "window" else out.view(n, n_rows, -1)

"""Device-resident candidate database for the phase-guided matcher.

What the reference keeps as Python/NumPy arrays inside CodeKNN
(GestureKNN.py:423-460) is laid out here once per database for the B200
kernels (DESIGN.md "Data layout in HBM"):

  * audio windows  mode A: float32 [W, 6*C] stacked WavLM features of the 26
                   candidate windows of every sequence (rows k = 0,6,...,150 of
                   wavlm_train_feat, GestureKNN.py:671-690), packed into 4 KiB
                   tiles + float64 squared norms
                   mode B: uint32 [W, 12] vq-wav2vec tokens (GestureKNN.py:58-60)
  * text windows   float32 [W, Dt] context_train[j, m] (GestureKNN.py:712-720)
  * labels         int32 [W] start code code_train[j, m]
  * code           int32 [N, 30];  phase_amp float32 [N, 240, 16]
  * pos_rank       int32 [512, 512]  rank of ||sig[last]-sig[c]|| (:532-540)
  * freq_rank      int32 [512]       rank of 1 - count/total (:481-499, :544)

Window ids are global: id = 26*j + m over the WHOLE database, so that a
row-sharded database (rank r holds sequences [j0, j1)) merges with a plain
lexicographic min and ties keep the reference's scan order.
"""
from __future__ import annotations

from collections import Counter
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _lib
from .constant import (STEP_SZ, WINDOWS_PER_SEQ, WAVVQ_FRAMES, codebook_size, num_frames, num_frames_code)

PAIR_DTYPE = np.dtype([("d", "<f8"), ("id", "<i8")])


# ----------------------------------------------------------------------------
# host-side one-off tables (same NumPy expressions as the reference, so that
# tie order inside argsort is whatever the reference gets on this machine)
# ----------------------------------------------------------------------------
def code_to_freq(train_code: np.ndarray) -> np.ndarray:
    """CodeKNN.code_to_freq (GestureKNN.py:481-499) -> freq_dist_cands [512]."""
    code = np.asarray(train_code).flatten()
    result = Counter(code.tolist())
    result_sorted = sorted(result.items(), key=lambda item: item[1], reverse=True)
    x = [d[0] for d in result_sorted]
    y = 1 - np.array([d[1] for d in result_sorted]) / sum(d[1] for d in result_sorted)
    pos = {c: i for i, c in enumerate(x)}
    return np.array([y[pos[i]] if i in pos else 1 for i in range(codebook_size)], dtype=np.float64)


def freq_rank_from_code(train_code: np.ndarray) -> np.ndarray:
    """np.array(freq_dist_cands).argsort().argsort() (GestureKNN.py:544)."""
    return np.array(list(code_to_freq(train_code))).argsort().argsort().astype(np.int32)


def pos_rank_table(signature: np.ndarray) -> np.ndarray:
    """pos_score for every possible last code (GestureKNN.py:532-540):
    row `last` = argsort(argsort([norm(sig[last]-sig[c]) or inf if c==last]))."""
    sig = np.asarray(signature)
    out = np.empty((codebook_size, codebook_size), dtype=np.int32)
    for last in range(codebook_size):
        pos = []
        s_last = sig[last]
        for c in range(codebook_size):
            if c == last:
                pos.append(1e10000)
                continue
            pos.append(np.linalg.norm(s_last - sig[c]))
        out[last] = np.array(pos).argsort().argsort()
    return out


def phase_to_dense(phase) -> np.ndarray:
    """phase column of the train npz -> float32 [N, 240, 16] = phase | amplitude
    (columns 0 and 2 of the (p, f, a, b) tuple, GestureKNN.py:633-635).  Accepts the
    reference's object array [N,240,4] of (1,8,1) tensors, or dense [N,240,4,8]."""
    if isinstance(phase, np.ndarray) and phase.dtype != object:
        ph = np.asarray(phase, dtype=np.float32)
        assert ph.ndim == 4 and ph.shape[2] == 4, "dense phase must be [N,240,4,8]"
        return np.ascontiguousarray(np.concatenate((ph[:, :, 0, :], ph[:, :, 2, :]), axis=2))
    n, t = phase.shape[0], phase.shape[1]
    # the reference pickles one (1, 8, 1) tensor per (sequence, frame, column): two torch.cat calls over the object
    # cells instead of n*t*2 tensor -> numpy round trips
    cols = []
    for col in (0, 2):
        cells = list(phase[:, :, col].reshape(-1))
        if cells and isinstance(cells[0], torch.Tensor):
            flat = torch.cat([c.detach().reshape(1, -1) for c in cells], dim=0).to(torch.float32).cpu().numpy()
        else:
            flat = np.stack([np.asarray(c, dtype=np.float32).reshape(-1) for c in cells]) if cells else np.zeros((0, 8), np.float32)
        cols.append(flat.reshape(n, t, -1))
    return np.ascontiguousarray(np.concatenate(cols, axis=2), dtype=np.float32)


def mode_b_window_frames(n_db_frm: int = WAVVQ_FRAMES):
    """int(k) and int(k/step_sz) of the float loop GestureKNN.py:671-690 (mode B)."""
    step = n_db_frm / num_frames_code
    ks, ms = [], []
    k = 0
    while k < n_db_frm - STEP_SZ * step:
        ks.append(int(k))
        ms.append(int(k / step))
        k += step
    return ks, ms


def wavvq_tokens(feat22: np.ndarray) -> np.ndarray:
    """[..., 22] stacked wavvq feature -> [..., 11] tokens g0*320+g1 (GestureKNN.py:58-60)."""
    f = np.asarray(feat22).reshape(feat22.shape[:-1] + (-1, 2))
    return (f[..., 0] * 320 + f[..., 1]).astype(np.int64)


def pad_tokens(tok: np.ndarray) -> np.ndarray:
    out = np.zeros(tok.shape[:-1] + (12,), dtype=np.uint32)
    out[..., :11] = tok.astype(np.uint32)
    return out


def phase_frame(k) -> int:
    return int(k / 398 * 240)  # GestureKNN.py:632,640 (398 is used in WavLM mode too)


# ----------------------------------------------------------------------------
@dataclass
class PackedRows:
    """float32 row table in the tile layout of qpg_pack_rows_f32."""
    packed: torch.Tensor      # float32 [G*NC*1024]
    sqnorm: torch.Tensor      # float64 [W]
    W: int
    D: int

    @staticmethod
    def from_rows(rows: torch.Tensor) -> "PackedRows":
        lib = _lib.load()
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2
        rows = rows.contiguous()
        W, D = rows.shape
        nbytes = lib.qpg_packed_bytes(W, D)
        packed = torch.empty(max(nbytes // 4, 1), dtype=torch.float32, device=rows.device)
        sqnorm = torch.empty(max(W, 1), dtype=torch.float64, device=rows.device)
        with torch.cuda.device(rows.device):
            _lib.check(lib.qpg_pack_rows_f32(_lib.ptr(rows), W, D, _lib.ptr(packed), _lib.ptr(sqnorm),
                                             _lib.stream_ptr()), "qpg_pack_rows_f32")
        return PackedRows(packed, sqnorm[:W], W, D)

    @property
    def nbytes(self) -> int:
        return self.packed.numel() * 4


def aligned_bytes(nbytes: int, device, align: int = 1024) -> torch.Tensor:
    """uint8 device buffer whose data pointer is `align`-byte aligned (the UMMA tile images need 1024)."""
    raw = torch.empty(max(int(nbytes), 1) + align, dtype=torch.uint8, device=device)
    off = (-raw.data_ptr()) % align
    return raw[off:off + max(int(nbytes), 1)]


def bin_order(labels: torch.Tensor):
    """Rows ordered by start code (stable, so ids ascend inside a bin); labels outside [0,512) go last.
    -> (order int32 [W]: sorted position -> source row, bin_start int32 [513])"""
    key = torch.where((labels >= 0) & (labels < codebook_size), labels, torch.full_like(labels, codebook_size))
    skey, order = torch.sort(key.to(torch.int64), stable=True)
    edges = torch.arange(codebook_size + 1, device=labels.device, dtype=torch.int64)
    bin_start = torch.searchsorted(skey, edges)
    return order.to(torch.int32).contiguous(), bin_start.to(torch.int32).contiguous()


def column_exponents(rows: torch.Tensor):
    """Per-column power-of-two scaling for the fixed-point copy: None when the columns already share a
    magnitude (binary exponents of the column maxima within 2 of each other), else int8 [D] =
    floor(log2(max |x[:, k]|)) - median."""
    colmax = torch.maximum(rows.amax(dim=0), -rows.amin(dim=0)).to(torch.float64)      # no |rows| temporary
    ok = colmax > 0
    if not bool(ok.any()):
        return None
    e = torch.floor(torch.log2(torch.where(ok, colmax, torch.ones_like(colmax))))
    med = e[ok].median()
    e = torch.where(ok, e - med, torch.zeros_like(e))
    if float(e.max() - e.min()) < 3:
        return None
    return e.clamp(-60, 60).to(torch.int8).contiguous()


@dataclass
class SlicedRows:
    """int8-sliced fixed-point copy of a float32 row table in bin order (csrc/sliced_scan.cu)."""
    slices: torch.Tensor      # uint8, 1024-byte aligned tile images
    row_info: torch.Tensor    # float64 [W, 2]
    order: torch.Tensor       # int32 [W]
    bin_start: torch.Tensor   # int32 [513]
    col_exp: Optional[torch.Tensor]
    W: int
    D: int

    @property
    def n_kblocks(self) -> int:
        return -(-self.D // 128)

    @property
    def Wpad(self) -> int:
        return -(-self.W // 128) * 128

    @staticmethod
    def from_rows(rows: torch.Tensor, sqnorm: torch.Tensor, order: torch.Tensor, bin_start: torch.Tensor,
                  column_scaling: bool = True) -> "SlicedRows":
        """rows float32 [W, D] on the device (this shard, source order); sqnorm float64 [W] by source row."""
        lib = _lib.load()
        assert rows.is_cuda and rows.dtype == torch.float32 and rows.dim() == 2
        rows = rows.contiguous()
        W, D = rows.shape
        dev = rows.device
        col_exp = column_exponents(rows) if column_scaling and W > 0 else None
        slices = aligned_bytes(lib.qpg_sliced_bytes(W, D), dev)
        row_info = torch.zeros((max(W, 1), 2), dtype=torch.float64, device=dev)
        with torch.cuda.device(dev):
            _lib.check(lib.qpg_slice_rows_i8(_lib.ptr(rows), W, D, _lib.ptr(order), _lib.ptr(col_exp),
                                             _lib.ptr(sqnorm.contiguous()), _lib.ptr(slices), _lib.ptr(row_info),
                                             _lib.stream_ptr()), "qpg_slice_rows_i8")
            torch.cuda.current_stream().synchronize()
        return SlicedRows(slices, row_info, order, bin_start, col_exp, W, D)

    @property
    def nbytes(self) -> int:
        return self.slices.numel()


class MatchDatabase:
    """Everything the matcher kernels read, resident on one GPU."""

    def __init__(self, mode: str, code: np.ndarray, signature: np.ndarray, phase_amp: np.ndarray,
                 txt_rows: np.ndarray, aud_rows: Optional[np.ndarray] = None,
                 aud_tokens: Optional[np.ndarray] = None, freq_code: Optional[np.ndarray] = None,
                 freq_rank: Optional[np.ndarray] = None, pos_rank: Optional[np.ndarray] = None,
                 device=None, seq_range=None, fuse_text=False, sliced=True, replicate_exact=False):
        """aud_rows [N*26, Da] float32 (mode A) or aud_tokens [N*26, 11] ints (mode B);
        txt_rows [N*26, Dt] float32.  `seq_range=(j0, j1)` keeps only the windows of
        sequences j0..j1-1 on this GPU (row shard); code / phase tables stay whole.
        `sliced`: build the int8-sliced copies the one-pass scan streams (mode A).
        `replicate_exact`: with a row shard, keep the float32 tables of ALL rows on this GPU (they are only
        touched to re-evaluate undecided candidates) while the scanned, sliced copy holds the shard - every
        rank can then settle cross-shard decisions itself after ONE all-gather."""
        assert mode in ("A", "B")
        _lib.load()
        self.mode = mode
        self.device = torch.device(device if device is not None else "cuda")
        code = np.asarray(code)
        self.n_seq = int(code.shape[0])
        self.code_host = code.astype(np.int64)
        self.signature = np.asarray(signature)
        j0, j1 = (0, self.n_seq) if seq_range is None else seq_range
        self.seq_range = (int(j0), int(j1))
        self.id_offset = int(j0) * WINDOWS_PER_SEQ
        w0, w1 = j0 * WINDOWS_PER_SEQ, j1 * WINDOWS_PER_SEQ
        self.W = w1 - w0
        dev = self.device
        self.replicated = bool(replicate_exact) and seq_range is not None
        x0, x1 = (0, self.n_seq * WINDOWS_PER_SEQ) if self.replicated else (w0, w1)
        self.exact_offset = x0               # global id of row 0 of the float32 tables
        self.row_base = w0 - x0              # shard row r is row r + row_base of the float32 tables

        labels = code[:, :WINDOWS_PER_SEQ].reshape(-1).astype(np.int32)          # code_train[j, m]
        self.labels = torch.from_numpy(np.ascontiguousarray(labels[x0:x1])).to(dev)   # rows of the float32 tables
        self.code = torch.from_numpy(code.astype(np.int32)).contiguous().to(dev)
        self.phase_amp_host = np.ascontiguousarray(phase_amp, dtype=np.float32)
        assert self.phase_amp_host.shape == (self.n_seq, num_frames, 16)
        self.phase_amp = torch.from_numpy(self.phase_amp_host).to(dev)

        def local_rows(rows):
            """host array of ALL windows (cut to [x0, x1) here) or a CUDA tensor holding exactly those rows"""
            if isinstance(rows, torch.Tensor):
                assert rows.is_cuda and rows.shape[0] == x1 - x0, "device rows must be the rows this rank keeps"
                return rows.to(device=dev, dtype=torch.float32)
            return torch.from_numpy(np.ascontiguousarray(np.asarray(rows, dtype=np.float32)[x0:x1])).to(dev)

        shard = slice(self.row_base, self.row_base + self.W)
        self.aud_s = self.txt_s = None
        if sliced and mode == "A":
            self.order, self.bin_start = bin_order(self.labels[shard])

        txt_dev = local_rows(txt_rows)
        self.txt = PackedRows.from_rows(txt_dev)
        if sliced and mode == "A":
            self.txt_s = SlicedRows.from_rows(txt_dev[shard], self.txt.sqnorm[shard], self.order, self.bin_start)
        self.aud = None
        self.tokens = None
        self.fused = None
        if mode == "A":
            aud_dev = local_rows(aud_rows)
            self.aud = PackedRows.from_rows(aud_dev)
            if sliced:
                self.aud_s = SlicedRows.from_rows(aud_dev[shard], self.aud.sqnorm[shard], self.order, self.bin_start)
            # audio | text in one table for the fused single-pass scan (qpg_cand_cosine2_minbycode)
            if fuse_text and aud_dev.shape[1] % 128 == 0 and txt_dev.shape[1] % 128 == 0 and \
                    self.W * 4 * (aud_dev.shape[1] + txt_dev.shape[1]) <= (8 << 30):
                self.fused = PackedRows.from_rows(torch.cat((aud_dev, txt_dev), dim=1))
            del aud_dev
            self.aud_k = [m * 6 for m in range(WINDOWS_PER_SEQ)]                    # k = 0,6,...,150
            self.n_db_frm, self.step_sz = 180, 6
        else:
            tok = pad_tokens(np.asarray(aud_tokens)[w0:w1])
            self.tokens = torch.from_numpy(tok.view(np.int32)).to(dev)             # bit pattern of uint32
            ks, ms = mode_b_window_frames()
            assert len(ks) == WINDOWS_PER_SEQ and ms == list(range(WINDOWS_PER_SEQ))
            self.aud_k = ks
            self.n_db_frm, self.step_sz = WAVVQ_FRAMES, WAVVQ_FRAMES / num_frames_code
        del txt_dev
        self.txt_k = [m * 8 for m in range(WINDOWS_PER_SEQ)]                        # k = 0,8,...,200
        self.aud_frame = torch.tensor([phase_frame(k) for k in self.aud_k], dtype=torch.int32, device=dev)
        self.txt_frame = torch.tensor([phase_frame(k) for k in self.txt_k], dtype=torch.int32, device=dev)

        if freq_rank is None:
            freq_rank = freq_rank_from_code(code if freq_code is None else freq_code)
        if pos_rank is None:
            pos_rank = pos_rank_table(self.signature)
        self.freq_rank_host = np.asarray(freq_rank, dtype=np.int32)
        self.pos_rank_host = np.asarray(pos_rank, dtype=np.int32)
        self.freq_rank = torch.from_numpy(self.freq_rank_host).to(dev)
        self.pos_rank = torch.from_numpy(np.ascontiguousarray(self.pos_rank_host)).to(dev)
        # transposed int16 copy for qpg_match_lookup: [c][last], coalesced over `last`
        self.pos_rank_t = torch.from_numpy(np.ascontiguousarray(self.pos_rank_host.T.astype(np.int16))).to(dev)
        torch.cuda.synchronize(dev)

    # -- bytes one query pass reads (the roofline's algorithmic bytes use D, not the padded D)
    def algorithmic_bytes(self, which: str) -> int:
        t = self.aud if which == "audio" else self.txt
        return self.W * (4 * t.D + 4)

    @property
    def n_exact_rows(self) -> int:
        return self.aud.W if self.aud is not None else self.txt.W

    def payload(self, w: int) -> np.ndarray:
        j, m = divmod(int(w), WINDOWS_PER_SEQ)
        return self.code_host[j, m:m + STEP_SZ]

    def aux(self, w: int, which: str):
        j, m = divmod(int(w), WINDOWS_PER_SEQ)
        return [j, int((self.aud_k if which == "audio" else self.txt_k)[m])]


_PACKED_VERSION = 2


def ensure_phase_stats(db):
    """Per-(window, table) phase statistics for the transition kernel (csrc/match_walk.cu, qpg_phase_stats): the
    squared norms and self-dots of the head / tail rows of the window's phase frame and the four rows of the cross
    term, 72 floats each.  Built once per database on the device (7.7 MB for the speaker-10 table)."""
    if getattr(db, "phase_stats", None) is None:
        lib = _lib.load()
        n = int(db.phase_amp.shape[0])
        stats = torch.empty((n * WINDOWS_PER_SEQ * 2 * int(lib.qpg_phase_stats_floats()),), dtype=torch.float32,
                            device=db.phase_amp.device)
        _lib.check(lib.qpg_phase_stats(_lib.ptr(db.phase_amp), n, _lib.ptr(db.aud_frame), _lib.ptr(db.txt_frame),
                                       _lib.ptr(stats), _lib.stream_ptr()), "qpg_phase_stats")
        db.phase_stats = stats
    return db.phase_stats


def save_packed_db(path: str, mode: str, code, signature, phase_amp, txt_rows, aud_rows=None, aud_tokens=None,
                   freq_code=None, device=None) -> None:
    """Row 8(f).1: one-file database for the matcher.  Everything load_db_codebook + CodeKNN.__init__ derive at
    start-up in the reference is built ONCE on the GPU and stored in its device layout: the float32 4 KiB tiles
    + squared norms, the int8-sliced tile images + per-row bound terms + bin order, dense phase|amplitude, the
    frequency and pose rank tables.  `load_packed_db` then needs one host->device copy per array and no kernel."""
    db = MatchDatabase(mode, code, signature, phase_amp, txt_rows, aud_rows=aud_rows, aud_tokens=aud_tokens,
                       freq_code=freq_code, device=device)
    rec = dict(version=np.array(_PACKED_VERSION), mode=np.array(mode), code=db.code_host,
               signature=np.asarray(signature, dtype=np.float32), phase_amp=db.phase_amp_host,
               freq_rank=db.freq_rank_host, pos_rank=db.pos_rank_host, labels=db.labels.cpu().numpy())

    def put_packed(name, t: PackedRows):
        rec[name + "_packed"], rec[name + "_sqnorm"] = t.packed.cpu().numpy(), t.sqnorm.cpu().numpy()
        rec[name + "_shape"] = np.array([t.W, t.D], dtype=np.int64)

    def put_sliced(name, t: SlicedRows):
        rec[name + "_slices"], rec[name + "_row_info"] = t.slices.cpu().numpy(), t.row_info.cpu().numpy()
        rec[name + "_col_exp"] = np.zeros((0,), np.int8) if t.col_exp is None else t.col_exp.cpu().numpy()
    put_packed("txt", db.txt)
    if mode == "A":
        put_packed("aud", db.aud)
        put_sliced("txt_s", db.txt_s)
        put_sliced("aud_s", db.aud_s)
        rec["order"], rec["bin_start"] = db.order.cpu().numpy(), db.bin_start.cpu().numpy()
    else:
        rec["tokens"] = db.tokens.cpu().numpy()
    np.savez(path, **rec)


def load_packed_db(path: str, device=None) -> "MatchDatabase":
    """Counterpart of save_packed_db: uploads the stored device images; no packing / slicing / sorting runs."""
    _lib.load()
    z = np.load(path)
    if int(z["version"]) != _PACKED_VERSION:
        raise ValueError(f"{path}: packed database version {int(z['version'])}, expected {_PACKED_VERSION}")
    dev = torch.device(device if device is not None else "cuda")
    db = object.__new__(MatchDatabase)
    db.mode, db.device = str(z["mode"]), dev
    code = z["code"]
    db.n_seq, db.code_host, db.signature = int(code.shape[0]), code.astype(np.int64), z["signature"]
    db.seq_range, db.id_offset, db.W = (0, db.n_seq), 0, db.n_seq * WINDOWS_PER_SEQ
    db.replicated, db.exact_offset, db.row_base = False, 0, 0
    up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    db.labels, db.code = up(z["labels"]), up(code.astype(np.int32))
    db.phase_amp_host = np.ascontiguousarray(z["phase_amp"], dtype=np.float32)
    db.phase_amp = up(db.phase_amp_host)

    def get_packed(name):
        W, D = (int(v) for v in z[name + "_shape"])
        return PackedRows(up(z[name + "_packed"]), up(z[name + "_sqnorm"]), W, D)

    def get_sliced(name, D):
        img = z[name + "_slices"]
        buf = aligned_bytes(img.size, dev)
        buf.copy_(torch.from_numpy(img))
        ce = z[name + "_col_exp"]
        return SlicedRows(buf, up(z[name + "_row_info"]), db.order, db.bin_start, up(ce) if ce.size else None, db.W, D)
    db.txt = get_packed("txt")
    db.aud = db.tokens = db.fused = db.aud_s = db.txt_s = None
    if db.mode == "A":
        db.aud = get_packed("aud")
        db.order, db.bin_start = up(z["order"]), up(z["bin_start"])
        db.txt_s, db.aud_s = get_sliced("txt_s", db.txt.D), get_sliced("aud_s", db.aud.D)
        db.aud_k = [m * 6 for m in range(WINDOWS_PER_SEQ)]
        db.n_db_frm, db.step_sz = 180, 6
    else:
        db.tokens = up(z["tokens"])
        db.aud_k, _ = mode_b_window_frames()
        db.n_db_frm, db.step_sz = WAVVQ_FRAMES, WAVVQ_FRAMES / num_frames_code
    db.txt_k = [m * 8 for m in range(WINDOWS_PER_SEQ)]
    db.aud_frame = torch.tensor([phase_frame(k) for k in db.aud_k], dtype=torch.int32, device=dev)
    db.txt_frame = torch.tensor([phase_frame(k) for k in db.txt_k], dtype=torch.int32, device=dev)
    db.freq_rank_host, db.pos_rank_host = z["freq_rank"].astype(np.int32), z["pos_rank"].astype(np.int32)
    db.freq_rank, db.pos_rank = up(db.freq_rank_host), up(db.pos_rank_host)
    db.pos_rank_t = up(db.pos_rank_host.T.astype(np.int16))
    torch.cuda.synchronize(dev)
    return db


def new_table(Q: int, device) -> torch.Tensor:
    """Uninitialised [Q, 512] table of qpg_pair_t (stored as int64 [Q,512,2])."""
    return torch.empty((Q, codebook_size, 2), dtype=torch.int64, device=device)


def table_to_numpy(table: torch.Tensor) -> np.ndarray:
    """-> structured array [Q, 512] with fields d (float64) and id (int64)."""
    return table.cpu().numpy().view(PAIR_DTYPE).reshape(table.shape[0], codebook_size)

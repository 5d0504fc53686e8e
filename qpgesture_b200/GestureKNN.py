"""B200 drop-in for the CodeKNN half of the reference's
codebook/Speech2GestureMatching/GestureKNN.py (:44-67, :422-845).

Same public names and argument meaning -- `wavvq_distances`, `CodeKNN`
(`search_code_knn`, `search_audio_cands`, `search_text_cands`,
`init_code_phase`, `code_to_signature`, `code_to_freq`),
`predict_code_from_audio`, `main_codebook`, and the CLI flags of
GestureKNN.sh -- but the candidate scans run as CUDA kernels of
libqpg_sm100.so over a database that stays resident in HBM, and all steps of
all segments are evaluated in one pass (they do not depend on the sequential
state).  There is no CPU fallback for the scans.

Unlike the reference, nothing is parsed at import: `args` is filled by
`main()` (or by `set_args`).
"""
from __future__ import annotations

import argparse
import os
import random
from types import SimpleNamespace
from typing import Optional

import numpy as np
import torch

from . import _lib
from .constant import (NUM_AUDIO_FEAT_FRAMES, SEED_VALUE, STEP_SZ, STEPS_PER_SEGMENT, WINDOWS_PER_SEQ, codebook_size,
                       num_frames, num_frames_code)
from .matchdb import (MatchDatabase, ensure_phase_stats, new_table, pad_tokens, phase_to_dense, table_to_numpy, wavvq_tokens)

args = None  # module-global like the reference's (GestureKNN.py:41); set by main()/set_args()


def build_parser() -> argparse.ArgumentParser:
    """Flags of GestureKNN.py:25-39 (GestureKNN.sh:7-18 passes 11 of them)."""
    p = argparse.ArgumentParser()
    d = "/path/to/training_db_data.npz"
    p.add_argument("-d", "--train_database", type=str, default=d)
    p.add_argument("-c", "--train_codebook", type=str, default=d)
    p.add_argument("-w", "--train_wavlm", type=str, default=d)
    p.add_argument("-wvq", "--train_wavvq", type=str, default=d)
    p.add_argument("-s", "--codebook_signature", type=str, default=d)
    p.add_argument("-e", "--test_data", type=str, default="/path/to/test_data.npz")
    p.add_argument("-tw", "--test_wavlm", type=str, default=d)
    p.add_argument("-twvq", "--test_wavvq", type=str, default=d)
    p.add_argument("-om", "--out_knn_filename", type=str, default="/path/to/knn_pred.npz")
    p.add_argument("-ov", "--out_video_path", type=str, default="/path/to/video/")
    p.add_argument("-k", "--desired_k", type=int, default=0)
    p.add_argument("-f", "--fake", type=bool, default=False)
    p.add_argument("-of", "--out_fake_knn_filename", type=str, default="/path/to/knn_pred.npz")
    p.add_argument("--max_frames", type=int, default=0)
    # additions (not in the reference)
    p.add_argument("--mode", choices=("A", "B"), default="A",
                   help="A = WavLM cosine (the literals shipped at GestureKNN.py:842-843), B = wavvq Levenshtein")
    p.add_argument("--tail", choices=("auto", "device", "numpy"), default="auto",
                   help="auto = device tail, clips whose result hinged on the order of exact ties are redone with the "
                        "reference's own NumPy argsort calls; device = stable order everywhere; numpy = NumPy tail")
    p.add_argument("--gpu", type=int, default=0)
    return p


def set_args(ns):
    global args
    args = ns


def seed_everything(seed_value: int = SEED_VALUE):
    """GestureKNN.py:19-22."""
    os.environ["PYTHONHASHSEED"] = str(seed_value)
    random.seed(seed_value)
    np.random.seed(seed_value)


# --------------------------------------------------------------------------
def _lev_pairs(a_tok: np.ndarray, b_tok: np.ndarray) -> np.ndarray:
    lib = _lib.load()
    a = torch.from_numpy(pad_tokens(a_tok).view(np.int32)).cuda()
    b = torch.from_numpy(pad_tokens(b_tok).view(np.int32)).cuda()
    out = torch.empty(a.shape[0], dtype=torch.int32, device=a.device)
    _lib.check(lib.qpg_lev_distance(_lib.ptr(a), _lib.ptr(b), a.shape[0], _lib.ptr(out), _lib.stream_ptr()),
               "qpg_lev_distance")
    return out.cpu().numpy()


def wavvq_distances(ls1, ls2, mode="sum"):
    """GestureKNN.py:44-67.  'combine': 11 tokens g0*320+g1, one edit distance.
    'sum': the two code groups as separate strings, distances added (the
    reference's reshape(NUM_AUDIO_FEAT_FRAMES, -1) only accepts 12 values)."""
    ls1, ls2 = np.asarray(ls1), np.asarray(ls2)
    if mode == "combine":
        return int(_lev_pairs(wavvq_tokens(ls1)[None, :], wavvq_tokens(ls2)[None, :])[0])
    if mode == "sum":
        a = ls1.reshape(NUM_AUDIO_FEAT_FRAMES, -1).transpose()
        b = ls2.reshape(NUM_AUDIO_FEAT_FRAMES, -1).transpose()
        pad = np.full((2, 11 - a.shape[1]), 0xFFFFFFFF, dtype=np.int64)  # common suffix: distance unchanged
        d = _lev_pairs(np.concatenate((a[:2].astype(np.int64), pad), 1),
                       np.concatenate((b[:2].astype(np.int64), pad), 1))
        return int(d[0] + d[1])
    return None


# --------------------------------------------------------------------------
class CodeKNN(object):
    """Reference constructor signature (GestureKNN.py:423-425) plus keyword-only
    extras.  Arrays are (N, time, feat) as in the reference at this point."""

    def __init__(self, mfcc_train=None, code_train=None, feat_train=None, wavlm_train=None, wavlm_train_feat=None,
                 speech_features=None, speech_features_feat=None, wavvq_train_feat=None, phase_train=None,
                 context_train=None, use_wavlm=False, use_wavvq=False, use_phase=False, use_txt=False, *,
                 codebook_signature=None, train_codebook=None, device=None, database: Optional[MatchDatabase] = None,
                 seq_range=None, process_group=None, tail="auto", aud_rows=None, aud_tokens=None):
        super().__init__()
        self.use_phase = use_phase
        self.phase_channels = 8
        self.tail = tail
        self.process_group = process_group
        self.code_train = None if code_train is None else np.asarray(code_train)
        self.c2s, self.c2f, self.freq_dist_cands = None, None, None
        if database is not None:
            self.db = database
            self.code_train = database.code_host
        else:
            if not (use_wavlm or use_wavvq):
                raise NotImplementedError("only the WavLM (use_wavlm) and wavvq (use_wavvq) matchers are built; the "
                                          "MFCC branch of CodeKNN is not on the shipped path (GestureKNN.py:842)")
            mode = "A" if use_wavlm else "B"
            sig_path = codebook_signature or (args.codebook_signature if args is not None else None)
            code_path = train_codebook or (args.train_codebook if args is not None else None)
            if sig_path is None:
                raise ValueError("codebook_signature path needed (reference reads args.codebook_signature, :476)")
            signature = np.load(sig_path)["signature"] if isinstance(sig_path, str) else np.asarray(sig_path)
            freq_code = self.code_train
            if code_path is not None:
                freq_code = np.load(code_path)["code"] if isinstance(code_path, str) else np.asarray(code_path)
            n = self.code_train.shape[0]
            ctx = np.asarray(context_train)
            txt_rows = ctx[:, :WINDOWS_PER_SEQ, :].reshape(n * WINDOWS_PER_SEQ, -1)
            if mode == "A" and aud_rows is None:
                f = np.asarray(wavlm_train_feat)                      # (N, 180, 6C)
                step = f.shape[1] // num_frames_code
                aud_rows = f[:, 0:step * WINDOWS_PER_SEQ:step, :].reshape(n * WINDOWS_PER_SEQ, -1)
            if mode == "B" and aud_tokens is None:
                from .matchdb import mode_b_window_frames
                ks, _ = mode_b_window_frames()
                aud_tokens = wavvq_tokens(np.asarray(wavvq_train_feat)[:, ks, :]).reshape(n * WINDOWS_PER_SEQ, -1)
            self.db = MatchDatabase(mode, self.code_train, signature, phase_to_dense(_phase_ntc(phase_train)),
                                    txt_rows, aud_rows=aud_rows, aud_tokens=aud_tokens, freq_code=freq_code,
                                    device=device, seq_range=seq_range)
        self.mode = self.db.mode
        self.step_sz = self.db.step_sz
        self.n_db_seq = self.db.n_seq
        self.n_db_frm = self.db.n_db_frm
        self.code_to_signature()
        self.code_to_freq()

    # ---- reference helpers ---------------------------------------------------
    def code_to_signature(self):
        self.c2s = {i: self.db.signature[i] for i in range(codebook_size)}       # GestureKNN.py:475-479

    def code_to_freq(self):
        """GestureKNN.py:481-499; the rank transform of it (:544) lives in self.db.freq_rank."""
        from .matchdb import code_to_freq
        self.freq_dist_cands = list(code_to_freq(self.code_train))
        self.c2f = dict(enumerate(self.freq_dist_cands))

    def init_code_phase(self):
        """GestureKNN.py:462-473: two draws from NumPy's global legacy RNG."""
        init_i = np.random.randint(0, self.n_db_seq)
        init_j = np.random.randint(0, self.n_db_frm - int(num_frames / num_frames_code))
        init_code = self.code_train[init_i, init_j // num_frames_code]
        if not self.use_phase:
            return init_code
        return init_code, self.db.phase_amp_host[init_i, init_j:init_j + int(num_frames / num_frames_code)]

    # ---- device scans ----------------------------------------------------------
    def _scan(self, which: str, q: torch.Tensor, table: torch.Tensor, stream=None):
        lib, db = _lib.load(), self.db
        Q = q.shape[0]
        sp = _lib.stream_ptr(stream)
        _lib.check(lib.qpg_table_init(_lib.ptr(table), Q * codebook_size, sp), "qpg_table_init")
        if which == "text" or db.mode == "A":
            t = db.txt if which == "text" else db.aud
            assert q.dtype == torch.float32 and q.shape[1] == t.D, (q.shape, t.D)
            _lib.check(lib.qpg_cand_cosine_minbycode_team(_lib.ptr(t.packed), _lib.ptr(t.sqnorm), _lib.ptr(db.labels),
                                                          t.W, t.D, db.exact_offset, _lib.ptr(q), Q, _lib.ptr(table), 0,
                                                          self._team_size(t.D), sp),
                       "qpg_cand_cosine_minbycode_team")
        else:
            assert q.dtype == torch.int32 and q.shape[1] == 12
            _lib.check(lib.qpg_cand_lev_minbycode(_lib.ptr(db.tokens), _lib.ptr(db.labels), db.W, db.id_offset,
                                                  _lib.ptr(q), Q, _lib.ptr(table), sp), "qpg_cand_lev_minbycode")
        if self._sharded_exact():
            table = self._merge_shards(table, stream)
        return table

    def _sharded_exact(self) -> bool:
        """True when the float32 tables hold only this rank's rows, i.e. exact scans need the cross-rank merge."""
        return self.process_group is not None and not self.db.replicated

    def _team_size(self, D: int) -> int:
        """Team size S for the cosine scan.  Single GPU: 0 (library picks).  Row-sharded: every rank must
        use the same S (it fixes the float64 summation order, see qpg_cand_cosine_minbycode_team), so it is
        derived from the largest shard of the WHOLE database, not from this rank's row count."""
        if not self._sharded_exact():
            return 0
        import torch.distributed as dist

        world = dist.get_world_size(self.process_group)
        max_seq = -(-self.db.n_seq // world)
        groups = -(-(max_seq * WINDOWS_PER_SEQ) // 8)
        n_chunks = -(-D // 128)
        sms = torch.cuda.get_device_properties(self.db.device).multi_processor_count
        best, S = None, 1
        for cand in (1, 2, 3, 4, 6):                      # same cost model as the library's automatic choice
            if cand > n_chunks:
                continue
            teams = sms * (12 // cand)
            cost = -(-groups // teams) * -(-n_chunks // cand)
            if best is None or cost < best:
                best, S = cost, cand
        return S

    def _merge_shards(self, table, stream=None):
        import torch.distributed as dist

        lib = _lib.load()
        world = dist.get_world_size(self.process_group)
        parts = torch.empty((world,) + tuple(table.shape), dtype=table.dtype, device=table.device)
        dist.all_gather_into_tensor(parts, table, group=self.process_group)
        out = torch.empty_like(table)
        _lib.check(lib.qpg_table_merge(_lib.ptr(parts), world, table.shape[0] * codebook_size, _lib.ptr(out),
                                       _lib.stream_ptr(stream)), "qpg_table_merge")
        return out

    def _audio_query_tensor(self, aud_q: np.ndarray) -> torch.Tensor:
        if self.db.mode == "A":
            return torch.from_numpy(np.ascontiguousarray(aud_q, dtype=np.float32))
        a = np.asarray(aud_q)
        tok = wavvq_tokens(a) if a.shape[-1] == 22 else a
        return torch.from_numpy(pad_tokens(tok).view(np.int32))

    def match_tables(self, aud_q, txt_q):
        """State-independent part for Q query steps: device tables [Q,512] of
        (best distance, best global window id) for audio and text."""
        dev = self.db.device
        with torch.cuda.device(dev):
            qa = self._audio_query_tensor(aud_q).to(dev, non_blocking=True)
            qt = torch.from_numpy(np.ascontiguousarray(txt_q, dtype=np.float32)).to(dev, non_blocking=True)
            ta = self._scan("audio", qa, new_table(qa.shape[0], dev))
            tt = self._scan("text", qt, new_table(qt.shape[0], dev))
        return ta, tt

    def _lists_from_table(self, row: np.ndarray, which: str):
        dist = [float(x) for x in row["d"]]
        index, aux = [], []
        for w in row["id"]:
            if w < 0:
                index.append([])
                aux.append([])
            else:
                index.append(self.db.payload(w))
                aux.append(self.db.aux(w, which))
        return dist, index, aux

    def search_audio_cands(self, clip_input, mode="audio"):
        """GestureKNN.py:666-691 for mode 'wavlm_feat' / 'wavvq_feat'."""
        want = {"A": "wavlm_feat", "B": "wavvq_feat"}[self.db.mode]
        if mode != want:
            raise NotImplementedError(f"database was built for mode {want!r}, got {mode!r}")
        dev = self.db.device
        with torch.cuda.device(dev):
            q = self._audio_query_tensor(np.asarray(clip_input)[None, :]).to(dev)
            table = self._scan("audio", q, new_table(1, dev))
        return self._lists_from_table(table_to_numpy(table)[0], "audio")

    def search_text_cands(self, clip_input, mode="wavvq_feat"):
        """GestureKNN.py:708-721."""
        dev = self.db.device
        with torch.cuda.device(dev):
            q = torch.from_numpy(np.ascontiguousarray(np.asarray(clip_input)[None, :], dtype=np.float32)).to(dev)
            table = self._scan("text", q, new_table(1, dev))
        return self._lists_from_table(table_to_numpy(table)[0], "text")

    # ---- preallocated plan: the whole step as a fixed launch sequence (optionally one CUDA graph) ------
    def make_plan(self, n_clips: int, n_seg: int, tail_clips=None, use_graph: bool = True, want_phase=False,
                  engine: Optional[str] = None, scan_priority: bool = False):
        """Static device buffers for `n_clips` clips x `n_seg` segments.  `tail_clips` = slice of the
        clips whose sequential tail this rank runs (default: all).  Fill plan.qa / plan.qt /
        plan.seed_code / plan.seed_phase, then call run_plan(plan); results land in plan.codes / plan.status.

        engine "sliced" (default in mode A): every query step of the batch in ONE pass over the int8-sliced
        table per 64 steps (tcgen05 kind::i8) + interval logic + float64 re-evaluation of undecided bins;
        engine "f64": the float64 streaming scans of round 1 (4 steps per pass; the only engine of mode B).
        `scan_priority`: launch the HBM-bound scan on a high-priority side stream (a fork/join inside the captured
        graph), so that in a pipeline of plans its persistent CTAs take freed SM resources before the small
        kernels of the other lanes do."""
        dev, db = self.db.device, self.db
        Q = n_clips * n_seg * STEPS_PER_SEGMENT
        tc = tail_clips if tail_clips is not None else slice(0, n_clips)
        n_tail = tc.stop - tc.start
        if engine is None:
            engine = "sliced" if (db.mode == "A" and db.aud_s is not None) else "f64"
        if engine == "sliced" and (db.mode != "A" or db.aud_s is None):
            raise ValueError("the sliced engine needs a mode-A database built with sliced=True")
        if engine == "sliced" and self.process_group is not None and not db.replicated:
            raise ValueError("row-sharded sliced scans need MatchDatabase(replicate_exact=True)")
        p = SimpleNamespace(n_clips=n_clips, n_seg=n_seg, Q=Q, tail=tc, n_tail=n_tail, graph=None, engine=engine)
        world = 1
        if self.process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(self.process_group)
        p.world = world
        lib = _lib.load()
        with torch.cuda.device(dev):
            Qt = n_tail * n_seg * STEPS_PER_SEGMENT
            p.Qt = Qt
            # all inputs of a step live in ONE device buffer (audio queries | text queries | seed phases | seed codes)
            # and all outputs in another (codes | status), so that a staged caller moves each with a single copy
            a_cols, a_dt = (db.aud.D, torch.float32) if db.mode == "A" else (12, torch.int32)
            sizes = [Q * a_cols * 4, Q * db.txt.D * 4, n_clips * 8 * 16 * 4, n_clips * 4]
            offs = [0]
            for sz in sizes:
                offs.append(offs[-1] + -(-sz // 256) * 256)
            p.inbuf = torch.zeros((offs[-1],), dtype=torch.uint8, device=dev)
            p.in_layout = list(zip(offs[:-1], sizes))
            p.qa = p.inbuf[offs[0]:offs[0] + sizes[0]].view(a_dt).view(Q, a_cols)
            p.qt = p.inbuf[offs[1]:offs[1] + sizes[1]].view(torch.float32).view(Q, db.txt.D)
            p.seed_phase = p.inbuf[offs[2]:offs[2] + sizes[2]].view(torch.float32).view(n_clips, 8, 16)
            p.seed_code = p.inbuf[offs[3]:offs[3] + sizes[3]].view(torch.int32)
            p.fused = False
            if engine == "sliced":
                n_pass = -(-Q // 64)
                per = -(-Q // n_pass)
                p.passes = []
                q0 = 0
                while q0 < Q:
                    nq = min(per, Q - q0)
                    n_pad = -(-nq // 16) * 16
                    from .matchdb import aligned_bytes
                    p.passes.append(SimpleNamespace(
                        q0=q0, nq=nq, n_pad=n_pad,
                        qs_a=aligned_bytes(lib.qpg_sliced_query_bytes(db.aud.D, n_pad), dev),
                        qs_t=aligned_bytes(lib.qpg_sliced_query_bytes(db.txt.D, n_pad), dev)))
                    q0 += nq
                n_pad_max = max(ps.n_pad for ps in p.passes)
                p.qinfo_a = torch.zeros((Q, 4), dtype=torch.float64, device=dev)
                p.qinfo_t = torch.zeros((Q, 4), dtype=torch.float64, device=dev)
                p.sacc_a = torch.zeros((n_pad_max, db.aud_s.Wpad), dtype=torch.int64, device=dev)
                p.sacc_t = torch.zeros((n_pad_max, db.txt_s.Wpad), dtype=torch.int64, device=dev)
                # per-bin records (qpg_bin_t, 32 bytes): [query][audio | text][code], so that the records of one rank's
                # clips are one contiguous block.  Row shards exchange them with ONE collective: an all-to-all when
                # the clips are split evenly (each rank receives only the records of its own clips), else an all-gather
                p.bins = torch.zeros((Q, 2, codebook_size, 4), dtype=torch.int64, device=dev)
                p.exchange = None
                if world == 1:
                    p.parts = p.bins[None]
                else:
                    import torch.distributed as dist
                    r = dist.get_rank(self.process_group)
                    even = n_clips % world == 0 and (tc.start, tc.stop) == (r * (n_clips // world), (r + 1) * (n_clips // world))
                    p.exchange = "all_to_all" if even else "all_gather"
                    p.parts = torch.zeros((world, Qt if even else Q, 2, codebook_size, 4), dtype=torch.int64, device=dev)
                p.stats = torch.zeros((2,), dtype=torch.int64, device=dev)
                p.ta, p.tt = new_table(Qt, dev), new_table(Qt, dev)
            else:
                p.fused = db.fused is not None and not self._sharded_exact() and \
                    -(-db.fused.W // 8) * 2 >= torch.cuda.get_device_properties(dev).multi_processor_count * 12
                if p.fused:
                    p.qf = torch.zeros((Q, db.aud.D + db.txt.D), dtype=torch.float32, device=dev)
                p.fa, p.ft = new_table(Q, dev), new_table(Q, dev)                 # tables of all Q steps
                if self._sharded_exact():
                    p.parts_f = torch.empty((world, 2, Q, codebook_size, 2), dtype=torch.int64, device=dev)
                    p.both = torch.empty((2, Q, codebook_size, 2), dtype=torch.int64, device=dev)
                    p.fa, p.ft = p.both[0], p.both[1]
                    p.merged = torch.empty((2, Q, codebook_size, 2), dtype=torch.int64, device=dev)
            p.ra = torch.empty((Qt, codebook_size), dtype=torch.int32, device=dev)
            p.rt = torch.empty((Qt, codebook_size), dtype=torch.int32, device=dev)
            p.qfa = torch.zeros((Qt,), dtype=torch.int32, device=dev)
            p.qft = torch.zeros((Qt,), dtype=torch.int32, device=dev)
            p.entries = torch.empty((max(Qt, 1), codebook_size, 4), dtype=torch.int64, device=dev)   # 32-byte entries
            # few clips: evaluate the phase pick for every reachable state in parallel, walk a table (lowest latency);
            # many clips: one warp per clip walks directly (the clips hide each other's latency)
            p.trans = torch.empty((Qt, 1024), dtype=torch.int16, device=dev) \
                if (0 < n_tail <= MAX_TABLE_WALK_CLIPS and n_seg * STEPS_PER_SEGMENT <= 104) else None
            n_codes = n_tail * n_seg * num_frames_code
            p.outbuf = torch.zeros((n_codes * 8 + n_tail * 4,), dtype=torch.uint8, device=dev)
            p.codes = p.outbuf[:n_codes * 8].view(torch.int64).view(n_tail, n_seg, num_frames_code)
            p.status = p.outbuf[n_codes * 8:].view(torch.int32)
            p.vote = torch.empty((n_tail, n_seg, STEPS_PER_SEGMENT), dtype=torch.int32, device=dev)
            p.phase = torch.empty((n_tail, n_seg, STEPS_PER_SEGMENT, 8, 16), dtype=torch.float32, device=dev) \
                if want_phase else None
            # the table walk's transition kernel reads per-window phase statistics (built once per database)
            p.phase_stats = ensure_phase_stats(db) if p.trans is not None else None
            p.scan_stream = torch.cuda.Stream(device=dev, priority=-1) if scan_priority else None
            if use_graph:
                self._launch_plan(p)                       # warm-up outside capture (sets function attributes)
                torch.cuda.synchronize(dev)
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._launch_plan(p)
                p.graph = g
        return p

    def _scan_fused(self, q, ta, tt, nq, sp):
        lib, db = _lib.load(), self.db
        _lib.check(lib.qpg_cand_cosine2_minbycode(_lib.ptr(db.fused.packed), _lib.ptr(db.aud.sqnorm), _lib.ptr(db.txt.sqnorm),
                                                  _lib.ptr(db.labels), db.fused.W, db.aud.D, db.txt.D, db.exact_offset,
                                                  _lib.ptr(q), nq, _lib.ptr(ta), _lib.ptr(tt), sp),
                   "qpg_cand_cosine2_minbycode")

    def _sliced_tables(self, p, q0, nq, bins_q0, for_resolve=False):
        """ctypes descriptors (audio, text) of qpg_sliced_table_t for queries [q0, q0+nq); `bins_q0` = first
        query of the bins block the descriptor points at."""
        db = self.db
        tabs = (_lib.SlicedTable * 2)()
        for x, (S, E, sacc, q, qi, tab, rk, qf) in enumerate((
                (db.aud_s, db.aud, p.sacc_a, p.qa, p.qinfo_a, p.ta, p.ra, p.qfa),
                (db.txt_s, db.txt, p.sacc_t, p.qt, p.qinfo_t, p.tt, p.rt, p.qft))):
            t = tabs[x]
            t.packed, t.row_sqnorm = _lib.dptr(E.packed), _lib.dptr(E.sqnorm)
            t.q, t.q_info, t.ldq, t.D = _lib.dptr(q[q0:q0 + nq]), _lib.dptr(qi[q0:q0 + nq]), E.D, E.D
            t.sacc, t.bin_start, t.row_info, t.order = _lib.dptr(sacc), _lib.dptr(S.bin_start), _lib.dptr(S.row_info), \
                _lib.dptr(S.order)
            t.bins = (p.parts[0, bins_q0, x] if for_resolve else p.bins[bins_q0, x]).data_ptr()
            t.bins_qstride = 2 * codebook_size
            if for_resolve:
                t.table, t.ranks, t.qflags = _lib.dptr(tab), _lib.dptr(rk), _lib.dptr(qf)
        return tabs

    def _launch_sliced(self, p, sp):
        """slice queries -> one tensor-core pass per <= 64 steps -> per-bin records -> [all-gather] -> resolve:
        four launches per pass-group on one GPU, audio and text handled together in each."""
        lib, db = _lib.load(), self.db
        A, T = db.aud_s, db.txt_s
        for ps in p.passes:
            jobs = (_lib.SliceJob * 2)()
            for x, (S, q, qi, qs) in enumerate(((A, p.qa, p.qinfo_a, ps.qs_a), (T, p.qt, p.qinfo_t, ps.qs_t))):
                jobs[x].q, jobs[x].col_exp = _lib.dptr(q[ps.q0:ps.q0 + ps.nq]), _lib.dptr(S.col_exp)
                jobs[x].q_slices, jobs[x].q_info = _lib.dptr(qs), _lib.dptr(qi[ps.q0:ps.q0 + ps.nq])
                jobs[x].ldq, jobs[x].D = S.D, S.D
            _lib.check(lib.qpg_slice_queries_i8(jobs, 2, ps.nq, ps.n_pad, sp), "qpg_slice_queries_i8")
            segs = (_lib.SlicedSeg * 2)()
            segs[0].db_slices, segs[0].q_slices, segs[0].sacc, segs[0].n_kblocks = \
                A.slices.data_ptr(), ps.qs_a.data_ptr(), p.sacc_a.data_ptr(), A.n_kblocks
            segs[1].db_slices, segs[1].q_slices, segs[1].sacc, segs[1].n_kblocks = \
                T.slices.data_ptr(), ps.qs_t.data_ptr(), p.sacc_t.data_ptr(), T.n_kblocks
            # sacc is all zero here: zero-initialised by make_plan and re-zeroed by every bins stage (consume = 1)
            side = getattr(p, "scan_stream", None)
            if side is None:
                _lib.check(lib.qpg_sliced_scan_i8(segs, 2, A.W, ps.n_pad, ps.nq, sp), "qpg_sliced_scan_i8")
            else:
                cur = torch.cuda.current_stream()
                fork, join = torch.cuda.Event(), torch.cuda.Event()
                fork.record(cur)
                side.wait_event(fork)
                _lib.check(lib.qpg_sliced_scan_i8(segs, 2, A.W, ps.n_pad, ps.nq, _lib.stream_ptr(side)),
                           "qpg_sliced_scan_i8")
                join.record(side)
                cur.wait_event(join)
            _lib.check(lib.qpg_sliced_bins(self._sliced_tables(p, ps.q0, ps.nq, ps.q0), 2, A.W, ps.nq, db.id_offset,
                                           db.row_base, 1, _lib.ptr(p.stats), sp), "qpg_sliced_bins")
        per_clip = p.n_seg * STEPS_PER_SEGMENT
        q0, q1 = p.tail.start * per_clip, p.tail.stop * per_clip
        part_q0 = q0
        if p.world > 1:                                                               # the ONE data-path collective
            import torch.distributed as dist
            if p.exchange == "all_to_all":
                dist.all_to_all_single(p.parts, p.bins, group=self.process_group)
                part_q0 = 0                                                            # parts hold this rank's clips only
            else:
                dist.all_gather_into_tensor(p.parts, p.bins, group=self.process_group)
        stride = p.parts.shape[1] * 2 * codebook_size                                  # records between two parts
        _lib.check(lib.qpg_sliced_resolve(self._sliced_tables(p, q0, q1 - q0, part_q0, for_resolve=True), 2, p.world,
                                          stride, q1 - q0, db.exact_offset, _lib.ptr(p.stats), sp), "qpg_sliced_resolve")
        return p.ta, p.tt

    def _launch_f64(self, p, sp):
        """round-1 float64 streaming scans (exact distances everywhere) + stable ranks with tie flags."""
        lib, db = _lib.load(), self.db
        if p.fused:
            p.qf[:, :db.aud.D].copy_(p.qa)
            p.qf[:, db.aud.D:].copy_(p.qt)
            _lib.check(lib.qpg_table_init(_lib.ptr(p.fa), p.Q * codebook_size, sp), "qpg_table_init")
            _lib.check(lib.qpg_table_init(_lib.ptr(p.ft), p.Q * codebook_size, sp), "qpg_table_init")
            self._scan_fused(p.qf, p.fa, p.ft, p.Q, sp)
        for which, q, tab in (() if p.fused else (("audio", p.qa, p.fa), ("text", p.qt, p.ft))):
            _lib.check(lib.qpg_table_init(_lib.ptr(tab), p.Q * codebook_size, sp), "qpg_table_init")
            if which == "text" or db.mode == "A":
                t = db.txt if which == "text" else db.aud
                _lib.check(lib.qpg_cand_cosine_minbycode_team(_lib.ptr(t.packed), _lib.ptr(t.sqnorm), _lib.ptr(db.labels),
                                                              t.W, t.D, db.exact_offset, _lib.ptr(q), p.Q, _lib.ptr(tab), 0,
                                                              self._team_size(t.D), sp), "qpg_cand_cosine_minbycode_team")
            else:
                _lib.check(lib.qpg_cand_lev_minbycode(_lib.ptr(db.tokens), _lib.ptr(db.labels), db.W, db.id_offset,
                                                      _lib.ptr(q), p.Q, _lib.ptr(tab), sp), "qpg_cand_lev_minbycode")
        fa, ft = p.fa, p.ft
        if self._sharded_exact():
            import torch.distributed as dist
            dist.all_gather_into_tensor(p.parts_f, p.both, group=self.process_group)   # both tables, one collective
            _lib.check(lib.qpg_table_merge(_lib.ptr(p.parts_f), p.world, 2 * p.Q * codebook_size, _lib.ptr(p.merged), sp),
                       "qpg_table_merge")
            fa, ft = p.merged[0], p.merged[1]
        per_clip = p.n_seg * STEPS_PER_SEGMENT
        q0, q1 = p.tail.start * per_clip, p.tail.stop * per_clip
        ta, tt = fa[q0:q1], ft[q0:q1]
        _lib.check(lib.qpg_rank512_ties(_lib.ptr(ta), q1 - q0, _lib.ptr(p.ra), _lib.ptr(p.qfa), sp), "qpg_rank512_ties")
        _lib.check(lib.qpg_rank512_ties(_lib.ptr(tt), q1 - q0, _lib.ptr(p.rt), _lib.ptr(p.qft), sp), "qpg_rank512_ties")
        p.ta, p.tt = ta, tt
        return ta, tt

    def _launch_plan(self, p):
        lib, db = _lib.load(), self.db
        sp = _lib.stream_ptr()
        ta, tt = self._launch_sliced(p, sp) if p.engine == "sliced" else self._launch_f64(p, sp)
        if p.Qt == 0:
            return
        sc, sph = p.seed_code[p.tail], p.seed_phase[p.tail]
        _lib.check(lib.qpg_match_lookup(_lib.ptr(ta), _lib.ptr(tt), _lib.ptr(p.ra), _lib.ptr(p.rt), _lib.ptr(db.pos_rank_t),
                                        _lib.ptr(db.freq_rank), _lib.ptr(db.code), db.n_seq, _lib.ptr(db.aud_frame),
                                        _lib.ptr(db.txt_frame), _lib.ptr(p.qfa), _lib.ptr(p.qft), p.Qt,
                                        _lib.ptr(p.entries), sp), "qpg_match_lookup")
        _lib.check(lib.qpg_match_walk_stats(_lib.ptr(p.entries), _lib.ptr(db.code), _lib.ptr(db.phase_amp),
                                            _lib.ptr(p.phase_stats), _lib.ptr(sc), _lib.ptr(sph), p.n_tail, p.n_seg,
                                            _lib.ptr(p.trans), _lib.ptr(p.codes), _lib.ptr(p.vote), _lib.ptr(p.phase),
                                            _lib.ptr(p.status), sp), "qpg_match_walk_stats")

    def make_pipeline(self, n_clips: int, n_seg: int, depth: int = 3, **plan_kwargs):
        """`depth` independent plans, each with its own stream, buffers, captured graph and pinned host mirrors:
        consecutive steps (independent batches of clips) issued round-robin overlap, so the latency-bound small
        kernels of one step run beside the HBM-bound scan of the next and the host<->device copies of a third.
        Returns a list of lanes with .plan, .stream, .io; use run_lane / stage_lane."""
        dev = self.db.device
        lanes = []
        for _ in range(depth):
            st = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(st):
                p = self.make_plan(n_clips, n_seg, **plan_kwargs)
            st.synchronize()
            lanes.append(SimpleNamespace(plan=p, stream=st, io=self.pinned_io(p)))
        return lanes

    def run_lane(self, lane):
        """one step of a pipeline lane on the lane's stream, inputs already in the plan's device buffers"""
        with torch.cuda.stream(lane.stream):
            self.run_plan(lane.plan)

    def stage_lane(self, lane):
        """one step of a pipeline lane from its pinned host buffers (H2D, step, D2H) on the lane's stream; the
        caller synchronises the lane's stream and looks at lane.io.status / lane.io.codes"""
        with torch.cuda.stream(lane.stream):
            self.match_staged(lane.plan, lane.io, sync=False)

    def pinned_io(self, p):
        """Pinned host mirrors of a plan's input and output buffers with typed views (qa, qt, seed_phase,
        seed_code; codes, status).  Fill the input views, call match_staged(p, io): one H2D copy, the captured
        step, one D2H copy."""
        io = SimpleNamespace(inp=torch.zeros(p.inbuf.shape, dtype=torch.uint8).pin_memory(),
                             out=torch.zeros(p.outbuf.shape, dtype=torch.uint8).pin_memory())
        (o0, s0), (o1, s1), (o2, s2), (o3, s3) = p.in_layout
        io.qa = io.inp[o0:o0 + s0].view(p.qa.dtype).view(p.qa.shape)
        io.qt = io.inp[o1:o1 + s1].view(torch.float32).view(p.qt.shape)
        io.seed_phase = io.inp[o2:o2 + s2].view(torch.float32).view(p.seed_phase.shape)
        io.seed_code = io.inp[o3:o3 + s3].view(torch.int32)
        n = p.codes.numel() * 8
        io.codes = io.out[:n].view(torch.int64).view(p.codes.shape)
        io.status = io.out[n:].view(torch.int32)
        return io

    def match_staged(self, p, io, sync=True):
        """One step from pinned host buffers: H2D of all inputs (one copy), the step, D2H of codes + status (one
        copy), all on the current stream.  With sync=False the caller synchronises and MUST look at io.status
        (bit 0: IndexError of GestureKNN.py:631 - that clip's remaining codes are -1; bit 1: tie dependent)."""
        with torch.cuda.device(self.db.device):
            p.inbuf.copy_(io.inp, non_blocking=True)
            self.run_plan(p)
            io.out.copy_(p.outbuf, non_blocking=True)
            if sync:
                torch.cuda.current_stream().synchronize()
                if int(io.status.max()) & 1:
                    raise IndexError("list index out of range")
        return io.codes

    def run_plan(self, p):
        """Enqueue one step on the current stream (graph replay when the plan was captured)."""
        with torch.cuda.device(self.db.device):
            if p.graph is not None:
                p.graph.replay()
            else:
                self._launch_plan(p)
        return p.codes

    # ---- sequential tail ---------------------------------------------------------
    def tail_device(self, ta, tt, seed_code, seed_phase, n_clips, n_seg, want_phase=True):
        """rank + lookup + walk kernels for n_clips x n_seg x 8 steps already scanned.
        -> (codes, vote, phase, status); status bit 0: IndexError of GestureKNN.py:631, bit 1: tie dependent."""
        lib, db, dev = _lib.load(), self.db, self.db.device
        Q = n_clips * n_seg * STEPS_PER_SEGMENT
        assert ta.shape[0] == Q and tt.shape[0] == Q
        with torch.cuda.device(dev):
            sp = _lib.stream_ptr()
            ra = torch.empty((Q, codebook_size), dtype=torch.int32, device=dev)
            rt = torch.empty((Q, codebook_size), dtype=torch.int32, device=dev)
            qfa = torch.empty((Q,), dtype=torch.int32, device=dev)
            qft = torch.empty((Q,), dtype=torch.int32, device=dev)
            _lib.check(lib.qpg_rank512_ties(_lib.ptr(ta), Q, _lib.ptr(ra), _lib.ptr(qfa), sp), "qpg_rank512_ties")
            _lib.check(lib.qpg_rank512_ties(_lib.ptr(tt), Q, _lib.ptr(rt), _lib.ptr(qft), sp), "qpg_rank512_ties")
            sc = torch.as_tensor(np.asarray(seed_code, dtype=np.int32).reshape(n_clips), device=dev)
            sph = torch.as_tensor(np.ascontiguousarray(seed_phase, dtype=np.float32).reshape(n_clips, 8, 16), device=dev)
            entries = torch.empty((Q, codebook_size, 4), dtype=torch.int64, device=dev)
            codes = torch.empty((n_clips, n_seg, num_frames_code), dtype=torch.int64, device=dev)
            vote = torch.empty((n_clips, n_seg, STEPS_PER_SEGMENT), dtype=torch.int32, device=dev)
            status = torch.empty((n_clips,), dtype=torch.int32, device=dev)
            phase = torch.empty((n_clips, n_seg, STEPS_PER_SEGMENT, 8, 16), dtype=torch.float32, device=dev) \
                if want_phase else None
            _lib.check(lib.qpg_match_lookup(_lib.ptr(ta), _lib.ptr(tt), _lib.ptr(ra), _lib.ptr(rt), _lib.ptr(db.pos_rank_t),
                                            _lib.ptr(db.freq_rank), _lib.ptr(db.code), db.n_seq, _lib.ptr(db.aud_frame),
                                            _lib.ptr(db.txt_frame), _lib.ptr(qfa), _lib.ptr(qft), Q, _lib.ptr(entries), sp),
                       "qpg_match_lookup")
            trans = torch.empty((Q, 1024), dtype=torch.int16, device=dev) \
                if (n_clips <= MAX_TABLE_WALK_CLIPS and n_seg * STEPS_PER_SEGMENT <= 104) else None
            _lib.check(lib.qpg_match_walk(_lib.ptr(entries), _lib.ptr(db.code), _lib.ptr(db.phase_amp), _lib.ptr(sc),
                                          _lib.ptr(sph), n_clips, n_seg, _lib.ptr(trans), _lib.ptr(codes), _lib.ptr(vote),
                                          _lib.ptr(phase), _lib.ptr(status), sp), "qpg_match_walk")
        return codes, vote, phase, status

    def _tail_numpy_segment(self, ta_np, tt_np, seed_code, seed_phase, desired_k=0, use_txt=True, use_aud=True):
        """The reference's own per-step NumPy calls (GestureKNN.py:528-660) on the
        device-computed tables -- inherits NumPy's tie order on this machine."""
        from sklearn.metrics.pairwise import paired_distances

        db = self.db
        result, result_phase, vote = [int(seed_code)], [np.asarray(seed_phase)], []
        for s in range(ta_np.shape[0]):
            pos_score = db.pos_rank_host[result[-1]].astype(np.int64)
            pos_score = pos_score + db.freq_rank_host.astype(np.int64) * 0.05
            picks = []
            if use_aud:
                aud_score = np.array([float(x) for x in ta_np[s]["d"]]).argsort().argsort()
                idx = np.argsort(pos_score + aud_score).tolist()
                picks += [("audio", ta_np[s], i) for i in (idx[:1] if use_txt else idx[:2])]
            if use_txt:
                txt_score = np.array([float(x) for x in tt_np[s]["d"]]).argsort().argsort()
                idx_ = np.argsort(pos_score + txt_score).tolist()
                picks += [("text", tt_np[s], i) for i in (idx_[:1] if use_aud else idx_[:2])]
            tmp_distance, tmp_phase_amp, wins = [], [], []
            for which, tab, c in picks:
                w = int(tab["id"][c])
                if w < 0:
                    raise IndexError("list index out of range")       # reference: aux[index] == [] at :631
                j, k = db.aux(w, which)
                f = int(k / 398 * 240)
                win = db.phase_amp_host[j, f:f + 32]
                head = win[:8]
                a = np.concatenate((result_phase[-1][-5:], head[:3]), axis=0).reshape(-1)
                b = np.concatenate((result_phase[-1][-3:], head[:5]), axis=0).reshape(-1)
                tmp_distance.append(paired_distances([a], [b], metric="cosine")[0])
                tmp_phase_amp.append(win[-8:])
                wins.append(w)
            final_index = tmp_distance.index(min(tmp_distance))
            result.extend(int(x) for x in db.payload(wins[final_index]))
            result_phase.append(tmp_phase_amp[final_index])
            vote.append(final_index)
        return (np.array(result)[1:1 + num_frames_code], np.array(result_phase[1:]), np.array(vote))

    def search_code_knn(self, clip_test, desired_k, use_feature=False, use_wavlm=False, use_freq=False,
                        seed_code=None, use_wavvq=False, use_phase=False, seed_phase=None, use_txt=False,
                        clip_context=None, use_aud=False):
        """One 4-s segment (GestureKNN.py:501-664).  clip_test: (180, 6C) stacked WavLM
        feature in mode A, (398, 22) stacked wavvq feature in mode B; clip_context (30, Dt)."""
        if not (use_phase and use_feature):
            raise NotImplementedError("only the phase-guided feature matcher (GestureKNN.py:842-843) is built")
        if not (use_aud or use_txt):
            raise ValueError("need use_aud and/or use_txt")
        clip_test = np.asarray(clip_test)
        if seed_code is not None:
            init_code, init_phase_amp = seed_code, seed_phase
        else:
            init_code, init_phase_amp = self.init_code_phase()
        n = len(clip_test)
        i_list, i = [], 0
        while i < n:                                                   # :528, :659
            i_list.append(i)
            i += STEP_SZ * self.step_sz
        aud_q = clip_test[[int(i) for i in i_list]]
        denom = n if self.db.mode == "A" else 398                       # :548-551
        ctx = np.asarray(clip_context)
        txt_q = ctx[[int(i / denom * 30) for i in i_list]]
        ta, tt = self.match_tables(aud_q, txt_q)
        if self.tail in ("device", "auto") and use_aud and use_txt and len(i_list) == STEPS_PER_SEGMENT:
            codes, vote, phase, status = self.tail_device(ta, tt, [init_code], np.asarray(init_phase_amp)[None], 1, 1)
            st = int(status.cpu()[0])
            self.last_status = np.array([st])
            if not (st & 2 and self.tail == "auto"):                  # tie dependent -> NumPy's own order below
                if st & 1:
                    raise IndexError("list index out of range")       # same failure as GestureKNN.py:631
                return codes[0, 0].cpu().numpy(), phase[0, 0].cpu().numpy(), vote[0, 0].cpu().numpy()
        return self._tail_numpy_segment(table_to_numpy(ta), table_to_numpy(tt), init_code, init_phase_amp,
                                        desired_k, use_txt=use_txt, use_aud=use_aud)

    # ---- batched entry: many clips at once ---------------------------------------
    def match_clips(self, aud_q, txt_q, seed_code=None, seed_phase=None, tail=None, out=None, sync=True,
                    tail_clips=None, status_out=None, engine=None):
        """aud_q [n_clips, n_seg, 8, Da] (or tokens [..., 11|22]), txt_q [n_clips, n_seg, 8, Dt]: host NumPy
        arrays or (pinned) host torch tensors.  Returns int64 codes [n_clips, n_seg, 30] on the host.

        tail: "auto" (default) = device tail; clips whose result depended on the order of exact ties are
        redone with the reference's own NumPy calls (`_tail_numpy_segment`) so that the platform's argsort
        order applies as it does in the reference; "device" = stable order (lower code first) everywhere;
        "numpy" = NumPy tail for every clip.  A chosen start code without window raises IndexError as
        GestureKNN.py:631 does.

        With `out` (+ `status_out`, pinned int64 / int32 host tensors) and sync=False the call only enqueues
        the H2D copies, the captured step and the D2H copies on the current stream: the caller synchronises
        and MUST look at status_out (bit 0: IndexError, bit 1: tie dependent) - rows of failed clips are -1."""
        tail = tail or self.tail
        is_t = isinstance(aud_q, torch.Tensor)
        n_clips, n_seg = aud_q.shape[0], aud_q.shape[1]
        Q = n_clips * n_seg * STEPS_PER_SEGMENT
        if seed_code is None:
            seeds = [self.init_code_phase() for _ in range(n_clips)]
            seed_code = [s[0] for s in seeds]
            seed_phase = np.stack([s[1] for s in seeds])
        if tail in ("device", "auto"):
            key = (n_clips, n_seg, None if tail_clips is None else (tail_clips.start, tail_clips.stop), engine)
            plans = self.__dict__.setdefault("_plans", {})
            if key not in plans:
                plans[key] = self.make_plan(n_clips, n_seg, tail_clips=tail_clips, engine=engine)
            p = plans[key]
            dev = self.db.device
            with torch.cuda.device(dev):
                if is_t:
                    qa_h, qt_h = aud_q.reshape(Q, -1), txt_q.reshape(Q, -1)
                else:
                    aud_q, txt_q = np.asarray(aud_q), np.asarray(txt_q)
                    qa_h = self._audio_query_tensor(aud_q.reshape((Q,) + aud_q.shape[3:]))
                    qt_h = torch.from_numpy(np.ascontiguousarray(txt_q.reshape(Q, -1), dtype=np.float32))
                sc_h = seed_code if isinstance(seed_code, torch.Tensor) else \
                    torch.from_numpy(np.asarray(seed_code, dtype=np.int32).reshape(n_clips))
                sp_h = seed_phase if isinstance(seed_phase, torch.Tensor) else \
                    torch.from_numpy(np.ascontiguousarray(seed_phase, dtype=np.float32).reshape(n_clips, 8, 16))
                p.qa.copy_(qa_h, non_blocking=True)
                p.qt.copy_(qt_h, non_blocking=True)
                p.seed_code.copy_(sc_h, non_blocking=True)
                p.seed_phase.copy_(sp_h, non_blocking=True)
                self.run_plan(p)
                if out is not None:
                    out.copy_(p.codes, non_blocking=True)
                    if status_out is not None:
                        status_out.copy_(p.status, non_blocking=True)
                    if not sync:
                        return out
                    torch.cuda.synchronize(dev)
                    codes_h = out.numpy()
                    status = status_out.numpy() if status_out is not None else p.status.cpu().numpy()
                else:
                    codes_h = p.codes.cpu().numpy()
                    status = p.status.cpu().numpy()
                self.last_status = status.copy()
                redo = [b for b in range(p.n_tail) if status[b] & 2] if tail == "auto" else []
                if any((status[b] & 1) and b not in redo for b in range(p.n_tail)):
                    raise IndexError("list index out of range")
                if redo:                                                  # NumPy's own tie order for these clips
                    codes_h = np.array(codes_h, copy=True)
                    ta_np = table_to_numpy(p.ta).reshape(p.n_tail, n_seg, STEPS_PER_SEGMENT, codebook_size)
                    tt_np = table_to_numpy(p.tt).reshape(p.n_tail, n_seg, STEPS_PER_SEGMENT, codebook_size)
                    sc_np, sp_np = np.asarray(sc_h), np.asarray(sp_h)
                    for b in redo:
                        code0, ph0 = int(sc_np[p.tail.start + b]), sp_np[p.tail.start + b]
                        for g in range(n_seg):
                            codes, phases, _ = self._tail_numpy_segment(ta_np[b, g], tt_np[b, g], code0, ph0)
                            codes_h[b, g] = codes
                            code0, ph0 = int(codes[-1]), phases[-1]
            return codes_h
        aud_q, txt_q = np.asarray(aud_q), np.asarray(txt_q)
        ta, tt = self.match_tables(aud_q.reshape((Q,) + aud_q.shape[3:]), txt_q.reshape(Q, -1))
        ta_np = table_to_numpy(ta).reshape(n_clips, n_seg, STEPS_PER_SEGMENT, codebook_size)
        tt_np = table_to_numpy(tt).reshape(n_clips, n_seg, STEPS_PER_SEGMENT, codebook_size)
        out = np.empty((n_clips, n_seg, num_frames_code), dtype=np.int64)
        for b in range(n_clips):
            code0, ph0 = seed_code[b], np.asarray(seed_phase[b])
            for g in range(n_seg):
                codes, phases, _ = self._tail_numpy_segment(ta_np[b, g], tt_np[b, g], code0, ph0)
                out[b, g] = codes
                code0, ph0 = int(codes[-1]), phases[-1]
        return out


MAX_TABLE_WALK_CLIPS = 2


class GestureKNN(object):
    """The reference's legacy pose-feature matcher (GestureKNN.py:70-284) on the device: same constructor, `init_frame`,
    `search_motion(feat_test, desired_k)` and `search_fake_motion(feat_test, desired_k)` with the reference's shapes
    and return values, plus batched forms that advance many clips together (csrc/legacy_knn.cu: four launches per
    8-frame step whatever the batch size).  `last_status` holds the per-clip status of the last call (bit 0: fewer
    candidates than desired_k + 1, where the reference raises IndexError - raised here too; bit 1: the result
    depended on the order of exact ties, which NumPy's unstable argsort decides in the reference)."""

    def __init__(self, feat_train, motn_train, control_mask, n_aud_feat=112, n_body_feat=96, n_joints=165, step_sz=8,
                 device=None, ties="numpy"):
        """ties: "numpy" (default) - the one choice per step among the rank sums (small integers, often tied) is made
        by np.argsort on the host, as in the reference (:139), at the cost of one small device->host copy per step;
        "stable" - everything on the device, lower sequence first among tied sums, ties reported in last_status."""
        if ties not in ("numpy", "stable"):
            raise ValueError("ties must be 'numpy' or 'stable'")
        self.ties = ties
        self.n_aud_feat, self.n_body_feat, self.n_joints, self.step_sz = n_aud_feat, n_body_feat, n_joints, step_sz
        self.feat_train, self.motn_train, self.control_mask = feat_train, motn_train, control_mask
        self.n_db_seq, self.n_db_frm = feat_train.shape[0], feat_train.shape[1]
        self.device = torch.device(device if device is not None else "cuda")
        _lib.load()                                           # fail loudly without the CUDA library
        if feat_train.shape[2] < n_aud_feat + n_body_feat or motn_train.shape[2] != n_joints:
            raise ValueError("feature / motion widths do not match n_aud_feat + n_body_feat / n_joints")
        self._feat = torch.as_tensor(np.ascontiguousarray(feat_train, dtype=np.float64), device=self.device)
        self._motn = torch.as_tensor(np.ascontiguousarray(motn_train, dtype=np.float64), device=self.device)
        self._mask = torch.as_tensor(np.ascontiguousarray(np.asarray(control_mask) != 0, dtype=np.int32), device=self.device)
        self.last_status = None
        self.last_chosen = None

    def init_frame(self):
        """GestureKNN.py:89-97: the same draws from NumPy's global generator."""
        init_seq = np.random.randint(0, self.n_db_seq)
        init_frm = np.random.randint(0, self.n_db_frm)
        while self.control_mask[init_seq, init_frm] != 1:
            init_seq = np.random.randint(0, self.n_db_seq)
            init_frm = np.random.randint(0, self.n_db_frm)
        return init_seq, init_frm

    def _run(self, feat_tests, desired_k, inits, fake):
        lib, dev = _lib.load(), self.device
        ft = torch.as_tensor(np.ascontiguousarray(feat_tests, dtype=np.float64), device=dev)     # [B, n_aud(+), n_frames]
        B, n_frames = ft.shape[0], ft.shape[2]
        na, nb, J, st, S = self.n_aud_feat, self.n_body_feat, self.n_joints, self.step_sz, self.n_db_seq
        aud = ft[:, :na, :].permute(0, 2, 1).contiguous()                                          # [B, n_frames, n_aud]
        dk_host = np.broadcast_to(np.asarray(desired_k, dtype=np.int32), (B,)).copy()
        dk = torch.as_tensor(dk_host, device=dev)
        starts = list(range(0, n_frames, st)) if fake else list(range(1, n_frames, st))
        out_frames = n_frames if fake else n_frames + 1
        pred = torch.zeros((B, J, out_frames), dtype=torch.float64, device=dev)
        cand_bytes = int(lib.qpg_legacy_cand_bytes())
        cands = torch.empty((B, S, cand_bytes), dtype=torch.uint8, device=dev)
        comb = torch.empty((B, S), dtype=torch.int32, device=dev)
        tie = torch.empty_like(comb)
        chosen = torch.zeros((B, 2), dtype=torch.int32, device=dev)
        n_found = torch.zeros((B,), dtype=torch.int32, device=dev)
        log = torch.full((B, len(starts), 2), -1, dtype=torch.int32, device=dev)
        status = torch.zeros((B,), dtype=torch.int32, device=dev)
        pose = None
        if not fake:
            idx = torch.as_tensor(np.asarray(inits, dtype=np.int64).reshape(B, 2), device=dev)
            pose = self._feat[idx[:, 0], idx[:, 1], na:na + nb].contiguous()                     # :111
        sp = _lib.stream_ptr()
        F = self._feat.shape[2]
        for s_i, j in enumerate(starts):
            a_q = aud[:, j if fake else j - 1].contiguous()           # feat_test' column j = feat_test column j - 1 (:105)
            query = a_q if fake else pose
            _lib.check(lib.qpg_legacy_candidates(
                _lib.ptr(self._feat), _lib.ptr(self._mask), S, self.n_db_frm, F, na, nb, st, B, 1 if fake else 0,
                _lib.ptr(query), query.shape[1], None if fake else _lib.ptr(a_q), _lib.ptr(cands), _lib.ptr(comb),
                _lib.ptr(tie), sp), "qpg_legacy_candidates")
            if self.ties == "numpy":
                # the reference's own call on the rank sums (GestureKNN.py:139): NumPy's order among tied sums
                comb_h = comb.cpu().numpy()
                frames_h = cands.cpu().numpy()[:, :, 16:20].copy().view(np.int32)[:, :, 0]
                ch, nf = np.zeros((B, 2), dtype=np.int32), np.zeros((B,), dtype=np.int32)
                for b in range(B):
                    found = np.nonzero(comb_h[b] >= 0)[0]
                    nf[b] = len(found)
                    if dk_host[b] < len(found):
                        w = found[np.argsort(comb_h[b][found].astype(np.int64))[dk_host[b]]]
                        ch[b] = (w, frames_h[b, w])
                chosen.copy_(torch.from_numpy(ch))
                n_found.copy_(torch.from_numpy(nf))
            else:
                _lib.check(lib.qpg_legacy_pick(_lib.ptr(cands), _lib.ptr(comb), _lib.ptr(tie), S, B, _lib.ptr(dk),
                                               _lib.ptr(chosen), _lib.ptr(n_found), _lib.ptr(status), sp), "qpg_legacy_pick")
            _lib.check(lib.qpg_legacy_gather(
                _lib.ptr(self._feat), _lib.ptr(self._motn), S, self.n_db_frm, F, J, na, nb, st, B, _lib.ptr(chosen),
                _lib.ptr(n_found), _lib.ptr(dk), j, out_frames, s_i, len(starts), _lib.ptr(pred),
                None if fake else _lib.ptr(pose), _lib.ptr(log), _lib.ptr(status), sp), "qpg_legacy_gather")
        self.last_status = status.cpu().numpy()
        self.last_chosen = log.cpu().numpy()
        if (self.last_status & 1).any():
            bad = int(np.nonzero(self.last_status & 1)[0][0])
            raise IndexError(f"clip {bad}: fewer candidates than desired_k + 1 (GestureKNN.py:144)")
        out = pred.cpu().numpy()
        return out if fake else out[:, :, 1:]

    def search_motion_batch(self, feat_tests, desired_k, inits=None):
        """feat_tests [B, n_aud_feat, n_frames]; inits [B, 2] = (sequence, frame) per clip (default: init_frame()
        per clip, in order).  Returns [B, n_joints, n_frames]."""
        B = len(feat_tests)
        if inits is None:
            inits = [self.init_frame() for _ in range(B)]
        return self._run(feat_tests, desired_k, inits, fake=False)

    def search_fake_motion_batch(self, feat_tests, desired_k):
        return self._run(feat_tests, desired_k, None, fake=True)

    def search_motion(self, feat_test, desired_k):
        """GestureKNN.py:100-152 -> pred_motion [n_joints, n_frames]."""
        return self.search_motion_batch(np.asarray(feat_test)[None], desired_k)[0]

    def search_fake_motion(self, feat_test, desired_k):
        """GestureKNN.py:217-243 -> pred_motion [n_joints, n_frames]."""
        return self.search_fake_motion_batch(np.asarray(feat_test)[None], desired_k)[0]


def predict_gesture_from_audio(feat_train, pose_train, feat_test, control_mask, data_stats, k=0, n_aud_feat=112,
                               n_body_feat=96, n_joints=165, step_sz=8, frames=0, fake=False, device=None):
    """GestureKNN.py:287-341 (`fake` replaces the module-global args.fake): normalise, build the matcher, match every
    test sequence - all of them in ONE batched call instead of a Python loop.  Returns [n_test, n_joints, n_frames]."""
    feat_mean, feat_std = data_stats['feat_mean'], data_stats['feat_std']
    norm = lambda d, m, sd: (d - m) / (sd + 1E-8)                                   # utils.normalize_data
    norm_feat_test = norm(feat_test, feat_mean[:, :n_aud_feat], feat_std[:, :n_aud_feat])
    norm_feat_train = norm(feat_train, feat_mean, feat_std).transpose((0, 2, 1))
    n_test_seq = frames if frames != 0 else feat_test.shape[0]
    knn = GestureKNN(feat_train=norm_feat_train, motn_train=pose_train.transpose((0, 2, 1)), control_mask=control_mask,
                     n_aud_feat=n_aud_feat, n_body_feat=n_body_feat, n_joints=n_joints, step_sz=step_sz, device=device)
    p = [0.5] + [0.5 / 14] * 14
    desired_k = np.random.choice(15, n_test_seq, p=p)                               # :322-323 (drawn in both modes)
    if fake:
        return knn.search_fake_motion_batch(norm_feat_test[:n_test_seq], desired_k)
    return knn.search_motion_batch(norm_feat_test[:n_test_seq], k)


def _phase_ntc(phase_train):
    """CodeKNN receives phase as (N, 240, 4[, 8]); accept the (N, 4, 240[, 8]) layout
    load_db_codebook returns as well."""
    p = phase_train
    if isinstance(p, np.ndarray) and p.ndim >= 3 and p.shape[1] == 4 and p.shape[2] != 4:
        p = np.swapaxes(p, 1, 2)
    return p


# --------------------------------------------------------------------------
def predict_code_from_audio(train_mfcc, train_code, test_mfcc, data_stats, train_feat, test_feat, train_wavlm,
                            test_wavlm, train_wavlm_feat, test_wavlm_feat, speech_features, test_speech_features,
                            train_speech_features_feat, test_speech_features_feat, train_wavvq_feat, test_wavvq_feat,
                            train_phase, test_phase, train_context, test_context, use_feature=False, use_wavlm=False,
                            use_freq=False, use_speechfeat=False, use_wavvq=False, use_phase=False, use_txt=False,
                            use_aud=False, frames=0, **knn_kwargs):
    """GestureKNN.py:724-813 with the reference's positional arguments (arrays in
    (N, feat, time) order as load_db_codebook returns them).  Segments are chained
    through (last code, last phase) exactly as :791,:800."""
    tr = lambda a: None if a is None else np.asarray(a).transpose((0, 2, 1))
    n_test_seq = frames if frames != 0 else test_wavvq_feat.shape[0]                    # :740
    knn = CodeKNN(code_train=train_code, wavlm_train=tr(train_wavlm), wavlm_train_feat=tr(train_wavlm_feat),
                  wavvq_train_feat=tr(train_wavvq_feat), phase_train=_phase_ntc(train_phase),
                  context_train=tr(train_context), use_wavlm=use_wavlm, use_wavvq=use_wavvq, use_phase=use_phase,
                  use_txt=use_txt, **knn_kwargs)
    clips = tr(test_wavlm_feat) if use_wavlm else tr(test_wavvq_feat)
    ctx = tr(test_context)
    motion_output, phase_output = [], []
    for i in range(n_test_seq):
        pred_motion, pred_phase, _ = knn.search_code_knn(
            clip_test=clips[i], desired_k=(args.desired_k if args is not None else 0), use_wavlm=use_wavlm,
            use_feature=use_feature, use_freq=use_freq, seed_code=motion_output[-1][-1] if i > 0 else None,
            use_wavvq=use_wavvq, use_phase=use_phase, seed_phase=phase_output[-1][-1] if i > 0 else None,
            use_txt=use_txt, clip_context=ctx[i] if use_txt else None, use_aud=use_aud)
        motion_output.append(pred_motion)
        phase_output.append(pred_phase)
    return np.array(motion_output)


def build_knn_from_files(a, mode="A", tail="auto", device=None, seq_range=None, process_group=None):
    """Lean construction used by the CLI and the bench: reads the 8 npz files and
    uploads only what the shipped matcher scans."""
    from .data_processing import load_match_inputs

    inp = load_match_inputs(a.train_database, a.train_codebook, a.test_data, a.train_wavlm, a.test_wavlm,
                            a.train_wavvq, a.test_wavvq, mode=mode,
                            device=(device if device is not None else "cuda") if seq_range is None else None)
    signature = np.load(a.codebook_signature)["signature"]
    db = MatchDatabase(mode, inp["code"], signature, phase_to_dense(inp["phase"]), inp["txt_rows"],
                       aud_rows=inp.get("aud_rows"), aud_tokens=inp.get("aud_tokens"),
                       freq_code=np.load(a.train_codebook)["code"], device=device, seq_range=seq_range)
    knn = CodeKNN(database=db, use_wavlm=mode == "A", use_wavvq=mode == "B", use_phase=True, use_txt=True, tail=tail,
                  process_group=process_group)
    return knn, inp


def main_codebook(maxFrames=0, mode=None, tail=None):
    """GestureKNN.py:816-845: load, match every test segment, write knn_pred."""
    mode = mode or getattr(args, "mode", "A")
    tail = tail or getattr(args, "tail", "auto")
    device = torch.device("cuda", getattr(args, "gpu", 0))
    knn, inp = build_knn_from_files(args, mode=mode, tail=tail, device=device)
    n_test_seq = maxFrames if maxFrames != 0 else inp["n_test"]                            # :740
    aud_q, txt_q = inp["aud_q"][:n_test_seq], inp["txt_q"][:n_test_seq]
    pred_seqs = knn.match_clips(aud_q[None], txt_q[None], tail=tail)[0]
    print(pred_seqs.shape)
    os.makedirs(os.path.dirname(os.path.abspath(args.out_knn_filename)), exist_ok=True)
    np.savez_compressed(args.out_knn_filename, knn_pred=pred_seqs)
    return pred_seqs


def main(argv=None):
    set_args(build_parser().parse_args(argv))
    seed_everything()
    return main_codebook(maxFrames=args.max_frames)


if __name__ == "__main__":
    main()

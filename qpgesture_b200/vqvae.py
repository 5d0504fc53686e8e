"""B200 drop-in for the inference surface of the reference VQ-VAE
(codebook/models/vqvae.py, bottleneck.py, encdec.py, resnet.py).

Same class and method names and argument meaning -- `VQVAE(hps, input_dim)`
with `.encode(x) -> [LongTensor]`, `.decode(zs) -> Tensor[B,T,C]`,
`.load_state_dict` accepting the reference's checkpoint keys (with or without
the DataParallel 'module.' prefix), `BottleneckBlock.quantise / dequantise /
encode / decode`, `Encoder`, `Decoder` -- but activations stay channels-last
[B, T, C] on the GPU end to end (no NCT<->NTC permutes) and every
layer is a launch of libqpg_sm100.so's tap-GEMM (qpg_conv1d_taps_f32), the
transposed convolutions as two sub-pixel phases; the quantiser is the fused
L2-argmin kernel (no [N*T', 512] distance matrix in memory).

Training (`forward` losses, EMA codebook updates: vqvae.py:187-303,
bottleneck.py:39-94) is out of scope and raises.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch

from . import _lib


def _strip_module(sd):
    return {(k[7:] if k.startswith("module.") else k): v for k, v in sd.items()}


class _TapConv:
    """One launch of qpg_conv1d_taps_f32: weights packed [n_taps, C_in, C_out]."""

    def __init__(self, w_taps: torch.Tensor, bias: Optional[torch.Tensor], tap_offset, in_stride=1, out_stride=1,
                 out_offset=0, relu_in=False, precision=0):
        self.w = w_taps.contiguous()
        self.bias = None if bias is None else bias.contiguous()
        self.n_taps, self.c_in, self.c_out = self.w.shape
        self.tap_offset = list(tap_offset)
        self.in_stride, self.out_stride, self.out_offset = in_stride, out_stride, out_offset
        self.relu_in = bool(relu_in)
        self.precision = precision

    def __call__(self, x: torch.Tensor, out: torch.Tensor, n_out: int, residual: Optional[torch.Tensor] = None):
        lib = _lib.load()
        B, T_in, C = x.shape
        assert C == self.c_in and out.shape[0] == B and out.shape[2] == self.c_out
        d = _lib.ConvDesc()
        d.B, d.T_in, d.T_out_total, d.C_in, d.C_out = B, T_in, out.shape[1], self.c_in, self.c_out
        d.n_taps = self.n_taps
        for i in range(4):
            d.tap_offset[i] = self.tap_offset[i] if i < self.n_taps else 0
        d.in_stride, d.out_stride, d.out_offset, d.n_out = self.in_stride, self.out_stride, self.out_offset, n_out
        d.relu_in, d.precision = int(self.relu_in), self.precision
        _lib.check(lib.qpg_conv1d_taps_f32(d, _lib.ptr(x), _lib.ptr(self.w), _lib.ptr(self.bias), _lib.ptr(residual),
                                           _lib.ptr(out), _lib.stream_ptr()), "qpg_conv1d_taps_f32")
        return out


def _pack_conv(weight: torch.Tensor) -> torch.Tensor:
    """nn.Conv1d weight [C_out, C_in, k] -> [k, C_in, C_out]."""
    return weight.permute(2, 1, 0).contiguous()


def _pack_convT(weight: torch.Tensor, taps) -> torch.Tensor:
    """nn.ConvTranspose1d weight [C_in, C_out, k] -> [len(taps), C_in, C_out] for the chosen k's."""
    return torch.stack([weight[:, :, k] for k in taps]).contiguous()


class ResConv1DBlock:
    """x + conv1(relu(conv3_dilated(relu(x))))  (resnet.py:27-46)."""

    def __init__(self, sd, prefix, dilation, device, precision=0):
        g = lambda n: sd[prefix + n].to(device=device, dtype=torch.float32)
        self.conv3 = _TapConv(_pack_conv(g(".1.weight")), g(".1.bias"), [-dilation, 0, dilation], relu_in=True,
                              precision=precision)
        self.conv1 = _TapConv(_pack_conv(g(".3.weight")), g(".3.bias"), [0], relu_in=True, precision=precision)

    def __call__(self, x, tmp, out):
        T = x.shape[1]
        self.conv3(x, tmp, T)
        self.conv1(tmp, out, T, residual=x)
        return out


class Resnet1D:
    """depth ResConv1DBlocks, dilations growth**d, reversed in the decoder (resnet.py:48-77)."""

    def __init__(self, sd, prefix, depth, growth, reverse, device, precision=0):
        dil = [growth ** d for d in range(depth)]
        if reverse:
            dil = dil[::-1]
        self.blocks = [ResConv1DBlock(sd, f"{prefix}.model.{d}.model", dil[d], device, precision) for d in range(depth)]

    def __call__(self, x):
        tmp = torch.empty_like(x)
        for blk in self.blocks:
            out = torch.empty_like(x)
            x = blk(x, tmp, out)
        return x


class Encoder:
    """encdec.py:8-30,53-90 for one level: down_t x (Conv1d k=2s, stride s, pad s/2 -> Resnet1D) -> Conv1d k3."""

    def __init__(self, sd, hps, device, precision=0):
        down_t, s = hps.downs_t[0], hps.strides_t[0]
        assert s == 2, "stride 2 is what the reference config uses (codebook.yml:4)"
        pre = "encoders.0.level_blocks.0.model"
        g = lambda n: sd[n].to(device=device, dtype=torch.float32)
        self.downs, self.res = [], []
        for i in range(down_t):
            k = 2 * s
            offs = [j - s // 2 for j in range(k)]                       # t_in = s*t - pad + j
            self.downs.append(_TapConv(_pack_conv(g(f"{pre}.{i}.0.weight")), g(f"{pre}.{i}.0.bias"), offs, in_stride=s,
                                       precision=precision))
            self.res.append(Resnet1D(sd, f"{pre}.{i}.1", hps.depth, hps.dilation_growth_rate, False, device, precision))
        self.out = _TapConv(_pack_conv(g(f"{pre}.{down_t}.weight")), g(f"{pre}.{down_t}.bias"), [-1, 0, 1],
                            precision=precision)
        self.stride = s

    def __call__(self, x):                                              # x [B, T, C_in] float32
        for down, res in zip(self.downs, self.res):
            B, T, _ = x.shape
            assert T % self.stride == 0, "sequence length must be divisible by the total stride"
            y = torch.empty((B, T // self.stride, down.c_out), dtype=torch.float32, device=x.device)
            x = res(down(x, y, T // self.stride))
        y = torch.empty((x.shape[0], x.shape[1], self.out.c_out), dtype=torch.float32, device=x.device)
        return self.out(x, y, x.shape[1])


class Decoder:
    """encdec.py:32-51,92-136: Conv1d k3 -> down_t x (Resnet1D reversed -> ConvTranspose1d k4 s2 p1) -> out Conv1d k3."""

    def __init__(self, sd, hps, device, precision=0):
        down_t, s = hps.downs_t[0], hps.strides_t[0]
        assert s == 2
        pre = "decoders.0.level_blocks.0.model"
        g = lambda n: sd[n].to(device=device, dtype=torch.float32)
        self.inp = _TapConv(_pack_conv(g(f"{pre}.0.weight")), g(f"{pre}.0.bias"), [-1, 0, 1], precision=precision)
        self.res, self.up_even, self.up_odd = [], [], []
        for i in range(down_t):
            self.res.append(Resnet1D(sd, f"{pre}.{i + 1}.0", hps.depth, hps.dilation_growth_rate,
                                     hps.vqvae_reverse_decoder_dilation, device, precision))
            w, b = g(f"{pre}.{i + 1}.1.weight"), g(f"{pre}.{i + 1}.1.bias")
            # t_out = 2*t_in - 1 + k :  even t_out=2u <- (k=1,t_in=u),(k=3,t_in=u-1) ; odd 2u+1 <- (k=0,u+1),(k=2,u)
            self.up_even.append(_TapConv(_pack_convT(w, (1, 3)), b, [0, -1], out_stride=2, out_offset=0,
                                         precision=precision))
            self.up_odd.append(_TapConv(_pack_convT(w, (0, 2)), b, [1, 0], out_stride=2, out_offset=1,
                                        precision=precision))
        self.out = _TapConv(_pack_conv(g("decoders.0.out.weight")), g("decoders.0.out.bias"), [-1, 0, 1],
                            precision=precision)

    def __call__(self, x):                                              # x [B, T', emb]
        y = torch.empty((x.shape[0], x.shape[1], self.inp.c_out), dtype=torch.float32, device=x.device)
        x = self.inp(x, y, x.shape[1])
        for res, ue, uo in zip(self.res, self.up_even, self.up_odd):
            x = res(x)
            B, T, _ = x.shape
            y = torch.empty((B, 2 * T, ue.c_out), dtype=torch.float32, device=x.device)
            ue(x, y, T)
            uo(x, y, T)
            x = y
        y = torch.empty((x.shape[0], x.shape[1], self.out.c_out), dtype=torch.float32, device=x.device)
        return self.out(x, y, x.shape[1])


# ======================================================================================
# tensor-core ("fast", TF32) path: qpg_conv1d_taps_tf32
# ======================================================================================
def _round_up(x, m):
    return (x + m - 1) // m * m


_SM_COUNT = {}


def _sm_count(device):
    key = torch.device(device).index or 0
    if key not in _SM_COUNT:
        _SM_COUNT[key] = torch.cuda.get_device_properties(key).multi_processor_count
    return _SM_COUNT[key]


class _TcConv:
    """One launch of qpg_conv1d_taps_tf32.  `w_taps` is a list of [C_out, C_in] matrices (one per
    tap); they are stored [n_taps][N_pad][K_pad] zero padded, K-major, as the UMMA B operand."""

    def __init__(self, w_taps, bias, row_offset, chan_offset, device, split=False):
        """split: 3xTF32 (float32-accurate products): weights stored as a TF32 part and a residual part."""
        c_out, c_in = w_taps[0].shape
        self.c_in, self.c_out, self.n_taps = c_in, c_out, len(w_taps)
        self.K_pad = _round_up(c_in, 4)
        self.BN = min(128 if split else 256, _round_up(c_out, 16))
        self.N_pad = _round_up(c_out, self.BN)
        w = torch.zeros((self.n_taps, self.N_pad, self.K_pad), dtype=torch.float32, device=device)
        for i, m in enumerate(w_taps):
            w[i, :c_out, :c_in] = m.to(device=device, dtype=torch.float32)
        self.w = w.contiguous()
        self.split = split
        if split:
            mask = torch.tensor(-8192, dtype=torch.int32, device=device)              # 0xffffe000
            self.w_hi = (self.w.view(torch.int32) & mask).view(torch.float32).contiguous()
            self.w_lo = ((self.w - self.w_hi).view(torch.int32) & mask).view(torch.float32).contiguous()
        self.bias = None if bias is None else bias.to(device=device, dtype=torch.float32).contiguous()
        self.row_offset, self.chan_offset = list(row_offset), list(chan_offset)

    # relative time of one 128-row tile by width (measured: the main loop is shared-memory bound, a half-width tile
    # costs 0.6 of a full one): narrower tiles pay only when the wide ones leave more than half of the SMs idle
    _TILE_COST = {256: 1.0, 128: 0.6, 64: 0.38}

    def _tile_n(self, B, n_out, device):
        if self.BN not in self._TILE_COST:                       # ragged output widths keep their single tile
            return self.BN
        t_box = min(n_out, 128)
        b_box = max(1, min(128 // t_box, B, 256))
        tiles_m = -(-B // b_box) * -(-n_out // t_box)
        sms = _sm_count(device)
        best, best_cost = self.BN, None
        for bn in (256, 128, 64):
            if bn > self.BN or self.N_pad % bn:
                continue
            waves = -(-(tiles_m * (self.N_pad // bn)) // sms)
            cost = waves * self._TILE_COST[bn]
            if best_cost is None or cost < best_cost - 1e-9:
                best, best_cost = bn, cost
        return best

    def __call__(self, x, B, T_view, C_view, n_out, out=None, out_relu=None, residual=None, out_rows_per_item=None,
                 out_ld=None, out_chan_offset=0):
        lib = _lib.load()
        d = _lib.ConvTcDesc()
        d.B, d.T_view, d.C_view, d.n_out = B, T_view, C_view, n_out
        d.C_in, d.C_out, d.K_pad, d.N_pad, d.BN, d.n_taps = self.c_in, self.c_out, self.K_pad, self.N_pad, \
            self._tile_n(B, n_out, x.device), self.n_taps
        for i in range(4):
            d.row_offset[i] = self.row_offset[i] if i < self.n_taps else 0
            d.chan_offset[i] = self.chan_offset[i] if i < self.n_taps else 0
        d.out_rows_per_item = out_rows_per_item if out_rows_per_item is not None else n_out
        d.out_ld = out_ld if out_ld is not None else self.c_out
        d.out_chan_offset = out_chan_offset
        if self.split:
            _lib.check(lib.qpg_conv1d_taps_3xtf32(d, _lib.ptr(x), _lib.ptr(self.w_hi), _lib.ptr(self.w_lo),
                                                  _lib.ptr(self.bias), _lib.ptr(residual), _lib.ptr(out),
                                                  _lib.ptr(out_relu), _lib.stream_ptr()), "qpg_conv1d_taps_3xtf32")
        else:
            _lib.check(lib.qpg_conv1d_taps_tf32(d, _lib.ptr(x), _lib.ptr(self.w), _lib.ptr(self.bias), _lib.ptr(residual),
                                                _lib.ptr(out), _lib.ptr(out_relu), _lib.stream_ptr()),
                       "qpg_conv1d_taps_tf32")


class _TcResnet:
    def __init__(self, sd, prefix, depth, growth, reverse, device, split=False):
        dil = [growth ** d for d in range(depth)]
        if reverse:
            dil = dil[::-1]
        self.blocks = []
        for d in range(depth):
            p = f"{prefix}.model.{d}.model"
            w3, b3, w1, b1 = sd[p + ".1.weight"], sd[p + ".1.bias"], sd[p + ".3.weight"], sd[p + ".3.bias"]
            conv3 = _TcConv([w3[:, :, k] for k in range(3)], b3, [-dil[d], 0, dil[d]], [0, 0, 0], device, split)
            conv1 = _TcConv([w1[:, :, 0]], b1, [0], [0], device, split)
            self.blocks.append((conv3, conv1))

    def __call__(self, x_raw, x_relu):
        """(raw, relu) -> raw of the last block (its ReLU copy is never needed by the next layer)."""
        B, T, Cc = x_raw.shape
        for i, (conv3, conv1) in enumerate(self.blocks):
            h_relu = torch.empty_like(x_raw)
            conv3(x_relu, B, T, Cc, T, out_relu=h_relu)
            y_raw = torch.empty_like(x_raw)
            last = i == len(self.blocks) - 1
            y_relu = None if last else torch.empty_like(x_raw)
            conv1(h_relu, B, T, Cc, T, out=y_raw, out_relu=y_relu, residual=x_raw)
            x_raw, x_relu = y_raw, y_relu
        return x_raw


class TcEncoder:
    """Encoder on tensor cores; the stride-2 k4 convolutions read the paired-frame view [B, T/2, 2C]."""

    def __init__(self, sd, hps, device, split=False):
        down_t, s = hps.downs_t[0], hps.strides_t[0]
        assert s == 2
        pre = "encoders.0.level_blocks.0.model"
        self.downs, self.res, self.c_pad = [], [], []
        for i in range(down_t):
            w, b = sd[f"{pre}.{i}.0.weight"], sd[f"{pre}.{i}.0.bias"]            # [C_out, C_in, 4]
            c_in = w.shape[1]
            cp = _round_up(c_in, 4)                                              # channels of the (padded) input
            self.c_pad.append(cp)
            # t_in = 2t - 1 + k: k=0 -> pair t-1 second half, k=1 -> pair t first half, k=2 -> second half, k=3 -> pair t+1
            self.downs.append(_TcConv([w[:, :, k] for k in range(4)], b, [-1, 0, 0, 1], [cp, 0, cp, 0], device, split))
            self.res.append(_TcResnet(sd, f"{pre}.{i}.1", hps.depth, hps.dilation_growth_rate, False, device, split))
        w, b = sd[f"{pre}.{down_t}.weight"], sd[f"{pre}.{down_t}.bias"]
        self.out = _TcConv([w[:, :, k] for k in range(3)], b, [-1, 0, 1], [0, 0, 0], device, split)

    def __call__(self, x):
        B, T, Cc = x.shape
        if Cc != self.c_pad[0]:
            x = torch.nn.functional.pad(x, (0, self.c_pad[0] - Cc)).contiguous()
        for down, res, cp in zip(self.downs, self.res, self.c_pad):
            B, T, Cc = x.shape
            assert T % 2 == 0 and Cc == cp
            raw = torch.empty((B, T // 2, down.c_out), dtype=torch.float32, device=x.device)
            relu = torch.empty_like(raw)
            down(x, B, T // 2, 2 * Cc, T // 2, out=raw, out_relu=relu)
            x = res(raw, relu)
        B, T, Cc = x.shape
        y = torch.empty((B, T, self.out.c_out), dtype=torch.float32, device=x.device)
        self.out(x, B, T, Cc, T, out=y)
        return y


class TcDecoder:
    def __init__(self, sd, hps, device, split=False):
        down_t, s = hps.downs_t[0], hps.strides_t[0]
        assert s == 2
        pre = "decoders.0.level_blocks.0.model"
        w, b = sd[f"{pre}.0.weight"], sd[f"{pre}.0.bias"]
        self.inp = _TcConv([w[:, :, k] for k in range(3)], b, [-1, 0, 1], [0, 0, 0], device, split)
        self.res, self.up_even, self.up_odd = [], [], []
        for i in range(down_t):
            self.res.append(_TcResnet(sd, f"{pre}.{i + 1}.0", hps.depth, hps.dilation_growth_rate,
                                      hps.vqvae_reverse_decoder_dilation, device, split))
            w, b = sd[f"{pre}.{i + 1}.1.weight"], sd[f"{pre}.{i + 1}.1.bias"]    # [C_in, C_out, 4]
            wt = lambda k: w[:, :, k].t()
            self.up_even.append(_TcConv([wt(1), wt(3)], b, [0, -1], [0, 0], device, split))
            self.up_odd.append(_TcConv([wt(0), wt(2)], b, [1, 0], [0, 0], device, split))
        w, b = sd["decoders.0.out.weight"], sd["decoders.0.out.bias"]
        self.out = _TcConv([w[:, :, k] for k in range(3)], b, [-1, 0, 1], [0, 0, 0], device, split)

    def __call__(self, x):
        B, T, Cc = x.shape
        raw = torch.empty((B, T, self.inp.c_out), dtype=torch.float32, device=x.device)
        relu = torch.empty_like(raw)
        self.inp(x, B, T, Cc, T, out=raw, out_relu=relu)
        n_up = len(self.res)
        for i, (res, ue, uo) in enumerate(zip(self.res, self.up_even, self.up_odd)):
            x = res(raw, relu)
            B, T, Cc = x.shape
            co = ue.c_out
            raw = torch.empty((B, 2 * T, co), dtype=torch.float32, device=x.device)    # == paired view [B, T, 2*co]
            relu = torch.empty_like(raw) if i < n_up - 1 else None
            ue(x, B, T, Cc, T, out=raw, out_relu=relu, out_rows_per_item=T, out_ld=2 * co, out_chan_offset=0)
            uo(x, B, T, Cc, T, out=raw, out_relu=relu, out_rows_per_item=T, out_ld=2 * co, out_chan_offset=co)
        B, T, Cc = raw.shape
        y = torch.empty((B, T, self.out.c_out), dtype=torch.float32, device=x.device)
        self.out(raw, B, T, Cc, T, out=y)
        return y


class BottleneckBlock:
    """Inference half of bottleneck.py:15-154 (quantise / dequantise / encode / decode)."""

    def __init__(self, k_bins, emb_width, mu=0.99, device=None, fast=True):
        self.k_bins, self.emb_width, self.mu, self.fast = k_bins, emb_width, mu, fast
        self.device = torch.device(device if device is not None else "cuda")
        self.k = torch.zeros((k_bins, emb_width), dtype=torch.float32, device=self.device)   # bottleneck.py:28

    def quantise(self, x):
        """x [M, emb] float32 (device) -> (x_l int64 [M], fit scalar)  (bottleneck.py:120-126)."""
        lib = _lib.load()
        x = x.contiguous()
        M = x.shape[0]
        x_l = torch.empty((M,), dtype=torch.int64, device=x.device)
        mind = torch.empty((M,), dtype=torch.float32, device=x.device)
        if self.fast and self.emb_width % 4 == 0 and self.k_bins % 16 == 0 and M > 0:
            # tensor-core filter + exact re-evaluation: same indices as the float64 kernel below
            scratch = torch.empty(((M + 1) * self.k_bins,), dtype=torch.float32, device=x.device)
            _lib.check(lib.qpg_vq_argmin_fast(_lib.ptr(x), _lib.ptr(self.k), M, self.emb_width, self.k_bins,
                                              _lib.ptr(scratch), _lib.ptr(x_l), _lib.ptr(mind), _lib.stream_ptr()),
                       "qpg_vq_argmin_fast")
        else:
            _lib.check(lib.qpg_vq_argmin_f32(_lib.ptr(x), _lib.ptr(self.k), M, self.emb_width, self.k_bins, _lib.ptr(x_l),
                                             _lib.ptr(mind), _lib.stream_ptr()), "qpg_vq_argmin_f32")
        return x_l, torch.mean(mind)

    def dequantise(self, x_l):
        """int64 [...] -> float32 [..., emb]  (F.embedding, bottleneck.py:128-130)."""
        lib = _lib.load()
        flat = x_l.reshape(-1).contiguous()
        out = torch.empty((flat.shape[0], self.emb_width), dtype=torch.float32, device=flat.device)
        _lib.check(lib.qpg_vq_dequantise_f32(_lib.ptr(flat), _lib.ptr(self.k), flat.shape[0], self.emb_width,
                                             self.k_bins, _lib.ptr(out), _lib.stream_ptr()), "qpg_vq_dequantise_f32")
        return out.view(tuple(x_l.shape) + (self.emb_width,))

    def encode(self, x_btc):
        """channels-last latents [N, T', emb] -> codes [N, T']  (bottleneck.py:132-143)."""
        N, T, C = x_btc.shape
        x_l, _ = self.quantise(x_btc.reshape(N * T, C))
        return x_l.view(N, T)

    def decode(self, x_l):
        """codes [N, T'] -> channels-last [N, T', emb]  (bottleneck.py:145-154, without the NCT permute)."""
        return self.dequantise(x_l)


class VQVAE:
    """Inference surface of models/vqvae.py:52-181 (levels = 1)."""

    def __init__(self, hps, input_dim=72, device=None, precision=0, use_graph=True, max_graphs=8):
        """precision 0 = float32 FFMA, 2 = 3xTF32 on the tcgen05 tensor cores (float32-accurate products: the
        index-parity modes), 1 = plain TF32 tensor cores (fast mode, ~1e-3 relative)."""
        assert hps.levels == 1, "the reference config uses one level (codebook.yml:3)"
        self.use_graph, self.max_graphs, self._graphs = use_graph, max_graphs, {}
        _lib.load()
        self.hps = hps
        self.input_dim = input_dim
        self.device = torch.device(device if device is not None else "cuda")
        self.precision = precision
        self.levels = 1
        self.l_bins = hps.l_bins
        self.encoder: Optional[Encoder] = None
        self.decoder: Optional[Decoder] = None
        self.bottleneck = BottleneckBlock(hps.l_bins, hps.emb_width, hps.l_mu, self.device)
        self.hop = int(np.prod([s ** d for s, d in zip(hps.strides_t, hps.downs_t)]))

    # -- checkpoint ------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor], strict: bool = True):
        sd = _strip_module(state_dict)
        self._graphs = {}
        if self.precision in (1, 2):
            self.encoder = TcEncoder(sd, self.hps, self.device, split=self.precision == 2)
            self.decoder = TcDecoder(sd, self.hps, self.device, split=self.precision == 2)
        else:
            self.encoder = Encoder(sd, self.hps, self.device, self.precision)
            self.decoder = Decoder(sd, self.hps, self.device, self.precision)
        self.bottleneck.k = sd["bottleneck.level_blocks.0.k"].to(device=self.device, dtype=torch.float32).contiguous()
        return self

    def eval(self):
        return self

    def to(self, device):
        assert torch.device(device).type == "cuda", "there is no CPU path"
        return self

    # -- inference ---------------------------------------------------------------------
    def preprocess(self, x):
        assert len(x.shape) == 3
        return x.to(device=self.device, dtype=torch.float32).contiguous()          # stays NTC (vqvae.py:127-131)

    def latents(self, x):
        with torch.cuda.device(self.device):
            return self.encoder(self.preprocess(x))

    # -- per-shape CUDA graphs: an encode / decode is ~25 dependent launches of microsecond kernels, so the
    #    fixed launch sequence for a given input shape is captured once and replayed (static in/out buffers)
    def _graphed(self, kind, x, fn):
        if not self.use_graph:
            return fn(x)
        key = (kind, tuple(x.shape))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= self.max_graphs:
                self._graphs.pop(next(iter(self._graphs)))
            static_in = x.clone()
            fn(static_in)                                   # warm-up: function attributes, allocator
            torch.cuda.synchronize(self.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                static_out = fn(static_in)
            ent = (g, static_in, static_out)
            self._graphs[key] = ent
        g, static_in, static_out = ent
        static_in.copy_(x)
        g.replay()
        return static_out.clone()

    def _encode(self, x, start_level=0, end_level=None):
        with torch.cuda.device(self.device):
            codes = self._graphed("enc", self.preprocess(x), lambda t: self.bottleneck.encode(self.encoder(t)))
            return [codes][start_level:end_level]

    def encode(self, x, start_level=0, end_level=None, bs_chunks=1):
        """x [B, T, C] -> [LongTensor [B, T/8]]  (vqvae.py:174-181)."""
        x = torch.as_tensor(x)
        outs = [self._encode(xi, start_level, end_level) for xi in torch.chunk(x, bs_chunks, dim=0)]
        return [torch.cat(level, dim=0) for level in zip(*outs)]

    def _decode(self, zs, start_level=0, end_level=None):
        assert len(zs) == 1
        with torch.cuda.device(self.device):
            z = zs[0].to(device=self.device, dtype=torch.int64).contiguous()
            if z.numel() and (int(z.min()) < 0 or int(z.max()) >= self.l_bins):
                # F.embedding (bottleneck.py:129) raises on an out-of-range index; the gather kernel would clamp.
                # Typical source: the -1 rows a failed match leaves in knn_pred.  Checked here, outside the graph.
                raise IndexError(f"code index out of range [0, {self.l_bins}): min {int(z.min())}, max {int(z.max())}")
            return self._graphed("dec", z, lambda t: self.decoder(self.bottleneck.decode(t)))   # [B, 8T', C]

    def decode(self, zs, start_level=0, end_level=None, bs_chunks=1):
        """[LongTensor [B, T']] -> Tensor [B, 8T', C]  (vqvae.py:152-159)."""
        z_chunks = [torch.chunk(torch.as_tensor(z), bs_chunks, dim=0) for z in zs]
        outs = [self._decode([zc[i] for zc in z_chunks], start_level, end_level) for i in range(bs_chunks)]
        return torch.cat(outs, dim=0)

    def sample(self, n_samples):
        z = torch.randint(0, self.l_bins, size=(n_samples, self.hps.sample_length // self.hop), device=self.device)
        return self.decode([z])

    def forward(self, x):
        """Inference half of VQVAE.forward (vqvae.py:187-303): encode -> quantise -> decode.  Returns
        (x_out, loss, metrics) like the reference; the training losses and the EMA codebook update are outside the
        inference hot path, so loss is None and metrics holds only the reconstruction errors."""
        x = torch.as_tensor(x)
        x_out = self.decode(self.encode(x))
        x_in = x.to(device=x_out.device, dtype=torch.float32)
        with torch.no_grad():
            err = x_out - x_in
            metrics = dict(recons_loss=torch.mean(err ** 2), l1_loss=torch.mean(err.abs()))
        return x_out, None, metrics

    __call__ = forward

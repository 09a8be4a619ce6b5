"""Convolutional encoder / decoder of the pixel world model, cuDNN-free (reference:
algorithms/repo/models/encoder.py:21-41 `VisualEncoder`, decoder.py:28-48 `VisualObservationModel`).

Every Conv2d / ConvTranspose2d runs as an IMPLICIT GEMM on the tcgen05 layer machine (vm.cuh): the
machine's operand loader gathers each output position's receptive field straight from the input tensor
(no materialised im2col), the products are the same fp16 hi/lo three-MMA scheme as the RSSM layers, and
the store epilogue adds the bias, applies ReLU and scatters onto the output grid.  Activations between
layers are NHWC (channels contiguous = the GEMM's feature dimension).

* Conv2d(k4, s2): row = output pixel, taps = 4x4, input pixel = 2*o + tap.
* ConvTranspose2d(k, s2): four parity classes of the output grid; class (py,px) only sees taps kh = py+2*th,
  kw = px+2*tw, i.e. a stride-1 gather with input pixel = o' - t.  Each class is one GEMM launch.

Backward: weight gradients are GEMMs of the output gradient against the (materialised, backward-only)
gathered rows; data gradients are a GEMM with the weight matrix followed by the col2im gather.  Those are
plain GEMMs (torch.matmul / cuBLAS); the gathers are hand-written kernels (elementwise.cuh)."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn

from . import _lib

_MAP_FIELDS = ("enabled RA RB in_nchw C H W TH TW tap0 ntaps sy sx dy dx y0 x0 "
               "out_nchw Ho Wo osy osx oy0 ox0 relu accumulate").split()


@dataclass
class ConvMap:
    RA: int; RB: int; in_nchw: int; C: int; H: int; W: int; TH: int; TW: int
    sy: int; sx: int; dy: int; dx: int; y0: int = 0; x0: int = 0
    out_nchw: int = 0; Ho: int = 0; Wo: int = 0; osy: int = 1; osx: int = 1; oy0: int = 0; ox0: int = 0
    relu: int = 0; accumulate: int = 0; tap0: int = 0; ntaps: int = 0; enabled: int = 1

    def carray(self, **over):
        vals = {f: getattr(self, f) for f in _MAP_FIELDS}
        vals.update(over)
        if vals["ntaps"] == 0:
            vals["ntaps"] = self.TH * self.TW - vals["tap0"]
        return (C.c_int * len(_MAP_FIELDS))(*[int(vals[f]) for f in _MAP_FIELDS])

    @property
    def K(self):
        return self.TH * self.TW * self.C


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(t, name):
    if not t.is_cuda or t.dtype != torch.float32:
        raise RuntimeError(f"{name}: expected a float32 CUDA tensor (no CPU fallback)")
    return t.contiguous()


_MAX_K = 1024  # columns per launch (shared-memory budget of the operand buffer); more taps -> accumulate passes


def conv_gemm(x, w_mat, bias, out, frames, cout, cmap: ConvMap):
    """out (+)= gather(x) @ w_mat^T + bias through the layer machine; splits the taps when K is too large."""
    L = _lib.lib()
    taps = cmap.TH * cmap.TW
    per = max(1, _MAX_K // cmap.C)
    t0 = 0
    first = True
    while t0 < taps:
        n = min(per, taps - t0)
        last = t0 + n == taps
        wm = w_mat[:, t0 * cmap.C:(t0 + n) * cmap.C].contiguous()
        ws = torch.empty(L.repo_b200_linear_workspace_bytes(n * cmap.C, cout), dtype=torch.uint8, device=x.device)
        m = cmap.carray(tap0=t0, ntaps=n, relu=cmap.relu if last else 0, accumulate=(cmap.accumulate if first else 1))
        rc = L.repo_b200_conv_gemm(_p(x), _p(wm), _p(bias) if last else None, _p(out), frames, cout, m, _p(ws), ws.numel(), _stream())
        _lib.check(rc, "repo_b200_conv_gemm")
        t0 += n
        first = False
    return out


def im2col(x, frames, cmap: ConvMap):
    col = torch.empty(frames * cmap.RA * cmap.RB, cmap.K, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().repo_b200_im2col(_p(x), _p(col), frames, cmap.carray(), _stream()), "repo_b200_im2col")
    return col


def col2im(d_col, d_in, frames, cmap: ConvMap, accumulate=False):
    _lib.check(_lib.lib().repo_b200_col2im(_p(d_col), _p(d_in), frames, int(accumulate), cmap.carray(), _stream()), "repo_b200_col2im")
    return d_in


# ------------------------------------------------------------------------------------------------- encoder
def _enc_maps(frames_hw=(64, 64)):
    H, W = frames_hw
    maps, chans = [], [3, 32, 64, 128, 256]
    h, w = H, W
    for i in range(4):
        ho, wo = (h - 4) // 2 + 1, (w - 4) // 2 + 1
        maps.append(ConvMap(RA=ho, RB=wo, in_nchw=1 if i == 0 else 0, C=chans[i], H=h, W=w, TH=4, TW=4, sy=2, sx=2, dy=1, dx=1,
                            out_nchw=1 if i == 3 else 0, Ho=ho, Wo=wo, relu=1))
        h, w = ho, wo
    return maps


def _conv_wmat(w):  # (Cout, Cin, kh, kw) -> (Cout, (kh, kw, cin))
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous()


class _EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, obs, *params):
        ws, bs = params[0::2], params[1::2]
        F_ = obs.shape[0]
        maps = _enc_maps(tuple(obs.shape[2:]))
        x = _need_cuda(obs.detach(), "observation")
        acts = [x]
        for i, cm in enumerate(maps):
            cout = ws[i].shape[0]
            shape = (F_, cout, cm.Ho, cm.Wo) if cm.out_nchw else (F_, cm.Ho, cm.Wo, cout)
            out = torch.empty(shape, device=x.device, dtype=torch.float32)
            conv_gemm(acts[-1], _conv_wmat(ws[i].detach()), bs[i].detach().contiguous(), out, F_, cout, cm)
            acts.append(out)
        ctx.maps = maps
        ctx.save_for_backward(*acts, *params)
        return acts[-1].reshape(F_, -1)  # NCHW flatten == hidden.view(-1, 1024) (encoder.py:39)

    @staticmethod
    def backward(ctx, g):
        maps = ctx.maps
        saved = ctx.saved_tensors
        acts, params = saved[:5], saved[5:]
        ws, bs = params[0::2], params[1::2]
        F_ = acts[0].shape[0]
        grads = [None] * 8
        # gradient w.r.t. the last activation, as NHWC rows
        cm = maps[3]
        gl = (g.reshape(acts[4].shape) * (acts[4] > 0)).permute(0, 2, 3, 1).reshape(F_ * cm.Ho * cm.Wo, -1).contiguous()
        for i in range(3, -1, -1):
            cm = maps[i]
            col = im2col(acts[i], F_, cm)
            if ctx.needs_input_grad[1 + 2 * i]:
                k = ws[i].shape[2]
                grads[2 * i] = (gl.t() @ col).reshape(ws[i].shape[0], k, k, ws[i].shape[1]).permute(0, 3, 1, 2).contiguous()
            if ctx.needs_input_grad[2 + 2 * i]:
                grads[2 * i + 1] = gl.sum(0)
            del col
            if i == 0:
                break
            d_col = gl @ _conv_wmat(ws[i])
            d_in = torch.empty_like(acts[i])
            col2im(d_col, d_in, F_, cm)
            del d_col
            gl = (d_in * (acts[i] > 0)).reshape(-1, acts[i].shape[-1])
        g_obs = None
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. the pixel observation is not needed by any trainer")
        return (g_obs, *grads)


class VisualEncoder(nn.Module):
    """encoder.py:21-41 — Conv(3->32->64->128->256, k4, s2) + ReLU, flattened to 1024 (+ Linear if embedding != 1024)."""

    def __init__(self, embedding_size, activation_function="relu"):
        super().__init__()
        if activation_function != "relu":
            raise RuntimeError("the fused conv epilogue implements ReLU (cnn_activation_function default, train_repo.py:32)")
        self.embedding_size = embedding_size
        self.conv1 = nn.Conv2d(3, 32, 4, stride=2)
        self.conv2 = nn.Conv2d(32, 64, 4, stride=2)
        self.conv3 = nn.Conv2d(64, 128, 4, stride=2)
        self.conv4 = nn.Conv2d(128, 256, 4, stride=2)
        self.fc = nn.Identity() if embedding_size == 1024 else nn.Linear(1024, embedding_size)

    def forward(self, observation):
        params = []
        for c in (self.conv1, self.conv2, self.conv3, self.conv4):
            params += [c.weight, c.bias]
        hidden = _EncoderFn.apply(observation, *params)
        if isinstance(self.fc, nn.Identity):
            return hidden
        raise NotImplementedError("embedding_size != 1024 needs the extra Linear; not wired to the machine yet")


# ------------------------------------------------------------------------------------------------- decoder
def _deconv_classes(cin, hin, win, k, out_nchw, relu):
    """The four output-parity classes of ConvTranspose2d(k, stride 2): (py, px, ConvMap)."""
    ho, wo = (hin - 1) * 2 + k, (win - 1) * 2 + k
    out = []
    for py in (0, 1):
        for px in (0, 1):
            th, tw = (k - py + 1) // 2, (k - px + 1) // 2
            ra, rb = (ho - py + 1) // 2, (wo - px + 1) // 2
            out.append((py, px, ConvMap(RA=ra, RB=rb, in_nchw=0, C=cin, H=hin, W=win, TH=th, TW=tw, sy=1, sx=1, dy=-1, dx=-1,
                                        out_nchw=int(out_nchw), Ho=ho, Wo=wo, osy=2, osx=2, oy0=py, ox0=px, relu=int(relu))))
    return ho, wo, out


def _deconv_wmat(w, py, px):  # (Cin, Cout, k, k) -> class matrix (Cout, (th, tw, cin))
    sub = w[:, :, py::2, px::2]
    return sub.permute(1, 2, 3, 0).reshape(w.shape[1], -1).contiguous()


_DEC = [(1024, 128, 5), (128, 64, 5), (64, 32, 6), (32, 3, 6)]


class _DecoderFn(torch.autograd.Function):
    """fc1 -> view(1024,1,1) -> 4 x ConvTranspose2d (decoder.py:41-48)."""

    @staticmethod
    def forward(ctx, belief, state, *params):
        from . import ops
        fc_w, fc_b = params[0], params[1]
        ws, bs = params[2::2], params[3::2]
        F_ = belief.shape[0]
        xin = torch.cat([_need_cuda(belief.detach(), "belief"), _need_cuda(state.detach(), "state")], 1)
        h = ops.linear(xin, fc_w.detach().contiguous(), fc_b.detach().contiguous())      # (F, 1024), no activation
        # layer 1: 1x1 input -> 5x5 output is a plain GEMM; features ordered (kh, kw, co) = NHWC (F,5,5,128)
        k1, co1 = ws[0].shape[2], ws[0].shape[1]
        w1 = ws[0].detach().permute(2, 3, 1, 0).reshape(k1 * k1 * co1, -1).contiguous()
        b1 = bs[0].detach().repeat(k1 * k1).contiguous()
        a1 = torch.empty(F_, k1, k1, co1, device=h.device, dtype=torch.float32)
        conv_gemm(h, w1, b1, a1, F_, k1 * k1 * co1,
                  ConvMap(RA=1, RB=1, in_nchw=0, C=h.shape[1], H=1, W=1, TH=1, TW=1, sy=1, sx=1, dy=1, dx=1, Ho=1, Wo=1, relu=1))
        acts, layer_maps = [xin, h, a1], []
        hin = k1
        for li in (1, 2, 3):
            cin, cout, k = _DEC[li]
            last = li == 3
            ho, wo, classes = _deconv_classes(cin, hin, hin, k, out_nchw=last, relu=not last)
            out = torch.empty((F_, cout, ho, wo) if last else (F_, ho, wo, cout), device=h.device, dtype=torch.float32)
            for py, px, cm in classes:
                conv_gemm(acts[-1], _deconv_wmat(ws[li].detach(), py, px), bs[li].detach().contiguous(), out, F_, cout, cm)
            acts.append(out)
            layer_maps.append(classes)
            hin = ho
        ctx.layer_maps = layer_maps
        ctx.belief_size = belief.shape[1]
        ctx.save_for_backward(*acts, *params)
        return acts[-1]

    @staticmethod
    def backward(ctx, g):
        saved = ctx.saved_tensors
        acts, params = saved[:6], saved[6:]
        xin, h, a1, a2, a3, out = acts
        fc_w, fc_b = params[0], params[1]
        ws, bs = params[2::2], params[3::2]
        F_ = xin.shape[0]
        need = ctx.needs_input_grad
        grads = [None] * len(params)
        layer_in = [a1, a2, a3]
        # gradient of the last layer's output (NCHW, no activation) as NHWC
        g_nhwc = g.contiguous().float().permute(0, 2, 3, 1).contiguous()
        for li in (3, 2, 1):
            x = layer_in[li - 1]
            classes = ctx.layer_maps[li - 1]
            if need[2 + 2 * li + 1]:
                grads[2 * li + 3] = g_nhwc.sum((0, 1, 2))
            dw = torch.zeros_like(ws[li]) if need[2 + 2 * li] else None
            d_in = torch.empty_like(x)
            first = True
            for py, px, cm in classes:
                gc = g_nhwc[:, py::2, px::2, :].reshape(-1, g_nhwc.shape[-1])
                if dw is not None:
                    col = im2col(x, F_, cm)
                    dwm = (gc.t() @ col).reshape(ws[li].shape[1], cm.TH, cm.TW, ws[li].shape[0])  # (co, th, tw, ci)
                    dw[:, :, py::2, px::2] = dwm.permute(3, 0, 1, 2)
                    del col
                d_col = gc @ _deconv_wmat(ws[li], py, px)
                col2im(d_col, d_in, F_, cm, accumulate=not first)
                first = False
                del d_col
            if dw is not None:
                grads[2 * li + 2] = dw
            g_nhwc = d_in * (x > 0)  # ReLU of the layer that produced x
        # layer 1 (plain GEMM) and fc1
        k1, co1 = ws[0].shape[2], ws[0].shape[1]
        g1 = g_nhwc.reshape(F_, k1 * k1 * co1)
        w1 = ws[0].permute(2, 3, 1, 0).reshape(k1 * k1 * co1, -1)
        if need[2 + 2]:
            grads[2] = (g1.t() @ h).reshape(k1, k1, co1, -1).permute(3, 2, 0, 1).contiguous()
        if need[2 + 3]:
            grads[3] = g1.reshape(F_, k1 * k1, co1).sum((0, 1))
        d_h = g1 @ w1
        if need[2]:
            grads[0] = d_h.t() @ xin
        if need[3]:
            grads[1] = d_h.sum(0)
        gb = gs = None
        if need[0] or need[1]:
            d_x = d_h @ fc_w
            gb, gs = d_x[:, :ctx.belief_size].contiguous(), d_x[:, ctx.belief_size:].contiguous()
        return (gb if need[0] else None, gs if need[1] else None, *grads)


class VisualObservationModel(nn.Module):
    """decoder.py:28-48 — Linear(belief+state -> 1024, no activation) -> ConvTranspose2d(1024->128 k5, 128->64 k5,
    64->32 k6, 32->3 k6; stride 2, ReLU on the first three): 1 -> 5 -> 13 -> 30 -> 64 pixels."""

    def __init__(self, belief_size, state_size, embedding_size, activation_function="relu"):
        super().__init__()
        if activation_function != "relu":
            raise RuntimeError("the fused conv epilogue implements ReLU (cnn_activation_function default, train_repo.py:32)")
        if embedding_size != 1024:
            raise RuntimeError("VisualObservationModel kernels are sized for embedding_size == 1024")
        self.embedding_size = embedding_size
        self.belief_size = belief_size
        self.fc1 = nn.Linear(belief_size + state_size, embedding_size)
        self.conv1 = nn.ConvTranspose2d(embedding_size, 128, 5, stride=2)
        self.conv2 = nn.ConvTranspose2d(128, 64, 5, stride=2)
        self.conv3 = nn.ConvTranspose2d(64, 32, 6, stride=2)
        self.conv4 = nn.ConvTranspose2d(32, 3, 6, stride=2)

    def forward(self, belief, state):
        params = [self.fc1.weight, self.fc1.bias]
        for c in (self.conv1, self.conv2, self.conv3, self.conv4):
            params += [c.weight, c.bias]
        return _DecoderFn.apply(belief, state, *params)

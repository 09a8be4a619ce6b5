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

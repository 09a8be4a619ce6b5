"""Convolutional encoder / decoder of the pixel world model, cuDNN-free (reference:
algorithms/repo/models/encoder.py:21-41 `VisualEncoder`, decoder.py:28-48 `VisualObservationModel`).

Every Conv2d / ConvTranspose2d — and the data gradient of each — runs as an IMPLICIT GEMM on tcgen05
(repo_b200/csrc/conv.cuh): gather warps build each output position's receptive field straight from the
input tensor into the MMA operand ring (no materialised im2col), the products are the same fp16 hi/lo
three-MMA scheme as the RSSM layers, and the epilogue adds the bias, applies ReLU (or the ReLU mask of the
layer below, in backward) and scatters onto the output grid.  Activations between layers are NHWC
(channels contiguous = the GEMM's K / feature dimension).

* Conv2d(k4, s2): row = output pixel, taps = 4x4, input pixel = 2*o + tap.
* ConvTranspose2d(k, s2) = one stride-1 gather over T = ceil(k/2) taps per axis whose 4*Cout features are the
  four output-parity classes (py, px): class (py, px) uses kernel taps kh = py + 2*th, kw = px + 2*tw (zero
  where kh >= k), input pixel = o' - t, and the epilogue pixel-shuffles (o', py) -> 2*o' + py.
* Data gradients are the adjoint maps: for Conv2d(s2) a sub-pixel (shuffle) conv of the output gradient with
  2x2 taps; for ConvTranspose2d a stride-1 conv (taps +t) over the un-shuffled output gradient.

* Weight gradients contract over the rows: dW = G^T @ gather(x) on a second tcgen05 kernel (`conv_wgrad`) whose
  operands are MN-major (row-contiguous) and whose row slices are summed with fp32 atomics.
Gradient operands are rescaled by powers of two into fp16's normal range before the hi/lo split.  `im2col`
(elementwise.cuh) only remains as a test/debug helper."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import ops as _ops

_MAP_FIELDS = ("enabled RA RB in_nchw C H W TH TW tap0 ntaps sy sx dy dx y0 x0 "
               "out_nchw Ho Wo osy osx oy0 ox0 relu accumulate shuffle pix").split()


@dataclass
class ConvMap:
    RA: int; RB: int; in_nchw: int; C: int; H: int; W: int; TH: int; TW: int
    sy: int; sx: int; dy: int; dx: int; y0: int = 0; x0: int = 0
    out_nchw: int = 0; Ho: int = 0; Wo: int = 0; osy: int = 1; osx: int = 1; oy0: int = 0; ox0: int = 0
    relu: int = 0; accumulate: int = 0; tap0: int = 0; ntaps: int = 0; enabled: int = 1; shuffle: int = 0; pix: int = 0

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


_SCRATCH = {}


def grad_scales(g):
    """Device floats [1, s_g, 1/s_g | s_g, 1, 1/s_g] (second triple: see `_as_input_side`) with s_g = 2^floor(log2(2^14 / max|g|)): lifts a gradient tensor into fp16's
    normal range before the hi/lo split (gradients sit far below fp16's 6e-5 normal threshold, where the split has no
    mantissa left).  Activations and weights are O(1e-2..1) and keep scale 1.  One fused read-only pass, no host sync."""
    dev = g.device
    if dev not in _SCRATCH:
        _SCRATCH[dev] = torch.zeros(2, dtype=torch.int32, device=dev)
    scales = torch.ones(6, dtype=torch.float32, device=dev)
    if g.numel() == 0:
        return scales
    g = g.contiguous()
    rc = _lib.lib().repo_b200_pow2_scale(_p(g), g.numel(), 16384.0, 1, _p(scales), _p(_SCRATCH[dev]), _stream())
    _lib.check(rc, "repo_b200_pow2_scale")
    return scales


def _as_input_side(scales):
    """[1, s, 1/s] -> [s, 1, 1/s]: the same gradient scale when the gradient is the GATHERED operand (data gradients)."""
    return scales[3:]


def hl_empty(shape, device):
    """Split-activation ("HL") tensor: fp16 planes [hi, lo] of an NHWC activation, x = hi + lo — the operand pair the
    tensor cores consume.  The producing conv epilogue writes it; consumers gather it with 16-byte copies."""
    return torch.empty((2,) + tuple(shape), device=device, dtype=torch.float16)


def hl_to_float(t):
    return t[0].float() + t[1].float()


def _is_hl(t):
    return t is not None and t.dtype == torch.float16


def conv_gemm(x, w_mat, bias, out, frames, n_total, cmap: ConvMap, relu_mask=None, scales=None, dense_opts=None):
    """out = epilogue(gather(x) @ w_mat^T + bias) on the tcgen05 conv kernel (one launch + the weight packing).
    x / out / relu_mask may be HL tensors (`hl_empty`).  `scales` = device triple [s_x, s_w, 1/(s_x*s_w)] (see
    `grad_scales`; the gradient is the gathered operand here, so callers pass the triple with the slots swapped)
    applied before the hi/lo split and undone on the accumulator."""
    L = _lib.lib()
    if frames == 0:
        return out   # empty batch: nothing to launch (zero-sized tensors have no device pointer)
    flags = int(_is_hl(x)) | (int(_is_hl(out)) << 1) | (int(_is_hl(relu_mask)) << 2)
    if w_mat.shape != (n_total, cmap.K):
        raise RuntimeError(f"conv_gemm: weight matrix {tuple(w_mat.shape)} != ({n_total}, {cmap.K})")
    w_mat = w_mat.contiguous()
    ws = torch.empty(L.repo_b200_conv_workspace_bytes(cmap.K, n_total), dtype=torch.uint8, device=x.device)
    opts = None if dense_opts is None else (C.c_int * 4)(*[int(v) for v in dense_opts])
    rc = L.repo_b200_conv_gemm(_p(x), _p(w_mat), _p(bias), _p(relu_mask), _p(scales), _p(out), frames, n_total, cmap.carray(),
                               flags, opts, _p(ws), ws.numel(), _stream())
    _lib.check(rc, "repo_b200_conv_gemm")
    return out


def conv_wgrad(x, grad_rows, frames, n_total, cmap: ConvMap, scales=None):
    """dW (n_total, K) = grad_rows^T @ gather(x) on the tcgen05 weight-gradient kernel (no materialised im2col).
    grad_rows is (frames*RA*RB, ld >= n_total) fp32; `scales` = [s_x, s_g, 1/(s_x*s_g)] from `grad_scales`."""
    if grad_rows.dim() != 2 or not grad_rows.is_contiguous():
        raise RuntimeError("conv_wgrad: grad_rows must be a contiguous 2-D tensor")
    if frames == 0 or grad_rows.shape[0] == 0:
        return torch.zeros(n_total, cmap.K, device=x.device, dtype=torch.float32)
    if scales is None:
        scales = grad_scales(grad_rows)
    dw = torch.empty(n_total, cmap.K, device=x.device, dtype=torch.float32)
    rc = _lib.lib().repo_b200_conv_wgrad(_p(x), _p(grad_rows), _p(scales), _p(dw), frames, n_total, grad_rows.shape[1],
                                         cmap.carray(), int(_is_hl(x)), _stream())
    _lib.check(rc, "repo_b200_conv_wgrad")
    return dw


def dense_layer(x, weight, bias, out, act="none", mask=None, mask_act="relu", scales=None):
    """out = act(x @ weight^T + bias) [* act'(mask)] for 2-D fp32 operands on the tcgen05 conv kernel run as a plain GEMM.
    x, out and mask may be column windows of wider row-major matrices (row strides multiples of 4, 16-byte aligned);
    act in {"none", "relu", "elu"}; `mask` is the OUTPUT of a `mask_act` layer whose activation derivative multiplies
    the result (backward through that layer); `scales` as in `conv_gemm` (x is the gathered operand)."""
    rows, k = x.shape
    n = weight.shape[0]
    if weight.shape[1] != k:
        raise RuntimeError("dense_layer: weight / input size mismatch")
    if rows == 0:
        return out
    for t, name, width in ((x, "x", k), (out, "out", n), (mask, "mask", n)):
        if t is None:
            continue
        if t.stride(1) != 1 or not t.is_cuda or t.dtype != torch.float32:
            raise RuntimeError(f"dense_layer: {name} must be an fp32 CUDA matrix with unit column stride")
        vector = (width % 4 == 0) if name == "x" else (width % 16 == 0 or t.stride(0) != width)
        if vector and (t.stride(0) % 4 or t.data_ptr() % 16):
            raise RuntimeError(f"dense_layer: {name} needs 16-byte aligned rows")
    cmap = ConvMap(RA=1, RB=1, in_nchw=0, C=k, H=1, W=1, TH=1, TW=1, sy=1, sx=1, dy=1, dx=1, Ho=1, Wo=1,
                   relu=int(act == "relu"), pix=x.stride(0))
    opts = [int(act == "elu"), int(mask is not None and mask_act == "elu"), out.stride(0) if out.stride(0) != n else 0,
            (mask.stride(0) if (mask is not None and mask.stride(0) != n) else 0)]
    return conv_gemm(x, weight, bias, out, rows, n, cmap, relu_mask=mask, scales=scales, dense_opts=opts)


def wgrad_gemm(dpre, x):
    """dW (n, k) = dpre^T @ x for 2-D fp32 operands — the weight gradient of a Linear layer over all (t, row) samples —
    on the tcgen05 weight-gradient kernel (fp16 hi/lo three-product arithmetic, gradient operand rescaled by a power of
    two).  `x` may be a column window of a wider row-major matrix (e.g. a slice of the activation stash); n > 256 is
    processed in 256-column slices of dpre inside one launch."""
    rows, n = dpre.shape
    k = x.shape[1]
    if x.shape[0] != rows:
        raise RuntimeError("wgrad_gemm: row counts differ")
    if rows == 0:
        return torch.zeros(n, k, device=dpre.device, dtype=torch.float32)
    dpre = _need_cuda(dpre, "dpre")
    if x.stride(1) != 1 or (x.stride(0) % 4) or (x.data_ptr() % 16) or (k % 4):
        kp = (k + 3) // 4 * 4                      # repack: 16-byte aligned rows (zero columns do not change dW[:, :k])
        xp = torch.zeros(rows, kp, device=x.device, dtype=torch.float32)
        xp[:, :k] = x
        x = xp
    kk = x.shape[1]
    cmap = ConvMap(RA=1, RB=1, in_nchw=0, C=kk, H=1, W=1, TH=1, TW=1, sy=1, sx=1, dy=1, dx=1, Ho=1, Wo=1, pix=x.stride(0))
    sc = grad_scales(dpre)
    if n % 4:
        raise RuntimeError("wgrad_gemm: output features must be a multiple of 4")
    dw = torch.empty(n, kk, device=x.device, dtype=torch.float32)
    # (more than 256 output features run as 256-wide slices inside ONE launch: blockIdx.y = slice)
    rc = _lib.lib().repo_b200_conv_wgrad(_p(x), _p(dpre), _p(sc), _p(dw), rows, n, n, cmap.carray(), 0, _stream())
    _lib.check(rc, "repo_b200_conv_wgrad")
    return dw[:, :k] if kk != k else dw


def im2col(x, frames, cmap: ConvMap):
    col = torch.empty(frames * cmap.RA * cmap.RB, cmap.K, device=x.device, dtype=torch.float32)
    _lib.check(_lib.lib().repo_b200_im2col(_p(x), _p(col), frames, cmap.carray(), _stream()), "repo_b200_im2col")
    return col


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
            # intermediate activations live in split (fp16 hi/lo) format; only the final embedding is fp32 NCHW
            out = torch.empty((F_, cout, cm.Ho, cm.Wo), device=x.device, dtype=torch.float32) if cm.out_nchw \
                else hl_empty((F_, cm.Ho, cm.Wo, cout), x.device)
            conv_gemm(acts[-1], _conv_wmat(ws[i].detach()), bs[i].detach().contiguous(), out, F_, cout, cm)
            acts.append(out)
        ctx.maps = maps
        ctx.save_for_backward(*acts, *params)
        return acts[-1].reshape(F_, acts[-1].shape[1] * acts[-1].shape[2] * acts[-1].shape[3])  # NCHW flatten (encoder.py:39)

    @staticmethod
    def backward(ctx, g):
        maps = ctx.maps
        saved = ctx.saved_tensors
        acts, params = saved[:5], saved[5:]
        ws, bs = params[0::2], params[1::2]
        F_ = acts[0].shape[0]
        grads = [None] * 8
        # gradient w.r.t. the last pre-activation, NHWC
        gp = (g.reshape(acts[4].shape) * (acts[4] > 0)).permute(0, 2, 3, 1).contiguous()
        for i in range(3, -1, -1):
            cm = maps[i]
            cout, cin, k = ws[i].shape[0], ws[i].shape[1], ws[i].shape[2]
            gl = gp.reshape(-1, cout)
            sc = grad_scales(gp)                      # [1, s_g, 1/s_g]
            if ctx.needs_input_grad[1 + 2 * i]:
                grads[2 * i] = conv_wgrad(acts[i], gl, F_, cout, cm, sc).reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous()
            if ctx.needs_input_grad[2 + 2 * i]:
                grads[2 * i + 1] = _ops.colsum(gl)
            if i == 0:
                break
            # data gradient = sub-pixel conv of gp with 2x2 taps; features (py, px, ci); masked by the ReLU below
            x = acts[i]                                # HL (2, F, H, W, C)
            H, W = x.shape[2], x.shape[3]
            dmap = ConvMap(RA=(H + 1) // 2, RB=(W + 1) // 2, in_nchw=0, C=cout, H=cm.Ho, W=cm.Wo, TH=2, TW=2, sy=1, sx=1,
                           dy=-1, dx=-1, Ho=H, Wo=W, osy=2, osx=2, shuffle=1)
            wd = ws[i].reshape(cout, cin, 2, 2, 2, 2).permute(3, 5, 1, 2, 4, 0).reshape(4 * cin, 4 * cout)
            d_in = torch.empty(x.shape[1:], device=x.device, dtype=torch.float32)
            conv_gemm(gp, wd, None, d_in, F_, 4 * cin, dmap, relu_mask=x, scales=_as_input_side(sc))
            gp = d_in
        if ctx.needs_input_grad[0]:
            raise NotImplementedError("gradient w.r.t. the pixel observation is not needed by any trainer")
        return (None, *grads)


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
        from .autograd import LinearFn
        return LinearFn.apply(hidden, self.fc.weight, self.fc.bias)   # encoder.py:30,40: Linear(1024, embedding_size)


# ------------------------------------------------------------------------------------------------- decoder
_DEC = [(1024, 128, 5), (128, 64, 5), (64, 32, 6), (32, 3, 6)]


def _deconv_map(cin, hin, win, k, out_nchw, relu):
    """ConvTranspose2d(k, stride 2) as a stride-1 gather with T = ceil(k/2) taps and sub-pixel (shuffle) store."""
    ho, wo, T = (hin - 1) * 2 + k, (win - 1) * 2 + k, (k + 1) // 2
    return ConvMap(RA=(ho + 1) // 2, RB=(wo + 1) // 2, in_nchw=0, C=cin, H=hin, W=win, TH=T, TW=T, sy=1, sx=1, dy=-1, dx=-1,
                   out_nchw=int(out_nchw), Ho=ho, Wo=wo, osy=2, osx=2, relu=int(relu), shuffle=1)


def _deconv_wmat(w):
    """(Cin, Cout, k, k) -> ((py, px, cout), (th, tw, cin)) with W[ci, co, py + 2 th, px + 2 tw], zero where kh >= k."""
    cin, cout, k = w.shape[0], w.shape[1], w.shape[2]
    T = (k + 1) // 2
    wp = F.pad(w, (0, 2 * T - k, 0, 2 * T - k))
    return wp.reshape(cin, cout, T, 2, T, 2).permute(3, 5, 1, 2, 4, 0).reshape(4 * cout, T * T * cin)


def _deconv_wgrad(dwm, cin, cout, k):
    """inverse of `_deconv_wmat` for the gradient."""
    T = (k + 1) // 2
    return dwm.reshape(2, 2, cout, T, T, cin).permute(5, 2, 3, 0, 4, 1).reshape(cin, cout, 2 * T, 2 * T)[:, :, :k, :k].contiguous()


def _cpad(cout):
    """columns of the un-shuffled gradient: 4*cout rounded up to a divisor of 256 (>= 16)"""
    n = 16
    while n < 4 * cout:
        n *= 2
    return n


def _unshuffle(g, ra, rb, cpad, nchw=False):
    """(F, Ho, Wo, C) NHWC (or (F, C, Ho, Wo) with nchw) -> G (F, RA, RB, cpad) rows of sub-pixel classes (py, px, c), zero
    outside the grid / in the padding, and the bias gradient sum(g) per channel — one fused pass."""
    g = g.contiguous()
    if nchw:
        F_, c, ho, wo = g.shape
    else:
        F_, ho, wo, c = g.shape
    G = torch.empty(F_, ra, rb, cpad, device=g.device, dtype=torch.float32)
    db = torch.empty(c, device=g.device, dtype=torch.float32)
    if F_ == 0:
        return G, db.zero_()
    rc = _lib.lib().repo_b200_grad_unshuffle(_p(g), int(nchw), _p(G), _p(db), F_, ra, rb, ho, wo, c, cpad, _stream())
    _lib.check(rc, "repo_b200_grad_unshuffle")
    return G, db


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
        # layer 1: 1x1 input -> k x k output is a plain GEMM; features ordered (kh, kw, co) = NHWC (F,k,k,128)
        k1, co1 = ws[0].shape[2], ws[0].shape[1]
        w1 = ws[0].detach().permute(2, 3, 1, 0).reshape(k1 * k1 * co1, -1)
        a1 = hl_empty((F_, k1, k1, co1), h.device)
        conv_gemm(h, w1, bs[0].detach().repeat(k1 * k1), a1, F_, k1 * k1 * co1,
                  ConvMap(RA=1, RB=1, in_nchw=0, C=h.shape[1], H=1, W=1, TH=1, TW=1, sy=1, sx=1, dy=1, dx=1, Ho=1, Wo=1, relu=1))
        acts, maps = [xin, h, a1], []
        hin = k1
        for li in (1, 2, 3):
            cin, cout, k = ws[li].shape[0], ws[li].shape[1], ws[li].shape[2]
            last = li == 3
            cm = _deconv_map(cin, hin, hin, k, out_nchw=last, relu=not last)
            out = torch.empty((F_, cout, cm.Ho, cm.Wo), device=h.device, dtype=torch.float32) if last \
                else hl_empty((F_, cm.Ho, cm.Wo, cout), h.device)
            conv_gemm(acts[-1], _deconv_wmat(ws[li].detach()), bs[li].detach().repeat(4), out, F_, 4 * cout, cm)
            acts.append(out)
            maps.append(cm)
            hin = cm.Ho
        ctx.maps = maps
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
        gp = g.contiguous().float()           # last layer has no activation; (F, C, 64, 64) NCHW, later layers NHWC
        for li in (3, 2, 1):
            x, cm = layer_in[li - 1], ctx.maps[li - 1]
            cin, cout, k = ws[li].shape[0], ws[li].shape[1], ws[li].shape[2]
            T = cm.TH
            cpad = _cpad(cout)
            G, db = _unshuffle(gp, cm.RA, cm.RB, cpad, nchw=(li == 3))  # (F, RA, RB, cpad) + bias gradient
            if need[2 + 2 * li + 1]:
                grads[2 * li + 3] = db
            sc = grad_scales(G)
            if need[2 + 2 * li]:
                dwm = conv_wgrad(x, G.reshape(-1, cpad), F_, 4 * cout, cm, sc)
                grads[2 * li + 2] = _deconv_wgrad(dwm, cin, cout, k)
            # data gradient: stride-1 conv over G with taps +t, masked by the ReLU that produced x
            wm = _deconv_wmat(ws[li]).reshape(4 * cout, T, T, cin)
            if cpad > 4 * cout:
                wm = F.pad(wm, (0, 0, 0, 0, 0, 0, 0, cpad - 4 * cout))
            wd = wm.permute(3, 1, 2, 0).reshape(cin, T * T * cpad)
            dmap = ConvMap(RA=cm.H, RB=cm.W, in_nchw=0, C=cpad, H=cm.RA, W=cm.RB, TH=T, TW=T, sy=1, sx=1, dy=1, dx=1,
                           Ho=cm.H, Wo=cm.W)
            d_in = torch.empty(x.shape[1:], device=x.device, dtype=torch.float32)
            conv_gemm(G, wd, None, d_in, F_, cin, dmap, relu_mask=x, scales=_as_input_side(sc))
            gp = d_in
        # layer 1 (plain GEMM) and fc1
        k1, co1 = ws[0].shape[2], ws[0].shape[1]
        g1 = gp.reshape(F_, k1 * k1 * co1)
        w1 = ws[0].permute(2, 3, 1, 0).reshape(k1 * k1 * co1, -1)
        # these three layers are plain GEMMs: same tcgen05 kernels (contraction over frames / over features), run dense
        big = F_ >= 256
        if need[2 + 2]:
            dw1 = wgrad_gemm(g1, h) if big else g1.t() @ h
            grads[2] = dw1.reshape(k1, k1, co1, -1).permute(3, 2, 0, 1).contiguous()
        if need[2 + 3]:
            grads[3] = g1.reshape(F_, k1 * k1, co1).sum((0, 1))
        if big:
            d_h = torch.empty(F_, w1.shape[1], device=g1.device, dtype=torch.float32)
            dense_layer(g1, w1.t().contiguous(), None, d_h, scales=_as_input_side(grad_scales(g1)))
        else:
            d_h = g1 @ w1
        if need[2]:
            grads[0] = wgrad_gemm(d_h, xin) if big else d_h.t() @ xin
        if need[3]:
            grads[1] = _ops.colsum(d_h)
        gb = gs = None
        if need[0] or need[1]:
            nin = fc_w.shape[1]
            if big:
                npad = (nin + 15) // 16 * 16
                d_x = torch.empty(F_, npad, device=g1.device, dtype=torch.float32)
                dense_layer(d_h, F.pad(fc_w.detach().t(), (0, 0, 0, npad - nin)).contiguous(), None, d_x,
                            scales=_as_input_side(grad_scales(d_h)))
            else:
                d_x = d_h @ fc_w
            gb, gs = d_x[:, :ctx.belief_size].contiguous(), d_x[:, ctx.belief_size:nin].contiguous()
        return (gb if need[0] else None, gs if need[1] else None, *grads)


class _TiaMixFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, t_out, d_out, weight, bias):
        t_out, d_out = _need_cuda(t_out.detach(), "t_out"), _need_cuda(d_out.detach(), "d_out")
        F_, _, H, W = t_out.shape
        recon = torch.empty(F_, 3, H, W, device=t_out.device, dtype=torch.float32)
        mask = torch.empty(F_, 1, H, W, device=t_out.device, dtype=torch.float32)
        w = weight.detach().reshape(-1).contiguous()
        rc = _lib.lib().repo_b200_tia_mix_fwd(_p(t_out), _p(d_out), _p(w), _p(bias.detach().contiguous()), _p(recon), _p(mask),
                                              F_, H * W, _stream())
        _lib.check(rc, "repo_b200_tia_mix_fwd")
        ctx.save_for_backward(t_out, d_out, w, mask)
        ctx.mark_non_differentiable(mask)
        return recon, mask

    @staticmethod
    def backward(ctx, g, _g_mask):
        t_out, d_out, w, mask = ctx.saved_tensors
        F_, _, H, W = t_out.shape
        g_t, g_d = torch.empty_like(t_out), torch.empty_like(d_out)
        g_wb = torch.empty(7, device=t_out.device, dtype=torch.float32)
        rc = _lib.lib().repo_b200_tia_mix_bwd(_p(t_out), _p(d_out), _p(w), _p(mask), _p(g.contiguous()), _p(g_t), _p(g_d), _p(g_wb),
                                              F_, H * W, _stream())
        _lib.check(rc, "repo_b200_tia_mix_bwd")
        return g_t, g_d, g_wb[:6].reshape(1, 6, 1, 1), g_wb[6:7]


def tia_mix(t_out, d_out, mask_head):
    """tia.py:124-127 in one kernel: `mask_head` is the reference's nn.Sequential(nn.Conv2d(6, 1, 1), nn.Sigmoid())
    (tia.py:72) used as the parameter holder; t_out / d_out are the un-chunked (F,6,H,W) decoder outputs.
    Returns (recon, mask)."""
    conv = mask_head[0]
    return _TiaMixFn.apply(t_out, d_out, conv.weight, conv.bias)


class VisualObservationModel(nn.Module):
    """decoder.py:28-48 — Linear(belief+state -> 1024, no activation) -> ConvTranspose2d(1024->128 k5, 128->64 k5,
    64->32 k6, 32->3 k6; stride 2, ReLU on the first three): 1 -> 5 -> 13 -> 30 -> 64 pixels."""

    def __init__(self, belief_size, state_size, embedding_size, activation_function="relu"):
        super().__init__()
        if activation_function != "relu":
            raise RuntimeError("the fused conv epilogue implements ReLU (cnn_activation_function default, train_repo.py:32)")
        if embedding_size % 16:
            raise RuntimeError("VisualObservationModel: embedding_size must be a multiple of 16")
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


class TIAObservationModel(VisualObservationModel):
    """decoder.py:154-175 — the same stack with 6 output channels, returned as (recon, mask) channel halves.
    `forward_full` returns the un-chunked (F,6,64,64) tensor for `tia_mix`."""

    def __init__(self, belief_size, state_size, embedding_size, activation_function="relu"):
        super().__init__(belief_size, state_size, embedding_size, activation_function)
        self.conv4 = nn.ConvTranspose2d(32, 6, 6, stride=2)

    def forward_full(self, belief, state):
        return super().forward(belief, state)

    def forward(self, belief, state):
        out = super().forward(belief, state)
        recon, mask = out.chunk(2, 1)
        return recon, mask

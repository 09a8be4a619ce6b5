"""CPU: the host-side algebra of the implicit-GEMM convolutions (repo_b200/conv.py) — ConvMap gathers, the sub-pixel
weight matrices of ConvTranspose2d, the adjoint (data-gradient) maps and the weight-gradient reshapes — checked against
torch's own conv ops with a dense emulation of what the kernels compute.  No GPU, no extension calls."""
import pytest
import torch
import torch.nn.functional as F

from repo_b200 import conv as cv


def emulate(x_nhwc, wmat, cm, n_total, bias=None):
    """out = gather(x) @ wmat^T (+ bias) with the ConvMap's row grid, taps and (optional) sub-pixel store."""
    Fr = x_nhwc.shape[0]
    cols = torch.zeros(Fr, cm.RA, cm.RB, cm.TH, cm.TW, cm.C, dtype=x_nhwc.dtype)
    for a in range(cm.RA):
        for b in range(cm.RB):
            for ty in range(cm.TH):
                for tx in range(cm.TW):
                    iy, ix = a * cm.sy + ty * cm.dy + cm.y0, b * cm.sx + tx * cm.dx + cm.x0
                    if 0 <= iy < cm.H and 0 <= ix < cm.W:
                        cols[:, a, b, ty, tx] = x_nhwc[:, iy, ix]
    y = cols.reshape(Fr * cm.RA * cm.RB, -1) @ wmat.t()
    if bias is not None:
        y = y + bias
    y = y.reshape(Fr, cm.RA, cm.RB, n_total)
    if cm.shuffle:
        co = n_total // 4
        out = torch.zeros(Fr, 2 * cm.RA, 2 * cm.RB, co, dtype=y.dtype)
        y = y.reshape(Fr, cm.RA, cm.RB, 2, 2, co)
        for py in range(2):
            for px in range(2):
                out[:, py::2, px::2] = y[:, :, :, py, px]
        return out[:, :cm.Ho, :cm.Wo], cols
    return y, cols


def unshuffle_ref(g_nhwc, ra, rb, cpad):
    Fr, ho, wo, c = g_nhwc.shape
    gp = F.pad(g_nhwc, (0, 0, 0, 2 * rb - wo, 0, 2 * ra - ho))
    gp = gp.reshape(Fr, ra, 2, rb, 2, c).permute(0, 1, 3, 2, 4, 5).reshape(Fr, ra, rb, 4 * c)
    return F.pad(gp, (0, cpad - 4 * c))


@pytest.mark.parametrize("cin,cout,k,hin", [(8, 4, 5, 5), (8, 3, 6, 7), (4, 6, 6, 3), (8, 2, 5, 1)])
def test_transposed_conv_as_subpixel_gemm(cin, cout, k, hin):
    torch.manual_seed(0)
    w, b = torch.randn(cin, cout, k, k, dtype=torch.float64), torch.randn(cout, dtype=torch.float64)
    x = torch.randn(2, cin, hin, hin, dtype=torch.float64, requires_grad=True)
    wr = w.clone().requires_grad_(True)
    ref = F.conv_transpose2d(x, wr, b, stride=2)
    cm = cv._deconv_map(cin, hin, hin, k, False, False)
    got, cols = emulate(x.detach().permute(0, 2, 3, 1), cv._deconv_wmat(w), cm, 4 * cout, b.repeat(4))
    torch.testing.assert_close(got.permute(0, 3, 1, 2), ref.detach())
    # backward: weight gradient = G^T @ gathered rows, data gradient = stride-1 conv over the un-shuffled gradient
    g = torch.randn_like(ref)
    (ref * g).sum().backward()
    cpad = cv._cpad(cout)
    G = unshuffle_ref(g.permute(0, 2, 3, 1).contiguous(), cm.RA, cm.RB, cpad)
    T = cm.TH
    dwm = G.reshape(-1, cpad)[:, :4 * cout].t() @ cols.reshape(-1, T * T * cin)
    torch.testing.assert_close(cv._deconv_wgrad(dwm, cin, cout, k), wr.grad)
    wm = F.pad(cv._deconv_wmat(w).reshape(4 * cout, T, T, cin), (0, 0, 0, 0, 0, 0, 0, cpad - 4 * cout))
    wd = wm.permute(3, 1, 2, 0).reshape(cin, T * T * cpad)
    dmap = cv.ConvMap(RA=cm.H, RB=cm.W, in_nchw=0, C=cpad, H=cm.RA, W=cm.RB, TH=T, TW=T, sy=1, sx=1, dy=1, dx=1, Ho=cm.H, Wo=cm.W)
    dx, _ = emulate(G, wd, dmap, cin)
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), x.grad)


@pytest.mark.parametrize("cin,cout,H", [(8, 16, 14), (8, 8, 31), (3, 4, 64)])
def test_strided_conv_and_its_subpixel_data_gradient(cin, cout, H):
    torch.manual_seed(1)
    w = torch.randn(cout, cin, 4, 4, dtype=torch.float64)
    x = torch.randn(2, cin, H, H, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w, None, stride=2)
    Ho = y.shape[2]
    cm = cv.ConvMap(RA=Ho, RB=Ho, in_nchw=0, C=cin, H=H, W=H, TH=4, TW=4, sy=2, sx=2, dy=1, dx=1, Ho=Ho, Wo=Ho)
    got, _ = emulate(x.detach().permute(0, 2, 3, 1), cv._conv_wmat(w), cm, cout)
    torch.testing.assert_close(got.permute(0, 3, 1, 2), y.detach())
    g = torch.randn_like(y)
    (y * g).sum().backward()
    dmap = cv.ConvMap(RA=(H + 1) // 2, RB=(H + 1) // 2, in_nchw=0, C=cout, H=Ho, W=Ho, TH=2, TW=2, sy=1, sx=1, dy=-1, dx=-1,
                      Ho=H, Wo=H, osy=2, osx=2, shuffle=1)
    wd = w.reshape(cout, cin, 2, 2, 2, 2).permute(3, 5, 1, 2, 4, 0).reshape(4 * cin, 4 * cout)
    dx, _ = emulate(g.permute(0, 2, 3, 1).contiguous(), wd, dmap, 4 * cin)
    torch.testing.assert_close(dx.permute(0, 3, 1, 2), x.grad)


def test_encoder_maps_and_padding_helpers():
    maps = cv._enc_maps((64, 64))
    assert [(m.Ho, m.Wo) for m in maps] == [(31, 31), (14, 14), (6, 6), (2, 2)]          # encoder.py:26-29 on 64x64 frames
    assert [m.in_nchw for m in maps] == [1, 0, 0, 0] and maps[3].out_nchw == 1
    assert [cv._cpad(c) for c in (3, 6, 32, 64)] == [16, 32, 128, 256]
    sizes, h = [], 1
    for cin, cout, k in [(1024, 128, 5), (128, 64, 5), (64, 32, 6), (32, 3, 6)]:
        h = (h - 1) * 2 + k
        sizes.append(h)
    assert sizes == [5, 13, 30, 64]                                                     # decoder.py:35-39
    m = cv._deconv_map(32, 30, 30, 6, True, False)
    assert (m.RA, m.RB, m.TH, m.TW, m.shuffle, m.out_nchw, m.relu) == (32, 32, 3, 3, 1, 1, 0)
    arr = m.carray()
    assert len(arr) == len(cv._MAP_FIELDS) == 28

"""Bring-up of the tcgen05 weight-gradient kernel: compare against im2col + fp64 matmul for a few maps,
with and without the LBO/SBO swap flag."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv, _lib

dev = torch.device("cuda:0")
L = _lib.lib()


def check(name, x, cm, n_total, Fr, pad_to=None):
    rows = Fr * cm.RA * cm.RB
    ld = pad_to or n_total
    g = torch.zeros(rows, ld, device=dev)
    g[:, :n_total] = torch.randn(rows, n_total, device=dev) * 1e-5
    col = cv.im2col(x, Fr, cm)
    want = (g[:, :n_total].double().t() @ col.double())
    for swap in (0,):
        L.repo_b200_debug_flags(swap)
        got = cv.conv_wgrad(x, g, Fr, n_total, cm)
        torch.cuda.synchronize()
        err = ((got.double() - want).abs().max() / want.abs().max()).item()
        print(f"{name:28s} swap={swap} rows={rows} K={cm.K} n={n_total} rel_err={err:.3e}", flush=True)
    L.repo_b200_debug_flags(0)


Fr = 5
maps = cv._enc_maps((64, 64))
check("enc2 (C=32,s2) n=64", torch.randn(Fr, 31, 31, 32, device=dev), maps[1], 64, Fr)
check("enc1 (NCHW C=3) n=32", torch.randn(Fr, 3, 64, 64, device=dev), maps[0], 32, Fr)
check("enc4 (K=2048) n=256", torch.randn(Fr, 6, 6, 128, device=dev), maps[3], 256, Fr)
cm = cv._deconv_map(64, 13, 13, 6, False, True)
check("dec3 (K=576) n=128", torch.randn(Fr, 13, 13, 64, device=dev), cm, 128, Fr)
cm = cv._deconv_map(32, 30, 30, 6, True, False)
check("dec4 (K=288) n=12 ld=16", torch.randn(Fr, 30, 30, 32, device=dev), cm, 12, Fr, pad_to=16)
cm = cv._deconv_map(128, 5, 5, 5, False, True)
check("dec2 (K=1152) n=256", torch.randn(Fr, 5, 5, 128, device=dev), cm, 256, Fr)

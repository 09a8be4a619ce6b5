import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv, _lib
dev = torch.device("cuda:0")
Fr = 2450
def to_hl(x):
    hi = x.half(); return torch.stack([hi, (x - hi.float()).half()]).contiguous()
cases = [("dec4 HL->nchw", cv._deconv_map(32, 30, 30, 6, True, False), (Fr, 30, 30, 32), 12, False),
         ("dec3 HL->HL", cv._deconv_map(64, 13, 13, 6, False, True), (Fr, 13, 13, 64), 128, True),
         ("enc2 HL->HL", cv._enc_maps((64, 64))[1], (Fr, 31, 31, 32), 64, True)]
for name, cm, xs, n, out_hl in cases:
    x = to_hl(torch.randn(xs, device=dev))
    w = torch.randn(n, cm.K, device=dev) * 0.05; b = torch.zeros(n, device=dev)
    cout = n // 4 if cm.shuffle else n
    out = cv.hl_empty((Fr, cm.Ho, cm.Wo, cout), dev) if out_hl else torch.empty((Fr, cout, cm.Ho, cm.Wo) if cm.out_nchw else (Fr, cm.Ho, cm.Wo, cout), device=dev)
    cv.conv_gemm(x, w, b, out, Fr, n, cm); torch.cuda.synchronize()
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    _lib.lib().repo_b200_debug_clock(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cv.conv_gemm(x, w, b, out, Fr, n, cm); e1.record(); torch.cuda.synchronize()
    _lib.lib().repo_b200_debug_clock(None)
    v = buf.tolist(); nch, nit = max(1, v[7]), max(1, v[8])
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms; CTA0 items {nit} chunks {nch}; per chunk: mma wait_full {v[1]/nch:.0f} issue {v[2]/nch:.0f} (wait_acc/item {v[0]/nit:.0f}); gather-group loop {v[3]/nch:.0f} x G per chunk it handles, wait_empty {v[4]/nch:.0f}; epilogue per item: wait {v[5]/nit:.0f} work {v[6]/nit:.0f}")

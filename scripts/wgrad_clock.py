import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv, _lib
dev = torch.device("cuda:0")
Fr = 2450
def to_hl(x):
    hi = x.half(); return torch.stack([hi, (x - hi.float()).half()]).contiguous()
for name, cm, xs, n, ld in [("dec4", cv._deconv_map(32, 30, 30, 6, True, False), (Fr, 30, 30, 32), 12, 16),
                            ("dec3", cv._deconv_map(64, 13, 13, 6, False, True), (Fr, 13, 13, 64), 128, 128),
                            ("enc1", cv._enc_maps((64, 64))[0], (Fr, 3, 64, 64), 32, 32),
                            ("enc2", cv._enc_maps((64, 64))[1], (Fr, 31, 31, 32), 64, 64)]:
    x = torch.randn(xs, device=dev) if name == "enc1" else to_hl(torch.randn(xs, device=dev))
    rows = Fr * cm.RA * cm.RB
    g = torch.randn(rows, ld, device=dev) * 1e-4
    sc = cv.grad_scales(g)
    cv.conv_wgrad(x, g, Fr, n, cm, sc); torch.cuda.synchronize()
    buf = torch.zeros(8, dtype=torch.int64, device=dev)
    _lib.lib().repo_b200_debug_clock(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cv.conv_wgrad(x, g, Fr, n, cm, sc); e1.record(); torch.cuda.synchronize()
    _lib.lib().repo_b200_debug_clock(None)
    b = buf.tolist()
    st = max(1, b[6])
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms, stages/CTA {st}; per stage cycles: producer issue {b[0]/st:.0f} wait_empty {b[1]/st:.0f} data+store {b[2]/st:.0f} arrive {b[3]/st:.0f} | mma wait_full {b[4]/st:.0f} issue {b[5]/st:.0f}")

"""Where the ~17 us floor of one conv / dense launch goes: global-timer stamps of CTA 0 (entry, prologue done, MMA warp done,
epilogue done, exit) for a one-tile dense layer, next to the kernel duration events see."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv, _lib
dev = torch.device("cuda:0")
for rows in (128, 2450):
    x = torch.randn(rows, 200, device=dev); w = torch.randn(200, 200, device=dev) * 0.05; b = torch.zeros(200, device=dev)
    out = torch.empty(rows, 200, device=dev)
    for _ in range(3):
        cv.dense_layer(x, w, b, out, act="elu")
    torch.cuda.synchronize()
    buf = torch.zeros(16, dtype=torch.int64, device=dev)
    _lib.lib().repo_b200_debug_clock(C.c_void_p(buf.data_ptr()))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); cv.dense_layer(x, w, b, out, act="elu"); e1.record(); torch.cuda.synchronize()
    _lib.lib().repo_b200_debug_clock(None)
    v = buf.tolist()
    t0 = v[10]
    print(f"rows {rows}: events around pack + kernel {e0.elapsed_time(e1)*1e3:.1f} us; CTA 0: prologue done +{(v[11]-t0)/1e3:.2f} us, MMA warp done +{(v[12]-t0)/1e3:.2f}, epilogue done +{(v[13]-t0)/1e3:.2f}, exit +{(v[14]-t0)/1e3:.2f}")

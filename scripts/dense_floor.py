import sys, os
sys.path.insert(0, "/root/repo")
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import conv as cv
dev = torch.device("cuda:0")
for rows in (128, 343, 2450, 4802, 34300):
    x = torch.randn(rows, 200, device=dev); w = torch.randn(200, 200, device=dev) * 0.05; b = torch.zeros(200, device=dev)
    out = torch.empty(rows, 200, device=dev)
    g = torch.randn(rows, 200, device=dev) * 1e-4
    for _ in range(3):
        cv.dense_layer(x, w, b, out, act="elu"); cv.wgrad_gemm(g, x)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        cv.dense_layer(x, w, b, out, act="elu"); cv.wgrad_gemm(g, x); torch.cuda.synchronize()
    print(rows, {e.key.split("(")[0][-32:]: round(e.self_device_time_total, 1) for e in prof.key_averages()})

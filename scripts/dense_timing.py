"""Time conv.dense_layer (one tcgen05 GEMM launch) on MLP-head shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv
dev = torch.device("cuda:0")
def timed(fn, n=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
N = int(sys.argv[1]) if len(sys.argv) > 1 else 34300
for (K, n, act, strided) in [(200, 200, "none", False), (200, 200, "elu", False), (200, 200, "elu", True), (232, 200, "elu", False),
                             (200, 16, "none", False), (200, 1, "none", False), (1024, 256, "none", False), (200, 128, "none", False), (200, 256, "none", False)]:
    x = torch.randn(N, K, device=dev)
    w = torch.randn(n, K, device=dev) * 0.05; b = torch.zeros(n, device=dev)
    if strided:
        big = torch.empty(N, 3 * n, device=dev); out = big[:, n:2 * n]
    else:
        out = torch.empty(N, n, device=dev)
    us = timed(lambda: cv.dense_layer(x, w, b, out, act=act))
    print(f"N={N} K={K} n={n} act={act} strided={strided}: {us:8.1f} us  {2e-6*N*K*n/us:8.1f} TFLOP/s")

# the same GEMM inside a CUDA graph (what Agent.graphed replays): kernel time without the host launch path
x = torch.randn(N, 200, device=dev); w = torch.randn(200, 200, device=dev) * 0.05; b = torch.zeros(200, device=dev); out = torch.empty(N, 200, device=dev)
s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3): cv.dense_layer(x, w, b, out, act="elu")
torch.cuda.current_stream().wait_stream(s)
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for _ in range(10): cv.dense_layer(x, w, b, out, act="elu")
us = timed(g.replay) / 10
print(f"graphed N={N} K=200 n=200 elu: {us:8.1f} us per layer  {2e-6*N*200*200/us:8.1f} TFLOP/s")

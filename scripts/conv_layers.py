"""Per-layer timing of the conv kernel (forward launches and data-gradient launches) with the FLOP / byte
figures that bound each, at the trainer's default frame count."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import conv as cv


ONCE = os.environ.get("CONV_ONCE") == "1"   # one launch per layer, for ncu captures


def timed(fn, n=10, warm=3):
    if ONCE:
        n, warm = 1, 0
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def main():
    dev = torch.device("cuda:0")
    Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 2450
    rows = []
    # encoder forward layers
    chans, h = [3, 32, 64, 128, 256], 64
    for i, cm in enumerate(cv._enc_maps((64, 64))):
        cin, cout = chans[i], chans[i + 1]
        x = torch.randn((Fr, cin, h, h) if i == 0 else (Fr, h, h, cin), device=dev)
        w = torch.randn(cout, cm.K, device=dev) * 0.05
        b = torch.zeros(cout, device=dev)
        out = torch.empty((Fr, cout, cm.Ho, cm.Wo) if cm.out_nchw else (Fr, cm.Ho, cm.Wo, cout), device=dev)
        ms = timed(lambda: cv.conv_gemm(x, w, b, out, Fr, cout, cm))
        n_rows = Fr * cm.RA * cm.RB
        rows.append(dict(layer=f"enc{i+1}", rows=n_rows, K=cm.K, N=cout, ms=ms, gflop=2e-9 * n_rows * cm.K * cout,
                         mb=(x.numel() + out.numel()) * 4e-6))
        h = cm.Ho
    # decoder forward layers
    hin = 5
    x = torch.randn(Fr, 1024, device=dev); w = torch.randn(3200, 1024, device=dev) * 0.03; b = torch.zeros(3200, device=dev)
    out = torch.empty(Fr, 5, 5, 128, device=dev)
    cm = cv.ConvMap(RA=1, RB=1, in_nchw=0, C=1024, H=1, W=1, TH=1, TW=1, sy=1, sx=1, dy=1, dx=1, Ho=1, Wo=1, relu=1)
    ms = timed(lambda: cv.conv_gemm(x, w, b, out, Fr, 3200, cm))
    rows.append(dict(layer="dec1", rows=Fr, K=1024, N=3200, ms=ms, gflop=2e-9 * Fr * 1024 * 3200, mb=(x.numel() + out.numel()) * 4e-6))
    for li in (1, 2, 3):
        cin, cout, k = cv._DEC[li]
        last = li == 3
        cm = cv._deconv_map(cin, hin, hin, k, out_nchw=last, relu=not last)
        x = torch.randn(Fr, hin, hin, cin, device=dev)
        w = torch.randn(4 * cout, cm.K, device=dev) * 0.05; b = torch.zeros(4 * cout, device=dev)
        out = torch.empty((Fr, cout, cm.Ho, cm.Wo) if last else (Fr, cm.Ho, cm.Wo, cout), device=dev)
        ms = timed(lambda: cv.conv_gemm(x, w, b, out, Fr, 4 * cout, cm))
        n_rows = Fr * cm.RA * cm.RB
        rows.append(dict(layer=f"dec{li+1}", rows=n_rows, K=cm.K, N=4 * cout, ms=ms, gflop=2e-9 * n_rows * cm.K * 4 * cout,
                         mb=(x.numel() + out.numel()) * 4e-6))
        hin = cm.Ho
    tot = 0
    for r in rows:
        r["tflops"] = r["gflop"] / r["ms"]
        r["gbps"] = r["mb"] / r["ms"]
        tot += r["ms"]
        print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}))
    print("total_ms", round(tot, 3))


if __name__ == "__main__":
    main()

"""Time the conv encoder/decoder (forward, forward+backward) at the trainer's default shape and print
cuDNN (torch.nn) times for the same layers beside them as a sanity line (not a product path)."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import synth
from repo_b200.conv import VisualEncoder, VisualObservationModel


def timed(fn, n=5, warm=2):
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
    F = int(sys.argv[1]) if len(sys.argv) > 1 else 2450
    enc = VisualEncoder(1024).to(dev); enc.load_state_dict(synth.make_conv_params("encoder", 1))
    dec = VisualObservationModel(200, 30, 1024).to(dev); dec.load_state_dict(synth.make_conv_params("decoder", 2))
    frames = synth.make_frames(3, F).to(dev)
    b = torch.randn(F, 200, device=dev); s = torch.randn(F, 30, device=dev)
    out = {"frames": F}
    with torch.no_grad():
        out["encoder_fwd_ms"] = timed(lambda: enc(frames))
        out["decoder_fwd_ms"] = timed(lambda: dec(b, s))

    def enc_fb():
        enc.zero_grad(set_to_none=True)
        enc(frames).square().mean().backward()

    def dec_fb():
        dec.zero_grad(set_to_none=True)
        dec(b, s).square().mean().backward()
    out["encoder_fwd_bwd_ms"] = timed(enc_fb, 3, 1)
    out["decoder_fwd_bwd_ms"] = timed(dec_fb, 3, 1)

    # cuDNN sanity line: the same layers through torch.nn (fp32, TF32 off)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    import torch.nn as nn
    tenc = nn.Sequential(nn.Conv2d(3, 32, 4, 2), nn.ReLU(), nn.Conv2d(32, 64, 4, 2), nn.ReLU(), nn.Conv2d(64, 128, 4, 2), nn.ReLU(),
                         nn.Conv2d(128, 256, 4, 2), nn.ReLU()).to(dev)
    tdec = nn.Sequential(nn.ConvTranspose2d(1024, 128, 5, 2), nn.ReLU(), nn.ConvTranspose2d(128, 64, 5, 2), nn.ReLU(),
                         nn.ConvTranspose2d(64, 32, 6, 2), nn.ReLU(), nn.ConvTranspose2d(32, 3, 6, 2)).to(dev)
    h = torch.randn(F, 1024, 1, 1, device=dev)
    with torch.no_grad():
        out["cudnn_encoder_fwd_ms"] = timed(lambda: tenc(frames))
        out["cudnn_decoder_fwd_ms"] = timed(lambda: tdec(h))

    def tenc_fb():
        tenc.zero_grad(set_to_none=True); tenc(frames).square().mean().backward()

    def tdec_fb():
        tdec.zero_grad(set_to_none=True); tdec(h).square().mean().backward()
    out["cudnn_encoder_fwd_bwd_ms"] = timed(tenc_fb, 3, 1)
    out["cudnn_decoder_fwd_bwd_ms"] = timed(tdec_fb, 3, 1)
    print(json.dumps({k: (round(v, 3) if isinstance(v, float) else v) for k, v in out.items()}))


if __name__ == "__main__":
    main()

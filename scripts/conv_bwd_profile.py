"""Kernel-level breakdown of encoder / decoder forward+backward (torch.profiler, CUDA time per kernel name)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import synth
from repo_b200.conv import VisualEncoder, VisualObservationModel

dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2450
enc = VisualEncoder(1024).to(dev); enc.load_state_dict(synth.make_conv_params("encoder", 1))
dec = VisualObservationModel(200, 30, 1024).to(dev); dec.load_state_dict(synth.make_conv_params("decoder", 2))
frames = synth.make_frames(3, F).to(dev)
b = torch.randn(F, 200, device=dev); s = torch.randn(F, 30, device=dev)


def enc_fb():
    enc.zero_grad(set_to_none=True); enc(frames).square().mean().backward()


def dec_fb():
    dec.zero_grad(set_to_none=True); dec(b, s).square().mean().backward()


for name, fn in (("encoder", enc_fb), ("decoder", dec_fb)):
    fn(); fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        fn(); torch.cuda.synchronize()
    print("=====", name)
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=18, max_name_column_width=70))

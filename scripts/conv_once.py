"""One encoder + decoder forward/backward at the default frame count (for ncu captures of the conv kernels)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import synth
from repo_b200.conv import VisualEncoder, VisualObservationModel

dev = torch.device("cuda:0")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 2450
enc = VisualEncoder(1024).to(dev); enc.load_state_dict(synth.make_conv_params("encoder", 1))
dec = VisualObservationModel(200, 30, 1024).to(dev); dec.load_state_dict(synth.make_conv_params("decoder", 2))
frames = synth.make_frames(3, F).to(dev)
b = torch.randn(F, 200, device=dev); s = torch.randn(F, 30, device=dev)
enc(frames).square().mean().backward()
dec(b, s).square().mean().backward()
torch.cuda.synchronize()
print("ok")

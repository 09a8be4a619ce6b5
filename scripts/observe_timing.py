"""observe at a batch that fills the GPU (18,944 sequences x 49 steps through the 128-row kernel): ms per call."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import rssm_oracle as O
from repo_b200 import ops
dev = torch.device("cuda:0")
P = {k: v.to(dev) for k, v in O.make_transition_params(0).items()}
OB, OT, D, S, A = 18944, 50, 200, 30, 6
gen = torch.Generator(device=dev).manual_seed(11)
rn = lambda *sh: torch.randn(*sh, device=dev, generator=gen)
big = [torch.zeros(OB, D, device=dev), torch.zeros(OB, S, device=dev), rn(OT - 1, OB, A).clamp_(-1, 1), rn(OT - 1, OB, 1024),
       torch.ones(OT - 1, OB, 1, device=dev), rn(OT - 1, OB, S), rn(OT - 1, OB, S)]
for _ in range(3):
    ops.observe_fwd(P, *big)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    ops.observe_fwd(P, *big)
e1.record()
torch.cuda.synchronize()
print("observe 18944x49: %.3f ms" % (e0.elapsed_time(e1) / 5))

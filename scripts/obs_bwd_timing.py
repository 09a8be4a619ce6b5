"""observe forward + backward at the default training shape (B = 50, T = 50): ms per call, kernel breakdown."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import synth as O
from repo_b200.rssm import TransitionModel
dev = torch.device("cuda:0")
tm = TransitionModel(200, 30, 6, 200, 1024, "elu").to(dev)
tm.load_state_dict(O.make_transition_params(1))
x = O.make_observe_inputs(4, 50, 50)
g = lambda k: x[k].to(dev)
emb = g("embeds").requires_grad_(True)
def step():
    outs = tm.observe(g("prev_belief"), g("prev_state"), g("actions"), emb, g("nonterms"), eps_prior=g("eps_prior"), eps_post=g("eps_post"))
    (outs[0].sum() + outs[4].sum() + outs[5].sum()).backward()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(); torch.cuda.synchronize()
for e in sorted(prof.key_averages(), key=lambda e: -e.self_device_time_total)[:6]:
    print(f"{e.self_device_time_total/1e3:8.3f} ms x{e.count:<3d} {e.key[:90]}")

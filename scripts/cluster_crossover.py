"""Where the cluster observe kernel stops winning: 49-step observe (no stash) at growing batches, cluster kernel
(row_tile=1) vs the 128-row kernel (row_tile=128)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from repo_b200 import ops
from oracle import rssm_oracle as O

dev = torch.device("cuda:0")
params = {k: v.to(dev) for k, v in O.make_transition_params(1).items()}
for B in (16, 50, 128, 144, 256, 384, 512, 768, 1024):
    x = O.make_observe_inputs(2, 50, B, p_done=0.05)
    args = [x[k].to(dev) for k in ("prev_belief", "prev_state", "actions", "embeds", "nonterms", "eps_prior", "eps_post")]
    res = []
    for rt in (1, 128):
        ws = None
        for _ in range(3):
            _, _, ws = ops.observe_fwd(params, *args, row_tile=rt, workspace=ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.observe_fwd(params, *args, row_tile=rt, workspace=ws)
        e1.record(); torch.cuda.synchronize()
        res.append(e0.elapsed_time(e1) / 10)
    print(f"B={B:5d}: cluster {res[0]:.3f} ms   128-row kernel {res[1]:.3f} ms", flush=True)

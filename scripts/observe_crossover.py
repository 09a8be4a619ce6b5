import os, sys, torch
sys.path.insert(0, "/root/repo")
from repo_b200 import ops, synth as O
dev = torch.device("cuda:0")
P = {k: v.to(dev) for k, v in O.make_transition_params(0).items()}
for B in (16, 50, 128, 512, 2048, 4096):
    xo = O.make_observe_inputs(1, 50, B)
    a = [P] + [xo[k].to(dev) for k in ("prev_belief", "prev_state", "actions", "embeds", "nonterms", "eps_prior", "eps_post")]
    line = f"observe B={B:5d} x 49:"
    for rt in (0, 16, 32, 128):
        for _ in range(3): ops.observe_fwd(*a, row_tile=rt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): ops.observe_fwd(*a, row_tile=rt)
        e1.record(); torch.cuda.synchronize()
        line += f"  rt{rt}: {e0.elapsed_time(e1) / 5:7.3f} ms"
    print(line)

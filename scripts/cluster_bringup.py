"""Bring-up of the cluster observe kernel (row_tile=1): outputs and stash against the vm kernel (row_tile=16) and the
oracle, then timing at the reference's operating point (50 sequences x 49 steps)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from repo_b200 import ops
from oracle import rssm_oracle as O

dev = torch.device("cuda:0")
cu = lambda p: {k: v.to(dev) for k, v in p.items()}


def run(params, x, rt, stash):
    g = lambda k: None if x[k] is None else x[k].to(dev)
    T1, B = x["actions"].shape[:2]
    d = ops.dims_of(cu(params))
    st = torch.zeros(T1, B, 5 * d.belief + 2 * d.hidden, device=dev) if stash else None
    outs, kl, _ = ops.observe_fwd(cu(params), g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"),
                                  g("eps_prior"), g("eps_post"), row_tile=rt, stash=st)
    torch.cuda.synchronize()
    return outs, kl, st


for (T, B, pd, with_obs) in [(3, 5, 0.0, True), (8, 10, 0.3, True), (49, 50, 0.1, True), (7, 33, 0.2, False), (5, 130, 0.1, True)]:
    params = O.make_transition_params(100 + T)
    x = O.make_observe_inputs(200 + B, T, B, p_done=pd)
    if not with_obs:
        x["embeds"] = None; x["eps_post"] = None
    want = O.observe(params, x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"], x["eps_prior"], x["eps_post"])
    o1, kl1, s1 = run(params, x, 1, True)
    o0, kl0, s0 = run(params, x, 16, True)
    worst = 0.0
    for a, b, w in zip(o1, o0, want):
        w = w.to(dev)
        err = ((a - w).abs() / (1e-5 + 1e-3 * w.abs())).max().item()
        worst = max(worst, err)
    ds = (s1 - s0).abs().max().item()
    dk = (kl1 - kl0).abs().max().item() if kl1 is not None else 0.0
    print(f"T={T} B={B} obs={with_obs}: worst err/tol vs oracle {worst:.4f}; stash maxdiff vs vm {ds:.3e}; kl maxdiff {dk:.3e}", flush=True)

params = cu(O.make_transition_params(1))
x = O.make_observe_inputs(2, 49, 50, p_done=0.05)
g = lambda k: x[k].to(dev)
args = (params, g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"), g("eps_prior"), g("eps_post"))
d = ops.dims_of(params)
for rt, stash in [(1, False), (1, True), (16, True), (128, False)]:
    st = torch.zeros(49, 50, 5 * d.belief + 2 * d.hidden, device=dev) if stash else None
    ws = None
    for _ in range(3):
        _, _, ws = ops.observe_fwd(*args, row_tile=rt, stash=st, workspace=ws)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.observe_fwd(*args, row_tile=rt, stash=st, workspace=ws)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"observe 50x49 row_tile={rt} stash={stash}: {ms:.3f} ms = {ms*1e3/49:.1f} us per time step", flush=True)

"""imagine forward (with stash) + backward through TransitionModel.imagine at small row counts (a data-parallel shard)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from repo_b200 import synth as O
from repo_b200.models import ActorModel
from repo_b200.rssm import TransitionModel

dev = torch.device("cuda:0")
tm = TransitionModel(200, 30, 6, 200, 1024, "elu").to(dev)
tm.load_state_dict(O.make_transition_params(1))
actor = ActorModel(200, 30, 200, 6, "elu").to(dev)
actor.load_state_dict(O.make_mlp_params(2, 230, 200, 12, 4))
for N in (100, 343, 600, 1225, 2450):
    x = O.make_imagine_inputs(5, N, 15)
    b, s = x["belief"].to(dev), x["state"].to(dev)

    def run():
        for p_ in list(tm.parameters()) + list(actor.parameters()):
            p_.grad = None
        outs = tm.imagine(b, s, actor, 15, eps_action=x["eps_action"].to(dev), eps_prior=x["eps_prior"].to(dev))
        (outs[0].sum() * 1e-3 + outs[1].sum() * 1e-3).backward()

    for _ in range(3):
        run()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        run(); torch.cuda.synchronize()
    ks = {e.key.split("(")[0][-40:]: e.self_device_time_total for e in prof.key_averages() if "imagine_bwd" in e.key or "rssm_vm" in e.key or "rssm_rows" in e.key}
    print(N, {k: round(v / 1e3, 3) for k, v in ks.items()}, flush=True)

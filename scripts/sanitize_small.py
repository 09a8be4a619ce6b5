"""Small forward + backward workload for compute-sanitizer (memcheck / racecheck)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repo_b200 import synth as O  # noqa: E402
from repo_b200.models import ActorModel, RewardModel  # noqa: E402
from repo_b200.rssm import TransitionModel  # noqa: E402

dev = torch.device("cuda:0")
tm = TransitionModel(200, 30, 6, 200, 1024, "elu").to(dev)
tm.load_state_dict(O.make_transition_params(1))
actor = ActorModel(200, 30, 200, 6, "elu").to(dev)
actor.load_state_dict(O.make_mlp_params(2, 230, 200, 12, 4))
reward = RewardModel(200, 30, 200, "elu").to(dev)
reward.load_state_dict(O.make_mlp_params(3, 230, 200, 1, 3))
x = O.make_observe_inputs(4, 4, 5)
g = lambda k: x[k].to(dev)
outs = tm.observe(g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"), eps_prior=g("eps_prior"), eps_post=g("eps_post"))
(outs[0].sum() + outs[4].sum() + tm.last_kl.sum() * 0).backward()
xi = O.make_imagine_inputs(5, 37, 4)
traj = tm.imagine(xi["belief"].to(dev), xi["state"].to(dev), actor, 4, eps_action=xi["eps_action"].to(dev), eps_prior=xi["eps_prior"].to(dev))
r = reward(traj[0].flatten(0, 1), traj[1].flatten(0, 1))
ent = actor.get_action_dist(traj[0].flatten(0, 1), traj[1].flatten(0, 1)).entropy().mean()
(r.sum() + ent).backward()
with torch.no_grad():
    big = O.make_imagine_inputs(6, 300, 4)
    from repo_b200 import ops
    named = lambda m: {k: v.detach() for k, v in m.named_parameters()}
    out = ops.imagine_fwd(named(tm), named(actor), named(reward), named(reward), big["belief"].to(dev), big["state"].to(dev),
                          big["eps_action"].to(dev), big["eps_prior"].to(dev), 4, row_tile=128)
    xo = O.make_observe_inputs(7, 4, 200)      # 128-row kernel, posterior + KL path, partial second tile
    go = lambda k: xo[k].to(dev)
    oo, kl, _ = ops.observe_fwd(named(tm), go("prev_belief"), go("prev_state"), go("actions"), go("embeds"), go("nonterms"),
                                go("eps_prior"), go("eps_post"), row_tile=128)
torch.cuda.synchronize()
print("sanitize workload done", float(out["returns"].sum()), float(kl.sum()))

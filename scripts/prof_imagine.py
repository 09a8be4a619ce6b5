"""One imagine (or observe) launch for ncu: python scripts/prof_imagine.py imagine N row_tile | observe B T row_tile"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rssm_oracle as O  # noqa: E402
from repo_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
cu = lambda p: {k: v.to(dev) for k, v in p.items()}
kind = sys.argv[1]
D = O.DEFAULT_DIMS
params = cu(O.make_transition_params(0))
if kind == "imagine":
    N, rt = int(sys.argv[2]), int(sys.argv[3])
    actor = cu(O.make_mlp_params(1, 230, 200, 12, 4))
    reward = cu(O.make_mlp_params(2, 230, 200, 1, 3))
    value = cu(O.make_mlp_params(3, 230, 200, 1, 3))
    x = O.make_imagine_inputs(1, N, 15)
    a = [params, actor, reward, value, x["belief"].to(dev), x["state"].to(dev), x["eps_action"].to(dev), x["eps_prior"].to(dev), 15]
    for _ in range(2):
        out = ops.imagine_fwd(*a, row_tile=rt)
    torch.cuda.synchronize()
else:
    B, T, rt = int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
    x = O.make_observe_inputs(1, T, B)
    g = lambda k: x[k].to(dev)
    for _ in range(2):
        ops.observe_fwd(params, g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"), g("eps_prior"), g("eps_post"), row_tile=rt)
    torch.cuda.synchronize()
print("done")

"""Forward imagine latency by row count for the row-tile choices (0 = library's pick, 16/32/64 = vm kernel, 128 = rows kernel):
where does the 128-row kernel start to win?  python scripts/tile_crossover.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from repo_b200 import ops, synth as O  # noqa: E402

dev = torch.device("cuda:0")
cu = lambda p: {k: v.to(dev) for k, v in p.items()}
params, actor = cu(O.make_transition_params(0)), cu(O.make_mlp_params(1, 230, 200, 12, 4))
reward, value = cu(O.make_mlp_params(2, 230, 200, 1, 3)), cu(O.make_mlp_params(3, 230, 200, 1, 3))
for N in (16, 128, 512, 1024, 2048, 2450, 4096, 18944):
    x = O.make_imagine_inputs(1, N, 15)
    a = [params, actor, reward, value, x["belief"].to(dev), x["state"].to(dev), x["eps_action"].to(dev), x["eps_prior"].to(dev), 15]
    line = f"N={N:6d}:"
    for rt in (0, 16, 32, 128):
        for _ in range(3):
            ops.imagine_fwd(*a, row_tile=rt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.imagine_fwd(*a, row_tile=rt)
        e1.record()
        torch.cuda.synchronize()
        line += f"  rt{rt}: {e0.elapsed_time(e1) / 10:7.3f} ms"
    print(line)

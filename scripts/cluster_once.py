"""One observe forward (with stash) and one backward of the cluster kernels at 50 sequences x 49 steps, for ncu."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cluster_bwd_bringup import fwd, bwd, cu, dev, O

T, B = 49, 50
params = cu(O.make_transition_params(1))
x = O.make_observe_inputs(2, T, B, p_done=0.05)
rs = np.random.RandomState(6)
G = [torch.from_numpy((4e-4 * rs.standard_normal((T, B, f))).astype(np.float32)).to(dev) for f in [200] + [30] * 6]
for _ in range(2):
    outs, st = fwd(params, x, True)
    bwd(params, x, outs, st, G, 1, True)
torch.cuda.synchronize()
print("done")

"""In-kernel clock64 stamps of the cluster observe backward (CTA 0 = rank 0 of cluster 0: an owner of state dimensions), from the
profiling build (python -m repo_b200.build --profiling).  Prints, averaged over the steps, when each event happens
relative to the top of the reverse step."""
import os, sys, ctypes as C
os.environ["REPO_B200_PROFILING"] = "1"
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from cluster_bwd_bringup import fwd, bwd, cu, dev, O
from repo_b200 import _lib

T, B = 49, 50
params = cu(O.make_transition_params(1))
x = O.make_observe_inputs(2, T, B, p_done=0.05)
outs, st = fwd(params, x, True)
rs = np.random.RandomState(6)
G = [torch.from_numpy((4e-4 * rs.standard_normal((T, B, f))).astype(np.float32)).to(dev) for f in [200] + [30] * 6]
bwd(params, x, outs, st, G, 1, True)
buf = torch.zeros(T * 32, dtype=torch.int64, device=dev)
_lib.lib().repo_b200_debug_clock(C.c_void_p(buf.data_ptr()))
bwd(params, x, outs, st, G, 1, True)
_lib.lib().repo_b200_debug_clock(None)
b = buf.cpu().numpy().reshape(T, 32)
names = {0: "epi: step top", 1: "epi: stage-1 slice sent", 16: "mma: DQP complete", 17: "mma: 2/3 issued", 2: "epi: acc 2/3 ready",
         3: "epi: DH slice sent", 18: "mma: DH complete", 19: "mma: 4 issued", 4: "epi: acc 4 ready", 5: "epi: DG slice sent",
         20: "mma: DG complete", 21: "mma: 6 issued", 6: "epi: acc 6 ready", 7: "epi: DE slice sent", 22: "mma: DE complete",
         23: "mma: 7 issued", 8: "epi: acc 7 ready"}
order = [0, 1, 16, 17, 2, 3, 18, 19, 4, 5, 20, 21, 6, 7, 22, 23, 8]
steps = range(5, T - 1)
rel = {s: np.mean([b[k, s] - b[k, 0] for k in steps]) for s in order}
prev = 0.0
for s in order:
    print(f"{names[s]:28s} +{rel[s]:8.0f} cycles   (delta {rel[s] - prev:7.0f})")
    prev = rel[s]
print(f"step period: {np.mean([b[k + 1, 0] - b[k, 0] for k in steps]):.0f} cycles")

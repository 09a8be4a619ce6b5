"""Per-stage timing of the rows kernel from in-kernel clock64 stamps (CTA 0)."""
import ctypes as C
import os
import sys

os.environ["REPO_B200_PROFILING"] = "1"   # the library variant built with -DRB_STAGE_CLOCK (python -m repo_b200.build --profiling)

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rssm_oracle as O  # noqa: E402
from repo_b200 import ops, _lib  # noqa: E402

dev = torch.device("cuda:0")
cu = lambda p: {k: v.to(dev) for k, v in p.items()}
OBSERVE = len(sys.argv) > 1 and sys.argv[1] == "observe"   # python scripts/stage_clock.py observe [sequences]: the observe program
if OBSERVE:
    sys.argv.pop(1)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 18944
params = cu(O.make_transition_params(0))
actor = cu(O.make_mlp_params(1, 230, 200, 12, 4))
reward = cu(O.make_mlp_params(2, 230, 200, 1, 3))
value = cu(O.make_mlp_params(3, 230, 200, 1, 3))
x = O.make_imagine_inputs(1, N, 15)
a = [params, actor, reward, value, x["belief"].to(dev), x["state"].to(dev), x["eps_action"].to(dev), x["eps_prior"].to(dev), 15]
_lib.lib().repo_b200_debug_flags(int(os.environ.get('RB_DBG', '0')))
if OBSERVE:
    xo = O.make_observe_inputs(1, 16, N)
    ao = [params] + [xo[k].to(dev) for k in ("prev_belief", "prev_state", "actions", "embeds", "nonterms", "eps_prior", "eps_post")]
    run = lambda: ops.observe_fwd(*ao, row_tile=128)
else:
    run = lambda: ops.imagine_fwd(*a, row_tile=128)
run()
torch.cuda.synchronize()
NS = 32
buf = torch.zeros(14 * NS * 2 + 256 + 1200, dtype=torch.int64, device=dev)
_lib.lib().repo_b200_debug_clock(C.c_void_p(buf.data_ptr()))
run()
torch.cuda.synchronize()
_lib.lib().repo_b200_debug_clock(None)
b = buf.cpu().numpy()
chunk = b[860:880].copy()
issuer = b[900:900 + 4 * NS].copy()
slabs = b[1100:1100 + 16].copy()
perwarp = b[1200:1200 + 16 * 64].copy().reshape(16, 32, 2)
extra = b[600:600 + 64].copy()
fine = b[700:700 + 8 * NS // 2 + 64].copy() if len(b) > 700 else None
b[600:] = 0
nz = (b != 0).sum() // (14 * 2)
b = b[: 14 * nz * 2].reshape(14, nz, 2)
print("stages per step:", nz)
names = ["E", "G0", "G1", "G2", "G3", "P1", "P2", "Q1", "Q2"] if OBSERVE else ["A1", "A2", "A3", "A4", "ACT", "E", "G0", "G1", "G2", "G3", "P1", "P2", "R1", "R2", "R3", "V1", "V2", "V3"]
t = 5
tot_epi = tot_mma = 0
for s in range(nz):
    epi = b[t, s, 1] - b[t, s, 0]
    prev_end = b[t, s - 1, 1] if s > 0 else b[t - 1, nz - 1, 1]
    mma = b[t, s, 0] - prev_end
    tot_epi += epi
    tot_mma += mma
    print(f"{names[s] if s < len(names) else s:>4}: mma_phase {mma:6d} cyc   epilogue {epi:6d} cyc   handoff->stage end {extra[2 * s] - b[t, s, 1]:6d}   prefetch {extra[2 * s + 1] - extra[2 * s]:6d}")
print("step total:", b[t + 1, 0, 0] - b[t, 0, 0], "cycles; mma", tot_mma, "epi", tot_epi)

if fine is not None and fine.any():
    print("# fine stamps of warp 4 (step 5): wait_acc = stage entry -> accumulators ready; sync = bias-staging barrier; body = epilogue math;")
    print("# st_wait = tcgen05.wait::st; proxy_fence = fence.proxy.async")
    for s in range(nz):
        f = fine[8 * s: 8 * s + 8]
        d = lambda a, c: int(f[c] - f[a]) if f[a] and f[c] else -1
        print(f"{names[s] if s < len(names) else s:>4}: wait_acc {d(0, 1):6d}  sync {d(1, 2):5d}  body {d(2, 3):6d}  (act_h loop {d(2, 5):6d} st_wait {d(5, 6):5d})  proxy_fence {d(3, 4):5d}")

if issuer.any():
    print("# MMA issuer (step 5), relative to the END of the previous stage's epilogue (warp 4): last k-slab group's inputs ready, its")
    print("# weights landed, stage committed; then the epilogue's view: accumulators ready")
    for s in range(nz):
        prev_end = b[t, s - 1, 1] if s > 0 else b[t - 1, nz - 1, 1]
        f = issuer[4 * s: 4 * s + 4]
        print(f"{names[s] if s < len(names) else s:>4}: inputs_ready {int(f[1] - prev_end):7d}  weights_landed {int(f[2] - prev_end):7d}  committed {int(f[3] - prev_end):7d}  epilogue_begins {int(b[t, s, 0] - prev_end):7d}")

if slabs.any():
    e0 = b[t, 0, 0]   # A1's epilogue begins (A2's MMAs are K-chained behind it)
    print("# stage A2: issue time of each k-slab's three MMAs relative to the BEGIN of A1's epilogue (which ends at %d; A2's epilogue begins at %d):" % (b[t, 0, 1] - e0, b[t, 1, 0] - e0))
    print("#  ", " ".join(f"{int(v - e0):6d}" for v in slabs if v))

if perwarp.any():
    print("# per epilogue warp (step 5): epilogue END of each stage relative to warp 4's (quadrant = warp % 4, part = warp // 4);")
    print("# rows = stages, columns = warps 0..15 of the epilogue (quadrant-major within a part)")
    for s in range(nz):
        ref = b[t, s, 1]
        print(f"{names[s] if s < len(names) else s:>4}: " + " ".join(f"{int(perwarp[w, s, 1] - ref):6d}" for w in range(16)))

if chunk.any():
    c = chunk[chunk != 0]
    print("# G1, warp 4: entry, handed off, [half j: math done, beliefs stored, scratch stored] x2, fenced:", [int(v - chunk[0]) if v else 0 for v in chunk[:9]])

"""Bring-up of the cluster observe backward (mode 1) against the per-sequence fp32 kernel (mode 2): every pre-activation
gradient through the C-ABI on identical inputs (tiny and normal gradient magnitudes), then timing at 50 x 49."""
import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from repo_b200 import ops, _lib
from oracle import rssm_oracle as O

dev = torch.device("cuda:0")
cu = lambda p: {k: v.to(dev) for k, v in p.items()}
L = _lib.lib()
p = ops._ptr


def fwd(params, x, with_obs):
    g = lambda k: None if x[k] is None else x[k].to(dev)
    T1, B = x["actions"].shape[:2]
    d = ops.dims_of(params)
    st = torch.zeros(T1, B, 5 * d.belief + 2 * d.hidden, device=dev)
    outs, kl, _ = ops.observe_fwd(params, g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"),
                                  g("eps_prior"), g("eps_post"), stash=st)
    return outs, st


def bwd(params, x, outs, st, G, mode, with_obs, reps=1):
    g = lambda k: None if x[k] is None else x[k].to(dev)
    T1, B = x["actions"].shape[:2]
    d = ops.dims_of(params)
    keep = ops._Keep()
    W = ops.rssm_struct(params, keep)
    D, S, Hd = d.belief, d.state, d.hidden
    mk = lambda f: torch.zeros(T1, B, f, device=dev)
    res = dict(d_q=mk(2 * S), d_hq=mk(Hd), d_p=mk(2 * S), d_hp=mk(Hd), d_gi=mk(3 * D), d_gh=mk(3 * D), d_e=mk(D),
               d_b0=torch.zeros(B, D, device=dev), d_s0=torch.zeros(B, S, device=dev))
    ws = torch.empty(L.repo_b200_observe_bwd_workspace_bytes(C.byref(d), B), dtype=torch.uint8, device=dev)
    nt = g("nonterms")
    nt = None if nt is None else nt.reshape(T1, B).contiguous()
    pb, e1_, e2_ = g("prev_belief"), g("eps_prior"), g("eps_post")   # kept alive: only raw pointers cross the C-ABI
    GG = list(G) + [None] * (7 - len(G))
    post_sd = outs[6] if with_obs else None
    def call():
        rc = L.repo_b200_observe_bwd_ws(
            C.byref(d), C.byref(W), p(pb), p(outs[0]), p(outs[3]), p(post_sd), p(e1_), p(e2_),
            p(nt), p(st), *[p(t) for t in GG], p(res["d_q"]) if with_obs else None, p(res["d_hq"]) if with_obs else None,
            p(res["d_p"]), p(res["d_hp"]), p(res["d_gi"]), p(res["d_gh"]), p(res["d_e"]), p(res["d_b0"]), p(res["d_s0"]),
            T1, B, int(with_obs), ops.act_kind("elu"), 0.1, p(ws), ws.numel(), mode, ops._stream())
        _lib.check(rc, "observe_bwd_ws")
    call()
    torch.cuda.synchronize()
    ms = None
    if reps > 1:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            call()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
    return res, ms


if __name__ == "__main__":
    for (T, B, pd, with_obs, gscale) in [(3, 5, 0.0, True, 1.0), (8, 10, 0.3, True, 1e-4), (49, 50, 0.1, True, 4e-4), (7, 33, 0.2, False, 1.0), (5, 130, 0.1, True, 1e-6)]:
        params = cu(O.make_transition_params(100 + T))
        x = O.make_observe_inputs(200 + B, T, B, p_done=pd)
        if not with_obs:
            x["embeds"] = None; x["eps_post"] = None
        outs, st = fwd(params, x, with_obs)
        rs = np.random.RandomState(5)
        feat = [200] + [30] * 6
        G = [torch.from_numpy((gscale * rs.standard_normal((T, B, f))).astype(np.float32)).to(dev) for f in feat[:7 if with_obs else 4]]
        r1, _ = bwd(params, x, outs, st, G, 1, with_obs)
        r2, _ = bwd(params, x, outs, st, G, 2, with_obs)
        msg = []
        for k in r1:
            if not with_obs and k in ("d_q", "d_hq"):
                continue
            sc = r2[k].abs().max().item() + 1e-30
            msg.append(f"{k} {((r1[k] - r2[k]).abs().max().item() / sc):.2e}")
        print(f"T={T} B={B} obs={with_obs} g~{gscale:g}: max|cluster - fp32| / max|fp32|: " + "  ".join(msg), flush=True)

    params = cu(O.make_transition_params(1))
    x = O.make_observe_inputs(2, 49, 50, p_done=0.05)
    outs, st = fwd(params, x, True)
    rs = np.random.RandomState(6)
    G = [torch.from_numpy((4e-4 * rs.standard_normal((49, 50, f))).astype(np.float32)).to(dev) for f in [200] + [30] * 6]
    for mode in (1, 2):
        _, ms = bwd(params, x, outs, st, G, mode, True, reps=20)
        print(f"observe backward 50x49 mode={mode}: {ms:.3f} ms = {ms*1e3/49:.1f} us per time step", flush=True)

"""GPU bring-up report: linear -> observe -> imagine against the CPU oracle.  Prints errors
instead of asserting so one gpurun call tells as much as possible."""
import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import rssm_oracle as O  # noqa: E402
from repo_b200 import ops  # noqa: E402
from tests import _cases as C  # noqa: E402

dev = torch.device("cuda:0")


def err(a, b):
    a = a.detach().double().cpu()
    b = b.detach().double().cpu()
    d = (a - b).abs()
    bad = (d > 1e-3 * b.abs() + 1e-4).double().mean().item()
    return "max_abs=%.2e max_rel=%.2e frac_out_of_tol=%.4f nan=%d" % (
        d.max().item(), (d / (b.abs() + 1e-6)).max().item(), bad, int(torch.isnan(a).sum()))


def cu(p):
    return {k: v.to(dev) for k, v in p.items()}


def test_linear():
    for rows, in_f, out_f, rt in [(64, 32, 16, 64), (64, 208, 200, 64), (50, 236, 200, 0), (37, 1024, 200, 16),
                                  (300, 230, 200, 32), (2450, 1024, 200, 0), (16, 16, 128, 16), (16, 48, 1, 16)]:
        g = torch.Generator().manual_seed(rows * 7 + in_f)
        x = torch.randn(rows, in_f, generator=g)
        w = torch.randn(out_f, in_f, generator=g) / in_f ** 0.5
        b = torch.randn(out_f, generator=g)
        want = x.double() @ w.double().t() + b.double()
        try:
            y = ops.linear(x.to(dev), w.to(dev), b.to(dev), row_tile=rt)
            torch.cuda.synchronize()
            print(f"linear rows={rows} in={in_f} out={out_f} rt={rt}: {err(y, want)}", flush=True)
        except Exception as e:
            print(f"linear rows={rows} in={in_f} out={out_f} rt={rt}: EXC {e}", flush=True)
            raise


def test_observe():
    for name in C.OBSERVE_CASES:
        params, x, gold, meta = C.observe_case(name)
        want = O.observe(params, x["prev_belief"], x["prev_state"], x["actions"], x["embeds"], x["nonterms"],
                         x["eps_prior"], x["eps_post"])
        g = lambda k: None if x[k] is None else x[k].to(dev)
        for rt in (64, 128):
            outs, kl, _ = ops.observe_fwd(cu(params), g("prev_belief"), g("prev_state"), g("actions"), g("embeds"),
                                          g("nonterms"), g("eps_prior"), g("eps_post"), row_tile=rt)
            torch.cuda.synchronize()
            for nm, o, w_ in zip(C.OBS_NAMES, outs, want):
                print(f"observe {name} rt={rt} {nm}: {err(o, w_)}", flush=True)
            if kl is not None:
                print(f"observe {name} rt={rt} kl: {err(kl, O.kl_sum(want[5], want[6], want[2], want[3]))}", flush=True)


def test_imagine():
    for name in C.IMAGINE_CASES:
        params, actor, reward, value, x, gold, meta = C.imagine_case(name)
        H = int(meta["H"])
        want = O.imagine(params, actor, x["belief"], x["state"], x["eps_action"], x["eps_prior"], H)
        for rt in (64, 128):
            out = ops.imagine_fwd(cu(params), cu(actor), cu(reward), cu(value), x["belief"].to(dev), x["state"].to(dev),
                                  x["eps_action"].to(dev), x["eps_prior"].to(dev), H, row_tile=rt)
            torch.cuda.synchronize()
            for nm, w_ in zip(C.IMG_NAMES + ["actions"], want):
                print(f"imagine {name} rt={rt} {nm}: {err(out[nm], w_)}", flush=True)
            for nm in ("rewards", "values", "returns"):
                print(f"imagine {name} rt={rt} {nm}: {err(out[nm], C.t(gold[nm]))}", flush=True)


def quick_timing():
    params, actor, reward, value, x, gold, meta = C.imagine_case("imagine_N8_H15")
    for N in (2450, 16384, 75776):
        xi = O.make_imagine_inputs(1, N, 15)
        a = [cu(params), cu(actor), cu(reward), cu(value), xi["belief"].to(dev), xi["state"].to(dev),
             xi["eps_action"].to(dev), xi["eps_prior"].to(dev), 15]
        for rt in (32, 64, 128):
            out = ops.imagine_fwd(*a, row_tile=rt)
            ws = out["workspace"]
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(3):
                ops.imagine_fwd(*a, row_tile=rt, workspace=ws, packed=True)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 3
            print(f"timing imagine N={N} rt={rt}: {ms:.3f} ms  {N * 14 / ms * 1e3:.3e} steps/s", flush=True)
    p2, x2, _, _ = C.observe_case("observe_default_tail")
    g = lambda k: x2[k].to(dev)
    a = [cu(p2), g("prev_belief"), g("prev_state"), g("actions"), g("embeds"), g("nonterms"), g("eps_prior"), g("eps_post")]
    for rt in (16, 64, 128):
        ops.observe_fwd(*a, row_tile=rt)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.observe_fwd(*a, row_tile=rt)
        e1.record()
        torch.cuda.synchronize()
        print(f"timing observe T=50 B=50 rt={rt}: {e0.elapsed_time(e1) / 3:.3f} ms", flush=True)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), flush=True)
    if os.environ.get("RB_DBG"):
        from repo_b200 import _lib
        _lib.lib().repo_b200_debug_flags(int(os.environ["RB_DBG"]))
        print("debug flags", os.environ["RB_DBG"], flush=True)
    which = sys.argv[1:] or ["linear", "observe", "imagine", "timing"]
    for nm, fn in [("linear", test_linear), ("observe", test_observe), ("imagine", test_imagine), ("timing", quick_timing)]:
        if nm in which:
            try:
                fn()
            except Exception:
                traceback.print_exc()
                print(f"SECTION {nm} FAILED", flush=True)
                break

#!/usr/bin/env python
"""bench.py — imagined latent steps/s of the RSSM hot path (BASELINE.json metric) on N GPUs.

A "step" = one pass of the hot path over one batch of synthetic input: `TransitionModel.imagine`
(actor -> GRU belief -> prior head, reward + value heads on every imagined state, lambda-return) over
ROWS start states per GPU for horizon 15, RePo default sizes (belief 200, state 30, hidden 200,
action 6).  Workload = BASELINE configs[4] (large-batch imagination sweep) at a fixed per-GPU row
count (weak scaling: rows are independent, no data-path collective); the RePo default-shape numbers
(configs[1]: 2450 rows x 14 steps imagine, 49 x 50 observe) ride along under "default_shape".

  python bench.py --gpus N --steps K --warmup W            (torchrun launches N ranks for N > 1)
  python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference)
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_STEP = 1_440_000      # algorithmic forward FLOPs per imagined latent step
BYTES_PER_STEP = 1_312         # algorithmic HBM bytes per imagined latent step
HORIZON = 15
# From the committed `ncu --set full` capture of this kernel at the bench shape (profiles/NCU_CAPTURE below): DRAM traffic of one
# launch (dram__bytes_read.sum + dram__bytes_write.sum) and the tensor-pipe activity.  ncu cannot run inside a timed bench, so
# these are constants tied to that file; everything else in `roofline` is measured live.
NCU_CAPTURE = {"file": "profiles/r02_rssm_rows_kernel_75776x14_ncu_full.txt", "rows": 75776,
               "traffic_bytes": 235_469_312 + 1_442_514_000, "tensor_pipe_active_pct": 54.5,
               "l2_to_sm_bytes": 27_600_537_000, "kernel_ms_under_ncu": 4.177}
MMA_ISSUE_FACTOR = 3.3   # tensor-pipe MACs issued per algorithmic MAC: 3 fp16 products per fp32-grade product x ~1.1 K/N padding
DIMS = dict(belief=200, state=30, action=6, hidden=200, embed=1024)
SWEEP_ROWS = (16384, 65536, 262144, 1048576)   # SURVEY 8(d) Config 5: total start states, split evenly over the GPUs


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="repo_b200", choices=["repo_b200", "reference"])
    ap.add_argument("--rows-per-gpu", type=int, default=75776, help="start states per GPU; default = 4 waves of 148 SMs x 128-row tiles")
    ap.add_argument("--cpu-rows", type=int, default=4096, help="rows per step of the bounded CPU sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the sweep / default-shape / data-parallel update blocks")
    ap.add_argument("--workload", default="imagine", choices=["imagine", "update"],
                    help="imagine = the headline sweep (default, what the driver runs); update = one full training iteration "
                         "(Agent.train_dynamics + train_actor_critic) with the replay batch sharded over the ranks (SURVEY 8d Config 2-4)")
    ap.add_argument("--algo", default="repo", choices=["repo", "dreamer", "tia"], help="--workload update: trainer wiring")
    ap.add_argument("--global-batch", type=int, default=50, help="--workload update: sequences per iteration over all ranks")
    return ap.parse_args()


def peaks():
    """(burst bf16 TF/s, sustained bf16 TF/s, HBM GB/s, source)"""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops", 1621.6), d.get("bf16_tflops_sustained", 1384.0), d.get("hbm_gbs", 6553.0), "measured (MEASURED_PEAKS.json)"
    return 1590.0, 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v == "Active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU arm
def cpu_imagine_step(O, params, actor, reward, value, x):
    outs = O.imagine(params, actor, x["belief"], x["state"], x["eps_action"], x["eps_prior"], HORIZON)
    T, N = outs[0].shape[:2]
    rew = O.head_forward(reward, outs[0].flatten(0, 1), outs[1].flatten(0, 1)).reshape(T, N)
    val = O.head_forward(value, outs[0].flatten(0, 1), outs[1].flatten(0, 1)).reshape(T, N)
    return O.imagine_returns(rew, val, 0.99, 0.95)


REF_DIR = "/root/reference"


def _reference_step_fn(rows):
    """The UNMODIFIED reference (TransitionModel.imagine + RewardModel / ValueModel + lambda_return, loaded by file path as
    oracle/make_golden.py does) when /root/reference exists on this machine; None on the GPU box, where it does not travel."""
    if not os.path.isdir(REF_DIR):
        return None
    try:
        from oracle import make_golden as MG
        from oracle import rssm_oracle as O
        R = MG.load_reference()
    except Exception:
        return None
    D, S, A, Hd = DIMS["belief"], DIMS["state"], DIMS["action"], DIMS["hidden"]
    tm = R.rssm.TransitionModel(D, S, A, Hd, DIMS["embed"], "elu")
    tm.load_state_dict(O.make_transition_params(0))
    actor = R.actor_critic.ActorModel(D, S, Hd, A, "elu")
    actor.load_state_dict(O.make_mlp_params(1, D + S, Hd, 2 * A, 4))
    reward = R.decoder.RewardModel(D, S, Hd, "elu")
    reward.load_state_dict(O.make_mlp_params(2, D + S, Hd, 1, 3))
    value = R.actor_critic.ValueModel(D, S, Hd, "elu")
    value.load_state_dict(O.make_mlp_params(3, D + S, Hd, 1, 3))
    x = O.make_imagine_inputs(4, rows, HORIZON)
    bottle = R.common_utils.bottle if hasattr(R.common_utils, "bottle") else None

    def step():   # dreamer.py:312-349 without the gradient bookkeeping
        beliefs, states, _, _ = tm.imagine(x["belief"], x["state"], actor, HORIZON)
        flat = lambda m: m(beliefs.flatten(0, 1), states.flatten(0, 1)).reshape(beliefs.shape[0], -1)
        rew, val = flat(reward), flat(value)
        disc = 0.99 * torch.ones_like(rew[:-1])
        return R.common_utils.lambda_return(rew[:-1], val[:-1], disc, val[-1], 0.95)

    return step


def cpu_baseline(rows, steps, warmup):
    """The reference's own CPU implementation of the path, on all host threads, on a bounded sample of the workload: the
    unmodified reference modules when /root/reference is present ("reference"), else the oracle, a torch-CPU restatement
    of it (nn.Linear / GRUCell arithmetic via ATen/MKL; "port")."""
    from oracle import rssm_oracle as O
    torch.set_num_threads(os.cpu_count() or 1)
    step = _reference_step_fn(rows)
    kind = "reference" if step is not None else "port"
    if step is None:
        params = O.make_transition_params(0)
        actor = O.make_mlp_params(1, 230, 200, 12, 4)
        reward = O.make_mlp_params(2, 230, 200, 1, 3)
        value = O.make_mlp_params(3, 230, 200, 1, 3)
        x = O.make_imagine_inputs(4, rows, HORIZON)
        step = lambda: cpu_imagine_step(O, params, actor, reward, value, x)
    with torch.no_grad():
        for _ in range(warmup):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = time.perf_counter() - t0
    return rows * (HORIZON - 1) * steps / dt, dt / steps, torch.get_num_threads(), kind


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(a.steps, 10))
    v, sec, cores, kind = cpu_baseline(a.cpu_rows, steps, max(1, min(a.warmup, 2)))
    sample = f"{a.cpu_rows} start rows x {HORIZON - 1} steps per step (bounded sample of the {a.rows_per_gpu}-row/GPU workload)"
    line = {
        "impl": "reference", "metric": "imagined latent steps/sec", "value": v, "unit": "steps/s", "n_gpus": a.gpus,
        "steps": steps, "warmup": max(1, min(a.warmup, 2)), "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(a),
        "cpu_baseline": {"value": v, "unit": "steps/s", "cores": cores, "kind": kind, "sample": sample},
        "e2e": {"value": v, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(a):
    return {"workload": f"imagine sweep (BASELINE configs[4]): {a.rows_per_gpu} start states per GPU x horizon {HORIZON}, "
                        "RePo RSSM default sizes (belief 200, state 30, hidden 200, action 6), actor + reward + value heads + lambda-return",
            "rows_per_gpu": a.rows_per_gpu, "horizon": HORIZON, "parallelism": f"rows sharded x{a.gpus}, no data-path collective",
            "l2": "per-step inputs+outputs (~1.4 GB) exceed the 126 MB L2; no explicit flush"}


# ------------------------------------------------------------------------------------------ GPU arm
def run_gpu(a):
    import torch.distributed as dist
    from repo_b200 import synth as O         # seeded synthetic weights/inputs (numpy RandomState); no oracle on this arm
    from repo_b200 import ops, _lib
    from repo_b200.models import ActorModel, RewardModel, ValueModel
    from repo_b200.rssm import TransitionModel

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != a.gpus:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch with torchrun --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        # every rank drives its own pinned-host -> device copies in the end-to-end loop: give each an own slice of the host
        # cores instead of letting all of them pile onto the launcher's affinity mask (round 1: e2e efficiency 0.966 at N = 8)
        try:
            cores = sorted(os.sched_getaffinity(0))
            per = max(1, len(cores) // world)
            os.sched_setaffinity(0, set(cores[local * per:(local + 1) * per]) or set(cores))
        except (AttributeError, OSError):
            pass

    N, T = a.rows_per_gpu, HORIZON - 1
    D, S, A, Hd = DIMS["belief"], DIMS["state"], DIMS["action"], DIMS["hidden"]
    # random-init weights of the reference architecture (seeded numpy; every rank holds a replica)
    model = TransitionModel(D, S, A, Hd, DIMS["embed"], "elu").to(dev)
    model.load_state_dict(O.make_transition_params(0))
    actor = ActorModel(D, S, Hd, A, "elu").to(dev)
    actor.load_state_dict(O.make_mlp_params(1, D + S, Hd, 2 * A, 4))
    reward = RewardModel(D, S, Hd, "elu").to(dev)
    reward.load_state_dict(O.make_mlp_params(2, D + S, Hd, 1, 3))
    value = ValueModel(D, S, Hd, "elu").to(dev)
    value.load_state_dict(O.make_mlp_params(3, D + S, Hd, 1, 3))
    named = lambda m: {k: v.detach() for k, v in m.named_parameters()}
    P, PA, PR, PV = named(model), named(actor), named(reward), named(value)

    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    belief = (torch.randn(N, D, device=dev, generator=g) * 0.3).clamp_(-1, 1)
    state = torch.randn(N, S, device=dev, generator=g)
    eps_a = torch.randn(T, N, A, device=dev, generator=g)
    eps_p = torch.randn(T, N, S, device=dev, generator=g)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ws = None

    def step():
        nonlocal ws
        out = ops.imagine_fwd(P, PA, PR, PV, belief, state, eps_a, eps_p, HORIZON, workspace=ws)  # packs weights every step
        ws = out["workspace"]
        return out

    # ---- device-resident throughput (value) ----
    for _ in range(max(a.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        step()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # kernel-only duration of the layer machine (the roofline numerator): the SAME back-to-back loop with the weights
    # already packed, i.e. a.steps launches of rssm_rows_kernel alone between two events on the launching stream
    last_out = None
    for _ in range(3):   # same retention pattern as the timed loop (two output sets alive), so that it allocates nothing
        last_out = ops.imagine_fwd(P, PA, PR, PV, belief, state, eps_a, eps_p, HORIZON, workspace=ws, packed=True)
    barrier()
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    for _ in range(a.steps):
        last_out = ops.imagine_fwd(P, PA, PR, PV, belief, state, eps_a, eps_p, HORIZON, workspace=ws, packed=True)
    k1.record()
    barrier()
    kernel_ms = k0.elapsed_time(k1) / a.steps

    # ---- end to end through the module API with HOST start states ----
    # Each step: H2D of that step's start states from pinned host memory, TransitionModel.imagine (device noise
    # draw + fused kernel), D2H of the lambda-returns.  Inputs are double-buffered on a copy stream so step i+1's
    # upload overlaps step i's kernel (the timed region still contains every copy of every step).
    hb = belief.cpu().pin_memory()
    hs = state.cpu().pin_memory()
    hret = [torch.empty(T - 1, N, dtype=torch.float32).pin_memory() for _ in range(2)]
    dbuf = [(torch.empty_like(belief), torch.empty_like(state)) for _ in range(2)]
    copy_stream = torch.cuda.Stream()
    main_stream = torch.cuda.current_stream()
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i & 1])        # the kernel that last read this buffer is done
            dbuf[i & 1][0].copy_(hb, non_blocking=True)
            dbuf[i & 1][1].copy_(hs, non_blocking=True)
            ready[i & 1].record(copy_stream)

    def e2e_loop(n):
        for ev in consumed:
            ev.record(main_stream)
        upload(0)
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)
            main_stream.wait_event(ready[i & 1])
            db, ds = dbuf[i & 1]
            with torch.no_grad():   # the public call: noise is drawn on the device inside imagine()
                traj, extra = model.imagine(db, ds, actor, HORIZON, reward_model=reward, value_model=value, return_extras=True)
            consumed[i & 1].record(main_stream)
            hret[i & 1].copy_(extra["returns"], non_blocking=True)

    model.prefetch_noise = True   # module option: the next call's noise is drawn on a side stream under this call's kernel
    e2e_loop(3)
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    e2e_loop(a.steps)
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    model.prefetch_noise = False
    clocks = sampler.stop() if rank == 0 else None  # sampled across both timed loops (device-resident and e2e)

    t = torch.tensor([ms, e2e_ms, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms, e2e_ms, kernel_ms = [float(v) for v in t.tolist()]

    # ---- sampled parity of the timed launch itself: 256 random rows of this rank's 75,776 vs the oracle (checker only;
    # rows are independent, so the oracle runs on just those rows with the same weights and noise; outside every timed region)
    parity = None
    if rank == 0 and not a.no_extras:
        from oracle import rssm_oracle as ORC
        gi = torch.Generator().manual_seed(7)
        idx = torch.randperm(N, generator=gi)[:256].sort().values
        di = idx.to(dev)
        cpu = lambda t_: t_.detach().float().cpu()
        Pc, PAc, PRc, PVc = ({k: cpu(v) for k, v in d_.items()} for d_ in (P, PA, PR, PV))
        want = ORC.imagine(Pc, PAc, cpu(belief[di]), cpu(state[di]), cpu(eps_a[:, di]), cpu(eps_p[:, di]), HORIZON)
        rew = ORC.head_forward(PRc, want[0].flatten(0, 1), want[1].flatten(0, 1)).reshape(T, -1)
        val = ORC.head_forward(PVc, want[0].flatten(0, 1), want[1].flatten(0, 1)).reshape(T, -1)
        wret = ORC.imagine_returns(rew, val, 0.99, 0.95)
        got = {"beliefs": last_out["beliefs"][:, di], "prior_states": last_out["prior_states"][:, di],
               "prior_means": last_out["prior_means"][:, di], "prior_std_devs": last_out["prior_std_devs"][:, di],
               "returns": last_out["returns"][:, di]}
        ref = {"beliefs": want[0], "prior_states": want[1], "prior_means": want[2], "prior_std_devs": want[3], "returns": wret}
        worst = {}
        for k_, w_ in ref.items():
            err = (cpu(got[k_]) - w_).abs()
            worst[k_] = float((err / (1e-5 + 1e-3 * w_.abs())).max())   # > 1 = outside rtol 1e-3 (+ atol 1e-5)
        parity = {"rows_checked": 256, "of_rows": N, "steps": T, "tolerance": "rtol 1e-3 + atol 1e-5 (north_star: rtol 1e-3)",
                  "worst_error_over_tolerance": worst, "ok": all(v < 1.0 for v in worst.values())}
        del want, got, ref

    # ---- SURVEY 8(d) Config 5: total start states split evenly over the GPUs (device-timed, max over ranks) ----
    sweep = None
    if not a.no_extras:
        sweep = {}
        for total in SWEEP_ROWS:
            n_loc = total // world
            gs_ = torch.Generator(device=dev).manual_seed(99 + rank)
            sb = (torch.randn(n_loc, D, device=dev, generator=gs_) * 0.3).clamp_(-1, 1)
            ss = torch.randn(n_loc, S, device=dev, generator=gs_)
            sea = torch.randn(T, n_loc, A, device=dev, generator=gs_)
            sep = torch.randn(T, n_loc, S, device=dev, generator=gs_)
            sw = None
            for _ in range(2):
                sw = ops.imagine_fwd(P, PA, PR, PV, sb, ss, sea, sep, HORIZON, workspace=sw)["workspace"]
            barrier()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 5 if total <= 262144 else 3
            s0.record()
            for _ in range(reps):
                ops.imagine_fwd(P, PA, PR, PV, sb, ss, sea, sep, HORIZON, workspace=sw)
            s1.record()
            barrier()
            tt = torch.tensor([s0.elapsed_time(s1) / reps], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            sweep[str(total)] = {"rows_per_gpu": n_loc, "ms": float(tt.item()), "steps_per_s": n_loc * world * T / float(tt.item()) * 1e3}
            del sb, ss, sea, sep, sw
        torch.cuda.empty_cache()

    # ---- the communicating path (N > 1): one data-parallel training iteration on ALL ranks ----
    dp_update = None
    if world > 1 and not a.no_extras:
        dp_update = dp_update_block(dev, rank, world, dist)

    # ---- RePo default shapes (configs[1]) : forward latency of the two kernels ----
    # (single-process runs only: the trainer-level updates all-reduce their gradient buckets, which would dead-lock
    # against the idle ranks of a multi-GPU imagine sweep)
    default_shape = None
    if rank == 0 and world == 1 and not a.no_extras:
        x = O.make_imagine_inputs(5, 2450, HORIZON)
        xa = [x["belief"].to(dev), x["state"].to(dev), x["eps_action"].to(dev), x["eps_prior"].to(dev)]
        xo = O.make_observe_inputs(6, 50, 50)
        go = lambda k: xo[k].to(dev)
        oa = [go("prev_belief"), go("prev_state"), go("actions"), go("embeds"), go("nonterms"), go("eps_prior"), go("eps_post")]

        def timeit(fn, n=10):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(n):
                fn()
            s1.record()
            torch.cuda.synchronize()
            return s0.elapsed_time(s1) / n

        img_ms = timeit(lambda: ops.imagine_fwd(P, PA, PR, PV, *xa, HORIZON))
        obs_ms = timeit(lambda: ops.observe_fwd(P, *oa))   # 50 sequences: the 16-CTA cluster kernel (csrc/cluster.cuh)

        # the same pass under autograd, as Agent.train_dynamics runs it: forward with the activation stash, BPTT through
        # the cluster backward kernel (csrc/cluster_bwd.cuh), the seven weight-gradient GEMMs and the bias sums
        from repo_b200.rssm import TransitionModel
        tm_ = TransitionModel(D, S, A, Hd, 1024, "elu").to(dev)
        tm_.load_state_dict({k_: v_ for k_, v_ in P.items()})
        gsum_ = [torch.randn_like(t_) * 4e-4 for t_ in ops.observe_fwd(P, *oa)[0]]

        def obs_train():
            for p_ in tm_.parameters():
                p_.grad = None
            outs_ = tm_.observe(oa[0], oa[1], oa[2], oa[3], oa[4], eps_prior=oa[5], eps_post=oa[6])
            torch.autograd.backward(list(outs_), gsum_)

        obs_train_ms = timeit(obs_train)
        del tm_, gsum_

        # observe at a batch that fills the GPU (the other half of the hot path; SURVEY §8d "Roofline - observe":
        # 1,112,000 FLOP and 5,884 B per row-step): 18,944 sequences x 49 steps through the 128-row kernel
        OB, OT = 18944, 50
        gen = torch.Generator(device=dev).manual_seed(11)
        rn = lambda *sh: torch.randn(*sh, device=dev, generator=gen)
        big = [torch.zeros(OB, D, device=dev), torch.zeros(OB, S, device=dev), rn(OT - 1, OB, A).clamp_(-1, 1), rn(OT - 1, OB, 1024),
               torch.ones(OT - 1, OB, 1, device=dev), rn(OT - 1, OB, S), rn(OT - 1, OB, S)]
        big_ms = timeit(lambda: ops.observe_fwd(P, *big), 3)
        big_steps = OB * (OT - 1)
        observe_large = {"sequences": OB, "steps": OT - 1, "ms": big_ms, "row_steps_per_s": big_steps / big_ms * 1e3,
                         "tflops_algorithmic": big_steps * 1_112_000 / big_ms / 1e9,
                         "hbm_gbs_algorithmic": big_steps * 5884 / big_ms / 1e6}
        # sampled parity of this very configuration: 64 random sequences of the 18,944 against the oracle (checker only;
        # sequences are independent, so the oracle runs on just those columns with the same weights and noise)
        from oracle import rssm_oracle as ORC
        outs_big, kl_big, _ = ops.observe_fwd(P, *big)
        ci = torch.randperm(OB, generator=torch.Generator().manual_seed(13))[:64].sort().values
        cd = ci.to(dev)
        cpu_ = lambda t_: t_.detach().float().cpu()
        Pc_ = {k_: cpu_(v_) for k_, v_ in P.items()}
        want_o = ORC.observe(Pc_, cpu_(big[0][cd]), cpu_(big[1][cd]), cpu_(big[2][:, cd]), cpu_(big[3][:, cd]), cpu_(big[4][:, cd]),
                             cpu_(big[5][:, cd]), cpu_(big[6][:, cd]))
        worst_o = {}
        for nm_, got_, w_ in zip(("beliefs", "prior_states", "prior_means", "prior_std_devs", "posterior_states", "posterior_means",
                                  "posterior_std_devs"), outs_big, want_o):
            worst_o[nm_] = float(((cpu_(got_[:, cd]) - w_).abs() / (1e-5 + 1e-3 * w_.abs())).max())
        wkl = ORC.kl_sum(want_o[5], want_o[6], want_o[2], want_o[3])
        worst_o["kl"] = float(((cpu_(kl_big[:, cd]) - wkl).abs() / (1e-4 + 1e-3 * wkl.abs())).max())
        observe_large["parity_sample"] = {"sequences_checked": 64, "tolerance": "rtol 1e-3 + atol 1e-5 (kl: atol 1e-4; it sums 30 terms)",
                                          "worst_error_over_tolerance": worst_o, "ok": all(v < 1.0 for v in worst_o.values())}
        del big, outs_big, kl_big, want_o

        # One full training iteration at the RePo default shapes (SURVEY §8d Metric 2 / Config 2), through the
        # trainer-level API (repo_b200/trainer.py): conv encoder -> observe (BPTT over 49 steps) -> conv decoder /
        # reward head / KL -> hand-written backward passes -> global-norm clip + Adam on flat buckets; then
        # Dreamer.train_actor_critic (imagine, heads, 100-sample entropy, lambda-return, both backward passes + Adam).
        from repo_b200.trainer import Agent, Config
        batch = {k: v.to(dev) for k, v in O.make_train_batch(7, 50, 50, A).items()}
        upd = {}
        for algo in ("repo", "dreamer", "tia"):
            agent = Agent(Config(), A, algo=algo, device=dev)
            agent.transition_model.load_state_dict(O.make_transition_params(1))
            if algo == "tia":
                agent.distractor_transition_model.load_state_dict(O.make_transition_params(2))
            agent.optimizers()
            state = {}

            def wm_update():
                state["b"], state["s"] = agent.train_dynamics(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])

            def ac_update():
                agent.train_actor_critic(state["b"].flatten(0, 1), state["s"].flatten(0, 1))

            upd[algo + "_world_model_update_ms"] = timeit(wm_update, 5)
            if algo == "repo":
                upd["actor_critic_update_ms"] = timeit(ac_update, 5)
                lat = list(agent.init_latent_and_action())
                frame = batch["obs"][0, :1].contiguous()

                def act_step():
                    lat[0], lat[1], lat[2] = agent.update_latent_and_select_action(lat[0], lat[1], lat[2], frame)

                upd["acting_step_ms"] = timeit(act_step, 20)
                # the same three calls captured once and replayed as CUDA graphs (repo_b200/graphs.py)
                g_act = agent.graphed_acting(frame)

                def act_graph():
                    lat[0], lat[1], lat[2] = g_act(lat[0], lat[1], lat[2], frame)

                upd["acting_step_graphed_ms"] = timeit(act_graph, 50)
            g_wm, g_ac = agent.graphed(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])
            gs = {}

            def wm_graph():
                gs["b"], gs["s"] = g_wm(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])

            upd[algo + "_world_model_update_graphed_ms"] = timeit(wm_graph, 10)
            if algo == "repo":
                upd["actor_critic_update_graphed_ms"] = timeit(lambda: g_ac(gs["b"].flatten(0, 1), gs["s"].flatten(0, 1)), 10)
            del g_wm, g_ac
            del agent
        ac_ms = upd["actor_critic_update_ms"]
        default_shape = {"imagine_2450x14_ms": img_ms, "imagine_steps_per_s": 2450 * 14 / img_ms * 1e3,
                         "observe_49x50_ms": obs_ms, "observe_row_steps_per_s": 2450 / obs_ms * 1e3,
                         "observe_us_per_time_step": obs_ms * 1e3 / 49, "observe_fwd_bwd_49x50_ms": obs_train_ms,
                         "observe_large_batch": observe_large,
                         **upd, "actor_critic_update_steps_per_s": 2450 * 14 / ac_ms * 1e3,
                         "note": "world_model_update = Agent.train_dynamics on a (50,50,3,64,64) batch: conv encoder + observe + conv "
                                 "decoder + reward/KL losses, all backward passes, clip + Adam; actor_critic_update = "
                                 "Agent.train_actor_critic on the 2450 resulting start rows incl. both Adam steps; *_graphed_ms = the "
                                 "same call captured once as a CUDA graph and replayed (Agent.graphed)"}

    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    total_steps = N * T * world
    value_sps = total_steps * a.steps / (ms / 1e3)
    e2e_sps = total_steps * a.steps / (e2e_ms / 1e3)
    peak_burst, peak_sust, peak_gbs, peak_src = peaks()
    achieved_tf = N * T * FLOP_PER_STEP / (kernel_ms / 1e3) / 1e12
    # burst peak when the sampled SM clock sat at its maximum during the timed loops (a short region, not power-capped
    # down), sustained otherwise
    at_max = bool(clocks and clocks.get("sm_mhz") and clocks.get("sm_max_mhz") and clocks["sm_mhz"] >= 0.97 * clocks["sm_max_mhz"])
    peak_tf = peak_burst if at_max else peak_sust
    line = {
        "metric": "imagined latent steps/sec", "value": value_sps, "unit": "steps/s", "n_gpus": world,
        "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f16x3 split (fp32-grade products, fp32 accumulate) on tcgen05",
        "data": "synthetic", "config": workload_config(a),
        "clocks": clocks,
        "e2e": {"value": e2e_sps, "unit": "steps/s", "h2d_bytes_per_step": N * (D + S) * 4, "d2h_bytes_per_step": (T - 1) * N * 4,
                "ms_per_step": e2e_ms / a.steps,
                "what": "per step: pinned-host start states -> H2D (double-buffered on a copy stream), TransitionModel.imagine (device "
                        "noise draw — with the module's prefetch_noise option the draw for step i+1 is enqueued on a side stream under "
                        "step i's kernel, still inside the timed region — fused kernel), D2H of the lambda-returns.  The imagined trajectories (1.2 GB per step) stay on "
                        "the device, as in the reference, where imagine() feeds the actor / value losses and only scalars leave the GPU "
                        "(dreamer.py:304-381)"},
        # per rank per timed step: pack_rows_weights_kernel + pack_rows_bias_kernel + rssm_rows_kernel (the weights are
        # re-packed every step because an optimiser step would have changed them; ~0.3 % of the step)
        "gpu_launches": 3 * a.steps,
        "roofline": {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf,
                     "peak_kind": "burst" if at_max else "sustained", "frac_of_burst": achieved_tf / peak_burst,
                     "frac_of_sustained": achieved_tf / peak_sust,
                     "traffic": NCU_CAPTURE["traffic_bytes"] if N == NCU_CAPTURE["rows"] else None, "traffic_source": NCU_CAPTURE["file"],
                     "kernel": "rssm_rows_kernel", "kernel_ms": kernel_ms,
                     "kernel_ms_how": "CUDA events around a.steps back-to-back launches with pre-packed weights, on the launching stream",
                     "algorithmic_flop_per_step": FLOP_PER_STEP, "algorithmic_bytes_per_step": BYTES_PER_STEP,
                     "mma_issue_factor": MMA_ISSUE_FACTOR, "tensor_pipe_active_pct_ncu": NCU_CAPTURE["tensor_pipe_active_pct"],
                     "l2_to_sm_bytes_ncu": NCU_CAPTURE["l2_to_sm_bytes"],
                     "tensor_pipe_busy_estimate": achieved_tf * MMA_ISSUE_FACTOR / 2250.0,
                     "hbm_gbs_achieved": N * T * BYTES_PER_STEP / (kernel_ms / 1e3) / 1e9, "hbm_peak_gbs": peak_gbs,
                     "peak_source": peak_src,
                     "note": "algorithmic fp32 FLOPs (SURVEY 8d: 1.44 MFLOP per imagined step); the kernel issues 3 fp16 MMAs per product "
                             "(hi*hi + lo*hi + hi*lo) plus K/N padding = mma_issue_factor x the algorithmic MACs on the tensor pipe; "
                             "tensor_pipe_busy_estimate = achieved x factor / 2250 nominal dense fp16 TFLOP/s"},
        "parity_sample": parity,
        "sweep": sweep,
        "dp_update": dp_update,
        "default_shape": default_shape,
    }
    if not a.no_cpu_baseline and world == 1:
        v, sec, cores, kind = cpu_baseline(2450, 3, 1)
        line["cpu_baseline"] = {"value": v, "unit": "steps/s", "cores": cores, "kind": kind,
                                "sample": "RePo default shape: 2450 start rows x 14 steps, 3 timed passes (torch CPU fp32, all host threads)"}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


DP_GRAPHS = os.environ.get("REPO_B200_DP_GRAPHS", "1") != "0"


def dp_update_block(dev, rank, world, dist):
    """One RePo training iteration (Agent.train_dynamics + train_actor_critic) with the replay batch sharded over ALL ranks:
    strong scaling (global batch 50, shards 7,7,6,6,... with losses weighted B_local / B) and weak scaling (50 sequences per
    GPU).  The one exchange is FlatAdam's flat-bucket gradient all-reduce (NCCL, one collective per parameter group); it is
    also timed alone on buffers of the same sizes."""
    from repo_b200 import parallel, synth
    from repo_b200.trainer import Agent, Config
    T, A = 50, DIMS["action"]
    out = {}

    def timed(fn, n):
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / n, dev)

    for mode, B in (("strong", 50), ("weak", 50 * world)):
        agent = Agent(Config(batch_size=B, chunk_size=T), A, algo="repo", device=dev)
        agent.optimizers()
        c0, cn = parallel.shard_rows(B, rank, world)
        full = synth.make_train_batch(7, T, 50, A)                 # every rank draws the same 50 sequences ...
        cols = [(c0 + j) % 50 for j in range(cn)]                  # ... and takes its columns (weak: the batch tiled x world)
        batch = {k: v[:, cols].contiguous().to(dev) for k, v in full.items()}
        st = {}

        def wm():
            st["b"], st["s"] = agent.train_dynamics(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])

        def ac():
            agent.train_actor_critic(st["b"].flatten(0, 1), st["s"].flatten(0, 1))

        for _ in range(3):
            wm(); ac()
        out[mode] = {"global_batch": B, "shards": [parallel.shard_rows(B, r, world)[1] for r in range(world)],
                     "iteration_ms": timed(lambda: (wm(), ac()), 5), "world_model_update_ms": timed(wm, 5),
                     "actor_critic_update_ms": timed(ac, 5)}
        buckets = {k: int(o.numel) * 4 for k, o in agent.optimizers().items() if hasattr(o, "numel")}
        # the same iteration as two CUDA-graph replays (Agent.graphed): the gradient all-reduces are captured inside the
        # graphs (NCCL collectives are capturable; every rank captures in lock-step).  With 6-7 sequences per GPU the eager
        # iteration is bound by the host issuing ~750 launches, not by the GPU.
        if DP_GRAPHS:
            try:
                g_wm, g_ac = agent.graphed(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])
                gs = {}

                def it_g():
                    gs["b"], gs["s"] = g_wm(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])
                    g_ac(gs["b"].flatten(0, 1), gs["s"].flatten(0, 1))

                for _ in range(3):
                    it_g()
                out[mode]["iteration_graphed_ms"] = timed(it_g, 10)
                del g_wm, g_ac
            except Exception as exc:  # noqa: BLE001 - reported in the line, the eager numbers above stand
                out[mode]["iteration_graphed_ms"] = None
                out[mode]["graph_capture_error"] = str(exc)[:200]
        del agent
    # the collective alone: SUM all-reduce of fp32 buffers of the bucket sizes (world model, actor, value)
    sizes = buckets if buckets else {"model": 5_170_420 * 4, "actor": 169_212 * 4, "value": 126_801 * 4}
    bufs = [torch.zeros(max(1, b // 4), device=dev) for b in sizes.values()]

    def ar():
        for t_ in bufs:
            dist.all_reduce(t_)

    for _ in range(3):
        ar()
    out["allreduce_bytes"] = int(sum(sizes.values()))
    out["allreduce_buckets"] = sizes
    out["allreduce_ms"] = timed(ar, 10)
    out["note"] = ("RePo defaults (batch x chunk 50 of 64x64x3 frames, horizon 15); iteration = train_dynamics + train_actor_critic "
                   "incl. all optimiser steps; max over ranks; allreduce_ms = the same buckets all-reduced alone")
    return out


def run_update(a):
    """One training iteration at the reference's default sizes, data parallel over the replay batch: every rank takes
    its columns of the (50, B, 3, 64, 64) batch (uneven shards allowed: 50 over 8 = 7,7,6,6,6,6,6,6), weights its losses
    by B_local / B, and the flat gradient buckets are all-reduced inside FlatAdam.step (one NCCL collective per
    parameter group).  Strong scaling: the global batch is fixed."""
    import torch
    import torch.distributed as dist
    from repo_b200 import parallel
    from repo_b200 import synth as O
    from repo_b200.trainer import Agent, Config
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    T, B, A = 50, a.global_batch, DIMS["action"]
    agent = Agent(Config(batch_size=B, chunk_size=T), A, algo=a.algo, device=dev)
    agent.optimizers()
    c0, cn = parallel.shard_rows(B, rank, world)
    full = O.make_train_batch(7, T, B, A)
    batch = {k: v[:, c0:c0 + cn].contiguous().to(dev) for k, v in full.items()}
    st = {}

    def wm():
        st["b"], st["s"] = agent.train_dynamics(batch["obs"], batch["actions"], batch["rewards"], batch["nonterms"])

    def ac():
        agent.train_actor_critic(st["b"].flatten(0, 1), st["s"].flatten(0, 1))

    def timed(fn, n):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return parallel.max_over_ranks(e0.elapsed_time(e1) / n, dev)

    for _ in range(max(a.warmup, 3)):
        wm(); ac()
    sampler = ClockSampler(local)
    sampler.start()
    it_ms = timed(lambda: (wm(), ac()), a.steps)
    clocks = sampler.stop()
    wm_ms = timed(wm, max(3, a.steps // 2))
    ac_ms = timed(ac, max(3, a.steps // 2))
    if rank == 0:
        rows = (T - 1) * B
        print(json.dumps({
            "metric": "training iteration ms (world-model update + actor-critic update)", "value": it_ms, "unit": "ms",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": it_ms, "higher_is_better": False,
            "scaling": "strong", "vs_baseline": None, "dtype": "f16x3 split (fp32-grade products, fp32 accumulate) on tcgen05",
            "data": "synthetic",
            "config": {"workload": f"{a.algo} training iteration at the reference defaults: batch {B} x chunk {T} of 64x64x3 frames, horizon {HORIZON}",
                       "algo": a.algo, "global_batch": B, "chunk": T, "horizon": HORIZON,
                       "parallelism": f"batch columns sharded x{world} (shards {[parallel.shard_rows(B, r, world)[1] for r in range(world)]}), "
                                      "flat-bucket gradient all-reduce over NCCL"},
            "world_model_update_ms": wm_ms, "actor_critic_update_ms": ac_ms,
            "imagined_steps_per_s": rows * (HORIZON - 1) / ac_ms * 1e3, "clocks": clocks}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.workload == "update":
        run_update(a)
    else:
        run_gpu(a)


if __name__ == "__main__":
    main()
